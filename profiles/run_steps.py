"""Tiny driver for ncu captures: python profiles/run_steps.py <obs> <policy> <steps> [envs] [steps_per_launch] [config]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402
from cases import large_config, readme_config  # noqa: E402

from collectivecrossing_b200 import BatchedCollectiveCrossing  # noqa: E402

obs, policy, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
T = int(sys.argv[5]) if len(sys.argv) > 5 else 1
cfg = large_config(512) if (len(sys.argv) > 6 and sys.argv[6] == "large") else readme_config()
env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=1, obs_dtype=obs, auto_reset=True)
env.reset()
for _ in range(steps):
    if T > 1:
        env.rollout_trajectory(T, policy=policy)
    else:
        env.step(policy=policy)
torch.cuda.synchronize()
env.check_error()
print(env.last_kernel_name, env.stats()["episodes"])
