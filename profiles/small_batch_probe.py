"""Small batches (BASELINE config 2 literal: 65,536 envs): per-step time of a Python loop of step() calls, of the C-side loop
cc_rollout (same launches, no Python between them) and of fused launches; float32 rows and the compact modes."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import torch  # noqa: E402
from cases import readme_config  # noqa: E402

from collectivecrossing_b200 import BatchedCollectiveCrossing  # noqa: E402

peak = 6525.2
cfg = readme_config()
for n in (65536, 262144):
    for obs in ("float32", "int8", "none"):
        env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=1, obs_dtype=obs, auto_reset=True)
        env.reset()
        for _ in range(30):
            env.step(policy="greedy")
        env.rollout_trajectory(20, policy="greedy")
        res = {}
        for mode in ("python_loop", "c_loop", "fused20"):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode == "python_loop":
                for _ in range(400):
                    env.step(policy="greedy")
            elif mode == "c_loop":
                env.rollout(400, policy="greedy")
            else:
                for _ in range(20):
                    env.rollout_trajectory(20, policy="greedy")
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / 400
        b = env.algorithmic_bytes_per_env_step()
        print(json.dumps({"envs": n, "obs": obs, "kernel": env.last_kernel_name, **{k: round(v * 1e3, 2) for k, v in res.items()},
                          "unit": "us per step", "frac_c_loop": round(b * n / (res["c_loop"] * 1e-3) / 1e9 / peak, 3),
                          "ideal_us": round(b * n / peak / 1e3, 2)}), flush=True)
        env.close()
