"""Compact output modes of the thread-per-env kernels at 1,048,576 README envs: ms per env-step (CUDA events) and the
fraction of the measured HBM peak that the algorithmic bytes reach.

    python profiles/compact_modes_probe.py            # the small-lattice kernel (cc_step_tpe2_kernel)
    CCB200_TPE2=0 python profiles/compact_modes_probe.py   # cc_step_tpe_kernel for comparison
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import torch  # noqa: E402
from cases import readme_config  # noqa: E402

from collectivecrossing_b200 import BatchedCollectiveCrossing  # noqa: E402

n = int(os.environ.get("ENVS", 1 << 20))
peak = 6525.2
pp = ROOT / "MEASURED_PEAKS.json"
if pp.exists():
    peak = float(json.loads(pp.read_text())["hbm_gbs"])
cfg = readme_config()
dev = torch.device("cuda:0")
for obs in ("none", "table", "int8"):
    for policy in ("greedy", "waiting", "random"):
        for T in (1, 20):
            env = BatchedCollectiveCrossing(cfg, n, dev, seed=1, obs_dtype=obs, auto_reset=True)
            env.reset()
            for _ in range(30):
                env.step(policy=policy)
            if T > 1:
                env.rollout_trajectory(T, policy=policy)
            reps = 100 if T == 1 else 6
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                if T == 1:
                    env.step(policy=policy)
                else:
                    env.rollout_trajectory(T, policy=policy)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (reps * T)
            b1 = env.algorithmic_bytes_per_env_step()
            bT = b1 - (2 * (3 * 8 + 4) + 8) * (1.0 - 1.0 / T)
            gbs = bT * n / (ms * 1e-3) / 1e9
            print(json.dumps({"obs": obs, "policy": policy, "steps_per_launch": T, "kernel": env.last_kernel_name, "ms_per_step": round(ms, 5),
                              "G_agent_steps_per_s": round(n * 8 / ms / 1e6, 1), "bytes_per_env_step": round(bT, 1), "GBps": round(gbs), "frac": round(gbs / peak, 3)}), flush=True)
            env.check_error()
            env.close()
            del env
            torch.cuda.empty_cache()
