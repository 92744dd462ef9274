"""Host-buffer path (cc_step_host / cc_rollout_host) throughput by delivery format and chunking.

    python profiles/e2e_probe.py [--envs N] [--steps K]

Each line: format, chunk, ms per step (wall clock around K synchronous calls), agent-steps/s, D2H bytes per step and the
PCIe rate they imply.  The loop is the bench's e2e loop: policy actions -> host (pinned) -> cc_step_host."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import torch  # noqa: E402
from cases import readme_config  # noqa: E402

from collectivecrossing_b200 import BatchedCollectiveCrossing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
n, A = args.envs, 8
cfg = readme_config()
dev = torch.device("cuda:0")


def run(tag, obs, chunk=0, expand=0, policy_on_device=False, T=1, pinned=True):
    env = BatchedCollectiveCrossing(cfg, n, dev, seed=1, obs_dtype=obs if obs else "none", auto_reset=True)
    env.set_host_chunk(chunk)
    env.set_host_expand(expand)
    env.reset()
    host = env.make_host_buffers(pinned=pinned, n_steps=None if T == 1 else T)
    dev_actions = torch.zeros((n, A), dtype=torch.int8, device=dev)

    def step():
        if T > 1:
            env.rollout_host(host, T, policy="greedy")
        elif policy_on_device:
            env.step_host(host, policy="greedy")
        else:
            env.policy_actions("greedy", out=dev_actions)
            host["actions"].copy_(dev_actions, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            env.step_host(host)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / (args.steps * T)
    call = env.last_host_call()
    d2h = call["d2h_bytes"] // T + (0 if policy_on_device or T > 1 else n * A)
    print(json.dumps({"tag": tag, "obs": obs, "pinned": pinned, "chunks": call["chunks"], "expand_threads": call["expand_threads"], "steps_per_call": T,
                      "ms_per_step": round(dt * 1e3, 3), "agent_steps_per_sec": round(n * A / dt), "d2h_bytes_per_step": d2h,
                      "pcie_GBps": round(d2h / dt / 1e9, 1)}), flush=True)
    env.close()
    del env, host
    torch.cuda.empty_cache()


AUTO = -2
run("fp32 rows, the handle's own delivery (table over PCIe, rows rebuilt by the host threads)", "float32", expand=AUTO)
run("fp32 rows over PCIe", "float32")
run("fp32 rows over PCIe, one chunk", "float32", chunk=n)
run("fp32 rows over PCIe, 32 chunks", "float32", chunk=n // 32)
run("fp32 rows rebuilt on the host, 8 threads", "float32", expand=8)
run("fp32 rows rebuilt on the host, 4 threads", "float32", expand=4)
run("fp32 rows rebuilt on the host, 4 chunks", "float32", expand=AUTO, chunk=n // 4)
run("fp32 rows rebuilt on the host, policy on device", "float32", expand=AUTO, policy_on_device=True)
run("int8 rows, the handle's own delivery", "int8", expand=AUTO)
run("int8 rows over PCIe", "int8")
run("table", "table", expand=AUTO)
run("table, one chunk", "table", chunk=n, expand=AUTO)
run("table, 4 chunks", "table", chunk=n // 4, expand=AUTO)
run("table, policy on device", "table", policy_on_device=True, expand=AUTO)
run("no observations", None, expand=AUTO)
run("rollout_host T=8 fp32", "float32", T=8, expand=AUTO)
run("rollout_host T=8 table", "table", T=8, expand=AUTO)
# ordinary (pageable) caller memory: through the handle's pinned mirrors, and with the host threads switched off
run("fp32 rows, pageable buffers", "float32", expand=AUTO, pinned=False, policy_on_device=True)
run("table, pageable buffers", "table", expand=AUTO, pinned=False, policy_on_device=True)
run("table, pageable buffers, no host threads (direct copies)", "table", expand=0, pinned=False, policy_on_device=True)
