import sys, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from cases import large_config, crew_config
from collectivecrossing_b200 import BatchedCollectiveCrossing
def t(cfg, n, obs, pol, steps=5):
    env = BatchedCollectiveCrossing(cfg, n, "cuda:0", seed=1, obs_dtype=obs, auto_reset=True)
    env.reset()
    for _ in range(3): env.step(policy=pol)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): env.step(policy=pol)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/steps
    b=env.algorithmic_bytes_per_env_step()*n
    print(f"A={env.num_agents} n={n} obs={obs} pol={pol} kernel={env.last_kernel}: {ms:.3f} ms  {b/ms/1e6:.0f} GB/s  {n*env.num_agents/ms/1e6:.2f} G agent-steps/s", flush=True)
    env.close()
for obs in ("none","int8","float32"):
    t(large_config(512), 1<<18, obs, "random")
t(large_config(512), 1<<18, "float32", "waiting")
t(crew_config(10,6), 1<<19, "float32", "greedy")
t(crew_config(20,12), 1<<18, "float32", "greedy")
t(crew_config(3,2), 1<<20, "float32", "greedy")
