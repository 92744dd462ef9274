"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden vectors.
Bit-exact for positions, flags, integer observations, infos and — because the kernel computes
rewards in float64 and rounds once — also for float32/float64 rewards."""

import numpy as np
import pytest
from cases import GOLDEN_CASES, cassette_config, crew_config, large_config, readme_config, readme_crew
from helpers import CASSETTE_BITS, assert_same, load_golden, random_states, replay_device

from collectivecrossing_b200 import _abi
from collectivecrossing_b200.lowering import lower_config

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

POLICY_CASES = [n for n, (_, kw) in GOLDEN_CASES.items() if kw["source"] in ("greedy", "waiting")]


def make_env(cfg, n, **kw):
    from collectivecrossing_b200 import BatchedCollectiveCrossing

    return BatchedCollectiveCrossing(cfg, n, "cuda:0", **kw)


# ---- golden vectors (reference outputs) ----------------------------------------------------------
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_device_replays_golden(name):
    cfg = GOLDEN_CASES[name][0]()
    rec = load_golden(name)
    assert_same(rec, replay_device(cfg, rec), f"device/{name}")


@pytest.mark.parametrize("name", POLICY_CASES)
def test_device_policies_match_reference_policies(name):
    cfg, kw = GOLDEN_CASES[name][0](), GOLDEN_CASES[name][1]
    rec = load_golden(name)
    assert_same(rec, replay_device(cfg, rec, policy=kw["source"]), f"device-policy/{name}", policy_actions=True)


@pytest.mark.parametrize("name", ["cassette_basic", "cassette_regression"])
def test_device_replays_reference_cassettes(name):
    rec = load_golden(name)
    assert_same(rec, replay_device(cassette_config(), rec), f"device/{name}", flag_mask=CASSETTE_BITS)


def test_device_fp32_outputs_match_golden_within_tolerance():
    """float32 observations are the same integers; float32 rewards within 1e-6 relative of the
    reference's float64 (north_star tolerance)."""
    name = "readme_random"
    cfg, rec = GOLDEN_CASES[name][0](), load_golden(name)
    got = replay_device(cfg, rec, obs_dtype="float32", reward_dtype="float32")
    assert_same(rec, got, "device-fp32", reward_rtol=1e-6)


# ---- the thread-per-env mapping against the REFERENCE's recordings -------------------------------
@pytest.mark.parametrize("name", ["readme_greedy", "readme_waiting", "readme_binary_all", "readme_simple_all", "readme_constneg_ind"])
def test_thread_per_env_kernel_replays_golden(name):
    """float32 rewards + agent order make a README step eligible for the thread-per-env kernel:
    positions, flags, observations, infos and policy actions bit-exact with the reference's
    recording, rewards within the north_star tolerance (1e-6 relative; they are in fact the
    correctly rounded float32 of the reference's float64)."""
    cfg, kw = GOLDEN_CASES[name][0](), GOLDEN_CASES[name][1]
    rec = load_golden(name)
    on_device = kw["source"] in ("greedy", "waiting")
    for obs_dtype in ("float32", "int8"):
        got = replay_device(cfg, rec, policy=kw["source"] if on_device else "external", use_order=False,
                            obs_dtype=obs_dtype, reward_dtype="float32", kernel="threads")
        assert got["_kernel"] == "threads"
        assert_same(rec, got, f"threads/{name}/{obs_dtype}", policy_actions=on_device, reward_rtol=1e-6)
        assert np.array_equal(got["reward"], rec["reward"].astype(np.float32).astype(np.float64)), "rewards are the rounded float64"


# ---- seeded comparison against the oracle, every kernel instantiation ----------------------------
ORACLE_CASES = {
    "A1": (lambda: crew_config(1, 0, max_steps=25), 37),
    "A2": (lambda: crew_config(1, 1, max_steps=30, reward="binary"), 75),
    "A3_cassette": (cassette_config, 101),
    "A4": (lambda: crew_config(3, 1, max_steps=40, reward="simple_distance", term="all"), 1000),
    "A5": (lambda: crew_config(3, 2, max_steps=40, term="all"), 258),
    "A6": (lambda: crew_config(4, 2, max_steps=40), 131),
    "A7": (lambda: crew_config(4, 3, max_steps=40, reward="constant_negative", term="all"), 260),
    "A4_small_lattice": (lambda: readme_crew(2, 2, term="all"), 333),
    "A6_small_lattice": (lambda: readme_crew(4, 2, reward="simple_distance"), 300),
    "A7_small_lattice": (lambda: readme_crew(4, 3), 129),
    "A8_readme": (lambda: readme_config(max_steps=60), 1031),
    "A12": (lambda: crew_config(7, 5, max_steps=40, reward="simple_distance"), 130),
    "A21_odd": (lambda: crew_config(13, 8, max_steps=40), 67),
    "A40": (lambda: crew_config(25, 15, max_steps=40, reward="binary"), 33),
    "A64_large": (lambda: large_config(40), 19),
    "A100": (lambda: crew_config(60, 40, max_steps=30, reward="constant_negative", term="all"), 9),
}


def _compare_step(env, orc, res, out, tag, obs_np):
    for k, a, b in (("x", env.x, orc.x), ("y", env.y, orc.y), ("flags", env.flags, orc.flags),
                    ("step", env.step_count, orc.step_count), ("episode_return", env.episode_return, orc.episode_return)):
        assert np.array_equal(a.cpu().numpy(), b), f"{tag}: state {k} differs"
    assert np.array_equal(out.reward.cpu().numpy(), res["reward"]), f"{tag}: reward"
    assert np.array_equal(out.agent_flags.cpu().numpy(), res["agent_flags"]), f"{tag}: agent_flags"
    assert np.array_equal(out.agent_info.cpu().numpy(), res["agent_info"]), f"{tag}: agent_info"
    assert np.array_equal(out.env_flags.cpu().numpy(), res["env_flags"]), f"{tag}: env_flags"
    assert np.array_equal(out.actions.cpu().numpy(), res["actions_out"]), f"{tag}: actions"
    assert np.array_equal(out.obs.cpu().numpy().astype(obs_np), res["obs"]), f"{tag}: obs"


# every case with the default mapping, and the crews the thread-per-env kernel takes also with the lane-group one
SMALL_CREWS = ("A1", "A2", "A3_cassette", "A4", "A5", "A6", "A7", "A4_small_lattice", "A6_small_lattice", "A7_small_lattice", "A8_readme")   # served by the thread-per-env kernel
CASE_KERNELS = [(c, "auto") for c in ORACLE_CASES] + [(c, "lanes") for c in SMALL_CREWS]


@pytest.mark.parametrize("policy", ["random", "greedy", "waiting"])
@pytest.mark.parametrize("case,kernel", CASE_KERNELS)
def test_device_matches_oracle_with_auto_reset(case, kernel, policy):
    """Same seeded start states, on-device policy, auto-reset on: every step, every field."""
    import oracle

    make_cfg, n = ORACLE_CASES[case]
    cfg = make_cfg()
    low = lower_config(cfg)
    rng = np.random.default_rng(hash((case, policy)) % 2**32)
    x, y, f, s = random_states(cfg, n, rng)
    seed, offset = 1234567 + n, 10_000_000_000 + n  # offset > 2**32: both counter words matter
    obs_dtype = "float32" if policy == "greedy" else "int8"
    if case in SMALL_CREWS and case != "A8_readme" and kernel == "auto":
        obs_dtype = "float32"   # int8 rows go through the thread-per-env kernel only for 8 agents (whole 16-byte vectors)
    env = make_env(cfg, n, seed=seed, global_env_offset=offset, obs_dtype=obs_dtype, auto_reset=True, with_info=True, kernel=kernel)
    orc = oracle.OracleEnvs(low, n, seed=seed, global_env_offset=offset)
    env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
    orc.set_state(x, y, f, s)
    steps = 70 if low.num_agents <= 21 else 45
    for t in range(steps):
        out = env.step(policy=policy)
        res = orc.step(policy=policy, auto_reset=True, obs_dtype=_abi.OBS_FP32 if obs_dtype == "float32" else _abi.OBS_INT8)
        _compare_step(env, orc, res, out, f"{case}/{policy}/t={t}", np.float32 if obs_dtype == "float32" else np.int8)
    env.check_error()
    assert env.last_kernel == ("threads" if kernel == "auto" and case in SMALL_CREWS else "lanes")
    st, want = env.stats(), orc.stats.as_dict()
    for k in ("env_steps", "episodes", "terminated_all", "truncated_all", "arrivals", "episode_length_sum"):
        assert st[k] == want[k], (k, st[k], want[k])
    for k in ("episode_return_sum", "reward_sum"):
        assert st[k] == pytest.approx(want[k], rel=1e-9, abs=1e-9), k
    assert st["episodes"] > 0 or low.max_steps > steps


@pytest.mark.parametrize("case", ["A5", "A8_readme", "A21_odd", "A64_large"])
def test_device_external_actions_and_custom_order_match_oracle(case):
    """Random action tensors incl. partial lists in shuffled dict order, no auto-reset."""
    import oracle

    make_cfg, n = ORACLE_CASES[case]
    cfg = make_cfg()
    low = lower_config(cfg)
    A = low.num_agents
    rng = np.random.default_rng(99)
    x, y, f, s = random_states(cfg, n, rng)
    env = make_env(cfg, n, obs_dtype="int8", reward_dtype="float64", auto_reset=False, with_info=True)
    orc = oracle.OracleEnvs(low, n)
    env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
    orc.set_state(x, y, f, s)
    for t in range(50):
        acts = rng.integers(0, 5, size=(n, A)).astype(np.int8)
        order = np.full((n, A), -1, np.int8)
        for e in range(n):
            k = int(rng.integers(0, A + 1))
            order[e, :k] = rng.permutation(A)[:k]
        use_order = t % 3 != 0
        out = env.step(torch.from_numpy(acts).cuda(), order=torch.from_numpy(order).cuda() if use_order else None)
        res = orc.step(acts, order=order if use_order else None, obs_dtype=_abi.OBS_INT8, reward_dtype=_abi.REWARD_F64)
        _compare_step(env, orc, res, out, f"{case}/t={t}", np.int8)
    env.check_error()


@pytest.mark.parametrize("obs_dtype", ["none", "int8", "float32"])
@pytest.mark.parametrize("case", list(SMALL_CREWS))
def test_thread_per_env_kernel_external_actions_match_oracle(case, obs_dtype):
    """External action tensors (incl. out-of-range values that must raise), no auto-reset, ragged
    env counts: the thread-per-env kernel against the oracle and against the lane-group kernel."""
    import oracle

    make_cfg, _ = ORACLE_CASES[case]
    cfg = make_cfg()
    low = lower_config(cfg)
    A = low.num_agents
    if obs_dtype == "int8" and A != 8:
        pytest.skip("int8 rows of other crews are not 16-byte multiples: served by the lane-group kernel")
    code = {"none": _abi.OBS_NONE, "int8": _abi.OBS_INT8, "float32": _abi.OBS_FP32}[obs_dtype]
    for n in (1, 31, 33, 517):
        rng = np.random.default_rng(n)
        x, y, f, s = random_states(cfg, n, rng, step_hi=low.max_steps - 5)
        envs = {k: make_env(cfg, n, obs_dtype=obs_dtype, auto_reset=False, with_info=True, kernel=k) for k in ("threads", "lanes")}
        orc = oracle.OracleEnvs(low, n)
        for e in envs.values():
            e.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
        orc.set_state(x, y, f, s)
        for t in range(30):
            acts = rng.integers(0, 5, size=(n, A)).astype(np.int8)
            res = orc.step(acts, obs_dtype=code)
            for k, e in envs.items():
                out = e.step(torch.from_numpy(acts).cuda())
                assert e.last_kernel == k
                for name, a, b in (("x", e.x, orc.x), ("y", e.y, orc.y), ("flags", e.flags, orc.flags), ("step", e.step_count, orc.step_count),
                                   ("episode_return", e.episode_return, orc.episode_return), ("reward", out.reward, res["reward"]),
                                   ("agent_flags", out.agent_flags, res["agent_flags"]), ("agent_info", out.agent_info, res["agent_info"]),
                                   ("env_flags", out.env_flags, res["env_flags"])):
                    assert np.array_equal(a.cpu().numpy(), b), f"{case}/{k}/n={n}/t={t}: {name}"
                if obs_dtype != "none":
                    assert np.array_equal(out.obs.cpu().numpy(), res["obs"]), f"{case}/{k}/n={n}/t={t}: obs"
        for e in envs.values():
            e.check_error()
        bad = rng.integers(0, 5, size=(n, A)).astype(np.int8)
        bad[n - 1, A - 1] = 7
        envs["threads"].step(torch.from_numpy(bad).cuda())
        with pytest.raises(ValueError, match="Invalid action"):
            envs["threads"].check_error()
        for e in envs.values():
            e.close()


def test_thread_per_env_request_on_ineligible_step_raises():
    env = make_env(crew_config(7, 5), 8, obs_dtype="none", kernel="threads")
    env.reset()
    with pytest.raises(NotImplementedError, match="CC_KERNEL_THREADS"):
        env.step(policy="greedy")
    env.close()


# ---- BASELINE config 2 at full size: 65,536 envs, greedy policy, bit-exact replay ----------------
def test_config2_65536_envs_greedy_bit_exact_vs_oracle():
    import oracle

    cfg = readme_config()
    low = lower_config(cfg)
    n = 65536
    rng = np.random.default_rng(2)
    x, y, f, s = random_states(cfg, n, rng)
    env = make_env(cfg, n, seed=7, obs_dtype="float32", auto_reset=True, with_info=True)
    orc = oracle.OracleEnvs(low, n, seed=7)
    env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
    orc.set_state(x, y, f, s)
    for t in range(120):  # beyond max_steps = 100: truncation-driven auto-reset is exercised
        out = env.step(policy="greedy")
        res = orc.step(policy="greedy", auto_reset=True, obs_dtype=_abi.OBS_FP32)
        if t % 10 == 9 or t >= 98:
            _compare_step(env, orc, res, out, f"cfg2/t={t}", np.float32)
    _compare_step(env, orc, res, out, "cfg2/final", np.float32)
    st = env.stats()
    assert st["episodes"] == orc.stats.episodes > 0 and st["truncated_all"] == orc.stats.truncated_all


# ---- size-independent properties at BASELINE's large sizes ----------------------------------------
def _check_invariants(env, cfg, out):
    """Properties every reachable state has, evaluated with torch ops on the device."""
    low = lower_config(cfg)
    A, B = low.num_agents, low.num_boarding
    x, y, fl = env.x.int(), env.y.int(), env.flags
    active = (fl & _abi.F_ACTIVE) != 0
    # inside the lattice, never on a wall (collectivecrossing.py:509-534)
    assert bool(((x >= 0) & (x <= low.width) & (y >= 0) & (y <= low.height)).all())
    on_div = y == low.division_y
    assert bool((~on_div | ((x > low.door_left) & (x < low.door_right))).all())
    assert bool(((y < low.division_y) | ((x > low.tram_left) & (x < low.tram_right))).all())
    # no two ACTIVE agents share a cell
    cell = torch.where(active, y * 128 + x, -1 - torch.arange(A, device=x.device).expand_as(x))
    srt = cell.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # arrived <=> inactive; step counter within bounds
    dest = torch.where(torch.arange(A, device=x.device) < B, low.boarding_dest_y, low.exiting_dest_y)
    assert bool(((y == dest) == ~active).all())
    assert bool(((env.step_count >= 0) & (env.step_count <= low.max_steps)).all())
    # observation rows are the state table with the own block masked (observations.py:62-94)
    obs = out.obs
    if obs is not None:
        o = obs.float()
        assert bool((o[:, :, 0] == x.float()).all() and (o[:, :, 1] == y.float()).all())
        door = torch.tensor([(low.door_left + low.door_right) // 2, low.division_y, low.door_left, low.door_right], device=o.device).float()
        assert bool((o[:, :, 2:6] == door).all())
        tab = torch.stack([x.float(), y.float(), (torch.arange(A, device=o.device) >= B).float().expand_as(x), active.float()], dim=2)
        want = tab[:, None, :, :].expand(-1, A, -1, -1).clone()
        idx = torch.arange(A, device=o.device)
        want[:, idx, idx, :] = -1.0
        assert bool((o[:, :, 6:].reshape(want.shape) == want).all())


def test_one_million_envs_invariants_and_shard_invariance():
    """BASELINE config 5 shape (12x8, waiting policy, auto-reset) at 1M envs: invariants hold, the
    episode statistics add up, and splitting the envs over two handles with global offsets gives
    the same trajectories (the counter RNG is keyed on the global env index)."""
    cfg = readme_config()
    n = 1 << 20
    whole = make_env(cfg, n, seed=5, obs_dtype="int8", auto_reset=True)
    whole.reset()
    halves = [make_env(cfg, n // 2, seed=5, global_env_offset=k * (n // 2), obs_dtype="int8", auto_reset=True) for k in range(2)]
    for h in halves:
        h.reset()
    assert torch.equal(whole.x, torch.cat([h.x for h in halves])) and torch.equal(whole.y, torch.cat([h.y for h in halves]))
    resets = 0
    for t in range(64):
        out = whole.step(policy="waiting")
        resets += int(out.was_reset.sum())
        for h in halves:
            h.step(policy="waiting")
        if t % 16 == 15:
            _check_invariants(whole, cfg, out)
            assert torch.equal(whole.x, torch.cat([h.x for h in halves]))
            assert torch.equal(whole.flags, torch.cat([h.flags for h in halves]))
            assert torch.equal(whole.obs, torch.cat([h.obs for h in halves]))
    st = whole.stats()
    assert st["env_steps"] == 64 * n and st["episodes"] == resets > 0
    assert st["terminated_all"] + st["truncated_all"] >= st["episodes"]
    parts = [h.stats() for h in halves]
    for k in ("episodes", "arrivals", "episode_length_sum", "terminated_all", "truncated_all"):
        assert st[k] == parts[0][k] + parts[1][k]
    whole.check_error()


def test_large_geometry_invariants_fp32():
    """BASELINE config 3 shape (64x32, 64 agents, SimpleDistance, AllAtDestination), random actions."""
    cfg = large_config(48)
    env = make_env(cfg, 4099, seed=3, obs_dtype="float32", auto_reset=True)
    env.reset()
    for t in range(60):
        out = env.step(policy="random")
    _check_invariants(env, cfg, out)
    st = env.stats()
    assert st["episodes"] == st["truncated_all"] > 0  # nobody gets 64 agents home in 48 random steps
    env.check_error()


# ---- API behaviour -------------------------------------------------------------------------------
def test_reset_kernel_matches_oracle_and_mask():
    import oracle

    cfg = readme_config()
    low = lower_config(cfg)
    n = 777
    env = make_env(cfg, n, seed=11, global_env_offset=5, obs_dtype="int8")
    orc = oracle.OracleEnvs(low, n, seed=11, global_env_offset=5)
    obs = env.reset()
    want = orc.reset()
    assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.y.cpu().numpy(), orc.y)
    assert np.array_equal(obs.cpu().numpy(), want)
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    before = env.x.clone()
    env.reset(torch.from_numpy(mask).cuda())
    orc.reset(mask)
    assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.flags.cpu().numpy(), orc.flags)
    assert torch.equal(env.x[1::3], before[1::3])


@pytest.mark.parametrize("name", ["readme_random", "large_random", "crew_60_40"])
def test_reset_seeded_is_bit_exact_with_reference_reset(name):
    """reset(seed) on the device == the reference's reset(seed) (golden initial states)."""
    cfg, rec = GOLDEN_CASES[name][0](), load_golden(name)
    env = make_env(cfg, len(rec["seeds"]), obs_dtype="int8")
    obs = env.reset_seeded(torch.from_numpy(rec["seeds"]).cuda())
    assert np.array_equal(env.x.cpu().numpy(), rec["init_x"]) and np.array_equal(env.y.cpu().numpy(), rec["init_y"])
    assert np.array_equal(obs.cpu().numpy(), rec["init_obs"])


def test_invalid_action_raises_value_error():
    cfg = readme_config()
    env = make_env(cfg, 16, auto_reset=False)
    env.reset()
    acts = torch.full((16, 8), 4, dtype=torch.int8, device="cuda")
    env.step(acts)
    env.check_error()
    acts[3, 2] = 7
    env.step(acts)
    with pytest.raises(ValueError, match="Invalid action"):
        env.check_error()
    env.check_error()  # cleared


def test_step_host_equals_device_path():
    cfg = readme_config()
    n = 513
    a = make_env(cfg, n, seed=1, obs_dtype="float32", with_info=True)
    b = make_env(cfg, n, seed=1, obs_dtype="float32", with_info=True)
    a.reset(); b.reset()
    host = b.make_host_buffers()
    rng = np.random.default_rng(0)
    for t in range(30):
        acts = torch.from_numpy(rng.integers(0, 5, size=(n, 8)).astype(np.int8))
        out = a.step(acts.cuda())
        host["actions"].copy_(acts)
        b.step_host(host)
        assert torch.equal(out.obs.cpu(), host["obs"]) and torch.equal(out.reward.cpu(), host["reward"])
        assert torch.equal(out.agent_flags.cpu(), host["agent_flags"]) and torch.equal(out.env_flags.cpu(), host["env_flags"])
        assert torch.equal(out.agent_info.cpu(), host["agent_info"])
    assert torch.equal(a.x, b.x)


def test_checkpoint_resume_reproduces_trajectory():
    cfg = readme_config(max_steps=30)
    env = make_env(cfg, 300, seed=9, obs_dtype="int8")
    env.reset()
    env.rollout(17, policy="waiting")
    snap = env.get_state()
    env.rollout(25, policy="waiting")
    want = (env.x.clone(), env.flags.clone(), env.step_count.clone(), env.episode_return.clone())
    other = make_env(cfg, 300, seed=9, obs_dtype="int8")
    other.load_state(snap)
    other.rollout(25, policy="waiting")
    for u, v in zip(want, (other.x, other.flags, other.step_count, other.episode_return)):
        assert torch.equal(u, v)


def test_policy_actions_kernel_matches_step_policy():
    cfg = readme_config()
    env = make_env(cfg, 1000, seed=2, obs_dtype="none")
    env.reset()
    env.rollout(9, policy="greedy")
    for pol in ("greedy", "waiting", "random"):
        standalone = env.policy_actions(pol, out=torch.zeros((1000, 8), dtype=torch.int8, device="cuda")).clone()
        snap = env.get_state()
        applied = env.step(policy=pol).actions.clone()
        if pol != "random":  # the random stream is keyed on the launch counter t, same for both here
            assert torch.equal(standalone, applied)
        else:
            assert torch.equal(standalone, applied)
        env.load_state(snap)


# ---- random geometries: every table the kernel builds (walkable map, x/y words, greedy rows, reward
# tables) against the oracle's straight-line arithmetic -----------------------------------------------
def _random_config(rng):
    from cases import REWARDS, TERMS, unchecked
    from collectivecrossing_b200.truncated_configs import MaxStepsTruncatedConfig

    while True:
        W, H = int(rng.integers(3, 60)), int(rng.integers(3, 40))
        D, L = int(rng.integers(1, H)), int(rng.integers(2, W + 1))
        dl = int(rng.integers(0, L))
        dr = int(rng.integers(dl, L))
        B, E = int(rng.integers(0, 12)), int(rng.integers(0, 9))
        half = L // 2
        tl, tr = W // 2 - half, W // 2 + half
        inside = max(0, tr - tl - 1)
        door = max(0, (tl + dr) - (tl + dl) - 1)
        free_tram = inside * (H - D - 1) + door
        free_wait = W * D - (dr - dl + 1)
        if B + E == 0 or E > free_tram // 2 or B > free_wait // 2:
            continue
        reward = list(REWARDS)[int(rng.integers(0, 4))]
        kw = {"default": dict(distance_penalty_factor=float(rng.choice([0.1, 0.37, 1.5])), tram_door_reward=7.5),
              "simple_distance": dict(distance_penalty_factor=0.3), "binary": dict(no_goal_reward=-0.25),
              "constant_negative": dict(step_penalty=-1.75)}[reward]
        return unchecked(
            width=W, height=H, division_y=D, tram_door_left=dl, tram_door_right=dr, tram_length=L,
            num_boarding_agents=B, num_exiting_agents=E, exiting_destination_area_y=int(rng.integers(0, D)),
            boarding_destination_area_y=int(rng.integers(D, H + 1)), reward_config=REWARDS[reward](**kw),
            terminated_config=list(TERMS.values())[int(rng.integers(0, 2))](),
            truncated_config=MaxStepsTruncatedConfig(max_steps=int(rng.integers(1, 50))))


@pytest.mark.parametrize("seed", range(12))
def test_random_geometries_match_oracle(seed):
    import oracle

    rng = np.random.default_rng(1000 + seed)
    cfg = _random_config(rng)
    low = lower_config(cfg)
    n = int(rng.integers(1, 200))
    x, y, f, s = random_states(cfg, n, rng)
    for policy in ("greedy", "waiting", "random"):
        env = make_env(cfg, n, seed=seed, global_env_offset=3, obs_dtype="int8", auto_reset=True, with_info=True)
        orc = oracle.OracleEnvs(low, n, seed=seed, global_env_offset=3)
        env.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
        orc.set_state(x, y, f, s)
        for t in range(40):
            out = env.step(policy=policy)
            res = orc.step(policy=policy, auto_reset=True, obs_dtype=_abi.OBS_INT8)
            _compare_step(env, orc, res, out, f"seed{seed}/{policy}/t={t} cfg={cfg.width}x{cfg.height} A={low.num_agents}", np.int8)
        env.check_error()
        env.close()


def test_impossible_placement_terminates_with_reset_stuck():
    """A tram with no interior cell: the reference's reset() would spin forever; the kernel stops after
    the attempt cap, places the agent on its last candidate (like the oracle) and reports the error."""
    import oracle
    from cases import unchecked

    cfg = unchecked(width=6, height=4, division_y=2, tram_door_left=0, tram_door_right=0, tram_length=1,
                    num_boarding_agents=1, num_exiting_agents=1, exiting_destination_area_y=0, boarding_destination_area_y=4)
    env = make_env(cfg, 5, seed=1, obs_dtype="int8")
    orc = oracle.OracleEnvs(lower_config(cfg), 5, seed=1)
    env.reset()
    with pytest.raises(RuntimeError):
        orc.reset()
    assert np.array_equal(env.x.cpu().numpy(), orc.x) and np.array_equal(env.y.cpu().numpy(), orc.y)
    with pytest.raises(Exception, match="no free valid cell"):
        env.check_error()


def test_rollout_equals_repeated_steps():
    cfg = readme_config(max_steps=40)
    a = make_env(cfg, 999, seed=4, obs_dtype="int8")
    b = make_env(cfg, 999, seed=4, obs_dtype="int8")
    a.reset(); b.reset()
    a.rollout(57, policy="greedy")
    for _ in range(57):
        out = b.step(policy="greedy")
    assert torch.equal(a.x, b.x) and torch.equal(a.flags, b.flags) and torch.equal(a.obs, out.obs)
    assert a.stats() == b.stats()


@pytest.mark.parametrize("case,policy,obs_dtype", [("A8_readme", "greedy", "float32"), ("A8_readme", "waiting", "int8"), ("A8_readme", "random", "none"),
                                                   ("A8_readme", "external", "float32"), ("A4", "greedy", "float32"), ("A5", "waiting", "int8"), ("A6", "greedy", "int8"),
                                                   ("A3_cassette", "greedy", "float32"), ("A7", "waiting", "float32"), ("A12", "greedy", "float32")])
def test_fused_rollout_equals_repeated_steps(case, policy, obs_dtype):
    """cc_rollout_fused (one launch, state in registers for T steps where the thread-per-env kernel
    applies; one launch per step otherwise) returns, for every step, exactly what T calls of step() return."""
    make_cfg, n = ORACLE_CASES[case]
    cfg = make_cfg()
    T = 37
    rng = np.random.default_rng(5)
    x, y, f, s = random_states(cfg, n, rng)
    envs = [make_env(cfg, n, seed=9, global_env_offset=77, obs_dtype=obs_dtype, auto_reset=True, with_info=True) for _ in range(2)]
    for e in envs:
        e.set_state(*(torch.from_numpy(v).cuda() for v in (x, y, f, s)))
    A = envs[0].num_agents
    acts = torch.from_numpy(rng.integers(0, 5, size=(T, n, A)).astype(np.int8)).cuda() if policy == "external" else None
    traj = envs[0].rollout_trajectory(T, policy=policy, actions=acts)
    fused = case in SMALL_CREWS and (obs_dtype != "int8" or case == "A8_readme")
    assert envs[0].last_kernel == ("threads" if fused else "lanes")
    assert envs[0].launch_count == (1 if fused else T)
    for t in range(T):
        out = envs[1].step(acts[t] if acts is not None else None, policy=policy)
        for name, got, want in (("reward", traj["reward"][t], out.reward), ("agent_flags", traj["agent_flags"][t], out.agent_flags),
                                ("agent_info", traj["agent_info"][t], out.agent_info), ("env_flags", traj["env_flags"][t], out.env_flags),
                                ("actions", traj["actions"][t], out.actions)):
            assert torch.equal(got, want), f"{case}/{policy}/t={t}: {name}"
        if obs_dtype != "none":
            assert torch.equal(traj["obs"][t], out.obs), f"{case}/{policy}/t={t}: obs"
    for name in ("x", "y", "flags", "step_count", "episode_return"):
        assert torch.equal(getattr(envs[0], name), getattr(envs[1], name)), name
    assert envs[0].stats() == pytest.approx(envs[1].stats()) and envs[0].step_counter == envs[1].step_counter
    for e in envs:
        e.check_error()
        e.close()


def test_fused_rollout_at_scale_equals_steps_and_keeps_invariants():
    """BASELINE config 2 / 5 shape at 262,144 envs: one fused launch of 12 steps against 12 single-step
    launches (every output tensor equal), then the reachable-state invariants on the result."""
    cfg = readme_config(max_steps=30)
    n, T = 1 << 18, 12
    a = make_env(cfg, n, seed=21, obs_dtype="float32", auto_reset=True, with_info=True)
    b = make_env(cfg, n, seed=21, obs_dtype="float32", auto_reset=True, with_info=True)
    a.reset()
    b.reset()
    for _ in range(25):       # into the regime where episodes end and envs are re-placed inside the launch
        a.step(policy="waiting")
        b.step(policy="waiting")
    traj = a.rollout_trajectory(T, policy="waiting")
    resets = 0
    for t in range(T):
        out = b.step(policy="waiting")
        resets += int(out.was_reset.sum())
        assert torch.equal(traj["obs"][t], out.obs) and torch.equal(traj["reward"][t], out.reward), t
        assert torch.equal(traj["agent_flags"][t], out.agent_flags) and torch.equal(traj["env_flags"][t], out.env_flags), t
        assert torch.equal(traj["agent_info"][t], out.agent_info) and torch.equal(traj["actions"][t], out.actions), t
    assert resets > 0
    for name in ("x", "y", "flags", "step_count", "episode_return"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    _check_invariants(a, cfg, b.step(policy="waiting") if False else out)
    sa, sb = a.stats(), b.stats()
    for k in ("env_steps", "episodes", "terminated_all", "truncated_all", "arrivals", "episode_length_sum"):
        assert sa[k] == sb[k], k
    a.check_error()
    a.close()
    b.close()


def test_envs_at_high_indices_equal_single_env_handles():
    """BASELINE config 5 scale on one GPU (4,194,304 envs): the RNG is keyed on the global env index, so env k of
    the big batch must follow exactly the trajectory of a one-env handle created with global_env_offset = k —
    checked at the first, middle and last indices, through single-step and fused launches."""
    cfg = readme_config(max_steps=20)
    n = 1 << 22
    big = make_env(cfg, n, seed=5, obs_dtype="float32", auto_reset=True)
    big.reset()
    ks = [0, 12345, n // 2 + 7, n - 33, n - 1]
    smalls = [make_env(cfg, 1, seed=5, global_env_offset=k, obs_dtype="float32", auto_reset=True) for k in ks]
    for s in smalls:
        s.reset()
    for t in range(26):
        out = big.step(policy="waiting")
        for k, s in zip(ks, smalls):
            o = s.step(policy="waiting")
            assert torch.equal(out.obs[k], o.obs[0]) and torch.equal(out.reward[k], o.reward[0]), (t, k)
            assert torch.equal(big.x[k], s.x[0]) and int(out.env_flags[k]) == int(o.env_flags[0]), (t, k)
    traj = big.rollout_trajectory(4, policy="waiting")
    for k, s in zip(ks, smalls):
        for t in range(4):
            o = s.step(policy="waiting")
            assert torch.equal(traj["obs"][t][k], o.obs[0]) and torch.equal(traj["reward"][t][k], o.reward[0]), (t, k)
    big.check_error()
    assert big.stats()["env_steps"] == 30 * n
    big.close()
    for s in smalls:
        s.close()


@pytest.mark.parametrize("name", ["readme_greedy", "readme_waiting", "readme_binary_all", "readme_simple_all"])
def test_fused_rollout_replays_reference_policy_recordings(name):
    """The reference's own rollouts (env + GreedyPolicy / WaitingPolicy at epsilon 0, recorded by
    tests/golden/make_golden.py) against ONE fused launch of all their steps: every step's positions are implied by
    the observations, so observations, flags, infos, policy actions are compared bit for bit and rewards as the
    correctly rounded float32 of the reference's float64."""
    cfg, kw = GOLDEN_CASES[name][0](), GOLDEN_CASES[name][1]
    rec = load_golden(name)
    T, N, A = rec["actions"].shape
    env = make_env(cfg, N, obs_dtype="float32", reward_dtype="float32", auto_reset=False, with_info=True)
    env.set_state(*(torch.from_numpy(rec[k]).cuda() for k in ("init_x", "init_y", "init_flags", "init_step")))
    traj = env.rollout_trajectory(T, policy=kw["source"])
    assert env.last_kernel == "threads" and env.launch_count == 1
    env.check_error()
    got = {"reward": traj["reward"].cpu().numpy().astype(np.float64), "agent_flags": traj["agent_flags"].cpu().numpy(),
           "agent_info": traj["agent_info"].cpu().numpy(), "env_flags": traj["env_flags"].cpu().numpy(),
           "obs": traj["obs"].cpu().numpy().astype(np.int8), "actions": traj["actions"].cpu().numpy(),
           "init_obs": rec["init_obs"]}
    slim = {k: v for k, v in rec.items() if k not in ("x", "y", "flags", "step")}   # per-step state is not kept by a fused launch
    assert_same(slim, got, f"fused/{name}", policy_actions=True, reward_rtol=1e-6)
    assert np.array_equal(got["reward"], rec["reward"].astype(np.float32).astype(np.float64))
    assert np.array_equal(env.x.cpu().numpy(), rec["x"][-1]) and np.array_equal(env.flags.cpu().numpy(), rec["flags"][-1])
    assert np.array_equal(env.step_count.cpu().numpy(), rec["step"][-1])
    env.close()
