// cc_api.cu — the extern "C" boundary declared in include/ccb200.h.
// Host-side plumbing only: handle, kernel dispatch, host<->device staging, statistics.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "cc_kernel_tpe.cuh"

#ifndef CCB_TPE_ALTERNATE
#define CCB_TPE_ALTERNATE 1   // single-step launches of the thread-per-env kernel alternate the direction in which groups are handed out
#endif
#include "cc_kernels.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CC_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) return fail(CC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct cc_handle {
    cc_config cfg;
    int64_t n_envs = 0;
    int device = 0;
    int64_t genv_offset = 0;
    uint64_t seed = 0;
    uint64_t t = 0;  // launches of step/reset so far: the RNG counter word
    int A = 0, lpe = 0, apl = 0, epw = 0, sm_count = 0;
    // persistent state (owned unless attached)
    int8_t *x = nullptr, *y = nullptr;
    uint8_t *flags = nullptr;
    int32_t *step = nullptr;
    float *ep_ret = nullptr;
    bool owns_state = false;
    void *own_block = nullptr;
    unsigned long long *stats = nullptr;
    int *err = nullptr;
    ccb::Pcg64State *gen = nullptr;  // per-env numpy-compatible generators (allocated on first seeded reset)
    bool gen_seeded = false;
    int64_t launches = 0;
    unsigned *tpe_counters = nullptr;   // two work counters of the thread-per-env kernel (launch parity)
    int64_t tpe_launches = 0;
    int variant = CC_KERNEL_AUTO;   // cc_set_kernel_variant
    int last_variant = 0;           // mapping of the last step launch (CC_KERNEL_LANES / CC_KERNEL_THREADS)
    // host-path staging
    void *stage_block = nullptr;
    size_t stage_bytes = 0;
    cudaStream_t host_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

using ccb::KParams;

void fill_params(const cc_handle *h, KParams &p, int obs_dtype, bool needs_bitmap) {
    const cc_config &c = h->cfg;
    memset(&p, 0, sizeof p);
    p.W = c.width; p.H = c.height; p.D = c.division_y; p.TL = c.tram_left; p.TR = c.tram_right;
    p.DL = c.door_left; p.DR = c.door_right; p.DC = (c.door_left + c.door_right) / 2;
    p.YB = c.boarding_dest_y; p.YE = c.exiting_dest_y; p.B = c.num_boarding; p.A = h->A;
    p.max_steps = c.max_steps; p.reward_kind = c.reward_kind; p.terminated_kind = c.terminated_kind;
    for (int i = 0; i < 4; ++i) { p.rp[i] = c.reward_params[i]; p.rpf[i] = (float)c.reward_params[i]; }
    p.reward_category_mask = c.reward_kind == CC_REWARD_DEFAULT ? 0xFu : 0u;
    p.n_envs = h->n_envs; p.genv_offset = (unsigned long long)h->genv_offset; p.seed = h->seed; p.t = (unsigned)h->t;
    p.x = h->x; p.y = h->y; p.flags = h->flags; p.step = h->step; p.ep_ret = h->ep_ret;
    p.stats = h->stats; p.err = h->err;
    p.R = 3 + 2 * h->A;
    p.pairs_per_env = h->A * p.R;
    p.lut_entries = obs_dtype == CC_OBS_NONE ? 0 : h->epw * p.pairs_per_env;
    p.stage_pairs = obs_dtype == CC_OBS_NONE ? 0 : round_up(h->epw * (2 * h->A + 4), 8);  // one row template per env
    p.walk_words = ((c.width + 3) * (c.height + 3) + 31) / 32;
    const int pair_bytes = obs_dtype == CC_OBS_FP32 ? 8 : 2;
    int off = round_up(p.lut_entries * 2, 16);
    p.off_stage = off;
    off += round_up(ccb::kWarpsPerCta * p.stage_pairs * pair_bytes, 16);
    p.off_bitmap = off;
    off += needs_bitmap ? round_up(ccb::kWarpsPerCta * h->epw * p.walk_words * 4, 16) : 0;
    p.off_desc = off;
    off += (obs_dtype != CC_OBS_NONE && h->lpe <= 16) ? ccb::kDescWords * 4 * ccb::kThreads : 0;
    // int8 rows of big crews (one env per warp): shifted template copies + vector lists (cc_kernels.cuh)
    const int env_bytes = h->A * (6 + 4 * h->A);
    if (obs_dtype == CC_OBS_INT8 && h->lpe == 32 && h->epw == 1 && env_bytes % 16 == 0 && env_bytes / 16 < 65536) {
        p.nvec_env = env_bytes / 16;
        p.shift_tst = round_up(14 + (2 * h->A + 4) * 2 + 16, 16);
        p.off_shift = off;
        off += ccb::kWarpsPerCta * 8 * p.shift_tst;
        p.off_vlist = off;
        off += round_up(p.nvec_env * 6 + 16, 16);
    }
    p.smem_total = off;
    p.n_groups = (h->n_envs + h->epw - 1) / h->epw;
}
int smem_bytes(const KParams &p) { return p.smem_total; }

template <int LPE, int APL, int OBS, int MODE>
int launch_t(cc_handle *h, const KParams &p, cudaStream_t s) {
    auto kern = ccb::cc_kernel<LPE, APL, OBS, MODE>;
    const int smem = smem_bytes(p);
    if (smem > 218 * 1024) return fail(CC_ERR_UNSUPPORTED, "configuration needs %d bytes of shared memory", smem);
    // the kernels also hold ~8 KB of static shared memory: opt in as soon as the sum can pass the 48 KB default
    if (smem > 36 * 1024) CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ccb::kThreads, smem));
    if (per_sm < 1) return fail(CC_ERR_UNSUPPORTED, "kernel does not fit on an SM (smem %d)", smem);
    // persistent grid: a whole number of waves of resident CTAs, never more CTAs than work
    long long want = (p.n_groups + ccb::kWarpsPerCta - 1) / ccb::kWarpsPerCta;
    long long cap = (long long)h->sm_count * per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, ccb::kThreads, smem, s>>>(p);
    CC_CUDA(cudaGetLastError());
    h->launches += 1;
    return CC_OK;
}

template <int LPE, int APL, int MODE>
int launch_obs(cc_handle *h, const KParams &p, int obs_dtype, cudaStream_t s) {
    if constexpr (MODE == ccb::kModePolicy) {
        return launch_t<LPE, APL, CC_OBS_NONE, MODE>(h, p, s);
    } else {
        switch (obs_dtype) {
        case CC_OBS_NONE:
            if constexpr (MODE == ccb::kModeObserve) return fail(CC_ERR_INVALID_ARG, "observe needs an observation dtype");
            else return launch_t<LPE, APL, CC_OBS_NONE, MODE>(h, p, s);
        case CC_OBS_INT8: return launch_t<LPE, APL, CC_OBS_INT8, MODE>(h, p, s);
        case CC_OBS_FP32: return launch_t<LPE, APL, CC_OBS_FP32, MODE>(h, p, s);
        }
        return fail(CC_ERR_INVALID_ARG, "unknown obs_dtype %d", obs_dtype);
    }
}

template <int MODE>
int launch(cc_handle *h, const KParams &p, int obs_dtype, cudaStream_t s) {
    if (MODE == ccb::kModePolicy) obs_dtype = CC_OBS_NONE;
    switch (h->lpe * 8 + h->apl) {
    case 4 * 8 + 1: return launch_obs<4, 1, MODE>(h, p, obs_dtype, s);
    case 8 * 8 + 1: return launch_obs<8, 1, MODE>(h, p, obs_dtype, s);
    case 16 * 8 + 1: return launch_obs<16, 1, MODE>(h, p, obs_dtype, s);
    case 32 * 8 + 1: return launch_obs<32, 1, MODE>(h, p, obs_dtype, s);
    case 32 * 8 + 2: return launch_obs<32, 2, MODE>(h, p, obs_dtype, s);
    case 32 * 8 + 4: return launch_obs<32, 4, MODE>(h, p, obs_dtype, s);
    }
    return fail(CC_ERR_UNSUPPORTED, "no kernel for %d agents", h->A);
}

// ---- thread-per-env step kernel (cc_kernel_tpe.cuh): crews of 4 or 8, agent order, float32 rewards ----
template <int A, int OBS>
int launch_tpe_t(cc_handle *h, KParams p, cudaStream_t s) {
    using L = ccb::TpeLayout<A, OBS>;
    auto kern = ccb::cc_step_tpe_kernel<A, OBS>;
    p.n_groups = (h->n_envs + 31) / 32;
    // greedy / waiting: a private lattice bitmap per thread while the lattice is small (README: 6 words)
    const bool on_device_policy = p.policy == CC_POLICY_GREEDY || p.policy == CC_POLICY_WAITING;
    // (with TMA rows the bitmap aliases the warp's image ring: 128 bytes per word)
    const int bm_cap = L::kTma ? (L::kImgRing * L::kImgBytes / 128 < ccb::kTpeMaxBitmapWords ? L::kImgRing * L::kImgBytes / 128 : ccb::kTpeMaxBitmapWords)
                               : ccb::kTpeMaxBitmapWords;
    p.tpe_bm_words = (on_device_policy && p.walk_words <= bm_cap) ? p.walk_words : 0;
    // one word per lattice row where the padded lattice has at most 32 columns and its rows fit the bitmap (README: 11 rows)
    if (on_device_policy && p.W + 3 <= 32 && p.H + 3 <= bm_cap) { p.tpe_bm_words = p.H + 3; p.tpe_bm_rows = 1; }
    // launch k counts its groups in counter k & 1 and zeroes the other one for launch k + 1 (launches of
    // one handle are stream-ordered by contract)
    if (p.n_steps < 1) p.n_steps = 1;
    p.slice_agents = (long long)h->n_envs * h->A;
    p.slice_envs = h->n_envs;
    p.slice_obs_bytes = OBS == CC_OBS_NONE ? 0 : (long long)h->n_envs * h->A * (6 + 4 * h->A) * (long long)OBS;
    p.tpe_reverse = CCB_TPE_ALTERNATE && p.n_steps == 1 ? (int)(h->tpe_launches & 1) : 0;
    p.tpe_counter = h->tpe_counters + (h->tpe_launches & 1);
    p.tpe_counter_next = h->tpe_counters + ((h->tpe_launches + 1) & 1);
    // (with TMA rows the bitmap aliases the warp's image ring)
    const int smem = L::kStageBytes + (L::kTma ? 0 : p.tpe_bm_words * ccb::kTpeThreads * 4);
    CC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ccb::kTpeThreads, smem));
    if (per_sm < 1) return fail(CC_ERR_UNSUPPORTED, "thread-per-env kernel does not fit on an SM (smem %d)", smem);
    long long want = (p.n_groups + ccb::kTpeWarps - 1) / ccb::kTpeWarps;
    long long cap = (long long)h->sm_count * per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, ccb::kTpeThreads, smem, s>>>(p);
    CC_CUDA(cudaGetLastError());
    h->launches += 1;
    h->tpe_launches += 1;
    return CC_OK;
}

bool aligned_to(const void *q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; }

// which step launches the thread-per-env kernel can serve
bool tpe_eligible(const cc_handle *h, const cc_step_io *io) {
    const int A = h->A;
    if (A < 1 || A > 8) return false;
    if (io->order || io->reward_dtype != CC_REWARD_F32) return false;      // dict order / float64 rewards: lane-group kernel
    if (io->obs_dtype == CC_OBS_INT8 && A != 8) return false;              // an env's int8 block must be whole 16-byte vectors
    if (A == 8 || A == 4) {                                                // rows are moved as one aligned word
        const void *rows[] = {h->x, h->y, h->flags, io->actions, io->actions_out, io->agent_flags, io->agent_info};
        for (const void *q : rows)
            if (q && !aligned_to(q, (uintptr_t)A)) return false;
    }
    return aligned_to(io->reward, A % 4 == 0 ? 16 : 4) && aligned_to(h->step, 4) && aligned_to(h->ep_ret, 4);
}

int launch_tpe(cc_handle *h, const KParams &p, int obs_dtype, cudaStream_t s) {
#define CCB_TPE_CASES(A_)                                                                   \
    case A_ * 8 + CC_OBS_NONE: return launch_tpe_t<A_, CC_OBS_NONE>(h, p, s);                   \
    case A_ * 8 + CC_OBS_FP32: return launch_tpe_t<A_, CC_OBS_FP32>(h, p, s);
    switch (h->A * 8 + obs_dtype) {
        CCB_TPE_CASES(1) CCB_TPE_CASES(2) CCB_TPE_CASES(3) CCB_TPE_CASES(4) CCB_TPE_CASES(5) CCB_TPE_CASES(6) CCB_TPE_CASES(7) CCB_TPE_CASES(8)
    case 8 * 8 + CC_OBS_INT8: return launch_tpe_t<8, CC_OBS_INT8>(h, p, s);
    }
#undef CCB_TPE_CASES
    return fail(CC_ERR_UNSUPPORTED, "no thread-per-env kernel for %d agents, obs_dtype %d", h->A, obs_dtype);
}

int check_io(const cc_handle *h, const cc_step_io *io) {
    if (!h || !io) return fail(CC_ERR_INVALID_ARG, "null handle or io");
    if (io->policy < CC_POLICY_EXTERNAL || io->policy > CC_POLICY_WAITING) return fail(CC_ERR_INVALID_ARG, "unknown policy %d", io->policy);
    if (io->policy == CC_POLICY_EXTERNAL && !io->actions) return fail(CC_ERR_INVALID_ARG, "policy EXTERNAL needs io->actions");
    if (io->obs_dtype != CC_OBS_NONE && io->obs_dtype != CC_OBS_INT8 && io->obs_dtype != CC_OBS_FP32)
        return fail(CC_ERR_INVALID_ARG, "unknown obs_dtype %d", io->obs_dtype);
    if (io->obs_dtype != CC_OBS_NONE && !io->obs) return fail(CC_ERR_INVALID_ARG, "obs_dtype set but io->obs is null");
    if (io->obs && (reinterpret_cast<uintptr_t>(io->obs) & 15)) return fail(CC_ERR_INVALID_ARG, "io->obs must be 16-byte aligned");
    if (io->reward_dtype != CC_REWARD_F32 && io->reward_dtype != CC_REWARD_F64) return fail(CC_ERR_INVALID_ARG, "unknown reward_dtype %d", io->reward_dtype);
    if (!io->reward || !io->agent_flags || !io->env_flags) return fail(CC_ERR_INVALID_ARG, "reward, agent_flags and env_flags are required");
    return CC_OK;
}

// n_steps > 1: one fused launch of the thread-per-env kernel; the caller has checked eligibility and that
// every output buffer of `io` is time-major [n_steps][...]
int step_on(cc_handle *h, const cc_step_io *io, cudaStream_t s, int n_steps = 1, long long obs_env_offset = 0) {
    KParams p;
    fill_params(h, p, io->obs_dtype, io->policy == CC_POLICY_GREEDY || io->policy == CC_POLICY_WAITING);
    p.n_steps = n_steps;
    p.obs_env_offset = obs_env_offset;
    p.actions = io->actions; p.order = io->order; p.actions_out = io->actions_out;
    p.obs = io->obs; p.reward = io->reward; p.agent_flags = io->agent_flags; p.agent_info = io->agent_info; p.env_flags = io->env_flags;
    p.policy = io->policy; p.auto_reset = io->auto_reset != 0; p.reward_f64 = io->reward_dtype == CC_REWARD_F64;
    const bool can_tpe = obs_env_offset == 0 && tpe_eligible(h, io);
    if (h->variant == CC_KERNEL_THREADS && !can_tpe)
        return fail(CC_ERR_UNSUPPORTED, "CC_KERNEL_THREADS was requested but this step is not eligible (needs at most 8 agents, agent order, "
                                        "float32 rewards, int8 rows only for 8 agents, word-aligned rows for 4 or 8 agents)");
    const bool use_tpe = can_tpe && h->variant != CC_KERNEL_LANES;
    if (!use_tpe) p.tpe_reverse = CCB_TPE_ALTERNATE ? (int)(h->t & 1) : 0;
    int rc = use_tpe ? launch_tpe(h, p, io->obs_dtype, s) : launch<ccb::kModeStep>(h, p, io->obs_dtype, s);
    if (n_steps > 1 && !use_tpe) return fail(CC_ERR_UNSUPPORTED, "fused multi-step launches need the thread-per-env kernel");
    if (rc == CC_OK) { h->t += (uint64_t)n_steps; h->last_variant = use_tpe ? CC_KERNEL_THREADS : CC_KERNEL_LANES; }
    return rc;
}

size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" {

const char *cc_last_error(void) { return g_err; }
int cc_abi_version(void) { return CCB200_ABI_VERSION; }

int cc_create(const cc_config *cfg, int64_t n_envs, int device, int64_t global_env_offset, uint64_t seed, cc_handle **out) {
    if (!cfg || !out) return fail(CC_ERR_INVALID_ARG, "null config or out pointer");
    *out = nullptr;
    const int A = cfg->num_boarding + cfg->num_exiting;
    if (cfg->num_boarding < 0 || cfg->num_exiting < 0 || A < 1 || A > CC_MAX_AGENTS)
        return fail(CC_ERR_UNSUPPORTED, "agents per env must be in 1..%d, got %d", CC_MAX_AGENTS, A);
    if (n_envs < 1) return fail(CC_ERR_INVALID_ARG, "n_envs must be positive");
    if (n_envs * (int64_t)A >= (int64_t)1 << 31) return fail(CC_ERR_UNSUPPORTED, "n_envs * agents must stay below 2^31 per handle (32-bit slot indices)");
    const int32_t geo[] = {cfg->width, cfg->height, cfg->division_y, cfg->tram_left, cfg->tram_right, cfg->door_left, cfg->door_right, cfg->boarding_dest_y, cfg->exiting_dest_y};
    for (int32_t v : geo)
        if (v < -1 || v > ccb::kMaxGeom) return fail(CC_ERR_UNSUPPORTED, "geometry value %d is outside the supported lattice (0..%d)", v, ccb::kMaxGeom);
    if (cfg->width < 1 || cfg->height < 1) return fail(CC_ERR_INVALID_ARG, "width and height must be positive");
    if (cfg->reward_kind < 0 || cfg->reward_kind > CC_REWARD_CONSTANT_NEGATIVE) return fail(CC_ERR_INVALID_ARG, "unknown reward_kind %d", cfg->reward_kind);
    if (cfg->terminated_kind < 0 || cfg->terminated_kind > CC_TERM_ALL_AT_DESTINATION) return fail(CC_ERR_INVALID_ARG, "unknown terminated_kind %d", cfg->terminated_kind);
    int count = 0;
    CC_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(CC_ERR_INVALID_ARG, "device %d out of range (%d visible)", device, count);
    DeviceGuard guard(device);

    cc_handle *h = new (std::nothrow) cc_handle();
    if (!h) return fail(CC_ERR_NOMEM, "out of host memory");
    h->cfg = *cfg; h->n_envs = n_envs; h->device = device; h->genv_offset = global_env_offset; h->seed = seed; h->A = A;
    h->lpe = 4; while (h->lpe < A && h->lpe < 32) h->lpe <<= 1;
    h->apl = A <= 32 ? 1 : (A <= 64 ? 2 : 4);
    h->epw = 32 / h->lpe;
    cudaError_t e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) { delete h; return fail(CC_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e)); }

    const size_t na = (size_t)n_envs * A;
    const size_t o_x = 0, o_y = align256(o_x + na), o_f = align256(o_y + na), o_s = align256(o_f + na),
                 o_r = align256(o_s + (size_t)n_envs * 4), o_st = align256(o_r + (size_t)n_envs * 4),
                 o_e = align256(o_st + ccb::kStCount * 8), o_c = align256(o_e + 4), total = align256(o_c + 8);
    e = cudaMalloc(&h->own_block, total);
    if (e != cudaSuccess) { delete h; return fail(CC_ERR_NOMEM, "cudaMalloc(%zu): %s", total, cudaGetErrorString(e)); }
    cudaMemset(h->own_block, 0, total);
    char *b = static_cast<char *>(h->own_block);
    h->x = reinterpret_cast<int8_t *>(b + o_x); h->y = reinterpret_cast<int8_t *>(b + o_y);
    h->flags = reinterpret_cast<uint8_t *>(b + o_f); h->step = reinterpret_cast<int32_t *>(b + o_s);
    h->ep_ret = reinterpret_cast<float *>(b + o_r); h->stats = reinterpret_cast<unsigned long long *>(b + o_st);
    h->err = reinterpret_cast<int *>(b + o_e);
    h->tpe_counters = reinterpret_cast<unsigned *>(b + o_c);
    h->owns_state = true;
    cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
    *out = h;
    return CC_OK;
}

void cc_destroy(cc_handle *h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    if (h->own_block) cudaFree(h->own_block);
    if (h->stage_block) cudaFree(h->stage_block);
    if (h->gen) cudaFree(h->gen);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
}

int cc_attach_state(cc_handle *h, int8_t *x, int8_t *y, uint8_t *flags, int32_t *step, float *episode_return) {
    if (!h || !x || !y || !flags || !step || !episode_return) return fail(CC_ERR_INVALID_ARG, "cc_attach_state: null pointer");
    h->x = x; h->y = y; h->flags = flags; h->step = step; h->ep_ret = episode_return;
    h->owns_state = false;
    return CC_OK;
}

static int copy_state(cc_handle *h, int8_t *x, int8_t *y, uint8_t *f, int32_t *st, bool to_device, cudaMemcpyKind kind, cudaStream_t s, bool sync) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(h->device);
    const size_t na = (size_t)h->n_envs * h->A;
    struct Item { void *dev; void *other; size_t bytes; } items[] = {
        {h->x, x, na}, {h->y, y, na}, {h->flags, f, na}, {h->step, st, (size_t)h->n_envs * 4}};
    for (auto &it : items) {
        if (!it.other) continue;
        if (to_device) CC_CUDA(cudaMemcpyAsync(it.dev, it.other, it.bytes, kind, s));
        else CC_CUDA(cudaMemcpyAsync(it.other, it.dev, it.bytes, kind, s));
    }
    if (to_device) CC_CUDA(cudaMemsetAsync(h->ep_ret, 0, (size_t)h->n_envs * 4, s));
    if (sync) CC_CUDA(cudaStreamSynchronize(s));
    return CC_OK;
}
int cc_set_state(cc_handle *h, const int8_t *x, const int8_t *y, const uint8_t *flags, const int32_t *step, void *stream) {
    return copy_state(h, const_cast<int8_t *>(x), const_cast<int8_t *>(y), const_cast<uint8_t *>(flags), const_cast<int32_t *>(step), true,
                      cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream), false);
}
int cc_get_state(cc_handle *h, int8_t *x, int8_t *y, uint8_t *flags, int32_t *step, void *stream) {
    return copy_state(h, x, y, flags, step, false, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream), false);
}
int cc_set_state_host(cc_handle *h, const int8_t *x, const int8_t *y, const uint8_t *flags, const int32_t *step) {
    return copy_state(h, const_cast<int8_t *>(x), const_cast<int8_t *>(y), const_cast<uint8_t *>(flags), const_cast<int32_t *>(step), true,
                      cudaMemcpyHostToDevice, nullptr, true);
}
int cc_get_state_host(cc_handle *h, int8_t *x, int8_t *y, uint8_t *flags, int32_t *step) {
    return copy_state(h, x, y, flags, step, false, cudaMemcpyDeviceToHost, nullptr, true);
}

int cc_step(cc_handle *h, const cc_step_io *io, void *stream) {
    int rc = check_io(h, io);
    if (rc != CC_OK) return rc;
    DeviceGuard guard(h->device);
    return step_on(h, io, static_cast<cudaStream_t>(stream));
}

int cc_rollout(cc_handle *h, const cc_step_io *io, int32_t n_steps, void *stream) {
    int rc = check_io(h, io);
    if (rc != CC_OK) return rc;
    if (n_steps < 0) return fail(CC_ERR_INVALID_ARG, "n_steps must be non-negative");
    if (io->policy == CC_POLICY_EXTERNAL) return fail(CC_ERR_INVALID_ARG, "cc_rollout needs an on-device policy");
    DeviceGuard guard(h->device);
    for (int i = 0; i < n_steps; ++i) {
        rc = step_on(h, io, static_cast<cudaStream_t>(stream));
        if (rc != CC_OK) return rc;
    }
    return CC_OK;
}

int cc_rollout_fused(cc_handle *h, const cc_step_io *io, int32_t n_steps, void *stream) {
    int rc = check_io(h, io);
    if (rc != CC_OK) return rc;
    if (n_steps < 1) return fail(CC_ERR_INVALID_ARG, "n_steps must be positive");
    if (io->order) return fail(CC_ERR_INVALID_ARG, "cc_rollout_fused moves agents in agent order (io->order must be NULL)");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (tpe_eligible(h, io) && h->variant != CC_KERNEL_LANES) return step_on(h, io, s, n_steps);   // state stays in registers
    if (h->variant == CC_KERNEL_THREADS) return fail(CC_ERR_UNSUPPORTED, "CC_KERNEL_THREADS was requested but this rollout is not eligible");
    // any other crew / dtype: one launch per step into the time slices
    const size_t na = (size_t)h->n_envs * h->A, n = (size_t)h->n_envs;
    for (int t = 0; t < n_steps; ++t) {
        cc_step_io d = *io;
        if (io->actions) d.actions = io->actions + t * na;
        if (io->actions_out) d.actions_out = io->actions_out + t * na;
        // (the observation slice is addressed through an env offset: t * obs_b need not be 16-byte aligned)
        d.reward = static_cast<char *>(io->reward) + t * na * (size_t)io->reward_dtype;
        d.agent_flags = io->agent_flags + t * na;
        if (io->agent_info) d.agent_info = io->agent_info + t * na;
        d.env_flags = io->env_flags + t * n;
        rc = step_on(h, &d, s, 1, (long long)t * h->n_envs);
        if (rc != CC_OK) return rc;
    }
    return CC_OK;
}

int cc_step_host(cc_handle *h, const cc_step_io *io) {
    int rc = check_io(h, io);
    if (rc != CC_OK) return rc;
    DeviceGuard guard(h->device);
    const size_t na = (size_t)h->n_envs * h->A, n = (size_t)h->n_envs;
    const size_t obs_b = io->obs_dtype == CC_OBS_NONE ? 0 : na * (6 + 4 * (size_t)h->A) * (size_t)io->obs_dtype;
    const size_t rew_b = na * (size_t)io->reward_dtype;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
    const size_t o_act = take(na), o_ord = take(na), o_ao = take(na), o_obs = take(obs_b), o_rew = take(rew_b), o_af = take(na), o_ai = take(na), o_ef = take(n);
    if (off > h->stage_bytes) {
        if (h->stage_block) CC_CUDA(cudaFree(h->stage_block));
        h->stage_block = nullptr; h->stage_bytes = 0;
        cudaError_t e = cudaMalloc(&h->stage_block, off);
        if (e != cudaSuccess) return fail(CC_ERR_NOMEM, "cudaMalloc(%zu) for host staging: %s", off, cudaGetErrorString(e));
        h->stage_bytes = off;
    }
    if (!h->host_stream) CC_CUDA(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->host_stream;
    char *b = static_cast<char *>(h->stage_block);
    cc_step_io d = *io;
    if (io->actions) { CC_CUDA(cudaMemcpyAsync(b + o_act, io->actions, na, cudaMemcpyHostToDevice, s)); d.actions = reinterpret_cast<int8_t *>(b + o_act); }
    if (io->order) { CC_CUDA(cudaMemcpyAsync(b + o_ord, io->order, na, cudaMemcpyHostToDevice, s)); d.order = reinterpret_cast<int8_t *>(b + o_ord); }
    d.actions_out = io->actions_out ? reinterpret_cast<int8_t *>(b + o_ao) : nullptr;
    d.obs = obs_b ? b + o_obs : nullptr;
    d.reward = b + o_rew;
    d.agent_flags = reinterpret_cast<uint8_t *>(b + o_af);
    d.agent_info = io->agent_info ? reinterpret_cast<uint8_t *>(b + o_ai) : nullptr;
    d.env_flags = reinterpret_cast<uint8_t *>(b + o_ef);
    rc = step_on(h, &d, s);
    if (rc != CC_OK) return rc;
    if (io->actions_out) CC_CUDA(cudaMemcpyAsync(io->actions_out, d.actions_out, na, cudaMemcpyDeviceToHost, s));
    if (obs_b) CC_CUDA(cudaMemcpyAsync(io->obs, d.obs, obs_b, cudaMemcpyDeviceToHost, s));
    CC_CUDA(cudaMemcpyAsync(io->reward, d.reward, rew_b, cudaMemcpyDeviceToHost, s));
    CC_CUDA(cudaMemcpyAsync(io->agent_flags, d.agent_flags, na, cudaMemcpyDeviceToHost, s));
    if (io->agent_info) CC_CUDA(cudaMemcpyAsync(io->agent_info, d.agent_info, na, cudaMemcpyDeviceToHost, s));
    CC_CUDA(cudaMemcpyAsync(io->env_flags, d.env_flags, n, cudaMemcpyDeviceToHost, s));
    CC_CUDA(cudaStreamSynchronize(s));
    return CC_OK;
}

int cc_reset(cc_handle *h, const uint8_t *mask, void *obs, int32_t obs_dtype, void *stream) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    if (obs_dtype != CC_OBS_NONE && !obs) return fail(CC_ERR_INVALID_ARG, "obs_dtype set but obs is null");
    if (!obs) obs_dtype = CC_OBS_NONE;
    if (obs && (reinterpret_cast<uintptr_t>(obs) & 15)) return fail(CC_ERR_INVALID_ARG, "obs must be 16-byte aligned");
    DeviceGuard guard(h->device);
    KParams p;
    fill_params(h, p, obs_dtype, false);
    p.mask = mask; p.obs = obs;
    int rc = launch<ccb::kModeReset>(h, p, obs_dtype, static_cast<cudaStream_t>(stream));
    if (rc == CC_OK) h->t += 1;
    return rc;
}

int cc_observe(cc_handle *h, void *obs, int32_t obs_dtype, void *stream) {
    if (!h || !obs) return fail(CC_ERR_INVALID_ARG, "null handle or obs");
    if (reinterpret_cast<uintptr_t>(obs) & 15) return fail(CC_ERR_INVALID_ARG, "obs must be 16-byte aligned");
    DeviceGuard guard(h->device);
    KParams p;
    fill_params(h, p, obs_dtype, false);
    p.obs = obs;
    return launch<ccb::kModeObserve>(h, p, obs_dtype, static_cast<cudaStream_t>(stream));
}

int cc_policy_actions(cc_handle *h, int32_t policy, int8_t *actions_out, void *stream) {
    if (!h || !actions_out) return fail(CC_ERR_INVALID_ARG, "null handle or actions_out");
    if (policy < CC_POLICY_RANDOM || policy > CC_POLICY_WAITING) return fail(CC_ERR_INVALID_ARG, "cc_policy_actions needs an on-device policy");
    DeviceGuard guard(h->device);
    KParams p;
    fill_params(h, p, CC_OBS_NONE, policy != CC_POLICY_RANDOM);
    p.policy = policy; p.actions_out = actions_out;
    return launch<ccb::kModePolicy>(h, p, CC_OBS_NONE, static_cast<cudaStream_t>(stream));
}

int cc_reset_seeded(cc_handle *h, const int64_t *seeds, void *obs, int32_t obs_dtype, void *stream) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    if (!seeds && !h->gen_seeded) return fail(CC_ERR_INVALID_ARG, "cc_reset_seeded(seeds = NULL) continues the stored generators: seed them first");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!h->gen) {
        cudaError_t e = cudaMalloc(&h->gen, (size_t)h->n_envs * sizeof(ccb::Pcg64State));
        if (e != cudaSuccess) return fail(CC_ERR_NOMEM, "cudaMalloc for generator states: %s", cudaGetErrorString(e));
    }
    KParams p;
    fill_params(h, p, CC_OBS_NONE, false);
    const int threads = 128;
    const long long blocks = (h->n_envs + threads - 1) / threads;
    ccb::cc_reset_seeded_kernel<<<(unsigned)blocks, threads, 0, s>>>(p, reinterpret_cast<const long long *>(seeds), h->gen);
    CC_CUDA(cudaGetLastError());
    h->launches += 1;
    h->gen_seeded = true;
    if (obs && obs_dtype != CC_OBS_NONE) return cc_observe(h, obs, obs_dtype, stream);
    return CC_OK;
}

int cc_stats_read(cc_handle *h, cc_stats *out, void *stream) {
    if (!h || !out) return fail(CC_ERR_INVALID_ARG, "null handle or out");
    DeviceGuard guard(h->device);
    static_assert(sizeof(cc_stats) == ccb::kStCount * 8, "cc_stats layout");
    CC_CUDA(cudaMemcpyAsync(out, h->stats, sizeof(cc_stats), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    CC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return CC_OK;
}
int cc_stats_reset(cc_handle *h, void *stream) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(h->device);
    CC_CUDA(cudaMemsetAsync(h->stats, 0, ccb::kStCount * 8, static_cast<cudaStream_t>(stream)));
    return CC_OK;
}
int cc_check_error(cc_handle *h, void *stream) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int bits = 0;
    CC_CUDA(cudaMemcpyAsync(&bits, h->err, sizeof bits, cudaMemcpyDeviceToHost, s));
    CC_CUDA(cudaStreamSynchronize(s));
    if (!bits) return CC_OK;
    CC_CUDA(cudaMemsetAsync(h->err, 0, sizeof bits, s));
    if (bits & ccb::kErrInvalidAction)
        return fail(CC_ERR_INVALID_ACTION, "Invalid action: some action is outside the valid actions [0, 1, 2, 3, 4] (or an order entry names an unknown agent)");
    return fail(CC_ERR_RESET_STUCK, "reset(): no free valid cell found within %d attempts for some agent", ccb::kResetAttemptCap);
}

int64_t cc_num_envs(const cc_handle *h) { return h ? h->n_envs : 0; }
int32_t cc_num_agents(const cc_handle *h) { return h ? h->A : 0; }
int32_t cc_obs_len(const cc_handle *h) { return h ? 6 + 4 * h->A : 0; }
uint64_t cc_step_counter(const cc_handle *h) { return h ? h->t : 0; }
int cc_set_step_counter(cc_handle *h, uint64_t t) { if (!h) return fail(CC_ERR_INVALID_ARG, "null handle"); h->t = t; return CC_OK; }
int64_t cc_launch_count(const cc_handle *h) { return h ? h->launches : 0; }
int cc_set_kernel_variant(cc_handle *h, int32_t variant) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    if (variant != CC_KERNEL_AUTO && variant != CC_KERNEL_LANES && variant != CC_KERNEL_THREADS) return fail(CC_ERR_INVALID_ARG, "unknown kernel variant %d", variant);
    h->variant = variant;
    return CC_OK;
}
int32_t cc_last_kernel_variant(const cc_handle *h) { return h ? h->last_variant : 0; }

int cc_timing_begin(cc_handle *h, void *stream) {
    if (!h) return fail(CC_ERR_INVALID_ARG, "null handle");
    DeviceGuard guard(h->device);
    CC_CUDA(cudaEventRecord(h->ev0, static_cast<cudaStream_t>(stream)));
    return CC_OK;
}
int cc_timing_end(cc_handle *h, void *stream, float *total_ms) {
    if (!h || !total_ms) return fail(CC_ERR_INVALID_ARG, "null handle or out");
    DeviceGuard guard(h->device);
    CC_CUDA(cudaEventRecord(h->ev1, static_cast<cudaStream_t>(stream)));
    CC_CUDA(cudaEventSynchronize(h->ev1));
    CC_CUDA(cudaEventElapsedTime(total_ms, h->ev0, h->ev1));
    return CC_OK;
}

}  // extern "C"
