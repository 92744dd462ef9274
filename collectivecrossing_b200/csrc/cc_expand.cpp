// cc_expand.cpp — host-side expansion of the compact observation table (CC_OBS_TABLE) into the reference's rows.
//
// observations.py:62-94 builds, for agent i of an env, the float32 vector
//      [x_i, y_i, door centre, division_y, door left, door right,  B_0, ..., B_(A-1)]
// with B_j = (x_j, y_j, type_j, active_j) for j != i and (-1, -1, -1, -1) for j == i.  All A rows of an env are
// therefore ONE table [A][4] with one block masked: the kernels can ship that table (4A bytes per env instead of
// 4A(6+4A)) and the consumer rebuilds the rows where it needs them.  This file is that consumer-side step for host
// memory: pure data movement, no env logic, no CUDA.  The rows it writes are bit-identical to what the step kernels
// write for CC_OBS_INT8 / CC_OBS_FP32 (tests/test_gpu_host_path.py).
//
// How a thread works: per env it converts the table once into a row TEMPLATE [0, 0, head, B_0 .. B_(A-1)]; a row is
// one copy of the template plus six patched values (own position, own block = -1).  Rows are assembled in a staging
// buffer whose offsets are congruent to the destination addresses modulo 64 and leave as whole cache lines through
// non-temporal stores (SSE2: every x86-64; 32-byte stores where the CPU has AVX2): the rows are written once and read
// by someone else, so read-for-ownership traffic would double the DRAM bytes.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif

#include "../../include/ccb200.h"

#if defined(__x86_64__)
#include <immintrin.h>
#define CCB_X86 1
#else
#define CCB_X86 0
#endif

int cc_fail(int code, const char *fmt, ...);

namespace {

constexpr size_t kMaxRowBytes = (size_t)(6 + 4 * CC_MAX_AGENTS) * sizeof(float);

#if CCB_X86
void stream_lines_sse2(unsigned char *dst, const unsigned char *src, size_t n) {   // n: multiple of 64; both 64-byte aligned
    for (size_t o = 0; o < n; o += 64) {
        const __m128i a = _mm_load_si128(reinterpret_cast<const __m128i *>(src + o)), b = _mm_load_si128(reinterpret_cast<const __m128i *>(src + o + 16)),
                      c = _mm_load_si128(reinterpret_cast<const __m128i *>(src + o + 32)), d = _mm_load_si128(reinterpret_cast<const __m128i *>(src + o + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + o), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + o + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + o + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(dst + o + 48), d);
    }
}
__attribute__((target("avx2"))) void stream_lines_avx2(unsigned char *dst, const unsigned char *src, size_t n) {
    for (size_t o = 0; o < n; o += 64) {
        const __m256i a = _mm256_load_si256(reinterpret_cast<const __m256i *>(src + o)), b = _mm256_load_si256(reinterpret_cast<const __m256i *>(src + o + 32));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + o), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + o + 32), b);
    }
}
using StreamFn = void (*)(unsigned char *, const unsigned char *, size_t);
StreamFn pick_stream() {
    __builtin_cpu_init();   // (a static initialiser of a dlopen'ed library: do not rely on libgcc's own constructor having run)
    return __builtin_cpu_supports("avx2") ? stream_lines_avx2 : stream_lines_sse2;
}
const StreamFn g_stream = pick_stream();
#endif

// Sequential writer of one thread's contiguous output range.  buf_[o] stands for the byte at base_ + o, base_ is
// 64-byte aligned: whole cache lines are drained with streaming stores, the ragged first and last bytes with memcpy.
class LineWriter {
  public:
    explicit LineWriter(unsigned char *dst) {
        lo_ = fill_ = (size_t)(reinterpret_cast<uintptr_t>(dst) & 63);
        base_ = dst - lo_;
    }
    unsigned char *cursor() { return buf_ + fill_; }          // room for one row of any crew
    void advance(size_t n) {
        fill_ += n;
        if (fill_ >= kDrain) drain(false);
    }
    void finish() { drain(true); }

  private:
    static constexpr size_t kDrain = 16384;
    void drain(bool last) {
        size_t o = lo_;
        if (o & 63) {   // the bytes in front of the first cache-line boundary
            const size_t k = std::min(fill_, (o + 63) & ~(size_t)63) - o;
            memcpy(base_ + o, buf_ + o, k);
            o += k;
        }
        const size_t lines = (fill_ - o) & ~(size_t)63;
#if CCB_X86
        g_stream(base_ + o, buf_ + o, lines);
#else
        memcpy(base_ + o, buf_ + o, lines);
#endif
        o += lines;
        if (last) {
            memcpy(base_ + o, buf_ + o, fill_ - o);
#if CCB_X86
            _mm_sfence();
#endif
            lo_ = fill_;
            return;
        }
        const size_t rest = fill_ - o;   // < 64: the open cache line moves to the front
        memcpy(buf_, buf_ + o, rest);
        base_ += o;
        lo_ = 0;
        fill_ = rest;
    }
    alignas(64) unsigned char buf_[kDrain + kMaxRowBytes + 64];
    unsigned char *base_;
    size_t lo_, fill_;
};

// AC: the crew size when it is a compile-time constant (the row copies become straight-line vector moves), 0 otherwise
template <typename T, int AC>
void expand_range(const cc_config *cfg, int64_t e0, int64_t e1, const int8_t *table, T *obs) {
    const int A = AC ? AC : cfg->num_boarding + cfg->num_exiting, L = 6 + 4 * A;
    const size_t row_bytes = (size_t)L * sizeof(T);
    alignas(64) T tmpl[6 + 4 * CC_MAX_AGENTS];
    // observations.py:66-75: door centre, division, door boundaries (per-config constants)
    tmpl[0] = tmpl[1] = (T)0;
    tmpl[2] = (T)((cfg->door_left + cfg->door_right) / 2); tmpl[3] = (T)cfg->division_y; tmpl[4] = (T)cfg->door_left; tmpl[5] = (T)cfg->door_right;
    const T masked[4] = {(T)-1, (T)-1, (T)-1, (T)-1};
    LineWriter w(reinterpret_cast<unsigned char *>(obs + e0 * (int64_t)A * L));
    for (int64_t e = e0; e < e1; ++e) {
        const int8_t *t = table + e * 4 * (int64_t)A;
        for (int k = 0; k < 4 * A; ++k) tmpl[6 + k] = (T)t[k];
        for (int i = 0; i < A; ++i) {
            unsigned char *p = w.cursor();
            memcpy(p, tmpl, row_bytes);
            memcpy(p, tmpl + 6 + 4 * i, 2 * sizeof(T));                       // own position (observations.py:62-64)
            memcpy(p + (size_t)(6 + 4 * i) * sizeof(T), masked, sizeof masked);   // the own block (observations.py:92-93)
            w.advance(row_bytes);
        }
    }
    w.finish();
}

#if CCB_X86
// int8 rows of the README crew (A = 8): an env's eight rows are 304 bytes = 19 vectors of 16, and every output byte is a table byte
// (one of 32), a per-config constant or -1.  Per vector: two byte shuffles of the table halves and a constant, no per-row work
// (the generic path spends ~55 cycles on each 38-byte row).  Needs SSSE3 and a 16-byte aligned destination.
struct Int8A8Plan {
    alignas(16) unsigned char lo[19][16], hi[19][16], fixed[19][16];
};
void plan_int8_a8(const cc_config *cfg, Int8A8Plan &pl) {
    const int8_t head[4] = {(int8_t)((cfg->door_left + cfg->door_right) / 2), (int8_t)cfg->division_y, (int8_t)cfg->door_left, (int8_t)cfg->door_right};
    for (int o = 0; o < 304; ++o) {
        const int i = o / 38, j = o % 38;
        int src = -1;                 // table byte, or -1: constant
        unsigned char c = 0;
        if (j < 2) src = 4 * i + j;                                   // own position (observations.py:62-64)
        else if (j < 6) c = (unsigned char)head[j - 2];               // observations.py:66-75
        else if ((j - 6) / 4 == i) c = 0xFF;                          // the own block: -1 (observations.py:92-93)
        else src = j - 6;
        pl.lo[o / 16][o % 16] = (src >= 0 && src < 16) ? (unsigned char)src : 0x80;
        pl.hi[o / 16][o % 16] = (src >= 16) ? (unsigned char)(src - 16) : 0x80;
        pl.fixed[o / 16][o % 16] = c;
    }
}
__attribute__((target("ssse3"))) void expand_int8_a8_ssse3(const Int8A8Plan &pl, int64_t e0, int64_t e1, const int8_t *table, int8_t *obs) {
    for (int64_t e = e0; e < e1; ++e) {
        const __m128i tlo = _mm_loadu_si128(reinterpret_cast<const __m128i *>(table + e * 32)),
                      thi = _mm_loadu_si128(reinterpret_cast<const __m128i *>(table + e * 32 + 16));
        __m128i *dst = reinterpret_cast<__m128i *>(obs + e * 304);
#pragma GCC unroll 19
        for (int q = 0; q < 19; ++q) {
            const __m128i v = _mm_or_si128(_mm_or_si128(_mm_shuffle_epi8(tlo, _mm_load_si128(reinterpret_cast<const __m128i *>(pl.lo[q]))),
                                                        _mm_shuffle_epi8(thi, _mm_load_si128(reinterpret_cast<const __m128i *>(pl.hi[q])))),
                                           _mm_load_si128(reinterpret_cast<const __m128i *>(pl.fixed[q])));
            _mm_stream_si128(dst + q, v);
        }
    }
    _mm_sfence();
}
const bool g_ssse3 = (__builtin_cpu_init(), __builtin_cpu_supports("ssse3"));
#endif

template <typename T>
void expand_any(const cc_config *cfg, int64_t e0, int64_t e1, const int8_t *table, T *obs) {
    switch (cfg->num_boarding + cfg->num_exiting) {
        case 8: return expand_range<T, 8>(cfg, e0, e1, table, obs);      // the README crew (BASELINE configs 1, 2, 4, 5)
        case 64: return expand_range<T, 64>(cfg, e0, e1, table, obs);    // BASELINE config 3
        default: return expand_range<T, 0>(cfg, e0, e1, table, obs);
    }
}

}  // namespace

// One thread's share (internal; cc_api.cu's host pipeline runs it on its own workers).
void cc_expand_rows_range(const cc_config *cfg, int64_t e0, int64_t e1, const int8_t *table, void *obs, int32_t obs_dtype) {
    if (e0 >= e1) return;
    if (obs_dtype == CC_OBS_FP32) { expand_any<float>(cfg, e0, e1, table, static_cast<float *>(obs)); return; }
#if CCB_X86
    if (g_ssse3 && cfg->num_boarding + cfg->num_exiting == 8 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {
        Int8A8Plan pl;
        plan_int8_a8(cfg, pl);
        expand_int8_a8_ssse3(pl, e0, e1, table, static_cast<int8_t *>(obs));
        return;
    }
#endif
    expand_any<int8_t>(cfg, e0, e1, table, static_cast<int8_t *>(obs));
}

// Whether rebuilding rows of this dtype on the host beats shipping them over PCIe (the automatic choice of the host path): float32 rows
// always (a thread streams ~10 GB/s of them), int8 rows where the shuffle path applies (27 GB/s per thread; the generic path: 4).
bool cc_expand_beats_pcie(const cc_config *cfg, int32_t obs_dtype) {
    if (obs_dtype == CC_OBS_FP32) return true;
#if CCB_X86
    return obs_dtype == CC_OBS_INT8 && g_ssse3 && cfg->num_boarding + cfg->num_exiting == 8;
#else
    (void)cfg;
    return false;
#endif
}

// Hardware threads this process may run on (the affinity mask where the OS has one: a rank bound to its GPU's NUMA node, a
// container pinned to a few cores), not the machine's total.
int cc_host_threads(void) {
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) return CPU_COUNT(&set);
#endif
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

// How many threads a job of n_envs gets when the caller asked for `requested` (<= 0: all this process may use).
int cc_expand_thread_count(int32_t requested, int64_t n_envs) {
    int threads = requested > 0 ? requested : cc_host_threads();
    if (threads < 1) threads = 1;
    const int64_t min_per_thread = 2048;   // below this a thread's start-up costs more than its share
    return (int)std::max<int64_t>(1, std::min<int64_t>(threads, (n_envs + min_per_thread - 1) / min_per_thread));
}

extern "C" int cc_expand_obs_host(const cc_config *cfg, int64_t n_envs, const int8_t *table, void *obs, int32_t obs_dtype, int32_t n_threads) {
    if (!cfg || !table || !obs) return cc_fail(CC_ERR_INVALID_ARG, "cc_expand_obs_host: null pointer");
    const int A = cfg->num_boarding + cfg->num_exiting;
    if (A < 1 || A > CC_MAX_AGENTS) return cc_fail(CC_ERR_UNSUPPORTED, "agents per env must be in 1..%d, got %d", CC_MAX_AGENTS, A);
    if (obs_dtype != CC_OBS_INT8 && obs_dtype != CC_OBS_FP32) return cc_fail(CC_ERR_INVALID_ARG, "cc_expand_obs_host writes CC_OBS_INT8 or CC_OBS_FP32 rows");
    if (n_envs <= 0) return CC_OK;
    const int threads = cc_expand_thread_count(n_threads, n_envs);
    if (threads <= 1) { cc_expand_rows_range(cfg, 0, n_envs, table, obs, obs_dtype); return CC_OK; }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    const int64_t per = (n_envs + threads - 1) / threads;
    for (int k = 1; k < threads; ++k) {
        const int64_t e0 = std::min<int64_t>(n_envs, k * per), e1 = std::min<int64_t>(n_envs, (k + 1) * per);
        if (e0 < e1) pool.emplace_back(cc_expand_rows_range, cfg, e0, e1, table, obs, obs_dtype);
    }
    cc_expand_rows_range(cfg, 0, std::min<int64_t>(n_envs, per), table, obs, obs_dtype);
    for (auto &th : pool) th.join();
    return CC_OK;
}
