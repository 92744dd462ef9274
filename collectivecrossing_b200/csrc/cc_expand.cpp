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
// Output is written with non-temporal 16-byte stores where the ISA has them (SSE2: every x86-64): the rows are
// written once and read by someone else, so read-for-ownership traffic would double the DRAM bytes.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/ccb200.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

int cc_fail(int code, const char *fmt, ...);

namespace {

// Sequential writer of one thread's contiguous output range: small pieces are appended to an aligned local buffer
// that is flushed with streaming stores.
class StreamWriter {
  public:
    explicit StreamWriter(unsigned char *dst) : dst_(dst), fill_(0) {
        // bytes in front of the first 16-byte boundary go out directly
        head_ = (size_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    }
    void push(const void *src, size_t n) {
        const unsigned char *s = static_cast<const unsigned char *>(src);
        while (head_ && n) { *dst_++ = *s++; --head_; --n; }
        while (n) {
            const size_t k = std::min(n, kBuf - fill_);
            memcpy(buf_ + fill_, s, k);
            fill_ += k; s += k; n -= k;
            if (fill_ == kBuf) flush_full();
        }
    }
    void finish() {
        const size_t whole = fill_ & ~(size_t)15;
        stream(whole);
        memcpy(dst_, buf_ + whole, fill_ - whole);
        dst_ += fill_ - whole;
        fill_ = 0;
#if defined(__SSE2__)
        _mm_sfence();
#endif
    }

  private:
    static constexpr size_t kBuf = 8192;
    void stream(size_t n) {   // n is a multiple of 16, dst_ is 16-byte aligned
#if defined(__SSE2__)
        for (size_t o = 0; o < n; o += 16)
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst_ + o), _mm_load_si128(reinterpret_cast<const __m128i *>(buf_ + o)));
#else
        memcpy(dst_, buf_, n);
#endif
        dst_ += n;
    }
    void flush_full() { stream(kBuf); fill_ = 0; }
    alignas(64) unsigned char buf_[kBuf];
    unsigned char *dst_;
    size_t fill_, head_;
};

template <typename T>
void expand_range(const cc_config *cfg, int64_t e0, int64_t e1, const int8_t *table, T *obs) {
    const int A = cfg->num_boarding + cfg->num_exiting, L = 6 + 4 * A;
    // observations.py:66-75: door centre, division, door boundaries (per-config constants)
    const T head[4] = {(T)((cfg->door_left + cfg->door_right) / 2), (T)cfg->division_y, (T)cfg->door_left, (T)cfg->door_right};
    const T masked[4] = {(T)-1, (T)-1, (T)-1, (T)-1};
    T blocks[4 * CC_MAX_AGENTS];
    StreamWriter w(reinterpret_cast<unsigned char *>(obs + e0 * (int64_t)A * L));
    for (int64_t e = e0; e < e1; ++e) {
        const int8_t *t = table + e * 4 * (int64_t)A;
        for (int k = 0; k < 4 * A; ++k) blocks[k] = (T)t[k];
        for (int i = 0; i < A; ++i) {
            w.push(blocks + 4 * i, 2 * sizeof(T));                 // own position (observations.py:62-64)
            w.push(head, sizeof head);
            if (i) w.push(blocks, (size_t)(4 * i) * sizeof(T));    // agents before i
            w.push(masked, sizeof masked);                         // the own block (observations.py:92-93)
            if (i + 1 < A) w.push(blocks + 4 * (i + 1), (size_t)(4 * (A - 1 - i)) * sizeof(T));
        }
    }
    w.finish();
}

}  // namespace

extern "C" int cc_expand_obs_host(const cc_config *cfg, int64_t n_envs, const int8_t *table, void *obs, int32_t obs_dtype, int32_t n_threads) {
    if (!cfg || !table || !obs) return cc_fail(CC_ERR_INVALID_ARG, "cc_expand_obs_host: null pointer");
    const int A = cfg->num_boarding + cfg->num_exiting;
    if (A < 1 || A > CC_MAX_AGENTS) return cc_fail(CC_ERR_UNSUPPORTED, "agents per env must be in 1..%d, got %d", CC_MAX_AGENTS, A);
    if (obs_dtype != CC_OBS_INT8 && obs_dtype != CC_OBS_FP32) return cc_fail(CC_ERR_INVALID_ARG, "cc_expand_obs_host writes CC_OBS_INT8 or CC_OBS_FP32 rows");
    if (n_envs <= 0) return CC_OK;
    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const int64_t min_per_thread = 2048;   // below this a thread's start-up costs more than its share
    threads = (int)std::min<int64_t>(threads, (n_envs + min_per_thread - 1) / min_per_thread);
    auto run = [&](int64_t e0, int64_t e1) {
        if (obs_dtype == CC_OBS_FP32) expand_range<float>(cfg, e0, e1, table, static_cast<float *>(obs));
        else expand_range<int8_t>(cfg, e0, e1, table, static_cast<int8_t *>(obs));
    };
    if (threads <= 1) { run(0, n_envs); return CC_OK; }
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    const int64_t per = (n_envs + threads - 1) / threads;
    for (int k = 1; k < threads; ++k) {
        const int64_t e0 = std::min<int64_t>(n_envs, k * per), e1 = std::min<int64_t>(n_envs, (k + 1) * per);
        if (e0 < e1) pool.emplace_back(run, e0, e1);
    }
    run(0, std::min<int64_t>(n_envs, per));
    for (auto &th : pool) th.join();
    return CC_OK;
}
