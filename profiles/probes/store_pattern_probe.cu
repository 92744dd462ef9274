// store_pattern_probe.cu — what HBM write bandwidth does the observation store pattern of the step
// kernel reach when nothing else runs?  (the ceiling the fused kernel is measured against.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern_probe store_pattern_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kChunkVec = 76 * 32;   // 16-byte vectors of one warp chunk: 32 README envs x 1216 B = 38,912 B

enum Hint { kWb = 0, kCs = 1, kCg = 2, kWt = 3 };
template <int HINT> __device__ __forceinline__ void st16(uint4 *p, uint4 v) {
    if (HINT == kCs) __stcs(p, v);
    else if (HINT == kCg) __stcg(p, v);
    else if (HINT == kWt) __stwt(p, v);
    else *p = v;
}

// P0: grid-stride, consecutive threads consecutive vectors (what a fill kernel does)
template <int HINT> __global__ void p_fill(uint4 *out, long long nvec) {
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) st16<HINT>(out + i, v);
}
// P1: persistent warps, warp w owns chunks w, w + W, ...; a chunk is written as 76 rows of 512 B
// (INTERLEAVE = 1: sequential rows; 4: four 9,728-byte blocks round-robin, as the kernel does)
template <int HINT, int INTERLEAVE> __global__ void p_chunks(uint4 *out, int n_chunks) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    for (int g = blockIdx.x * wpc + warp; g < n_chunks; g += gridDim.x * wpc) {
        uint4 *o = out + (long long)g * kChunkVec + lane;
        constexpr int JB = 76 / INTERLEAVE;
#pragma unroll 4
        for (int j = 0; j < JB; ++j)
#pragma unroll
            for (int b = 0; b < INTERLEAVE; ++b) st16<HINT>(o + b * JB * 32 + 32 * j, v);
    }
}
// P4: persistent warps, but every warp starts its chunk at a different row (desynchronised streams)
template <int HINT> __global__ void p_chunks_rot(uint4 *out, int n_chunks) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    const int w_global = blockIdx.x * wpc + warp;
    int it = 0;
    for (int g = w_global; g < n_chunks; g += gridDim.x * wpc, ++it) {
        uint4 *o = out + (long long)g * kChunkVec + lane;
        int j = (w_global * 29 + it * 13) % 76;
#pragma unroll 4
        for (int k = 0; k < 76; ++k) { st16<HINT>(o + 32 * j, v); j = j + 1 == 76 ? 0 : j + 1; }
    }
}
// P5: non-persistent, each warp writes K consecutive... chunks c*8K + k*8 + warp (CTA-contiguous)
template <int HINT> __global__ void p_chunks_k(uint4 *out, int n_chunks, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    for (int k = 0; k < K; ++k) {
        const int g = (blockIdx.x * K + k) * wpc + warp;
        if (g >= n_chunks) break;
        uint4 *o = out + (long long)g * kChunkVec + lane;
#pragma unroll 4
        for (int j = 0; j < 76; ++j) st16<HINT>(o + 32 * j, v);
    }
}
// P6: persistent CTAs, dynamic chunk assignment through an atomic counter (8 chunks per grab)
template <int HINT> __global__ void p_chunks_dyn(uint4 *out, int n_chunks, int *counter) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    __shared__ int base;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) base = atomicAdd(counter, wpc);
        __syncthreads();
        const int g = base + warp;
        if (base >= n_chunks) break;
        if (g < n_chunks) {
            uint4 *o = out + (long long)g * kChunkVec + lane;
#pragma unroll 4
            for (int j = 0; j < 76; ++j) st16<HINT>(o + 32 * j, v);
        }
    }
}

// P2: like P1 but the CTA's warps write ONE chunk at a time together (8 warps x 512 B rows): fewer, faster streams
template <int HINT> __global__ void p_cta_chunks(uint4 *out, int n_chunks) {
    const uint4 v = make_uint4(1, 2, 3, threadIdx.x);
    for (int g = blockIdx.x; g < n_chunks; g += gridDim.x) {
        uint4 *o = out + (long long)g * kChunkVec;
        for (int i = threadIdx.x; i < kChunkVec; i += blockDim.x) st16<HINT>(o + i, v);
    }
}
// P3: 8-byte stores, lane <-> pair (256 B per instruction)
template <int HINT> __global__ void p_chunks_pair(uint2 *out, int n_chunks) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint2 v = make_uint2(1, threadIdx.x);
    for (int g = blockIdx.x * wpc + warp; g < n_chunks; g += gridDim.x * wpc) {
        uint2 *o = out + (long long)g * kChunkVec * 2 + lane;
#pragma unroll 8
        for (int j = 0; j < 152; ++j) {
            if (HINT == kCs) __stcs(o + 32 * j, v); else o[32 * j] = v;
        }
    }
}

template <typename F> float time_ms(F launch, int reps = 20) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int n_chunks = 32768;                        // 1,048,576 envs
    const long long nvec = (long long)n_chunks * kChunkVec;
    const double gb = nvec * 16 / 1e9;
    uint4 *out; cudaMalloc(&out, nvec * 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char *name, float ms) { printf("%-58s %8.4f ms %8.1f GB/s\n", name, ms, gb / (ms * 1e-3)); fflush(stdout); };
    report("P0 fill grid-stride wb (sms*8 x 256)", time_ms([&] { p_fill<kWb><<<sms * 8, 256>>>(out, nvec); }));
    report("P0 fill grid-stride cs", time_ms([&] { p_fill<kCs><<<sms * 8, 256>>>(out, nvec); }));
    for (int cps : {2, 3, 4, 6, 8}) {
        char nm[128];
        snprintf(nm, sizeof nm, "P1 warp chunks seq rows cs, %d CTAs/SM x 8 warps", cps);
        report(nm, time_ms([&] { p_chunks<kCs, 1><<<sms * cps, 256>>>(out, n_chunks); }));
    }
    report("P1 warp chunks seq rows wb, 3 CTAs/SM", time_ms([&] { p_chunks<kWb, 1><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P1 warp chunks seq rows cg, 3 CTAs/SM", time_ms([&] { p_chunks<kCg, 1><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P1 warp chunks seq rows wt, 3 CTAs/SM", time_ms([&] { p_chunks<kWt, 1><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P1 warp chunks 4-way interleave cs, 3 CTAs/SM", time_ms([&] { p_chunks<kCs, 4><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P1 warp chunks 4-way interleave wb, 3 CTAs/SM", time_ms([&] { p_chunks<kWb, 4><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P1 non-persistent (1 chunk per warp) cs", time_ms([&] { p_chunks<kCs, 1><<<n_chunks / 8, 256>>>(out, n_chunks); }));
    report("P2 CTA chunks cs, 3 CTAs/SM", time_ms([&] { p_cta_chunks<kCs><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P2 CTA chunks cs, 8 CTAs/SM", time_ms([&] { p_cta_chunks<kCs><<<sms * 8, 256>>>(out, n_chunks); }));
    report("P2 CTA chunks wb, non-persistent", time_ms([&] { p_cta_chunks<kWb><<<n_chunks, 256>>>(out, n_chunks); }));
    report("P3 warp chunks 8-byte pairs cs, 3 CTAs/SM", time_ms([&] { p_chunks_pair<kCs><<<sms * 3, 256>>>((uint2 *)out, n_chunks); }));
    report("P3 warp chunks 8-byte pairs wb, 3 CTAs/SM", time_ms([&] { p_chunks_pair<kWb><<<sms * 3, 256>>>((uint2 *)out, n_chunks); }));
    report("P4 persistent rotated rows cs, 3 CTAs/SM", time_ms([&] { p_chunks_rot<kCs><<<sms * 3, 256>>>(out, n_chunks); }));
    report("P4 persistent rotated rows wb, 3 CTAs/SM", time_ms([&] { p_chunks_rot<kWb><<<sms * 3, 256>>>(out, n_chunks); }));
    for (int K : {1, 2, 3, 4, 8}) {
        char nm[128];
        snprintf(nm, sizeof nm, "P5 non-persistent, %d chunks per warp, cs", K);
        report(nm, time_ms([&] { p_chunks_k<kCs><<<(n_chunks / 8 + K - 1) / K, 256>>>(out, n_chunks, K); }));
        snprintf(nm, sizeof nm, "P5 non-persistent, %d chunks per warp, wb", K);
        report(nm, time_ms([&] { p_chunks_k<kWb><<<(n_chunks / 8 + K - 1) / K, 256>>>(out, n_chunks, K); }));
    }
    int *counter; cudaMalloc(&counter, 4);
    report("P6 persistent dynamic (atomic counter) cs, 3 CTAs/SM", time_ms([&] { cudaMemsetAsync(counter, 0, 4); p_chunks_dyn<kCs><<<sms * 3, 256>>>(out, n_chunks, counter); }));
    report("P6 persistent dynamic (atomic counter) cs, 6 CTAs/SM", time_ms([&] { cudaMemsetAsync(counter, 0, 4); p_chunks_dyn<kCs><<<sms * 6, 256>>>(out, n_chunks, counter); }));
    for (int thr : {32, 64, 128, 256, 512})
        for (int cps : {1, 2}) {
            char nm[128];
            snprintf(nm, sizeof nm, "P6 dynamic cs, %d CTA/SM x %d threads (%d storing warps/SM)", cps, thr, cps * thr / 32);
            report(nm, time_ms([&] { cudaMemsetAsync(counter, 0, 4); p_chunks_dyn<kCs><<<sms * cps, thr>>>(out, n_chunks, counter); }));
        }
    report("P1 non-persistent 128-thread CTAs cs", time_ms([&] { p_chunks<kCs, 1><<<n_chunks / 4, 128>>>(out, n_chunks); }));
    report("P1 non-persistent 512-thread CTAs cs", time_ms([&] { p_chunks<kCs, 1><<<n_chunks / 16, 512>>>(out, n_chunks); }));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
