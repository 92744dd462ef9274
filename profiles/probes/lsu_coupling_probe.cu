// lsu_coupling_probe.cu — do streaming global stores that are throttled by HBM slow down OTHER warps'
// shared-memory work on the same SM?  Half the warps of every CTA run a fixed dependent chain of
// shared-memory loads mixed with ALU work ("compute"), the other half stream 1.3 GB to HBM either with
// st.global (LSU) or by staging 4,864-byte images in shared memory and issuing cp.async.bulk (TMA).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lsu_coupling_probe lsu_coupling_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChunkBytes = 38912;   // 32 README envs of float32 observations
constexpr int kImgBytes = 4864;      // 4 envs
constexpr int kThreads = 256;

__device__ __forceinline__ unsigned chase(const unsigned *tab, unsigned idx, int n) {
    // dependent chain: LDS -> few ALU -> LDS ...
    for (int i = 0; i < n; ++i) idx = tab[idx & 1023] * 2654435761u + (idx >> 3) + 1u;
    return idx;
}

// mode 0: compute warps only; 1: store warps only (st.global); 2: both, st.global; 3: store warps only (TMA); 4: both, TMA
template <int MODE>
__global__ void __launch_bounds__(kThreads) probe(uint2 *out, int n_chunks, int *counter, int chase_len, int compute_iters,
                                                  unsigned *sink, unsigned long long *compute_cycles) {
    __shared__ unsigned tab[1024];
    extern __shared__ __align__(128) unsigned char dyn[];   // TMA images: [4 store warps][2][kImgBytes]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = i * 7919u + 13u;
    __syncthreads();
    const bool is_compute = warp < 4;
    if (is_compute) {
        if (MODE == 1 || MODE == 3) return;
        const long long t0 = clock64();
        unsigned acc = threadIdx.x;
        for (int it = 0; it < compute_iters; ++it) acc = chase(tab, acc + it, chase_len);
        const long long t1 = clock64();
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
        if (lane == 0) atomicAdd(compute_cycles, (unsigned long long)(t1 - t0));
        return;
    }
    if (MODE == 0) return;
    const uint2 v = make_uint2(1, threadIdx.x);
    unsigned char *img = dyn + (warp - 4) * 2 * kImgBytes;
    int buf = 0;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        if (MODE <= 2) {
            uint2 *o = out + (long long)g * (kChunkBytes / 8) + lane;
#pragma unroll 8
            for (int j = 0; j < kChunkBytes / 256; ++j) __stcs(o + 32 * j, v);
        } else {
            for (int b = 0; b < 8; ++b) {
                // the buffer must have been read by the bulk copy issued two blocks ago
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                uint2 *s = reinterpret_cast<uint2 *>(img + buf * kImgBytes) + lane;
#pragma unroll
                for (int j = 0; j < kImgBytes / 256; ++j) s[32 * j] = v;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    unsigned char *dst = reinterpret_cast<unsigned char *>(out) + (long long)g * kChunkBytes + b * kImgBytes;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(img + buf * kImgBytes);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(kImgBytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                buf ^= 1;
            }
        }
    }
    if (MODE >= 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int n_chunks = 32768;
    uint2 *out; cudaMalloc(&out, (size_t)n_chunks * kChunkBytes);
    int *counter; cudaMalloc(&counter, 4);
    unsigned *sink; cudaMalloc(&sink, 4 * 148 * 8 * kThreads);
    unsigned long long *cyc; cudaMalloc(&cyc, 8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int dyn = 4 * 2 * kImgBytes;
    auto run = [&](auto kern, const char *name, int cps, int chase_len, int iters) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9; unsigned long long c = 0;
        for (int rep = 0; rep < 6; ++rep) {
            cudaMemsetAsync(counter, 0, 4); cudaMemsetAsync(cyc, 0, 8);
            cudaEventRecord(e0);
            kern<<<sms * cps, kThreads, dyn>>>(out, n_chunks, counter, chase_len, iters, sink, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 2 && ms < best) { best = ms; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
        }
        printf("%-44s CTAs/SM %d: %8.4f ms, compute warps avg %9.0f cycles  (%s)\n", name, cps, best, (double)c / (sms * cps * 4.0), cudaGetErrorString(cudaGetLastError()));
        fflush(stdout);
    };
    for (int cps : {2, 3}) {
        // compute sized so that, alone, it takes roughly half of the store time
        const int chase_len = 64, iters = cps == 2 ? 60 : 40;
        run(probe<0>, "compute only", cps, chase_len, iters);
        run(probe<1>, "stores only, st.global.cs", cps, chase_len, iters);
        run(probe<2>, "compute + st.global.cs", cps, chase_len, iters);
        run(probe<3>, "stores only, smem image + cp.async.bulk", cps, chase_len, iters);
        run(probe<4>, "compute + smem image + cp.async.bulk", cps, chase_len, iters);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
