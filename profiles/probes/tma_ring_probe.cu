// tma_ring_probe.cu — how many bytes must a warp keep in flight for cp.async.bulk stores to saturate HBM?
// Every warp alternates a "compute" phase (dependent shared-memory chain, ~9 us) and a store phase that
// writes its 38,912-byte chunk as images of IMG bytes through a ring of RING shared-memory buffers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_ring_probe tma_ring_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChunkBytes = 38912;
constexpr int kThreads = 256;

template <int IMG, int RING, bool COMPUTE, int SMALL = 0>
__global__ void __launch_bounds__(kThreads) probe(unsigned char *out, int n_chunks, int *counter, int chase_len, unsigned *sink, unsigned char *small = nullptr) {
    __shared__ unsigned tab[1024];
    extern __shared__ __align__(128) unsigned char dyn[];   // [8 warps][RING][IMG]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = i * 7919u + 13u;
    __syncthreads();
    const uint2 v = make_uint2(1, threadIdx.x);
    if ((SMALL & 64) && threadIdx.x < 5) {
        // bulk L2 prefetch of this CTA's share of the five state arrays while HBM is still idle
        const size_t N = (size_t)n_chunks * 32;
        const size_t bytes = threadIdx.x < 3 ? N * 8 : N * 4;
        const unsigned char *base = threadIdx.x < 3 ? small + threadIdx.x * N * 8 : small + 3 * N * 8 + (threadIdx.x - 3) * N * 4;
        const size_t share = ((bytes + gridDim.x - 1) / gridDim.x + 15) / 16 * 16;
        const size_t lo = (size_t)blockIdx.x * share;
        if (lo < bytes) {
            const size_t len = lo + share <= bytes ? share : bytes - lo;
            for (size_t o = 0; o < len; o += 16384) {
                const unsigned sz = (unsigned)(len - o < 16384 ? len - o : 16384);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + lo + o), "r"(sz) : "memory");
            }
        }
    }
    unsigned char *ring = dyn + warp * RING * IMG;
    unsigned acc = threadIdx.x;
    int buf = 0;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= n_chunks) break;
        if (COMPUTE) for (int i = 0; i < chase_len; ++i) acc = tab[acc & 1023] * 2654435761u + (acc >> 3) + 1u;
        if (SMALL) {
            // the step kernel's other traffic.  state: x, y, flags (8 B per env each), step, return (4 B each), read and
            // written in place; outputs: actions_out, agent_flags, agent_info (8 B), reward (32 B), env_flags (1 B)
            const size_t n = (SMALL & 32) ? (((size_t)g * 32 + lane) & 8191) : ((size_t)g * 32 + lane), N = (size_t)n_chunks * 32;
            unsigned long long pol_last = 0, pol_first = 0;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
            unsigned long long r = n;
            if (SMALL & 128) {
                // state record of the group: [x 256 | y 256 | flags 256 | step 128 | return 128] = 1024 B, read and rewritten;
                // output record: [reward 1024 | actions 256 | agent_flags 256 | agent_info 256 | env_flags 32] = 1824 B
                unsigned char *srec = small + (size_t)g * 1024, *orec = small + (size_t)n_chunks * 1024 + (size_t)g * 1824;
                for (int a = 0; a < 3; ++a) r += *reinterpret_cast<const unsigned long long *>(srec + a * 256 + lane * 8);
                r += *reinterpret_cast<const unsigned *>(srec + 768 + lane * 4) + *reinterpret_cast<const unsigned *>(srec + 896 + lane * 4);
                acc += (unsigned)r;
                for (int a = 0; a < 3; ++a) *reinterpret_cast<unsigned long long *>(srec + a * 256 + lane * 8) = r + a;
                *reinterpret_cast<unsigned *>(srec + 768 + lane * 4) = acc;
                *reinterpret_cast<unsigned *>(srec + 896 + lane * 4) = acc + 1;
                reinterpret_cast<uint4 *>(orec)[2 * lane] = make_uint4(acc, 1, 2, 3);
                reinterpret_cast<uint4 *>(orec)[2 * lane + 1] = make_uint4(acc, 1, 2, 3);
                for (int a = 0; a < 3; ++a) *reinterpret_cast<unsigned long long *>(orec + 1024 + a * 256 + lane * 8) = r + a;
                orec[1792 + lane] = (unsigned char)acc;
            } else {
            auto ld8 = [&](const void *q) { unsigned long long v;
                if (SMALL & 8) asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(q), "l"(pol_last));
                else v = *reinterpret_cast<const unsigned long long *>(q);
                return v; };
            auto st8 = [&](void *q, unsigned long long v) {
                if (SMALL & 8) asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(q), "l"(v), "l"(pol_last) : "memory");
                else *reinterpret_cast<unsigned long long *>(q) = v; };
            if (SMALL & 1) {
                for (int a = 0; a < 3; ++a) r += ld8(small + a * N * 8 + n * 8);
                r += ld8(small + 3 * N * 8 + (n / 2) * 8);   // step + return (two 4-byte arrays, modelled as 8 bytes per 2 envs x 2)
                r += ld8(small + 3 * N * 8 + N * 4 + (n / 2) * 8);
            }
            acc += (unsigned)r;
            if (SMALL & 2) {
                for (int a = 0; a < 3; ++a) st8(small + a * N * 8 + n * 8, r + a);
                reinterpret_cast<unsigned *>(small + 3 * N * 8)[n] = acc;
                reinterpret_cast<unsigned *>(small + 3 * N * 8 + N * 4)[n] = acc + 1;
            }
            if (SMALL & 4) {
                for (int a = 4; a < 7; ++a) *reinterpret_cast<unsigned long long *>(small + a * N * 8 + n * 8) = r + a;
                reinterpret_cast<uint4 *>(small + 7 * N * 8)[2 * n] = make_uint4(acc, 1, 2, 3);
                reinterpret_cast<uint4 *>(small + 7 * N * 8)[2 * n + 1] = make_uint4(acc, 1, 2, 3);
                small[11 * N * 8 + n] = (unsigned char)acc;
            }
            }
        }
        for (int b = 0; b < kChunkBytes / IMG; ++b) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(RING - 1) : "memory");
            __syncwarp();
            uint2 *s = reinterpret_cast<uint2 *>(ring + buf * IMG) + lane;
#pragma unroll
            for (int j = 0; j < (IMG + 255) / 256; ++j)
                if (j * 256 + 256 <= IMG || lane * 8 + j * 256 < IMG) s[32 * j] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                unsigned char *dst = out + (long long)g * kChunkBytes + b * IMG;
                const unsigned src = (unsigned)__cvta_generic_to_shared(ring + buf * IMG);
                if (SMALL & 16) {
                    unsigned long long pf;
                    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "n"(IMG), "l"(pf) : "memory");
                } else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(IMG) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf = buf + 1 == RING ? 0 : buf + 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int IMG, int RING, bool COMPUTE, int SMALL = 0>
void run(unsigned char *out, int *counter, unsigned *sink, int sms, int cps, int chase_len, unsigned char *small = nullptr) {
    const int n_chunks = 32768, dyn = 8 * RING * IMG;
    auto kern = probe<IMG, RING, COMPUTE, SMALL>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(counter, 0, 4);
        cudaEventRecord(e0);
        kern<<<sms * cps, kThreads, dyn>>>(out, n_chunks, counter, chase_len, sink, small);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    printf("side traffic mode %2d, image %5d B x ring %d (%6d B in flight per warp), %d CTAs/SM, compute %d: %8.4f ms %7.1f GB/s (%s)\n", SMALL, IMG, RING, IMG * RING, cps,
           (int)COMPUTE, best, 32768.0 * kChunkBytes / 1e9 / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main() {
    unsigned char *out; cudaMalloc(&out, (size_t)32768 * kChunkBytes);
    int *counter; cudaMalloc(&counter, 4);
    unsigned *sink; cudaMalloc(&sink, 4 * 148 * 8 * kThreads);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int chase = 600;   // ~ 600 dependent LDS+ALU steps per chunk
    unsigned char *small; cudaMalloc(&small, (size_t)32768 * 32 * 8 * 12); cudaMemset(small, 0, (size_t)32768 * 32 * 8 * 12);
    {
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
        printf("L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
    }
    run<1216, 2, false, 0>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 1>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 2>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 3>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 4>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 8>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 16>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 8 + 16>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 1 + 32>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 32>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 6>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 6 + 32>(out, counter, sink, sms, 3, 0, small);
    for (size_t carve_mb : {36, 48, 64}) {
        // persisting L2 carve-out sized for the state arrays only (first 32 MB of `small`), normal policy for everything else
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve_mb << 20);
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.base_ptr = small;
        attr.accessPolicyWindow.num_bytes = (size_t)32 << 20;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cudaError_t e = cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &attr);
        printf("state window 32 MB persisting, carve-out %zu MB: %s\n", carve_mb, cudaGetErrorString(e));
        run<1216, 2, false, 3>(out, counter, sink, sms, 3, 0, small);
        run<1216, 2, false, 7>(out, counter, sink, sms, 3, 0, small);
        run<1216, 2, false, 7 + 16>(out, counter, sink, sms, 3, 0, small);
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
    }
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    run<1216, 2, false, 7>(out, counter, sink, sms, 3, 0, small);
    return 0;
    run<1216, 2, false, 128>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 128 + 16>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, true, 128 + 16>(out, counter, sink, sms, 3, 100, small);
    run<1216, 2, false, 1 + 64>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 64>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, false, 7 + 16 + 64>(out, counter, sink, sms, 3, 0, small);
    run<1216, 2, true, 7 + 16 + 64>(out, counter, sink, sms, 3, 100, small);
    run<1216, 2, true, 7 + 16>(out, counter, sink, sms, 3, 100, small);
    return 0;
    {
        // persisting L2 window over the state arrays (x, y, flags, step, return = first 32 MB) and then over state + outputs
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
        for (size_t win_mb : {32, 64, 97}) {
            size_t carve = (size_t)prop.persistingL2CacheMaxSize;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = small;
            attr.accessPolicyWindow.num_bytes = win_mb << 20;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t e = cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &attr);
            printf("persisting window %zu MB (carve-out %zu MB): %s\n", win_mb, carve >> 20, cudaGetErrorString(e));
            run<1216, 2, false, 7>(out, counter, sink, sms, 3, 0, small);
            run<1216, 2, false, 7 + 16>(out, counter, sink, sms, 3, 0, small);
            run<1216, 2, true, 7 + 16>(out, counter, sink, sms, 3, 100, small);
        }
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
    }
    return 0;
    for (int cps : {3}) {
        run<1216, 2, false>(out, counter, sink, sms, cps, chase);
        run<1216, 2, true>(out, counter, sink, sms, cps, chase);
        run<1216, 4, true>(out, counter, sink, sms, cps, chase);
        run<2432, 2, true>(out, counter, sink, sms, cps, chase);
        run<4864, 2, true>(out, counter, sink, sms, cps, chase);
        run<9728, 2, true>(out, counter, sink, sms, cps, chase);
        run<4864, 1, true>(out, counter, sink, sms, cps, chase);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
