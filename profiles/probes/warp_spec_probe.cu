// warp_spec_probe.cu — skeleton of the warp-specialised step kernel: P producer warps per CTA run a
// latency-bound "compute" phase per chunk and publish a 4,864-byte template (one slot per producer);
// E emitter warps expand ready templates 8x into 4,864-byte images and stream them out with
// cp.async.bulk.  Question: does the handoff through shared memory give full compute / store overlap?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o warp_spec_probe warp_spec_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChunkBytes = 38912, kTplBytes = 4864, kImgBytes = 4864;

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory"); return v;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// slot state: 0 = empty, 1 + g = template of chunk g ready, 0xFFFFFFFF = producer finished
template <int P, int E>
__global__ void __launch_bounds__((P + E) * 32) probe(unsigned char *out, int n_chunks, int *counter, int chase_len, unsigned *sink) {
    __shared__ unsigned tab[1024];
    __shared__ unsigned state[P];
    extern __shared__ __align__(128) unsigned char dyn[];   // [P][kTplBytes] templates, then [E][2][kImgBytes] images
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = i * 7919u + 13u;
    if (threadIdx.x < P) state[threadIdx.x] = 0;
    __syncthreads();
    if (warp < P) {
        // ---------------- producer ----------------
        unsigned char *tpl = dyn + warp * kTplBytes;
        unsigned acc = threadIdx.x;
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(counter, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= n_chunks) break;
            for (int i = 0; i < chase_len; ++i) acc = tab[acc & 1023] * 2654435761u + (acc >> 3) + 1u;
            if (lane == 0) while (ld_acquire(&state[warp]) != 0u) __nanosleep(64);
            __syncwarp();
            uint2 *t = reinterpret_cast<uint2 *>(tpl) + lane;
#pragma unroll
            for (int j = 0; j < kTplBytes / 256; ++j) t[32 * j] = make_uint2(acc, g);
            __syncwarp();
            if (lane == 0) st_release(&state[warp], 1u + (unsigned)g);
        }
        if (lane == 0) { while (ld_acquire(&state[warp]) != 0u) __nanosleep(64); st_release(&state[warp], 0xFFFFFFFFu); }
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    } else {
        // ---------------- emitter: serves producers e, e + E, ... ----------------
        const int e = warp - P;
        unsigned char *ring = dyn + P * kTplBytes + e * 2 * kImgBytes;
        int buf = 0, live = 0;
        for (int q = e; q < P; q += E) ++live;
        while (live > 0) {
            bool idle = true;
            for (int q = e; q < P; q += E) {
                unsigned s = 0;
                if (lane == 0) s = ld_acquire(&state[q]);
                s = __shfl_sync(0xffffffffu, s, 0);
                if (s == 0u || s == 0xFFFFFFFEu) continue;
                if (s == 0xFFFFFFFFu) { if (lane == 0) st_release(&state[q], 0xFFFFFFFEu); --live; continue; }
                idle = false;
                const int g = (int)(s - 1u);
                const uint2 *t = reinterpret_cast<const uint2 *>(dyn + q * kTplBytes) + lane;
                for (int b = 0; b < kChunkBytes / kImgBytes; ++b) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    uint2 *im = reinterpret_cast<uint2 *>(ring + buf * kImgBytes) + lane;
#pragma unroll
                    for (int j = 0; j < kImgBytes / 256; ++j) im[32 * j] = t[32 * j];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        unsigned char *dst = out + (long long)g * kChunkBytes + b * kImgBytes;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                                     "r"((unsigned)__cvta_generic_to_shared(ring + buf * kImgBytes)), "n"(kImgBytes) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    buf ^= 1;
                }
                __syncwarp();
                if (lane == 0) st_release(&state[q], 0u);
            }
            if (idle) __nanosleep(100);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int P, int E>
void run(unsigned char *out, int *counter, unsigned *sink, int sms, int cps, int chase_len) {
    const int n_chunks = 32768, dyn = P * kTplBytes + E * 2 * kImgBytes;
    auto kern = probe<P, E>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
        cudaMemsetAsync(counter, 0, 4);
        cudaEventRecord(e0);
        kern<<<sms * cps, (P + E) * 32, dyn>>>(out, n_chunks, counter, chase_len, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    printf("P %2d producers + E %d emitters per CTA, %d CTA/SM (smem %3d KB), chase %4d: %8.4f ms %7.1f GB/s (%s)\n", P, E, cps, dyn / 1024, chase_len, best,
           32768.0 * kChunkBytes / 1e9 / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main() {
    unsigned char *out; cudaMalloc(&out, (size_t)32768 * kChunkBytes);
    int *counter; cudaMalloc(&counter, 4);
    unsigned *sink; cudaMalloc(&sink, 4 * 148 * 4 * 1024);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int chase : {0, 300, 600}) {
        run<20, 4>(out, counter, sink, sms, 1, chase);
        run<16, 8>(out, counter, sink, sms, 1, chase);
        run<24, 8>(out, counter, sink, sms, 1, chase);
        run<10, 2>(out, counter, sink, sms, 2, chase);
        run<6, 2>(out, counter, sink, sms, 3, chase);
        run<12, 4>(out, counter, sink, sms, 2, chase);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
