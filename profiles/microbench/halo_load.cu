// Microbenchmark: how fast can 128 producer threads per SM move a stride-2 conv halo (64-byte pixel segments, every other
// pixel, 33 x 17 pixels per tile) from HBM into shared memory?  Variants: cp.async 16 B (LDGSTS), LDG.128 + STS.128 in
// batches, the same with 256 threads, and 128-byte segments (full pixels).  Build: nvcc -arch=sm_100a -O3 -o halo_load halo_load.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT, int SEG>   // SEG = bytes per pixel segment (64 or 128)
__global__ void __launch_bounds__(256) k(const uint8_t* __restrict__ in, int H, int W, int tiles, unsigned long long* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int CH = SEG / 16;                // 16-byte chunks per pixel
    const int n_px = 561, n_copy = n_px * CH;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tiles_x = W / 16, tiles_y = H / 32;
    uint32_t acc = 0;
    int stage = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int ty = (t / tiles_x) % tiles_y, tx = t % tiles_x, n = t / (tiles_x * tiles_y);
        const uint8_t* org = in + (((size_t)n * H + ty * 32) * W + tx * 16) * 128;
        uint8_t* dst = sm + stage * (SEG == 64 ? 40960 : 81920);
        if (VARIANT == 0) {
            for (int i = tid; i < n_copy; i += nt) {
                const int p = i / CH, c = i % CH;
                const int r = p / 17, x = p % 17;
                const uint8_t* src = org + ((size_t)r * W + x) * 128 + c * 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i * 16)), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else if (VARIANT == 2 || VARIANT == 3) {
            for (int half = 0; half < 2; ++half) {
                uint8_t* d2 = sm + (stage ^ half) * 40960;
                for (int i = tid; i < n_px * 4; i += nt) {
                    const int p = i >> 2, c = i & 3;
                    const int r = p / 17, x = p % 17;
                    const uint8_t* src = org + ((size_t)r * W + x) * 128 + half * 64 + c * 16;
                    if (VARIANT == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d2 + i * 16)), "l"(src) : "memory");
                    else asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(smem_u32(d2 + i * 16)), "l"(src) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            }
        } else {
            constexpr int B = 8;
            for (int i0 = tid; i0 < n_copy; i0 += nt * B) {
                uint4 v[B];
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int i = i0 + b * nt;
                    if (i < n_copy) {
                        const int p = i / CH, c = i % CH;
                        const int r = p / 17, x = p % 17;
                        v[b] = __ldg(reinterpret_cast<const uint4*>(org + ((size_t)r * W + x) * 128 + c * 16));
                    }
                }
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const int i = i0 + b * nt;
                    if (i < n_copy) *reinterpret_cast<uint4*>(dst + i * 16) = v[b];
                }
            }
        }
        stage ^= 1;
        acc += sm[(tid * 16) % 4096];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (acc == 0xffffffffu) *sink = acc;
}

template <int VARIANT, int SEG>
static void run(const char* name, const uint8_t* in, int N, int H, int W, int threads, int ctas_per_sm, unsigned long long* sink) {
    const int tiles = N * (H / 32) * (W / 16);
    cudaFuncSetAttribute(k<VARIANT, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 2; ++it) {
        cudaEventRecord(e0);
        k<VARIANT, SEG><<<148 * ctas_per_sm, threads, (SEG == 64 || ctas_per_sm == 2) ? 81920 : 163840>>>(in, H, W, tiles, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)tiles * 561 * SEG;
    printf("%-34s threads=%3d ctas/SM=%d  %.3f ms  %.0f GB/s  (%s)\n", name, threads, ctas_per_sm, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int N = 8, H = 352, W = 640;      // 8 x 352 x 640 pixels x 128 B = 230 MB (> L2)
    uint8_t* in; unsigned long long* sink;
    cudaMalloc(&in, (size_t)N * (H + 40) * W * 128);
    cudaMalloc(&sink, 8);
    cudaMemset(in, 1, (size_t)N * (H + 40) * W * 128);
    run<0, 64>("cp.async 16B, 64B segments", in, N, H, W, 128, 1, sink);
    run<0, 64>("cp.async 16B, 64B segments", in, N, H, W, 256, 1, sink);
    run<0, 64>("cp.async 16B, 64B segments", in, N, H, W, 128, 2, sink);
    run<1, 64>("LDG.128+STS.128 x8, 64B segments", in, N, H, W, 128, 1, sink);
    run<1, 64>("LDG.128+STS.128 x8, 64B segments", in, N, H, W, 256, 1, sink);
    run<1, 64>("LDG.128+STS.128 x8, 64B segments", in, N, H, W, 128, 2, sink);
    run<2, 128>("cp.async 2 x 64B half passes", in, N, H, W, 128, 1, sink);
    run<3, 128>("cp.async 2 x 64B half passes +L2::128B", in, N, H, W, 128, 1, sink);
    run<2, 128>("cp.async 2 x 64B half passes", in, N, H, W, 256, 1, sink);
    run<3, 128>("cp.async 2 x 64B half passes +L2::128B", in, N, H, W, 256, 1, sink);
    run<0, 128>("cp.async 16B, 128B segments", in, N, H, W, 128, 1, sink);
    run<0, 128>("cp.async 16B, 128B segments", in, N, H, W, 256, 1, sink);
    run<1, 128>("LDG.128+STS.128 x8, 128B segments", in, N, H, W, 128, 1, sink);
    run<1, 128>("LDG.128+STS.128 x8, 128B segments", in, N, H, W, 256, 1, sink);
    return 0;
}
