// Microbenchmark (companion of halo_load.cu): the same 33 x 17 pixel halo tiles fetched by the TMA engine.
//   variant A: box {32 ch (64 B), 17, 33}           -> pixel-major [y][x][64 B]
//   variant B: box {8 ch (16 B), 17, 33, 4 k-chunks}  (k-chunk = extra dimension with a 16-byte stride)
//              -> [k-chunk][y][x][16 B], the K-major SWIZZLE_NONE "plane" layout conv_umma.cu feeds to tcgen05.mma
//   variant C: like B with element strides {1, 2, 2, 1}: every other pixel of a 33 x 17 window (one parity plane of a
//              stride-2 conv), 4 boxes per stage
// Build: nvcc -arch=sm_100a -O3 -o halo_load_tma halo_load_tma.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma4(uint32_t dst, const void* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(tm), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma5(uint32_t dst, const void* tm, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(tm), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm, int H, int W, int tiles, int tx_bytes, int stage_bytes, int nstage, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[8];
    const int tid = threadIdx.x;
    if (tid == 0) for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int tiles_x = W / 16, tiles_y = H / 32;
    auto issue = [&](int t, int s) {
        const int ty = (t / tiles_x) % tiles_y, tx = t % tiles_x, n = t / (tiles_x * tiles_y);
        const uint32_t dst = smem_u32(sm) + s * stage_bytes;
        mbar_expect(&bar[s], tx_bytes);
        if (VARIANT == 0) tma4(dst, &tm, 0, tx * 16 - 1, ty * 32 - 1, n, &bar[s]);
        else if (VARIANT == 1) tma5(dst, &tm, 0, tx * 16 - 1, ty * 32 - 1, 0, n, &bar[s]);
        else {   // 4 parity planes: (row parity, column parity) windows of 17/16 x 9/8 pixels; boxes are all 17 x 9 here
            for (int q = 0; q < 4; ++q) tma5(dst + q * (((tx_bytes / 4) + 127) / 128 * 128), &tm, 0, tx * 16 - 1 + (q & 1), ty * 32 - 1 + (q >> 1), 0, n, &bar[s]);
        }
    };
    uint32_t acc = 0;
    int it = 0;
    // prologue: fill nstage - 1 stages
    int t_issue = blockIdx.x, issued = 0;
    for (; issued < nstage - 1 && t_issue < tiles; ++issued, t_issue += gridDim.x)
        if (tid == 0) issue(t_issue, issued % nstage);
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        if (t_issue < tiles) {
            if (tid == 0) issue(t_issue, issued % nstage);
            ++issued; t_issue += gridDim.x;
        }
        const int s = it % nstage;
        mbar_wait(&bar[s], (it / nstage) & 1);
        acc += sm[s * stage_bytes + (tid * 16) % 4096];
        __syncthreads();   // everybody is done with stage s before it is refilled
    }
    if (acc == 0xffffffffu) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 8, H = 352, W = 640, C = 64;      // bf16 NHWC, 128 B per pixel
    uint8_t* in; unsigned long long* sink;
    cudaMalloc(&in, (size_t)N * H * W * C * 2);
    cudaMalloc(&sink, 8);
    cudaMemset(in, 1, (size_t)N * H * W * C * 2);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int tiles = N * (H / 32) * (W / 16);
    for (int variant = 0; variant < 3; ++variant) {
        CUtensorMap tm;
        CUresult r;
        int stage_bytes;
        if (variant == 0) {
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            cuuint32_t box[4] = {64, 17, 33, 1}, es[4] = {1, 1, 1, 1};
            r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            stage_bytes = 33 * 17 * 128;
        } else {
            // dims: (8 channels, x, y, k-chunk, n); the k-chunk dimension has a 16-byte stride
            cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C / 8, (cuuint64_t)N};
            cuuint64_t str[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, 16, (cuuint64_t)H * W * C * 2};
            cuuint32_t box[5] = {8, 17, 33, 8, 1}, es[5] = {1, 1, 1, 1, 1};
            if (variant == 2) { es[1] = 2; es[2] = 2; }
            r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, in, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            stage_bytes = variant == 1 ? 8 * 33 * 17 * 16 : 4 * (8 * 17 * 9 * 16);
        }
        if (r != CUDA_SUCCESS) { printf("variant %d: encode failed (%d)\n", variant, (int)r); continue; }
        stage_bytes = (stage_bytes + 1023) / 1024 * 1024;   // expect_tx must match what lands: see below
        for (int nstage = 2; nstage <= 3; ++nstage) {
            const int real = variant == 0 ? 33 * 17 * 128 : (variant == 1 ? 8 * 33 * 17 * 16 : 4 * 8 * 17 * 9 * 16);
            auto kern = variant == 0 ? k<0> : (variant == 1 ? k<1> : k<2>);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float ms = 0;
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(e0);
                const int pitch = (real + 2047) / 1024 * 1024;
                kern<<<148, 128, nstage * pitch + 1024>>>(tm, H, W, tiles, real, pitch, nstage, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            cudaEventElapsedTime(&ms, e0, e1);
            printf("variant %d (%s) stages=%d  %.3f ms  %.0f GB/s  (%s)\n", variant,
                   variant == 0 ? "box 128B x17x33" : (variant == 1 ? "box 16B x17x33x8kc" : "4 parity boxes 16B x17x9(x2 strides)x8kc"), nstage, ms,
                   (double)tiles * real / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
