// conv_umma.cu -- implicit-GEMM convolution on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces, for the LFD hot path, every nn.Conv2d(+BatchNorm2d)(+residual)(+ReLU) of the reference's
// backbone / neck / head towers (lfd/model/backbone/lfd_resnet.py:96-154,354-473,
// lfd/model/neck/simple_neck.py:35-47, lfd/model/head/lfd_head.py:85-135), which the reference runs
// as separate cuDNN / ATen kernels in NCHW fp32.
//
// Formulation (NHWC bf16 activations, fp32 accumulate in TMEM):
//   D[128 output pixels, Cout] = sum over taps (kh,kw) and channel chunks of  A_tap[128, 16] * W_tap[16, Cout]
//   * persistent CTAs (1 per SM for 3x3, 2 per SM for the 1x1 layers and the stem), warp-specialised:
//       warps 0..E-1  epilogue   : TMEM -> regs -> (+residual) (+ReLU) -> bf16 -> the warp's own staging rows -> the warp's
//                                  own TMA tensor store (+ optional GroupNorm partial statistics of the stored tensor);
//                                  E = 4 (one warp per TMEM lane quarter) or 8 (two warps per quarter, each half the columns)
//       warp  E       MMA issuer : one lane issues tcgen05.mma with pre-built descriptors, commits to mbarriers
//       warps E+1..   producers  : cp.async (16 B, zero-fill = conv padding) of the input halo tile into the A ring;
//                                  completion is signalled by cp.async.mbarrier.arrive (no thread waits on its own copies)
//   * BatchNorm is folded on the host: scale into the bf16 weights, shift into one extra K=16 MMA against a constant
//     "ones" operand, so the accumulator already holds scale * conv + shift.
//   * the input halo tile is loaded ONCE per (tile, channel-chunk) into "pixel planes"
//       plane[k-chunk][pixel][8 channels = 16 B]
//     which is exactly the UMMA K-major / no-swizzle canonical layout (8-row core matrices of 16 B rows,
//     SBO between 8-row groups, LBO between 16-byte K chunks).  A 3x3 tap is then just a *shifted view*
//     (start address += tap offset, SBO = halo row pitch), so the 9 taps re-read shared memory, never L2.
//     Stride-2 convolutions de-interleave the halo into 4 row/column parity planes so that every tap is
//     again a unit-stride view.  The plane pitch (LBO) is an ODD multiple of 16 B so that the 8 channel
//     chunks of one pixel land in 8 different bank groups (conflict-free cp.async writes).
//   * MODE_STEM: the 3-channel stem conv's im2col is done by the same address generator (see kStem* below).
//   * optional second GEMM in the same launch: a trailing 1x1 conv fed from shared memory ("tail": stem0->stem1,
//     stem2->stem3), or the residual block's 1x1/s2 shortcut conv on the centre tap (MODE_3X3S2, second output tensor).
//   * weights: pre-packed on the host in [channel-chunk][tap][k-chunk][Cout][8] order and brought in by
//     the TMA engine as 1-D bulk copies (cp.async.bulk -> UBLKCP), either once (resident) or per stage
//     (streamed, for 3x3x128x128 which does not fit next to the A ring).
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * programmatic dependent launch: the prologue (barriers, TMEM allocation, weight fetch) overlaps the previous layer.
#include <stdlib.h>

#include "conv_common.cuh"
#include "ptx.cuh"

namespace lfd {

static constexpr int kProdThreads = 128;
// The 3x3/s2 layers are bound by the issue latency of the halo copies (561 pixels x Cc/8 scattered 16-byte cp.async per stage, each behind a
// shared-memory table read).  Measured (profiles/r02_tuning_notes.md): 128 / 192 / 256 producer threads give 0.109 / 0.107 / 0.106 ms for the
// 64->64 @180x320 layer -- the layer is NOT issue-bound but bound by 64-byte (half-line) segment fetches; four warps stay the default.
#ifndef LFD_B200_S2_PROD
#define LFD_B200_S2_PROD 128
#endif
template <int MODE> struct ProdThreads { static constexpr int value = (MODE == MODE_3X3S2) ? LFD_B200_S2_PROD : kProdThreads; };

// The role bodies are lambdas that capture ~30 locals by reference.  If the compiler decides NOT to inline one of them (it did,
// as soon as a lambda had three call sites or a second instantiation of the template existed) the closure is materialised in
// local memory and the kernel runs 2-3x slower (a 230-byte stack frame in ptxas -v is the symptom).  Force it.
#define LFD_LAMBDA_INLINE __attribute__((always_inline))

// clock64() timeline of CTA 0 (tests/debug_trace.py); compiled in only with -DLFD_B200_TRACE (LFD_B200_TRACE=1 python build.py)
#ifdef LFD_B200_TRACE
#define LFD_TRACE(role, idx, slot) \
    do { if (p.trace && blockIdx.x == 0 && (idx) < 32) p.trace[((role) * 32 + (idx)) * 4 + (slot)] = clock64(); } while (0)
#else
#define LFD_TRACE(role, idx, slot) ((void)0)
#endif

// floor(x / d) for 0 <= x < 2^24 via one 32x32->64 multiply; m = ceil(2^40 / d), exact for d < 2^16
LFD_DEVINL int fast_div(int x, uint64_t magic) { return (int)(((uint64_t)(uint32_t)x * magic) >> 40); }

struct PxEntry {  // one halo pixel: where it comes from and where it goes
    uint32_t src_off;  // byte offset of the pixel relative to the halo's top-left pixel (tile origin + (dy_min, dx_min))
    uint32_t dst_off;  // byte offset of its 16-byte slot inside a plane
};
struct PxDelta { int8_t dy, dx; };  // the same pixel relative to the tile's input origin, for tiles that touch the border

// MODE_STEM: the im2col of the 3-channel stem conv is done by the UMMA address generator.  The producers write the raw-image
// patch of one 16x8 output tile (33 rows x 18 pixels) to shared memory as bf16 with the channels padded to 4
// ([row][pixel][b g r 0] = 8 B per pixel, already normalised / rounded = rounding point R0).  Output pixel (oy, ox) and filter row
// kh need the 3 input pixels 2ox-1 .. 2ox+1 of input row 2oy+kh-1 = 12 of the 16 consecutive bf16 values that start at patch
// pixel (2oy + kh, 2ox); consecutive ox are 16 B apart, consecutive oy 2 patch rows.  That is exactly a K-major SWIZZLE_NONE A
// operand with K = 16: core-matrix rows 16 B apart, SBO = 2 rows, LBO (second K chunk) = 16 B (the chunks of neighbouring rows
// overlap, which a read-only view may).  So the conv is 3 MMAs (one per kh, K = 16) per tile against weights packed
// [kh][2][Cout][8] with zeros in the 4th pixel / 4th channel positions; the 18th patch column only meets zero weights but
// must hold finite values.
static constexpr int kStemRows = 33, kStemCols = 18, kStemPix = kStemRows * kStemCols;          // 594
static constexpr int kStemRowBytes = kStemCols * 8;                                             // 144
static constexpr int kStemPerThread = (kStemPix + kProdThreads - 1) / kProdThreads;             // 5
static constexpr int kStemPatchBytes = ((kStemRows * kStemRowBytes + 127) / 128) * 128;         // 4864

// pitch (bytes) between the 16-byte channel chunks of the fused tail's A operand [chunk][128 rows + 1][16 B]
static constexpr uint32_t kA2Pitch = 129 * 16;

// 8 accumulator columns (+ 8 residual values) -> packed bf16; ReLU is fused into the conversion (cvt.rn.relu)
template <bool RELU, bool F16>
LFD_DEVINL uint4 pack8(const float* v) {
    uint4 o;
    if (RELU) { o.x = pack2_relu<F16>(v[0], v[1]); o.y = pack2_relu<F16>(v[2], v[3]); o.z = pack2_relu<F16>(v[4], v[5]); o.w = pack2_relu<F16>(v[6], v[7]); }
    else { o.x = pack2<F16>(v[0], v[1]); o.y = pack2<F16>(v[2], v[3]); o.z = pack2<F16>(v[4], v[5]); o.w = pack2<F16>(v[6], v[7]); }
    return o;
}
template <bool RELU, bool F16>
LFD_DEVINL uint4 pack8_res(const float* v, uint4 rv) {
    float o[8] = {v[0] + up_lo<F16>(rv.x), v[1] + up_hi<F16>(rv.x), v[2] + up_lo<F16>(rv.y), v[3] + up_hi<F16>(rv.y),
                  v[4] + up_lo<F16>(rv.z), v[5] + up_hi<F16>(rv.z), v[6] + up_lo<F16>(rv.w), v[7] + up_hi<F16>(rv.w)};
    return pack8<RELU, F16>(o);
}

// Epilogue inner loop of one warp: NC accumulator columns of its TMEM lane quarter (row = lane) -> (+residual) (+ReLU) -> bf16
// -> shared memory, fully unrolled so that every address is `base` plus / xor a literal.
//   OPERAND = false: staging row in TMA swizzle layout, chunk k (8 channels, 16 B) at (base ^ ((k & 7) << 4)) + (k >> 3) * 4096;
//                    RES adds the residual chunk found at the same place (TMA-loaded before), STATS accumulates the sum and
//                    the sum of squares of the stored values per chunk into st[k] / st[NC/8 + k] (rows with !valid count 0)
//   OPERAND = true : K-major A operand of the fused tail, chunk k at base + k * kA2Pitch
template <int NC, bool RELU, bool RES, bool STATS, bool OPERAND, int MAXB, bool F16>
LFD_DEVINL void drain(uint32_t taddr, uint32_t base, bool valid, float* st) {
    constexpr int BATCH = NC < MAXB ? NC : MAXB;   // columns in flight per tcgen05.wait::ld
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += BATCH) {
        float v[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; j += 16) tmem_ld16(taddr + c0 + j, v + j);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < BATCH / 8; ++h) {
            const int k = (c0 >> 3) + h;
            const uint32_t addr = OPERAND ? base + k * kA2Pitch : (base ^ (uint32_t)((k & 7) << 4)) + (uint32_t)(k >> 3) * 4096u;
            const uint4 o = RES ? pack8_res<RELU, F16>(v + h * 8, lds128(addr)) : pack8<RELU, F16>(v + h * 8);
            sts128(addr, o);
            if (STATS) {
                const float f[8] = {up_lo<F16>(o.x), up_hi<F16>(o.x), up_lo<F16>(o.y), up_hi<F16>(o.y), up_lo<F16>(o.z), up_hi<F16>(o.z), up_lo<F16>(o.w), up_hi<F16>(o.w)};
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
                st[k] = valid ? s1 : 0.f;
                st[NC / 8 + k] = valid ? s2 : 0.f;
            }
        }
    }
}

// Sums NV per-lane values over the 32 lanes of a warp with NV - 1 + log2(32 / NV) shuffles (transposing butterfly: every step
// halves the number of live values).  Returns, in every lane, the total of value index (lane * NV / 32).
template <int NV>
LFD_DEVINL float warp_multi_reduce(float* val, int lane) {
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? val[i] : val[i + n / 2];
            const float keep = upper ? val[i + n / 2] : val[i];
            val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = val[0];
#pragma unroll
    for (int o = 16 / NV; o >= 1; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

// GroupNorm partial statistics of one warp's NC stored channels: st = [sum per 8-channel chunk | sum of squares per chunk]
template <int NC>
LFD_DEVINL void stats_flush(float* st, int lane, double* dst) {   // dst: (sum, sumsq) pair of the warp's first group
    constexpr int NV = NC / 4, PER = 32 / NV, G = NV / 2;
    const float r = warp_multi_reduce<NV>(st, lane);
    if ((lane & (PER - 1)) == 0) {
        const int idx = lane / PER;
        atomicAdd(dst + (idx & (G - 1)) * 2 + (idx >= G ? 1 : 0), (double)r);
    }
}

template <int MODE>
__device__ __forceinline__ constexpr int tap_view(int tap) {  // pixel offset of tap's shifted view inside a plane
    if (MODE == MODE_3X3S1) return (tap / 3) * 10 + (tap % 3);
    if (MODE == MODE_3X3S2) {
        const int kh = tap / 3, kw = tap % 3;
        return (kh == 1 ? 0 : 288) + (kw == 1 ? 0 : (kh == 1 ? 144 : 153)) + (kh == 2 ? 9 : 0) + (kw == 2 ? 1 : 0);
    }
    if (MODE == MODE_STEM) return tap * (kStemRowBytes / 16);   // "tap" = filter row kh: one patch row further down
    return 0;
}

template <int MODE, int EPI_WARPS, bool F16>
__global__ void __launch_bounds__(EPI_WARPS * 32 + 32 + ProdThreads<MODE>::value, (EPI_WARPS == 4 ? 2 : 1))
conv_umma_kernel(const __grid_constant__ UmmaConvParams p) {
    constexpr int kEpiThreads = EPI_WARPS * 32;
    constexpr int kMmaWarp = EPI_WARPS;
    constexpr int kProd = ProdThreads<MODE>::value;     // producer threads of this instantiation (the stem code below assumes kProdThreads)
    constexpr int kThreads = kEpiThreads + 32 + kProd;
    constexpr int TAPS = (MODE == MODE_3X3S1 || MODE == MODE_3X3S2) ? 9 : (MODE == MODE_STEM ? 3 : 1);
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBarOff);
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint64_t* a2_full = wbar + 1;     // tail: intermediate operand written by the epilogue warps
    uint64_t* a2_empty = a2_full + 2; //       ... consumed by the tail MMAs
    uint64_t* tfull2 = a2_empty + 2;  //       tail accumulator ready
    uint64_t* tempty2 = tfull2 + 2;   //       tail accumulator drained
    uint64_t* res_bar = tempty2 + 2;  // [8] residual rows of one epilogue warp landed (TMA load)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);
    PxEntry* table = reinterpret_cast<PxEntry*>(smem + p.smem_table_off);
    PxDelta* delta = reinterpret_cast<PxDelta*>(smem + p.smem_table_off + (size_t)p.n_px * sizeof(PxEntry));
    uint8_t* staging = smem + p.smem_staging_off;
    uint8_t* wres = smem + p.smem_w_off;        // resident weights (if any)
    uint8_t* ring = smem + p.smem_ring_off;     // stages: [A chunk | B slice (streaming only)]

    // Programmatic dependent launch: let the next kernel of the stream start its prologue (barrier init, TMEM allocation,
    // weight fetch) while this one is still running; everything that touches upstream results waits at pdl_wait().
    pdl_launch_dependents();
    LFD_TL_BEGIN(p.tl);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int SA = p.stages;
    const int cpc = p.Cc >> 3;  // 16-byte chunks per pixel per stage

    // ------------------------------------------------------------------ one-time setup
    if (tid == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(&full[i], kProd + (p.b_resident ? 0 : 1));
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], EPI_WARPS);
        }
        mbar_init(wbar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a2_full[i], EPI_WARPS);
            mbar_init(&a2_empty[i], 1);
            mbar_init(&tfull2[i], 1);
            mbar_init(&tempty2[i], EPI_WARPS);
        }
        for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, p.tmem_cols);
    // The per-channel shift (folded BatchNorm / bias) is added ON THE TENSOR CORE: one extra K=16 MMA per tile of a
    // constant A operand (column 0 = 1) with a B operand whose k = 0 row holds the bf16 shift.  The epilogue is then just
    // (+residual) ReLU + convert.
    for (int i = tid; i < 256; i += kThreads)
        reinterpret_cast<uint4*>(smem + kSmemOnesOff)[i] = i < 128 ? make_uint4(F16 ? 0x00003C00u : 0x00003F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * p.Cout; i += kThreads) {
        const uint32_t b = (i < p.Cout && p.shift) ? bits16<F16>(p.shift[i]) : 0u;
        reinterpret_cast<uint4*>(smem + p.smem_bias_off)[i] = make_uint4(b, 0u, 0u, 0u);
    }
    const int cn2 = p.Cout2 + p.Cout3;      // second GEMM of the launch: fused 1x1 tail or fused 1x1/s2 shortcut (never both)
    for (int i = tid; i < 2 * cn2; i += kThreads) {
        const uint32_t b = (i < cn2 && p.shift2) ? bits16<F16>(p.shift2[i]) : 0u;
        reinterpret_cast<uint4*>(smem + p.smem_bias2_off)[i] = make_uint4(b, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();   // these operands are read by tcgen05.mma (async proxy)
    // halo pixel table (tile independent)
    if (MODE != MODE_FLAT && MODE != MODE_STEM) {
        for (int i = tid; i < p.n_px; i += kThreads) {
            int dy, dx, slot;
            if (MODE == MODE_3X3S1) {
                int r = i / 10, c = i % 10;
                dy = r - 1; dx = c - 1; slot = i;
            } else if (MODE == MODE_1X1S2) {
                int r = i >> 3, c = i & 7;
                dy = 2 * r; dx = 2 * c; slot = i;
            } else {  // MODE_3X3S2: EE(16x8) | EO(16x9) | OE(17x8) | OO(17x9); all planes use pitch 9
                int j = i, r, c, base, rodd, codd;
                if (j < 128) { r = j >> 3; c = j & 7; base = 0; rodd = 0; codd = 0; }
                else if ((j -= 128) < 144) { r = j / 9; c = j % 9; base = 144; rodd = 0; codd = 1; }
                else if ((j -= 144) < 136) { r = j >> 3; c = j & 7; base = 288; rodd = 1; codd = 0; }
                else { j -= 136; r = j / 9; c = j % 9; base = 441; rodd = 1; codd = 1; }
                dy = 2 * r - rodd; dx = 2 * c - codd; slot = base + r * 9 + c;
            }
            constexpr int kMin = (MODE == MODE_1X1S2) ? 0 : -1;   // smallest dy / dx of the mode
            PxEntry e;
            e.src_off = (uint32_t)(((dy - kMin) * p.W + (dx - kMin)) * p.Cin * 2);
            e.dst_off = (uint32_t)slot * 16u;
            table[i] = e;
            PxDelta d;
            d.dy = (int8_t)dy; d.dx = (int8_t)dx;
            delta[i] = d;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int HW = p.H * p.W;
    const int n_cc = p.Cin / p.Cc;

    if (warp < EPI_WARPS) {
        // ============================================================== EPILOGUE
        // Every warp works on its own: TMEM lane quarter (warp % 4) x a 1/WPQ slice of the columns -> bf16 rows in its OWN
        // staging region -> its own TMA tensor store (32 rows x cw channels).  No CTA-wide barrier is involved, and with two
        // staging buffers the store of tile t is still being read by the TMA engine while tile t+1 is converted.
        pdl_wait();                                   // residual reads, output stores and statistics depend on upstream kernels
        constexpr int WPQ = EPI_WARPS / 4;
        constexpr int MAXB = EPI_WARPS == 4 ? 32 : 64;   // the two-CTAs-per-SM kernels have 96 registers per thread
        const int quarter = warp & 3, chalf = warp >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;   // a warp may only touch TMEM lanes 32 * (warp % 4) ..
        const int cw = p.Cf / WPQ;                    // stored channels handled by this warp (16, 32, 64 or 128)
        const int cw1 = p.Cout / WPQ;                 // tail: channels of the intermediate handled by this warp
        const int ch0 = chalf * cw;
        // Staging rows are laid out the way the TMA engine expects for its swizzle modes: 128-byte panels
        // [panel][32 rows][128 B] with the 16-byte chunk index XORed by (row & 7) (SWIZZLE_128B); 64 / 32-byte rows use the
        // 64B / 32B patterns.  Row offsets are multiples of the row size, so "+" is "^" and chunk k of this lane's row is at
        //   (pre ^ ((k & 7) << 4)) + (k >> 3) * 4096        with a per-thread constant `pre`.
        const int row_bytes = cw >= 64 ? 128 : cw * 2;
        const uint32_t warp_stg = 64u * cw;           // bytes of one staging buffer of this warp: 32 rows x cw x 2
        const uint32_t stg0 = smem_u32(staging) + (uint32_t)warp * p.stg_nbuf * warp_stg;
        const uint32_t swz = row_bytes == 128 ? (lane & 7) : (row_bytes == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
        const uint32_t pre = (uint32_t)(lane * row_bytes) ^ (swz << 4);
        const int n_panels = cw >= 64 ? (cw >> 6) : 1;
        const int HoWo = p.Ho * p.Wo;
        const bool has_res = p.res != nullptr;
        // code variant of the conversion loop: log2(cw / 16) | relu << 2 | residual << 3, or 16 + log2(cw / 16) with statistics
        const int l2cw = cw == 16 ? 0 : (cw == 32 ? 1 : (cw == 64 ? 2 : 3));
        const int l2cw1 = cw1 == 16 ? 0 : (cw1 == 32 ? 1 : (cw1 == 64 ? 2 : 3));
        const int fin_relu = p.Cout2 ? p.relu2 : p.relu;
        const int variant = p.stats ? 16 + l2cw : (l2cw | (fin_relu ? 4 : 0) | (has_res ? 8 : 0));
        const int variant1 = l2cw1 | (p.relu ? 4 : 0);
        if (lane == 0 && (stg0 & 1023u)) __trap();    // swizzle atoms need 1024-byte aligned staging regions

        // ---- tail phase 1: main accumulator -> bf16 operand of the fused 1x1 conv (never leaves the SM)
        auto mid_tile = [&](uint32_t tc) LFD_LAMBDA_INLINE {
            const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
            const uint32_t b = p.n_a2 == 2 ? (tc & 1) : 0;
            const uint32_t use = p.n_a2 == 2 ? (tc >> 1) : tc;       // how often this operand buffer has been filled before
            mbar_wait(&tfull[a], aph);
            mbar_wait(&a2_empty[b], (use & 1) ^ 1);                   // tail MMAs of the previous user of this buffer are done
            tc_fence_after_sync();
            const uint32_t trow = tmem_base + lane_base + a * p.Cout + chalf * cw1;
            const uint32_t dst = smem_u32(smem + p.smem_a2_off) + b * p.a2_bytes + (uint32_t)(quarter * 32 + lane) * 16 + (uint32_t)((chalf * cw1) >> 3) * kA2Pitch;
            switch (variant1) {
                case 0: drain<16, false, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 1: drain<32, false, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 2: drain<64, false, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 3: drain<128, false, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 4: drain<16, true, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 5: drain<32, true, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                case 6: drain<64, true, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
                default: drain<128, true, false, false, true, MAXB, F16>(trow, dst, true, nullptr); break;
            }
            tc_fence_before_sync();
            fence_proxy_async_smem();           // st.shared (generic proxy) -> tcgen05.mma (async proxy)
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a2_full[b]);
                mbar_arrive(&tempty[a]);
            }
        };

        // ---- final phase: accumulator (+residual) (+ReLU) -> bf16 staging rows -> TMA store (+ GroupNorm statistics)
        uint32_t store_count = 0;   // staging buffers alternate per STORE (a launch with a fused shortcut stores twice per tile)
        auto finish_tile = [&](int tile, uint32_t tc, uint64_t* bar_full, uint64_t* bar_empty, uint32_t col_base, const CUtensorMap* tmap,
                               int var, bool first, bool last) LFD_LAMBDA_INLINE {
            const int n = fast_div(tile, p.magic_tpi);
            const int t = tile - n * p.tiles_per_img;
            int c1, c2 = 0;      // coordinates of this warp's first row: pixel index (flat) or (x, y)
            bool valid;          // this lane's row lies inside the feature map (statistics only; the TMA store clips)
            if (MODE == MODE_FLAT) { c1 = t * 128 + quarter * 32; valid = c1 + lane < HoWo; }
            else {
                const int ty = fast_div(t, p.magic_tx);
                c2 = ty * 16 + quarter * 4; c1 = (t - ty * p.tiles_x) * 8;
                valid = (c2 + (lane >> 3) < p.Ho) && (c1 + (lane & 7) < p.Wo);
            }
            const uint32_t a = tc & 1, aph = (tc >> 1) & 1;
            // (only the 3x3/s2 mode can store twice per tile; the other instantiations keep using the tile counter, which costs
            //  them no extra live register -- the two-CTAs-per-SM kernels sit right at their 96-register budget)
            const uint32_t sbuf = stg0 + (p.stg_nbuf == 2 ? ((MODE == MODE_3X3S2 ? store_count : tc) & 1) * warp_stg : 0u);
            if (MODE == MODE_3X3S2) ++store_count;
            if (lane == 0) {
                // this staging buffer was the source of an earlier store: the TMA engine must be done reading it
                if (p.stg_nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                if (has_res) {   // residual rows -> staging (out-of-map rows / columns arrive as zeros), added in place below
                    mbar_arrive_expect_tx(&res_bar[warp], warp_stg);
                    for (int pn = 0; pn < n_panels; ++pn) {
                        if (MODE == MODE_FLAT) tma_load_3d(sbuf + pn * 4096, &p.tm_res, ch0 + pn * 64, c1, n, &res_bar[warp]);
                        else tma_load_4d(sbuf + pn * 4096, &p.tm_res, ch0 + pn * 64, c1, c2, n, &res_bar[warp]);
                    }
                }
            }
            if (tid == 0) LFD_TRACE(2, tc, 0);
            if (first) {
                mbar_wait(&bar_full[a], aph);
                tc_fence_after_sync();
            }
            if (tid == 0) LFD_TRACE(2, tc, 1);
            if (has_res) mbar_wait(&res_bar[warp], tc & 1);
            else __syncwarp();                    // lane 0 has seen the buffer free
            const uint32_t trow = tmem_base + lane_base + col_base + a * p.Cf + ch0;
            const uint32_t base = sbuf + pre;
            auto publish = [&]() LFD_LAMBDA_INLINE {
                tc_fence_before_sync();
                fence_proxy_async_smem();             // st.shared (generic proxy) -> TMA store (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (last) mbar_arrive(&bar_empty[a]);   // accumulator stage may be overwritten by the next-but-one tile
                    for (int pn = 0; pn < n_panels; ++pn) {   // rows / columns outside the map are clipped by the TMA engine
                        if (MODE == MODE_FLAT) tma_store_3d(tmap, sbuf + pn * 4096, ch0 + pn * 64, c1, n);
                        else tma_store_4d(tmap, sbuf + pn * 4096, ch0 + pn * 64, c1, c2, n);
                    }
                    bulk_commit();
                }
            };
            // GroupNorm partial sums are taken over the STORED (bf16) values; one group = one 16-byte chunk (8 channels)
            double* sdst = p.stats ? p.stats + ((size_t)n * p.gn_groups + (ch0 >> 3)) * 2 : nullptr;
            switch (var) {
                case 0: drain<16, false, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 1: drain<32, false, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 2: drain<64, false, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 3: drain<128, false, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 4: drain<16, true, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 5: drain<32, true, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 6: drain<64, true, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 7: drain<128, true, false, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 8: drain<16, false, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 9: drain<32, false, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 10: drain<64, false, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 11: drain<128, false, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 12: drain<16, true, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 13: drain<32, true, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 14: drain<64, true, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 15: drain<128, true, true, false, false, MAXB, F16>(trow, base, valid, nullptr); break;
                case 16: { float st[4]; drain<16, false, false, true, false, MAXB, F16>(trow, base, valid, st); publish(); stats_flush<16>(st, lane, sdst); } break;
                case 17: { float st[8]; drain<32, false, false, true, false, MAXB, F16>(trow, base, valid, st); publish(); stats_flush<32>(st, lane, sdst); } break;
                case 18: { float st[16]; drain<64, false, false, true, false, MAXB, F16>(trow, base, valid, st); publish(); stats_flush<64>(st, lane, sdst); } break;
                default: { float st[32]; drain<128, false, false, true, false, MAXB, F16>(trow, base, valid, st); publish(); stats_flush<128>(st, lane, sdst); } break;
            }
            if (var < 16) publish();
            if (tid == 0) LFD_TRACE(2, tc, 2);
            if (tid == 0) LFD_TRACE(2, tc, 3);
        };

        if (!p.Cout2) {
            uint32_t tcount = 0;
            // fused shortcut (3x3/s2 only): a second pass drains the second accumulator of the same stage (same barriers) into
            // its own tensor, without ReLU.  One call site, so that the lambda stays inlined in every instantiation.
            const int n_pass = (MODE == MODE_3X3S2 && p.Cout3) ? 2 : 1;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount)
                for (int ps = 0; ps < n_pass; ++ps)
                    finish_tile(tile, tcount, tfull, tempty, ps ? 2 * p.Cout : 0, ps ? &p.tm_out3 : &p.tm_out, ps ? l2cw : variant, ps == 0,
                                ps == n_pass - 1);
        } else {
            // software pipelined: intermediate of tile t, then the finished tail of tile t-1
            uint32_t tcount = 0;
            for (int tile = blockIdx.x;; tile += gridDim.x, ++tcount) {
                const bool has = tile < p.num_tiles;
                if (has) mid_tile(tcount);
                if (tcount >= 1) finish_tile(tile - (int)gridDim.x, tcount - 1, tfull2, tempty2, 2 * p.Cout, &p.tm_out, variant, true, true);
                if (!has) break;
            }
        }
        if (lane == 0) bulk_wait_all();   // all tile stores have been performed before the CTA retires
    } else if (warp == kMmaWarp) {
        // ============================================================== MMA ISSUER
        // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers; one elected
        // lane issues the tcgen05 instructions.
        const uint32_t idesc = umma_idesc_16(128, p.Cout, F16);
        const uint32_t lbo_b = p.Cout * 16;
        // descriptors differ only in the 14-bit start-address field (bytes >> 4): pre-compute everything else
        const uint64_t adesc0 = umma_smem_desc(0, p.lbo_a, p.sbo_a);
        const uint64_t bdesc0 = umma_smem_desc(0, lbo_b, 128);
        const uint32_t a_k16 = (2 * p.lbo_a) >> 4;      // address-field step per 16 input channels (A)
        const uint32_t b_k16 = (2 * lbo_b) >> 4;        //   (B)
        const uint32_t b_tap = (cpc * lbo_b) >> 4;      // address-field step per tap (B)
        const int nk16 = p.Cc >> 4;
        const uint32_t w2_bytes = p.Cout2 ? (uint32_t)(p.Cout * p.Cout2 * 2) : (p.Cout3 ? (uint32_t)(p.Cin * p.Cout3 * 2) : 0u);
        if (p.b_resident || w2_bytes) {
            if (elect_one_sync()) {
                const uint32_t w1_bytes = p.b_resident ? p.w_total_bytes : 0u;
                mbar_arrive_expect_tx(wbar, w1_bytes + w2_bytes);
                for (uint32_t off = 0; off < w1_bytes; off += 32768) {
                    uint32_t nb = w1_bytes - off < 32768 ? w1_bytes - off : 32768;
                    bulk_g2s(smem_u32(wres) + off, reinterpret_cast<const uint8_t*>(p.w) + off, nb, wbar);
                }
                if (w2_bytes) bulk_g2s(smem_u32(smem + p.smem_w2_off), p.w2, w2_bytes, wbar);
            }
            __syncwarp();
            mbar_wait(wbar, 0);
        }
        // fused 1x1 tail: D2[128 x Cout2] = A2[128 x Cout] . W2, A2 written by the epilogue warps (mid_tile)
        const uint32_t idesc2 = umma_idesc_16(128, cn2 ? cn2 : 16, F16);
        const uint64_t ones_desc = umma_smem_desc(smem_u32(smem + kSmemOnesOff), 2048, 128);
        const uint64_t bias_desc = umma_smem_desc(smem_u32(smem + p.smem_bias_off), lbo_b, 128);
        const uint64_t bias2_desc = umma_smem_desc(smem_u32(smem + p.smem_bias2_off), cn2 * 16, 128);
        const uint64_t a2desc0 = umma_smem_desc(0, kA2Pitch, 128);
        const uint64_t b2desc0 = umma_smem_desc(smem_u32(smem + p.smem_w2_off), cn2 * 16, 128);
        auto issue_tail = [&](uint32_t u) LFD_LAMBDA_INLINE {
            const uint32_t b = p.n_a2 == 2 ? (u & 1) : 0, use = p.n_a2 == 2 ? (u >> 1) : u;
            const uint32_t a2s = u & 1, a2ph = (u >> 1) & 1;
            mbar_wait(&a2_full[b], use & 1);
            mbar_wait(&tempty2[a2s], a2ph ^ 1);
            tc_fence_after_sync();
            fence_proxy_async_smem();
            if (elect_one_sync()) {
                const uint64_t ad2 = a2desc0 + ((smem_u32(smem + p.smem_a2_off) + b * p.a2_bytes) >> 4);
                const uint32_t d2 = tmem_base + 2 * p.Cout + a2s * p.Cout2;
                for (int k16 = 0; k16 < (p.Cout >> 4); ++k16)
                    umma_bf16(d2, ad2 + (uint32_t)(k16 * ((2 * kA2Pitch) >> 4)), b2desc0 + (uint32_t)(k16 * ((2 * p.Cout2 * 16) >> 4)), idesc2, k16 != 0);
                if (p.shift2) umma_bf16(d2, ones_desc, bias2_desc, idesc2, 1);
                umma_commit(&tfull2[a2s]);
                umma_commit(&a2_empty[b]);
            }
            __syncwarp();
        };
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
            if (lane == 0) LFD_TRACE(1, tcount, 0);
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after_sync();
            if (lane == 0) LFD_TRACE(1, tcount, 1);
            const uint32_t d_tmem = tmem_base + a * p.Cout;
            for (int cc = 0; cc < n_cc; ++cc, ++it) {
                const uint32_t s = it % SA, ph = (it / SA) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                fence_proxy_async_smem();   // cp.async (generic proxy) writes -> tcgen05.mma (async proxy) reads
                if (cc == 0 && lane == 0) LFD_TRACE(1, tcount, 2);
                const uint32_t a_base = smem_u32(ring) + s * p.stage_bytes;
                const uint32_t b_base = p.b_resident ? smem_u32(wres) + cc * p.b_slice_bytes : a_base + p.a_stage_bytes;
                const uint64_t ad = adesc0 + (a_base >> 4), bd = bdesc0 + (b_base >> 4);
                if (elect_one_sync()) {
                    for (int k16 = 0; k16 < nk16; ++k16) {
                        const uint64_t adk = ad + (uint32_t)(k16 * a_k16), bdk = bd + (uint32_t)(k16 * b_k16);
#pragma unroll
                        for (int tap = 0; tap < TAPS; ++tap)
                            umma_bf16(d_tmem, adk + (uint32_t)tap_view<MODE>(tap), bdk + (uint32_t)(tap * b_tap), idesc, (cc | k16 | tap) != 0);
                    }
                    if (MODE == MODE_3X3S2 && p.Cout3) {
                        // fused 1x1/s2 shortcut conv of the residual block: its input pixel is this conv's centre tap, so it
                        // is one more MMA per 16 channels on the operand that is already in shared memory
                        const uint32_t d3 = tmem_base + 2 * p.Cout + a * p.Cout3;
                        for (int k16 = 0; k16 < nk16; ++k16)
                            umma_bf16(d3, ad + (uint32_t)(k16 * a_k16) + (uint32_t)tap_view<MODE>(4),
                                      b2desc0 + (uint32_t)(((cc * cpc + 2 * k16) * p.Cout3 * 16) >> 4), idesc2, (cc | k16) != 0);
                        if (cc == n_cc - 1 && p.shift2) umma_bf16(d3, ones_desc, bias2_desc, idesc2, 1);
                    }
                    umma_commit(&empty[s]);
                    if (cc == n_cc - 1) {
                        if (p.shift) umma_bf16(d_tmem, ones_desc, bias_desc, idesc, 1);
                        umma_commit(&tfull[a]);
                    }
                }
                __syncwarp();
                if (cc == n_cc - 1 && lane == 0) LFD_TRACE(1, tcount, 3);
            }
            if (p.Cout2 && tcount >= 1) issue_tail(tcount - 1);
        }
        if (p.Cout2 && tcount >= 1) issue_tail(tcount - 1);
    } else {
        // ============================================================== PRODUCERS
        const int ptid = tid - (kEpiThreads + 32);
        if (MODE == MODE_STEM) {
            // Raw image patch -> normalised bf16 [row][pixel][b g r 0] in the ring stage (see kStem* above).  Every thread owns
            // up to 5 patch pixels; the raw values of the NEXT tile are fetched into registers right after the current patch
            // has been written, so the loads fly while the thread waits for the next free stage.
            const bool u8 = p.input_format == 1;
            const int plane = p.H * p.W;
            int rel[kStemPerThread];          // source offset of pixel j relative to the patch origin (bytes for u8, elements for fp32)
            int prc[kStemPerThread];          // (row << 8) | col
#pragma unroll
            for (int j = 0; j < kStemPerThread; ++j) {
                const int q = min(ptid + j * kProdThreads, kStemPix - 1);
                const int r = q / kStemCols, c = q - r * kStemCols;
                prc[j] = (r << 8) | c;
                rel[j] = u8 ? (r * p.W + c) * 3 : r * p.W + c;
            }
            const bool last_ok = ptid + (kStemPerThread - 1) * kProdThreads < kStemPix;   // this thread owns a pixel in the last round
            uint32_t raw[kStemPerThread][3];
            uint32_t okmask = 0;
            auto fetch = [&](int tile) LFD_LAMBDA_INLINE {
                const int n = fast_div(tile, p.magic_tpi), t = tile - n * p.tiles_per_img;
                const int ty = fast_div(t, p.magic_tx);
                const int iy0 = 2 * ty * 16 - 1, ix0 = 2 * (t - ty * p.tiles_x) * 8 - 1;
                const bool interior = iy0 >= 0 && ix0 >= 0 && iy0 + kStemRows <= p.H && ix0 + kStemCols <= p.W;
                okmask = 0;
                if (u8) {
                    const uint8_t* img = reinterpret_cast<const uint8_t*>(p.in_raw) + (size_t)n * plane * 3;
                    const uint8_t* org = img + ((ptrdiff_t)iy0 * p.W + ix0) * 3;
#pragma unroll
                    for (int j = 0; j < kStemPerThread; ++j) {
                        if (j == kStemPerThread - 1 && !last_ok) break;
                        const uint8_t* src = org + rel[j];
                        bool ok = true;
                        if (!interior) {
                            const int y = iy0 + (prc[j] >> 8), x = ix0 + (prc[j] & 255);
                            ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
                            if (!ok) src = img;
                        }
                        okmask |= (ok ? 1u : 0u) << j;
                        raw[j][0] = __ldg(src); raw[j][1] = __ldg(src + 1); raw[j][2] = __ldg(src + 2);
                    }
                } else {
                    const float* img = reinterpret_cast<const float*>(p.in_raw) + (size_t)n * plane * 3;
                    const float* org = img + (ptrdiff_t)iy0 * p.W + ix0;
#pragma unroll
                    for (int j = 0; j < kStemPerThread; ++j) {
                        if (j == kStemPerThread - 1 && !last_ok) break;
                        const float* src = org + rel[j];
                        bool ok = true;
                        if (!interior) {
                            const int y = iy0 + (prc[j] >> 8), x = ix0 + (prc[j] & 255);
                            ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
                            if (!ok) src = img;
                        }
                        okmask |= (ok ? 1u : 0u) << j;
                        raw[j][0] = __float_as_uint(__ldg(src)); raw[j][1] = __float_as_uint(__ldg(src + plane));
                        raw[j][2] = __float_as_uint(__ldg(src + 2 * plane));
                    }
                }
            };
            uint32_t it = 0;
            pdl_wait();
            if ((int)blockIdx.x < p.num_tiles) fetch(blockIdx.x);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % SA, ph = (it / SA) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const uint32_t dst0 = smem_u32(ring) + s * p.stage_bytes + ptid * 8;
#pragma unroll
                for (int j = 0; j < kStemPerThread; ++j) {
                    if (j == kStemPerThread - 1 && !last_ok) break;
                    float f[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        f[k] = u8 ? ((float)raw[j][k] - 127.5f) * (1.0f / 127.5f) : __uint_as_float(raw[j][k]);
                        if (!((okmask >> j) & 1u)) f[k] = 0.f;        // conv zero padding (of the normalised image)
                    }
                    const uint32_t lo = pack2<F16>(f[0], f[1]), hi = pack2<F16>(f[2], 0.f);   // rounding point R0
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst0 + j * (kProdThreads * 8)), "r"(lo), "r"(hi) : "memory");
                }
                fence_proxy_async_smem();       // generic-proxy st.shared -> tcgen05.mma reads
                mbar_arrive(&full[s]);
                if (tile + (int)gridDim.x < p.num_tiles) fetch(tile + gridDim.x);
            }
        } else {
        // every thread owns ONE 16-byte channel chunk (cpc divides 128) and walks the halo pixels with a fixed stride
        const int ch = ptid & (cpc - 1);
        const int px0 = ptid >> p.log2_cpc;
        const int pstep = kProd >> p.log2_cpc;
        const uint32_t ch_dst = ch * p.lbo_a;
        const int my_cnt = (p.n_px - px0 + pstep - 1) / pstep;   // halo pixels this thread copies per stage
        uint32_t it = 0;
        pdl_wait();                                   // the input activations are produced by the previous kernel
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n = fast_div(tile, p.magic_tpi);
            const int t = tile - n * p.tiles_per_img;
            int iy0 = 0, ix0 = 0;
            if (MODE == MODE_FLAT) ix0 = t * 128;
            else {
                const int ty = fast_div(t, p.magic_tx);
                int oy0 = ty * 16, ox0 = (t - ty * p.tiles_x) * 8;
                iy0 = (MODE == MODE_3X3S1) ? oy0 : 2 * oy0;
                ix0 = (MODE == MODE_3X3S1) ? ox0 : 2 * ox0;
            }
            const __nv_bfloat16* img = p.in + (size_t)n * HW * p.Cin + ch * 8;
            // tiles whose halo lies inside the image (the vast majority) copy without bounds checks: source = tile origin +
            // a per-pixel offset from the table
            constexpr int kDyMin = (MODE == MODE_1X1S2) ? 0 : -1, kDxMin = kDyMin;
            constexpr int kDyMax = (MODE == MODE_3X3S1) ? 16 : ((MODE == MODE_3X3S2) ? 31 : 30);
            constexpr int kDxMax = (MODE == MODE_3X3S1) ? 8 : ((MODE == MODE_3X3S2) ? 15 : 14);
            const bool interior = MODE == MODE_FLAT ? (ix0 + 128 <= HW)
                                                    : (iy0 + kDyMin >= 0 && ix0 + kDxMin >= 0 && iy0 + kDyMax < p.H && ix0 + kDxMax < p.W);
            // flat: first pixel of the tile; otherwise the halo's top-left pixel (in range for interior tiles)
            const __nv_bfloat16* org = img + (ptrdiff_t)(MODE == MODE_FLAT ? ix0 : (iy0 + kDyMin) * p.W + ix0 + kDxMin) * p.Cin;
            for (int cc = 0; cc < n_cc; ++cc, ++it) {
                const uint32_t s = it % SA, ph = (it / SA) & 1;
                if (ptid == 0) LFD_TRACE(0, it, 0);
                mbar_wait(&empty[s], ph ^ 1);
                if (ptid == 0) LFD_TRACE(0, it, 1);
                const uint32_t a_base = smem_u32(ring) + s * p.stage_bytes;
                if (!p.b_resident && ptid == 0) {
                    mbar_arrive_expect_tx(&full[s], p.b_slice_bytes);
                    bulk_g2s(a_base + p.a_stage_bytes, reinterpret_cast<const uint8_t*>(p.w) + (size_t)cc * p.b_slice_bytes,
                             p.b_slice_bytes, &full[s]);
                }
                const __nv_bfloat16* src_cc = img + cc * p.Cc;
                const uint32_t dst_cc = a_base + ch_dst;
                if (MODE == MODE_FLAT) {
                    if (interior) {
                        const __nv_bfloat16* src = org + cc * p.Cc + (size_t)px0 * p.Cin;
                        const size_t sstep = (size_t)pstep * p.Cin;
                        uint32_t dst = dst_cc + px0 * 16;
#pragma unroll 4
                        for (int pxi = px0; pxi < 128; pxi += pstep, src += sstep, dst += pstep * 16) cp_async16_full(dst, src);
                    } else {
#pragma unroll 4
                        for (int pxi = px0; pxi < 128; pxi += pstep) {
                            const int q = ix0 + pxi;
                            const bool ok = q < HW;
                            cp_async16(dst_cc + pxi * 16, src_cc + (ok ? q : 0) * p.Cin, ok);
                        }
                    }
                } else if (interior) {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(org + cc * p.Cc);
                    const PxEntry* tp = table + px0;
#pragma unroll 4
                    for (int k = 0; k < my_cnt; ++k, tp += pstep) {
                        const uint2 pe = *reinterpret_cast<const uint2*>(tp);
                        cp_async16_full(dst_cc + pe.y, src + pe.x);
                    }
                } else {
#pragma unroll 2
                    for (int pxi = px0; pxi < p.n_px; pxi += pstep) {
                        const PxDelta pd = delta[pxi];
                        const int y = iy0 + pd.dy, x = ix0 + pd.dx;
                        const bool ok = ((unsigned)y < (unsigned)p.H) && ((unsigned)x < (unsigned)p.W);
                        cp_async16(dst_cc + table[pxi].dst_off, src_cc + (ok ? (y * p.W + x) : 0) * p.Cin, ok);
                    }
                }
                cp_async_mbar_arrive(&full[s]);   // arrives when this thread's copies have landed
                if (ptid == 0) LFD_TRACE(0, it, 2);
            }
        }
        cp_async_wait_all();
        }
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    LFD_TL_END(p.tl);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int epi_warps_of(int mode) { return (mode == MODE_3X3S1 || mode == MODE_3X3S2) ? 8 : 4; }

static int configure_with(const ConvGeom& g, int num_sms, int nbuf, UmmaConvParams* out, size_t* smem_bytes, int* grid) {
    UmmaConvParams p;
    memset(&p, 0, sizeof(p));
    p.stg_nbuf = nbuf;
    int mode;
    if (g.stem) {
        if (g.ksize != 3 || g.stride != 2 || g.Cin != 16) return -1;   // "Cin" = K of one filter row: 4 pixels x 4 padded channels
        mode = MODE_STEM;
    } else if (g.ksize == 1 && g.stride == 1) mode = MODE_FLAT;
    else if (g.ksize == 3 && g.stride == 1) mode = MODE_3X3S1;
    else if (g.ksize == 3 && g.stride == 2) mode = MODE_3X3S2;
    else if (g.ksize == 1 && g.stride == 2) mode = MODE_1X1S2;
    else return -1;
    if (g.Cin % 16 || g.Cout % 16 || g.Cout > 128 || g.Cout < 16) return -1;
    if (epi_warps_of(mode) == 8 && g.Cout < 32) return -1;   // two epilogue warps per lane quarter need >= 16 columns each
    p.mode = mode;
    p.N = g.N; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.Ho = g.Ho; p.Wo = g.Wo; p.Cout = g.Cout;
    const int taps = g.stem ? 3 : g.ksize * g.ksize;   // stem: one MMA group per filter row
    int px_slots;  // plane size in pixels (slots), n_px = pixels actually loaded
    if (mode == MODE_FLAT) {
        p.tiles_x = 0; p.tiles_per_img = (g.Ho * g.Wo + 127) / 128;
        p.n_px = 128; px_slots = 128; p.sbo_a = 128;
    } else {
        p.tiles_x = (g.Wo + 7) / 8;
        p.tiles_per_img = p.tiles_x * ((g.Ho + 15) / 16);
        if (mode == MODE_3X3S1) { p.n_px = 180; px_slots = 180; p.sbo_a = 160; }
        else if (mode == MODE_3X3S2) { p.n_px = 561; px_slots = 594; p.sbo_a = 144; }
        else { p.n_px = 128; px_slots = 128; p.sbo_a = 128; }   // MODE_1X1S2 (and, overridden below, MODE_STEM)
    }
    // plane pitch = odd multiple of 16 B: the 8 chunks of a pixel fall into 8 distinct 16-byte bank groups
    const int plane_slots = px_slots | 1;
    p.lbo_a = plane_slots * 16;
    if (mode == MODE_STEM) { p.lbo_a = 16; p.sbo_a = 2 * kStemRowBytes; }   // overlapping view of the 4-channel patch
    p.num_tiles = p.tiles_per_img * g.N;
    if (p.num_tiles >= (1 << 24) || p.tiles_per_img >= (1 << 16)) return -5;
    p.magic_tpi = ((1ull << 40) + p.tiles_per_img - 1) / p.tiles_per_img;
    p.magic_tx = p.tiles_x ? ((1ull << 40) + p.tiles_x - 1) / p.tiles_x : 0;
    const int Cf = g.tail_cout > 0 ? g.tail_cout : g.Cout;
    if (g.ds_cout && (mode != MODE_3X3S2 || g.tail_cout || g.ds_cout != g.Cout)) return -6;
    if (g.tail_cout) {
        if (g.tail_cout % 16 || g.tail_cout > 128 || g.tail_cout < 16 || 2 * (g.Cout + g.tail_cout) > 512) return -4;
        if (epi_warps_of(mode) == 8 && g.tail_cout < 32) return -4;
    }
    const size_t staging = (size_t)nbuf * 128 * Cf * 2;   // [epilogue warp][buffer][32 rows][Cf / (warps per quarter)]
    // fixed head of the shared-memory map: barriers | ones operand | [halo table] | bias | [tail bias] | staging (1 KB aligned)
    size_t hoff = kSmemOnesOff + 4096;
    p.smem_table_off = (uint32_t)hoff;
    if (mode != MODE_FLAT && mode != MODE_STEM) hoff += ((size_t)p.n_px * 10 + 127) & ~(size_t)127;   // PxEntry[n_px] | PxDelta[n_px]
    p.smem_bias_off = (uint32_t)hoff; hoff += (size_t)g.Cout * 32;
    p.smem_bias2_off = (uint32_t)hoff; hoff += (size_t)(g.tail_cout + g.ds_cout) * 32;
    p.smem_staging_off = (uint32_t)((hoff + 1023) & ~(size_t)1023);
    const size_t fixed = p.smem_staging_off + staging;
    // everything that lives behind the ring: fused-tail weights + two operand buffers
    const size_t a2_bytes = g.tail_cout ? ((size_t)(g.Cout / 8) * 129 * 16 + 127) & ~(size_t)127 : 0;
    const size_t w2_bytes = g.tail_cout ? ((size_t)g.Cout * g.tail_cout * 2 + 127) & ~(size_t)127
                                        : (g.ds_cout ? ((size_t)g.Cin * g.ds_cout * 2 + 127) & ~(size_t)127 : 0);
    const size_t post = w2_bytes + 2 * a2_bytes;
    const size_t budget = 224 * 1024 - post;
    const size_t w_total = (size_t)taps * g.Cin * g.Cout * 2;
    // Choose the channel chunk Cc, weight residency and ring depth.  Preference order:
    //   1. resident weights (loaded once per CTA) with >= 3 A stages, the largest Cc first;
    //   2. otherwise streamed weights: the largest Cc that still gives >= 4 stages (deep ring hides the per-stage
    //      weight fetch), then >= 3, then >= 2.
    // Memory-bound 1x1 layers additionally cap the ring so that two CTAs fit on one SM (latency hiding).
    const int cands[3] = {64, 32, 16};
    int best_cc = 0, best_res = 0, best_st = 0;
    auto a_bytes = [&](int cc) -> size_t {
        if (mode == MODE_STEM) return kStemPatchBytes;
        return (((size_t)plane_slots * (cc / 8) * 16) + 127) & ~(size_t)127;
    };
    auto stages_for = [&](int cc, int resident) -> int {
        if (g.Cin % cc) return 0;
        const size_t b_slice = (size_t)taps * cc * g.Cout * 2;
        const size_t stage = a_bytes(cc) + (resident ? 0 : b_slice);
        if (fixed + (resident ? w_total : 0) + 2 * stage > budget) return 0;
        int st = (int)((budget - fixed - (resident ? w_total : 0)) / stage);
        if (st > kMaxStages) st = kMaxStages;
        if (st > 4 && stage >= 8192) st = 4;
        return st;
    };
    // (16-channel chunks mean 32-byte global segments per pixel and twice the barrier round trips per tile: measured
    //  1600 cycles per 19 KB stage for the 3x3/s2 stem conv, so a 2-deep ring of 32-channel stages beats a 4-deep one of 16)
    static const int force_cc = getenv("LFD_B200_FORCE_CC") ? atoi(getenv("LFD_B200_FORCE_CC")) : 0;   // experiments only
    if (force_cc && mode == MODE_3X3S2) {
        int st = stages_for(force_cc, 1);
        if (st >= 2) { best_cc = force_cc; best_res = 1; best_st = st; }
    }
    for (int pass = 0; pass < 3 && !best_cc; ++pass)
        for (int ci = 0; ci < 3 && !best_cc; ++ci) {
            const int want = pass == 0 ? 3 : 2;
            if (pass < 2 && cands[ci] < 32) continue;
            int st = stages_for(cands[ci], 1);
            if (st >= want) { best_cc = cands[ci]; best_res = 1; best_st = st; }
        }
    // Streaming re-reads the whole filter bank from L2 for every tile, so it is taken only when the weights cannot be
    // resident next to a ring of >= 2 stages of >= 32 channels (measured on the 3x3/s2 64->64 + tail layer at 180x320:
    // resident / 32 / 2 stages beats streamed / 16 / 4 stages).
    if (!best_cc || (best_st < 3 && best_cc < 32)) {
        for (int want = 4; want >= 2; --want) {
            bool found = false;
            for (int ci = 0; ci < 3 && !found; ++ci) {
                int st = stages_for(cands[ci], 0);
                if (st >= want && (!best_cc || st > best_st)) { best_cc = cands[ci]; best_res = 0; best_st = st; found = true; }
            }
            if (found) break;
        }
    }
    if (!best_cc) return -2;
    p.Cc = best_cc; p.b_resident = best_res; p.stages = best_st;
    p.a_stage_bytes = (uint32_t)a_bytes(best_cc);
    p.b_slice_bytes = (uint32_t)((size_t)taps * best_cc * g.Cout * 2);
    p.stage_bytes = p.a_stage_bytes + (best_res ? 0 : p.b_slice_bytes);
    p.w_total_bytes = (uint32_t)w_total;
    p.smem_w_off = (uint32_t)fixed;
    p.smem_ring_off = (uint32_t)(fixed + (best_res ? w_total : 0));
    // two CTAs per SM when a >= 2-deep ring fits in half of the shared memory (1x1 / stem layers; 4 epilogue warps each)
    p.ctas_per_sm = 1;
    if (epi_warps_of(mode) == 4) {
        const size_t half = (227 * 1024) / 2 - 1024;
        const size_t base = p.smem_ring_off + post;
        if (base + 2 * (size_t)p.stage_bytes <= half) {
            int st = (int)((half - base) / p.stage_bytes);
            if (st > p.stages) st = p.stages;
            if (st >= 2) { p.stages = st; p.ctas_per_sm = 2; }
        }
    }
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    p.log2_cpc = ilog2(p.Cc / 8);
    p.Cf = Cf;
    p.Cout2 = g.tail_cout;
    p.Cout3 = g.ds_cout;
    p.log2_cpr = ilog2(Cf / 8);
    if ((1 << p.log2_cpr) != Cf / 8) return -3;       // stored channel count must be 16/32/64/128
    p.log2_rp128 = Cf * 2 >= 128 ? 0 : ilog2(128 / (Cf * 2));
    {   // TMEM: two accumulator stages of the conv (+ two of the tail); power of two >= 32
        int need = 2 * g.Cout + 2 * (g.tail_cout + g.ds_cout), cols = 32;
        while (cols < need) cols <<= 1;
        p.tmem_cols = cols;
    }
    size_t off = p.smem_ring_off + (size_t)p.stages * p.stage_bytes;
    if (g.tail_cout) {
        p.smem_w2_off = (uint32_t)off; off += w2_bytes;
        p.smem_a2_off = (uint32_t)off; off += 2 * a2_bytes;
        p.a2_bytes = (uint32_t)a2_bytes; p.n_a2 = 2;
    } else if (g.ds_cout) {
        p.smem_w2_off = (uint32_t)off; off += w2_bytes;
    }
    *smem_bytes = off;
    if (p.ctas_per_sm == 2 && 2 * p.tmem_cols > 512) p.ctas_per_sm = 1;
    const int max_ctas = num_sms * p.ctas_per_sm;
    *grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
    *out = p;
    return 0;
}

int umma_conv_configure(const ConvGeom& g, int num_sms, UmmaConvParams* out, size_t* smem_bytes, int* grid) {
    // a second staging buffer per epilogue warp is taken when it costs neither occupancy, weight residency nor ring depth
    UmmaConvParams p1, p2;
    size_t s1 = 0, s2 = 0;
    int g1 = 0, g2 = 0;
    const int rc = configure_with(g, num_sms, 1, &p1, &s1, &g1);
    if (rc) return rc;
    if (configure_with(g, num_sms, 2, &p2, &s2, &g2) == 0 && p2.ctas_per_sm == p1.ctas_per_sm && p2.b_resident == p1.b_resident &&
        p2.Cc == p1.Cc && p2.stages >= (p1.stages < 3 ? p1.stages : 3)) {
        *out = p2; *smem_bytes = s2; *grid = g2;
    } else {
        *out = p1; *smem_bytes = s1; *grid = g1;
    }
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// One tensor map per stored tensor; the box is what ONE epilogue warp moves: 32 tile rows (32 pixels of a flat tile, 8 x 4 of
// a spatial one) x min(64, its channel slice) channels, shared-memory side in the matching swizzle mode.
static int encode_one(const UmmaConvParams& p, const void* ptr, CUtensorMap* tm) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return -1;
    const cuuint64_t Cf = p.Cf, HoWo = (cuuint64_t)p.Ho * p.Wo;
    const int cw = p.Cf / (epi_warps_of(p.mode) / 4);
    const cuuint32_t inner = cw < 64 ? cw : 64;
    const CUtensorMapSwizzle swz = inner * 2 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
    cuuint32_t rank;
    if (p.mode == MODE_FLAT) {
        rank = 3;
        dims[0] = Cf; dims[1] = HoWo; dims[2] = p.N;
        strides[0] = Cf * 2; strides[1] = HoWo * Cf * 2;
        box[0] = inner; box[1] = 32; box[2] = 1;
    } else {
        rank = 4;
        dims[0] = Cf; dims[1] = p.Wo; dims[2] = p.Ho; dims[3] = p.N;
        strides[0] = Cf * 2; strides[1] = (cuuint64_t)p.Wo * Cf * 2; strides[2] = HoWo * Cf * 2;
        box[0] = inner; box[1] = 8; box[2] = 4; box[3] = 1;
    }
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

int umma_conv_encode_maps(UmmaConvParams* p) {
    if (encode_one(*p, p->out, &p->tm_out)) return -1;
    if (p->res && encode_one(*p, p->res, &p->tm_res)) return -1;
    if (p->Cout3 && encode_one(*p, p->out3, &p->tm_out3)) return -1;
    return 0;
}

template <int MODE, int EPI_WARPS, bool F16>
static cudaError_t launch_mode_t(const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: one flag per device ordinal
    static bool configured[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<MODE, EPI_WARPS, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    static const bool use_pdl = getenv("LFD_B200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(EPI_WARPS * 32 + 32 + ProdThreads<MODE>::value);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, conv_umma_kernel<MODE, EPI_WARPS, F16>, p);
}

template <int MODE, int EPI_WARPS>
static cudaError_t launch_mode(const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
    return p.f16 ? launch_mode_t<MODE, EPI_WARPS, true>(p, smem, grid, st) : launch_mode_t<MODE, EPI_WARPS, false>(p, smem, grid, st);
}

cudaError_t umma_conv_launch(const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
    switch (p.mode) {
        case MODE_FLAT: return launch_mode<MODE_FLAT, 4>(p, smem, grid, st);
        case MODE_3X3S1: return launch_mode<MODE_3X3S1, 8>(p, smem, grid, st);
        case MODE_3X3S2: return launch_mode<MODE_3X3S2, 8>(p, smem, grid, st);
        case MODE_1X1S2: return launch_mode<MODE_1X1S2, 4>(p, smem, grid, st);
        case MODE_STEM: return launch_mode<MODE_STEM, 4>(p, smem, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace lfd
