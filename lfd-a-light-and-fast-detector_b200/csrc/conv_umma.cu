// conv_umma.cu -- implicit-GEMM convolution on 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces, for the LFD hot path, every nn.Conv2d(+BatchNorm2d)(+residual)(+ReLU) of the reference's
// backbone / neck / head towers (lfd/model/backbone/lfd_resnet.py:96-154,354-473,
// lfd/model/neck/simple_neck.py:35-47, lfd/model/head/lfd_head.py:85-135), which the reference runs
// as separate cuDNN / ATen kernels in NCHW fp32.
//
// Formulation (NHWC bf16 activations, fp32 accumulate in TMEM):
//   D[128 output pixels, Cout] = sum over taps (kh,kw) and channel chunks of  A_tap[128, 16] * W_tap[16, Cout]
//   * one CTA per SM, persistent over output tiles, warp-specialised:
//       warps 0-3  epilogue   : TMEM -> regs -> scale/shift (+residual) (+ReLU) -> bf16 -> smem -> coalesced store
//                               (+ optional GroupNorm partial statistics of the stored tensor)
//       warp  4    MMA issuer : one elected lane issues tcgen05.mma, commits to mbarriers
//       warps 5-8  producers  : cp.async (16 B, zero-fill = conv padding) of the input halo tile into the A ring
//   * the input halo tile is loaded ONCE per (tile, channel-chunk) into "pixel planes"
//       plane[k-chunk][pixel][8 channels = 16 B]
//     which is exactly the UMMA K-major / no-swizzle canonical layout (8-row core matrices of 16 B rows,
//     SBO between 8-row groups, LBO between 16-byte K chunks).  A 3x3 tap is then just a *shifted view*
//     (start address += tap offset, SBO = halo row pitch), so the 9 taps re-read shared memory, never L2.
//     Stride-2 convolutions de-interleave the halo into 4 row/column parity planes so that every tap is
//     again a unit-stride view.
//   * weights: pre-packed on the host in [channel-chunk][tap][k-chunk][Cout][8] order and brought in by
//     the TMA engine as 1-D bulk copies (cp.async.bulk -> UBLKCP), either once (resident) or per stage
//     (streamed, for 3x3x128x128 which does not fit next to the A ring).
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv_common.cuh"
#include "ptx.cuh"

namespace lfd {

static constexpr int kEpiThreads = 128;
static constexpr int kMmaWarp = 4;
static constexpr int kProdThreads = 128;
static constexpr int kThreads = kEpiThreads + 32 + kProdThreads;  // 288
static constexpr int kLag = 2;                                     // cp.async groups kept in flight per producer thread

#define LFD_TRACE(role, idx, slot) \
    do { if (p.trace && blockIdx.x == 0 && (idx) < 32) p.trace[((role) * 32 + (idx)) * 4 + (slot)] = clock64(); } while (0)

struct PxEntry {  // one halo pixel: where it comes from (relative to the tile's input origin) and where it goes
    int16_t dy, dx;
    uint16_t slot;
    uint16_t pad;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) conv_umma_kernel(const UmmaConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmemBarOff);
    uint64_t* empty = full + kMaxStages;
    uint64_t* tfull = empty + kMaxStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* s_scale = reinterpret_cast<float*>(smem + kSmemScaleOff);
    float* s_shift = s_scale + 128;
    PxEntry* table = reinterpret_cast<PxEntry*>(smem + kSmemTableOff);
    uint8_t* staging = smem + kSmemStagingOff;
    uint8_t* wres = smem + p.smem_w_off;        // resident weights (if any)
    uint8_t* ring = smem + p.smem_ring_off;     // stages: [A chunk | B slice (streaming only)]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int SA = p.stages;
    const int cpc = p.Cc >> 3;  // 16-byte chunks per pixel per stage

    // ------------------------------------------------------------------ one-time setup
    if (tid == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(&full[i], kProdThreads + (p.b_resident ? 0 : 1));
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiThreads);
        }
        mbar_init(wbar, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, p.tmem_cols);
    for (int c = tid; c < p.Cout; c += kThreads) {
        s_scale[c] = p.scale[c];
        s_shift[c] = p.shift[c];
    }
    // halo pixel table (tile independent)
    for (int i = tid; i < p.n_px; i += kThreads) {
        PxEntry e;
        e.pad = 0;
        if (MODE == MODE_FLAT) {
            e.dy = 0; e.dx = (int16_t)i; e.slot = (uint16_t)i;
        } else if (MODE == MODE_3X3S1) {
            int r = i / 10, c = i % 10;
            e.dy = (int16_t)(r - 1); e.dx = (int16_t)(c - 1); e.slot = (uint16_t)i;
        } else if (MODE == MODE_1X1S2) {
            int r = i >> 3, c = i & 7;
            e.dy = (int16_t)(2 * r); e.dx = (int16_t)(2 * c); e.slot = (uint16_t)i;
        } else {  // MODE_3X3S2: EE(16x8) | EO(16x9) | OE(17x8) | OO(17x9); all planes use pitch 9
            int j = i, r, c, base, rodd, codd;
            if (j < 128) { r = j >> 3; c = j & 7; base = 0; rodd = 0; codd = 0; }
            else if ((j -= 128) < 144) { r = j / 9; c = j % 9; base = 144; rodd = 0; codd = 1; }
            else if ((j -= 144) < 136) { r = j >> 3; c = j & 7; base = 288; rodd = 1; codd = 0; }
            else { j -= 136; r = j / 9; c = j % 9; base = 441; rodd = 1; codd = 1; }
            e.dy = (int16_t)(2 * r - rodd); e.dx = (int16_t)(2 * c - codd); e.slot = (uint16_t)(base + r * 9 + c);
        }
        table[i] = e;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int HW = p.H * p.W;
    const int n_cc = p.Cin / p.Cc;
    const int taps = (MODE == MODE_3X3S1 || MODE == MODE_3X3S2) ? 9 : 1;

    if (warp < 4) {
        // ============================================================== EPILOGUE
        const int m = tid;  // D row == TMEM lane
        const int cpr = p.Cout >> 3;          // 16 B chunks per staged row (power of two)
        const int l2cpr = p.log2_cpr;
        const int row_bytes = p.Cout * 2;
        const int l2rp = p.log2_rp128;         // log2(rows per 128 B): swizzle granularity for rows shorter than 128 B
        const int swz_mask = (cpr < 8 ? cpr : 8) - 1;
        const int HoWo = p.Ho * p.Wo;
        const uint32_t stg = smem_u32(staging);
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
            const int n = tile / p.tiles_per_img;
            const int t = tile - n * p.tiles_per_img;
            int oy0 = 0, ox0 = 0, p0 = 0;
            if (MODE == MODE_FLAT) p0 = t * 128;
            else { oy0 = (t / p.tiles_x) * 16; ox0 = (t % p.tiles_x) * 8; }
            const size_t img_out = (size_t)n * p.Ho * p.Wo;
            // pixel index (within the image) of staged row r, or -1 when outside the feature map
            auto row_pixel = [&](int r) -> int {
                if (MODE == MODE_FLAT) { int q = p0 + r; return q < HoWo ? q : -1; }
                int y = oy0 + (r >> 3), x = ox0 + (r & 7);
                return (y < p.Ho && x < p.Wo) ? y * p.Wo + x : -1;
            };
            const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
            if (p.res) {  // residual tile -> staging (coalesced), consumed row-wise below
                for (int e = tid; e < 128 * cpr; e += kEpiThreads) {
                    const int r = e >> l2cpr, c = e & (cpr - 1);
                    const int q = row_pixel(r);
                    const __nv_bfloat16* src = p.res + ((img_out + (q < 0 ? 0 : q)) * p.Cout + c * 8);
                    cp_async16(stg + r * row_bytes + ((c ^ ((r >> l2rp) & swz_mask)) << 4), src, q >= 0);
                }
                cp_async_commit();
            }
            if (tid == 0) LFD_TRACE(2, tcount, 0);
            mbar_wait(&tfull[a], aph);
            tc_fence_after_sync();
            if (tid == 0) LFD_TRACE(2, tcount, 1);
            if (p.res) {
                cp_async_wait<0>();
                named_bar_sync(1, kEpiThreads);
            }
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + a * p.Cout;
            uint8_t* my_row = staging + m * row_bytes;
            const int my_swz = (m >> l2rp) & swz_mask;
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                float v[32];
                tmem_ld16(trow + c0, v);
                if (c0 + 16 < p.Cout) tmem_ld16(trow + c0 + 16, v + 16);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    if (c0 + h * 8 >= p.Cout) break;
                    const int chunk = (c0 >> 3) + h;
                    uint4* slot = reinterpret_cast<uint4*>(my_row + ((chunk ^ my_swz) << 4));
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(v[h * 8 + j], s_scale[c0 + h * 8 + j], s_shift[c0 + h * 8 + j]);
                    if (p.res) {
                        uint4 rv = *slot;
                        o[0] += bf16_lo(rv.x); o[1] += bf16_hi(rv.x); o[2] += bf16_lo(rv.y); o[3] += bf16_hi(rv.y);
                        o[4] += bf16_lo(rv.z); o[5] += bf16_hi(rv.z); o[6] += bf16_lo(rv.w); o[7] += bf16_hi(rv.w);
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
                    }
                    uint4 ov;
                    ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]);
                    ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
                    *slot = ov;
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&tempty[a]);  // accumulator stage may be overwritten by the next-but-one tile
            if (tid == 0) LFD_TRACE(2, tcount, 2);
            named_bar_sync(1, kEpiThreads);
            if (p.stats) {  // GroupNorm partial sums over the STORED (bf16) values; group = one 16 B chunk
                const int g = tid >> 3, sl = tid & 7;  // host guarantees Cout/groups == 8 and groups == 16
                float s1 = 0.f, s2 = 0.f;
                for (int r = sl; r < 128; r += 8) {
                    if (row_pixel(r) < 0) continue;
                    uint4 q = *reinterpret_cast<const uint4*>(staging + r * row_bytes + ((g ^ ((r >> l2rp) & swz_mask)) << 4));
                    float f[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y),
                                  bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
                }
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (sl == 0) {
                    double* dst = p.stats + ((size_t)n * p.gn_groups + g) * 2;
                    atomicAdd(dst, (double)s1);
                    atomicAdd(dst + 1, (double)s2);
                }
            }
            for (int e = tid; e < 128 * cpr; e += kEpiThreads) {  // coalesced store
                const int r = e >> l2cpr, c = e & (cpr - 1);
                const int q = row_pixel(r);
                if (q < 0) continue;
                const uint4 val = *reinterpret_cast<const uint4*>(staging + r * row_bytes + ((c ^ ((r >> l2rp) & swz_mask)) << 4));
                *reinterpret_cast<uint4*>(p.out + ((img_out + q) * p.Cout + c * 8)) = val;
            }
            named_bar_sync(1, kEpiThreads);  // staging free again
            if (tid == 0) LFD_TRACE(2, tcount, 3);
        }
    } else if (warp == kMmaWarp) {
        // ============================================================== MMA ISSUER
        const uint32_t idesc = umma_idesc_bf16(128, p.Cout);
        const uint32_t lbo_a = p.lbo_a, sbo_a = p.sbo_a;
        const uint32_t lbo_b = p.Cout * 16, sbo_b = 128;
        if (p.b_resident) {
            if (lane == 0) {
                mbar_arrive_expect_tx(wbar, p.w_total_bytes);
                for (uint32_t off = 0; off < p.w_total_bytes; off += 32768) {
                    uint32_t n = p.w_total_bytes - off < 32768 ? p.w_total_bytes - off : 32768;
                    bulk_g2s(smem_u32(wres) + off, reinterpret_cast<const uint8_t*>(p.w) + off, n, wbar);
                }
            }
            if (lane == 0) mbar_wait(wbar, 0);
        }
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; lane == 0 && tile < p.num_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
            LFD_TRACE(1, tcount, 0);
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after_sync();
            LFD_TRACE(1, tcount, 1);
            const uint32_t d_tmem = tmem_base + a * p.Cout;
            for (int cc = 0; cc < n_cc; ++cc, ++it) {
                const uint32_t s = it % SA, ph = (it / SA) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after_sync();
                fence_proxy_async_smem();
                if (cc == 0) LFD_TRACE(1, tcount, 2);
                {
                    const uint32_t a_base = smem_u32(ring) + s * p.stage_bytes;
                    const uint32_t b_base = p.b_resident ? smem_u32(wres) + cc * p.b_slice_bytes
                                                         : a_base + p.a_stage_bytes;
                    for (int tap = 0; tap < taps; ++tap) {
                        uint32_t view = 0;
                        if (MODE == MODE_3X3S1) view = (tap / 3) * 10 + (tap % 3);
                        if (MODE == MODE_3X3S2) {
                            const int kh = tap / 3, kw = tap % 3;
                            const int base = (kh == 1 ? 0 : 288) + (kw == 1 ? 0 : (kh == 1 ? 144 : 153));
                            view = base + (kh == 2 ? 9 : 0) + (kw == 2 ? 1 : 0);
                        }
                        for (int k16 = 0; k16 < p.Cc / 16; ++k16) {
                            uint64_t ad = umma_smem_desc(a_base + view * 16 + 2 * k16 * lbo_a, lbo_a, sbo_a);
                            uint64_t bd = umma_smem_desc(b_base + (tap * cpc + 2 * k16) * lbo_b, lbo_b, sbo_b);
                            umma_bf16(d_tmem, ad, bd, idesc, (cc | tap | k16) != 0);
                        }
                    }
                    umma_commit(&empty[s]);
                    if (cc == n_cc - 1) { umma_commit(&tfull[a]); LFD_TRACE(1, tcount, 3); }
                }
            }
        }
        __syncwarp();
    } else {
        // ============================================================== PRODUCERS
        const int ptid = tid - (kEpiThreads + 32);
        // every thread owns ONE 16-byte channel chunk (cpc divides 128) and walks the halo pixels with a fixed stride
        const int ch = ptid & (cpc - 1);
        const int px0 = ptid >> p.log2_cpc;
        const int pstep = kProdThreads >> p.log2_cpc;
        const uint32_t ch_dst = ch * p.lbo_a;
        // A full-barrier arrival for stage-iteration j is made at the end of iteration j+lag; the empty wait of
        // iteration j+SA must come later than that, hence lag <= SA-1.
        const uint32_t lag = SA >= 3 ? (uint32_t)kLag : 1u;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n = tile / p.tiles_per_img;
            const int t = tile - n * p.tiles_per_img;
            int iy0 = 0, ix0 = 0;
            if (MODE == MODE_FLAT) ix0 = t * 128;
            else {
                int oy0 = (t / p.tiles_x) * 16, ox0 = (t % p.tiles_x) * 8;
                iy0 = (MODE == MODE_3X3S1) ? oy0 : 2 * oy0;
                ix0 = (MODE == MODE_3X3S1) ? ox0 : 2 * ox0;
            }
            const __nv_bfloat16* img = p.in + (size_t)n * HW * p.Cin + ch * 8;
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
#pragma unroll 4
                    for (int pxi = px0; pxi < 128; pxi += pstep) {
                        const int q = ix0 + pxi;
                        const bool ok = q < HW;
                        cp_async16(dst_cc + pxi * 16, src_cc + (ok ? q : 0) * p.Cin, ok);
                    }
                } else {
#pragma unroll 4
                    for (int pxi = px0; pxi < p.n_px; pxi += pstep) {
                        const PxEntry pe = table[pxi];
                        const int y = iy0 + pe.dy, x = ix0 + pe.dx;
                        const bool ok = ((unsigned)y < (unsigned)p.H) && ((unsigned)x < (unsigned)p.W);
                        cp_async16(dst_cc + pe.slot * 16, src_cc + (ok ? (y * p.W + x) : 0) * p.Cin, ok);
                    }
                }
                cp_async_commit();
                if (ptid == 0) LFD_TRACE(0, it, 2);
                if (it >= lag) {
                    if (lag == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                    fence_proxy_async_smem();
                    mbar_arrive(&full[(it - lag) % SA]);
                    if (ptid == 0) LFD_TRACE(0, it - lag, 3);
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async_smem();
        for (uint32_t k = (it > lag ? it - lag : 0); k < it; ++k) mbar_arrive(&full[k % SA]);
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int umma_conv_configure(const ConvGeom& g, int num_sms, UmmaConvParams* out, size_t* smem_bytes, int* grid) {
    UmmaConvParams p;
    memset(&p, 0, sizeof(p));
    int mode;
    if (g.ksize == 1 && g.stride == 1) mode = MODE_FLAT;
    else if (g.ksize == 3 && g.stride == 1) mode = MODE_3X3S1;
    else if (g.ksize == 3 && g.stride == 2) mode = MODE_3X3S2;
    else if (g.ksize == 1 && g.stride == 2) mode = MODE_1X1S2;
    else return -1;
    if (g.Cin % 16 || g.Cout % 16 || g.Cout > 128 || g.Cout < 16) return -1;
    p.mode = mode;
    p.N = g.N; p.H = g.H; p.W = g.W; p.Cin = g.Cin; p.Ho = g.Ho; p.Wo = g.Wo; p.Cout = g.Cout;
    const int taps = g.ksize * g.ksize;
    int px_slots;  // plane size in pixels (slots), n_px = pixels actually loaded
    if (mode == MODE_FLAT) {
        p.tiles_x = 0; p.tiles_per_img = (g.Ho * g.Wo + 127) / 128;
        p.n_px = 128; px_slots = 128; p.sbo_a = 128;
    } else {
        p.tiles_x = (g.Wo + 7) / 8;
        p.tiles_per_img = p.tiles_x * ((g.Ho + 15) / 16);
        if (mode == MODE_3X3S1) { p.n_px = 180; px_slots = 180; p.sbo_a = 160; }
        else if (mode == MODE_3X3S2) { p.n_px = 561; px_slots = 594; p.sbo_a = 144; }
        else { p.n_px = 128; px_slots = 128; p.sbo_a = 128; }
    }
    p.lbo_a = px_slots * 16;
    p.num_tiles = p.tiles_per_img * g.N;
    const size_t staging = (size_t)128 * g.Cout * 2;
    const size_t fixed = kSmemStagingOff + staging;
    const size_t budget = 224 * 1024;
    const size_t w_total = (size_t)taps * g.Cin * g.Cout * 2;
    // Choose the channel chunk Cc, weight residency and ring depth.  Preference order:
    //   1. resident weights (loaded once per CTA) with >= 3 A stages, the largest Cc first;
    //   2. otherwise streamed weights: the largest Cc that still gives >= 4 stages (deep ring hides the per-stage
    //      weight fetch), then >= 3, then >= 2.
    // Memory-bound 1x1 layers additionally cap the ring so that two CTAs fit on one SM (latency hiding).
    const int cands[3] = {64, 32, 16};
    int best_cc = 0, best_res = 0, best_st = 0;
    auto stages_for = [&](int cc, int resident) -> int {
        if (g.Cin % cc) return 0;
        const size_t a_stage = (size_t)px_slots * cc * 2;
        const size_t b_slice = (size_t)taps * cc * g.Cout * 2;
        const size_t stage = a_stage + (resident ? 0 : b_slice);
        if (fixed + (resident ? w_total : 0) + 2 * stage > budget) return 0;
        int st = (int)((budget - fixed - (resident ? w_total : 0)) / stage);
        if (st > kMaxStages) st = kMaxStages;
        if (st > 4 && stage >= 8192) st = 4;
        return st;
    };
    for (int want = 3; want >= 2 && !best_cc; --want)
        for (int ci = 0; ci < 3 && !best_cc; ++ci) {
            int st = stages_for(cands[ci], 1);
            if (st >= want) { best_cc = cands[ci]; best_res = 1; best_st = st; }
        }
    if (!best_cc || best_st < 3) {
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
    p.a_stage_bytes = (uint32_t)((size_t)px_slots * best_cc * 2);
    p.b_slice_bytes = (uint32_t)((size_t)taps * best_cc * g.Cout * 2);
    p.stage_bytes = p.a_stage_bytes + (best_res ? 0 : p.b_slice_bytes);
    p.w_total_bytes = (uint32_t)w_total;
    p.smem_w_off = (uint32_t)fixed;
    p.smem_ring_off = (uint32_t)(fixed + (best_res ? w_total : 0));
    // two CTAs per SM when a 3-deep ring fits in half of the shared memory (1x1 layers)
    p.ctas_per_sm = 1;
    {
        const size_t half = (227 * 1024) / 2 - 1024;
        const size_t base = p.smem_ring_off;
        if (base + 2 * (size_t)p.stage_bytes <= half) {
            int st = (int)((half - base) / p.stage_bytes);
            if (st > p.stages) st = p.stages;
            if (st >= 2) { p.stages = st; p.ctas_per_sm = 2; }
        }
    }
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    p.log2_cpc = ilog2(p.Cc / 8);
    p.log2_cpr = ilog2(g.Cout / 8);
    if ((1 << p.log2_cpr) != g.Cout / 8) return -3;   // Cout must be 16/32/64/128
    p.log2_rp128 = g.Cout * 2 >= 128 ? 0 : ilog2(128 / (g.Cout * 2));
    p.tmem_cols = 2 * g.Cout < 32 ? 32 : 2 * g.Cout;  // two accumulator stages; power of two because Cout is
    *smem_bytes = p.smem_ring_off + (size_t)p.stages * p.stage_bytes;
    const int max_ctas = num_sms * p.ctas_per_sm;
    *grid = p.num_tiles < max_ctas ? p.num_tiles : max_ctas;
    *out = p;
    return 0;
}

template <int MODE>
static cudaError_t launch_mode(const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(224 * 1024));
        if (e != cudaSuccess) return e;
        configured = 224 * 1024;
    }
    conv_umma_kernel<MODE><<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t umma_conv_launch(const UmmaConvParams& p, size_t smem, int grid, cudaStream_t st) {
    switch (p.mode) {
        case MODE_FLAT: return launch_mode<MODE_FLAT>(p, smem, grid, st);
        case MODE_3X3S1: return launch_mode<MODE_3X3S1>(p, smem, grid, st);
        case MODE_3X3S2: return launch_mode<MODE_3X3S2>(p, smem, grid, st);
        case MODE_1X1S2: return launch_mode<MODE_1X1S2>(p, smem, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace lfd
