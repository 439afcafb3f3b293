// ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, cp.async, bulk copy (TMA 1-D), tcgen05
// (alloc / mma / commit / ld / fences), plus UMMA descriptor builders.
//
// Descriptor bit layouts follow the sm_100 UMMA conventions (shared-memory matrix descriptor:
// start>>4 @[0,14), LBO>>4 @[16,30), SBO>>4 @[32,46), version=1 @[46,48), layout type @[61,64);
// instruction descriptor: c_format @[4,6), a/b_format @[7,10)/[10,13), a/b_major @15/@16,
// N>>3 @[17,23), M>>4 @[24,29)).  All operands here are K-major, SWIZZLE_NONE ("interleaved" 8x16B
// core matrices): element (row r, k) of an operand lives at
//     start + (r%8)*16 + (r/8)*SBO + (k/8)*LBO + (k%8)*2        (bf16)
// which makes arbitrary 16-byte-aligned *shifted views* of a pixel plane legal operands -- the
// property the implicit-GEMM convolution in conv_umma.cu is built on.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace lfd {

#define LFD_DEVINL __device__ __forceinline__

// Watchdog: every spin-wait is bounded so that a protocol bug traps (context error, process exits)
// instead of hanging the GPU box.
#ifndef LFD_SPIN_LIMIT
#define LFD_SPIN_LIMIT (1u << 27)
#endif

LFD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
LFD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
LFD_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
LFD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LFD_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
LFD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
LFD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > LFD_SPIN_LIMIT) __trap();
    }
}

// ---------------------------------------------------------------- proxies / fences
LFD_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LFD_DEVINL void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
LFD_DEVINL void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- programmatic dependent launch
LFD_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
LFD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- cp.async (LDGSTS) 16 B with zero fill
LFD_DEVINL void cp_async16(uint32_t dst_smem, const void* src, bool valid) {
    uint32_t sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
LFD_DEVINL void cp_async16_full(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
LFD_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
LFD_DEVINL void cp_async4(uint32_t dst_smem, const void* src, bool valid) {   // 4 bytes, zero fill when !valid
    const uint32_t sz = valid ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
LFD_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
LFD_DEVINL void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// The mbarrier receives one arrival (counted against its expected-arrival count, hence .noinc) once ALL cp.async
// operations issued so far by this thread have completed -- the thread itself does not wait.
LFD_DEVINL void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- 1-D bulk copy (TMA engine, UBLKCP)
LFD_DEVINL void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// true on exactly one (the lowest active) lane of a converged warp
LFD_DEVINL bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- TMA tensor copies (cp.async.bulk.tensor)
// tmap: generic address of a CUtensorMap living in kernel-parameter (__grid_constant__) space
LFD_DEVINL void tma_store_3d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(c2), "r"(src_smem)
                 : "memory");
}
LFD_DEVINL void tma_store_4d(const void* tmap, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(tmap), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(src_smem)
                 : "memory");
}
LFD_DEVINL void tma_load_3d(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
LFD_DEVINL void tma_load_4d(uint32_t dst_smem, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst_smem),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
LFD_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
LFD_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// at most N of this thread's bulk groups may still be READING their shared-memory source
template <int N>
LFD_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- tcgen05: TMEM alloc
// cols: power of two in [32, 512]
LFD_DEVINL void tmem_alloc(uint32_t* slot_in_smem, uint32_t cols) {  // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
LFD_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t cols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// ---------------------------------------------------------------- tcgen05: descriptors
// K-major, no swizzle.  lbo/sbo in bytes (multiples of 16).
LFD_DEVINL uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// bf16 x bf16 (or fp16 x fp16) -> fp32, both operands K-major, M = 128, N = n.
LFD_DEVINL constexpr uint32_t umma_idesc_16(uint32_t m, uint32_t n, bool f16) {
    return (1u << 4)                    // c_format  = F32
           | ((f16 ? 0u : 1u) << 7)     // a_format  = F16 (0) / BF16 (1)
           | ((f16 ? 0u : 1u) << 10)    // b_format
           | ((n >> 3) << 17)           // N / 8
           | ((m >> 4) << 24);          // M / 16
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
LFD_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread have completed.
LFD_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- tcgen05: TMEM -> registers
// 32 lanes x 32-bit, 16 consecutive columns; thread t of warp w reads lane 32*(w%4)+t.
LFD_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
LFD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- misc
// Kernel time-line for tests/debug_timeline.py (only with -DLFD_B200_TIMELINE): tl[0] = earliest CTA start, tl[1] = latest CTA end,
// in %globaltimer nanoseconds; works inside CUDA-graph replays where events cannot be placed.
#ifdef LFD_B200_TIMELINE
LFD_DEVINL unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define LFD_TL_BEGIN(tl) do { if ((tl) && threadIdx.x == 0) atomicMin((tl), globaltimer_ns()); } while (0)
#define LFD_TL_END(tl) do { if ((tl) && threadIdx.x == 0) atomicMax((tl) + 1, globaltimer_ns()); } while (0)
#else
#define LFD_TL_BEGIN(tl) ((void)0)
#define LFD_TL_END(tl) ((void)0)
#endif
LFD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// max(x, 0) fused into the conversion (cvt.rn.relu)
LFD_DEVINL uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
LFD_DEVINL void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
LFD_DEVINL float lds32f(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
LFD_DEVINL uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
LFD_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
LFD_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
LFD_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---------------------------------------------------------------- 16-bit activation type: bf16 (F16 = false) or IEEE fp16 (true)
// Same bytes, same tensor-core rate; fp16 carries 3 more mantissa bits (the values on this path are post-BatchNorm / ReLU
// activations and folded weights of O(1), far inside the fp16 range).
LFD_DEVINL uint32_t pack_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // saturates at +-65504 instead of producing inf
    return d;
}
LFD_DEVINL uint32_t pack_f16x2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
LFD_DEVINL float f16_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))); }
LFD_DEVINL float f16_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }
template <bool F16> LFD_DEVINL uint32_t pack2(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
template <bool F16> LFD_DEVINL uint32_t pack2_relu(float lo, float hi) { return F16 ? pack_f16x2_relu(lo, hi) : pack_bf16x2_relu(lo, hi); }
template <bool F16> LFD_DEVINL float up_lo(uint32_t v) { return F16 ? f16_lo(v) : bf16_lo(v); }
template <bool F16> LFD_DEVINL float up_hi(uint32_t v) { return F16 ? f16_hi(v) : bf16_hi(v); }
template <bool F16> LFD_DEVINL float round16(float x) { return F16 ? __half2float(__float2half_rn(x)) : bf16_round(x); }
// the 16 raw bits of x rounded to the activation type
template <bool F16> LFD_DEVINL uint32_t bits16(float x) {
    return F16 ? (uint32_t)__half_as_ushort(__float2half_rn(x)) : (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x));
}
// run-time flavours for the cross-check kernels
LFD_DEVINL float up_lo_rt(uint32_t v, bool f16) { return f16 ? f16_lo(v) : bf16_lo(v); }
LFD_DEVINL float up_hi_rt(uint32_t v, bool f16) { return f16 ? f16_hi(v) : bf16_hi(v); }
LFD_DEVINL uint32_t pack2_rt(float lo, float hi, bool f16) { return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
LFD_DEVINL float round16_rt(float x, bool f16) { return f16 ? __half2float(__float2half_rn(x)) : bf16_round(x); }

}  // namespace lfd
