// slr_device.cuh — device-side building blocks shared by the sm_100a kernels:
//   * strict-mode (reference-exact) and corrected-mode phase arithmetic
//   * the match predicate and the Q-matrix emitter
//   * mbarrier / TMA-bulk (cp.async.bulk) wrappers and streaming load/store helpers
//
// Strict-mode arithmetic uses the explicit round-to-nearest intrinsics (__fadd_rn, __dadd_rn, ...)
// so nvcc can never contract a multiply-add into an FMA: results are bit-identical to the IEEE
// evaluation of the reference expressions (Duke/mfreconstruct.cpp:231-269) on the host.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "slr_internal.h"

namespace slr {

__device__ __forceinline__ float qnan() { return __uint_as_float(SLR_QNAN_BITS); }

// ------------------------------------------------------------------------------------------------
// streaming global memory access
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream_f4(void *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copies (1-D cp.async.bulk; SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the smem sources of all but the newest `N` bulk groups have been read
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// strict mode: Duke/mfreconstruct.cpp:231-269
// ------------------------------------------------------------------------------------------------
#define SLR_ATAN_LUT_SIZE 511

// Wrapped phase of one frequency from a = G4-G2 and b = G1-G3 (:246-261, first matching branch wins).
// lut[q+255] = atan(float(q)).  Returns false on the degenerate branch (:254-255).
__device__ __forceinline__ bool wrapped_phase_strict(int a, int b, const float *__restrict__ lut, float &P)
{
    constexpr float PI = SLR_PI_DEC;
    constexpr float PI_3_2 = 3.0f * SLR_PI_DEC / 2.0f;  // 3*PI/2  (:250)
    constexpr float PI_1_2 = SLR_PI_DEC / 2.0f;          // PI/2    (:252)
    constexpr float PI_2 = 2.0f * SLR_PI_DEC;            // 2*PI    (:259)
    if (a == 0) {
        if (b == 0) return false;
        P = (b > 0) ? 0.0f : PI;
        return true;
    }
    if (b == 0) {
        P = (a > 0) ? PI_3_2 : PI_1_2;
        return true;
    }
    // C++ int division (truncation toward zero).  |a|,|b| <= 255: a non-integer quotient is at least
    // 1/255 away from the nearest integer, so truncating the correctly rounded float quotient is exact.
    const int q = __float2int_rz(__fdiv_rn((float)a, (float)b));
    const float at = lut[q + 255];
    if (b < 0)
        P = __fadd_rn(at, PI);
    else if (a > 0)
        P = __fadd_rn(at, PI_2);
    else
        P = at;
    return true;
}

// Heterodyne of :265-268.  P12/P23 are computed in double and narrowed; P123 and the scale in float.
__device__ __forceinline__ float heterodyne_strict(float P0, float P1, float P2)
{
    constexpr float PI_2 = 2.0f * SLR_PI_DEC;
    const double d01 = __dsub_rn((double)P0, (double)P1);
    const double d12 = __dsub_rn((double)P1, (double)P2);
    const float P12 = __double2float_rn((P0 > P1) ? d01 : __dadd_rn(d01, (double)PI_2));
    const float P23 = __double2float_rn((P1 > P2) ? d12 : __dadd_rn(d12, (double)PI_2));
    const float d = __fsub_rn(P12, P23);
    const float P123 = (P12 > P23) ? d : __fadd_rn(d, PI_2);
    return __fmul_rn(__fdiv_rn(P123, PI_2), 255.0f);
}

// One pixel, G[4*f+s] as ints.  Returns false if the pixel is dropped (degenerate branch).
__device__ __forceinline__ bool phase_strict(const int (&G)[12], const float *__restrict__ lut, float &phase)
{
    float P[3];
    bool ok = true;
#pragma unroll
    for (int f = 0; f < 3; f++) {
        float p = 0.0f;
        ok &= wrapped_phase_strict(G[4 * f + 3] - G[4 * f + 1], G[4 * f + 0] - G[4 * f + 2], lut, p);
        P[f] = p;
    }
    phase = heterodyne_strict(P[0], P[1], P[2]);
    return ok;
}

// ------------------------------------------------------------------------------------------------
// strict mode, table-driven form used by the kernels (tables built by slr_build_strict_tables, k_fused.cu)
// ------------------------------------------------------------------------------------------------
// byte i of a 32-bit word, zero extended: one PRMT
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0, 0x4440 + i); }

// ---- strict decode of one pixel from table lookups (Duke/mfreconstruct.cpp:239-268) ----------------
// Returns the wrapped phase of one frequency as a double holding the reference's float value.
// ptab rows (512 doubles each, entry 256 + signed quotient):
//   0: b > 0, a <= 0 -> atan(q)          1: b < 0 -> atan(q) + PI        2: b > 0, a > 0 -> atan(q) + 2PI
//   3: b == 0 -> 3PI/2 (a > 0) / PI/2 (a < 0); mtab[0] = 65536 makes the "quotient" there equal to a.
__device__ __forceinline__ double wrapped_strict_tab(int G1, int G2, int G3, int G4, const double *ptab,
                                                     const uint32_t *mtab, bool &ok)
{
    const int a = G4 - G2, b = G1 - G3;
    const int ua = abs(a), ub = abs(b);
    // floor(ua/ub) for 0 <= ua,ub <= 255 via M = floor(65536/ub)+1: the excess ua/65536 < 1/256 <= 1/ub can never
    // reach the next integer because a non-integer quotient has a fractional part <= 1 - 1/ub.
    const int q = (int)(((uint32_t)ua * mtab[ub]) >> 16);
    const int sg = (a ^ b) >> 31;                       // C++ int division truncates toward zero
    const int qs = (q ^ sg) - sg;
    int row = (b < 0) ? 512 : ((a > 0) ? 1024 : 0);
    row = (ub == 0) ? 1536 : row;                       // :250 / :252
    ok = ok && ((ua | ub) != 0);                        // :254 degenerate
    return ptab[row + 256 + qs];
}

__device__ __forceinline__ float heterodyne_strict_d(double P0, double P1, double P2)
{
    constexpr float PI_2 = 2.0f * SLR_PI_DEC;
    constexpr float RPI_2 = 1.0f / PI_2;                // RN(1/(2*PI))
    const double c = (double)PI_2;
    double d01 = __dsub_rn(P0, P1);
    double d12 = __dsub_rn(P1, P2);
    if (!(P0 > P1)) d01 = __dadd_rn(d01, c);
    if (!(P1 > P2)) d12 = __dadd_rn(d12, c);
    const float P12 = __double2float_rn(d01);
    const float P23 = __double2float_rn(d12);
    const float d = __fsub_rn(P12, P23);
    const float P123 = (P12 > P23) ? d : __fadd_rn(d, PI_2);
    // P123 / (2*PI), correctly rounded, without the generic division routine: with rc = RN(1/c),
    // q0 = RN(x*rc), rem = x - c*q0 (exact in an FMA), RN(q0 + rem*rc) == RN(x/c).  Verified
    // exhaustively on the host against IEEE division for every float with 2^-100 <= |x| < 32 and x = +0
    // (scratch/div_check.c); P123 is +0-free of sign issues and a multiple of 2^-24, so it is in range.
    const float q0 = __fmul_rn(P123, RPI_2);
    const float rem = __fmaf_rn(-q0, PI_2, P123);
    const float quo = __fmaf_rn(rem, RPI_2, q0);
    return __fmul_rn(quo, 255.0f);
}

// ------------------------------------------------------------------------------------------------
// corrected mode (atan2 + cascade of wrapped differences; no reference counterpart)
// ------------------------------------------------------------------------------------------------
#define SLR_TWO_PI_F 6.28318530717958647692f

__device__ __forceinline__ float wrap_2pi(float d) { return (d < 0.0f) ? __fadd_rn(d, SLR_TWO_PI_F) : d; }

// ------------------------------------------------------------------------------------------------
// match predicate: Duke/mfreconstruct.cpp:295  fabs(pL - pR) < 0.1  (float difference, float fabs,
// compared against the double 0.1 — equivalent to the float compare against 0.1f because
// pred(0.1f) < 0.1 < 0.1f).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool phase_match(float pl, float pr) { return fabsf(__fsub_rn(pl, pr)) < 0.1f; }

// Bucket of width 1/8 (> 0.1): any pR matching pL lies in bucket(pL)-1 .. bucket(pL)+1.
// Saturating conversion keeps huge / infinite phases in the end buckets (still adjacent-safe).
__device__ __forceinline__ int phase_bucket(float p) { return __float2int_rd(__fmul_rn(p, 8.0f)); }

// ------------------------------------------------------------------------------------------------
// emitters
// ------------------------------------------------------------------------------------------------
// p = Q * [x y d 1]^T in double (sums left to right), /w, narrowed to float, then the optional
// 3x4 rigid transform.  Duke/mfreconstruct.cpp:299-323, Duke/reconstruct.cpp:570-594.
__device__ __forceinline__ void reproject_q(const slr_calib_dev &c, double x, double y, double d, float &ox,
                                            float &oy, float &oz)
{
    double r[4];
    if (c.q_std && !(d == 0.0 && c.Q[15] == 0.0)) {
        // Q = [q00 0 0 q03; 0 q11 0 q13; 0 0 0 q23; 0 0 q32 q33] with q03, q13, q23 != 0 (what cv::stereoRectify
        // produces).  The skipped products are +-0 and x + (+-0) == x unless x is itself a zero, in which case the
        // following addition of a non-zero constant absorbs the sign: the four sums below are bit-identical to
        // the dense evaluation.  (d == 0 with q33 == 0 gives W = +-0, whose sign the dense order decides.)
        r[0] = __dadd_rn(__dmul_rn(c.Q[0], x), c.Q[3]);
        r[1] = __dadd_rn(__dmul_rn(c.Q[5], y), c.Q[7]);
        r[2] = c.Q[11];
        r[3] = __dadd_rn(__dmul_rn(c.Q[14], d), c.Q[15]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double s = __dmul_rn(c.Q[4 * i + 0], x);
            s = __dadd_rn(s, __dmul_rn(c.Q[4 * i + 1], y));
            s = __dadd_rn(s, __dmul_rn(c.Q[4 * i + 2], d));
            s = __dadd_rn(s, c.Q[4 * i + 3]);  // Q[i][3] * 1 is exact
            r[i] = s;
        }
    }
    float px = __double2float_rn(__ddiv_rn(r[0], r[3]));
    float py = __double2float_rn(__ddiv_rn(r[1], r[3]));
    float pz = __double2float_rn(__ddiv_rn(r[2], r[3]));
    if (c.has_rigid) {
        float o[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double s = __dmul_rn((double)c.rigid[4 * i + 0], (double)px);
            s = __dadd_rn(s, __dmul_rn((double)c.rigid[4 * i + 1], (double)py));
            s = __dadd_rn(s, __dmul_rn((double)c.rigid[4 * i + 2], (double)pz));
            s = __dadd_rn(s, (double)c.rigid[4 * i + 3]);
            o[i] = __double2float_rn(s);
        }
        px = o[0];
        py = o[1];
        pz = o[2];
    }
    ox = px;
    oy = py;
    oz = pz;
}

__device__ __forceinline__ unsigned long long warp_sum_u32(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace slr
