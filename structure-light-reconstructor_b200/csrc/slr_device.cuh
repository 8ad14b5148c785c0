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
#include <limits.h>
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
// strict mode, table-driven integer form used by the kernels (tables built by slr_build_strict_tables, k_fused.cu)
// ------------------------------------------------------------------------------------------------
// byte i of a 32-bit word, zero extended: one PRMT
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0, 0x4440 + i); }

// SC * (byte i of x - byte i of y) + c for two words of four pixels, SC in -128..127: two integer dot products with a
// one-hot weight word (IDP.4A, unsigned bytes times signed weights, on the FMA pipe) instead of two byte extractions and
// a subtraction on the ALU pipe, which is the busiest pipe of the strict decode.  i is a compile-time constant
// wherever this is called (unrolled pixel loops), so the weight words are immediates.
template <int SC>
__device__ __forceinline__ int byte_diff(uint32_t x, uint32_t y, int i, int c = 0)
{
    const int wp = (int)((uint32_t)(SC & 0xff) << (8 * i)), wn = (int)((uint32_t)(-SC & 0xff) << (8 * i));
    int r;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(y), "r"(wn), "r"(c));
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(wp), "r"(r));
    return r;
}

// Every wrapped phase the reference can produce (Duke/mfreconstruct.cpp:246-261) is one of a few hundred floats in
// [-1.6, 7.9], each zero or >= 0.78 in magnitude, i.e. an integer multiple of 2^-24.  The tables hold them as
// int32 in units of 2^-24 ("fixed point", exact), so that the double-precision steps of :265-266 become exact
// integer arithmetic followed by ONE rounding (I2F), which is what narrowing the exact double to float does.
//
//   btab[b + 256], b = G1-G3 in [-255, 255]:  bits 0..16  M = floor(65536/|b|) + 1   (b == 0: 65536)
//                                             bits 17..   byte offset of the centre of b's group in ptab
//   idx = (a * M - 1) >> 16 (arithmetic shift, a = G4-G2 signed) tells the sign of a AND the quotient apart:
//       a >  0:  idx =  q        a <= 0:  idx = -(q + 1)        with q = floor(|a| / |b|) (b == 0: see below)
//     |a| * M / 65536 exceeds |a| / |b| by less than |a| / 65536 < 1/256 <= 1/|b|, and a non-integer quotient has a
//     fractional part <= 1 - 1/|b|: the scaled product never reaches q + 1, and for a != 0 it is never an integer
//     (it is strictly above |a| / |b| >= q).  So flooring (a*M - 1) / 65536 gives q for a > 0, -(q + 1) for a < 0 (the
//     negated ceiling) and -1 for a == 0, which shares its entry with a < 0, q == 0: both are atan(+-0) (+ PI).
//     C++ int division truncates toward zero, so the reference's quotient is sign(a)*sign(b)*q.
//   ptab groups of 512 entries around a centre (entry = centre + idx):
//     b > 0:  idx >= 0  atan(float( q)) + 2*PI  (:259)         idx < 0  atan(float(-q))       (:261, :246 via a == 0)
//     b < 0:  idx >= 0  atan(float(-q)) + PI    (:257)         idx < 0  atan(float( q)) + PI  (:257, :248 via a == 0)
//     b == 0 (M = 65536, idx = a - 1):  idx >= 0  3*PI/2 (:250)     idx == -1  SLR_PTAB_DEGENERATE (:254)
//                                       idx <= -2  PI/2 (:252)
// The centres of the first two groups sit 16 banks apart: the quotient is 0..7 for nine pixels in ten, and those
// entries of the four sign cases a warp mixes freely (idx -8..7 of both groups) occupy 32 distinct shared-memory banks.
// slr_build_strict_tables (k_fused.cu) checks every (a, b) pair against the branch form before an engine starts.
#define SLR_PTAB_C_POS 256                 // centre of the b > 0 group: entries 0..511
#define SLR_PTAB_C_NEG (512 + 16 + 256)    // centre of the b < 0 group: entries 528..1039 (bank 16 where C_POS is bank 0)
#define SLR_PTAB_C_ZERO (1040 + 256)       // centre of the b == 0 group: entries 1040..1551
#define SLR_PTAB_SIZE 1552
#define SLR_BTAB_SIZE 512
#define SLR_PTAB_DEGENERATE INT_MIN

// t = btab[b + 256]
__host__ __device__ __forceinline__ int strict_lookup(int a, uint32_t t, const int *__restrict__ ptab)
{
    const int sp = a * (int)(t & 0x1FFFFu) - 1;
    return *reinterpret_cast<const int *>(reinterpret_cast<const char *>(ptab) + (t >> 17) + ((sp >> 16) << 2));
}
__host__ __device__ __forceinline__ int wrapped_strict_fx(int a, int b, const int *__restrict__ ptab,
                                                          const uint32_t *__restrict__ btab)
{
    return strict_lookup(a, btab[b + 256], ptab);
}
// The same for pixel i of the four plane words G1..G4 of one frequency (a = G4 - G2, b = G1 - G3, :239-242), tables
// in shared memory: the address of b's btab entry comes straight out of the dot products.
__device__ __forceinline__ int wrapped_strict_fx_px(uint32_t g1, uint32_t g2, uint32_t g3, uint32_t g4, int i,
                                                    const int *__restrict__ s_ptab, const uint32_t *__restrict__ s_btab)
{
    uint32_t t;
    asm("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(byte_diff<4>(g1, g3, i, (int)smem_u32(s_btab + 256))));
    return strict_lookup(byte_diff<1>(g4, g2, i), t, s_ptab);
}

// Heterodyne of :265-268 on fixed-point wrapped phases.  I0 - I1 (+ 2*PI) is exact in int32 (|I| < 2^27), and
// (float)D * 2^-24 == float(double result of the reference) because scaling by a power of two commutes with
// rounding.  P123 and the division stay in fp32 as in the reference, carried in units of 2^-24 throughout (every
// operation below is exactly the reference's operation on scaled operands; no overflow, no subnormals).
// P123 / (2*PI), correctly rounded, without the generic division routine: with rc = RN(1/c), q0 = RN(x*rc),
// rem = x - c*q0 (exact in an FMA), RN(q0 + rem*rc) == RN(x/c).  Verified exhaustively on the host against IEEE
// division for every float with 2^-100 <= |x| < 32 and x = +0 (scratch/div_check.c).
__device__ __forceinline__ float heterodyne_strict_fx(int I0, int I1, int I2, bool &ok)
{
    constexpr float PI_2 = 2.0f * SLR_PI_DEC;
    constexpr float SC = 16777216.0f;                   // 2^24
    constexpr float C_S = PI_2 * SC;                    // exact
    constexpr float RC_S = (1.0f / PI_2) / SC;          // RN(1/(2*PI)) * 2^-24, exact scaling
    const int C_I = (int)C_S;                           // 2*PI in units of 2^-24 (a multiple of 8)
    ok = ok && (min(I0, min(I1, I2)) != SLR_PTAB_DEGENERATE);
    // (unsigned arithmetic: a degenerate entry would overflow int; its pixel is dropped anyway)
    // (no overflow for table values, so I0 > I1 is the sign of the difference: one compare-and-add on the difference)
    int d01 = (int)((unsigned)I0 - (unsigned)I1), d12 = (int)((unsigned)I1 - (unsigned)I2);
    if (d01 <= 0) d01 = (int)((unsigned)d01 + (unsigned)C_I);
    if (d12 <= 0) d12 = (int)((unsigned)d12 + (unsigned)C_I);
    const float P12 = __int2float_rn(d01);              // = P12 * 2^24
    const float P23 = __int2float_rn(d12);
    const float d = __fsub_rn(P12, P23);
    const float P123 = (P12 > P23) ? d : __fadd_rn(d, C_S);
    const float q0 = __fmul_rn(P123, RC_S);
    const float rem = __fmaf_rn(-q0, C_S, P123);
    const float quo = __fmaf_rn(rem, RC_S, q0);
    return __fmul_rn(quo, 255.0f);
}

// ------------------------------------------------------------------------------------------------
// corrected mode (atan2 + cascade of wrapped differences; no reference counterpart)
// ------------------------------------------------------------------------------------------------
#define SLR_TWO_PI_F 6.28318530717958647692f

__device__ __forceinline__ float wrap_2pi(float d) { return (d < 0.0f) ? __fadd_rn(d, SLR_TWO_PI_F) : d; }

// corrected-mode scale of the unwrapped phase to 0..255 (one multiply; the oracle's d / 2pi * 255 differs by <= 1 ulp)
__device__ __forceinline__ float phase_scale_corrected(float d) { return __fmul_rn(d, 255.0f / SLR_TWO_PI_F); }

// atan2(a, b) mapped to [0, 2*pi]: octant reduction, one approximate reciprocal (MUFU.RCP) and a degree-6 minimax
// polynomial in t^2 (max error 3.4e-7 rad on [0, 1], about one float ulp of the result) instead of libdevice's
// atan2f (a full-precision division with its slow path, plus special-case branches): ~20 instructions, no
// branches.  Corrected mode has no reference counterpart and is checked against the oracle's libm atan2f within the
// north star's 1e-4.  a == b == 0 gives NaN; those pixels are dropped by the caller.
__device__ __forceinline__ float atan2_pos(float a, float b)
{
    const float ax = fabsf(a), bx = fabsf(b);
    const float mx = fmaxf(ax, bx), mn = fminf(ax, bx);
    const float t = __fdividef(mn, mx);
    const float u = t * t;
    float q = 0.006811774335801601f;
    q = fmaf(q, u, -0.033604156225919724f);
    q = fmaf(q, u, 0.07962359488010406f);
    q = fmaf(q, u, -0.13233336806297302f);
    q = fmaf(q, u, 0.19807814061641693f);
    q = fmaf(q, u, -0.3331736922264099f);
    q = fmaf(q, u, 0.9999961256980896f);
    float r = q * t;
    if (ax > bx) r = 1.57079632679489661923f - r;
    if (b < 0.0f) r = 3.14159265358979323846f - r;
    if (a < 0.0f) r = SLR_TWO_PI_F - r;
    return r;
}

// ------------------------------------------------------------------------------------------------
// byte-parallel Gray decode helpers (k2_gray_decode.cu, k_fused_ge.cu): one 32-bit word = 4 pixels of one plane
// ------------------------------------------------------------------------------------------------
// bit 7 of every byte = (that byte of a > that byte of d), unsigned; the other bits are garbage
__device__ __forceinline__ uint32_t gt7(uint32_t a, uint32_t d)
{
    const uint32_t s = (a & 0x7f7f7f7fu) + (~d & 0x7f7f7f7fu);   // carry into bit 7 = (low 7 bits of a > low 7 bits of d)
    return (a & ~d) | (~(a ^ d) & s);
}

// Gray accumulator of 4 pixels x 8 bits: planes are visited LSB first, the new bit enters at bit 7 of each byte
__device__ __forceinline__ uint32_t push_bit(uint32_t acc, uint32_t t7)
{
    return ((acc >> 1) & 0x7f7f7f7fu) | (t7 & 0x80808080u);
}

// GrayCodes::grayToDec on two 16-bit lanes: prefix XOR from the MSB down
__device__ __forceinline__ uint32_t gray_to_binary_x2(uint32_t g)
{
    g ^= (g >> 1) & 0x7fff7fffu;
    g ^= (g >> 2) & 0x3fff3fffu;
    g ^= (g >> 4) & 0x0fff0fffu;
    g ^= (g >> 8) & 0x00ff00ffu;
    return g;
}

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
// float(RN(a0 / w)), float(RN(a1 / w)), float(RN(a2 / w)) — the three double divisions of the reprojection narrowed
// to float, bit-identical to IEEE division followed by the cast (Duke/mfreconstruct.cpp:303-309 on cv::Mat doubles).
// One shared reciprocal (MUFU seed + two Newton steps), one correction step per quotient: q is then within one
// double ulp of a / w, so (float)q equals (float)RN(a / w) unless q sits within a few double ulps of a float
// rounding boundary (the 29 dropped mantissa bits read 0x10000000 +- 8) — those lanes, and any operand or result
// outside the normal range (zero disparity with Q[3][3] == 0 gives w = 0), take the IEEE divisions.
__device__ __forceinline__ void div3_narrow(double a0, double a1, double a2, double w, float &f0, float &f1, float &f2)
{
    const unsigned ew = ((unsigned)__double2hiint(w) >> 20) & 0x7ffu;
    bool slow = (ew - 523u) > 1000u;                       // |w| outside [2^-500, 2^500], zero, inf, nan
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
    double e = fma(-w, r, 1.0);
    r = fma(r, e, r);
    e = fma(-w, r, 1.0);
    r = fma(r, e, r);
    auto one = [&](double a, float &f) {
        double q = a * r;
        const double rem = fma(-q, w, a);
        q = fma(rem, r, q);
        f = __double2float_rn(q);
        // float rounding boundary within 8 double ulps: the 29 dropped bits in 0x0FFFFFF8 .. 0x10000008 — tested as
        // "bits 5..28 of (dropped - 0x0FFFFFF8) are clear", i.e. .. 0x10000017: a superset, one add and one test
        slow |= (((unsigned)__double2loint(q) - 0x0FFFFFF8u) & 0x1FFFFFE0u) == 0u;
    };
    one(a0, f0);
    one(a1, f1);
    one(a2, f2);
    // a zero, subnormal or infinite result: smallest and largest magnitude of the three against the normal range (a NaN
    // quotient needs a NaN numerator here, w being finite and non-zero, and is a NaN on either path)
    slow |= !(fminf(fminf(fabsf(f0), fabsf(f1)), fabsf(f2)) >= 1.17549435e-38f &&
              fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fabsf(f2)) <= 3.40282347e+38f);
    if (slow) {
        f0 = __double2float_rn(__ddiv_rn(a0, w));
        f1 = __double2float_rn(__ddiv_rn(a1, w));
        f2 = __double2float_rn(__ddiv_rn(a2, w));
    }
}

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
    float px, py, pz;
    div3_narrow(r[0], r[1], r[2], r[3], px, py, pz);
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
