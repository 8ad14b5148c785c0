// k1_mf_decode.cu — K1: shadow mask + multi-frequency phase decode + heterodyne unwrap.
// Replaces MFReconstruct::computeShadows / decodePatterns / getPhase (Duke/mfreconstruct.cpp:190-269).
//
// Streaming kernel: each thread owns 16 consecutive pixels of one camera view, issues all
// N = 2 + F*S 128-bit plane loads up front (ld.global.nc, L1 no-allocate), decodes in registers and
// writes 4 x float4 phase + 1 x uint4 mask.  Algorithmic traffic: N + 4 + 1 bytes per pixel.
#include <stdlib.h>

#include "slr_device.cuh"

namespace {

constexpr int K1_THREADS = 256;

// strict mode, F = 3, S = 4, 16 pixels per thread, table-driven exact decode (slr_device.cuh: the integer quotient by
// reciprocal multiplication, the wrapped phase from the table of the reference's float values held in exact
// 2^-24 fixed point); persistent grid-stride CTAs load the 8 KB of tables into shared memory once
template <int NW>  // 32-bit words (4 pixels each) per thread per plane: 4 = 128-bit loads, 2 = 64-bit, 1 = 32-bit
__global__ void __launch_bounds__(K1_THREADS, NW == 4 ? 2 : NW == 2 ? 3 : 5)
k1_mf_decode_strict(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
                    int black_thr, const int *__restrict__ g_ptab, const uint32_t *__restrict__ g_btab,
                    float *__restrict__ phase, uint8_t *__restrict__ mask)
{
    __shared__ int s_ptab[SLR_PTAB_SIZE];
    __shared__ uint32_t s_btab[SLR_BTAB_SIZE];
    for (int i = threadIdx.x; i < SLR_PTAB_SIZE; i += K1_THREADS) s_ptab[i] = g_ptab[i];
    for (int i = threadIdx.x; i < SLR_BTAB_SIZE; i += K1_THREADS) s_btab[i] = g_btab[i];
    __syncthreads();

    for (long long chunk = (long long)blockIdx.x * K1_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K1_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * 14 * P + (size_t)c * (4 * NW);
        uint32_t img[14][NW];
#pragma unroll
        for (int n = 0; n < 14; n++) {
            if (NW == 4) {
                const uint4 v = slr::ldg_stream_u4(src + (size_t)n * P);
                img[n][0] = v.x, img[n][NW > 1 ? 1 : 0] = v.y, img[n][NW > 2 ? NW - 2 : 0] = v.z, img[n][NW - 1] = v.w;
            } else if (NW == 2) {
                const uint2 v = slr::ldg_stream_u2(src + (size_t)n * P);
                img[n][0] = v.x, img[n][NW - 1] = v.y;
            } else {
                img[n][0] = slr::ldg_stream_u32(src + (size_t)n * P);
            }
        }

        const size_t o = (size_t)view * P + (size_t)c * (4 * NW);
        uint32_t mk[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {  // one 32-bit word = 4 pixels of every plane
            auto word = [&](int n) { return img[n][w]; };
            float ph[4];
            uint32_t m4 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                bool ok = slr::byte_diff<1>(word(0), word(1), i) > black_thr;  // computeShadows (:199-204)
                int Pw[3];
#pragma unroll
                for (int f = 0; f < 3; f++)   // a = G4 - G2, b = G1 - G3 (:239-242)
                    Pw[f] = slr::wrapped_strict_fx_px(word(2 + 4 * f), word(3 + 4 * f), word(4 + 4 * f), word(5 + 4 * f), i,
                                                      s_ptab, s_btab);
                const float p = slr::heterodyne_strict_fx(Pw[0], Pw[1], Pw[2], ok);
                ph[i] = ok ? p : slr::qnan();
                m4 |= (ok ? 1u : 0u) << (8 * i);
            }
            slr::stg_stream_f4(phase + o + 4 * w, make_float4(ph[0], ph[1], ph[2], ph[3]));
            mk[w] = m4;
        }
        if (NW == 4)
            slr::stg_stream_u4(mask + o, make_uint4(mk[0], mk[NW > 1 ? 1 : 0], mk[NW > 2 ? NW - 2 : 0], mk[NW - 1]));
        else if (NW == 2)
            *reinterpret_cast<uint2 *>(mask + o) = make_uint2(mk[0], mk[NW - 1]);
        else
            *reinterpret_cast<uint32_t *>(mask + o) = mk[0];
    }
}

// strict mode, scalar fallback for P % 16 != 0 (one pixel per thread)
__global__ void __launch_bounds__(K1_THREADS)
k1_mf_decode_strict_scalar(const uint8_t *__restrict__ stack, size_t P, long long total, int black_thr,
                           const float *__restrict__ g_lut, float *__restrict__ phase, uint8_t *__restrict__ mask)
{
    __shared__ float lut[SLR_ATAN_LUT_SIZE + 1];
    for (int i = threadIdx.x; i < SLR_ATAN_LUT_SIZE; i += K1_THREADS) lut[i] = g_lut[i];
    __syncthreads();
    for (long long idx = (long long)blockIdx.x * K1_THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * K1_THREADS) {
        const long long view = idx / (long long)P;
        const size_t p = (size_t)(idx - view * (long long)P);
        const uint8_t *src = stack + (size_t)view * 14 * P + p;
        bool m = ((int)src[0] - (int)src[P]) > black_thr;
        int G[12];
#pragma unroll
        for (int n = 0; n < 12; n++) G[n] = src[(size_t)(2 + n) * P];
        float ph;
        const bool ok = slr::phase_strict(G, lut, ph);
        m = m && ok;
        phase[idx] = m ? ph : slr::qnan();
        mask[idx] = m ? 1 : 0;
    }
}

// corrected mode for the reference's stack shape (3 frequencies x 4 steps): 16 pixels per thread, all 14 plane loads
// issued up front as 128-bit streaming loads, atan2 by the branch-free polynomial of slr_device.cuh, cascade and scale
// in registers, vector stores.  Same arithmetic, pixel for pixel, as the generic kernel below and as the fused
// kernel's corrected branch.
template <int NW>
__global__ void __launch_bounds__(K1_THREADS, NW == 4 ? 2 : 3)
k1_mf_decode_corrected_3x4(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
                           int black_thr, float *__restrict__ phase, uint8_t *__restrict__ mask)
{
    for (long long chunk = (long long)blockIdx.x * K1_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K1_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * 14 * P + (size_t)c * (4 * NW);
        uint32_t img[14][NW];
#pragma unroll
        for (int n = 0; n < 14; n++) {
            if (NW == 4) {
                const uint4 v = slr::ldg_stream_u4(src + (size_t)n * P);
                img[n][0] = v.x, img[n][1] = v.y, img[n][NW - 2] = v.z, img[n][NW - 1] = v.w;
            } else {
                const uint2 v = slr::ldg_stream_u2(src + (size_t)n * P);
                img[n][0] = v.x, img[n][1] = v.y;
            }
        }
        const size_t o = (size_t)view * P + (size_t)c * (4 * NW);
        uint32_t mk[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            float ph[4];
            uint32_t m4 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                bool ok = slr::byte_diff<1>(img[0][w], img[1][w], i) > black_thr;
                float l[3];
#pragma unroll
                for (int f = 0; f < 3; f++) {
                    const int a = slr::byte_diff<1>(img[5 + 4 * f][w], img[3 + 4 * f][w], i);   // G4 - G2
                    const int b = slr::byte_diff<1>(img[2 + 4 * f][w], img[4 + 4 * f][w], i);   // G1 - G3
                    ok = ok && ((a | b) != 0);
                    l[f] = slr::atan2_pos((float)a, (float)b);
                }
                const float d01 = slr::wrap_2pi(__fsub_rn(l[0], l[1]));
                const float d12 = slr::wrap_2pi(__fsub_rn(l[1], l[2]));
                const float p = slr::phase_scale_corrected(slr::wrap_2pi(__fsub_rn(d01, d12)));
                ph[i] = ok ? p : slr::qnan();
                m4 |= (ok ? 1u : 0u) << (8 * i);
            }
            slr::stg_stream_f4(phase + o + 4 * w, make_float4(ph[0], ph[1], ph[2], ph[3]));
            mk[w] = m4;
        }
        if (NW == 4)
            slr::stg_stream_u4(mask + o, make_uint4(mk[0], mk[1], mk[NW - 2], mk[NW - 1]));
        else
            *reinterpret_cast<uint2 *>(mask + o) = make_uint2(mk[0], mk[1]);
    }
}

// corrected mode, generic F (<= 8) and S (3..16): 4 pixels per thread (32-bit loads).
struct CorrectedCoef {
    float cs[16];
    float sn[16];
};

// corrected mode with compile-time F and S (BASELINE config 5: 4 frequencies x 8 steps): 8 pixels per thread, all
// 2 + F*S plane loads issued up front as 64-bit streaming loads, the N-step sums, the cascade and its levels in
// registers.  Same arithmetic, pixel for pixel, as the generic kernel below.
template <int F, int S>
__global__ void __launch_bounds__(K1_THREADS, 2)
k1_mf_decode_corrected_fs(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
                          int black_thr, CorrectedCoef coef, float *__restrict__ phase, uint8_t *__restrict__ mask)
{
    constexpr int N = 2 + F * S;
    for (long long chunk = (long long)blockIdx.x * K1_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K1_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * N * P + (size_t)c * 8;
        uint32_t img[N][2];
#pragma unroll
        for (int n = 0; n < N; n++) {
            const uint2 v = slr::ldg_stream_u2(src + (size_t)n * P);
            img[n][0] = v.x, img[n][1] = v.y;
        }
        const size_t o = (size_t)view * P + (size_t)c * 8;
        uint32_t mk[2];
#pragma unroll
        for (int w = 0; w < 2; w++) {
            float ph[4];
            uint32_t m4 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                bool ok = slr::byte_diff<1>(img[0][w], img[1][w], i) > black_thr;
                float lvl[F];
#pragma unroll
                for (int f = 0; f < F; f++) {
                    float nn = 0.0f, dd = 0.0f;
                    if (S == 4) {   // exact integer form: num = G4-G2, den = G1-G3
                        const int in = slr::byte_diff<1>(img[2 + 4 * f + 3][w], img[2 + 4 * f + 1][w], i);
                        const int id = slr::byte_diff<1>(img[2 + 4 * f + 0][w], img[2 + 4 * f + 2][w], i);
                        if (in == 0 && id == 0) ok = false;
                        nn = (float)in, dd = (float)id;
                    } else {
#pragma unroll
                        for (int s2 = 0; s2 < S; s2++) {
                            const float g = (float)(int)slr::byte_of(img[2 + S * f + s2][w], i);
                            nn = __fsub_rn(nn, __fmul_rn(g, coef.sn[s2]));
                            dd = __fadd_rn(dd, __fmul_rn(g, coef.cs[s2]));
                        }
                        if (__fadd_rn(__fmul_rn(nn, nn), __fmul_rn(dd, dd)) < 0.25f) ok = false;
                    }
                    lvl[f] = slr::atan2_pos(nn, dd);
                }
#pragma unroll
                for (int n2 = F; n2 > 1; n2--)
#pragma unroll
                    for (int j = 0; j + 1 < n2; j++) lvl[j] = slr::wrap_2pi(__fsub_rn(lvl[j], lvl[j + 1]));
                const float p = slr::phase_scale_corrected(lvl[0]);
                ph[i] = ok ? p : slr::qnan();
                m4 |= (ok ? 1u : 0u) << (8 * i);
            }
            slr::stg_stream_f4(phase + o + 4 * w, make_float4(ph[0], ph[1], ph[2], ph[3]));
            mk[w] = m4;
        }
        *reinterpret_cast<uint2 *>(mask + o) = make_uint2(mk[0], mk[1]);
    }
}

template <int PX>
__global__ void __launch_bounds__(K1_THREADS)
k1_mf_decode_corrected(const uint8_t *__restrict__ stack, size_t P, long long chunks_per_view, long long total_chunks,
                       int F, int S, int black_thr, CorrectedCoef coef, float *__restrict__ phase,
                       uint8_t *__restrict__ mask)
{
    const int N = 2 + F * S;
    for (long long chunk = (long long)blockIdx.x * K1_THREADS + threadIdx.x; chunk < total_chunks;
         chunk += (long long)gridDim.x * K1_THREADS) {
        const long long view = chunk / chunks_per_view;
        const long long c = chunk - view * chunks_per_view;
        const uint8_t *src = stack + (size_t)view * N * P + (size_t)c * PX;
        uint32_t wv, bv;
        if (PX == 4) {
            wv = slr::ldg_stream_u32(src);
            bv = slr::ldg_stream_u32(src + P);
        } else {
            wv = src[0];
            bv = src[P];
        }
        float lvl[8][PX];
        bool ok[PX];
#pragma unroll
        for (int i = 0; i < PX; i++) ok[i] = (int)((wv >> (8 * i)) & 0xff) - (int)((bv >> (8 * i)) & 0xff) > black_thr;
        for (int f = 0; f < F; f++) {
            float num[PX], den[PX];
            int inum[PX], iden[PX];
#pragma unroll
            for (int i = 0; i < PX; i++) num[i] = den[i] = 0.0f, inum[i] = iden[i] = 0;
            for (int s = 0; s < S; s++) {
                const uint8_t *pp = src + (size_t)(2 + S * f + s) * P;
                const uint32_t v = (PX == 4) ? slr::ldg_stream_u32(pp) : (uint32_t)pp[0];
#pragma unroll
                for (int i = 0; i < PX; i++) {
                    const int g = (int)((v >> (8 * i)) & 0xff);
                    if (S == 4) {  // exact integer form: num = G4-G2, den = G1-G3
                        inum[i] += (s == 3) ? g : (s == 1) ? -g : 0;
                        iden[i] += (s == 0) ? g : (s == 2) ? -g : 0;
                    } else {
                        num[i] = __fsub_rn(num[i], __fmul_rn((float)g, coef.sn[s]));
                        den[i] = __fadd_rn(den[i], __fmul_rn((float)g, coef.cs[s]));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PX; i++) {
                float nn = num[i], dd = den[i];
                if (S == 4) {
                    nn = (float)inum[i];
                    dd = (float)iden[i];
                    if (inum[i] == 0 && iden[i] == 0) ok[i] = false;
                } else if (__fadd_rn(__fmul_rn(nn, nn), __fmul_rn(dd, dd)) < 0.25f) {
                    ok[i] = false;
                }
                lvl[f][i] = slr::atan2_pos(nn, dd);
            }
        }
        for (int n = F; n > 1; n--)
            for (int j = 0; j + 1 < n; j++)
#pragma unroll
                for (int i = 0; i < PX; i++) lvl[j][i] = slr::wrap_2pi(__fsub_rn(lvl[j][i], lvl[j + 1][i]));
        const size_t o = (size_t)view * P + (size_t)c * PX;
#pragma unroll
        for (int i = 0; i < PX; i++) {
            const float p = slr::phase_scale_corrected(lvl[0][i]);
            phase[o + i] = ok[i] ? p : slr::qnan();
            mask[o + i] = ok[i] ? 1 : 0;
        }
    }
}

}  // namespace

slr_status slr_launch_mf_decode(slr_engine *e, const uint8_t *d_stack, int views, int F, int S, int black_thr,
                                int mode, float *d_phase, uint8_t *d_mask)
{
    const size_t P = (size_t)e->W * e->H;
    if (mode == SLR_MODE_STRICT) {
        SLR_REQUIRE(F == 3 && S == 4, "strict mode reproduces the reference's hard-coded 3 frequencies x 4 steps "
                                      "(Duke/mfreconstruct.cpp:237); got F=%d S=%d", F, S);
        const bool vec = (P % 16 == 0) && (((uintptr_t)d_stack | (uintptr_t)d_phase | (uintptr_t)d_mask) % 16 == 0);
        if (vec) {
            int nw = 2;   // 8 pixels per thread, three CTAs per SM measured best (0.091 vs 0.097 ms per 8 scans for 16 pixels)
            if (const char *ev = getenv("SLR_K1_NW")) nw = atoi(ev) == 1 ? 1 : atoi(ev) == 2 ? 2 : 4;   // tuning knob
            const long long cpv = (long long)(P / (4 * nw));
            const long long total = cpv * views;
            long long blocks = (total + K1_THREADS - 1) / K1_THREADS;
            const long long cap = (long long)e->num_sms * (nw == 4 ? 2 : nw == 2 ? 3 : 5);   // persistent: tables loaded once per CTA
            if (blocks > cap) blocks = cap;
            if (blocks < 1) blocks = 1;
            if (nw == 4)
                k1_mf_decode_strict<4><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, black_thr,
                                                                                       e->d_ptab, e->d_btab, d_phase, d_mask);
            else if (nw == 1)
                k1_mf_decode_strict<1><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, black_thr,
                                                                                       e->d_ptab, e->d_btab, d_phase, d_mask);
            else
                k1_mf_decode_strict<2><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, black_thr,
                                                                                       e->d_ptab, e->d_btab, d_phase, d_mask);
        } else {
            const long long total = (long long)P * views;
            long long blocks = (total + K1_THREADS - 1) / K1_THREADS;
            const long long cap = (long long)e->num_sms * 32;
            if (blocks > cap) blocks = cap;
            if (blocks < 1) blocks = 1;
            k1_mf_decode_strict_scalar<<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, total, black_thr,
                                                                                       e->d_atan_lut, d_phase, d_mask);
        }
        SLR_CHECK_LAUNCH(e);
        return SLR_OK;
    }
    SLR_REQUIRE(mode == SLR_MODE_CORRECTED, "unknown mode %d", mode);
    SLR_REQUIRE(F >= 1 && F <= 8 && S >= 3 && S <= 16, "corrected mode supports 1<=F<=8, 3<=S<=16; got F=%d S=%d", F, S);
    CorrectedCoef coef;
    for (int s = 0; s < 16; s++) {
        coef.cs[s] = (s < S) ? (float)cos(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
        coef.sn[s] = (s < S) ? (float)sin(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
    }
    if (F == 3 && S == 4 && P % 16 == 0 && (((uintptr_t)d_stack | (uintptr_t)d_phase | (uintptr_t)d_mask) % 16) == 0) {
        const long long cpv = (long long)(P / 16), total = cpv * views;   // 16 pixels per thread (8 spills: 0.115 vs 0.099 ms)
        long long blocks = (total + K1_THREADS - 1) / K1_THREADS;
        const long long cap = (long long)e->num_sms * 2 * 4;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        k1_mf_decode_corrected_3x4<4><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, black_thr,
                                                                                     d_phase, d_mask);
        SLR_CHECK_LAUNCH(e);
        return SLR_OK;
    }
    if (F == 4 && S == 8 && P % 8 == 0 && (((uintptr_t)d_stack | (uintptr_t)d_mask) % 8) == 0 && ((uintptr_t)d_phase % 16) == 0) {
        const long long cpv = (long long)(P / 8), total = cpv * views;
        long long blocks = (total + K1_THREADS - 1) / K1_THREADS;
        const long long cap = (long long)e->num_sms * 2 * 4;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        k1_mf_decode_corrected_fs<4, 8><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, black_thr, coef,
                                                                                       d_phase, d_mask);
        SLR_CHECK_LAUNCH(e);
        return SLR_OK;
    }
    const bool vec = (P % 4 == 0) && (((uintptr_t)d_stack) % 4 == 0);
    const int px = vec ? 4 : 1;
    const long long cpv = (long long)(P / px);
    const long long total = cpv * views;
    long long blocks = (total + K1_THREADS - 1) / K1_THREADS;
    const long long cap = (long long)e->num_sms * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (vec)
        k1_mf_decode_corrected<4><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, F, S,
                                                                                  black_thr, coef, d_phase, d_mask);
    else
        k1_mf_decode_corrected<1><<<(unsigned)blocks, K1_THREADS, 0, e->stream>>>(d_stack, P, cpv, total, F, S,
                                                                                  black_thr, coef, d_phase, d_mask);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
