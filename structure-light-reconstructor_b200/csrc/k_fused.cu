// k_fused.cu — the fused multi-frequency pipeline: shadow mask + phase decode + heterodyne +
// per-row phase correspondence + Q-matrix triangulation in ONE kernel.
// Replaces MFReconstruct::runReconstruction minus image IO (Duke/mfreconstruct.cpp:160-334).
//
// HBM traffic is the algorithmic minimum: every stack byte is read once (TMA bulk copies into
// shared memory), every output byte written once (TMA bulk stores); the phase maps never exist in
// global memory.  One persistent CTA owns one rectified row of one scan at a time:
//
//   [TMA]   the 2*N image rows (both cameras, all planes) of the NEXT row stream into the stage
//           buffer while the current row is matched and triangulated;
//   decode  4 pixels per thread from shared memory, strict mode through two small tables
//           (exact integer quotient by reciprocal multiplication, and the finite set of wrapped
//           phase values atan(float(q)) + {0, PI, 2PI} held as doubles), heterodyne in fp64/fp32
//           exactly as the reference evaluates it;
//   match   right-row phases go into an open-addressing table keyed by the exact float value
//           (atomicCAS) holding the minimum column (atomicMin): strict-mode phases repeat heavily
//           (a few values occupy ~7 % of a row each), and only the smallest k of equal values can be
//           "the first k".  Each distinct value is then filed under the one or two phase buckets
//           (width 1/4) its +-0.1 match window touches, so a left pixel walks a single short chain,
//           applies the exact predicate and keeps the minimum k  ==  the reference's first-k linear
//           scan, exactly;
//   emit    warps take 32-pixel groups of the left row from a shared counter (chain lengths vary along
//           a row, so static assignment leaves warps idle at the barrier); Q reprojection in fp64 with
//           the precomputed undistortPoints maps; XYZ / valid / match_k written straight from registers.
//
// Rows are visited row-index-major (all scans' row i back to back) so the undistort-map row stays
// hot in L2 while the image stacks stream through with an evict-first policy.
#include <limits.h>
#include <stdlib.h>

#include "slr_device.cuh"

slr_status slr_unfused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr, int mode,
                          float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points);
slr_status slr_unfused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                          int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                          int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points);

namespace {

constexpr int FUSED_MAX_THREADS = 512;
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr int MODE_PHASE_INPUT = 2;  // rows of already decoded phase + mask (slr_match_triangulate_phase) instead of images

struct FusedParams {
    const uint8_t *stack;  // [batch][2][N][H][W]
    const float *phase;    // MODE_PHASE_INPUT: [batch][2][H][W]
    const uint8_t *mask;   // MODE_PHASE_INPUT: [batch][2][H][W]
    int W, H, batch, F, S, N;
    int T, logT;           // dedupe table / bucket heads size (power of two)
    int black_thr;
    int stagger_ns, num_sms;
    const float *lx, *ly, *rx;
    const double *ptab;    // strict: [4][512] wrapped-phase values; see build_strict_tables()
    const uint32_t *mtab;  // strict: [256] reciprocal multipliers
    float cs[16], sn[16];  // corrected: cos/sin(2 pi s / S)
    float *xyz;
    uint8_t *valid;
    int32_t *match_k;
    unsigned long long *n_points;
    slr_calib_dev calib;
};

__device__ __forceinline__ uint64_t make_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            slr::smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(slr::smem_u32(bar)), "l"(policy)
        : "memory");
}

// Window buckets.  Bucket width 1/4; a right value pR is filed under every bucket that the interval
// [pR - 0.11, pR + 0.11] touches (one or two), so a left value only probes its own bucket: any pL with
// fabs(pL - pR) < 0.1 lies inside that interval, and the clamp keeps the mapping monotone for huge values
// (where float spacing exceeds the margin every such value shares the end bucket).
__device__ __forceinline__ int window_bucket(float p)
{
    return __float2int_rd(__fmul_rn(fminf(fmaxf(p, -30000.0f), 30000.0f), 4.0f));
}

template <int MODE>
__device__ __forceinline__ void decode_chunk(const uint8_t *__restrict__ rows, int W, int c, const FusedParams &p,
                                             const double *s_ptab, const uint32_t *s_mtab, float (&ph)[4],
                                             bool (&ok)[4])
{
    // rows = [N][W] u8 for one camera in shared memory; c = 4-pixel chunk index
    const uint32_t *base = reinterpret_cast<const uint32_t *>(rows) + c;
    const int wstride = W >> 2;
    const uint32_t wv = base[0], bv = base[wstride];
#pragma unroll
    for (int i = 0; i < 4; i++) ok[i] = (int)slr::byte_of(wv, i) - (int)slr::byte_of(bv, i) > p.black_thr;  // computeShadows
    if (MODE == SLR_MODE_STRICT) {
        double P[3][4];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            const uint32_t g1 = base[(2 + 4 * f) * wstride], g2 = base[(3 + 4 * f) * wstride];
            const uint32_t g3 = base[(4 + 4 * f) * wstride], g4 = base[(5 + 4 * f) * wstride];
#pragma unroll
            for (int i = 0; i < 4; i++)
                P[f][i] = slr::wrapped_strict_tab(slr::byte_of(g1, i), slr::byte_of(g2, i), slr::byte_of(g3, i), slr::byte_of(g4, i), s_ptab,
                                             s_mtab, ok[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) ph[i] = slr::heterodyne_strict_d(P[0][i], P[1][i], P[2][i]);
    } else if (p.F == 3 && p.S == 4) {
        // the reference's 3 frequencies x 4 steps, fully unrolled (same arithmetic as the generic branch below)
        float l[3][4];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            const uint32_t g1 = base[(2 + 4 * f) * wstride], g2 = base[(3 + 4 * f) * wstride];
            const uint32_t g3 = base[(4 + 4 * f) * wstride], g4 = base[(5 + 4 * f) * wstride];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int a = (int)slr::byte_of(g4, i) - (int)slr::byte_of(g2, i);
                const int b = (int)slr::byte_of(g1, i) - (int)slr::byte_of(g3, i);
                ok[i] = ok[i] && ((a | b) != 0);
                float ang = atan2f((float)a, (float)b);
                if (ang < 0.0f) ang = __fadd_rn(ang, SLR_TWO_PI_F);
                l[f][i] = ang;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d01 = slr::wrap_2pi(__fsub_rn(l[0][i], l[1][i]));
            const float d12 = slr::wrap_2pi(__fsub_rn(l[1][i], l[2][i]));
            const float d = slr::wrap_2pi(__fsub_rn(d01, d12));
            ph[i] = __fmul_rn(__fdiv_rn(d, SLR_TWO_PI_F), 255.0f);
        }
    } else {
        float lvl[8][4];
        const int F = p.F, S = p.S;
        for (int f = 0; f < F; f++) {
            float num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};
            int inum[4] = {0, 0, 0, 0}, iden[4] = {0, 0, 0, 0};
            for (int s = 0; s < S; s++) {
                const uint32_t v = base[(2 + S * f + s) * wstride];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int g = (int)slr::byte_of(v, i);
                    if (S == 4) {
                        inum[i] += (s == 3) ? g : (s == 1) ? -g : 0;
                        iden[i] += (s == 0) ? g : (s == 2) ? -g : 0;
                    } else {
                        num[i] = __fsub_rn(num[i], __fmul_rn((float)g, p.sn[s]));
                        den[i] = __fadd_rn(den[i], __fmul_rn((float)g, p.cs[s]));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float nn = num[i], dd = den[i];
                if (S == 4) {
                    nn = (float)inum[i];
                    dd = (float)iden[i];
                    if (inum[i] == 0 && iden[i] == 0) ok[i] = false;
                } else if (__fadd_rn(__fmul_rn(nn, nn), __fmul_rn(dd, dd)) < 0.25f) {
                    ok[i] = false;
                }
                float a = atan2f(nn, dd);
                if (a < 0.0f) a = __fadd_rn(a, SLR_TWO_PI_F);
                lvl[f][i] = a;
            }
        }
        for (int n = F; n > 1; n--)
            for (int j = 0; j + 1 < n; j++)
#pragma unroll
                for (int i = 0; i < 4; i++) lvl[j][i] = slr::wrap_2pi(__fsub_rn(lvl[j][i], lvl[j + 1][i]));
#pragma unroll
        for (int i = 0; i < 4; i++) ph[i] = __fmul_rn(__fdiv_rn(lvl[0][i], SLR_TWO_PI_F), 255.0f);
    }
}

// MAXT/MINB: launch bounds.  Rows up to 1280 wide run 320-thread CTAs, two per SM (<= 102 registers);
// wider rows run up to 512 threads, one CTA per SM.
template <int MODE, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
k_fused_mf(const FusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = p.W, N = p.N, T = p.T;
    const long long rows = (long long)p.batch * p.H;
    if ((long long)blockIdx.x >= rows) return;
    const int tid = threadIdx.x, nthr = blockDim.x;

    // ---- shared memory carve-up ----
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    int *grp_ctr = reinterpret_cast<int *>(smem + 8);               // dynamic query-group counter
    unsigned char *stage = smem + 16;                               // [2][N][W] u8
    const size_t stage_bytes = (MODE == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * N * W;
    // [T] entries {x = distinct right phase (float bits), y = smallest right column carrying it}
    uint2 *ent = reinterpret_cast<uint2 *>(stage + stage_bytes);
    int *head = reinterpret_cast<int *>(ent + T);                   // [HB = 2T] bucket heads
    const int HB = 2 * T;
    int *nxt = head + HB;                                           // [2T]  node n = entry + T*(0|1)
    float *s_pl = reinterpret_cast<float *>(nxt + 2 * T);           // [W]   left phases (NaN = none)
    double *s_ptab = reinterpret_cast<double *>(s_pl + W);          // [2048] (strict)
    uint32_t *s_mtab = reinterpret_cast<uint32_t *>(s_ptab + 2048); // [256]  (strict)

    if (MODE == SLR_MODE_STRICT) {
        for (int i = tid; i < 2048; i += nthr) s_ptab[i] = p.ptab[i];
        for (int i = tid; i < 256; i += nthr) s_mtab[i] = p.mtab[i];
    }
    uint64_t policy = 0;
    if (tid == 0) {
        slr::mbar_init(bar, 1);
        slr::mbar_fence_init();
        policy = make_evict_first_policy();
    }
    __syncthreads();

    // row-index-major visiting order: r -> (i = r / batch, b = r % batch)
    auto issue_row = [&](unsigned r) {
        const int i = (int)(r / (unsigned)p.batch);
        const int b = (int)(r - (unsigned)i * (unsigned)p.batch);
        slr::mbar_expect_tx(bar, (uint32_t)stage_bytes);
        if (MODE == MODE_PHASE_INPUT) {  // stage = pL f32[W] | pR f32[W] | mL u8[W] | mR u8[W]
            const size_t offL = ((size_t)(b * 2 + 0) * p.H + i) * W, offR = ((size_t)(b * 2 + 1) * p.H + i) * W;
            tma_load_1d_hint(stage, p.phase + offL, 4u * W, bar, policy);
            tma_load_1d_hint(stage + 4 * W, p.phase + offR, 4u * W, bar, policy);
            tma_load_1d_hint(stage + 8 * W, p.mask + offL, (uint32_t)W, bar, policy);
            tma_load_1d_hint(stage + 9 * W, p.mask + offR, (uint32_t)W, bar, policy);
            return;
        }
        const uint8_t *src = p.stack + ((size_t)b * 2 * N * p.H + i) * W;
        for (int v = 0; v < 2 * N; v++)   // plane v of this scan (cam-major, then image index)
            tma_load_1d_hint(stage + (size_t)v * W, src + (size_t)v * p.H * W, (uint32_t)W, bar, policy);
    };
    if (tid == 0) issue_row(blockIdx.x);
    // CTAs that share an SM would otherwise march in lock step (identical rows): the ALU-bound decode phases and
    // the shared-memory-bound match phases of both would coincide.  Delay the second wave by part of a row.
    if (p.stagger_ns > 0 && blockIdx.x >= (unsigned)p.num_sms) {
        for (int w = 0; w < p.stagger_ns; w += 500) __nanosleep(500);
    }

    const int nchunks = W >> 2;
    const int ngroups = (W + 63) >> 6;
    const int lane = tid & 31;
    unsigned n_local = 0;
    int it = 0;
    for (unsigned r = blockIdx.x; r < (unsigned)rows; r += gridDim.x, ++it) {
        const int i = (int)(r / (unsigned)p.batch);
        const int b = (int)(r - (unsigned)i * (unsigned)p.batch);

        // ---- clear the tables while the stage fills: keys = EMPTY, mink = INT_MAX, heads = -1 ----
        {
            uint4 *e4p = reinterpret_cast<uint4 *>(ent);   // two entries per vector
            uint4 *h4 = reinterpret_cast<uint4 *>(head);
            const uint4 e4 = make_uint4(KEY_EMPTY, 0x7fffffffu, KEY_EMPTY, 0x7fffffffu);
            const uint4 m4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            for (int t = tid; t < (T >> 1); t += nthr) {
                e4p[t] = e4;
                h4[t] = m4;
            }
            if (tid == 0) *grp_ctr = 0;
        }
        slr::mbar_wait(bar, it & 1);
        __syncthreads();  // tables cleared, stage landed

        // ---- decode: tasks [0, nchunks) = right 4-pixel chunks, [nchunks, 2 nchunks) = left chunks ----
        for (int task = tid; task < 2 * nchunks; task += nthr) {
            float ph[4];
            bool ok[4];
            if (task < nchunks) {
                const int c = task;
                if (MODE == MODE_PHASE_INPUT) {
                    const float4 v = reinterpret_cast<const float4 *>(stage + 4 * W)[c];
                    const uint32_t m = reinterpret_cast<const uint32_t *>(stage + 9 * W)[c];
                    ph[0] = v.x, ph[1] = v.y, ph[2] = v.z, ph[3] = v.w;
#pragma unroll
                    for (int q = 0; q < 4; q++) ok[q] = slr::byte_of(m, q) != 0 && ph[q] == ph[q];  // NaN never matches
                } else {
                    decode_chunk<MODE>(stage + (size_t)N * W, W, c, p, s_ptab, s_mtab, ph, ok);
                }
                // value -> min column, deduplicated.  The four pixels' first probes are issued back to back
                // (independent atomics in flight); the thread that claims a new value also files it under the
                // bucket(s) its +-0.1 match window touches.
                uint32_t key[4], h[4], old[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    key[q] = __float_as_uint(__fadd_rn(ph[q], 0.0f));  // -0 -> +0
                    h[q] = (key[q] * 2654435761u) >> (32 - p.logT);
                    old[q] = ok[q] ? atomicCAS(&ent[h[q]].x, KEY_EMPTY, key[q]) : key[q];
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    while (old[q] != KEY_EMPTY && old[q] != key[q]) {  // collision: linear probing
                        h[q] = (h[q] + 1) & (T - 1);
                        old[q] = atomicCAS(&ent[h[q]].x, KEY_EMPTY, key[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (ok[q]) atomicMin(reinterpret_cast<int *>(&ent[h[q]].y), 4 * c + q);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (ok[q] && old[q] == KEY_EMPTY) {
                        const float v = __uint_as_float(key[q]);
                        const int lo = window_bucket(__fsub_rn(v, 0.11f)), hi = window_bucket(__fadd_rn(v, 0.11f));
                        nxt[h[q]] = atomicExch(&head[lo & (HB - 1)], (int)h[q]);
                        if (hi != lo) nxt[h[q] + T] = atomicExch(&head[hi & (HB - 1)], (int)h[q] + T);
                    }
                }
            } else {
                const int c = task - nchunks;
                if (MODE == MODE_PHASE_INPUT) {
                    const float4 v = reinterpret_cast<const float4 *>(stage)[c];
                    const uint32_t m = reinterpret_cast<const uint32_t *>(stage + 8 * W)[c];
                    ph[0] = v.x, ph[1] = v.y, ph[2] = v.z, ph[3] = v.w;
#pragma unroll
                    for (int q = 0; q < 4; q++) ok[q] = slr::byte_of(m, q) != 0;
                } else {
                    decode_chunk<MODE>(stage, W, c, p, s_ptab, s_mtab, ph, ok);
                }
                reinterpret_cast<float4 *>(s_pl)[c] = make_float4(ok[0] ? ph[0] : slr::qnan(), ok[1] ? ph[1] : slr::qnan(),
                                                                  ok[2] ? ph[2] : slr::qnan(), ok[3] ? ph[3] : slr::qnan());
            }
        }
        __syncthreads();  // stage consumed, table + chains + left phases complete

        if (tid == 0 && r + gridDim.x < (unsigned)rows) issue_row(r + gridDim.x);  // prefetch the next row

        // ---- query + emit: warps take 64-pixel groups dynamically (chain lengths vary along the row);
        //      each lane owns two pixels of the group ----
        const float *lx_row = p.lx + (size_t)i * W, *ly_row = p.ly + (size_t)i * W, *rx_row = p.rx + (size_t)i * W;
        const size_t orow = ((size_t)b * p.H + i) * W;
        float *xyz_row = p.xyz + orow * 3;
        uint8_t *valid_row = p.valid + orow;
        int32_t *k_row = p.match_k ? p.match_k + orow : nullptr;
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(grp_ctr, 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= ngroups) break;
            const int j0 = (g << 6) + lane, j1 = j0 + 32;
            const float v0 = (j0 < W) ? s_pl[j0] : slr::qnan();
            const float v1 = (j1 < W) ? s_pl[j1] : slr::qnan();
            // the undistort-map values of both pixels are requested now so that their L2 latency overlaps the walk
            const float ulx0 = (v0 == v0) ? __ldg(lx_row + j0) : 0.0f, uly0 = (v0 == v0) ? __ldg(ly_row + j0) : 0.0f;
            const float ulx1 = (v1 == v1) ? __ldg(lx_row + j1) : 0.0f, uly1 = (v1 == v1) ? __ldg(ly_row + j1) : 0.0f;
            int best0 = INT_MAX, best1 = INT_MAX;
            if (v0 == v0) {
                int t = head[window_bucket(v0) & (HB - 1)];
                while (t >= 0) {
                    const uint2 e = ent[t & (T - 1)];
                    t = nxt[t];
                    if (slr::phase_match(v0, __uint_as_float(e.x))) best0 = min(best0, (int)e.y);
                }
            }
            if (v1 == v1) {
                int t = head[window_bucket(v1) & (HB - 1)];
                while (t >= 0) {
                    const uint2 e = ent[t & (T - 1)];
                    t = nxt[t];
                    if (slr::phase_match(v1, __uint_as_float(e.x))) best1 = min(best1, (int)e.y);
                }
            }
            // both pixels are reprojected unconditionally (dummy inputs where there is no match) so that the two
            // independent fp64 chains interleave; misses are turned into NaN afterwards
            const bool hit0 = (best0 != INT_MAX), hit1 = (best1 != INT_MAX);
            const float urx0 = hit0 ? __ldg(rx_row + best0) : 0.0f, urx1 = hit1 ? __ldg(rx_row + best1) : 0.0f;
            float X0, Y0, Z0, X1, Y1, Z1;
            slr::reproject_q(p.calib, (double)ulx0, (double)uly0, (double)__fsub_rn(ulx0, urx0), X0, Y0, Z0);
            slr::reproject_q(p.calib, (double)ulx1, (double)uly1, (double)__fsub_rn(ulx1, urx1), X1, Y1, Z1);
            n_local += (hit0 ? 1u : 0u) + (hit1 ? 1u : 0u);
            if (j0 < W) {
                float *dst = xyz_row + 3 * j0;
                dst[0] = hit0 ? X0 : slr::qnan();
                dst[1] = hit0 ? Y0 : slr::qnan();
                dst[2] = hit0 ? Z0 : slr::qnan();
                valid_row[j0] = hit0 ? 1 : 0;
                if (k_row) k_row[j0] = hit0 ? best0 : -1;
            }
            if (j1 < W) {
                float *dst = xyz_row + 3 * j1;
                dst[0] = hit1 ? X1 : slr::qnan();
                dst[1] = hit1 ? Y1 : slr::qnan();
                dst[2] = hit1 ? Z1 : slr::qnan();
                valid_row[j1] = hit1 ? 1 : 0;
                if (k_row) k_row[j1] = hit1 ? best1 : -1;
            }
        }
        // the next iteration clears the tables: every warp must be done probing them
        __syncthreads();
    }
    if (p.n_points) {
        const unsigned long long s = slr::warp_sum_u32(n_local);
        if ((tid & 31) == 0 && s) atomicAdd(p.n_points, s);
    }
}

}  // namespace

// strict-mode tables, built once per engine on the host with the host libm:
//   ptab[cs*512 + 256 + q]  (double holding the reference's float wrapped phase)
//      cs = 0: b > 0, a <= 0   atan(float(q))            (Duke/mfreconstruct.cpp:261, and :246 via q = 0)
//      cs = 1: b < 0           atan(float(q)) + PI       (:257, and :248 via q = 0)
//      cs = 2: b > 0, a > 0    atan(float(q)) + 2*PI     (:259)
//      cs = 3: b == 0          256 + a: PI/2 (a < 0, :252), 3*PI/2 (a > 0, :250)
//   mtab[ub] = floor(65536/ub) + 1  (mtab[0] = 65536)
slr_status slr_build_strict_tables(slr_engine *e)
{
    static double ptab[2048];
    static uint32_t mtab[256];
    const float PI = SLR_PI_DEC;
    for (int i = 0; i < 2048; i++) ptab[i] = 0.0;
    for (int q = -255; q <= 255; q++) {
        const float at = atanf((float)q);
        ptab[0 * 512 + 256 + q] = (double)at;
        ptab[1 * 512 + 256 + q] = (double)(at + PI);
        ptab[2 * 512 + 256 + q] = (double)(at + 2.0f * PI);
    }
    for (int q = 1; q <= 255; q++) {
        ptab[1536 + 256 + q] = (double)(3.0f * PI / 2.0f);   // b == 0, a > 0 (:250)
        ptab[1536 + 256 - q] = (double)(PI / 2.0f);          // b == 0, a < 0 (:252)
    }
    mtab[0] = 65536u;   // b == 0: the "quotient" is |a| itself (selects within row 3)
    for (int ub = 1; ub < 256; ub++) mtab[ub] = 65536u / (uint32_t)ub + 1u;
    if (!e->d_ptab) SLR_CHECK_CUDA(cudaMalloc(&e->d_ptab, sizeof(ptab)));
    if (!e->d_mtab) SLR_CHECK_CUDA(cudaMalloc(&e->d_mtab, sizeof(mtab)));
    SLR_CHECK_CUDA(cudaMemcpy(e->d_ptab, ptab, sizeof(ptab), cudaMemcpyHostToDevice));
    SLR_CHECK_CUDA(cudaMemcpy(e->d_mtab, mtab, sizeof(mtab), cudaMemcpyHostToDevice));
    return SLR_OK;
}

static size_t fused_smem_bytes(size_t stage_bytes, int W, int T)
{
    return 16 + stage_bytes + (size_t)24 * T + (size_t)4 * W + 2048 * 8 + 256 * 4;
}

// shared launcher: mode = SLR_MODE_STRICT | SLR_MODE_CORRECTED (image stacks) or MODE_PHASE_INPUT (phase + mask rows)
static slr_status launch_fused(slr_engine *e, int mode, const uint8_t *d_stack, const float *d_phase,
                               const uint8_t *d_mask, int batch, int F, int S, int black_thr, float *d_xyz,
                               uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points, bool *handled)
{
    *handled = false;
    const int W = e->W, N = 2 + F * S;
    int T = 64, logT = 6;
    while (T < W || (double)W / T > 0.7) T <<= 1, logT++;
    const size_t stage_bytes = (mode == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * N * W;
    const size_t smem = fused_smem_bytes(stage_bytes, W, T);
    const int nchunks = W / 4;
    const bool aligned = ((uintptr_t)d_stack | (uintptr_t)d_phase | (uintptr_t)d_mask | (uintptr_t)d_xyz |
                          (uintptr_t)d_valid | (uintptr_t)d_match_k) % 16 == 0;
    if (W % 16 != 0 || smem > 227 * 1024 || !aligned || batch > 65535 || (long long)batch * e->H >= (1LL << 31))
        return SLR_OK;  // not handled: the caller falls back to the un-fused kernels
    *handled = true;

    // decode tasks per row = 2*nchunks; a little more than one even share per round measured best (the extra
    // warps help the shared-memory-latency-bound match phase): 1280-wide rows -> 384 threads, two CTAs per SM
    const int tasks = 2 * nchunks;
    const int rounds = (tasks + FUSED_MAX_THREADS - 1) / FUSED_MAX_THREADS;
    int threads = ((tasks + rounds - 1) / rounds + 31) / 32 * 32 + 64;
    if (threads > FUSED_MAX_THREADS) threads = FUSED_MAX_THREADS;
    if (threads < 64) threads = 64;
    if (const char *ev = getenv("SLR_FUSED_THREADS")) {  // tuning knob (bench experiments)
        const int t = atoi(ev);
        if (t >= 64 && t <= FUSED_MAX_THREADS && t % 32 == 0) threads = t;
    }
    FusedParams p;
    p.stack = d_stack;
    p.phase = d_phase;
    p.mask = d_mask;
    p.W = W;
    p.H = e->H;
    p.batch = batch;
    p.F = F;
    p.S = S;
    p.N = N;
    p.T = T;
    p.logT = logT;
    p.black_thr = black_thr;
    p.num_sms = e->num_sms;
    p.stagger_ns = 0;
    if (const char *ev = getenv("SLR_FUSED_STAGGER_NS")) p.stagger_ns = atoi(ev);
    p.lx = e->d_undist_lx;
    p.ly = e->d_undist_ly;
    p.rx = e->d_undist_rx;
    p.ptab = e->d_ptab;
    p.mtab = e->d_mtab;
    for (int s = 0; s < 16; s++) {
        p.cs[s] = (s < S) ? (float)cos(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
        p.sn[s] = (s < S) ? (float)sin(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
    }
    p.xyz = d_xyz;
    p.valid = d_valid;
    p.match_k = d_match_k;
    p.n_points = d_n_points;
    p.calib = e->calib;

    void (*kern)(const FusedParams);
    const bool two = 2 * smem + 4096 <= 227 * 1024;
#define SLR_PICK(MAXT, MINB)                                                                                      \
    kern = (mode == SLR_MODE_STRICT)      ? k_fused_mf<SLR_MODE_STRICT, MAXT, MINB>                                 \
           : (mode == SLR_MODE_CORRECTED) ? k_fused_mf<SLR_MODE_CORRECTED, MAXT, MINB>                              \
                                          : k_fused_mf<MODE_PHASE_INPUT, MAXT, MINB>
    if (threads <= 320) {
        SLR_PICK(320, 2);
    } else if (threads <= 384) {
        SLR_PICK(384, 2);
    } else if (threads <= 448 && two) {
        SLR_PICK(448, 2);
    } else if (two) {
        SLR_PICK(512, 2);
    } else {
        SLR_PICK(FUSED_MAX_THREADS, 1);
    }
#undef SLR_PICK
    SLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SLR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)e->num_sms * occ;
    const long long rows = (long long)batch * e->H;
    if (grid > rows) grid = rows;
    if (grid < 1) return SLR_OK;
    kern<<<(unsigned)grid, threads, smem, e->stream>>>(p);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}

slr_status slr_launch_fused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr,
                               int mode, float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                               unsigned long long *d_n_points)
{
    if (mode == SLR_MODE_STRICT)
        SLR_REQUIRE(F == 3 && S == 4, "strict mode reproduces the reference's hard-coded 3 frequencies x 4 steps "
                                      "(Duke/mfreconstruct.cpp:237); got F=%d S=%d", F, S);
    else
        SLR_REQUIRE(mode == SLR_MODE_CORRECTED && F >= 1 && F <= 8 && S >= 3 && S <= 16,
                    "corrected mode supports 1<=F<=8, 3<=S<=16; got mode %d F=%d S=%d", mode, F, S);
    bool handled = false;
    slr_status st = launch_fused(e, mode, d_stack, nullptr, nullptr, batch, F, S, black_thr, d_xyz, d_valid, d_match_k,
                                 d_n_points, &handled);
    if (st != SLR_OK || handled) return st;
    return slr_unfused_mf(e, d_stack, batch, F, S, black_thr, mode, d_xyz, d_valid, d_match_k, d_n_points);
}

// K3a through the same kernel: rows of decoded phase + mask in, XYZ out (slr_match_triangulate_phase).
// Returns handled = false when the shape needs the plain k3a kernel (k3_match_phase.cu).
slr_status slr_launch_match_phase_fast(slr_engine *e, const float *d_phase, const uint8_t *d_mask, int batch,
                                       float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                       unsigned long long *d_n_points, bool *handled)
{
    return launch_fused(e, MODE_PHASE_INPUT, nullptr, d_phase, d_mask, batch, 1, 3, 0, d_xyz, d_valid, d_match_k,
                        d_n_points, handled);
}

slr_status slr_launch_fused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                               int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                               int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points)
{
    return slr_unfused_ge(e, d_stack, batch, nbits_col, black_thr, white_thr, scan_w, have_color, d_xyz, d_valid,
                          d_match_k, d_color, d_n_points);
}
