// k_fused.cu — the fused multi-frequency pipeline: shadow mask + phase decode + heterodyne +
// per-row phase correspondence + Q-matrix triangulation in ONE kernel.
// Replaces MFReconstruct::runReconstruction minus image IO (Duke/mfreconstruct.cpp:160-334).
//
// HBM traffic is the algorithmic minimum: every stack byte is read once (TMA bulk copies into
// shared memory), every output byte written once; the phase maps never exist in global memory.
// One persistent CTA owns one rectified row of one scan at a time:
//
//   [TMA]   the 2*N image rows (both cameras, all planes) of the NEXT row stream into the stage
//           buffer while the current row is matched and triangulated; the three undistort-map rows
//           (left x, left y, right x) of image row i are bulk-copied into shared memory once per i
//           (a CTA walks a contiguous range of (i, scan) pairs, scan fastest);
//   decode  4 pixels per thread from shared memory in full rounds, the remainder of the row one pixel
//           per thread so that every thread carries the same load.  Strict mode goes through two
//           small tables (exact integer quotient by reciprocal multiplication, and the finite set of
//           wrapped phase values atan(float(q)) + {0, PI, 2PI} held in exact 2^-24 fixed point), the
//           heterodyne follows the reference's double/float mix exactly (slr_device.cuh);
//   match   right-row phases go into an open-addressing table keyed by the exact float value
//           holding the minimum column: strict-mode phases repeat heavily (a few values occupy ~7 %
//           of a row each), and only the smallest k of equal values can be "the first k".  A plain
//           64-bit load of the entry settles the common case (value already present with a smaller
//           column); atomicCAS / atomicMin only run for new values and new minima.  Each distinct
//           value is then filed under the one or two phase buckets (width 1/4) its +-0.1 match
//           window touches, so a left pixel walks a single short chain, applies the exact predicate
//           and keeps the minimum k  ==  the reference's first-k linear scan, exactly;
//   emit    warps take pixel groups of the left row from a shared counter (chain lengths vary along
//           a row, so static assignment leaves warps idle at the barrier); Q reprojection in fp64 with
//           the undistortPoints maps from shared memory; XYZ / valid / match_k written from registers.
#include <limits.h>
#include <stdlib.h>

#include "k_fused_common.cuh"

slr_status slr_unfused_mf(slr_engine *e, const uint8_t *d_stack, int batch, int F, int S, int black_thr, int mode,
                          float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points);
slr_status slr_unfused_ge(slr_engine *e, const uint8_t *d_stack, int batch, int nbits_col, int black_thr,
                          int white_thr, int scan_w, int have_color, float *d_xyz, uint8_t *d_valid,
                          int32_t *d_match_k, uint8_t *d_color, unsigned long long *d_n_points);

// k_fused_flow.cu: the barrier-free dataflow schedule of the same pipeline (preferred when its row contexts fit)
slr_status slr_launch_fused_flow(slr_engine *e, int mode, const slr_fused::FusedParams &p, bool *handled);

namespace {

using namespace slr_fused;

// MAXT/MINB: launch bounds.  QPX: left pixels per lane in one query group (group = 32*QPX pixels).
// WCT > 0: the row width (table size, strict-mode plane count) as compile-time constants, as in k_fused_flow;
// instantiated for BASELINE config 4's 2048 pixels.
template <int MODE, int MAXT, int MINB, int QPX, int WCT = 0>
__global__ void __launch_bounds__(MAXT, MINB)
k_fused_mf(const FusedParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool CLAMP = MODE == MODE_PHASE_INPUT;
    const int W = WCT ? WCT : p.W, N = (WCT && MODE == SLR_MODE_STRICT) ? 14 : p.N, T = (WCT == 2048) ? 4096 : p.T;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;

    // this CTA's contiguous range of rows r = i * batch + b (scan index b fastest, so consecutive rows share the
    // undistort-map row i)
    const long long rows = (long long)p.batch * p.H;
    const long long r_begin = rows * blockIdx.x / gridDim.x, r_end = rows * (blockIdx.x + 1) / gridDim.x;
    if (r_begin >= r_end) return;

    // ---- shared memory carve-up ----
    uint64_t *bar_stage = reinterpret_cast<uint64_t *>(smem);
    uint64_t *bar_maps = reinterpret_cast<uint64_t *>(smem + 8);
    int *grp_ctr = reinterpret_cast<int *>(smem + 16);              // dynamic query-group counter
    unsigned char *stage = smem + 32;                               // [2][N][W] u8
    const size_t stage_bytes = (MODE == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * N * W;
    RowTables tab;
    tab.T = T;
    tab.logT = (WCT == 2048) ? 12 : p.logT;
    tab.HB = 2 * T;
    tab.ent = reinterpret_cast<uint2 *>(stage + stage_bytes);
    tab.head = reinterpret_cast<int *>(tab.ent + T);
    tab.nxt = tab.head + 2 * T;
    float *s_pl = reinterpret_cast<float *>(tab.nxt + 2 * T);       // [W]   left phases (NaN = none)
    float *s_lx = s_pl + W, *s_ly = s_lx + W, *s_rx = s_ly + W;     // [W]   undistort-map rows of image row i
    int *s_ptab = reinterpret_cast<int *>(s_rx + W);                // [SLR_PTAB_SIZE] (strict)
    uint32_t *s_btab = reinterpret_cast<uint32_t *>(s_ptab + SLR_PTAB_SIZE);  // [SLR_BTAB_SIZE] (strict)

    if (MODE == SLR_MODE_STRICT) {
        for (int k = tid; k < SLR_PTAB_SIZE; k += nthr) s_ptab[k] = p.ptab[k];
        for (int k = tid; k < SLR_BTAB_SIZE; k += nthr) s_btab[k] = p.btab[k];
    }
    uint64_t policy = 0;
    if (tid == 0) {
        slr::mbar_init(bar_stage, 1);
        slr::mbar_init(bar_maps, 1);
        slr::mbar_fence_init();
        policy = make_evict_first_policy();
    }
    __syncthreads();

    auto issue_row = [&](int i, int b) {
        slr::mbar_expect_tx(bar_stage, (uint32_t)stage_bytes);
        if (MODE == MODE_PHASE_INPUT) {  // stage = pL f32[W] | pR f32[W] | mL u8[W] | mR u8[W]
            const size_t offL = ((size_t)(b * 2 + 0) * p.H + i) * W, offR = ((size_t)(b * 2 + 1) * p.H + i) * W;
            tma_load_1d_hint(stage, p.phase + offL, 4u * W, bar_stage, policy);
            tma_load_1d_hint(stage + 4 * W, p.phase + offR, 4u * W, bar_stage, policy);
            tma_load_1d_hint(stage + 8 * W, p.mask + offL, (uint32_t)W, bar_stage, policy);
            tma_load_1d_hint(stage + 9 * W, p.mask + offR, (uint32_t)W, bar_stage, policy);
            return;
        }
        const uint8_t *src = p.stack + ((size_t)b * 2 * N * p.H + i) * W;
        for (int v = 0; v < 2 * N; v++)   // plane v of this scan (cam-major, then image index)
            tma_load_1d_hint(stage + (size_t)v * W, src + (size_t)v * p.H * W, (uint32_t)W, bar_stage, policy);
    };
    auto issue_maps = [&](int i) {  // L2-resident (15 MB for all rows), default policy
        slr::mbar_expect_tx(bar_maps, 12u * W);
        slr::tma_load_1d(s_lx, p.lx + (size_t)i * W, 4u * W, bar_maps);
        slr::tma_load_1d(s_ly, p.ly + (size_t)i * W, 4u * W, bar_maps);
        slr::tma_load_1d(s_rx, p.rx + (size_t)i * W, 4u * W, bar_maps);
    };

    int i = (int)(r_begin / p.batch), b = (int)(r_begin - (long long)i * p.batch);
    size_t orow = ((size_t)b * p.H + i) * W;            // first output pixel of row (b, i)
    const size_t scan_px = (size_t)p.H * W;
    if (tid == 0) {
        issue_row(i, b);
        issue_maps(i);
    }
    bool maps_pending = true;
    uint32_t maps_parity = 0;

    const int ntasks = W >> 1;                          // 4-pixel chunks, right and left interleaved
    const int full_tasks = (ntasks / nthr) * nthr;      // done in full rounds; the rest goes pixel by pixel
    const int rem_x0 = 2 * full_tasks;                  // first column left over (per camera)
    const int rem_px = W - rem_x0, rem_pad = (rem_px + 15) & ~15;
    constexpr int GPX = 32 * QPX;
    const int ngroups = (W + GPX - 1) / GPX;
    unsigned n_local = 0;
    int it = 0;
    for (long long r = r_begin; r < r_end; ++r, ++it) {
        // ---- clear the tables while the stage fills: keys = EMPTY, mink = INT_MAX, heads = -1 ----
        {
            uint4 *e4p = reinterpret_cast<uint4 *>(tab.ent);   // two entries per vector
            uint4 *h4 = reinterpret_cast<uint4 *>(tab.head);
            const uint4 e4 = make_uint4(KEY_EMPTY, 0x7fffffffu, KEY_EMPTY, 0x7fffffffu);
            const uint4 m4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            for (int t = tid; t < (T >> 1); t += nthr) {
                e4p[t] = e4;
                h4[t] = m4;
            }
            if (tid == 0) *grp_ctr = 0;
        }
        SLR_STAMP(0);
#ifdef SLR_PHASE_CLOCKS
        if (tid == 0 && it == DBG_SKIP) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.dbg[(((size_t)blockIdx.x * DBG_ROWS) * 16) * DBG_PTS + 6] = smid;
        }
#endif
        slr::mbar_wait(bar_stage, it & 1);
        __syncthreads();  // tables cleared, stage landed
        SLR_STAMP(1);

        // ---- decode: 4-pixel chunks, right and left chunk of the same columns on neighbouring lanes.  The right
        //      lane hands its upper two phases to the left lane, so every lane decodes 4 pixels and files 2 right
        //      pixels in the tables: all lanes, hence all warps, carry the same load ----
        for (int task = tid; task < full_tasks; task += nthr) {
            float ph[4];
            bool ok[4];
            const int x0 = 4 * (task >> 1);
            const bool right = (task & 1) == 0;
            load_phases<MODE, 4>(stage, W, N, x0, right, p, s_ptab, s_btab, ph, ok);
            const float n2 = __shfl_xor_sync(0xffffffffu, ph[2], 1), n3 = __shfl_xor_sync(0xffffffffu, ph[3], 1);
            const unsigned okb = __shfl_xor_sync(0xffffffffu, (ok[2] ? 1u : 0u) | (ok[3] ? 2u : 0u), 1);
            float ip[2] = {right ? ph[0] : n2, right ? ph[1] : n3};
            bool io[2] = {right ? ok[0] : (okb & 1u) != 0, right ? ok[1] : (okb & 2u) != 0};
            if (!SLR_ABLATE(4)) insert_right<2, CLAMP>(tab, ip, io, right ? x0 : x0 + 2);
            if (!right)
                reinterpret_cast<float4 *>(s_pl)[x0 >> 2] = make_float4(ok[0] ? ph[0] : slr::qnan(), ok[1] ? ph[1] : slr::qnan(),
                                                                        ok[2] ? ph[2] : slr::qnan(), ok[3] ? ph[3] : slr::qnan());
        }
        // ---- the remaining columns one pixel per thread: lanes 0-15 of a warp take right pixels, lanes 16-31 the
        //      left pixels of the same columns ----
        for (int s = tid; s < 2 * rem_pad; s += nthr) {
            float ph[1];
            bool ok[1];
            const bool right = (s & 16) == 0;
            const int xi = ((s >> 5) << 4) | (s & 15);
            const bool live = xi < rem_px;
            const int x = rem_x0 + (live ? xi : 0);
            load_phases<MODE, 1>(stage, W, N, x, right, p, s_ptab, s_btab, ph, ok);
            ok[0] = ok[0] && live;
            if (right && SLR_ABLATE(4))
                ;
            else if (right)
                insert_right<1, CLAMP>(tab, ph, ok, x);
            else if (live)
                s_pl[x] = ok[0] ? ph[0] : slr::qnan();
        }
        SLR_STAMP(2);
        __syncthreads();  // stage consumed, table + chains + left phases complete
        SLR_STAMP(3);

        // next row of this CTA; its image rows stream in while this row is matched
        int ni = i, nb = b + 1;
        if (nb == p.batch) nb = 0, ni = i + 1;
        const bool more = r + 1 < r_end;
        if (tid == 0 && more) issue_row(ni, nb);
        if (maps_pending) {
            slr::mbar_wait(bar_maps, maps_parity);
            maps_parity ^= 1u;
            maps_pending = false;
        }

        // ---- query + emit: warps take pixel groups dynamically (chain lengths vary along the row) ----
        float *xyz_row = p.xyz + orow * 3;
        uint8_t *valid_row = p.valid + orow;
        int32_t *k_row = p.match_k ? p.match_k + orow : nullptr;
        for (;;) {
            int g = 0;
            // (atom.inc in plain PTX: ptxas turns a single-lane atomicAdd into a 14-instruction warp-aggregation sequence)
            if (lane == 0) asm volatile("atom.relaxed.cta.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(g) : "r"(slr::smem_u32(grp_ctr)) : "memory");
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= ngroups) break;
            int j[QPX], best[QPX];
            float v[QPX];
#pragma unroll
            for (int u = 0; u < QPX; u++) {
                j[u] = g * GPX + 32 * u + lane;
                v[u] = (j[u] < W) ? s_pl[j[u]] : slr::qnan();
            }
#pragma unroll
            for (int u = 0; u < QPX; u++)
                best[u] = (v[u] != v[u]) ? INT_MAX : SLR_ABLATE(2) ? max(j[u] - 7, 0) : first_match<CLAMP>(tab, v[u]);
            // every pixel is reprojected unconditionally (dummy inputs where there is no match) so that independent
            // fp64 chains interleave; misses are turned into NaN afterwards
            float X[QPX], Y[QPX], Z[QPX];
#pragma unroll
            for (int u = 0; u < QPX; u++) {
                const bool hit = best[u] != INT_MAX;
                const float ulx = hit ? s_lx[j[u]] : 0.0f, uly = hit ? s_ly[j[u]] : 0.0f, urx = hit ? s_rx[best[u]] : -1.0f;  // misses: disparity 1
                if (SLR_ABLATE(1))
                    X[u] = ulx, Y[u] = uly, Z[u] = urx;
                else
                    slr::reproject_q(p.calib, (double)ulx, (double)uly, (double)__fsub_rn(ulx, urx), X[u], Y[u], Z[u]);
            }
#pragma unroll
            for (int u = 0; u < QPX; u++) {
                const bool hit = best[u] != INT_MAX;
                n_local += hit ? 1u : 0u;
                if (j[u] < W && !(SLR_ABLATE(16) && X[u] != 12345.0f)) {
                    float *dst = xyz_row + 3 * j[u];
                    dst[0] = hit ? X[u] : slr::qnan();
                    dst[1] = hit ? Y[u] : slr::qnan();
                    dst[2] = hit ? Z[u] : slr::qnan();
                    valid_row[j[u]] = hit ? 1 : 0;
                    if (k_row) k_row[j[u]] = hit ? best[u] : -1;
                }
            }
        }
        // the next iteration clears the tables: every warp must be done probing them
        SLR_STAMP(4);
        __syncthreads();
        SLR_STAMP(5);
        if (more && ni != i) {  // nobody reads the old map rows any more
            if (tid == 0) issue_maps(ni);
            maps_pending = true;
        }
        orow = (ni == i) ? orow + scan_px : orow - (size_t)(p.batch - 1) * scan_px + W;
        i = ni;
        b = nb;
    }
    if (p.n_points) {
        const unsigned long long s = slr::warp_sum_u32(n_local);
        if (lane == 0 && s) atomicAdd(p.n_points, s);
    }
}

}  // namespace

// strict-mode tables (layout: slr_device.cuh), built on the host with the host libm and checked for every (a, b)
static slr_status strict_tables_host(int32_t *ptab, uint32_t *btab)
{
    const float PI = SLR_PI_DEC;
    bool exact = true;
    auto fx = [&](float v) -> int32_t {  // v in units of 2^-24; exact for every value the reference can produce
        const double s = (double)v * 16777216.0;
        const int32_t r = (int32_t)s;
        if ((double)r != s) exact = false;
        return r;
    };
    for (int k = 0; k < SLR_PTAB_SIZE; k++) ptab[k] = SLR_PTAB_DEGENERATE;   // (entries no (a, b) pair reaches)
    for (int q = 0; q <= 255; q++) {
        const float atp = atanf((float)q), atn = atanf((float)(-q));
        ptab[SLR_PTAB_C_POS + q] = fx(atp + 2 * PI);                   // b > 0, a > 0  (:259)
        ptab[SLR_PTAB_C_POS - (q + 1)] = fx(atn);                      // b > 0, a <= 0 (:261 / :246)
        ptab[SLR_PTAB_C_NEG + q] = fx(atn + PI);                       // b < 0, a > 0  (:257)
        ptab[SLR_PTAB_C_NEG - (q + 1)] = fx(atp + PI);                 // b < 0, a <= 0 (:257 / :248)
        ptab[SLR_PTAB_C_ZERO + q] = fx(3 * PI / 2);                    // b == 0, a > 0 (:250): idx = a - 1
        ptab[SLR_PTAB_C_ZERO - (q + 1)] = fx(PI / 2);                  // b == 0, a < 0 (:252): idx = a - 1 <= -2
    }
    ptab[SLR_PTAB_C_ZERO - 1] = SLR_PTAB_DEGENERATE;                   // a == 0 and b == 0 (:254)
    if (!exact) {
        slr_set_error("strict tables: a wrapped phase is not a multiple of 2^-24 (host libm atanf out of spec?)");
        return SLR_ERR_INVALID;
    }
    for (int b = -255; b <= 255; b++) {
        const uint32_t ub = (uint32_t)abs(b);
        const uint32_t M = ub ? 65536u / ub + 1u : 65536u;
        const uint32_t centre = (b > 0) ? SLR_PTAB_C_POS : (b < 0) ? SLR_PTAB_C_NEG : SLR_PTAB_C_ZERO;
        btab[b + 256] = M | ((centre * 4u) << 17);
    }
    // every (a, b) pair through the device's own lookup against the branch form of :246-261 (C++ int division)
    for (int b = -255; b <= 255; b++)
        for (int a = -255; a <= 255; a++) {
            int32_t want;
            if (a == 0 && b == 0) want = SLR_PTAB_DEGENERATE;
            else if (a == 0) want = fx(b > 0 ? 0.0f : PI);
            else if (b == 0) want = fx(a > 0 ? 3 * PI / 2 : PI / 2);
            else {
                const float at = atanf((float)(a / b));
                want = fx(b < 0 ? at + PI : a > 0 ? at + 2 * PI : at);
            }
            if (slr::wrapped_strict_fx(a, b, ptab, btab) != want) {
                slr_set_error("strict tables: lookup disagrees with the branch form at a = %d, b = %d", a, b);
                return SLR_ERR_INVALID;
            }
        }
    btab[0] = 0;
    return SLR_OK;
}

// once per engine
slr_status slr_build_strict_tables(slr_engine *e)
{
    static int32_t ptab[SLR_PTAB_SIZE];
    static uint32_t btab[SLR_BTAB_SIZE];
    const slr_status st = strict_tables_host(ptab, btab);
    if (st != SLR_OK) return st;
    if (!e->d_ptab) SLR_CHECK_CUDA(cudaMalloc(&e->d_ptab, sizeof(ptab)));
    if (!e->d_btab) SLR_CHECK_CUDA(cudaMalloc(&e->d_btab, sizeof(btab)));
    SLR_CHECK_CUDA(cudaMemcpy(e->d_ptab, ptab, sizeof(ptab), cudaMemcpyHostToDevice));
    SLR_CHECK_CUDA(cudaMemcpy(e->d_btab, btab, sizeof(btab), cudaMemcpyHostToDevice));
    return SLR_OK;
}

// Host-only view of the same tables for tests: h_fx[(b + 255) * 511 + (a + 255)] = the kernels' lookup for a = G4 - G2,
// b = G1 - G3 (needs no GPU and no engine).
extern "C" slr_status slr_strict_tables_check(int32_t *h_fx)
{
    static int32_t ptab[SLR_PTAB_SIZE];
    static uint32_t btab[SLR_BTAB_SIZE];
    const slr_status st = strict_tables_host(ptab, btab);
    if (st != SLR_OK || !h_fx) return st;
    for (int b = -255; b <= 255; b++)
        for (int a = -255; a <= 255; a++) h_fx[(b + 255) * 511 + (a + 255)] = slr::wrapped_strict_fx(a, b, ptab, btab);
    return SLR_OK;
}

static size_t fused_smem_bytes(size_t stage_bytes, int W, int T)
{
    return 32 + stage_bytes + (size_t)24 * T + (size_t)16 * W + SLR_PTAB_SIZE * 4 + SLR_BTAB_SIZE * 4;
}

// shared launcher: mode = SLR_MODE_STRICT | SLR_MODE_CORRECTED (image stacks) or MODE_PHASE_INPUT (phase + mask rows)
static slr_status launch_fused(slr_engine *e, int mode, const uint8_t *d_stack, const float *d_phase,
                               const uint8_t *d_mask, int batch, int F, int S, int black_thr, float *d_xyz,
                               uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points, bool *handled,
                               bool raw = false)
{
    *handled = false;
    const int W = e->W, N = 2 + F * S;
    int T = 64, logT = 6;
    while (T < W || (double)W / T > 0.7) T <<= 1, logT++;
    const size_t stage_bytes = (mode == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * N * W;
    const size_t smem = fused_smem_bytes(stage_bytes, W, T);
    const bool aligned = ((uintptr_t)d_stack | (uintptr_t)d_phase | (uintptr_t)d_mask | (uintptr_t)d_xyz |
                          (uintptr_t)d_valid | (uintptr_t)d_match_k) % 16 == 0;
    if (W % 16 != 0 || smem > 227 * 1024 || !aligned || batch > 65535 || (long long)batch * e->H >= (1LL << 31))
        return SLR_OK;  // not handled: the caller falls back to the un-fused kernels
    *handled = true;

    // 512 threads, two CTAs per SM (64 registers) when two row contexts fit in shared memory: 32 warps per SM hide
    // the shared-memory and fp64 latencies of the match phase; narrow rows get one thread per 4-pixel chunk.
    const bool two = 2 * (smem + 1024) <= 228 * 1024;
    int threads = (W / 2 + 31) / 32 * 32;   // one 4-pixel decode task per thread, or the widest CTA
    if (threads > (two ? 512 : FUSED_MAX_THREADS)) threads = two ? 512 : FUSED_MAX_THREADS;
    if (threads < 64) threads = 64;
    if (const char *ev = getenv("SLR_FUSED_THREADS")) {  // tuning knob (bench experiments)
        const int t = atoi(ev);
        if (t >= 64 && t <= FUSED_MAX_THREADS && t % 32 == 0) threads = t;
    }
    int qpx = (W + 63) / 64 >= 3 * (threads / 32) ? 2 : 1;  // 64-pixel groups only when every warp still gets >= 3
    if (const char *ev = getenv("SLR_FUSED_QPX")) qpx = atoi(ev) == 2 ? 2 : 1;
    FusedParams p;
    p.stack = d_stack;
    p.map1 = raw ? reinterpret_cast<const short2 *>(e->d_map1) : nullptr;
    p.map2 = raw ? e->d_map2 : nullptr;
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
    p.lx = e->d_undist_lx;
    p.ly = e->d_undist_ly;
    p.rx = e->d_undist_rx;
    p.ptab = e->d_ptab;
    p.btab = e->d_btab;
    for (int s = 0; s < 16; s++) {
        p.cs[s] = (s < S) ? (float)cos(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
        p.sn[s] = (s < S) ? (float)sin(2.0 * 3.14159265358979323846 * s / S) : 0.0f;
    }
    p.xyz = d_xyz;
    p.valid = d_valid;
    p.n_t = 1;
    p.xyz_t[0] = d_xyz;
    p.valid_t[0] = d_valid;
    {
        // writing this engine's block of its own assembled cloud with peers registered: the dataflow kernel stores
        // every row to the same block on all of them (an all-gather folded into the epilogue)
        const size_t off = (size_t)e->target_first_scan * e->W * e->H;
        if (e->n_targets > 1 && mode != MODE_PHASE_INPUT && d_xyz == e->xyz_t[0] + off * 3 && d_valid == e->valid_t[0] + off) {
            p.n_t = e->n_targets;
            for (int t = 1; t < e->n_targets; t++) {
                p.xyz_t[t] = e->xyz_t[t] + off * 3;
                p.valid_t[t] = e->valid_t[t] + off;
            }
        }
    }
    p.match_k = d_match_k;
    p.n_points = d_n_points;
    p.calib = e->calib;
#if !defined(SLR_ABLATION) && !defined(SLR_PHASE_CLOCKS)
    {
        bool flow = false;
        const slr_status st = slr_launch_fused_flow(e, mode, p, &flow);
        if (st == SLR_OK && flow && p.n_t > 1) e->targets_written = true;
        if (st != SLR_OK || flow) return st;
    }
#endif
    if (raw) {   // only the dataflow kernel rectifies on load
        *handled = false;
        return SLR_OK;
    }

    void (*kern)(const FusedParams);
#define SLR_PICK_Q(MAXT, MINB, Q)                                                                                 \
    kern = (mode == SLR_MODE_STRICT)      ? k_fused_mf<SLR_MODE_STRICT, MAXT, MINB, Q>                           \
           : (mode == SLR_MODE_CORRECTED) ? k_fused_mf<SLR_MODE_CORRECTED, MAXT, MINB, Q>                        \
                                          : k_fused_mf<MODE_PHASE_INPUT, MAXT, MINB, Q>
#define SLR_PICK(MAXT, MINB)          \
    if (qpx == 2) {                   \
        SLR_PICK_Q(MAXT, MINB, 2);    \
    } else {                          \
        SLR_PICK_Q(MAXT, MINB, 1);    \
    }
    if (threads <= 384 && two) {
        SLR_PICK(384, 2)
    } else if (two) {
        SLR_PICK(512, 2)
    } else if (threads <= 512) {
        SLR_PICK(512, 1)
    } else {
        SLR_PICK(1024, 1)
    }
#undef SLR_PICK
#undef SLR_PICK_Q
    if (mode == SLR_MODE_STRICT && W == 2048 && T == 4096 && N == 14 && threads == FUSED_MAX_THREADS && qpx == 1 && !two &&
        !getenv("SLR_FUSED_GENERIC_W"))
        kern = k_fused_mf<SLR_MODE_STRICT, 1024, 1, 1, 2048>;
    SLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SLR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)e->num_sms * occ;
    const long long rows = (long long)batch * e->H;
    if (grid > rows) grid = rows;
    if (grid < 1) return SLR_OK;
#ifdef SLR_ABLATION
    p.ablate = getenv("SLR_ABLATE") ? atoi(getenv("SLR_ABLATE")) : 0;
#endif
#ifdef SLR_PHASE_CLOCKS
    const size_t dbg_n = (size_t)grid * DBG_ROWS * 16 * DBG_PTS;
    SLR_CHECK_CUDA(cudaMalloc(&p.dbg, dbg_n * sizeof(long long)));
    SLR_CHECK_CUDA(cudaMemsetAsync(p.dbg, 0, dbg_n * sizeof(long long), e->stream));
#endif
    kern<<<(unsigned)grid, threads, smem, e->stream>>>(p);
    SLR_CHECK_LAUNCH(e);
#ifdef SLR_PHASE_CLOCKS
    if (const char *path = getenv("SLR_PHASE_CLOCKS_OUT")) {
        long long *h = (long long *)malloc(dbg_n * sizeof(long long));
        SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
        SLR_CHECK_CUDA(cudaMemcpy(h, p.dbg, dbg_n * sizeof(long long), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(path, "wb")) {
            int hdr[4] = {(int)grid, DBG_ROWS, 16, DBG_PTS};
            fwrite(hdr, sizeof(hdr), 1, f);
            fwrite(h, sizeof(long long), dbg_n, f);
            fclose(f);
        }
        free(h);
    }
    cudaFree(p.dbg);
#endif
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
    // TMA rows need 16-byte multiples and aligned bases: zero-padded / aligned copies on a child engine (slr_engine.cu)
    if (e->W % 16 != 0 || slr_misaligned16(d_stack, d_xyz, d_valid, d_match_k))
        return slr_padded_run(e, 0, d_stack, nullptr, batch, 2 + F * S, F, S, black_thr, mode, 0, d_xyz, d_valid, d_match_k,
                              nullptr, d_n_points);
    bool handled = false;
    slr_status st = launch_fused(e, mode, d_stack, nullptr, nullptr, batch, F, S, black_thr, d_xyz, d_valid, d_match_k,
                                 d_n_points, &handled);
    if (st != SLR_OK || handled) return st;
    return slr_unfused_mf(e, d_stack, batch, F, S, black_thr, mode, d_xyz, d_valid, d_match_k, d_n_points);
}

// RAW camera stacks in, XYZ out: stereoRect::doStereoRectify folded into the fused kernel's stage fill (k_fused_flow<RAW>).
// Shapes that kernel does not take are rectified by K0 into engine scratch, scan by scan, and then run as usual.
slr_status slr_launch_fused_mf_raw(slr_engine *e, const uint8_t *d_raw, int batch, int F, int S, int black_thr, int mode,
                                   float *d_xyz, uint8_t *d_valid, int32_t *d_match_k, unsigned long long *d_n_points)
{
    const size_t P = (size_t)e->W * e->H, in_bytes = (size_t)2 * (2 + F * S) * P;
    const bool shape_ok = (mode == SLR_MODE_STRICT) ? (F == 3 && S == 4)
                                                     : (mode == SLR_MODE_CORRECTED && F >= 1 && F <= 8 && S >= 3 && S <= 16);
    if (shape_ok && e->W % 16 == 0 && !slr_misaligned16(d_raw, d_xyz, d_valid, d_match_k)) {
        bool handled = false;
        const slr_status st = launch_fused(e, mode, d_raw, nullptr, nullptr, batch, F, S, black_thr, d_xyz, d_valid,
                                           d_match_k, d_n_points, &handled, true);
        if (st != SLR_OK || handled) return st;
    }
    if (e->stage_rect_bytes < in_bytes) {
        for (int k = 0; k < 2; k++) {
            if (e->d_stage_rect[k]) SLR_CHECK_CUDA(cudaFree(e->d_stage_rect[k]));
            e->d_stage_rect[k] = nullptr;
            SLR_CHECK_CUDA(cudaMalloc(&e->d_stage_rect[k], in_bytes));
        }
        e->stage_rect_bytes = in_bytes;
    }
    for (int b = 0; b < batch; b++) {
        slr_status st = slr_launch_rectify(e, d_raw + (size_t)b * in_bytes, 1, 2 + F * S, e->d_stage_rect[0]);
        if (st != SLR_OK) return st;
        st = slr_launch_fused_mf(e, e->d_stage_rect[0], 1, F, S, black_thr, mode, d_xyz + (size_t)b * P * 3,
                                 d_valid + (size_t)b * P, d_match_k ? d_match_k + (size_t)b * P : nullptr, d_n_points);
        if (st != SLR_OK) return st;
    }
    return SLR_OK;
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
    if (e->W % 16 != 0 || slr_misaligned16(d_stack, d_xyz, d_valid, d_match_k, have_color ? d_color : nullptr))
        return slr_padded_run(e, 1, d_stack, nullptr, batch, 2 + 2 * nbits_col, nbits_col, black_thr, white_thr, scan_w,
                              have_color, d_xyz, d_valid, d_match_k, have_color ? d_color : nullptr, d_n_points);
    return slr_unfused_ge(e, d_stack, batch, nbits_col, black_thr, white_thr, scan_w, have_color, d_xyz, d_valid,
                          d_match_k, d_color, d_n_points);
}
