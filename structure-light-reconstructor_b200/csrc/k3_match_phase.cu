// k3_match_phase.cu — K3a: per-row phase correspondence search + Q-matrix triangulation.
// Replaces MFReconstruct::triangulation (Duke/mfreconstruct.cpp:272-334).
//
// Reference semantics: for every left pixel (i,j) that carries a phase, take the FIRST right column
// k in [0,W) of the same rectified row that carries a phase with fabs(pL - pR) < 0.1.  The reference
// scans k linearly (O(W^2) per row).  Here each persistent CTA owns one row at a time:
//   1. the L/R phase + mask rows arrive in shared memory by TMA bulk copies (cp.async.bulk +
//      mbarrier), double buffered so the next row streams in while this one is matched;
//   2. the right row is hashed by phase bucket (width 1/8 > 0.1) into chained lists in smem, pushed in
//      descending column blocks so that chains are ordered by block and walks can stop early;
//   3. each left pixel probes the three buckets that can hold a match, applies the exact
//      predicate and keeps the minimum k  ==  the reference's "first k" (exact, not approximate);
//   4. matched pixels are reprojected with Q in fp64 using the precomputed undistortPoints maps,
//      the row of XYZ / valid / match_k is staged in smem and leaves by TMA bulk stores.
#include <limits.h>
#include <stdlib.h>

#include "slr_device.cuh"

namespace {

constexpr int K3_MAX_THREADS = 1024;   // rows whose context allows one CTA per SM only run the widest CTA

struct K3aParams {
    const float *phase;     // [batch][2][H][W]
    const uint8_t *mask;    // [batch][2][H][W]
    int W, H, batch, HB;    // HB = hash buckets (power of two)
    const float *lx, *ly, *rx;  // undistort maps [H][W]
    float *xyz;
    uint8_t *valid;
    int32_t *match_k;       // may be null
    unsigned long long *n_points;  // may be null
    slr_calib_dev calib;
};

__global__ void __launch_bounds__(K3_MAX_THREADS)
k3a_phase_match(const K3aParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = p.W, HB = p.HB;
    const long long rows = (long long)p.batch * p.H;
    if ((long long)blockIdx.x >= rows) return;

    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    unsigned char *stage0 = smem + 16;
    const size_t stage_bytes = (size_t)10 * W;
    int *head = reinterpret_cast<int *>(stage0 + 2 * stage_bytes);
    int *next = head + HB;
    float *o_xyz = reinterpret_cast<float *>(next + W);
    int *o_k = reinterpret_cast<int *>(o_xyz + 3 * W);
    uint8_t *o_valid = reinterpret_cast<uint8_t *>(o_k + W);

    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) {
        slr::mbar_init(&bar[0], 1);
        slr::mbar_init(&bar[1], 1);
        slr::mbar_fence_init();
    }
    __syncthreads();

    auto issue_row = [&](long long r, int s) {
        const long long b = r / p.H;
        const int i = (int)(r - b * p.H);
        const size_t offL = ((size_t)(b * 2 + 0) * p.H + i) * W;
        const size_t offR = ((size_t)(b * 2 + 1) * p.H + i) * W;
        unsigned char *st = stage0 + s * stage_bytes;
        slr::mbar_expect_tx(&bar[s], (uint32_t)stage_bytes);
        slr::tma_load_1d(st, p.phase + offL, 4 * W, &bar[s]);
        slr::tma_load_1d(st + 4 * W, p.phase + offR, 4 * W, &bar[s]);
        slr::tma_load_1d(st + 8 * W, p.mask + offL, W, &bar[s]);
        slr::tma_load_1d(st + 9 * W, p.mask + offR, W, &bar[s]);
    };

    if (tid == 0) issue_row(blockIdx.x, 0);

    unsigned n_local = 0;
    int it = 0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x, ++it) {
        const int s = it & 1;
        if (tid == 0 && r + gridDim.x < rows) issue_row(r + gridDim.x, s ^ 1);

        for (int h = tid; h < HB; h += nthr) head[h] = -1;
        slr::mbar_wait(&bar[s], (it >> 1) & 1);
        const unsigned char *st = stage0 + s * stage_bytes;
        const float *pL = reinterpret_cast<const float *>(st);
        const float *pR = reinterpret_cast<const float *>(st + 4 * W);
        const uint8_t *mL = st + 8 * W;
        const uint8_t *mR = st + 9 * W;
        __syncthreads();

        // hash the right row: chained lists keyed by phase bucket.  Columns are pushed in descending blocks of
        // blockDim.x (one barrier per block), so every chain lists block 0's columns first, then block 1's, ...:
        // a walk can stop at the first entry of a later block than its best match (rows whose phases repeat
        // hundreds of times would otherwise cost a full chain per left pixel)
        for (int kb = (W - 1) / nthr; kb >= 0; kb--) {
            const int k = kb * nthr + tid;
            if (k < W && mR[k]) {
                const int slot = slr::phase_bucket(pR[k]) & (HB - 1);
                next[k] = atomicExch(&head[slot], k);
            }
            __syncthreads();
        }
        if (tid == 0) slr::tma_store_wait_read<0>();  // previous row's staged outputs have left smem
        __syncthreads();

        const long long b = r / p.H;
        const int i = (int)(r - b * p.H);
        const size_t map_row = (size_t)i * W;
        for (int j = tid; j < W; j += nthr) {
            int best = INT_MAX;
            if (mL[j]) {
                const float pl = pL[j];
                const int b0 = slr::phase_bucket(pl);
#pragma unroll
                for (int db = -1; db <= 1; db++) {
                    int k = head[(b0 + db) & (HB - 1)];
                    while (k >= 0) {
                        if (best != INT_MAX && k / nthr > best / nthr) break;  // only later blocks follow
                        if (slr::phase_match(pl, pR[k])) best = min(best, k);
                        k = next[k];
                    }
                }
            }
            float X = slr::qnan(), Y = slr::qnan(), Z = slr::qnan();
            const bool hit = (best != INT_MAX);
            if (hit) {
                // Utilities::undistortPoints((j,i),cam1) / ((k,i),cam2): precomputed maps
                const float ulx = __ldg(p.lx + map_row + j);
                const float uly = __ldg(p.ly + map_row + j);
                const float urx = __ldg(p.rx + map_row + best);
                const float disp = __fsub_rn(ulx, urx);  // float difference (:299)
                slr::reproject_q(p.calib, (double)ulx, (double)uly, (double)disp, X, Y, Z);
                n_local++;
            }
            o_xyz[3 * j + 0] = X;
            o_xyz[3 * j + 1] = Y;
            o_xyz[3 * j + 2] = Z;
            o_k[j] = hit ? best : -1;
            o_valid[j] = hit ? 1 : 0;
        }
        slr::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            slr::tma_store_1d(p.xyz + (size_t)r * W * 3, o_xyz, 12 * W);
            slr::tma_store_1d(p.valid + (size_t)r * W, o_valid, W);
            if (p.match_k) slr::tma_store_1d(p.match_k + (size_t)r * W, o_k, 4 * W);
            slr::tma_store_commit();
        }
    }
    if (tid == 0) slr::tma_store_wait_all<0>();

    if (p.n_points) {
        const unsigned long long s = slr::warp_sum_u32(n_local);
        if ((tid & 31) == 0 && s) atomicAdd(p.n_points, s);
    }
}

}  // namespace

static int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

slr_status slr_launch_match_phase_fast(slr_engine *e, const float *d_phase, const uint8_t *d_mask, int batch,
                                       float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                       unsigned long long *d_n_points, bool *handled);

slr_status slr_launch_match_phase(slr_engine *e, const float *d_phase, const uint8_t *d_mask, int batch,
                                  float *d_xyz, uint8_t *d_valid, int32_t *d_match_k,
                                  unsigned long long *d_n_points)
{
    // preferred: the fused kernel's match stage fed with phase rows (value-deduplicating hash, single-chain
    // probes); this plain kernel remains for shapes that one does not take and as the simple reference form
    if (!getenv("SLR_K3A_PLAIN")) {
        bool handled = false;
        slr_status st = slr_launch_match_phase_fast(e, d_phase, d_mask, batch, d_xyz, d_valid, d_match_k, d_n_points, &handled);
        if (st != SLR_OK || handled) return st;
    }
    SLR_REQUIRE(e->W % 16 == 0, "image width must be a multiple of 16 (TMA bulk rows); got %d", e->W);
    SLR_REQUIRE(((uintptr_t)d_phase | (uintptr_t)d_mask | (uintptr_t)d_xyz | (uintptr_t)d_valid |
                 (uintptr_t)d_match_k) % 16 == 0, "device buffers must be 16-byte aligned");
    K3aParams p;
    p.phase = d_phase;
    p.mask = d_mask;
    p.W = e->W;
    p.H = e->H;
    p.batch = batch;
    p.HB = next_pow2(e->W < 64 ? 64 : e->W);
    p.lx = e->d_undist_lx;
    p.ly = e->d_undist_ly;
    p.rx = e->d_undist_rx;
    p.xyz = d_xyz;
    p.valid = d_valid;
    p.match_k = d_match_k;
    p.n_points = d_n_points;
    p.calib = e->calib;
    const size_t smem = 16 + (size_t)20 * e->W + (size_t)4 * p.HB + (size_t)4 * e->W  // stages, head, next
                        + (size_t)12 * e->W + (size_t)4 * e->W + (size_t)e->W;        // xyz, k, valid
    SLR_REQUIRE(smem <= 227 * 1024, "image width %d needs %zu bytes of shared memory per row", e->W, smem);
    SLR_CHECK_CUDA(cudaFuncSetAttribute(k3a_phase_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 256-thread CTAs while several row contexts fit an SM; a single resident CTA gets 1024 threads
    int threads = 256, occ = 0;
    SLR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3a_phase_match, threads, smem));
    if (occ <= 1) {
        threads = (e->W >= 2048) ? 1024 : 512;
        SLR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3a_phase_match, threads, smem));
    }
    if (occ < 1) occ = 1;
    long long grid = (long long)e->num_sms * occ;
    const long long rows = (long long)batch * e->H;
    if (grid > rows) grid = rows;
    if (grid < 1) return SLR_OK;
    k3a_phase_match<<<(unsigned)grid, threads, smem, e->stream>>>(p);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
