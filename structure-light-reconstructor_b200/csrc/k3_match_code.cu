// k3_match_code.cu — K3b: Gray-EPI code correspondence with the reference's `kstart` chain +
// Q-matrix triangulation.  Replaces Reconstruct::triangulation_ge (Duke/reconstruct.cpp:555-611).
//
// Reference semantics per rectified row: kstart = 0; for each left column j (ascending) that
// carries a code, take the first right column k >= kstart carrying the same code; on a match emit
// Q*[j, i, j-k, 1] and set kstart = k.  The chain makes the row sequential on the CPU.
//
// Here one persistent CTA owns a row at a time (rows arrive by TMA bulk copies, double buffered):
//   1. the right row is hashed by code into chained lists in shared memory;
//   2. the left row is cut into one short chunk per thread; every thread runs the chain over its
//      chunk from a guessed start state, then the CTA iterates to the fixed point: the start of a
//      chunk is the end state of the nearest earlier chunk that matched anything (block-wide
//      max-scan), and only chunks whose start changed are re-run.  After t rounds the first t
//      chunks are final, so the fixed point is the sequential answer (exact); on real rows it is
//      reached in 2-3 rounds because the chain state merges at the first matched pixel;
//   3. matches are reprojected with Q in fp64; XYZ is written from registers, valid / k / colour leave through
//      smem staging + TMA bulk stores (47 KB of shared memory per row context: four CTAs per SM).
#include <limits.h>

#include "slr_device.cuh"

namespace {

constexpr int K3B_THREADS = 256;

struct K3bParams {
    const int32_t *code;   // [batch][2][H][W]
    const uint8_t *mask;   // [batch][2][H][W]
    const uint8_t *white;  // white image of view 0 of scan 0 (or null); views are white_stride bytes apart
    size_t white_stride;
    int W, H, batch, HB, chunk;
    float *xyz;
    uint8_t *valid;
    int32_t *match_k;  // may be null
    uint8_t *color;    // may be null
    unsigned long long *n_points;
    slr_calib_dev calib;
};

// first k >= s in the chain of `code` (chains are unordered: take the minimum)
__device__ __forceinline__ int first_at_or_after(const int *head, const int *next, const int *cR, int HB, int code,
                                                 int s)
{
    int best = INT_MAX;
    int k = head[code & (HB - 1)];
    while (k >= 0) {
        if (cR[k] == code && k >= s) best = min(best, k);
        k = next[k];
    }
    return best;
}

__global__ void __launch_bounds__(K3B_THREADS, 4)
k3b_code_match(const K3bParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int W = p.W, HB = p.HB;
    const long long rows = (long long)p.batch * p.H;
    if ((long long)blockIdx.x >= rows) return;

    __shared__ int s_end[K3B_THREADS];
    __shared__ int s_warp[K3B_THREADS / 32];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    unsigned char *stage0 = smem + 64;
    const size_t stage_bytes = (size_t)10 * W;  // cL i32[W] | cR i32[W] | mL u8[W] | mR u8[W]
    int *head = reinterpret_cast<int *>(stage0 + 2 * stage_bytes);
    int *next = head + HB;
    int *o_k = next + W;
    uint8_t *o_valid = reinterpret_cast<uint8_t *>(o_k + W);
    uint8_t *o_color = o_valid + W;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
        slr::tma_load_1d(st, p.code + offL, 4 * W, &bar[s]);
        slr::tma_load_1d(st + 4 * W, p.code + offR, 4 * W, &bar[s]);
        slr::tma_load_1d(st + 8 * W, p.mask + offL, W, &bar[s]);
        slr::tma_load_1d(st + 9 * W, p.mask + offR, W, &bar[s]);
    };
    if (tid == 0) issue_row(blockIdx.x, 0);

    unsigned n_local = 0;
    int it = 0;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x, ++it) {
        const int s = it & 1;
        if (tid == 0 && r + gridDim.x < rows) issue_row(r + gridDim.x, s ^ 1);
        for (int h = tid; h < HB; h += K3B_THREADS) head[h] = -1;
        slr::mbar_wait(&bar[s], (it >> 1) & 1);
        const unsigned char *st = stage0 + s * stage_bytes;
        const int *cL = reinterpret_cast<const int *>(st);
        const int *cR = reinterpret_cast<const int *>(st + 4 * W);
        const uint8_t *mL = st + 8 * W;
        const uint8_t *mR = st + 9 * W;
        __syncthreads();

        for (int k = tid; k < W; k += K3B_THREADS)
            if (mR[k]) next[k] = atomicExch(&head[cR[k] & (HB - 1)], k);
        if (tid == 0) slr::tma_store_wait_read<0>();
        __syncthreads();

        // ---- chain: chunk `tid` covers left columns [j0, j1) ----
        const int j0 = min(tid * p.chunk, W), j1 = min(j0 + p.chunk, W);
        int start = 0, end = 0;
        bool need_run = true, any = false;
        while (true) {
            if (need_run) {
                end = start;
                any = false;
                for (int j = j0; j < j1; j++) {
                    int m = -1;
                    if (mL[j]) {
                        const int k = first_at_or_after(head, next, cR, HB, cL[j], end);
                        if (k != INT_MAX) {
                            m = k;
                            end = k;  // kstart = k (:604)
                            any = true;
                        }
                    }
                    o_k[j] = m;
                }
            }
            // start of chunk t = end state of the nearest earlier chunk that matched anything, else 0:
            // block-wide max-scan over (matched ? chunk index : -1)
            int key = any ? tid : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, key, o);
                if (lane >= o) key = max(key, v);
            }
            if (lane == 31) s_warp[warp] = key;
            s_end[tid] = end;
            __syncthreads();
            int carry = -1;
            for (int w2 = 0; w2 < warp; w2++) carry = max(carry, s_warp[w2]);
            int excl = __shfl_up_sync(0xffffffffu, key, 1);
            if (lane == 0) excl = -1;
            excl = max(excl, carry);
            const int new_start = (excl >= 0) ? s_end[excl] : 0;
            need_run = (new_start != start);
            start = new_start;
            if (!__syncthreads_or(need_run ? 1 : 0)) break;
        }

        // ---- emit ----
        const long long b = r / p.H;
        const int i = (int)(r - b * p.H);
        for (int j = tid; j < W; j += K3B_THREADS) {
            const int k = o_k[j];
            float X = slr::qnan(), Y = slr::qnan(), Z = slr::qnan();
            uint8_t colr = 0;
            if (k >= 0) {
                slr::reproject_q(p.calib, (double)j, (double)(i + p.calib.row0), (double)(j - k), X, Y, Z);  // :570
                if (p.color) {
                    const size_t vL = (size_t)(b * 2 + 0) * p.white_stride + (size_t)i * W;
                    const size_t vR = (size_t)(b * 2 + 1) * p.white_stride + (size_t)i * W;
                    colr = (uint8_t)(((int)p.white[vL + j] + (int)p.white[vR + k]) / 2);  // :598
                }
                n_local++;
            }
            float *dst = p.xyz + ((size_t)r * W + j) * 3;   // 12 bytes per lane, contiguous across the warp
            dst[0] = X;
            dst[1] = Y;
            dst[2] = Z;
            o_valid[j] = (k >= 0) ? 1 : 0;
            o_color[j] = colr;
        }
        slr::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            slr::tma_store_1d(p.valid + (size_t)r * W, o_valid, W);
            if (p.match_k) slr::tma_store_1d(p.match_k + (size_t)r * W, o_k, 4 * W);
            if (p.color) slr::tma_store_1d(p.color + (size_t)r * W, o_color, W);
            slr::tma_store_commit();
        }
    }
    if (tid == 0) slr::tma_store_wait_all<0>();
    if (p.n_points) {
        const unsigned long long sum = slr::warp_sum_u32(n_local);
        if (lane == 0 && sum) atomicAdd(p.n_points, sum);
    }
}

}  // namespace

static int next_pow2_i(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

slr_status slr_launch_match_code(slr_engine *e, const int32_t *d_col, const uint8_t *d_mask, int batch,
                                 const uint8_t *d_white, size_t white_view_stride, float *d_xyz,
                                 uint8_t *d_valid, int32_t *d_match_k, uint8_t *d_color,
                                 unsigned long long *d_n_points)
{
    SLR_REQUIRE(e->W % 16 == 0, "image width must be a multiple of 16 (TMA bulk rows); got %d", e->W);
    SLR_REQUIRE(((uintptr_t)d_col | (uintptr_t)d_mask | (uintptr_t)d_xyz | (uintptr_t)d_valid | (uintptr_t)d_match_k |
                 (uintptr_t)d_color) % 16 == 0, "device buffers must be 16-byte aligned");
    K3bParams p;
    p.code = d_col;
    p.mask = d_mask;
    p.white = d_color ? d_white : nullptr;
    p.white_stride = white_view_stride;
    p.W = e->W;
    p.H = e->H;
    p.batch = batch;
    p.HB = next_pow2_i(e->W < 64 ? 64 : e->W);
    p.chunk = (e->W + K3B_THREADS - 1) / K3B_THREADS;
    p.xyz = d_xyz;
    p.valid = d_valid;
    p.match_k = d_match_k;
    p.color = d_color;
    p.n_points = d_n_points;
    p.calib = e->calib;
    const size_t smem = 64 + (size_t)20 * e->W + (size_t)4 * p.HB + (size_t)4 * e->W + (size_t)4 * e->W + (size_t)2 * e->W;
    SLR_REQUIRE(smem <= 226 * 1024, "image width %d needs %zu bytes of shared memory per row", e->W, smem);
    SLR_CHECK_CUDA(cudaFuncSetAttribute(k3b_code_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    SLR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3b_code_match, K3B_THREADS, smem));
    if (occ < 1) occ = 1;
    long long grid = (long long)e->num_sms * occ;
    const long long rows = (long long)batch * e->H;
    if (grid > rows) grid = rows;
    if (grid < 1) return SLR_OK;
    k3b_code_match<<<(unsigned)grid, K3B_THREADS, smem, e->stream>>>(p);
    SLR_CHECK_LAUNCH(e);
    return SLR_OK;
}
