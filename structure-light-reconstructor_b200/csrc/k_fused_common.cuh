// k_fused_common.cuh — device code of the fused multi-frequency kernel (k_fused.cu), kept apart from its schedule:
// kernel parameters, strict / corrected decode of staged image rows, the value-deduplicating hash table with
// bucket chains, the first-match query.  See k_fused.cu for the algorithm.
#pragma once
#include <limits.h>
#include <stdlib.h>

#include "slr_device.cuh"

namespace slr_fused {


constexpr int FUSED_MAX_THREADS = 1024;   // one CTA per SM (rows too wide for two row contexts): 64 registers as well
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr int MODE_PHASE_INPUT = 2;  // rows of already decoded phase + mask (slr_match_triangulate_phase) instead of images

struct FusedParams {
    const uint8_t *stack;  // [batch][2][N][H][W]
    // k_fused_flow<RAW>: `stack` holds the RAW camera images and is rectified on the way into shared memory with the
    // CV_16SC2 maps of cv::initUndistortRectifyMap, [cam][H][W] each (NULL: the stack is rectified already)
    const short2 *map1;
    const uint16_t *map2;
    const float *phase;    // MODE_PHASE_INPUT: [batch][2][H][W]
    const uint8_t *mask;   // MODE_PHASE_INPUT: [batch][2][H][W]
    int W, H, batch, F, S, N;
    int T, logT;           // dedupe table size (power of two); 2T bucket heads
    int black_thr;
    const float *lx, *ly, *rx;
    const int *ptab;       // strict: [SLR_PTAB_SIZE] wrapped-phase values, 2^-24 fixed point (slr_device.cuh)
    const uint32_t *btab;  // strict: [SLR_BTAB_SIZE] reciprocal multipliers + row bases
    float cs[16], sn[16];  // corrected: cos/sin(2 pi s / S)
    float *xyz;
    uint8_t *valid;
    // k_fused_flow: the clouds every output row is stored to — [0] = xyz / valid above, [1..] = the same block of the
    // assembled cloud on every peer GPU (NVLink peer mappings, slr_set_gather_targets)
    int n_t;
    float *xyz_t[SLR_MAX_TARGETS];
    uint8_t *valid_t[SLR_MAX_TARGETS];
    int32_t *match_k;
    unsigned long long *n_points;
    slr_calib_dev calib;
#ifdef SLR_PHASE_CLOCKS
    long long *dbg;        // [grid][DBG_ROWS][16 warps][DBG_PTS] clock64 stamps (debug builds only)
#endif
#ifdef SLR_ABLATION
    int ablate;            // SLR_ABLATE bit mask: skip 1 reprojection, 2 chain walk, 4 insert, 8 decode math, 16 stores
#endif
};

#ifdef SLR_PHASE_CLOCKS
constexpr int DBG_ROWS = 8, DBG_PTS = 7, DBG_SKIP = 4;
#define SLR_STAMP(pt)                                                                                          \
    do {                                                                                                       \
        __syncwarp();                                                                                          \
        if (lane == 0 && tid < 512 && it >= DBG_SKIP && it < DBG_SKIP + DBG_ROWS)                                          \
            p.dbg[(((size_t)blockIdx.x * DBG_ROWS + (it - DBG_SKIP)) * 16 + (tid >> 5)) * DBG_PTS + (pt)] = clock64(); \
    } while (0)
#else
#define SLR_STAMP(pt) do { } while (0)
#endif
#ifdef SLR_ABLATION   // cost-attribution builds (csrc/Makefile ABLATE=1): results are wrong by construction
#define SLR_ABLATE(bit) ((p.ablate & (bit)) != 0)
#else
#define SLR_ABLATE(bit) false
#endif

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
template <bool CLAMP>
__device__ __forceinline__ int window_bucket(float p)
{
    // decoded phases are bounded (|p| < 1000); caller-supplied phase maps (MODE_PHASE_INPUT) are arbitrary floats
    if (CLAMP) p = fminf(fmaxf(p, -30000.0f), 30000.0f);
    return __float2int_rd(__fmul_rn(p, 4.0f));
}

// Slot of a phase value.  Two multiply rounds: a single multiplicative hash maps the arithmetic progressions that
// smooth (corrected-mode) phase ramps form in float-bit space onto a handful of slots for unlucky strides, and
// probing then degenerates into scans of hundreds of entries.
__device__ __forceinline__ uint32_t slot_of(uint32_t key, int logT)
{
    uint32_t h = key * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    return h >> (32 - logT);
}

// PX consecutive pixels (4 / 2 / 1: one 32- / 16- / 8-bit load per plane) of one camera row held in shared
// memory as rows = [N][W] u8; x0 = first pixel (a multiple of PX).
template <int MODE, int PX>
__device__ __forceinline__ void decode_px(const uint8_t *__restrict__ rows, int W, int x0, const FusedParams &p,
                                          const int *s_ptab, const uint32_t *s_btab, float (&ph)[PX], bool (&ok)[PX])
{
    auto plane = [&](int n) -> uint32_t {
        if (PX == 4) return *reinterpret_cast<const uint32_t *>(rows + (size_t)n * W + x0);
        if (PX == 2) return *reinterpret_cast<const uint16_t *>(rows + (size_t)n * W + x0);
        return rows[(size_t)n * W + x0];
    };
    const uint32_t wv = plane(0), bv = plane(1);
#pragma unroll
    for (int i = 0; i < PX; i++) ok[i] = slr::byte_diff<1>(wv, bv, i) > p.black_thr;  // computeShadows
    if (MODE == SLR_MODE_STRICT) {
        int P[3][PX];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            const uint32_t g1 = plane(2 + 4 * f), g2 = plane(3 + 4 * f), g3 = plane(4 + 4 * f), g4 = plane(5 + 4 * f);
#pragma unroll
            for (int i = 0; i < PX; i++)
                P[f][i] = slr::wrapped_strict_fx_px(g1, g2, g3, g4, i, s_ptab, s_btab);
        }
#pragma unroll
        for (int i = 0; i < PX; i++) ph[i] = slr::heterodyne_strict_fx(P[0][i], P[1][i], P[2][i], ok[i]);
    } else if (p.F == 3 && p.S == 4) {
        // the reference's 3 frequencies x 4 steps, fully unrolled (same arithmetic as the generic branch below)
        float l[3][PX];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            const uint32_t g1 = plane(2 + 4 * f), g2 = plane(3 + 4 * f), g3 = plane(4 + 4 * f), g4 = plane(5 + 4 * f);
#pragma unroll
            for (int i = 0; i < PX; i++) {
                const int a = slr::byte_diff<1>(g4, g2, i), b = slr::byte_diff<1>(g1, g3, i);
                ok[i] = ok[i] && ((a | b) != 0);
                l[f][i] = slr::atan2_pos((float)a, (float)b);
            }
        }
#pragma unroll
        for (int i = 0; i < PX; i++) {
            const float d01 = slr::wrap_2pi(__fsub_rn(l[0][i], l[1][i]));
            const float d12 = slr::wrap_2pi(__fsub_rn(l[1][i], l[2][i]));
            const float d = slr::wrap_2pi(__fsub_rn(d01, d12));
            ph[i] = slr::phase_scale_corrected(d);
        }
    } else {
        float lvl[8][PX];
        const int F = p.F, S = p.S;
        for (int f = 0; f < F; f++) {
            float num[PX], den[PX];
            int inum[PX], iden[PX];
#pragma unroll
            for (int i = 0; i < PX; i++) num[i] = den[i] = 0.0f, inum[i] = iden[i] = 0;
            for (int s = 0; s < S; s++) {
                const uint32_t v = plane(2 + S * f + s);
#pragma unroll
                for (int i = 0; i < PX; i++) {
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
#pragma unroll
        for (int i = 0; i < PX; i++) ph[i] = slr::phase_scale_corrected(lvl[0][i]);
    }
}

// phases of PX pixels of the right or left row, from the staged image rows (or the staged phase + mask rows)
// cam_pad = bytes between the end of the left camera's planes and the start of the right camera's in the stage buffer
template <int MODE, int PX>
__device__ __forceinline__ void load_phases(const unsigned char *stage, int W, int N, int x0, bool right,
                                            const FusedParams &p, const int *s_ptab, const uint32_t *s_btab,
                                            float (&ph)[PX], bool (&ok)[PX], int cam_pad = 0)
{
    if (MODE == MODE_PHASE_INPUT) {  // stage = pL f32[W] | pR f32[W] | mL u8[W] | mR u8[W]
        const float *src = reinterpret_cast<const float *>(stage + (right ? 4 * W : 0)) + x0;
        const unsigned char *m = stage + (right ? 9 * W : 8 * W) + x0;
#pragma unroll
        for (int q = 0; q < PX; q++) {
            ph[q] = src[q];
            ok[q] = m[q] != 0 && (!right || ph[q] == ph[q]);  // a NaN on the right never matches
        }
    } else if (SLR_ABLATE(8)) {
        const unsigned char *r0 = stage + (right ? (size_t)N * W + cam_pad : 0) + x0;
#pragma unroll
        for (int q = 0; q < PX; q++) ph[q] = (float)r0[2 * W + q] + 0.01f * (float)r0[6 * W + q], ok[q] = r0[q] > r0[W + q];
    } else {
        decode_px<MODE, PX>(stage + (right ? (size_t)N * W + cam_pad : 0), W, x0, p, s_ptab, s_btab, ph, ok);
    }
}

// Shared-memory tables of one row.  LinkT = int (k_fused_mf) or int16_t (k_fused_flow: four row contexts per SM).
// A chain link (bucket head or nxt entry) names a node as L = 4 * entry + 2 * side (side 0 = the value's lower bucket,
// 1 = its upper one), -1 ends a chain: the entry's byte offset is 2 * (L & ~3) and the node's int16 link sits at byte L
// (int links: 2 * L), so a chain step costs one mask-and-shift and one add instead of separate index arithmetic.
template <typename LinkT>
struct RowTablesT {
    uint2 *ent;   // [T]  {x = distinct right phase (float bits), y = smallest right column carrying it}
    int *head;    // [HB] bucket heads (link, -1 = empty); a bucket index is taken modulo HB (a power of two)
    LinkT *nxt;   // [2T] link of node (entry, side) at index 2 * entry + side
    int T, logT, HB;
    __device__ __forceinline__ const uint2 *entry_of(int L) const
    {
        int m;   // (opaque to the front end, which would otherwise rewrite 2 * (L & ~3) as ((L + L) & ~7): three instructions)
        asm("and.b32 %0, %1, 0xfffffffc;" : "=r"(m) : "r"(L));
        return reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(ent) + 2 * m);
    }
    __device__ __forceinline__ int next_of(int L) const
    {
        return (int)*reinterpret_cast<const LinkT *>(reinterpret_cast<const char *>(nxt) + L * (int)(sizeof(LinkT) / 2));
    }
    // One step of the walk for left phase v along a chain that may have ended already: where L names a node (L >= 0),
    // best = min(best, the node's column) if the node's phase matches v (slr::phase_match: fabsf(v - pR) < 0.1f), and
    // L = the node's link; an ended chain keeps its registers.  In PTX because the C++ form (a ? load : default)
    // materialises a default for every predicated load on every step.  The entry {phase bits, smallest column} is one
    // 64-bit load: two steps in five match and need the column anyway, and a second load of the same entry costs an
    // instruction and as many shared-memory wavefronts as the first.
    __device__ __forceinline__ void chain_step(float v, int &L, int &best, uint32_t &key, int &col) const
    {
        const uint32_t ea = slr::smem_u32(entry_of(L)), la = slr::smem_u32(nxt) + (uint32_t)(L * (int)(sizeof(LinkT) / 2));
#define SLR_CHAIN_STEP(LD_LINK)                                                                               \
    asm volatile("{\n.reg .pred p, q;\n.reg .f32 d;\n"                                                       \
                 "setp.ge.s32 p, %0, 0;\n"                                                                     \
                 "@p ld.shared.v2.u32 {%2, %3}, [%5];\n"                                                       \
                 "@p " LD_LINK " %0, [%6];\n"                                                                  \
                 "sub.rn.f32 d, %4, %2;\nabs.f32 d, d;\nsetp.lt.and.f32 q, d, 0f3DCCCCCD, p;\n"                \
                 "@q min.s32 %1, %1, %3;\n}"                                                                   \
                 : "+r"(L), "+r"(best), "+r"(key), "+r"(col)                                                   \
                 : "f"(v), "r"(ea), "r"(la))
        if (sizeof(LinkT) == 2)
            SLR_CHAIN_STEP("ld.shared.s16");
        else
            SLR_CHAIN_STEP("ld.shared.s32");
#undef SLR_CHAIN_STEP
    }
};
using RowTables = RowTablesT<int>;

// value -> min column, deduplicated, for PX right pixels of columns col0 .. col0+PX-1.  The pixels advance in lock
// step (all first probes, then all claims, then all minima, then all bucket links) so that their shared-memory
// round trips overlap.  The thread that claims a new value also files it under the bucket(s) its +-0.1 match
// window touches.
template <int PX, bool CLAMP, typename LinkT>
__device__ __forceinline__ void insert_right(const RowTablesT<LinkT> &t, const float (&ph)[PX], const bool (&ok)[PX], int col0)
{
    const int T = t.T, HB = t.HB;
    uint32_t key[PX], h[PX], cur[PX];
    int mk[PX];
    bool need[PX], claimed[PX];
#pragma unroll
    for (int q = 0; q < PX; q++) {
        key[q] = __float_as_uint(__fadd_rn(ph[q], 0.0f));  // -0 -> +0
        h[q] = slot_of(key[q], t.logT);
        // the same value one column to the left is already filed with a smaller column
        need[q] = ok[q] && !(q > 0 && ok[q > 0 ? q - 1 : 0] && key[q] == key[q > 0 ? q - 1 : 0]);
    }
#pragma unroll
    for (int q = 0; q < PX; q++) {
        const uint2 e = need[q] ? t.ent[h[q]] : make_uint2(key[q], 0u);
        cur[q] = e.x;
        mk[q] = (int)e.y;
    }
#pragma unroll
    for (int q = 0; q < PX; q++) {
        claimed[q] = false;
        if (need[q] && cur[q] == KEY_EMPTY) {
            cur[q] = atomicCAS(&t.ent[h[q]].x, KEY_EMPTY, key[q]);
            mk[q] = INT_MAX;
            claimed[q] = cur[q] == KEY_EMPTY;
            if (claimed[q]) cur[q] = key[q];
        }
    }
#pragma unroll
    for (int q = 0; q < PX; q++) {
        // collision (need[q] is set here): double hashing, odd stride.  Linear probing's longest run at the 0.6 load
        // of a row of all-distinct phases is ~75 slots, and one such lane stalls its warp for thousands of cycles.
        const uint32_t stride = ((key[q] * 0x7FEB352Du) >> 9) | 1u;
        while (cur[q] != key[q]) {
            h[q] = (h[q] + stride) & (T - 1);
            const uint2 n = t.ent[h[q]];
            cur[q] = n.x;
            mk[q] = (int)n.y;
            if (cur[q] == KEY_EMPTY) {
                cur[q] = atomicCAS(&t.ent[h[q]].x, KEY_EMPTY, key[q]);
                mk[q] = INT_MAX;
                claimed[q] = cur[q] == KEY_EMPTY;
                if (claimed[q]) cur[q] = key[q];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PX; q++)
        if (need[q] && col0 + q < mk[q]) atomicMin(reinterpret_cast<int *>(&t.ent[h[q]].y), col0 + q);
#pragma unroll
    for (int q = 0; q < PX; q++) {
        if (claimed[q]) {
            const float v = __uint_as_float(key[q]);
            const int lo = window_bucket<CLAMP>(__fsub_rn(v, 0.11f)), hi = window_bucket<CLAMP>(__fadd_rn(v, 0.11f));
            t.nxt[2 * h[q]] = (LinkT)atomicExch(&t.head[lo & (HB - 1)], (int)(4 * h[q]));
            if (hi != lo) t.nxt[2 * h[q] + 1] = (LinkT)atomicExch(&t.head[hi & (HB - 1)], (int)(4 * h[q] + 2));
        }
    }
}

// smallest right column whose phase matches v (INT_MAX = none): walk the chain of v's bucket
template <bool CLAMP, typename LinkT>
__device__ __forceinline__ int first_match(const RowTablesT<LinkT> &t, float v)
{
    int best = INT_MAX;
    int n = t.head[window_bucket<CLAMP>(v) & (t.HB - 1)];
    while (n >= 0) {
        const uint2 e = *t.entry_of(n);
        n = t.next_of(n);
        if (slr::phase_match(v, __uint_as_float(e.x))) best = min(best, (int)e.y);
    }
    return best;
}

// The same for two left pixels at once (NaN = no pixel): both chains advance in one loop, so a warp iterates
// max(len0, len1) over its lanes instead of max(len0) + max(len1), and the two walks' shared-memory round trips overlap.
template <bool CLAMP, typename LinkT>
__device__ __forceinline__ void first_match_x2(const RowTablesT<LinkT> &t, float v0, float v1, int &best0, int &best1)
{
    best0 = best1 = INT_MAX;
    int n0 = (v0 == v0) ? t.head[window_bucket<CLAMP>(v0) & (t.HB - 1)] : -1;
    int n1 = (v1 == v1) ? t.head[window_bucket<CLAMP>(v1) & (t.HB - 1)] : -1;
    uint32_t k0 = 0u, k1 = 0u;   // (scratch registers of the steps)
    int c0 = INT_MAX, c1 = INT_MAX;
    while ((n0 & n1) >= 0) {   // at least one chain has a node left (an ended chain's link is -1)
        t.chain_step(v0, n0, best0, k0, c0);
        t.chain_step(v1, n1, best1, k1, c1);
    }
}

}  // namespace slr_fused
