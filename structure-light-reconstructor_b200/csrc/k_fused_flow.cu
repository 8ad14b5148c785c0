// k_fused_flow.cu — the fused multi-frequency pipeline as a barrier-free dataflow kernel.
// Same work and the same device arithmetic as k_fused_mf (k_fused.cu: shadow mask + phase decode + heterodyne +
// per-row first-k phase match + Q triangulation, Duke/mfreconstruct.cpp:160-334); what changes is the schedule.
//
// k_fused_mf runs a row in CTA-wide phases (clear | decode + insert | query + emit) separated by __syncthreads();
// its profile shows a quarter of the resident warps parked at those barriers and the issue slots 67 % busy.
// Here one persistent CTA per SM keeps FOUR row contexts in flight and no CTA-wide barrier in steady state:
//
//   row r+1   TMA bulk copies of its image rows in flight      (stage buffer (r+1) % 2)
//   row r     decode jobs: image bytes -> phases, right phases filed in the tables, left phases parked
//             (context r % 4, stage buffer r % 2)
//   row r-1   decoded, waiting (its last decode jobs may still be running)
//   row r-2   query jobs: first-k match of 64 left pixels + Q reprojection + stores   (context (r-2) % 4)
//   row r-3   its tables being cleared by the warp that finished its last query job  (context (r-3) % 4)
//
// Work is cut into warp-sized jobs; every warp draws the next job from ONE shared counter whose order lists, per
// step, the decode jobs of row r and then the query jobs of row r-2, so table-lookup-bound, latency-bound and
// fp64-bound instruction streams share the SM at all times.  A job waits only on per-context mbarriers that jobs
// preceding it in the global order arrive on, which makes the schedule deadlock-free: decode(r) needs the clear after
// query(r-4) and the stage re-armed after decode(r-2); query(r) needs decode(r).  A waiting warp sleeps in
// mbarrier.try_wait (no issue slots, no shared-memory polling).  The warp whose decode job is the last to have read a
// row's stage buffer re-arms it with the TMA copies of row r+2.
//
// RAW = true (slr_run_mf_raw): the stack holds the RAW camera images and the stage buffer is not
// filled by bulk copies but by a third kind of job, listed one step ahead of the row's decode jobs: rectify jobs
// evaluate stereoRect::doStereoRectify = cv::remap(INTER_LINEAR, CV_16SC2 maps) (Duke/stereorect.cpp:26-34) for 128
// pixels of one camera's row and all N planes (slr::rectify_job, slr_rectify.cuh — the code the stand-alone K0 runs too)
// straight into the stage buffer, so the rectified images never exist in HBM: raw bytes in, XYZ out, one kernel.  Rows
// are walked image-row-fastest there, so that the two source rows an output row blends are still in L2 for the next
// output row.
#include <limits.h>
#include <stdlib.h>

#include "k_fused_common.cuh"
#include "slr_rectify.cuh"

namespace {

using namespace slr_fused;

constexpr int FLOW_CTX = 4;        // row contexts (tables + parked left phases); a power of two
constexpr int FLOW_LAG = 2;        // steps between a row's decode jobs and its query jobs
constexpr int FLOW_STAGES = 2;     // TMA stage buffers
constexpr int FLOW_QPX = 2;        // left pixels per lane in one query job
constexpr int FLOW_HEADER = 256;   // mbarriers, counters, row descriptors
constexpr int FLOW_RAW_LEAD = 3;   // RAW: rows between the L2 prefetch of a source row and the row that blends it
constexpr int FLOW_CAM_PAD = 64;   // bytes between the two cameras' planes in a stage buffer: N * W is a multiple of 128
                                   // for the usual widths, and the left and the right lane of a pair read the same
                                   // columns, i.e. the same banks, of their cameras in one load

using FlowTables = RowTablesT<int16_t>;

#ifdef SLR_FLOW_TRACE   // debug build (csrc/Makefile TRACE=1): per-job clock stamps of CTA 0, read by profiles/flow_trace.py
constexpr int TRACE_JOBS = 8192;
__device__ long long g_flow_trace[TRACE_JOBS * 4];   // {type | row << 8 | warp << 40, t_draw, t_ready, t_end}
#define FLOW_TRACE_DRAW() const long long trc_draw = clock64()
#define FLOW_TRACE_READY() const long long trc_ready = clock64()
#define FLOW_TRACE_END(type, row)                                                                               \
    do {                                                                                                        \
        if (blockIdx.x == 0 && lane == 0 && g < TRACE_JOBS) {                                                   \
            g_flow_trace[4 * g + 0] = (long long)(type) | ((long long)(row) << 8) | ((long long)(tid >> 5) << 40); \
            g_flow_trace[4 * g + 1] = trc_draw;                                                                 \
            g_flow_trace[4 * g + 2] = trc_ready;                                                                \
            g_flow_trace[4 * g + 3] = clock64();                                                                \
        }                                                                                                       \
    } while (0)
#else
#define FLOW_TRACE_DRAW() do { } while (0)
#define FLOW_TRACE_READY() do { } while (0)
#define FLOW_TRACE_END(type, row) do { } while (0)
#endif

// The single-lane steps of a job: lane 0 arrives (release) on an mbarrier that counts jobs and / or bumps a counter for
// its warp.  counter += 1 returns the old value and elects the job that finishes a row's stage reads / queries last;
// the ORDERING between the jobs' shared-memory accesses and what the elected job does next (bulk copies into the stage,
// the table clear) comes from the mbarrier every job arrives on (release) and the elected job waits on (acquire) — an
// acq_rel atomic costs a MEMBAR.ALL.CTA that also waits for the job's global stores.  (atom.inc, not atom.add: ptxas
// wraps an add of a warp-uniform operand into its warp-aggregation sequence — vote, popc, lane masks — which costs more
// than the atomic when a single lane executes it.)  Written as predicated PTX rather than `if (lane == 0)`: no default
// value to materialise for the other lanes and one convergence region where a job does both (1.7 % on rectified input).
// Results are defined on lane 0 only.  The RAW instantiations keep the branch form below: with their rectify jobs they
// sit at the 64-register limit, and the predicated form costs them two more spilled registers (4.31 vs 4.16 ms).
__device__ __forceinline__ int count_job(int *p)
{
    unsigned old;
    asm volatile("atom.relaxed.cta.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "r"(slr::smem_u32(p)) : "memory");
    return (int)old;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(slr::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_lane0(int lane, uint64_t *bar)
{
    asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %1, 0;\n@p mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n}"
                 ::"r"(slr::smem_u32(bar)), "r"(lane) : "memory");
}
__device__ __forceinline__ int count_lane0(int lane, int *ctr)
{
    unsigned old;
    asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %2, 0;\n@p atom.relaxed.cta.shared.inc.u32 %0, [%1], 0x7fffffff;\n}"
                 : "=r"(old) : "r"(slr::smem_u32(ctr)), "r"(lane) : "memory");
    return (int)old;
}
__device__ __forceinline__ int arrive_count_lane0(int lane, uint64_t *bar, int *ctr)   // both, in this order
{
    unsigned old;
    asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %3, 0;\n@p mbarrier.arrive.release.cta.shared::cta.b64 _, [%1];\n"
                 "@p atom.relaxed.cta.shared.inc.u32 %0, [%2], 0x7fffffff;\n}"
                 : "=r"(old) : "r"(slr::smem_u32(bar)), "r"(slr::smem_u32(ctr)), "r"(lane) : "memory");
    return (int)old;
}

// shared-memory bytes of one row context: ent[T] (8 B) + head[T] (4 B) + nxt[2T] (2 B) + left phases [W] (4 B)
__host__ __device__ inline size_t flow_ctx_bytes(int W, int T) { return (size_t)16 * T + (size_t)4 * W; }

// first pixel of a row in the outputs / in the undistort maps, written by the warp that issues the row's TMA copies
struct RowInfo {
    unsigned out_px, map_px;
};

// WCT > 0: the row width (and with it the table size and, in strict mode, the plane count) as compile-time constants —
// plane loads become [base + immediate], bounds tests fold away.  Instantiated for the reference camera's 1280 pixels.
template <int MODE, bool RAW, int WCT>
__global__ void __launch_bounds__(1024, 1)
k_fused_flow(const FusedParams p, const int n_r_rt, const int n_d_rt, const int n_q_rt, const unsigned js_magic)
{
    // jobs per row: rectify (RAW), decode, query
    const int n_r = WCT ? (RAW ? 2 * ((WCT + 127) / 128) : 0) : n_r_rt;
    const int n_d = WCT ? (WCT / 2 + 31) / 32 : n_d_rt;
    const int n_q = WCT ? (WCT + 32 * FLOW_QPX - 1) / (32 * FLOW_QPX) : n_q_rt;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool CLAMP = MODE == MODE_PHASE_INPUT;
    const int W = WCT ? WCT : p.W, N = (WCT && MODE == SLR_MODE_STRICT) ? 14 : p.N, T = (WCT == 1280) ? 2048 : p.T;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;

    // this CTA's contiguous range of rows rg = i * batch + b (scan index b fastest; RAW: rg = b * H + i)
    // (batch * H < 2^31, checked by the launcher: row indices are 32-bit, and so are the divisions that split them)
    const long long rows = (long long)p.batch * p.H;
    const unsigned r_begin = (unsigned)(rows * blockIdx.x / gridDim.x), r_end = (unsigned)(rows * (blockIdx.x + 1) / gridDim.x);
    const int R = (int)(r_end - r_begin);
    if (R <= 0) return;

    // ---- shared memory carve-up ----
    uint64_t *bar_stage = reinterpret_cast<uint64_t *>(smem);              // [FLOW_STAGES] a row's bulk copies have landed
    uint64_t *bar_dec = bar_stage + FLOW_STAGES;                           // [FLOW_CTX] phase u: row 4u+c is decoded (n_d arrivals)
    uint64_t *bar_clr = bar_dec + FLOW_CTX;                                // [FLOW_CTX] phase u: tables cleared after row 4u+c
    uint64_t *bar_free = reinterpret_cast<uint64_t *>(smem + 192);         // [FLOW_STAGES] a row's decode jobs have read the stage
    uint64_t *bar_qry = bar_free + FLOW_STAGES;                            // [FLOW_CTX] a row's query jobs are done with its tables
    int *job_ctr = reinterpret_cast<int *>(bar_clr + FLOW_CTX);
    int *done_q = job_ctr + 1;                                             // [FLOW_CTX] query jobs completed
    int *done_l = done_q + FLOW_CTX;                                       // [FLOW_CTX] decode jobs done reading the stage
    RowInfo *rowinfo = reinterpret_cast<RowInfo *>(smem + 128);            // [8] ring, indexed by row & 7
    const size_t tx_bytes = (MODE == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * N * W;   // bytes a row's copies bring
    const int cam_pad = (MODE == MODE_PHASE_INPUT) ? 0 : FLOW_CAM_PAD;
    const size_t stage_bytes = tx_bytes + cam_pad;
    unsigned char *stage0 = smem + FLOW_HEADER;
    unsigned char *ctx0 = stage0 + FLOW_STAGES * stage_bytes;
    const unsigned ctx_bytes = (unsigned)flow_ctx_bytes(W, T);
    int *s_ptab = reinterpret_cast<int *>(ctx0 + (size_t)FLOW_CTX * ctx_bytes);  // [SLR_PTAB_SIZE] (strict)
    uint32_t *s_btab = reinterpret_cast<uint32_t *>(s_ptab + SLR_PTAB_SIZE);     // [SLR_BTAB_SIZE] (strict)

    auto tables_of = [&](int c, float *&s_pl) {
        FlowTables t;
        unsigned char *base = ctx0 + (unsigned)c * ctx_bytes;
        t.T = T;
        t.logT = (WCT == 1280) ? 11 : p.logT;
        t.HB = T;
        t.ent = reinterpret_cast<uint2 *>(base);
        t.head = reinterpret_cast<int *>(t.ent + T);
        t.nxt = reinterpret_cast<int16_t *>(t.head + T);
        s_pl = reinterpret_cast<float *>(t.nxt + 2 * T);
        return t;
    };
    // a context's tables: keys = EMPTY, min column = INT_MAX, heads = -1
    auto clear_tables = [&](int c, int first, int stride) {
        float *unused;
        const FlowTables t = tables_of(c, unused);
        uint4 *e4p = reinterpret_cast<uint4 *>(t.ent);   // two entries per vector
        uint4 *h4 = reinterpret_cast<uint4 *>(t.head);   // four heads per vector
        const uint4 e4 = make_uint4(KEY_EMPTY, 0x7fffffffu, KEY_EMPTY, 0x7fffffffu);
        const uint4 m4 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        for (int q = first; q < (T >> 1); q += stride) e4p[q] = e4;
        for (int q = first; q < (T >> 2); q += stride) h4[q] = m4;
    };

    const int JS = n_r + n_d + n_q;
    if (MODE == SLR_MODE_STRICT) {
        for (int k = tid; k < SLR_PTAB_SIZE; k += nthr) s_ptab[k] = p.ptab[k];
        for (int k = tid; k < SLR_BTAB_SIZE; k += nthr) s_btab[k] = p.btab[k];
    }
    for (int c = 0; c < FLOW_CTX; c++) clear_tables(c, tid, nthr);
    if (tid < 1 + 2 * FLOW_CTX) job_ctr[tid] = 0;
    if (tid == 0) {
        for (int s = 0; s < FLOW_STAGES; s++) slr::mbar_init(&bar_stage[s], RAW ? (uint32_t)n_r : 1u);
        for (int s = 0; s < FLOW_STAGES; s++) slr::mbar_init(&bar_free[s], (uint32_t)n_d);
        for (int c = 0; c < FLOW_CTX; c++) slr::mbar_init(&bar_qry[c], (uint32_t)n_q);
        for (int c = 0; c < FLOW_CTX; c++) slr::mbar_init(&bar_dec[c], (uint32_t)n_d), slr::mbar_init(&bar_clr[c], 1);
        slr::mbar_fence_init();
    }
    const uint64_t policy = make_evict_first_policy();
    __syncthreads();

    // TMA bulk copies of row r (CTA-local index) into stage r % FLOW_STAGES, issued by one whole warp
    auto issue_row = [&](int r) {
        const unsigned rg = r_begin + (unsigned)r;
        const int i = (int)(rg / (unsigned)p.batch), b = (int)(rg - (unsigned)i * (unsigned)p.batch);
        uint64_t *bar = &bar_stage[r % FLOW_STAGES];
        unsigned char *stage = stage0 + (size_t)(r % FLOW_STAGES) * stage_bytes;
        if (lane == 0) {
            rowinfo[r & 7].out_px = (unsigned)(((size_t)b * p.H + i) * W);
            rowinfo[r & 7].map_px = (unsigned)((size_t)i * W);
            slr::mbar_expect_tx(bar, (uint32_t)tx_bytes);   // release: the descriptor is visible to every waiter
        }
        __syncwarp();
        if (MODE == MODE_PHASE_INPUT) {  // stage = pL f32[W] | pR f32[W] | mL u8[W] | mR u8[W]
            const size_t offL = ((size_t)(b * 2 + 0) * p.H + i) * W, offR = ((size_t)(b * 2 + 1) * p.H + i) * W;
            if (lane == 0) tma_load_1d_hint(stage, p.phase + offL, 4u * W, bar, policy);
            if (lane == 1) tma_load_1d_hint(stage + 4 * W, p.phase + offR, 4u * W, bar, policy);
            if (lane == 2) tma_load_1d_hint(stage + 8 * W, p.mask + offL, (uint32_t)W, bar, policy);
            if (lane == 3) tma_load_1d_hint(stage + 9 * W, p.mask + offR, (uint32_t)W, bar, policy);
            return;
        }
        const uint8_t *src = p.stack + ((size_t)b * 2 * N * p.H + i) * W;
        for (int v = lane; v < 2 * N; v += 32)   // plane v of this scan (cam-major, then image index)
            tma_load_1d_hint(stage + (size_t)v * W + (v >= N ? cam_pad : 0), src + (size_t)v * p.H * W, (uint32_t)W, bar, policy);
        // (An L2 prefetch of the row after next used to follow here; with the mbarrier schedule the bulk copies are issued
        // a full step before their row is decoded and the prefetch only cost issue slots: 1 % faster without it.)
    };
    if (!RAW && tid < 32) {
        issue_row(0);
        if (R > 1) issue_row(1);
    }

    const int ntasks = W >> 1;   // 4-pixel chunks, right and left chunk of the same columns on neighbouring lanes
    const int total_jobs = (R + 1 + FLOW_LAG) * JS;
    unsigned n_local = 0;

    // Step t lists the n_d decode jobs of row t - 1, then the n_q query jobs of row t - 1 - FLOW_LAG: a job takes most
    // of a step from draw to completion (8 warps share a scheduler), so a row's queries are drawn two steps after
    // its decode jobs and practically never wait.
    for (;;) {
        // (atom.inc in plain PTX: atomicAdd / atom.add under `lane == 0` becomes a 14-instruction warp-aggregation sequence)
        int g = 0;
        if (RAW) {
            if (lane == 0) g = count_job(job_ctr);
        } else {
            g = count_lane0(lane, job_ctr);
        }
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= total_jobs) break;
        FLOW_TRACE_DRAW();
        const int t = (int)__umulhi((unsigned)g, js_magic);   // g / JS (exact for g * JS < 2^32, checked by the launcher)
        int s = g - t * JS;
        // (Alternating the job kinds inside a step instead of listing them in blocks was measured: 5 % slower for
        // rectified input, 16 % slower for raw input — a row's jobs of one kind share table lines and stage rows.)

        if (RAW && s < n_r) {
            // ================= rectify job s of row r = t: 128 pixels of one camera, all planes =================
            const int r = t;
            if (r >= R) continue;
            // the stage buffer is free once every decode job of row r - 2 has read its bytes
            if (r >= FLOW_STAGES) slr::mbar_wait(&bar_free[r % FLOW_STAGES], (uint32_t)((r / FLOW_STAGES - 1) & 1));
            FLOW_TRACE_READY();
            const unsigned rg = r_begin + (unsigned)r;
            const int b = (int)(rg / (unsigned)p.H), i = (int)(rg - (unsigned)b * (unsigned)p.H);
            const int per_cam = n_r >> 1, cam = s >= per_cam ? 1 : 0;
            const int x = ((s - cam * per_cam) * 32 + lane) * 4;
            if (s == 0 && lane == 0) {
                rowinfo[r & 7].out_px = (unsigned)(((size_t)b * p.H + i) * W);
                rowinfo[r & 7].map_px = (unsigned)((size_t)i * W);
            }
            {
                const size_t P = (size_t)W * p.H;
                unsigned char *dst = stage0 + (size_t)(r % FLOW_STAGES) * stage_bytes + (size_t)cam * ((size_t)N * W + cam_pad) + x;
                slr::rectify_job<FLOW_RAW_LEAD>(p.stack + ((size_t)b * 2 + cam) * N * P, p.map1 + (size_t)cam * P,
                                                p.map2 + (size_t)cam * P, W, p.H, N, i, x, x < W, dst, (size_t)W, lane);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_stage[r % FLOW_STAGES]);
            FLOW_TRACE_END(3, r);
            continue;
        }
        s -= n_r;
        if (s < n_d) {
            // ================= decode job s of row r = t - 1 =================
            const int r = t - 1;
            if (r < 0 || r >= R) continue;
            const int c = r & (FLOW_CTX - 1), u = r / FLOW_CTX;
            // this row's tables were cleared by the last query job of row r - FLOW_CTX (first rows: by the prologue)
            if (u > 0) slr::mbar_wait(&bar_clr[c], (uint32_t)((u - 1) & 1));
            slr::mbar_wait(&bar_stage[r % FLOW_STAGES], (uint32_t)((r / FLOW_STAGES) & 1));
            FLOW_TRACE_READY();
            float *s_pl;
            const FlowTables tab = tables_of(c, s_pl);
            const unsigned char *stage = stage0 + (size_t)(r % FLOW_STAGES) * stage_bytes;
            {
                const int task = s * 32 + lane;
                const bool live = task < ntasks;
                const int x0 = live ? 4 * (task >> 1) : 0;
                const bool right = (task & 1) == 0;
                float ph[4];
                bool ok[4];
                load_phases<MODE, 4>(stage, W, N, x0, right, p, s_ptab, s_btab, ph, ok, cam_pad);
                // this job has read its bytes of the stage buffer (arrival = release of the loads).  The warp that counts
                // the last reader streams row r + 2 into the buffer: a third of a decode job earlier than its completion,
                // which is the slack the bulk copies need to land before row r + 2 is drawn.
                __syncwarp();
                if (RAW) {
                    if (lane == 0) mbar_arrive(&bar_free[r % FLOW_STAGES]);
                } else {
                    const int readers = __shfl_sync(0xffffffffu, arrive_count_lane0(lane, &bar_free[r % FLOW_STAGES], &done_l[c]), 0);
                    if (readers + 1 == (u + 1) * n_d && r + FLOW_STAGES < R) {
                        slr::mbar_wait(&bar_free[r % FLOW_STAGES], (uint32_t)((r / FLOW_STAGES) & 1));   // acquire: every reader's loads
                        issue_row(r + FLOW_STAGES);
                    }
                }
                // the right lane hands its upper two phases to the left lane: every lane decodes 4 pixels and files 2.
                // A pixel without a phase travels as NaN (a decoded phase never is one; a caller-supplied NaN on the
                // right is no phase either, load_phases), so the masks need no shuffle of their own.
#pragma unroll
                for (int q = 0; q < 4; q++) ph[q] = ok[q] ? ph[q] : slr::qnan();
                const float n2 = __shfl_xor_sync(0xffffffffu, ph[2], 1), n3 = __shfl_xor_sync(0xffffffffu, ph[3], 1);
                float ip[2] = {right ? ph[0] : n2, right ? ph[1] : n3};
                bool io[2] = {live && ip[0] == ip[0], live && ip[1] == ip[1]};
                insert_right<2, CLAMP>(tab, ip, io, right ? x0 : x0 + 2);
                if (live && !right) reinterpret_cast<float4 *>(s_pl)[x0 >> 2] = make_float4(ph[0], ph[1], ph[2], ph[3]);
            }
            __syncwarp();
            if (RAW) {
                if (lane == 0) mbar_arrive(&bar_dec[c]);
            } else {
                mbar_arrive_lane0(lane, &bar_dec[c]);
            }
            FLOW_TRACE_END(0, r);
        } else {
            // ================= query + emit job (FLOW_QPX * 32 left pixels) of row r = t - 1 - FLOW_LAG =================
            // (measured with the trace build, profiles/flow_trace.py: jobs wait 5 % of their time in this order; drawing
            // half of a row's queries half a step earlier made the ring's slack even on paper and the kernel 5 % slower)
            const int r = t - 1 - FLOW_LAG;
            const int qidx = s - n_d;
            if (r < 0 || r >= R) continue;
            const int c = r & (FLOW_CTX - 1), u = r / FLOW_CTX;
            slr::mbar_wait(&bar_dec[c], (uint32_t)(u & 1));
            FLOW_TRACE_READY();
            const RowInfo ri = rowinfo[r & 7];
            float *s_pl;
            const FlowTables tab = tables_of(c, s_pl);
            int j[FLOW_QPX], best[FLOW_QPX];
            float ulx[FLOW_QPX], uly[FLOW_QPX], v[FLOW_QPX];
#pragma unroll
            for (int q = 0; q < FLOW_QPX; q++) {
                j[q] = (qidx * FLOW_QPX + q) * 32 + lane;
                const bool inside = j[q] < W;
                // undistortPoints maps of the left pixel (L2-resident, coalesced): in flight during the table walk
                // (one 32-bit pixel index per load, widened once: pointer + unsigned + int is two 64-bit additions)
                ulx[q] = inside ? __ldg(p.lx + (ri.map_px + (unsigned)j[q])) : 0.0f;
                uly[q] = inside ? __ldg(p.ly + (ri.map_px + (unsigned)j[q])) : 0.0f;
                v[q] = inside ? s_pl[j[q]] : slr::qnan();
            }
            static_assert(FLOW_QPX == 2, "the query job walks two chains per lane");
            first_match_x2<CLAMP>(tab, v[0], v[1], best[0], best[1]);
            float d[FLOW_QPX], X[FLOW_QPX], Y[FLOW_QPX], Z[FLOW_QPX];
#pragma unroll
            for (int q = 0; q < FLOW_QPX; q++) {
                // every pixel is reprojected unconditionally; misses get harmless inputs (disparity 1) and become NaN
                const bool hit = best[q] != INT_MAX;
                d[q] = 1.0f;
                if (hit)
                    d[q] = __fsub_rn(ulx[q], __ldg(p.rx + (ri.map_px + (unsigned)best[q])));
                else
                    ulx[q] = 0.0f, uly[q] = 0.0f;
            }
#pragma unroll
            for (int q = 0; q < FLOW_QPX; q++)
                slr::reproject_q(p.calib, (double)ulx[q], (double)uly[q], (double)d[q], X[q], Y[q], Z[q]);
#pragma unroll
            for (int q = 0; q < FLOW_QPX; q++) {
                const bool hit = best[q] != INT_MAX;
                n_local += hit ? 1u : 0u;
                const float ox = hit ? X[q] : slr::qnan(), oy = hit ? Y[q] : slr::qnan(), oz = hit ? Z[q] : slr::qnan();
                const unsigned o = ri.out_px + (unsigned)j[q];   // pixel offsets fit 32 bits (checked by the launcher)
                if (p.n_t == 1) {
                    if (j[q] < W) {
                        float *dst = p.xyz + (size_t)o * 3;
                        dst[0] = ox;
                        dst[1] = oy;
                        dst[2] = oz;
                        p.valid[o] = hit ? 1 : 0;
                    }
                } else {
                    // Assembly on every GPU: target 0 is this GPU's cloud, the others are the peers' (NVLink stores
                    // straight from registers, so the all-gather overlaps the kernel that produces the data).  The 32
                    // float3 of the group are transposed across the warp (4 x select + shuffle) into 24 float4, so
                    // that every peer store is a full 16-byte write instead of three 4-byte writes 12 bytes apart:
                    // a third of the NVLink packets, each with a full payload.
                    float4 v4;
                    float *o4 = &v4.x;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int cmp = (k - 3 * lane) & 3;            // the component lane `lane` exposes in round k
                        const float sel = (cmp == 0) ? ox : (cmp == 1) ? oy : oz;
                        o4[k] = __shfl_sync(0xffffffffu, sel, min((4 * lane + k) / 3, 31));
                    }
                    const int j0 = j[q] - lane, n_f = 3 * min(32, W - j0);   // floats of this group inside the row
                    const size_t f0 = ((size_t)ri.out_px + j0) * 3;
#pragma unroll 1
                    for (int tg = 0; tg < p.n_t; tg++) {
                        if (4 * lane + 3 < n_f) *reinterpret_cast<float4 *>(p.xyz_t[tg] + f0 + 4 * lane) = v4;
                        if (j[q] < W) p.valid_t[tg][o] = hit ? 1 : 0;
                    }
                }
                if (j[q] < W && p.match_k) p.match_k[o] = hit ? best[q] : -1;
            }
            __syncwarp();
            int last = 0;
            if (RAW) {
                if (lane == 0) {
                    mbar_arrive(&bar_qry[c]);
                    last = count_job(&done_q[c]) + 1 == (u + 1) * n_q;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
            } else {
                last = __shfl_sync(0xffffffffu, arrive_count_lane0(lane, &bar_qry[c], &done_q[c]), 0) + 1 == (u + 1) * n_q;
            }
            if (last) {  // nobody reads this row's tables any more: clear them for row r + FLOW_CTX
                slr::mbar_wait(&bar_qry[c], (uint32_t)(u & 1));   // acquire: every query job's table reads
                clear_tables(c, lane, 32);
                __syncwarp();
                if (RAW) {
                    if (lane == 0) mbar_arrive(&bar_clr[c]);
                } else {
                    mbar_arrive_lane0(lane, &bar_clr[c]);
                }
            }
            FLOW_TRACE_END(last ? 2 : 1, r);
        }
    }
    if (p.n_points) {
        const unsigned long long s = slr::warp_sum_u32(n_local);
        if (lane == 0 && s) atomicAdd(p.n_points, s);
    }
}

}  // namespace

// Launch the dataflow kernel when the shape fits (three row contexts + two stage buffers in one SM's shared memory);
// *handled = false sends the caller to k_fused_mf / the un-fused kernels.
slr_status slr_launch_fused_flow(slr_engine *e, int mode, const FusedParams &p_in, bool *handled)
{
    *handled = false;
    if (const char *ev = getenv("SLR_FUSED_FLOW"))
        if (atoi(ev) == 0) return SLR_OK;
    FusedParams p = p_in;
    const int W = p.W;
    const size_t stage_bytes = (mode == MODE_PHASE_INPUT) ? (size_t)10 * W : (size_t)2 * p.N * W + FLOW_CAM_PAD;
    const size_t smem = FLOW_HEADER + FLOW_STAGES * stage_bytes + FLOW_CTX * flow_ctx_bytes(W, p.T) +
                        SLR_PTAB_SIZE * 4 + SLR_BTAB_SIZE * 4;
    const bool raw = p.map1 != nullptr;   // the stack holds raw camera images: rectify jobs fill the stage buffer
    const int n_d = (W / 2 + 31) / 32, n_q = (W + 32 * FLOW_QPX - 1) / (32 * FLOW_QPX);
    const int n_r = raw ? 2 * ((W + 127) / 128) : 0;
    if (smem > 227 * 1024 || 4 * p.T > 32768 || (stage_bytes % 16) != 0) return SLR_OK;   // (16-bit links hold 4 * entry + 2)
    if (raw && (mode == MODE_PHASE_INPUT || p.calib.row0 != 0)) return SLR_OK;
    // rows per CTA must keep the job counter inside int32, pixel offsets inside uint32
    const long long rows = (long long)p.batch * p.H;
    const long long JS = n_r + n_d + n_q;
    if ((rows / e->num_sms + 8) * JS * JS >= (1LL << 32) || rows * W >= (1LL << 32)) return SLR_OK;
    const unsigned js_magic = (unsigned)(((1ULL << 32) + JS - 1) / JS);
    *handled = true;

    // one CTA per SM, up to 32 warps; narrow rows (few jobs per step) run two smaller CTAs per SM when they fit
    const int ctas = (2 * (smem + 1024) <= 228 * 1024) ? 2 : 1;
    int warps = 32 / ctas;
    while (warps > 2 && warps > n_d + n_q) warps >>= 1;   // (n_r + n_d >= n_d + n_q: a rectify job covers 128 pixels, a query job 64)
    // (the mbarrier phases stay unambiguous only while a CTA has no more warps than one step has jobs: a stuck row then
    // blocks every warp before any of them can reach the jobs of the row that reuses its context)
    if (const char *ev = getenv("SLR_FLOW_WARPS")) {
        const int w = atoi(ev);
        if (w >= 1 && w <= 32 / ctas && w <= (n_d + n_q > 2 ? n_d + n_q : 2)) warps = w;
    }
    void (*kern)(const FusedParams, int, int, int, unsigned);
    if (raw)
        kern = (mode == SLR_MODE_STRICT) ? k_fused_flow<SLR_MODE_STRICT, true, 0> : k_fused_flow<SLR_MODE_CORRECTED, true, 0>;
    else
        kern = (mode == SLR_MODE_STRICT)      ? k_fused_flow<SLR_MODE_STRICT, false, 0>
               : (mode == SLR_MODE_CORRECTED) ? k_fused_flow<SLR_MODE_CORRECTED, false, 0>
                                              : k_fused_flow<MODE_PHASE_INPUT, false, 0>;
    if (W == 1280 && p.T == 2048 && !getenv("SLR_FLOW_GENERIC_W")) {   // the reference camera's width, specialised
        if (mode == SLR_MODE_STRICT && p.N == 14) kern = raw ? k_fused_flow<SLR_MODE_STRICT, true, 1280> : k_fused_flow<SLR_MODE_STRICT, false, 1280>;
        if (mode == SLR_MODE_CORRECTED) kern = raw ? k_fused_flow<SLR_MODE_CORRECTED, true, 1280> : k_fused_flow<SLR_MODE_CORRECTED, false, 1280>;
    }
    SLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (long long)e->num_sms * ctas;
    if (grid > rows) grid = rows;
    if (grid < 1) return SLR_OK;
    kern<<<(unsigned)grid, warps * 32, smem, e->stream>>>(p, n_r, n_d, n_q, js_magic);
    SLR_CHECK_LAUNCH(e);
#ifdef SLR_FLOW_TRACE
    if (const char *path = getenv("SLR_FLOW_TRACE_OUT")) {
        static long long h[TRACE_JOBS * 4];
        SLR_CHECK_CUDA(cudaStreamSynchronize(e->stream));
        SLR_CHECK_CUDA(cudaMemcpyFromSymbol(h, g_flow_trace, sizeof(h)));
        if (FILE *f = fopen(path, "wb")) {
            int hdr[4] = {TRACE_JOBS, n_d, n_q, warps};
            fwrite(hdr, sizeof(hdr), 1, f);
            fwrite(h, sizeof(h), 1, f);
            fclose(f);
        }
    }
#endif
    return SLR_OK;
}
