// TMA-staged message passing for small formulas (bf16 tensor-core path): persistent CTAs, one per SM,
// loop over (chain, feature-slice) work items.  The two tables an item gathers from are brought into
// shared memory by cp.async.bulk.tensor (dense boxes, one mbarrier per buffer) while the previous
// item is being computed (double buffering), so every table row is read from HBM exactly once and all
// per-edge reads hit shared memory.
//   clause side : tables LIT[n][2Q] and SP[n][2Q]      -> CROW[:, F:F+2Q] = clause_messages | 4*clauses_loss
//   literal side: tables CL4[m][W] and MSG[m][W]       -> VROW[:, F+16:]  = variables_grad | loss_pos | loss_neg
#pragma once
#include "dsat_gemm_tc.cuh"
#include "dsat_message.cuh"

namespace dsat {
namespace tg {

constexpr int THREADS = 1024;

struct GatherPlan {
    int width;          // feature slice width W (columns per item)
    int slices;         // Q / W
    int box_rows;       // rows per TMA box (<= 256)
    int boxes;          // boxes per table
    int table_bytes;    // bytes of one table buffer (boxes * box_rows * W * 2)
    int smem_bytes;     // 2 buffers x 2 tables + barriers
};

inline bool plan_gather(int table_rows, int Q, int max_width, GatherPlan* g) {
    const int widths[3] = {256, 128, 64};
    for (int w : widths) {
        if (w > max_width || w > Q || Q % w) continue;
        const int boxes = (table_rows + 255) / 256;
        const int box_rows = (table_rows + boxes - 1) / boxes;
        const int table_bytes = boxes * box_rows * w * 2;
        const int total = 4 * table_bytes + 64;
        if (total <= 226 * 1024) {
            g->width = w; g->slices = Q / w; g->box_rows = box_rows; g->boxes = boxes;
            g->table_bytes = table_bytes; g->smem_bytes = total + 128;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ void issue_tables(const CUtensorMap* m0, const CUtensorMap* m1, uint8_t* buf, uint64_t* bar,
                                             const GatherPlan& gp, int col, long long row0) {
    using namespace tc;
    const int box_bytes = gp.box_rows * gp.width * 2;
    mbar_expect_tx(bar, (uint32_t)(2 * gp.boxes * box_bytes));
    for (int b = 0; b < gp.boxes; ++b) {
        tma_load_2d(buf + (size_t)b * box_bytes, m0, bar, col, (int)(row0 + (long long)b * gp.box_rows));
        tma_load_2d(buf + gp.table_bytes + (size_t)b * box_bytes, m1, bar, col, (int)(row0 + (long long)b * gp.box_rows));
    }
}

// ---------------------------------------------------------------------------------- clause side
// map_lit: LITb viewed as [N rows, 2Q cols]; map_sp: QSb[:, Q:3Q] viewed as [N rows, 2Q cols]; one item = one chain,
// width = 2Q (both literal signs of every variable): table row v holds [positive Q | negative Q] = codes 2v, 2v+1.
template <int VS>
__global__ void __launch_bounds__(THREADS, 1)
clause_gather_tma_kernel(const __grid_constant__ CUtensorMap map_lit, const __grid_constant__ CUtensorMap map_sp,
                         UnitGraphDev g, int chains, GatherPlan gp, __nv_bfloat16* __restrict__ OUT, int ld_out, int out_off) {
    using namespace tc;
    using T = __nv_bfloat16;
    constexpr int Q = 32 * VS;
    extern __shared__ uint8_t tsm_raw[];
    uint8_t* tsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm_raw) + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(tsm + 4 * (size_t)gp.table_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = THREADS / 32;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int items = chains;
    int k = 0;
    if (tid == 0 && blockIdx.x < items)
        issue_tables(&map_lit, &map_sp, tsm, &full[0], gp, 0, (long long)blockIdx.x * g.n);
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int buf = k & 1;
        if (tid == 0 && item + (int)gridDim.x < items)
            issue_tables(&map_lit, &map_sp, tsm + (size_t)(buf ^ 1) * 2 * gp.table_bytes, &full[buf ^ 1], gp, 0,
                         (long long)(item + gridDim.x) * g.n);
        mbar_wait(&full[buf], (uint32_t)((k >> 1) & 1));
        const T* t_lit = reinterpret_cast<const T*>(tsm + (size_t)buf * 2 * gp.table_bytes);
        const T* t_sp = reinterpret_cast<const T*>(tsm + (size_t)buf * 2 * gp.table_bytes + gp.table_bytes);
        for (int j = warp; j < g.m; j += nwarps) {
            const int e0 = __ldg(g.cl_rowptr + j), e1 = __ldg(g.cl_rowptr + j + 1);
            LaneVec<VS> acc_l, acc_s;
#pragma unroll
            for (int i = 0; i < VS; ++i) { acc_l.v[i] = 0.f; acc_s.v[i] = 0.f; }
            for (int e = e0; e < e1; ++e) {
                const int code = __ldg(g.cl_lit + e);
                LaneVec<VS> l0 = lane_load_rw_t<VS, T>(t_lit + (size_t)code * Q, lane);
                LaneVec<VS> s0 = lane_load_rw_t<VS, T>(t_sp + (size_t)code * Q, lane);
#pragma unroll
                for (int i = 0; i < VS; ++i) { acc_l.v[i] += l0.v[i]; acc_s.v[i] += s0.v[i]; }
            }
            const float rw = __ldg(g.rev_w + j);
#pragma unroll
            for (int i = 0; i < VS; ++i) {
                acc_l.v[i] *= rw;
                acc_s.v[i] = 4.0f * __expf(-acc_s.v[i]);
            }
            T* dst = OUT + ((size_t)item * g.m + j) * ld_out + out_off;
            lane_store_t<VS, T>(dst, lane, acc_l);
            lane_store_t<VS, T>(dst + Q, lane, acc_s);
        }
        __syncthreads();          // everyone is done with this buffer before it is refilled two items later
    }
}

// --------------------------------------------------------------------------------- literal side
// map_cl: CROWb[:, F+Q:] viewed as [M rows, Q cols]; map_ms: COUTb[:, :Q] viewed as [M rows, Q cols];
// item = (chain, slice of W = 32*VS features).
template <int VS>
__global__ void __launch_bounds__(THREADS, 1)
literal_gather_tma_kernel(const __grid_constant__ CUtensorMap map_cl, const __grid_constant__ CUtensorMap map_ms,
                          UnitGraphDev g, int chains, int Q, GatherPlan gp,
                          const __nv_bfloat16* __restrict__ QRY, int ld_q,
                          __nv_bfloat16* __restrict__ OUT, int ld_out, int out_off) {
    using namespace tc;
    using T = __nv_bfloat16;
    constexpr int W = 32 * VS;
    extern __shared__ uint8_t tsm_raw[];
    uint8_t* tsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tsm_raw) + 127) & ~(uintptr_t)127);
    uint64_t* full = reinterpret_cast<uint64_t*>(tsm + 4 * (size_t)gp.table_bytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = THREADS / 32;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int items = chains * gp.slices;
    int k = 0;
    if (tid == 0 && blockIdx.x < items)
        issue_tables(&map_cl, &map_ms, tsm, &full[0], gp, (blockIdx.x % gp.slices) * W,
                     (long long)(blockIdx.x / gp.slices) * g.m);
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int buf = k & 1;
        const int next = item + gridDim.x;
        if (tid == 0 && next < items)
            issue_tables(&map_cl, &map_ms, tsm + (size_t)(buf ^ 1) * 2 * gp.table_bytes, &full[buf ^ 1], gp,
                         (next % gp.slices) * W, (long long)(next / gp.slices) * g.m);
        mbar_wait(&full[buf], (uint32_t)((k >> 1) & 1));
        const int chain = item / gp.slices, slice = item % gp.slices;
        const T* t_cl = reinterpret_cast<const T*>(tsm + (size_t)buf * 2 * gp.table_bytes);
        const T* t_ms = reinterpret_cast<const T*>(tsm + (size_t)buf * 2 * gp.table_bytes + gp.table_bytes);
        for (int v = warp; v < g.n; v += nwarps) {
            LaneVec<VS> s4[2], ms[2];
#pragma unroll
            for (int sgn = 0; sgn < 2; ++sgn) {
#pragma unroll
                for (int i = 0; i < VS; ++i) { s4[sgn].v[i] = 0.f; ms[sgn].v[i] = 0.f; }
                const int code = 2 * v + sgn;
                const int e0 = __ldg(g.lit_rowptr + code), e1 = __ldg(g.lit_rowptr + code + 1);
                for (int e = e0; e < e1; ++e) {
                    const int j = __ldg(g.lit_clause + e);
                    LaneVec<VS> a0 = lane_load_rw_t<VS, T>(t_cl + (size_t)j * W, lane);
                    LaneVec<VS> b0 = lane_load_rw_t<VS, T>(t_ms + (size_t)j * W, lane);
#pragma unroll
                    for (int i = 0; i < VS; ++i) { s4[sgn].v[i] += a0.v[i]; ms[sgn].v[i] += b0.v[i]; }
                }
            }
            const size_t row = (size_t)chain * g.n + v;
            LaneVec<VS> q = lane_load_t<VS, T>(QRY + row * ld_q + slice * W, lane);
            const float vw = __ldg(g.vdeg_w + v);
            const float dwp = __ldg(g.deg_w + 2 * v), dwn = __ldg(g.deg_w + 2 * v + 1);
            LaneVec<VS> grad;
#pragma unroll
            for (int i = 0; i < VS; ++i) {
                const float sg = 1.0f / (1.0f + __expf(-q.v[i]));
                grad.v[i] = (-sg * s4[0].v[i] + (1.0f - sg) * s4[1].v[i]) * vw;
                ms[0].v[i] *= dwp;
                ms[1].v[i] *= dwn;
            }
            T* dst = OUT + row * ld_out + out_off + slice * W;
            lane_store_t<VS, T>(dst, lane, grad);
            lane_store_t<VS, T>(dst + Q, lane, ms[0]);
            lane_store_t<VS, T>(dst + 2 * Q, lane, ms[1]);
        }
        __syncthreads();
    }
}

}  // namespace tg
}  // namespace dsat
