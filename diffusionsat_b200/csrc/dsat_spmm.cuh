// Segment-sum SpMM on caller buffers (dsat_spmm; roofline sweeps of SURVEY.md section 8d):
//   Y[c, r, :] = scale[r] * sum_{e in row r} X[c, colidx[e], :]        for every chain c
// with ONE adjacency (the unit graph's CSR or CSC) shared by all chains.  These are the cores of reference
// model/query_sat.py:255-256 (clause <- literal) and :269-273 (literal <- clause) without the model's fused
// epilogues.  HBM-bound: every input row should come from DRAM once, every output row is written once.
//
// Work split: a lane always moves 16 bytes per gathered row (4 fp32 or 8 bf16 features), so a feature row of
// ROW_BYTES takes LPR = ROW_BYTES/16 lanes (at most 32; wider rows give a lane several chunks) and a warp produces
// 32/LPR output rows per pass.  Narrow rows (bf16, or 64 features) therefore issue as few instructions per byte
// as wide ones instead of leaving half of every request empty.
//
// Locality: the grid is exactly one resident wave and each warp strides over the flattened (chain, row) space,
// so at any moment all SMs work on the same one or two chains and a chain's gathered table (<= 22 MB at
// n = 10000) is fetched from HBM once and re-read from L2; `order` additionally makes rows that share their first
// gathered row neighbours.  The position is advanced with a decomposed stride (RowCursor): a 64-bit division per
// row had made the kernel issue-bound.  Sums run in entry order, fp32 accumulation.
#pragma once
#include "dsat_message.cuh"

namespace dsat {

template <bool BF16>
__device__ __forceinline__ void spmm_add_chunk(float* acc, const uint4& x) {
    if constexpr (BF16) {
        const uint32_t u[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[2 * i] = add_bf16_lo(acc[2 * i], u[i]);
            acc[2 * i + 1] = add_bf16_hi(acc[2 * i + 1], u[i]);
        }
    } else {
        acc[0] += __uint_as_float(x.x);
        acc[1] += __uint_as_float(x.y);
        acc[2] += __uint_as_float(x.z);
        acc[3] += __uint_as_float(x.w);
    }
}

template <bool BF16>
__device__ __forceinline__ uint4 spmm_pack_chunk(const float* acc, float s) {
    uint4 o;
    if constexpr (BF16) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[2 * i] * s, acc[2 * i + 1] * s);
    } else {
        o.x = __float_as_uint(acc[0] * s);
        o.y = __float_as_uint(acc[1] * s);
        o.z = __float_as_uint(acc[2] * s);
        o.w = __float_as_uint(acc[3] * s);
    }
    return o;
}

// Row descriptor, two int4 per output row in processing order:
//   {output row, length, bits of the row's scale, first entry in colidx}, {first four gathered rows (or -1)}
// A row of at most four entries (every clause of k<=4-SAT) needs no colidx access at all: descriptor -> gathers is
// the whole dependent chain, and all its gathers are issued together.  Longer rows walk colidx from the first entry.
template <int ROW_BYTES, bool BF16>
__global__ void __launch_bounds__(GATHER_WARPS * 32, 8)
spmm_rows_kernel(const int4* __restrict__ rowdesc, const int* __restrict__ colidx, int rows_out, int rows_in, int chains,
                 const void* __restrict__ Xv, void* __restrict__ Yv) {
    constexpr int CH = ROW_BYTES / 16;           // 16-byte chunks per feature row
    constexpr int LPR = CH < 32 ? CH : 32;       // lanes per row (fewer lanes per row with several strided chunks per lane were
                                                 // tried in round 2: 12-35 % of the copy rate -- the 32-register budget spills)
    constexpr int RPW = 32 / LPR;                // output rows per warp pass
    constexpr int CPL = CH / LPR;                // chunks per lane
    constexpr int EPC = BF16 ? 8 : 4;            // features per chunk
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane % LPR;
    const size_t x_chain = (size_t)rows_in * ROW_BYTES, y_chain = (size_t)rows_out * ROW_BYTES;
    RowCursor pass = row_cursor(((long long)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5)) * RPW,
                                (long long)gridDim.x * GATHER_WARPS * RPW, rows_out);
    for (; pass.c < chains; row_cursor_step(pass, rows_out)) {
        int c = pass.c, pos = pass.pos + sub;
        if constexpr (RPW > 1) {
            while (pos >= rows_out) { pos -= rows_out; ++c; }
            if (c >= chains) continue;
        }
        const int4 d0 = __ldg(rowdesc + 2 * pos), d1 = __ldg(rowdesc + 2 * pos + 1);
        const int len = d0.y;
        const char* xc = reinterpret_cast<const char*>(Xv) + (size_t)c * x_chain + l * 16;
        float acc[CPL][EPC];
#pragma unroll
        for (int k = 0; k < CPL; ++k)
#pragma unroll
            for (int i = 0; i < EPC; ++i) acc[k][i] = 0.f;
        // several rows per warp: take the short path only when all of them can, or both paths run one after the other
        const bool short_rows = RPW == 1 ? len <= 4 : __all_sync(__activemask(), len <= 4);
        if (short_rows) {
            const int col[4] = {d1.x, d1.y, d1.z, d1.w};
            uint4 x[4][CPL];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    x[j][k] = make_uint4(0u, 0u, 0u, 0u);
                    if (j < len)
                        x[j][k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)col[j] * ROW_BYTES) + k * LPR);
                }
            const bool any4 = RPW == 1 ? len == 4 : __any_sync(__activemask(), len == 4);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < 3 || any4) {        // 3-SAT: the fourth (all-zero) chunk is not added
#pragma unroll
                    for (int k = 0; k < CPL; ++k) spmm_add_chunk<BF16>(acc[k], x[j][k]);
                }
        } else {
            int e = d0.w;
            const int e1 = e + len;
            for (; e + 4 <= e1; e += 4) {
                const int c0 = __ldg(colidx + e), c1 = __ldg(colidx + e + 1), c2 = __ldg(colidx + e + 2), c3 = __ldg(colidx + e + 3);
                uint4 x0[CPL], x1[CPL], x2[CPL], x3[CPL];
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    x0[k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)c0 * ROW_BYTES) + k * LPR);
                    x1[k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)c1 * ROW_BYTES) + k * LPR);
                    x2[k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)c2 * ROW_BYTES) + k * LPR);
                    x3[k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)c3 * ROW_BYTES) + k * LPR);
                }
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    spmm_add_chunk<BF16>(acc[k], x0[k]); spmm_add_chunk<BF16>(acc[k], x1[k]);
                    spmm_add_chunk<BF16>(acc[k], x2[k]); spmm_add_chunk<BF16>(acc[k], x3[k]);
                }
            }
            for (; e < e1; ++e) {
                const int c0 = __ldg(colidx + e);
#pragma unroll
                for (int k = 0; k < CPL; ++k)
                    spmm_add_chunk<BF16>(acc[k], __ldg(reinterpret_cast<const uint4*>(xc + (size_t)c0 * ROW_BYTES) + k * LPR));
            }
        }
        const float s = __int_as_float(d0.z);
        uint4* Y = reinterpret_cast<uint4*>(reinterpret_cast<char*>(Yv) + (size_t)c * y_chain + (size_t)d0.x * ROW_BYTES + l * 16);
#pragma unroll
        for (int k = 0; k < CPL; ++k) __stcs(Y + k * LPR, spmm_pack_chunk<BF16>(acc[k], s));   // streamed: never re-read
    }
}

}  // namespace dsat
