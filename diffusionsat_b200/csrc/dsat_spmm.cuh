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
// n = 10000) is fetched from HBM once and re-read from L2; the processing order additionally groups rows by length and, inside
// a length, makes rows that share their first gathered row neighbours.  The position is advanced with a decomposed stride (RowCursor): a 64-bit division per
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

template <int N> struct SpmmCount { static constexpr int value = N; };

// One 16-byte word of a row descriptor.  Volatile: otherwise the compiler sinks the loads of the column words below the
// branch on the row length (word 0), which turns descriptor -> gathers into word 0 -> column words -> gathers.
__device__ __forceinline__ int4 spmm_ld_desc(const int4* p) {
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// Row descriptor, DW int4 per output row in processing order:
//   {output row, length, bits of the row's scale, first entry in colidx}, then 4 * (DW - 1) gathered rows (or -1)
// A row whose entries all sit in its descriptor needs no colidx access at all: descriptor -> gathers is the whole dependent
// chain, and all its gathers are issued together.  DW = 2 (four entries: every clause of k<=4-SAT) on the clause side,
// DW = 4 (twelve entries: 98.5 % of the literals of a ratio-4.3 3-SAT formula) on the literal side, where walking colidx
// (descriptor -> colidx -> gathers, once per four entries) had left the narrow-row shapes latency-bound.  Longer rows
// walk colidx from the first entry.  The processing order groups rows of equal length (ensure_spmm_desc), so the rows that
// share a warp pass (RPW > 1) run the same number of steps.
//
// MINB = resident CTAs per SM the register allocation is held to (8 -> 32 registers, 6 -> 40, 5 -> 48, 4 -> 64): the
// instantiations whose rows give a lane two chunks or whose adds are bf16 -> fp32 spill at 32 registers (ptxas: 32 bytes of
// stack, 20-24 local loads / stores per pass) and ran at 43-55 % of the copy rate for it; `spmm_plan` in dsat_api.cu picks
// MINB per shape from the measured sweep (profiles/r2_spmm_variants.txt).
// PF: the descriptor of the NEXT pass is fetched before this pass's gathers are issued, so a pass's dependent chain is
// gathers -> store instead of descriptor -> gathers -> store (one L2 latency less per pass; pays on the clause side
// for rows of >= 512 bytes, costs registers everywhere).
template <int ROW_BYTES, bool BF16, int MINB = 8, bool PF = false, int DW = 2, int LPRT = 0>
__global__ void __launch_bounds__(GATHER_WARPS * 32, MINB)
spmm_rows_kernel(const int4* __restrict__ rowdesc, const int* __restrict__ colidx, int rows_out, int rows_in, int chains,
                 const void* __restrict__ Xv, void* __restrict__ Yv) {
    static_assert(DW == 2 || (DW == 4 && !PF), "descriptor prefetch is built for two-int4 descriptors only");
    constexpr int CH = ROW_BYTES / 16;           // 16-byte chunks per feature row
    constexpr int LPR = LPRT ? LPRT : (CH < 32 ? CH : 32);   // lanes per row; LPRT < CH gives a lane several strided chunks and
                                                 // the warp more rows per pass (more gathers in flight per warp, the pass
                                                 // overhead shared by more rows; needs the 64-register budget)
    constexpr int RPW = 32 / LPR;                // output rows per warp pass
    constexpr int CPL = CH / LPR;                // chunks per lane
    constexpr int EPC = BF16 ? 8 : 4;            // features per chunk
    constexpr int NC = 4 * (DW - 1);             // entries held by the descriptor
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane % LPR;
    const size_t x_chain = (size_t)rows_in * ROW_BYTES, y_chain = (size_t)rows_out * ROW_BYTES;
    RowCursor pass = row_cursor(((long long)blockIdx.x * GATHER_WARPS + (threadIdx.x >> 5)) * RPW,
                                (long long)gridDim.x * GATHER_WARPS * RPW, rows_out);
    // (chain, row position) of this lane's row in a pass; false past the last chain
    auto locate = [&](const RowCursor& k, int& c, int& pos) {
        c = k.c;
        pos = k.pos + sub;
        if constexpr (RPW > 1) {
            while (pos >= rows_out) { pos -= rows_out; ++c; }
        }
        return c < chains;
    };
    int4 n0 = make_int4(0, 0, 0, 0), n1 = make_int4(-1, -1, -1, -1);
    int nc = 0, npos = 0;
    bool nvalid = false;
    if constexpr (PF) {
        if (pass.c < chains && (nvalid = locate(pass, nc, npos))) {
            n0 = spmm_ld_desc(rowdesc + 2 * npos);
            n1 = spmm_ld_desc(rowdesc + 2 * npos + 1);
        }
    }
    for (; pass.c < chains;) {
        int c, pos;
        int4 d[DW];
        bool valid;
        if constexpr (PF) {
            c = nc; valid = nvalid; d[0] = n0; d[1] = n1;
            row_cursor_step(pass, rows_out);
            nvalid = false;
            if (pass.c < chains && (nvalid = locate(pass, nc, npos))) {
                n0 = spmm_ld_desc(rowdesc + 2 * npos);
                n1 = spmm_ld_desc(rowdesc + 2 * npos + 1);
            }
            if (!valid) continue;
        } else {
            valid = locate(pass, c, pos);
            row_cursor_step(pass, rows_out);
            if (!valid) continue;
#pragma unroll
            for (int w = 0; w < DW; ++w) d[w] = spmm_ld_desc(rowdesc + DW * pos + w);
        }
        const int len = d[0].y;
        DSAT_CHECK(c >= 0 && c < chains && d[0].x >= 0 && d[0].x < rows_out && len >= 0);
        const char* xc = reinterpret_cast<const char*>(Xv) + (size_t)c * x_chain + l * 16;
        asm("" : "+l"(xc));      // one full pointer: a gather address is then a single IMAD.WIDE (column x ROW_BYTES + xc)
        float acc[CPL][EPC];
#pragma unroll
        for (int k = 0; k < CPL; ++k)
#pragma unroll
            for (int i = 0; i < EPC; ++i) acc[k][i] = 0.f;
        // several rows per warp: the longest of them sets the (warp-uniform) number of steps, and the descriptor path is
        // taken only when all of them can, or both paths would run one after the other
        const unsigned active = RPW == 1 ? 0xffffffffu : __activemask();
        const int maxlen = RPW == 1 ? len : __reduce_max_sync(active, len);
        // rows of one pass almost always have the same length (processing order): then the loads need no predicates
        const bool same_len = RPW == 1 ? true : __all_sync(active, len == maxlen);
        if (maxlen <= NC && same_len) {
#pragma unroll
            for (int g = 0; g < DW - 1; ++g) {
                if (4 * g < maxlen) {
                    const int col[4] = {d[g + 1].x, d[g + 1].y, d[g + 1].z, d[g + 1].w};
                    const int here = maxlen - 4 * g;        // entries in this group of four (warp-uniform)
                    auto group = [&](auto cnt) {
                        constexpr int N = decltype(cnt)::value;
                        uint4 x[N][CPL];
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int k = 0; k < CPL; ++k) {
                                DSAT_CHECK(col[j] >= 0 && col[j] < rows_in);
                                x[j][k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)col[j] * ROW_BYTES) + k * LPR);
                            }
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int k = 0; k < CPL; ++k) spmm_add_chunk<BF16>(acc[k], x[j][k]);
                    };
                    if (here == 3) group(SpmmCount<3>{});          // 3-SAT clauses first
                    else if (here >= 4) group(SpmmCount<4>{});
                    else if (here == 2) group(SpmmCount<2>{});
                    else group(SpmmCount<1>{});
                }
            }
        } else if (maxlen <= NC) {
#pragma unroll
            for (int g = 0; g < DW - 1; ++g) {
                if (g == 0 || 4 * g < maxlen) {
                    const int col[4] = {d[g + 1].x, d[g + 1].y, d[g + 1].z, d[g + 1].w};
                    uint4 x[4][CPL];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            x[j][k] = make_uint4(0u, 0u, 0u, 0u);
                            if (4 * g + j < len)
                                x[j][k] = __ldg(reinterpret_cast<const uint4*>(xc + (size_t)col[j] * ROW_BYTES) + k * LPR);
                        }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (4 * g + j < maxlen) {   // warp-uniform
#pragma unroll
                            for (int k = 0; k < CPL; ++k) spmm_add_chunk<BF16>(acc[k], x[j][k]);
                        }
                }
            }
        } else {
            int e = d[0].w;
            const int e1 = e + len;
            for (; e + 4 <= e1; e += 4) {
                const int c0 = __ldg(colidx + e), c1 = __ldg(colidx + e + 1), c2 = __ldg(colidx + e + 2), c3 = __ldg(colidx + e + 3);
                DSAT_CHECK(c0 >= 0 && c0 < rows_in && c1 >= 0 && c1 < rows_in && c2 >= 0 && c2 < rows_in && c3 >= 0 && c3 < rows_in);
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
        const float s = __int_as_float(d[0].z);
        uint4* Y = reinterpret_cast<uint4*>(reinterpret_cast<char*>(Yv) + (size_t)c * y_chain + (size_t)d[0].x * ROW_BYTES + l * 16);
#pragma unroll
        for (int k = 0; k < CPL; ++k) __stcs(Y + k * LPR, spmm_pack_chunk<BF16>(acc[k], s));   // streamed: never re-read
    }
}

}  // namespace dsat
