// Whole-MLP persistent kernel on the 5th-gen tensor cores: the 2 or 3 Dense layers of one QuerySAT MLP
// (reference model/mlp.py:42-50) for tiles of 128 rows, one CTA per SM looping over its tiles.
// Hidden activations never leave the SM: each layer's fp32 accumulator is drained from TMEM by the
// epilogue warps (bias, leaky-relu, bf16) straight into shared memory in the K-major SWIZZLE_128B
// layout that the next layer's tcgen05.mma reads as its A operand.
//
//   shared memory   AH region   : the input tile A [128, K0] (TMA, 16 KB per 64 columns); after layer l
//                                 finished it is overwritten by that layer's hidden activations
//                   weight ring : 2-4 slots; slot = W_l^T[n_half*256 .. +256, kb*64 .. +64] (from L2)
//                   out staging : 4 x 4.5 KB, per-warp transpose so that the output leaves in 64/128-byte rows
//   tensor memory   512 columns; when every layer is at most 256 wide consecutive layers alternate between
//                   the two halves, so the first layer of tile i+1 accumulates while tile i's output drains
//   warps           0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue (lane quadrant = warp%4)
//
// mbarriers:  a_full (TMA->MMA, per tile)      ah_free (MMA->TMA, AH may take the next input tile)
//             ring full/empty (TMA<->MMA)       tmem_full[buf] (MMA->epilogue)   tmem_empty[buf] (epilogue->MMA)
//             h_full (epilogue->MMA, hidden activations of a layer are in shared memory)
#pragma once
#include "dsat_gemm_tc.cuh"

namespace dsat {
namespace fm {

using tc::TcOut;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int THREADS = 192;
constexpr int MAX_LAYERS = 3;
constexpr int MAX_SLOTS = 8;
constexpr int AH_BLOCK_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB: 128 rows x 64 bf16
constexpr int TMEM_COLS = 512;
constexpr int STAGE_ROW = 128 + 16;                        // bytes per staged row (one 32-column fp32 chunk + pad)
constexpr int STAGE_BYTES = 4 * 32 * STAGE_ROW;            // four epilogue warps
constexpr int BAR_BYTES = 256;
constexpr int SMEM_LIMIT = 227 * 1024;

struct FmLayer {
    int K, N;              // multiples of 16; N <= 512
    int epi;               // tc::TC_LINEAR / TC_LRELU / TC_QUERY (query only on the last layer)
    int box_rows;          // TMA box height of this layer's weight map (min(256, N))
    int bias_off;          // offset of this layer's biases in the shared bias array
    const float* bias;
};

struct FmParams {
    int n_layers;
    FmLayer layer[MAX_LAYERS];
    int rows, n_tiles, a_box_rows, ah_blocks, slots, slot_bytes, two_bufs, qmaps;
    TcOut out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// one 32-column chunk of one output row per lane -> global memory through a per-warp transpose buffer
template <bool BF16>
__device__ __forceinline__ void store_chunk_coalesced(uint8_t* stage, int lane, const float (&v)[32], int valid,
                                                      uint8_t* gbase, size_t row_pitch_bytes, size_t row_first,
                                                      int rows_left, int col) {
    uint8_t* mine = stage + lane * STAGE_ROW;
    if (BF16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * q], v[8 * q + 1]), p1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]), p3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
            uint4 pack;
            pack.x = *reinterpret_cast<uint32_t*>(&p0); pack.y = *reinterpret_cast<uint32_t*>(&p1);
            pack.z = *reinterpret_cast<uint32_t*>(&p2); pack.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(mine + 16 * q) = pack;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(mine + 16 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncwarp();
    constexpr int ES = BF16 ? 2 : 4;
    constexpr int PIECES = BF16 ? 4 : 8;                    // 16-byte pieces per staged row
    const int valid_pieces = valid * ES / 16;
#pragma unroll
    for (int k = 0; k < PIECES; ++k) {
        const int pidx = lane + 32 * k;
        const int rr = pidx / PIECES, piece = pidx % PIECES;
        if (rr < rows_left && piece < valid_pieces)
            *reinterpret_cast<uint4*>(gbase + (row_first + rr) * row_pitch_bytes + (size_t)col * ES + 16 * piece) =
                *reinterpret_cast<const uint4*>(stage + rr * STAGE_ROW + 16 * piece);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(THREADS, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w0,
                 const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, FmParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ah = smem;
    uint8_t* ring = smem + (size_t)p.ah_blocks * AH_BLOCK_BYTES;
    uint8_t* stage_all = ring + (size_t)p.slots * p.slot_bytes;
    uint8_t* tail = stage_all + STAGE_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* ah_free = a_full + 1;
    uint64_t* h_full = a_full + 2;
    uint64_t* tmem_full = a_full + 3;                  // [2]
    uint64_t* tmem_empty = a_full + 5;                 // [2]
    uint64_t* ring_full = a_full + 7;                  // [MAX_SLOTS]
    uint64_t* ring_empty = a_full + 7 + MAX_SLOTS;     // [MAX_SLOTS]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 7 + 2 * MAX_SLOTS);
    float* bias_s = reinterpret_cast<float*>(tail + BAR_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CUtensorMap* map_w[MAX_LAYERS] = {&map_w0, &map_w1, &map_w2};

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(ah_free, 1);
        mbar_init(h_full, 128);
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 128); }
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 2) {
        for (int l = 0; l < p.n_layers; ++l)
            for (int i = threadIdx.x - 64; i < p.layer[l].N; i += 128)
                bias_s[p.layer[l].bias_off + i] = __ldg(p.layer[l].bias + i);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_layers = p.n_layers;

    if (warp == 0) {
        if (lane == 0) {   // ================================ TMA producer
            const int k0_blocks = (p.layer[0].K + BLOCK_K - 1) / BLOCK_K;
            int slot = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                if (it > 0) mbar_wait(ah_free, (uint32_t)((it - 1) & 1));       // tile it-1 no longer reads AH
                mbar_expect_tx(a_full, (uint32_t)(k0_blocks * p.a_box_rows) * (BLOCK_K * 2));
                for (int kb = 0; kb < k0_blocks; ++kb)
                    tma_load_2d(ah + (size_t)kb * AH_BLOCK_BYTES, &map_a, a_full, kb * BLOCK_K, tile * BLOCK_M);
                for (int l = 0; l < n_layers; ++l) {
                    const int kbs = (p.layer[l].K + BLOCK_K - 1) / BLOCK_K;
                    const int halves = (p.layer[l].N + 255) / 256;
                    for (int kb = 0; kb < kbs; ++kb)
                        for (int h = 0; h < halves; ++h) {
                            mbar_wait(&ring_empty[slot], phase ^ 1);
                            mbar_expect_tx(&ring_full[slot], (uint32_t)p.layer[l].box_rows * (BLOCK_K * 2));
                            tma_load_2d(ring + (size_t)slot * p.slot_bytes, map_w[l], &ring_full[slot], kb * BLOCK_K, h * 256);
                            if (++slot == p.slots) { slot = 0; phase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ================================ MMA issuer
            int slot = 0; uint32_t phase = 0;
            int it = 0, g = 0, hcount = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                for (int l = 0; l < n_layers; ++l, ++g) {
                    const int buf = p.two_bufs ? (g & 1) : 0;
                    const int use = p.two_bufs ? (g >> 1) : g;
                    mbar_wait(&tmem_empty[buf], (uint32_t)((use & 1) ^ 1));     // accumulator drained (first use passes)
                    if (l == 0) mbar_wait(a_full, (uint32_t)(it & 1));
                    else { mbar_wait(h_full, (uint32_t)(hcount & 1)); ++hcount; }
                    tcgen05_fence_after();
                    const int K = p.layer[l].K, N = p.layer[l].N;
                    const int kbs = (K + BLOCK_K - 1) / BLOCK_K;
                    const int halves = (N + 255) / 256;
                    const uint32_t acc = tmem_base + (uint32_t)(buf * 256);
                    for (int kb = 0; kb < kbs; ++kb) {
                        const uint64_t da = make_smem_desc_sw128(smem_u32(ah + (size_t)kb * AH_BLOCK_BYTES));
                        const int ksteps = min(BLOCK_K / 16, (K - kb * BLOCK_K + 15) / 16);
                        for (int h = 0; h < halves; ++h) {
                            const int bn = min(256, N - h * 256);
                            const uint32_t idesc = make_idesc_bf16(BLOCK_M, bn);
                            mbar_wait(&ring_full[slot], phase);
                            tcgen05_fence_after();
                            const uint64_t db = make_smem_desc_sw128(smem_u32(ring + (size_t)slot * p.slot_bytes));
                            for (int k = 0; k < ksteps; ++k)
                                umma_bf16(acc + (uint32_t)(h * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                          (kb | k) != 0);
                            tcgen05_commit(&ring_empty[slot]);
                            if (++slot == p.slots) { slot = 0; phase ^= 1; }
                        }
                    }
                    tcgen05_commit(&tmem_full[buf]);
                    if (l == n_layers - 1) tcgen05_commit(ah_free);
                }
            }
        }
    } else {               // ================================ epilogue warps 2..5
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                    // row inside the tile = TMEM lane
        uint8_t* stage = stage_all + (warp - 2) * (32 * STAGE_ROW);
        int g = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int row0 = tile * BLOCK_M;
            for (int l = 0; l < n_layers; ++l, ++g) {
                const int buf = p.two_bufs ? (g & 1) : 0;
                const int use = p.two_bufs ? (g >> 1) : g;
                const int N = p.layer[l].N, epi = p.layer[l].epi;
                const float* bl = bias_s + p.layer[l].bias_off;
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256);
                mbar_wait(&tmem_full[buf], (uint32_t)(use & 1));
                tcgen05_fence_after();
                if (l + 1 < n_layers) {
                    // hidden activations -> AH region, K-major SWIZZLE_128B: 16-byte chunk j of row r of block kb
                    // sits at kb*16KB + r*128 + ((j ^ (r & 7)) << 4)
                    for (int c = 0; c < N; c += 32) {
                        uint32_t raw[32];
                        tmem_ld_32cols(lane_addr + (uint32_t)c, raw);
                        uint8_t* blk = ah + (size_t)(c >> 6) * AH_BLOCK_BYTES + (size_t)r * 128;
                        const int j0 = (c & 63) >> 3;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (c + 8 * q < N) {
                                float v[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float x = __uint_as_float(raw[8 * q + e]) + bl[c + 8 * q + e];
                                    v[e] = (epi == TC_LRELU) ? (x > 0.f ? x : 0.2f * x) : x;
                                }
                                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                                uint4 pack;
                                pack.x = *reinterpret_cast<uint32_t*>(&p0); pack.y = *reinterpret_cast<uint32_t*>(&p1);
                                pack.z = *reinterpret_cast<uint32_t*>(&p2); pack.w = *reinterpret_cast<uint32_t*>(&p3);
                                *reinterpret_cast<uint4*>(blk + (((j0 + q) ^ (r & 7)) << 4)) = pack;
                            }
                        }
                    }
                    tcgen05_fence_before();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> visible to the MMA (async proxy)
                    mbar_arrive(&tmem_empty[buf]);
                    mbar_arrive(h_full);
                } else {
                    const bool split = p.out.ptr1 != nullptr;
                    const size_t row_first = (size_t)row0 + quad * 32;
                    const int rows_left = p.rows - (int)row_first;          // rows of this warp that exist (may be <= 0)
                    for (int c = 0; c < N; c += 32) {
                        uint32_t raw[32];
                        tmem_ld_32cols(lane_addr + (uint32_t)c, raw);
                        if (c + 32 >= N) {                  // last TMEM read of this accumulator: hand it back early
                            tcgen05_fence_before();
                            mbar_arrive(&tmem_empty[buf]);
                        }
                        float v[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float x = __uint_as_float(raw[e]) + ((c + e < N) ? bl[c + e] : 0.f);
                            v[e] = (epi == TC_LRELU) ? (x > 0.f ? x : 0.2f * x) : x;
                        }
                        const bool second = split && c >= p.out.split;
                        uint8_t* gbase = reinterpret_cast<uint8_t*>(second ? p.out.ptr1 : p.out.ptr0);
                        const int is_bf16 = second ? p.out.bf16_1 : p.out.bf16_0;
                        const size_t pitch = (size_t)(second ? p.out.ld1 : p.out.ld0) * (is_bf16 ? 2 : 4);
                        const int cc = second ? c - p.out.split : c;
                        const int valid = min(32, N - c);
                        if (is_bf16) store_chunk_coalesced<true>(stage, lane, v, valid, gbase, pitch, row_first, rows_left, cc);
                        else store_chunk_coalesced<false>(stage, lane, v, valid, gbase, pitch, row_first, rows_left, cc);
                        if (epi == TC_QUERY) {
                            float sp[32], sn[32];
#pragma unroll
                            for (int e = 0; e < 32; ++e) {
                                const float t = __logf(1.0f + __expf(-fabsf(v[e])));
                                sp[e] = fmaxf(v[e], 0.f) + t;
                                sn[e] = fmaxf(-v[e], 0.f) + t;
                            }
                            if (is_bf16) {
                                store_chunk_coalesced<true>(stage, lane, sp, valid, gbase, pitch, row_first, rows_left, cc + p.qmaps);
                                store_chunk_coalesced<true>(stage, lane, sn, valid, gbase, pitch, row_first, rows_left, cc + 2 * p.qmaps);
                            } else {
                                store_chunk_coalesced<false>(stage, lane, sp, valid, gbase, pitch, row_first, rows_left, cc + p.qmaps);
                                store_chunk_coalesced<false>(stage, lane, sn, valid, gbase, pitch, row_first, rows_left, cc + 2 * p.qmaps);
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

struct FusedMlp {
    CUtensorMap map_a;
    CUtensorMap map_w[MAX_LAYERS];
    FmParams p;
    int smem_bytes;
    // CTA-pair variant (dsat_mlp_pair.cuh): weight maps with half-height boxes, its own ring plan
    CUtensorMap map_wp[MAX_LAYERS];
    FmParams pp;
    int smem_bytes_pair;
    bool pair_ok;
};

// shared-memory plan; returns false when the MLP does not fit
inline bool plan_fused(FusedMlp& f) {
    FmParams& p = f.p;
    int blocks = (p.layer[0].K + 63) / 64;
    int bias_total = 0, max_box = 0;
    p.two_bufs = 1;
    for (int l = 0; l < p.n_layers; ++l) {
        if (l + 1 < p.n_layers) blocks = max(blocks, (p.layer[l].N + 63) / 64);
        p.layer[l].bias_off = bias_total;
        bias_total += (p.layer[l].N + 31) / 32 * 32;
        max_box = max(max_box, p.layer[l].box_rows);
        if (p.layer[l].N > 256) p.two_bufs = 0;
    }
    p.ah_blocks = blocks;
    p.slot_bytes = (max_box * BLOCK_K * 2 + 1023) / 1024 * 1024;
    p.n_tiles = ceil_div(p.rows, BLOCK_M);
    // pair variant: each CTA holds half of every weight block, so slots are half as large
    f.pp = p;
    f.pp.slot_bytes = (max_box / 2 * BLOCK_K * 2 + 1023) / 1024 * 1024;
    f.pp.slots = 0;
    for (int slots = MAX_SLOTS; slots >= 2; --slots) {
        const int total = 1024 + blocks * AH_BLOCK_BYTES + slots * f.pp.slot_bytes + STAGE_BYTES + BAR_BYTES + bias_total * 4;
        if (total <= SMEM_LIMIT) { f.pp.slots = slots; f.smem_bytes_pair = total; break; }
    }
    for (int slots = MAX_SLOTS; slots >= 2; --slots) {
        const int total = 1024 + blocks * AH_BLOCK_BYTES + slots * p.slot_bytes + STAGE_BYTES + BAR_BYTES + bias_total * 4;
        if (total <= SMEM_LIMIT) {
            p.slots = slots;
            f.pp.ah_blocks = p.ah_blocks; f.pp.n_tiles = p.n_tiles; f.pp.two_bufs = p.two_bufs;
            for (int l = 0; l < p.n_layers; ++l) f.pp.layer[l].bias_off = p.layer[l].bias_off;
            f.smem_bytes = total;
            return true;
        }
    }
    return false;
}

inline cudaError_t launch_fused(const FusedMlp& f, int sm_count, cudaStream_t stream) {
    if (f.p.rows <= 0) return cudaSuccess;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const unsigned grid = (unsigned)(f.p.n_tiles < sm_count ? f.p.n_tiles : sm_count);
    fused_mlp_kernel<<<grid, THREADS, f.smem_bytes, stream>>>(f.map_a, f.map_w[0], f.map_w[1],
                                                               f.map_w[f.p.n_layers > 2 ? 2 : 1], f.p);
    return cudaGetLastError();
}

}  // namespace fm
}  // namespace dsat
