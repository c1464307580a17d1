// Whole-MLP persistent kernel, CTA-PAIR version (tcgen05 cta_group::2, thread-block clusters of 2).
//
// Why: with one CTA per 128-row tile every tile re-streams the full weights of the MLP from L2
// (288 KB per tile for clause_update, ~1 MB for lit_query): at cfg2 that is ~9 GB per round and makes the
// MLP kernels L2-bandwidth bound.  Here two CTAs on the two SMs of a TPC work on one 256-row tile:
// each CTA keeps its own 128 rows of activations, loads only HALF of every weight block
// (rows [rank*N/2, (rank+1)*N/2) of W^T) and the leader CTA issues tcgen05.mma.cta_group::2 with M = 256,
// which reads A and B from both CTAs' shared memory and writes each CTA's 128 accumulator rows into its own
// tensor memory.  Weight traffic per row halves; everything else follows dsat_mlp_fused.cuh.
//
// Barriers that the leader's MMA thread waits on live in the LEADER's shared memory and are arrived on
// remotely by the peer (mapa + shared::cluster addressing): a_full, ring_full[], h_full, tmem_empty[].
// Barriers that producers / epilogues wait on are per CTA and are signalled by the leader with
// tcgen05.commit ... multicast::cluster to both CTAs: ring_empty[], tmem_full[], ah_free.
#pragma once
#include "dsat_mlp_fused.cuh"

namespace dsat {
namespace fm2 {

using namespace fm;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_C:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_C;\n"
        "bra WAIT_LOOP_C;\n"
        "WAIT_DONE_C:\n"
        "}\n" ::"r"(tc::smem_u32(bar)), "r"(parity) : "memory");
}
// TMA load into this CTA's shared memory, completion signalled on a barrier given by its shared::cluster address
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs once all prior MMAs completed
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
fused_mlp_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w0,
                      const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, FmParams p) {
    using tc::mbar_init; using tc::mbar_wait; using tc::smem_u32; using tc::tmem_ld_32cols;
    using tc::tcgen05_fence_before; using tc::tcgen05_fence_after; using tc::make_smem_desc_sw128;
    using tc::make_idesc_bf16; using tc::TC_LRELU; using tc::TC_QUERY;
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
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_macro = (p.n_tiles + 1) / 2;
    const CUtensorMap* map_w[MAX_LAYERS] = {&map_w0, &map_w1, &map_w2};

    if (threadIdx.x == 0) {
        mbar_init(a_full, 2);                          // one arrive.expect_tx from each CTA's producer
        mbar_init(ah_free, 1);
        mbar_init(h_full, 256);                        // epilogue threads of both CTAs
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 256); }
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&ring_full[s], 2); mbar_init(&ring_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (warp >= 2) {
        for (int l = 0; l < p.n_layers; ++l)
            for (int i = threadIdx.x - 64; i < p.layer[l].N; i += 128)
                bias_s[p.layer[l].bias_off + i] = __ldg(p.layer[l].bias + i);
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                 // both CTAs' barriers are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_layers = p.n_layers;

    if (warp == 0) {
        if (lane == 0) {   // ================================ TMA producer (one per CTA)
            const uint32_t a_full_leader = map_to_rank(smem_u32(a_full), 0);
            const int k0_blocks = (p.layer[0].K + BLOCK_K - 1) / BLOCK_K;
            int slot = 0; uint32_t phase = 0;
            int it = 0;
            for (int macro = pair; macro < n_macro; macro += n_pairs, ++it) {
                const int tile = 2 * macro + (int)rank;
                if (it > 0) mbar_wait(ah_free, (uint32_t)((it - 1) & 1));
                mbar_expect_tx_cluster(a_full_leader, (uint32_t)(k0_blocks * p.a_box_rows) * (BLOCK_K * 2));
                for (int kb = 0; kb < k0_blocks; ++kb)
                    tma_load_2d_pair(ah + (size_t)kb * AH_BLOCK_BYTES, &map_a, a_full_leader, kb * BLOCK_K, tile * BLOCK_M);
                for (int l = 0; l < n_layers; ++l) {
                    const int kbs = (p.layer[l].K + BLOCK_K - 1) / BLOCK_K;
                    const int N = p.layer[l].N;
                    const int halves = (N + 255) / 256;
                    for (int kb = 0; kb < kbs; ++kb)
                        for (int h = 0; h < halves; ++h) {
                            const int bn = min(256, N - h * 256);
                            mbar_wait(&ring_empty[slot], phase ^ 1);
                            const uint32_t full_leader = map_to_rank(smem_u32(&ring_full[slot]), 0);
                            mbar_expect_tx_cluster(full_leader, (uint32_t)(bn / 2) * (BLOCK_K * 2));
                            // this CTA's half of the weight block: rows [h*256 + rank*bn/2, +bn/2) of W^T
                            tma_load_2d_pair(ring + (size_t)slot * p.slot_bytes, map_w[l], full_leader, kb * BLOCK_K,
                                             h * 256 + (int)rank * (bn / 2));
                            if (++slot == p.slots) { slot = 0; phase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {   // ================================ MMA issuer (leader CTA only)
            int slot = 0; uint32_t phase = 0;
            int it = 0, g = 0, hcount = 0;
            for (int macro = pair; macro < n_macro; macro += n_pairs, ++it) {
                for (int l = 0; l < n_layers; ++l, ++g) {
                    const int buf = p.two_bufs ? (g & 1) : 0;
                    const int use = p.two_bufs ? (g >> 1) : g;
                    mbar_wait_cluster(&tmem_empty[buf], (uint32_t)((use & 1) ^ 1));
                    if (l == 0) mbar_wait_cluster(a_full, (uint32_t)(it & 1));
                    else { mbar_wait_cluster(h_full, (uint32_t)(hcount & 1)); ++hcount; }
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
                            const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, bn);
                            mbar_wait_cluster(&ring_full[slot], phase);
                            tcgen05_fence_after();
                            const uint64_t db = make_smem_desc_sw128(smem_u32(ring + (size_t)slot * p.slot_bytes));
                            for (int k = 0; k < ksteps; ++k)
                                umma_bf16_pair(acc + (uint32_t)(h * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                               (kb | k) != 0);
                            tcgen05_commit_pair(&ring_empty[slot]);
                            if (++slot == p.slots) { slot = 0; phase ^= 1; }
                        }
                    }
                    tcgen05_commit_pair(&tmem_full[buf]);
                    if (l == n_layers - 1) tcgen05_commit_pair(ah_free);
                }
            }
        }
    } else {               // ================================ epilogue warps 2..5 (both CTAs)
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        uint8_t* stage = stage_all + (warp - 2) * (32 * STAGE_ROW);
        const uint32_t h_full_leader = map_to_rank(smem_u32(h_full), 0);
        const uint32_t tmem_empty_leader[2] = {map_to_rank(smem_u32(&tmem_empty[0]), 0), map_to_rank(smem_u32(&tmem_empty[1]), 0)};
        int g = 0;
        for (int macro = pair; macro < n_macro; macro += n_pairs) {
            const int row0 = (2 * macro + (int)rank) * BLOCK_M;
            for (int l = 0; l < n_layers; ++l, ++g) {
                const int buf = p.two_bufs ? (g & 1) : 0;
                const int use = p.two_bufs ? (g >> 1) : g;
                const int N = p.layer[l].N, epi = p.layer[l].epi;
                const float* bl = bias_s + p.layer[l].bias_off;
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256);
                mbar_wait(&tmem_full[buf], (uint32_t)(use & 1));
                tcgen05_fence_after();
                if (l + 1 < n_layers) {
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
                    asm volatile("fence.proxy.async;" ::: "memory");      // st.shared -> visible to the (leader-issued) MMA
                    mbar_arrive_cluster(tmem_empty_leader[buf]);
                    mbar_arrive_cluster(h_full_leader);
                } else {
                    const bool split = p.out.ptr1 != nullptr;
                    const size_t row_first = (size_t)row0 + quad * 32;
                    const int rows_left = p.rows - (int)row_first;
                    for (int c = 0; c < N; c += 32) {
                        uint32_t raw[32];
                        tmem_ld_32cols(lane_addr + (uint32_t)c, raw);
                        if (c + 32 >= N) {
                            tcgen05_fence_before();
                            mbar_arrive_cluster(tmem_empty_leader[buf]);
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
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                 // the peer may still be reading this CTA's shared memory through the MMA
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// the pair kernel needs every layer to be <= 256 wide or a multiple of 256 (equal weight halves per MMA)
inline bool pair_supported(const FusedMlp& f) {
    for (int l = 0; l < f.p.n_layers; ++l) {
        const int N = f.p.layer[l].N;
        if (N % 16) return false;
        if (N > 256 && N % 256) return false;
    }
    return true;
}

inline cudaError_t launch_fused_pair(const FusedMlp& f, int sm_count, cudaStream_t stream) {
    if (f.p.rows <= 0) return cudaSuccess;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_mlp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int n_macro = (f.p.n_tiles + 1) / 2;
    const int pairs = n_macro < sm_count / 2 ? n_macro : sm_count / 2;
    fused_mlp_pair_kernel<<<2 * pairs, THREADS, f.smem_bytes, stream>>>(f.map_a, f.map_w[0], f.map_w[1],
                                                                         f.map_w[f.p.n_layers > 2 ? 2 : 1], f.p);
    return cudaGetLastError();
}

}  // namespace fm2
}  // namespace dsat
