// Whole-MLP kernel on the 5th-gen tensor cores: 2 or 3 Dense layers of one QuerySAT MLP
// (reference model/mlp.py:42-50) for a tile of 128 rows in ONE launch.  Hidden activations never leave
// the SM: each layer's fp32 accumulator is drained from TMEM by the epilogue warps (bias, leaky-relu,
// bf16) straight into shared memory in the K-major SWIZZLE_128B layout that the next layer's
// tcgen05.mma reads as its A operand.
//
//   shared memory   AH region   : the input tile A [128, K0] (TMA, 16 KB per 64 columns); after layer l
//                                 finished it is overwritten by that layer's hidden activations
//                   weight ring : 2-3 slots of 32 KB; slot = W_l^T[n_half*256 .. +256, kb*64 .. +64]
//                   staging     : the final output tile (reuses the regions above), copied out with
//                                 fully coalesced 16-byte stores
//   tensor memory   512 columns; layer l accumulates into columns [0, N_l)
//   warps           0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..5 = epilogue (lane quadrant = warp%4)
//
// Synchronisation (all mbarriers, no __syncthreads in the steady state):
//   a_full        TMA  -> MMA   input tile landed
//   ring full/empty     TMA <-> MMA   weight slots
//   tmem_full     MMA  -> epilogue   layer l accumulated (one phase per layer)
//   h_full[l]     epilogue -> MMA    hidden activations of layer l are in shared memory and TMEM is drained
#pragma once
#include "dsat_gemm_tc.cuh"

namespace dsat {
namespace fm {

using tc::TcOut;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int THREADS = 192;
constexpr int MAX_LAYERS = 3;
constexpr int MAX_SLOTS = 3;
constexpr int AH_BLOCK_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB: 128 rows x 64 bf16
constexpr int SLOT_BYTES = 256 * BLOCK_K * 2;              // 32 KB: 256 weight rows x 64 bf16
constexpr int TMEM_COLS = 512;
constexpr int BIAS_FLOATS = 1536;
constexpr int TAIL_BYTES = 256 + BIAS_FLOATS * 4;          // barriers + biases
constexpr int SMEM_LIMIT = 227 * 1024;

struct FmLayer {
    int K, N;              // multiples of 16; N <= 512
    int epi;               // tc::TC_LINEAR / TC_LRELU / TC_QUERY (query only on the last layer)
    int box_rows;          // TMA box height of this layer's weight map (min(256, N))
    const float* bias;
};

struct FmParams {
    int n_layers;
    FmLayer layer[MAX_LAYERS];
    int rows, a_box_rows, ah_blocks, slots, qmaps;
    TcOut out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w0,
                 const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, FmParams p) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* ah = smem;
    uint8_t* ring = smem + (size_t)p.ah_blocks * AH_BLOCK_BYTES;
    uint8_t* tail = ring + (size_t)p.slots * SLOT_BYTES;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* tmem_full = a_full + 1;
    uint64_t* h_full = a_full + 2;                     // [MAX_LAYERS - 1]
    uint64_t* ring_full = a_full + 4;                  // [MAX_SLOTS]
    uint64_t* ring_empty = a_full + 7;                 // [MAX_SLOTS]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 10);
    float* bias_s = reinterpret_cast<float*>(tail + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * BLOCK_M;
    const CUtensorMap* map_w[MAX_LAYERS] = {&map_w0, &map_w1, &map_w2};

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(tmem_full, 1);
        mbar_init(&h_full[0], 128);
        mbar_init(&h_full[1], 128);
        for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(&ring_full[s], 1); mbar_init(&ring_empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp >= 2) {       // biases of all layers -> shared memory (layer l at offset 512*l)
        for (int l = 0; l < p.n_layers; ++l)
            for (int i = threadIdx.x - 64; i < p.layer[l].N; i += 128) bias_s[512 * l + i] = __ldg(p.layer[l].bias + i);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ================================ TMA producer
            const int k0_blocks = (p.layer[0].K + BLOCK_K - 1) / BLOCK_K;
            mbar_expect_tx(a_full, (uint32_t)(k0_blocks * p.a_box_rows) * (BLOCK_K * 2));
            for (int kb = 0; kb < k0_blocks; ++kb)
                tma_load_2d(ah + (size_t)kb * AH_BLOCK_BYTES, &map_a, a_full, kb * BLOCK_K, row0);
            int slot = 0; uint32_t phase = 0;
            for (int l = 0; l < p.n_layers; ++l) {
                const int kbs = (p.layer[l].K + BLOCK_K - 1) / BLOCK_K;
                const int halves = (p.layer[l].N + 255) / 256;
                for (int kb = 0; kb < kbs; ++kb)
                    for (int h = 0; h < halves; ++h) {
                        mbar_wait(&ring_empty[slot], phase ^ 1);
                        mbar_expect_tx(&ring_full[slot], (uint32_t)p.layer[l].box_rows * (BLOCK_K * 2));
                        tma_load_2d(ring + (size_t)slot * SLOT_BYTES, map_w[l], &ring_full[slot], kb * BLOCK_K, h * 256);
                        if (++slot == p.slots) { slot = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ================================ MMA issuer
            int slot = 0; uint32_t phase = 0;
            for (int l = 0; l < p.n_layers; ++l) {
                if (l == 0) mbar_wait(a_full, 0);
                else mbar_wait(&h_full[l - 1], 0);
                tcgen05_fence_after();
                const int K = p.layer[l].K, N = p.layer[l].N;
                const int kbs = (K + BLOCK_K - 1) / BLOCK_K;
                const int halves = (N + 255) / 256;
                for (int kb = 0; kb < kbs; ++kb) {
                    const uint64_t da = make_smem_desc_sw128(smem_u32(ah + (size_t)kb * AH_BLOCK_BYTES));
                    const int ksteps = min(BLOCK_K / 16, (K - kb * BLOCK_K + 15) / 16);
                    for (int h = 0; h < halves; ++h) {
                        const int bn = min(256, N - h * 256);
                        const uint32_t idesc = make_idesc_bf16(BLOCK_M, bn);
                        mbar_wait(&ring_full[slot], phase);
                        tcgen05_fence_after();
                        const uint64_t db = make_smem_desc_sw128(smem_u32(ring + (size_t)slot * SLOT_BYTES));
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(tmem_base + (uint32_t)(h * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                      (kb | k) != 0);
                        tcgen05_commit(&ring_empty[slot]);
                        if (++slot == p.slots) { slot = 0; phase ^= 1; }
                    }
                }
                tcgen05_commit(tmem_full);
            }
        }
    } else {               // ================================ epilogue warps 2..5
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                    // row inside the tile = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int l = 0; l < p.n_layers; ++l) {
            const int N = p.layer[l].N, epi = p.layer[l].epi;
            const float* bl = bias_s + 512 * l;
            mbar_wait(tmem_full, (uint32_t)(l & 1));
            tcgen05_fence_after();
            if (l + 1 < p.n_layers) {
                // hidden activations -> AH region, K-major SWIZZLE_128B: 16-byte chunk j of row r of block kb
                // sits at kb*16KB + r*128 + ((j ^ (r & 7)) << 4)
                for (int c = 0; c < N; c += 32) {
                    uint32_t raw[32];
                    tmem_ld_32cols(lane_addr + (uint32_t)c, raw);
                    uint8_t* blk = ah + (size_t)(c >> 6) * AH_BLOCK_BYTES + (size_t)r * 128;
                    const int j0 = (c & 63) >> 3;          // first 16-byte chunk of this 32-column group (0 or 4)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float x = __uint_as_float(raw[8 * q + e]) + bl[c + 8 * q + e];
                            v[e] = (epi == TC_LRELU) ? (x > 0.f ? x : 0.2f * x) : x;
                        }
                        if (c + 8 * q < N) {
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
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
                mbar_arrive(&h_full[l]);
            } else {
                // final layer: stage the output tile in shared memory, then copy out coalesced
                const bool split = p.out.ptr1 != nullptr;
                const int cols0 = split ? p.out.split : (epi == TC_QUERY ? 3 * N : N);
                const int cols1 = split ? N - p.out.split : 0;
                const int es0 = p.out.bf16_0 ? 2 : 4, es1 = p.out.bf16_1 ? 2 : 4;
                const int stride0 = cols0 * es0 + 16, stride1 = cols1 * es1 + 16;
                uint8_t* area0 = smem;
                uint8_t* area1 = smem + (size_t)BLOCK_M * stride0;
                for (int c = 0; c < N; c += 32) {
                    uint32_t raw[32];
                    tmem_ld_32cols(lane_addr + (uint32_t)c, raw);
                    float v[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        float x = __uint_as_float(raw[e]) + ((c + e < N) ? bl[c + e] : 0.f);
                        v[e] = (epi == TC_LRELU) ? (x > 0.f ? x : 0.2f * x) : x;
                    }
                    const bool second = split && c >= p.out.split;
                    uint8_t* dst = second ? area1 + (size_t)r * stride1 : area0 + (size_t)r * stride0;
                    const int cc = second ? c - p.out.split : c;
                    const int es = second ? es1 : es0;
                    const int valid = min(32, N - c);              // multiple of 16
                    auto put = [&](const float (&w)[32], int col) {
                        if (es == 2) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (8 * q < valid) {
                                    __nv_bfloat162 p0 = __floats2bfloat162_rn(w[8 * q], w[8 * q + 1]);
                                    __nv_bfloat162 p1 = __floats2bfloat162_rn(w[8 * q + 2], w[8 * q + 3]);
                                    __nv_bfloat162 p2 = __floats2bfloat162_rn(w[8 * q + 4], w[8 * q + 5]);
                                    __nv_bfloat162 p3 = __floats2bfloat162_rn(w[8 * q + 6], w[8 * q + 7]);
                                    uint4 pack;
                                    pack.x = *reinterpret_cast<uint32_t*>(&p0); pack.y = *reinterpret_cast<uint32_t*>(&p1);
                                    pack.z = *reinterpret_cast<uint32_t*>(&p2); pack.w = *reinterpret_cast<uint32_t*>(&p3);
                                    *reinterpret_cast<uint4*>(dst + (size_t)col * 2 + 16 * q) = pack;
                                }
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (4 * q < valid)
                                    *reinterpret_cast<float4*>(dst + (size_t)col * 4 + 16 * q) =
                                        make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                        }
                    };
                    put(v, cc);
                    if (epi == TC_QUERY) {
                        float sp[32], sn[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float t = __logf(1.0f + __expf(-fabsf(v[e])));
                            sp[e] = fmaxf(v[e], 0.f) + t;
                            sn[e] = fmaxf(-v[e], 0.f) + t;
                        }
                        put(sp, cc + p.qmaps);
                        put(sn, cc + 2 * p.qmaps);
                    }
                }
                tcgen05_fence_before();
                asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
                const int ew = warp - 2;
                for (int a = 0; a < (split ? 2 : 1); ++a) {
                    const uint8_t* area = a ? area1 : area0;
                    const int stride = a ? stride1 : stride0;
                    const int row_bytes = a ? cols1 * es1 : cols0 * es0;
                    uint8_t* gbase = reinterpret_cast<uint8_t*>(a ? p.out.ptr1 : p.out.ptr0);
                    const size_t grow = (size_t)(a ? p.out.ld1 : p.out.ld0) * (a ? es1 : es0);
                    for (int rr = ew; rr < BLOCK_M; rr += 4) {
                        if (row0 + rr >= p.rows) break;
                        for (int off = lane * 16; off < row_bytes; off += 512)
                            *reinterpret_cast<uint4*>(gbase + (size_t)(row0 + rr) * grow + off) =
                                *reinterpret_cast<const uint4*>(area + (size_t)rr * stride + off);
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
};

// shared-memory plan: AH blocks, ring slots, staging; returns false when the MLP does not fit
inline bool plan_fused(FusedMlp& f) {
    FmParams& p = f.p;
    int blocks = (p.layer[0].K + 63) / 64;
    for (int l = 0; l + 1 < p.n_layers; ++l) blocks = max(blocks, (p.layer[l].N + 63) / 64);
    p.ah_blocks = blocks;
    const FmLayer& last = p.layer[p.n_layers - 1];
    const bool split = p.out.ptr1 != nullptr;
    const int cols0 = split ? p.out.split : (last.epi == tc::TC_QUERY ? 3 * last.N : last.N);
    const int cols1 = split ? last.N - p.out.split : 0;
    const int staging = BLOCK_M * (cols0 * (p.out.bf16_0 ? 2 : 4) + 16) + (split ? BLOCK_M * (cols1 * (p.out.bf16_1 ? 2 : 4) + 16) : 0);
    // the staging area reuses AH + ring and must end below the tail (barriers, biases)
    for (int slots = MAX_SLOTS; slots >= 2; --slots) {
        const int body = blocks * AH_BLOCK_BYTES + slots * SLOT_BYTES;
        if (1024 + body + TAIL_BYTES <= SMEM_LIMIT && staging <= body) {
            p.slots = slots;
            f.smem_bytes = 1024 + body + TAIL_BYTES;
            return true;
        }
    }
    return false;
}

inline cudaError_t launch_fused(const FusedMlp& f, cudaStream_t stream) {
    if (f.p.rows <= 0) return cudaSuccess;
    static int configured = 0;
    if (configured < f.smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(fused_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) return e;
        configured = SMEM_LIMIT;
    }
    const unsigned grid = (unsigned)ceil_div(f.p.rows, BLOCK_M);
    fused_mlp_kernel<<<grid, THREADS, f.smem_bytes, stream>>>(f.map_a, f.map_w[0], f.map_w[1],
                                                               f.map_w[f.p.n_layers > 2 ? 2 : 1], f.p);
    return cudaGetLastError();
}

}  // namespace fm
}  // namespace dsat
