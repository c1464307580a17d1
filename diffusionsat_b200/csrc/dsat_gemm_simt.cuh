// fp32 CUDA-core fused linear layer  Y = epi(A @ W + b)   (the fp32 parity path of the MLPs,
// reference model/mlp.py:42-50 with Dense = x @ kernel + bias, hidden activation leaky_relu 0.2).
//
// A  [rows, K]  row-major, leading dimension lda (a column slice of a wider row buffer is fine)
// W  [K, N]     row-major, leading dimension ldw; K and N are multiples of 16 (zero padded)
// Y  [rows, N]  leading dimension ldy
// Tiling: 128x128x16 per CTA of 256 threads, 8x8 outputs per thread split as 2x2 blocks of 4x4 so that
// shared-memory reads are conflict-free float4; register-staged double buffering of the K tiles.
#pragma once
#include "dsat_common.cuh"

namespace dsat {

enum Epilogue : int { EPI_LINEAR = 0, EPI_LRELU = 1, EPI_QUERY = 2 };

template <int EPI>
__global__ void __launch_bounds__(256, 2)
sgemm128_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                const float* __restrict__ bias, float* __restrict__ Y, int ldy,
                int rows, int K, int N, int qmaps, int n_tiles) {
    constexpr int BM = 128, BN = 128, BK = 16, APAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    // 1-D grid, N tiles fastest: the CTAs that share an A row tile run back to back (L2 reuse)
    const int row0 = (blockIdx.x / n_tiles) * BM;
    const int n0 = (blockIdx.x % n_tiles) * BN;
    const int a_row = tid >> 2, a_k = (tid & 3) * 4;
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;
    const int ty = tid >> 4, tx = tid & 15;

    float4 ra[2], rb[2];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = row0 + a_row + i * 64;
            ra[i] = (r < rows) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)r * lda + k0 + a_k)) : zero4;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int kk = k0 + b_k + i * 8;
            int nn = n0 + b_n;
            rb[i] = (nn < N) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)kk * ldw + nn)) : zero4;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = a_row + i * 64;
            As[buf][a_k + 0][r] = ra[i].x;
            As[buf][a_k + 1][r] = ra[i].y;
            As[buf][a_k + 2][r] = ra[i].z;
            As[buf][a_k + 3][r] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
            *reinterpret_cast<float4*>(&Bs[buf][b_k + i * 8][b_n]) = rb[i];
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = K / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();

    for (int kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(cur ^ 1);
        __syncthreads();
    }

    // epilogue: bias, activation, store (float4 per 4 columns)
#pragma unroll
    for (int hj = 0; hj < 2; ++hj) {
        const int col = n0 + hj * 64 + tx * 4;
        if (col >= N) continue;
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = row0 + (i >> 2) * 64 + ty * 4 + (i & 3);
            if (r >= rows) continue;
            float v[4] = {acc[i][hj * 4 + 0] + bb.x, acc[i][hj * 4 + 1] + bb.y,
                          acc[i][hj * 4 + 2] + bb.z, acc[i][hj * 4 + 3] + bb.w};
            if (EPI == EPI_LRELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
            }
            float* dst = Y + (size_t)r * ldy + col;
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            if (EPI == EPI_QUERY) {
                // query head: also emit softplus(+q) and softplus(-q) for the clause
                // "unsat probability" gather (reference loss/sat.py:132-133)
                float sp[4], sn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { sp[j] = softplus_f(v[j]); sn[j] = softplus_f(-v[j]); }
                *reinterpret_cast<float4*>(dst + qmaps) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                *reinterpret_cast<float4*>(dst + 2 * qmaps) = make_float4(sn[0], sn[1], sn[2], sn[3]);
            }
        }
    }
}

struct LinearOp {
    const float* A; int lda;
    const float* W; int ldw;
    const float* bias;
    float* Y; int ldy;
    int rows, K, N, epi, qmaps;
};

inline cudaError_t launch_sgemm(const LinearOp& op, cudaStream_t stream) {
    if (op.rows <= 0) return cudaSuccess;
    const int n_tiles = ceil_div(op.N, 128);
    dim3 grid((unsigned)(n_tiles * ceil_div(op.rows, 128)));
    switch (op.epi) {
        case EPI_LINEAR:
            sgemm128_kernel<EPI_LINEAR><<<grid, 256, 0, stream>>>(op.A, op.lda, op.W, op.ldw, op.bias, op.Y, op.ldy,
                                                                  op.rows, op.K, op.N, op.qmaps, n_tiles);
            break;
        case EPI_LRELU:
            sgemm128_kernel<EPI_LRELU><<<grid, 256, 0, stream>>>(op.A, op.lda, op.W, op.ldw, op.bias, op.Y, op.ldy,
                                                                 op.rows, op.K, op.N, op.qmaps, n_tiles);
            break;
        default:
            sgemm128_kernel<EPI_QUERY><<<grid, 256, 0, stream>>>(op.A, op.lda, op.W, op.ldw, op.bias, op.Y, op.ldy,
                                                                 op.rows, op.K, op.N, op.qmaps, n_tiles);
            break;
    }
    return cudaGetLastError();
}

}  // namespace dsat
