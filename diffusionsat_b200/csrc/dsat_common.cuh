// Shared device helpers for the DiffusionSAT sampling kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define DSAT_AUX_PAD 16      // aux columns of v1 (9 used: normal4, noisy2, noise_scale, denoised2)
#define DSAT_LOGIT_MAPS 8    // reference model/query_sat.py:99
#define DSAT_LOGIT_PAD 16

// Debug build (`DSAT_NVCC_FLAGS=-DDSAT_ASSERT python -m diffusionsat_b200.build --force`): bounds checks on every index the
// gathers read and on the tile / row ranges of the whole-MLP kernels; a violated check traps (the launch fails with
// cudaErrorLaunchFailure instead of reading or writing out of range).  compute-sanitizer is not available on this pool; the
// GPU test-suite is run once per round under this build instead (profiles/r2_assert_build.txt).
#ifdef DSAT_ASSERT
// a trap discards the device printf buffer and poisons the context, so the failing check leaves its line number in
// host-mapped memory first (dsat_create allocates it; CK_CUDA appends it to the error message)
__device__ int* g_dsat_assert_slot = nullptr;
#define DSAT_CHECK(cond)                                                                    \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            if (g_dsat_assert_slot) { g_dsat_assert_slot[0] = __LINE__; __threadfence_system(); } \
            __trap();                                                                       \
        }                                                                                   \
    } while (0)
#else
#define DSAT_CHECK(cond) do { } while (0)
#endif

namespace dsat {

__host__ __device__ inline int pad16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the shared-memory base probe apply to the CURRENT device only:
// a process that holds one context per GPU has to repeat them per device (every C-ABI entry calls cudaSetDevice first).
struct PerDeviceOnce {
    unsigned long long mask = 0;
    static int current() { int d = 0; return cudaGetDevice(&d) == cudaSuccess && d >= 0 && d < 64 ? d : -1; }
    bool done() const { const int d = current(); return d >= 0 && ((mask >> d) & 1ull); }
    void mark() { const int d = current(); if (d >= 0) mask |= 1ull << d; }
};

// Early exit of a reference batch (model/query_sat.py:330-338): once every graph of an early-exit group is satisfied the
// group's round loop breaks.  `done[group]` is set on the device at the end of the breaking round; every later kernel of the
// same model call skips the rows of that group's chains (their buffers keep the state of the breaking round, which nothing
// reads any more).  done == nullptr: groups do not consist of whole chains, nothing is skipped.
struct SkipInfo {
    const int* done;
    int chains_per_group;
};
__device__ __forceinline__ bool chain_done(const SkipInfo& s, int chain) {
    return s.done != nullptr && __ldg(s.done + chain / s.chains_per_group) != 0;
}

// ---------------------------------------------------------------------------------- math
// softplus with the usual large-argument guard (log1p(exp(x)) otherwise); reference
// loss/sat.py:132 uses tf.nn.softplus.
__device__ __forceinline__ float softplus_f(float x) {
    return x > 20.0f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// tf.round(tf.sigmoid(z)) as a bit: half-to-even makes sigma == 0.5 round to 0
// (reference utils/sat.py:119, satuniformity/DiffusionSampler.py:154).
__device__ __forceinline__ int sigmoid_bit(float z) { return sigmoid_f(z) > 0.5f ? 1 : 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------- per-lane row vectors
// A feature row of width W = 32*V floats is spread over a warp: lane l owns, for V >= 4, the
// float4 chunks at columns c*128 + 4*l (c < V/4); for V == 2 the float2 at column 2*l.
template <int V> struct LaneVec { float v[V]; };

template <int V>
__device__ __forceinline__ LaneVec<V> lane_load(const float* __restrict__ row, int lane) {
    LaneVec<V> r;
    if constexpr (V == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(row) + lane);
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            float4 t = __ldg(reinterpret_cast<const float4*>(row + c * 128) + lane);
            r.v[4 * c + 0] = t.x; r.v[4 * c + 1] = t.y; r.v[4 * c + 2] = t.z; r.v[4 * c + 3] = t.w;
        }
    }
    return r;
}

// plain (coherent) load for buffers written earlier in the same kernel
template <int V>
__device__ __forceinline__ LaneVec<V> lane_load_rw(const float* row, int lane) {
    LaneVec<V> r;
    if constexpr (V == 2) {
        float2 t = reinterpret_cast<const float2*>(row)[lane];
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            float4 t = reinterpret_cast<const float4*>(row + c * 128)[lane];
            r.v[4 * c + 0] = t.x; r.v[4 * c + 1] = t.y; r.v[4 * c + 2] = t.z; r.v[4 * c + 3] = t.w;
        }
    }
    return r;
}

template <int V>
__device__ __forceinline__ void lane_store(float* __restrict__ row, int lane, const LaneVec<V>& r) {
    if constexpr (V == 2) {
        reinterpret_cast<float2*>(row)[lane] = make_float2(r.v[0], r.v[1]);
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c)
            reinterpret_cast<float4*>(row + c * 128)[lane] =
                make_float4(r.v[4 * c + 0], r.v[4 * c + 1], r.v[4 * c + 2], r.v[4 * c + 3]);
    }
}

// bf16 rows: lane owns the same columns as in the fp32 layout (8-byte loads for V == 4)
template <int V>
__device__ __forceinline__ LaneVec<V> lane_load_bf16(const __nv_bfloat16* __restrict__ row, int lane) {
    LaneVec<V> r;
    if constexpr (V == 2) {
        __nv_bfloat162 t = reinterpret_cast<const __nv_bfloat162*>(row)[lane];
        r.v[0] = __low2float(t); r.v[1] = __high2float(t);
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            uint2 raw = __ldg(reinterpret_cast<const uint2*>(row + c * 128) + lane);
            __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
            __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
            r.v[4 * c + 0] = __low2float(a); r.v[4 * c + 1] = __high2float(a);
            r.v[4 * c + 2] = __low2float(b); r.v[4 * c + 3] = __high2float(b);
        }
    }
    return r;
}

template <int V>
__device__ __forceinline__ void lane_store_bf16(__nv_bfloat16* __restrict__ row, int lane, const LaneVec<V>& r) {
    if constexpr (V == 2) {
        reinterpret_cast<__nv_bfloat162*>(row)[lane] = __floats2bfloat162_rn(r.v[0], r.v[1]);
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            __nv_bfloat162 a = __floats2bfloat162_rn(r.v[4 * c + 0], r.v[4 * c + 1]);
            __nv_bfloat162 b = __floats2bfloat162_rn(r.v[4 * c + 2], r.v[4 * c + 3]);
            uint2 raw;
            raw.x = *reinterpret_cast<uint32_t*>(&a);
            raw.y = *reinterpret_cast<uint32_t*>(&b);
            reinterpret_cast<uint2*>(row + c * 128)[lane] = raw;
        }
    }
}

// Split-plane storage (fp32-accurate tensor-core path, dsat_mlp_x3.cuh): a value v lives as hi = bf16(v) in one
// plane and lo = bf16(v - hi) in a second plane `plane` elements further on (v = hi + lo up to 2^-17 |v|); both
// planes have the layout of a bf16 row buffer, so the MLP kernels fetch them as two bf16 A operands.
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<uint32_t*>(&h);
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<uint32_t*>(&l);
}
template <int V>
__device__ __forceinline__ void lane_store_split(__nv_bfloat16* __restrict__ row_hi, size_t plane, int lane, const LaneVec<V>& r) {
    if constexpr (V == 2) {
        uint32_t hi, lo;
        split_bf16x2(r.v[0], r.v[1], hi, lo);
        reinterpret_cast<uint32_t*>(row_hi)[lane] = hi;
        reinterpret_cast<uint32_t*>(row_hi + plane)[lane] = lo;
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            uint2 hi, lo;
            split_bf16x2(r.v[4 * c + 0], r.v[4 * c + 1], hi.x, lo.x);
            split_bf16x2(r.v[4 * c + 2], r.v[4 * c + 3], hi.y, lo.y);
            reinterpret_cast<uint2*>(row_hi + c * 128)[lane] = hi;
            reinterpret_cast<uint2*>(row_hi + plane + c * 128)[lane] = lo;
        }
    }
}
// coherent load of hi + lo (rows that the same kernel also writes)
template <int V>
__device__ __forceinline__ LaneVec<V> lane_load_split_rw(const __nv_bfloat16* row_hi, size_t plane, int lane) {
    LaneVec<V> r;
    if constexpr (V == 2) {
        const uint32_t hi = reinterpret_cast<const uint32_t*>(row_hi)[lane];
        const uint32_t lo = reinterpret_cast<const uint32_t*>(row_hi + plane)[lane];
        r.v[0] = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
        r.v[1] = __uint_as_float(hi & 0xffff0000u) + __uint_as_float(lo & 0xffff0000u);
    } else {
#pragma unroll
        for (int c = 0; c < V / 4; ++c) {
            const uint2 hi = reinterpret_cast<const uint2*>(row_hi + c * 128)[lane];
            const uint2 lo = reinterpret_cast<const uint2*>(row_hi + plane + c * 128)[lane];
            r.v[4 * c + 0] = __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
            r.v[4 * c + 1] = __uint_as_float(hi.x & 0xffff0000u) + __uint_as_float(lo.x & 0xffff0000u);
            r.v[4 * c + 2] = __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
            r.v[4 * c + 3] = __uint_as_float(hi.y & 0xffff0000u) + __uint_as_float(lo.y & 0xffff0000u);
        }
    }
    return r;
}

// storage-type generic front ends (T = float or __nv_bfloat16)
template <int V, typename T>
__device__ __forceinline__ LaneVec<V> lane_load_t(const T* __restrict__ row, int lane) {
    if constexpr (sizeof(T) == 2) return lane_load_bf16<V>(reinterpret_cast<const __nv_bfloat16*>(row), lane);
    else return lane_load<V>(reinterpret_cast<const float*>(row), lane);
}
// coherent loads for rows that the same kernel also writes
template <int V, typename T>
__device__ __forceinline__ LaneVec<V> lane_load_rw_t(const T* row, int lane) {
    if constexpr (sizeof(T) == 2) {
        LaneVec<V> r;
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(row);
        if constexpr (V == 2) {
            __nv_bfloat162 t = reinterpret_cast<const __nv_bfloat162*>(p)[lane];
            r.v[0] = __low2float(t); r.v[1] = __high2float(t);
        } else {
#pragma unroll
            for (int c = 0; c < V / 4; ++c) {
                uint2 raw = reinterpret_cast<const uint2*>(p + c * 128)[lane];
                __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
                __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
                r.v[4 * c + 0] = __low2float(a); r.v[4 * c + 1] = __high2float(a);
                r.v[4 * c + 2] = __low2float(b); r.v[4 * c + 3] = __high2float(b);
            }
        }
        return r;
    } else {
        return lane_load_rw<V>(reinterpret_cast<const float*>(row), lane);
    }
}
template <int V, typename T>
__device__ __forceinline__ void lane_store_t(T* __restrict__ row, int lane, const LaneVec<V>& r) {
    if constexpr (sizeof(T) == 2) lane_store_bf16<V>(reinterpret_cast<__nv_bfloat16*>(row), lane, r);
    else lane_store<V>(reinterpret_cast<float*>(row), lane, r);
}

// -------------------------------------------------------------------------------- Philox
// Philox4x32-10 counter-based generator; the same function is restated in numpy in
// diffusionsat_b200/philox.py so that host-side tests can inject identical noise.
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 23 random mantissa bits -> [0,1)
__host__ __device__ inline float u32_to_unit_float(uint32_t x) {
    uint32_t bits = (x & 0x7fffffu) | 0x3f800000u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(bits) - 1.0f;
#else
    float f; memcpy(&f, &bits, 4); return f - 1.0f;
#endif
}

enum PhiloxStream : uint32_t { STREAM_NORMAL = 0, STREAM_UNIFORM = 1, STREAM_LABEL = 2 };

// counter = (element lo, element hi, step<<16 | round, stream); key = seed
__device__ __forceinline__ Philox4 noise_draw(uint64_t seed, uint64_t element, uint32_t step, uint32_t round,
                                              uint32_t stream) {
    return philox4x32_10((uint32_t)element, (uint32_t)(element >> 32), (step << 16) | round, stream,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u1 = fmaxf(u32_to_unit_float(a), 1.0e-7f);
    float ang = 6.283185307179586f * u32_to_unit_float(b);
    float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(ang, &s, &c);
    n0 = s * rad; n1 = c * rad;
}

}  // namespace dsat
