// hg_common.cuh — shared helpers of libhuygens_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include "../../include/huygens_cuda.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhuygens_cuda is written for sm_100a (B200) only"
#endif

namespace hg {

// thread-local error text behind hg_last_error()
char* err_buf();
hg_status fail(hg_status code, const char* fmt, ...);

#define HG_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return hg::fail(HG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,     \
                            cudaGetErrorString(e__));                                     \
    } while (0)

#define HG_TRY(expr)                          \
    do {                                      \
        hg_status s__ = (expr);               \
        if (s__ != HG_OK) return s__;         \
    } while (0)

inline size_t dtype_size(int dtype) { return dtype == HG_F32 ? 4 : 8; }

// true if p is a device (or managed) pointer usable from kernels of the current device
bool is_device_ptr(const void* p);

// ---------------------------------------------------------------------------------------------
// Arithmetic policies.  STRICT never lets the compiler contract (intrinsics with explicit
// rounding are never fused by nvcc); FMA uses the one fixed contraction pattern shared with
// oracle/huygens_oracle.hpp.  The mixdown term is `muladd(y, g, acc)` in both modes: fl(fl(y*g) + acc) strict,
// fma(y, g, acc) fused — accumulated in filter-index order (chain.cuh).  `acc` (always fused) is only used by
// the tolerance-mode scan.
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> struct Ar;

template <bool FMA> struct Ar<float, FMA> {
    typedef float T;
    static __device__ __forceinline__ T mul(T a, T b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __fsub_rn(a, b); }
    // fl(a*b + c) fused, or fl(fl(a*b) + c)
    static __device__ __forceinline__ T muladd(T a, T b, T c) {
        return FMA ? __fmaf_rn(a, b, c) : __fadd_rn(__fmul_rn(a, b), c);
    }
    // fl(a*b - c)
    static __device__ __forceinline__ T mulsub(T a, T b, T c) {
        return FMA ? __fmaf_rn(a, b, -c) : __fsub_rn(__fmul_rn(a, b), c);
    }
    static __device__ __forceinline__ T acc(T a, T b, T c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ T div(T a, T b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ T sqrt(T a) { return __fsqrt_rn(a); }
};

template <bool FMA> struct Ar<double, FMA> {
    typedef double T;
    static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ T muladd(T a, T b, T c) {
        return FMA ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
    }
    static __device__ __forceinline__ T mulsub(T a, T b, T c) {
        return FMA ? __fma_rn(a, b, -c) : __dsub_rn(__dmul_rn(a, b), c);
    }
    static __device__ __forceinline__ T acc(T a, T b, T c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ T div(T a, T b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ T sqrt(T a) { return __dsqrt_rn(a); }
};

// tests/filterbank.cpp:158-171 softclip(sample, width = 0.125), evaluated in double like the
// reference's `double softclip(double)` and rounded to T.  Device atan is within 2 ulp of the
// correctly rounded value (CUDA math API), so this path is tolerance-bounded, not bit-exact.
// __noinline__: one copy of the double-precision atan per kernel instead of one per unrolled call site — the hot
// kernels' code must stay small enough for the instruction caches (the recurrence warps stream ~10 KB loops).
template <typename T> __device__ __noinline__ T softclip_dev(T v) {
    const double PI_REF = 3.14159265359, width = 0.125;       // src/includes.h:30
    const double sample = (double)v;
    if (fabs(sample) < width) return v;
    const double sign = (0.0 < sample) - (sample < 0.0);
    const double gap = sample - sign * width;
    return (T)(sign * width + (1 - width) * 2.0 / PI_REF * atan(PI_REF * gap / (2 * (1 - width))));
}

// Post-processing of the finished bank sum, fused into the store of out[t] (the caller loops' last step):
//   HG_EPI_SOFTCLIP  out = softclip(dry * in + sum, width)   tests/filterbank.cpp:205-207 (the outer softclip(., 0.9))
//   HG_EPI_LIMITER   out = 2/PI * atan(dry * in + sum)       src/wave.h:150 `limiter` (tests/harmbank.cpp:76), evaluated
//                    directly — the reference reads it from a 4096-point interpolated table
// evaluated in double like the reference's callers and rounded to T; device atan: tolerance-bounded, not bit-exact.
struct Epilogue { int kind; double dry, width; };
template <typename T> __device__ __noinline__ T epilogue_dev(const Epilogue e, T sum, T in) {
    const double PI_REF = 3.14159265359;
    const double v = e.dry * (double)in + (double)sum;
    if (e.kind == 2) return (T)(2.0 / PI_REF * atan(v));
    if (fabs(v) < e.width) return (T)v;
    const double sign = (0.0 < v) - (v < 0.0);
    const double gap = v - sign * e.width;
    return (T)(sign * e.width + (1 - e.width) * 2.0 / PI_REF * atan(PI_REF * gap / (2 * (1 - e.width))));
}

// Transposed CTA reduction of one tile of per-thread partial sums.
//   rb[t * RS + tid], RS = NT + NW  (RS == NW mod 32: the per-sample STS of all threads and the
//   strided reads below are both bank-conflict free).  Warp w owns samples [w*SPW, (w+1)*SPW),
//   NW lanes per sample, each adding 32 of the NT partials in a fixed order, then a shuffle tree:
//   deterministic for a given NT.  out[t] is written for t < tn (tile of at most 32 samples).
template <typename T, int NT>
__device__ __forceinline__ void cta_tile_reduce(const T* rb, int tn, T* out, int out_step = 1) {
    constexpr int NW = NT / 32, RS = NT + NW, SPW = 32 / NW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ls = lane / NW, jq = lane % NW;
    const int t = warp * SPW + ls;
    const T* row = rb + t * RS + jq;
    T sum = (T)0;
    if (t < tn) {
        // four independent partial sums keep the dependent-add chain 8 deep instead of 32
        T s0 = (T)0, s1 = (T)0, s2 = (T)0, s3 = (T)0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            s0 += row[(i + 0) * NW]; s1 += row[(i + 1) * NW]; s2 += row[(i + 2) * NW]; s3 += row[(i + 3) * NW];
        }
        sum = (s0 + s1) + (s2 + s3);
    }
#pragma unroll
    for (int o = NW / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (jq == 0 && t < tn) out[(size_t)t * out_step] = sum;
}

// staging buffer that grows on demand (device memory)
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    hg_status reserve(size_t bytes);
    void release();
};

}  // namespace hg
