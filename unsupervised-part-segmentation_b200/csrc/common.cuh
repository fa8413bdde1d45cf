// Shared host/device helpers for libups_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ups_b200.h"
#include "canon_math.cuh"

namespace ups {

constexpr int NUM_SMS = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
void count_launch();
int after_launch(const char* what);  // cudaGetLastError -> status, counts the launch

int fused_pix_per_cta(int B, int P);  // step_fused.cu
size_t pool_ws_bytes(int B, int P, int KF);  // parts_ops.cu
size_t decode_bwd_tma_ws_bytes(int B, int P, int K, int F);  // step_decode_bwd_tma.cu
size_t moments_ws_bytes(int B, int P, int K);  // stats_ops.cu

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define UPS_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::ups::set_error(__VA_ARGS__);     \
            return UPS_E_INVALID;              \
        }                                      \
    } while (0)

#define UPS_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::ups::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
            return UPS_E_CUDA;                                                      \
        }                                                                           \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

#if defined(__CUDACC__)
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming variants: read-once inputs / write-once outputs should not displace reusable lines
__device__ __forceinline__ float4 ld4_stream(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4_stream(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

template <int W>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ int group_min_int(int v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ int group_sum_int(int v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// canonical adjacent-pair tree across W lanes (oracle/canon.py::sum_tree): NOT contracted
template <int W>
__device__ __forceinline__ float group_sum_canon(float v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Softmax of one pixel spread over LPP lanes, 4 consecutive parts per lane (K = 4*LPP), in the
// canonical order.  Returns probabilities; pmax = max prob of the pixel; arg = first index of
// the max; nmax = number of tied maxima.  `c` = lane's chunk index within the pixel.
template <int LPP>
__device__ __forceinline__ float4 softmax4(float4 v, int c, float& pmax, int& arg, int& nmax) {
    float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    m = group_max<LPP>(m);
    float4 e;
    e.x = exp_canon(__fsub_rn(v.x, m));
    e.y = exp_canon(__fsub_rn(v.y, m));
    e.z = exp_canon(__fsub_rn(v.z, m));
    e.w = exp_canon(__fsub_rn(v.w, m));
    float s = __fadd_rn(__fadd_rn(e.x, e.y), __fadd_rn(e.z, e.w));
    s = group_sum_canon<LPP>(s);
    const float rs = __frcp_rn(s);
    float4 p;
    p.x = __fmul_rn(e.x, rs);
    p.y = __fmul_rn(e.y, rs);
    p.z = __fmul_rn(e.z, rs);
    p.w = __fmul_rn(e.w, rs);
    pmax = group_max<LPP>(fmaxf(fmaxf(p.x, p.y), fmaxf(p.z, p.w)));
    int a = 1 << 30;
    a = (p.w == pmax) ? 4 * c + 3 : a;
    a = (p.z == pmax) ? 4 * c + 2 : a;
    a = (p.y == pmax) ? 4 * c + 1 : a;
    a = (p.x == pmax) ? 4 * c + 0 : a;
    arg = group_min_int<LPP>(a);
    const int n = (p.x == pmax) + (p.y == pmax) + (p.z == pmax) + (p.w == pmax);
    nmax = group_sum_int<LPP>(n);
    return p;
}

// straight_through_estimator(hard_max(p), p) for the lane's 4 parts
__device__ __forceinline__ float4 hard_st4(float4 p, float pmax) {
    float4 h;
    h.x = st_value(p.x == pmax ? 1.0f : 0.0f, p.x);
    h.y = st_value(p.y == pmax ? 1.0f : 0.0f, p.y);
    h.z = st_value(p.z == pmax ? 1.0f : 0.0f, p.z);
    h.w = st_value(p.w == pmax ? 1.0f : 0.0f, p.w);
    return h;
}
#endif

}  // namespace ups
