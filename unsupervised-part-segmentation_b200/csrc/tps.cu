// Thin-plate-spline equivariance warp: parameter transform, 11x11 solve, fused
// grid-evaluation + bilinear gather (forward) and scatter-add (backward).
// Reference: baselines/unsupervised-disentangling/transformations.py:59-77,93-244
// (the vendored copy of eddata.utils.tps called from cub/code/SB_model48i/model.py:300-309).
#include "common.cuh"

namespace ups {

// ------------------------------------------------------------------ a2: make_input_tps_param
__global__ void tps_input_param_kernel(const float* __restrict__ coord, const float* __restrict__ vector,
                                       const float* __restrict__ offset, const float* __restrict__ offset_2,
                                       const float* __restrict__ t_scal, const float* __restrict__ rot,
                                       float* __restrict__ t_vector, int N) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    float c[16], v[16], o[2], o2[2], ts[2], r[4], tv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i] = coord[b * 16 + i]; v[i] = vector[b * 16 + i]; }
#pragma unroll
    for (int i = 0; i < 2; ++i) { o[i] = offset[b * 2 + i]; o2[i] = offset_2[b * 2 + i]; ts[i] = t_scal[b * 2 + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = rot[b * 4 + i];
    tps_input_param(c, v, o, o2, ts, r, tv);
#pragma unroll
    for (int i = 0; i < 16; ++i) t_vector[b * 16 + i] = tv[i];
}

// ------------------------------------------------------------------ a3(ii): _solve_system
// One thread per sample; the 11x13 augmented system lives in shared memory interleaved by
// thread (element e of thread t at [e*SOLVE_TPB + t]) so every access is conflict-free.
constexpr int SOLVE_TPB = 32;
struct SmemAcc {
    double* base;
    __device__ __forceinline__ double& operator()(int i, int j) const { return base[(i * 13 + j) * SOLVE_TPB]; }
};

__global__ void __launch_bounds__(SOLVE_TPB) tps_solve_kernel(const float* __restrict__ coord,
                                                              const float* __restrict__ vector,
                                                              float* __restrict__ T, int N) {
    __shared__ double A[11 * 13 * SOLVE_TPB];
    const int b = blockIdx.x * SOLVE_TPB + threadIdx.x;
    if (b >= N) return;
    float c[16], v[16], t[22];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i] = coord[b * 16 + i]; v[i] = vector[b * 16 + i]; }
    tps_solve(c, v, SmemAcc{A + threadIdx.x}, t);
#pragma unroll
    for (int i = 0; i < 22; ++i) T[b * 22 + i] = t[i];
}

// ------------------------------------------------------------------ a3(iii-v): grid + sample
constexpr int WARP_TPB = 128;  // threads per CTA
constexpr int WARP_PPT = 4;    // output pixels per thread (strided by the CTA width)

struct SampleConst {
    float T[22];
    float qx[8], qy[8];
    float sy, sx, my, mx;  // optional move/scal branch (transformations.py:202-208)
};

__device__ __forceinline__ void load_sample_const(SampleConst& k, float* sm, const float* coord, const float* T,
                                                  const float* move, const float* scal, int b) {
    // 22 T + 16 coord + 4 move/scal through shared memory, then into registers
    const int t = threadIdx.x;
    if (t < 22) sm[t] = T[b * 22 + t];
    else if (t < 38) sm[t] = coord[b * 16 + (t - 22)];
    else if (t < 40) sm[t] = scal ? scal[b * 2 + (t - 38)] : 1.0f;
    else if (t < 42) sm[t] = move ? move[b * 2 + (t - 40)] : 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 22; ++i) k.T[i] = sm[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { k.qx[i] = sm[22 + 2 * i + 1]; k.qy[i] = sm[22 + 2 * i + 0]; }  // ::-1 flip
    k.sy = sm[38]; k.sx = sm[39]; k.my = sm[40]; k.mx = sm[41];
}

__device__ __forceinline__ void sample_position(const SampleConst& k, bool has_move, int i, int j, float step_h,
                                                float step_w, float& x_s, float& y_s) {
    tps_coords(k.T, k.qx, k.qy, lin_at(j, step_w), lin_at(i, step_h), x_s, y_s);
    if (has_move) {
        y_s = __fadd_rn(__fmul_rn(y_s, k.sy), k.my);
        x_s = __fadd_rn(__fmul_rn(x_s, k.sx), k.mx);
    }
}

template <int C>
__global__ void __launch_bounds__(WARP_TPB) tps_warp_fwd_kernel(const float* __restrict__ U,
                                                                const float* __restrict__ coord,
                                                                const float* __restrict__ T,
                                                                const float* __restrict__ move,
                                                                const float* __restrict__ scal,
                                                                float* __restrict__ out, float* __restrict__ mesh,
                                                                int H, int W, int Crt, int oh, int ow) {
    __shared__ float sm_const[42];
    extern __shared__ float sm_out[];  // WARP_TPB * C floats
    const int Cc = (C > 0) ? C : Crt;
    const int b = blockIdx.y;
    SampleConst k;
    load_sample_const(k, sm_const, coord, T, move, scal, b);
    const bool has_move = (move != nullptr);
    const float step_w = lin_step(ow), step_h = lin_step(oh);
    const int OP = oh * ow;
    const float* Ub = U + (size_t)b * H * W * Cc;
    const int tile0 = blockIdx.x * (WARP_TPB * WARP_PPT);
#pragma unroll 1
    for (int it = 0; it < WARP_PPT; ++it) {
        const int base = tile0 + it * WARP_TPB;
        if (base >= OP) break;
        const int pix = base + threadIdx.x;
        const bool live = pix < OP;
        if (live) {
            const int i = pix / ow, j = pix - i * ow;
            float x_s, y_s;
            sample_position(k, has_move, i, j, step_h, step_w, x_s, y_s);
            if (mesh) reinterpret_cast<float2*>(mesh)[(size_t)b * OP + pix] = make_float2(y_s, x_s);
            const Bilinear s = bilinear_stencil(x_s, y_s, W, H);
            const float* pa = Ub + (size_t)(s.y0 * W + s.x0) * Cc;
            const float* pb = Ub + (size_t)(s.y1 * W + s.x0) * Cc;
            const float* pc = Ub + (size_t)(s.y0 * W + s.x1) * Cc;
            const float* pd = Ub + (size_t)(s.y1 * W + s.x1) * Cc;
            if (C > 0) {
#pragma unroll
                for (int c = 0; c < (C > 0 ? C : 1); ++c)
                    sm_out[threadIdx.x * Cc + c] = bilinear_mix(s, __ldg(pa + c), __ldg(pb + c), __ldg(pc + c), __ldg(pd + c));
            } else {
                for (int c = 0; c < Cc; ++c)
                    sm_out[threadIdx.x * Cc + c] = bilinear_mix(s, __ldg(pa + c), __ldg(pb + c), __ldg(pc + c), __ldg(pd + c));
            }
        }
        __syncthreads();
        // coalesced write of the tile: WARP_TPB*C contiguous floats
        const int n_live = min(WARP_TPB, OP - base) * Cc;
        float* ob = out + ((size_t)b * OP + base) * Cc;
        for (int e = threadIdx.x; e < n_live; e += WARP_TPB) __stcs(ob + e, sm_out[e]);
        __syncthreads();
    }
}

template <int C>
__global__ void __launch_bounds__(WARP_TPB) tps_warp_bwd_kernel(const float* __restrict__ g_out,
                                                                const float* __restrict__ coord,
                                                                const float* __restrict__ T,
                                                                const float* __restrict__ move,
                                                                const float* __restrict__ scal, float* __restrict__ dU,
                                                                int H, int W, int Crt, int oh, int ow) {
    __shared__ float sm_const[42];
    extern __shared__ float sm_g[];  // WARP_TPB * C floats
    const int Cc = (C > 0) ? C : Crt;
    const int b = blockIdx.y;
    SampleConst k;
    load_sample_const(k, sm_const, coord, T, move, scal, b);
    const bool has_move = (move != nullptr);
    const float step_w = lin_step(ow), step_h = lin_step(oh);
    const int OP = oh * ow;
    float* Ub = dU + (size_t)b * H * W * Cc;
    const int tile0 = blockIdx.x * (WARP_TPB * WARP_PPT);
#pragma unroll 1
    for (int it = 0; it < WARP_PPT; ++it) {
        const int base = tile0 + it * WARP_TPB;
        if (base >= OP) break;
        const int n_live = min(WARP_TPB, OP - base) * Cc;
        const float* gb = g_out + ((size_t)b * OP + base) * Cc;
        for (int e = threadIdx.x; e < n_live; e += WARP_TPB) sm_g[e] = __ldcs(gb + e);
        __syncthreads();
        const int pix = base + threadIdx.x;
        if (pix < OP) {
            const int i = pix / ow, j = pix - i * ow;
            float x_s, y_s;
            sample_position(k, has_move, i, j, step_h, step_w, x_s, y_s);
            const Bilinear s = bilinear_stencil(x_s, y_s, W, H);
            float* pa = Ub + (size_t)(s.y0 * W + s.x0) * Cc;
            float* pb = Ub + (size_t)(s.y1 * W + s.x0) * Cc;
            float* pc = Ub + (size_t)(s.y0 * W + s.x1) * Cc;
            float* pd = Ub + (size_t)(s.y1 * W + s.x1) * Cc;
            for (int c = 0; c < Cc; ++c) {
                const float g = sm_g[threadIdx.x * Cc + c];
                // out-of-range samples have coinciding clipped corners whose weights cancel:
                // keep all four adds so that the sum matches autodiff of the forward exactly.
                atomicAdd(pa + c, s.wa * g);
                atomicAdd(pb + c, s.wb * g);
                atomicAdd(pc + c, s.wc * g);
                atomicAdd(pd + c, s.wd * g);
            }
        }
        __syncthreads();
    }
}

}  // namespace ups

using namespace ups;

extern "C" int ups_tps_input_param(const float* coord, const float* vector, const float* offset,
                                   const float* offset_2, const float* t_scal, const float* rot_mat, float* t_vector,
                                   int N, void* stream) {
    UPS_REQUIRE(coord && vector && offset && offset_2 && t_scal && rot_mat && t_vector, "tps_input_param: null pointer");
    UPS_REQUIRE(N >= 0, "tps_input_param: N=%d", N);
    if (N == 0) return UPS_OK;
    tps_input_param_kernel<<<(unsigned)cdiv(N, 64), 64, 0, as_stream(stream)>>>(coord, vector, offset, offset_2, t_scal,
                                                                                rot_mat, t_vector, N);
    return after_launch("tps_input_param_kernel");
}

extern "C" int ups_tps_solve(const float* coord, const float* vector, float* T, int N, void* stream) {
    UPS_REQUIRE(coord && vector && T, "tps_solve: null pointer");
    UPS_REQUIRE(N >= 0, "tps_solve: N=%d", N);
    if (N == 0) return UPS_OK;
    tps_solve_kernel<<<(unsigned)cdiv(N, SOLVE_TPB), SOLVE_TPB, 0, as_stream(stream)>>>(coord, vector, T, N);
    return after_launch("tps_solve_kernel");
}

static int check_warp_args(const void* U, const void* coord, const void* T, const void* move, const void* scal,
                           const void* out, int N, int H, int W, int C, int oh, int ow) {
    UPS_REQUIRE(U && coord && T && out, "tps_warp: null pointer");
    UPS_REQUIRE((move == nullptr) == (scal == nullptr), "tps_warp: move and scal must be given together");
    UPS_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0 && oh > 1 && ow > 1, "tps_warp: bad dims N=%d H=%d W=%d C=%d out=%dx%d",
                N, H, W, C, oh, ow);
    UPS_REQUIRE(N <= 65535, "tps_warp: N=%d exceeds grid.y limit; split the batch", N);
    UPS_REQUIRE((long long)H * W * C < (1ll << 31) && (long long)oh * ow * C < (1ll << 31), "tps_warp: image too large");
    UPS_REQUIRE(C <= 64, "tps_warp: C=%d > 64 unsupported", C);
    return UPS_OK;
}

extern "C" int ups_tps_warp_fwd(const float* U, const float* coord, const float* T, const float* move,
                                const float* scal, float* out, float* mesh, int N, int H, int W, int C, int out_h,
                                int out_w, void* stream) {
    int rc = check_warp_args(U, coord, T, move, scal, out, N, H, W, C, out_h, out_w);
    if (rc) return rc;
    if (N == 0) return UPS_OK;
    UPS_REQUIRE(mesh == nullptr || (reinterpret_cast<uintptr_t>(mesh) & 7u) == 0, "tps_warp_fwd: mesh must be 8-byte aligned");
    dim3 grid((unsigned)cdiv((long long)out_h * out_w, WARP_TPB * WARP_PPT), (unsigned)N);
    const size_t sm = (size_t)WARP_TPB * C * sizeof(float);
    if (C == 3)
        tps_warp_fwd_kernel<3><<<grid, WARP_TPB, sm, as_stream(stream)>>>(U, coord, T, move, scal, out, mesh, H, W, C, out_h, out_w);
    else
        tps_warp_fwd_kernel<0><<<grid, WARP_TPB, sm, as_stream(stream)>>>(U, coord, T, move, scal, out, mesh, H, W, C, out_h, out_w);
    return after_launch("tps_warp_fwd_kernel");
}

extern "C" int ups_tps_warp_bwd(const float* g_out, const float* coord, const float* T, const float* move,
                                const float* scal, float* dU, int N, int H, int W, int C, int out_h, int out_w,
                                void* stream) {
    int rc = check_warp_args(g_out, coord, T, move, scal, dU, N, H, W, C, out_h, out_w);
    if (rc) return rc;
    if (N == 0) return UPS_OK;
    UPS_CUDA(cudaMemsetAsync(dU, 0, (size_t)N * H * W * C * sizeof(float), as_stream(stream)));
    dim3 grid((unsigned)cdiv((long long)out_h * out_w, WARP_TPB * WARP_PPT), (unsigned)N);
    const size_t sm = (size_t)WARP_TPB * C * sizeof(float);
    if (C == 3)
        tps_warp_bwd_kernel<3><<<grid, WARP_TPB, sm, as_stream(stream)>>>(g_out, coord, T, move, scal, dU, H, W, C, out_h, out_w);
    else
        tps_warp_bwd_kernel<0><<<grid, WARP_TPB, sm, as_stream(stream)>>>(g_out, coord, T, move, scal, dU, H, W, C, out_h, out_w);
    return after_launch("tps_warp_bwd_kernel");
}
