// Mask statistics on the part probabilities: the consumers of the [B,H,W,K]
// probability maps inside the training step, same HW reductions as the pooling kernels.
//   probs_to_mu_sigma(probs, scaling_factor)  cub/code/nn.py:1541-1587  (call sites cub/code/SB_model48i/model.py:440,458,689)
//   categorical_kl(probs)                     cub/code/SB_model48i/model.py:21-25 (call site :659-661)
// Both are HBM-bound single passes over probs; reductions are two-stage and fixed-order
// (bit-reproducible run to run, no atomics).
#include "common.cuh"

namespace ups {

constexpr int ST_TPB = 256;
constexpr int NMOM = 5;  // sum p*y, p*x, p*y*y, p*y*x, p*x*x

int moments_splits(int B, int P) {
    // ~16 CTAs per SM in total (several waves: short tail), each with at least 1024 pixels
    long long want = cdiv(16ll * NUM_SMS, B > 0 ? B : 1);
    const long long maxs = cdiv(P, 1024);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}
size_t moments_ws_bytes(int B, int P, int K) { return (size_t)B * moments_splits(B, P) * K * NMOM * sizeof(float) + 256; }

// thread t < TU = (TPB/K)*K owns part k = t % K and pixel phase t / K: element index (p0 + phase)*K + k = p0*K + t,
// so the scalar loads of a warp are contiguous for any K (25 included)
__global__ void __launch_bounds__(ST_TPB) mask_moments_partial_kernel(const float* __restrict__ probs,
                                                                      float* __restrict__ partial, int P, int H, int W,
                                                                      int K, int pix_per_cta) {
    extern __shared__ float red[];  // [PP][K][NMOM]
    const int b = blockIdx.y, t = threadIdx.x;
    const int PP = ST_TPB / K, TU = PP * K;
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const float step_h = lin_step(H), step_w = lin_step(W);
    if (t < TU) {
        const int k = t % K, ph = t / K;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        const float* src = probs + (size_t)b * P * K + k;
        for (int p = p_begin + ph; p < p_end; p += PP) {
            const float v = __ldcs(src + (size_t)p * K);
            const int i = p / W, j = p - i * W;
            const float y = lin_at(i, step_h), x = lin_at(j, step_w);
            const float vy = v * y, vx = v * x;
            a0 += vy; a1 += vx;
            a2 = fmaf(vy, y, a2); a3 = fmaf(vy, x, a3); a4 = fmaf(vx, x, a4);
        }
        float* r = red + (ph * K + k) * NMOM;
        r[0] = a0; r[1] = a1; r[2] = a2; r[3] = a3; r[4] = a4;
    }
    __syncthreads();
    for (int e = t; e < K * NMOM; e += ST_TPB) {
        float s = 0.f;
        for (int ph = 0; ph < PP; ++ph) s += red[ph * K * NMOM + e];
        partial[((size_t)b * gridDim.x + blockIdx.x) * (K * NMOM) + e] = s;
    }
}

// Fast path for K = 4*LPP: LPP lanes per pixel, one float4 (4 parts) per lane, 4 pixels per lane in flight.
// Lane-private accumulators, then a fixed shuffle tree over the warp's pixel groups and a fixed-order sum
// over the CTA's warps.
template <int LPP>
__global__ void __launch_bounds__(ST_TPB) mask_moments_partial_vec_kernel(const float* __restrict__ probs,
                                                                          float* __restrict__ partial, int P, int H,
                                                                          int W, int pix_per_cta) {
    constexpr int K = 4 * LPP, PW = 32 / LPP, NW = ST_TPB / 32, UN = 4;
    __shared__ float red[NW][K * NMOM];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane & (LPP - 1), q = lane / LPP;
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const float step_h = lin_step(H), step_w = lin_step(W);
    float a[4][NMOM];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int m = 0; m < NMOM; ++m) a[i][m] = 0.f;
    const float* src = probs + (size_t)b * P * K + 4 * c;
    // register double buffer: the next 4 pixels per lane are in flight while the current ones are reduced
    auto load = [&](float4 (&v)[UN], int p0) {
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int p = p0 + u * PW + q;
            v[u] = p < p_end ? ld4_stream(src + (size_t)p * K) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    float4 v[UN], nv[UN];
    int p0 = p_begin + warp * PW * UN;
    load(v, p0);
    for (; p0 < p_end; p0 += NW * PW * UN) {
        load(nv, p0 + NW * PW * UN);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int p = min(p0 + u * PW + q, P - 1);
            const int i = p / W, j = p - i * W;
            const float y = lin_at(i, step_h), x = lin_at(j, step_w);
            const float yy = y * y, yx = y * x, xx = x * x;
            const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[e][0] = fmaf(vv[e], y, a[e][0]); a[e][1] = fmaf(vv[e], x, a[e][1]);
                a[e][2] = fmaf(vv[e], yy, a[e][2]); a[e][3] = fmaf(vv[e], yx, a[e][3]);
                a[e][4] = fmaf(vv[e], xx, a[e][4]);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) v[u] = nv[u];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int m = 0; m < NMOM; ++m) {
            float s = a[e][m];
#pragma unroll
            for (int o = LPP; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (q == 0) red[warp][(4 * c + e) * NMOM + m] = s;
        }
    __syncthreads();
    for (int e = threadIdx.x; e < K * NMOM; e += ST_TPB) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w][e];
        partial[((size_t)b * gridDim.x + blockIdx.x) * (K * NMOM) + e] = s;
    }
}

// one thread per (b, k): sum the splits in ascending order, then mu / sigma as nn.py:1577-1586
__global__ void mask_moments_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ scaling,
                                             float* __restrict__ moments, float* __restrict__ mu,
                                             float* __restrict__ sigma, int splits, int K, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / K;
    const int k = (int)(i % K);
    float m[NMOM];
#pragma unroll
    for (int c = 0; c < NMOM; ++c) {
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)b * splits + sp) * (K * NMOM) + k * NMOM + c];
        m[c] = s;
        moments[i * NMOM + c] = s;
    }
    const float s1 = scaling[i], s2 = s1 * s1;
    mu[i * 2 + 0] = m[0] * s1;
    mu[i * 2 + 1] = m[1] * s1;
    sigma[i * 4 + 0] = s2 * (m[2] - m[0] * m[0]);
    sigma[i * 4 + 1] = s2 * (m[3] - m[0] * m[1]);
    sigma[i * 4 + 2] = s2 * (m[3] - m[1] * m[0]);
    sigma[i * 4 + 3] = s2 * (m[4] - m[1] * m[1]);
}

// dprobs[b,p,k] = c0*y + c1*x + c2*y*y + c3*y*x + c4*x*x with per-(b,k) coefficients from the cotangents
__global__ void __launch_bounds__(ST_TPB) mask_moments_bwd_kernel(const float* __restrict__ g_mu,
                                                                  const float* __restrict__ g_sigma,
                                                                  const float* __restrict__ scaling,
                                                                  const float* __restrict__ moments,
                                                                  float* __restrict__ dprobs, int P, int H, int W, int K,
                                                                  int pix_per_cta) {
    const int b = blockIdx.y, t = threadIdx.x;
    const int PP = ST_TPB / K, TU = PP * K;
    if (t >= TU) return;
    const int k = t % K, ph = t / K;
    const size_t bk = (size_t)b * K + k;
    const float s1 = scaling[bk], s2 = s1 * s1;
    const float m0 = moments[bk * NMOM], m1 = moments[bk * NMOM + 1];
    const float g00 = g_sigma ? g_sigma[bk * 4] : 0.f, g01 = g_sigma ? g_sigma[bk * 4 + 1] + g_sigma[bk * 4 + 2] : 0.f,
                g11 = g_sigma ? g_sigma[bk * 4 + 3] : 0.f;
    const float gm0 = g_mu ? g_mu[bk * 2] : 0.f, gm1 = g_mu ? g_mu[bk * 2 + 1] : 0.f;
    const float c2 = s2 * g00, c3 = s2 * g01, c4 = s2 * g11;
    const float c0 = s1 * gm0 - s2 * (2.f * g00 * m0 + g01 * m1);
    const float c1 = s1 * gm1 - s2 * (2.f * g11 * m1 + g01 * m0);
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const float step_h = lin_step(H), step_w = lin_step(W);
    float* dst = dprobs + (size_t)b * P * K + k;
    for (int p = p_begin + ph; p < p_end; p += PP) {
        const int i = p / W, j = p - i * W;
        const float y = lin_at(i, step_h), x = lin_at(j, step_w);
        const float v = fmaf(y, fmaf(c2, y, fmaf(c3, x, c0)), x * fmaf(c4, x, c1));
        __stcs(dst + (size_t)p * K, v);
    }
}

template <int LPP>
__global__ void __launch_bounds__(ST_TPB) mask_moments_bwd_vec_kernel(const float* __restrict__ g_mu,
                                                                      const float* __restrict__ g_sigma,
                                                                      const float* __restrict__ scaling,
                                                                      const float* __restrict__ moments,
                                                                      float* __restrict__ dprobs, int P, int H, int W,
                                                                      int pix_per_cta) {
    constexpr int K = 4 * LPP, PW = 32 / LPP, NW = ST_TPB / 32;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = lane & (LPP - 1), q = lane / LPP;
    float c0[4], c1[4], c2[4], c3[4], c4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const size_t bk = (size_t)b * K + 4 * c + e;
        const float s1 = scaling[bk], s2 = s1 * s1;
        const float m0 = moments[bk * NMOM], m1 = moments[bk * NMOM + 1];
        const float g00 = g_sigma ? g_sigma[bk * 4] : 0.f, g01 = g_sigma ? g_sigma[bk * 4 + 1] + g_sigma[bk * 4 + 2] : 0.f,
                    g11 = g_sigma ? g_sigma[bk * 4 + 3] : 0.f;
        const float gm0 = g_mu ? g_mu[bk * 2] : 0.f, gm1 = g_mu ? g_mu[bk * 2 + 1] : 0.f;
        c2[e] = s2 * g00; c3[e] = s2 * g01; c4[e] = s2 * g11;
        c0[e] = s1 * gm0 - s2 * (2.f * g00 * m0 + g01 * m1);
        c1[e] = s1 * gm1 - s2 * (2.f * g11 * m1 + g01 * m0);
    }
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const float step_h = lin_step(H), step_w = lin_step(W);
    float* dst = dprobs + (size_t)b * P * K + 4 * c;
    for (int p = p_begin + warp * PW + q; p < p_end; p += NW * PW) {
        const int i = p / W, j = p - i * W;
        const float y = lin_at(i, step_h), x = lin_at(j, step_w);
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaf(y, fmaf(c2[e], y, fmaf(c3[e], x, c0[e])), x * fmaf(c4[e], x, c1[e]));
        st4_stream(dst + (size_t)p * K, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// ------------------------------------------------------------------ categorical KL
constexpr int KL_BLOCKS = NUM_SMS * 8;

__global__ void __launch_bounds__(ST_TPB) categorical_kl_partial_kernel(const float* __restrict__ probs,
                                                                        float* __restrict__ partial, long long n,
                                                                        float kf) {
    __shared__ float red[ST_TPB / 32];
    float acc = 0.f;
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * ST_TPB;
    for (long long i = (long long)blockIdx.x * ST_TPB + threadIdx.x; i < n4; i += stride) {
        const float4 p = ld4_stream(probs + 4 * i);
        // __logf (MUFU.LG2): abs error ~2^-21 on the log, far inside the 1e-4 / 1e-5 tolerance of the mean
        acc = fmaf(p.x, __logf(fmaf(kf, p.x, 1e-20f)), acc);
        acc = fmaf(p.y, __logf(fmaf(kf, p.y, 1e-20f)), acc);
        acc = fmaf(p.z, __logf(fmaf(kf, p.z, 1e-20f)), acc);
        acc = fmaf(p.w, __logf(fmaf(kf, p.w, 1e-20f)), acc);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {   // tail (n not a multiple of 4)
        const float p = probs[(n4 << 2) + threadIdx.x];
        acc = fmaf(p, __logf(fmaf(kf, p, 1e-20f)), acc);
    }
    acc = group_sum<32>(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < ST_TPB / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void categorical_kl_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblocks,
                                               float inv_count) {
    // one warp, fixed order: lane l sums blocks l, l+32, ... then a shuffle tree
    float s = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
    s = group_sum<32>(s);
    if (threadIdx.x == 0) out[0] = s * inv_count;
}

__global__ void __launch_bounds__(ST_TPB) categorical_kl_bwd_kernel(const float* __restrict__ probs,
                                                                    const float* __restrict__ g_out,
                                                                    float* __restrict__ dprobs, long long n, float kf,
                                                                    float inv_count) {
    const long long i = (long long)blockIdx.x * ST_TPB + threadIdx.x;
    const float g = g_out[0] * inv_count;
    const long long n4 = n >> 2;
    auto d = [&](float p) {
        const float u = fmaf(kf, p, 1e-20f);
        return g * (__logf(u) + __fdividef(kf * p, u));
    };
    if (i < n4) {
        const float4 p = ld4_stream(probs + 4 * i);
        st4_stream(dprobs + 4 * i, make_float4(d(p.x), d(p.y), d(p.z), d(p.w)));
    } else if (i == n4) {
        for (long long e = n4 << 2; e < n; ++e) dprobs[e] = d(probs[e]);
    }
}

}  // namespace ups

using namespace ups;

static int moments_checks(const char* what, int B, int H, int W, int K) {
    UPS_REQUIRE(B >= 0 && B <= 65535, "%s: B=%d out of range", what, B);
    UPS_REQUIRE(H > 1 && W > 1 && (long long)H * W < (1ll << 31), "%s: bad H=%d W=%d", what, H, W);
    UPS_REQUIRE(K >= 1 && K <= ST_TPB, "%s: K=%d not in [1, %d]", what, K, ST_TPB);
    return UPS_OK;
}

// ---------------------------------------------------------------- patch masks (draw_rect)
// out[n, i, j] = 1 inside the ph x pw rectangle centred on (cy, cx) = centers[n] (int32, pixels), 0 outside:
// rows [cy - ph/2, cy - ph/2 + ph), columns [cx - pw/2, cx - pw/2 + pw), clipped to the image.
namespace ups {
__global__ void draw_rect_kernel(const int* __restrict__ centers, float* __restrict__ out, int ph, int pw, int H, int W,
                                 long long n_total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const long long n = i / ((long long)H * W);
    const int r = (int)(i - n * H * W);
    const int y = r / W, x = r - y * W;
    const int y0 = centers[2 * n] - ph / 2, x0 = centers[2 * n + 1] - pw / 2;
    out[i] = (y >= y0 && y < y0 + ph && x >= x0 && x < x0 + pw) ? 1.0f : 0.0f;
}
}  // namespace ups

extern "C" int ups_draw_rect_fwd(const int* centers, float* out, int N, int ph, int pw, int H, int W, void* stream) {
    UPS_REQUIRE(N >= 0 && ph >= 0 && pw >= 0 && H > 0 && W > 0, "draw_rect: bad dims N=%d ph=%d pw=%d H=%d W=%d", N, ph, pw, H, W);
    if (N == 0) return UPS_OK;
    UPS_REQUIRE(centers && out, "draw_rect: null pointer");
    const long long n_total = (long long)N * H * W;
    UPS_REQUIRE(cdiv(n_total, 256) < (1ll << 31), "draw_rect: too large");
    ups::draw_rect_kernel<<<(unsigned)cdiv(n_total, 256), 256, 0, as_stream(stream)>>>(centers, out, ph, pw, H, W, n_total);
    return after_launch("draw_rect_kernel");
}

extern "C" int ups_mask_moments_fwd(const float* probs, const float* scaling, float* mu, float* sigma, float* moments,
                                    int B, int H, int W, int K, void* ws, size_t ws_bytes, void* stream) {
    UPS_REQUIRE(probs && scaling && mu && sigma && moments, "mask_moments_fwd: null pointer");
    if (int rc = moments_checks("mask_moments_fwd", B, H, W, K)) return rc;
    if (B == 0) return UPS_OK;
    const int P = H * W;
    const int splits = moments_splits(B, P);
    const size_t need = moments_ws_bytes(B, P, K);
    if (!ws || ws_bytes < need) { set_error("mask_moments_fwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    const int per = (int)cdiv(P, splits);
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    const dim3 grid(splits, B);
    if ((K == 4 || K == 8 || K == 16 || K == 32) && aligned16(probs)) {
        if (K == 4) mask_moments_partial_vec_kernel<1><<<grid, ST_TPB, 0, s>>>(probs, partial, P, H, W, per);
        else if (K == 8) mask_moments_partial_vec_kernel<2><<<grid, ST_TPB, 0, s>>>(probs, partial, P, H, W, per);
        else if (K == 16) mask_moments_partial_vec_kernel<4><<<grid, ST_TPB, 0, s>>>(probs, partial, P, H, W, per);
        else mask_moments_partial_vec_kernel<8><<<grid, ST_TPB, 0, s>>>(probs, partial, P, H, W, per);
    } else {
        const size_t sm = (size_t)(ST_TPB / K) * K * NMOM * sizeof(float);
        mask_moments_partial_kernel<<<grid, ST_TPB, sm, s>>>(probs, partial, P, H, W, K, per);
    }
    if (int rc = after_launch("mask_moments_partial_kernel")) return rc;
    const long long n = (long long)B * K;
    mask_moments_finalize_kernel<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(partial, scaling, moments, mu, sigma, splits, K, n);
    return after_launch("mask_moments_finalize_kernel");
}

extern "C" int ups_mask_moments_bwd(const float* g_mu, const float* g_sigma, const float* scaling, const float* moments,
                                    float* dprobs, int B, int H, int W, int K, void* stream) {
    UPS_REQUIRE(scaling && moments && dprobs, "mask_moments_bwd: null pointer");
    if (int rc = moments_checks("mask_moments_bwd", B, H, W, K)) return rc;
    if (B == 0) return UPS_OK;
    const int P = H * W;
    const int splits = moments_splits(B, P);
    const int per = (int)cdiv(P, splits);
    const dim3 grid(splits, B);
    cudaStream_t s = as_stream(stream);
    if ((K == 4 || K == 8 || K == 16 || K == 32) && aligned16(dprobs)) {
        if (K == 4) mask_moments_bwd_vec_kernel<1><<<grid, ST_TPB, 0, s>>>(g_mu, g_sigma, scaling, moments, dprobs, P, H, W, per);
        else if (K == 8) mask_moments_bwd_vec_kernel<2><<<grid, ST_TPB, 0, s>>>(g_mu, g_sigma, scaling, moments, dprobs, P, H, W, per);
        else if (K == 16) mask_moments_bwd_vec_kernel<4><<<grid, ST_TPB, 0, s>>>(g_mu, g_sigma, scaling, moments, dprobs, P, H, W, per);
        else mask_moments_bwd_vec_kernel<8><<<grid, ST_TPB, 0, s>>>(g_mu, g_sigma, scaling, moments, dprobs, P, H, W, per);
    } else {
        mask_moments_bwd_kernel<<<grid, ST_TPB, 0, s>>>(g_mu, g_sigma, scaling, moments, dprobs, P, H, W, K, per);
    }
    return after_launch("mask_moments_bwd_kernel");
}

extern "C" int ups_categorical_kl_fwd(const float* probs, float* out, long long n_pix, int K, void* ws, size_t ws_bytes,
                                      void* stream) {
    UPS_REQUIRE(probs && out, "categorical_kl_fwd: null pointer");
    UPS_REQUIRE(n_pix > 0 && K >= 1, "categorical_kl_fwd: n_pix=%lld K=%d", n_pix, K);
    UPS_REQUIRE(aligned16(probs), "categorical_kl_fwd: 16-byte alignment");
    const size_t need = (size_t)KL_BLOCKS * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("categorical_kl_fwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    const long long n = n_pix * K;
    long long blocks = cdiv(cdiv(n, 4), ST_TPB);
    if (blocks > KL_BLOCKS) blocks = KL_BLOCKS;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    categorical_kl_partial_kernel<<<(unsigned)blocks, ST_TPB, 0, s>>>(probs, partial, n, (float)K);
    if (int rc = after_launch("categorical_kl_partial_kernel")) return rc;
    categorical_kl_finalize_kernel<<<1, 32, 0, s>>>(partial, out, (int)blocks, 1.0f / (float)n_pix);
    return after_launch("categorical_kl_finalize_kernel");
}

extern "C" int ups_categorical_kl_bwd(const float* probs, const float* g_out, float* dprobs, long long n_pix, int K,
                                      void* stream) {
    UPS_REQUIRE(probs && g_out && dprobs, "categorical_kl_bwd: null pointer");
    UPS_REQUIRE(n_pix > 0 && K >= 1, "categorical_kl_bwd: n_pix=%lld K=%d", n_pix, K);
    UPS_REQUIRE(aligned16(probs) && aligned16(dprobs), "categorical_kl_bwd: 16-byte alignment");
    const long long n = n_pix * K;
    const long long blocks = cdiv((n >> 2) + 1, ST_TPB);
    categorical_kl_bwd_kernel<<<(unsigned)blocks, ST_TPB, 0, as_stream(stream)>>>(probs, g_out, dprobs, n, (float)K,
                                                                                  1.0f / (float)n_pix);
    return after_launch("categorical_kl_bwd_kernel");
}
