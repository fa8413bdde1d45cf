// Mask priors and mean-field sampling around the part-map softmax: the other per-step
// consumers / producers of the [B,H,W,K] logits and probabilities.
//   mumford_shah / edge_set / tf_squared_grad      cub/code/nn.py:1357-1392   (call site cub/code/SB_model48i/model.py:744-769)
//   MeanFieldDistribution.sample / kl / kl_improper_gmrf / kl_tv   cub/code/nn.py:1395-1451   (model.py:413-421,1071)
//   weak cross entropy of the logits with their own hard / soft assignment   cub/code/SB_model48i/model.py:667-681
//   mask2rgb / mask2hotmask                         cub/code/nn.py:2067-2089   (eval_01.py:276-281)
// All of them are single HBM-bound passes (3-point stencils read their neighbours through L1/L2).  Elementwise
// maps use one correctly rounded operation per step of the reference's expression, so they are bit-identical to
// oracle/priors.py; spatial sums are two-stage and fixed-order (bit-reproducible run to run, no atomics).
#include "common.cuh"

namespace ups {

constexpr int PR_TPB = 256;
constexpr int MS_NSUM = 4;  // sums over (h,w) of r, smoothness_cost, contour_cost, x

int priors_splits(int B, int P) {
    long long want = cdiv(16ll * NUM_SMS, B > 0 ? B : 1);  // ~16 CTAs per SM in total
    const long long maxs = cdiv(P, 512);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}
size_t mumford_shah_ws_bytes(int B, int P, int K) {
    return (size_t)B * priors_splits(B, P) * K * MS_NSUM * sizeof(float) + 256;
}
constexpr int PR_BLOCKS = NUM_SMS * 8;
size_t priors_scalar_ws_bytes() { return (size_t)PR_BLOCKS * 4 * sizeof(float) + 256; }

template <int VEC>
struct Vec {
    float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> vload(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = ld4(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> vzero() {
    Vec<VEC> r;
#pragma unroll
    for (int e = 0; e < VEC; ++e) r.v[e] = 0.f;
    return r;
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const Vec<VEC>& r) {
    if constexpr (VEC == 4) st4_stream(p, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    else __stcs(p, r.v[0]);
}

// g = gx*gx + gy*gy with gx = 0.25*(a - right), gy = 0.25*(a - down)  (nn.py:1357-1378), every step rounded once
__device__ __forceinline__ float sq_grad(float a, float right, float down, float& gx, float& gy) {
    gx = __fmul_rn(0.25f, __fsub_rn(a, right));
    gy = __fmul_rn(0.25f, __fsub_rn(a, down));
    return __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
}

// ------------------------------------------------------------------ Mumford-Shah forward
// Thread t < TU = (TPB/KV)*KV owns chunk t % KV (VEC consecutive parts) and pixel phase t / KV, so that the
// loads of a warp are contiguous for any K.  Any of r / smooth / contour / edges / partial may be null.
template <int VEC>
__global__ void __launch_bounds__(PR_TPB) mumford_shah_fwd_kernel(const float* __restrict__ x, float alpha, float lam,
                                                                  float thr, float* __restrict__ r,
                                                                  float* __restrict__ smooth, float* __restrict__ contour,
                                                                  float* __restrict__ edges, float* __restrict__ partial,
                                                                  int P, int H, int W, int K, int pix_per_cta) {
    extern __shared__ float red[];  // [PP][K][MS_NSUM] when partial != null
    const int b = blockIdx.y, t = threadIdx.x;
    const int KV = K / VEC, PP = PR_TPB / KV, TU = PP * KV;
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const size_t base = (size_t)b * P * K;
    if (t < TU) {
        const int c = t % KV, ph = t / KV;
        float acc[VEC][MS_NSUM];
#pragma unroll
        for (int e = 0; e < VEC; ++e)
#pragma unroll
            for (int m = 0; m < MS_NSUM; ++m) acc[e][m] = 0.f;
        for (int p = p_begin + ph; p < p_end; p += PP) {
            const int i = p / W, j = p - i * W;
            const size_t o = base + (size_t)p * K + c * VEC;
            const Vec<VEC> a = vload<VEC>(x + o);
            const Vec<VEC> rt = j + 1 < W ? vload<VEC>(x + o + K) : vzero<VEC>();
            const Vec<VEC> dn = i + 1 < H ? vload<VEC>(x + o + (size_t)W * K) : vzero<VEC>();
            Vec<VEC> vr, vs, vc, ve;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float gx, gy;
                const float g = sq_grad(a.v[e], rt.v[e], dn.v[e], gx, gy);
                const float ag = __fmul_rn(alpha, g);
                vr.v[e] = ag <= lam ? ag : lam;
                vs.v[e] = ag < lam ? vr.v[e] : 0.f;
                vc.v[e] = ag >= lam ? vr.v[e] : 0.f;
                ve.v[e] = g > thr ? 1.f : 0.f;
                acc[e][0] += vr.v[e]; acc[e][1] += vs.v[e]; acc[e][2] += vc.v[e]; acc[e][3] += a.v[e];
            }
            if (r) vstore<VEC>(r + o, vr);
            if (smooth) vstore<VEC>(smooth + o, vs);
            if (contour) vstore<VEC>(contour + o, vc);
            if (edges) vstore<VEC>(edges + o, ve);
        }
        if (partial) {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
#pragma unroll
                for (int m = 0; m < MS_NSUM; ++m) red[(ph * K + c * VEC + e) * MS_NSUM + m] = acc[e][m];
        }
    }
    if (!partial) return;
    __syncthreads();
    for (int e = t; e < K * MS_NSUM; e += PR_TPB) {
        float s = 0.f;
        for (int ph = 0; ph < PP; ++ph) s += red[ph * K * MS_NSUM + e];
        partial[((size_t)b * gridDim.x + blockIdx.x) * (K * MS_NSUM) + e] = s;
    }
}

// sums [B, 4, K]: splits added in ascending order
__global__ void mumford_shah_finalize_kernel(const float* __restrict__ partial, float* __restrict__ sums, int splits,
                                             int K, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*K*4 in partial's (k, m) order
    if (i >= n) return;
    const long long b = i / (K * MS_NSUM);
    const int km = (int)(i % (K * MS_NSUM)), k = km / MS_NSUM, m = km % MS_NSUM;
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)b * splits + sp) * (K * MS_NSUM) + km];
    sums[((size_t)b * MS_NSUM + m) * K + k] = s;
}

// ------------------------------------------------------------------ Mumford-Shah backward
// dx[p] = 0.5*G(p)*(gx(p)+gy(p)) - 0.5*G(left)*gx(left) - 0.5*G(up)*gy(up) + g_sum_x, with
// G(q) = alpha*(g_r(q)*[ag<=lam] + g_smooth(q)*[ag<lam] + g_contour(q)*[ag==lam]) (tf.minimum / tf.where gradients),
// each cotangent = elementwise map (optional) + per-(b,k) cotangent of the spatial sum (optional).
template <int VEC>
struct MsCot {
    const float *g_r, *g_s, *g_c;  // elementwise, nullable
    float sr[VEC], ss[VEC], sc[VEC], sx[VEC];
};

template <int VEC>
__device__ __forceinline__ void ms_weight(const MsCot<VEC>& ct, size_t o, const Vec<VEC>& a, const Vec<VEC>& rt,
                                          const Vec<VEC>& dn, float alpha, float lam, float (&wx)[VEC],
                                          float (&wy)[VEC]) {
    // wx = 0.5*G*gx, wy = 0.5*G*gy at the pixel whose centre / right / down values are a / rt / dn
    Vec<VEC> er = vzero<VEC>(), es = vzero<VEC>(), ec = vzero<VEC>();
    if (ct.g_r) er = vload<VEC>(ct.g_r + o);
    if (ct.g_s) es = vload<VEC>(ct.g_s + o);
    if (ct.g_c) ec = vload<VEC>(ct.g_c + o);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        float gx, gy;
        const float g = sq_grad(a.v[e], rt.v[e], dn.v[e], gx, gy);
        const float ag = __fmul_rn(alpha, g);
        float G = 0.f;
        if (ag <= lam) G += er.v[e] + ct.sr[e];
        if (ag < lam) G += es.v[e] + ct.ss[e];
        if (ag == lam) G += ec.v[e] + ct.sc[e];
        G *= 0.5f * alpha;
        wx[e] = G * gx;
        wy[e] = G * gy;
    }
}

template <int VEC>
__global__ void __launch_bounds__(PR_TPB) mumford_shah_bwd_kernel(const float* __restrict__ x, float alpha, float lam,
                                                                  const float* __restrict__ g_r,
                                                                  const float* __restrict__ g_smooth,
                                                                  const float* __restrict__ g_contour,
                                                                  const float* __restrict__ g_sums,
                                                                  float* __restrict__ dx, int P, int H, int W, int K,
                                                                  int pix_per_cta) {
    const int b = blockIdx.y, t = threadIdx.x;
    const int KV = K / VEC, PP = PR_TPB / KV, TU = PP * KV;
    if (t >= TU) return;
    const int c = t % KV, ph = t / KV;
    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    const size_t base = (size_t)b * P * K;
    MsCot<VEC> ct;
    ct.g_r = g_r; ct.g_s = g_smooth; ct.g_c = g_contour;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const int k = c * VEC + e;
        ct.sr[e] = g_sums ? g_sums[((size_t)b * MS_NSUM + 0) * K + k] : 0.f;
        ct.ss[e] = g_sums ? g_sums[((size_t)b * MS_NSUM + 1) * K + k] : 0.f;
        ct.sc[e] = g_sums ? g_sums[((size_t)b * MS_NSUM + 2) * K + k] : 0.f;
        ct.sx[e] = g_sums ? g_sums[((size_t)b * MS_NSUM + 3) * K + k] : 0.f;
    }
    const size_t row = (size_t)W * K;
    for (int p = p_begin + ph; p < p_end; p += PP) {
        const int i = p / W, j = p - i * W;
        const size_t o = base + (size_t)p * K + c * VEC;
        const bool has_r = j + 1 < W, has_d = i + 1 < H, has_l = j > 0, has_u = i > 0;
        const Vec<VEC> a = vload<VEC>(x + o);
        const Vec<VEC> rt = has_r ? vload<VEC>(x + o + K) : vzero<VEC>();
        const Vec<VEC> dn = has_d ? vload<VEC>(x + o + row) : vzero<VEC>();
        float wx[VEC], wy[VEC], lx[VEC], ly[VEC], ux[VEC], uy[VEC];
        ms_weight<VEC>(ct, o, a, rt, dn, alpha, lam, wx, wy);
        Vec<VEC> out;
#pragma unroll
        for (int e = 0; e < VEC; ++e) out.v[e] = wx[e] + wy[e] + ct.sx[e];
        if (has_l) {   // the left pixel's gx reads this pixel
            const Vec<VEC> la = vload<VEC>(x + o - K);
            const Vec<VEC> ld = has_d ? vload<VEC>(x + o - K + row) : vzero<VEC>();
            ms_weight<VEC>(ct, o - K, la, a, ld, alpha, lam, lx, ly);
#pragma unroll
            for (int e = 0; e < VEC; ++e) out.v[e] -= lx[e];
        }
        if (has_u) {   // the upper pixel's gy reads this pixel
            const Vec<VEC> ua = vload<VEC>(x + o - row);
            const Vec<VEC> ur = has_r ? vload<VEC>(x + o - row + K) : vzero<VEC>();
            ms_weight<VEC>(ct, o - row, ua, ur, a, alpha, lam, ux, uy);
#pragma unroll
            for (int e = 0; e < VEC; ++e) out.v[e] -= uy[e];
        }
        vstore<VEC>(dx + o, out);
    }
}

// ------------------------------------------------------------------ priors on the logits: kl, GMRF, TV
// partial[block][3] = sum x^2, sum dy^2+dx^2, sum |dy|+|dx| (forward differences, zero in the last row / column)
__device__ __forceinline__ void block_sum3(float (&a)[3], float* __restrict__ partial) {
    __shared__ float red[PR_TPB / 32][3];
#pragma unroll
    for (int m = 0; m < 3; ++m) a[m] = group_sum<32>(a[m]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int m = 0; m < 3; ++m) red[threadIdx.x >> 5][m] = a[m];
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.f;
        for (int w = 0; w < PR_TPB / 32; ++w) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 3 + threadIdx.x] = s;
    }
}

template <int VEC>
__global__ void __launch_bounds__(PR_TPB) logit_priors_fwd_kernel(const float* __restrict__ x,
                                                                  float* __restrict__ partial, long long n_vec, int H,
                                                                  int W, int K) {
    // flat index over B*H*W*K/VEC vector elements, grid-stride
    const int KV = K / VEC;
    float acc[3] = {0.f, 0.f, 0.f};
    const long long stride = (long long)gridDim.x * PR_TPB;
    for (long long v = (long long)blockIdx.x * PR_TPB + threadIdx.x; v < n_vec; v += stride) {
        const long long pix = v / KV;
        const int j = (int)(pix % W), i = (int)((pix / W) % H);
        const size_t o = (size_t)v * VEC;
        const Vec<VEC> a = vload<VEC>(x + o);
        const Vec<VEC> rt = j + 1 < W ? vload<VEC>(x + o + K) : a;
        const Vec<VEC> dn = i + 1 < H ? vload<VEC>(x + o + (size_t)W * K) : a;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float dy = dn.v[e] - a.v[e], dx = rt.v[e] - a.v[e];
            acc[0] = fmaf(a.v[e], a.v[e], acc[0]);
            acc[1] = fmaf(dy, dy, fmaf(dx, dx, acc[1]));
            acc[2] += fabsf(dy) + fabsf(dx);
        }
    }
    block_sum3(acc, partial);
}

__global__ void logit_priors_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblocks,
                                             float inv_batch) {
    // one warp; lane l sums blocks l, l+32, ... in double, then a fixed shuffle tree
    double s[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += 32)
#pragma unroll
        for (int m = 0; m < 3; ++m) s[m] += (double)partial[(size_t)i * 3 + m];
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) s[m] += __shfl_xor_sync(0xffffffffu, s[m], o);
    if (threadIdx.x == 0) {
        out[0] = (float)(0.5 * s[0] * inv_batch);   // MeanFieldDistribution.kl
        out[1] = (float)(0.5 * s[1] * inv_batch);   // kl_improper_gmrf
        out[2] = (float)(s[2] * inv_batch);         // kl_tv
    }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

template <int VEC>
__global__ void __launch_bounds__(PR_TPB) logit_priors_bwd_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ g_out,
                                                                  float* __restrict__ dx_out, long long n_vec, int H,
                                                                  int W, int K, float inv_batch) {
    const int KV = K / VEC;
    const long long v = (long long)blockIdx.x * PR_TPB + threadIdx.x;
    if (v >= n_vec) return;
    const float g0 = g_out[0] * inv_batch, g1 = g_out[1] * inv_batch, g2 = g_out[2] * inv_batch;
    const long long pix = v / KV;
    const int j = (int)(pix % W), i = (int)((pix / W) % H);
    const size_t o = (size_t)v * VEC, row = (size_t)W * K;
    const Vec<VEC> a = vload<VEC>(x + o);
    const Vec<VEC> rt = j + 1 < W ? vload<VEC>(x + o + K) : a;
    const Vec<VEC> dn = i + 1 < H ? vload<VEC>(x + o + row) : a;
    const Vec<VEC> lf = j > 0 ? vload<VEC>(x + o - K) : a;
    const Vec<VEC> up = i > 0 ? vload<VEC>(x + o - row) : a;
    Vec<VEC> out;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        const float dy = dn.v[e] - a.v[e], dxx = rt.v[e] - a.v[e], dyu = a.v[e] - up.v[e], dxl = a.v[e] - lf.v[e];
        out.v[e] = g0 * a.v[e] + g1 * ((dyu - dy) + (dxl - dxx)) + g2 * ((sgn(dyu) - sgn(dy)) + (sgn(dxl) - sgn(dxx)));
    }
    vstore<VEC>(dx_out + o, out);
}

// ------------------------------------------------------------------ mean-field sample (nn.py:1421-1427)
__global__ void __launch_bounds__(PR_TPB) mean_field_sample_kernel(const float* __restrict__ mean,
                                                                   const float* __restrict__ eps, float noise,
                                                                   float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * PR_TPB + threadIdx.x;
    const long long n4 = n >> 2;
    if (i < n4) {
        const float4 m = ld4_stream(mean + 4 * i), e = ld4_stream(eps + 4 * i);
        st4_stream(out + 4 * i, make_float4(__fadd_rn(m.x, __fmul_rn(noise, e.x)), __fadd_rn(m.y, __fmul_rn(noise, e.y)),
                                            __fadd_rn(m.z, __fmul_rn(noise, e.z)), __fadd_rn(m.w, __fmul_rn(noise, e.w))));
    } else if (i == n4) {
        for (long long k = n4 << 2; k < n; ++k) out[k] = __fadd_rn(mean[k], __fmul_rn(noise, eps[k]));
    }
}

// ------------------------------------------------------------------ weak cross entropy (model.py:667-681)
// per pixel: p = softmax(x) (canonical), lsm = (x - max) - log(sum), labels = ST(hard_max(p), p) [mode 0] or p [mode 1];
// loss = -sum_k labels_k*lsm_k.  Backward with TF's registered gradient of softmax_cross_entropy_with_logits_v2
// (into logits AND labels) chained through the softmax: dx_k = g*((p_k - lab_k) - p_k*(lsm_k - sum_j p_j*lsm_j)).
struct XentPix {
    float4 p, lab, lsm;
};
template <int LPP>
__device__ __forceinline__ XentPix xent_pixel(float4 v, int c, int mode) {
    const float m = group_max<LPP>(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    const float4 d = make_float4(__fsub_rn(v.x, m), __fsub_rn(v.y, m), __fsub_rn(v.z, m), __fsub_rn(v.w, m));
    const float4 e = make_float4(exp_canon(d.x), exp_canon(d.y), exp_canon(d.z), exp_canon(d.w));
    float s = __fadd_rn(__fadd_rn(e.x, e.y), __fadd_rn(e.z, e.w));
    s = group_sum_canon<LPP>(s);
    const float rs = __frcp_rn(s), lse = log_canon(s);
    XentPix o;
    o.p = make_float4(__fmul_rn(e.x, rs), __fmul_rn(e.y, rs), __fmul_rn(e.z, rs), __fmul_rn(e.w, rs));
    o.lsm = make_float4(d.x - lse, d.y - lse, d.z - lse, d.w - lse);
    if (mode == 0) {
        const float pmax = group_max<LPP>(fmaxf(fmaxf(o.p.x, o.p.y), fmaxf(o.p.z, o.p.w)));
        o.lab = hard_st4(o.p, pmax);
    } else {
        o.lab = o.p;
    }
    (void)c;
    return o;
}

template <int LPP>
__global__ void __launch_bounds__(PR_TPB) weak_xent_fwd_kernel(const float* __restrict__ logits,
                                                               float* __restrict__ partial, long long n4, int mode) {
    __shared__ float red[PR_TPB / 32];
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * PR_TPB;
    const long long n4r = (n4 + PR_TPB - 1) / PR_TPB * PR_TPB;  // whole warps stay alive for the shuffles
    for (long long i = (long long)blockIdx.x * PR_TPB + threadIdx.x; i < n4r; i += stride) {
        const bool live = i < n4;
        const float4 v = ld4_stream(logits + 4 * (live ? i : n4 - 1));
        const XentPix q = xent_pixel<LPP>(v, (int)(i & (LPP - 1)), mode);
        const float l = q.lab.x * q.lsm.x + q.lab.y * q.lsm.y + q.lab.z * q.lsm.z + q.lab.w * q.lsm.w;
        if (live) acc -= l;
    }
    acc = group_sum<32>(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < PR_TPB / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

template <int LPP>
__global__ void __launch_bounds__(PR_TPB) weak_xent_bwd_kernel(const float* __restrict__ logits,
                                                               const float* __restrict__ g_out,
                                                               float* __restrict__ dlogits, long long n4, int mode,
                                                               float inv_count) {
    const long long i = (long long)blockIdx.x * PR_TPB + threadIdx.x;
    const bool live = i < n4;
    const float g = g_out[0] * inv_count;
    const float4 v = ld4_stream(logits + 4 * (live ? i : n4 - 1));
    const XentPix q = xent_pixel<LPP>(v, (int)(i & (LPP - 1)), mode);
    float dot = q.p.x * q.lsm.x + q.p.y * q.lsm.y + q.p.z * q.lsm.z + q.p.w * q.lsm.w;
    dot = group_sum<LPP>(dot);
    if (live)
        st4_stream(dlogits + 4 * i, make_float4(g * ((q.p.x - q.lab.x) - q.p.x * (q.lsm.x - dot)),
                                                g * ((q.p.y - q.lab.y) - q.p.y * (q.lsm.y - dot)),
                                                g * ((q.p.z - q.lab.z) - q.p.z * (q.lsm.z - dot)),
                                                g * ((q.p.w - q.lab.w) - q.p.w * (q.lsm.w - dot))));
}

// generic K (e.g. 25): one thread per pixel; DO_BWD writes dlogits, else accumulates the loss
template <bool DO_BWD>
__global__ void __launch_bounds__(128) weak_xent_generic_kernel(const float* __restrict__ logits,
                                                                const float* __restrict__ g_out,
                                                                float* __restrict__ out, long long n_pix, int K, int mode,
                                                                float inv_count) {
    __shared__ float red[4];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    if (i < n_pix) {
        float x[KMAX], p[KMAX], e[KMAX];
        for (int k = 0; k < K; ++k) x[k] = logits[i * K + k];
        float mx = x[0];
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k]);
        const float pmax = softmax_row_canon(x, p, e, K);   // e[0] = canonical sum of the exponentials
        const float lse = log_canon(e[0]);
        float dot = 0.f;
        for (int k = 0; k < K; ++k) {
            const float lsm = __fsub_rn(x[k], mx) - lse;
            const float lab = mode == 0 ? st_value(p[k] == pmax ? 1.f : 0.f, p[k]) : p[k];
            loss -= lab * lsm;
            dot += p[k] * lsm;
        }
        if (DO_BWD) {
            const float g = g_out[0] * inv_count;
            for (int k = 0; k < K; ++k) {
                const float lsm = __fsub_rn(x[k], mx) - lse;
                const float lab = mode == 0 ? st_value(p[k] == pmax ? 1.f : 0.f, p[k]) : p[k];
                out[i * K + k] = g * ((p[k] - lab) - p[k] * (lsm - dot));
            }
        }
    }
    if (!DO_BWD) {
        loss = group_sum<32>(loss);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
        __syncthreads();
        if (threadIdx.x == 0) out[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

__global__ void scalar_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, long long nblocks,
                                       float scale) {
    double s = 0.0;
    for (long long i = threadIdx.x; i < nblocks; i += 32) s += (double)partial[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[0] = (float)(s * scale);
}

// ------------------------------------------------------------------ mask2rgb (nn.py:2067-2089)
// table [K,3] = (colors - 0.5)*2 prepared by the caller; hot: the first maximum's colour (one_hot(argmax) . table);
// soft: sum_k mask_k * table_k in ascending k.  One thread per pixel; the 12-byte pixels go through shared memory
// so that the stores of a CTA are contiguous.
constexpr int RGB_TPB = 128;
__global__ void __launch_bounds__(RGB_TPB) mask2rgb_kernel(const float* __restrict__ mask,
                                                           const float* __restrict__ table, float* __restrict__ out,
                                                           long long n_pix, int K, int make_hot) {
    extern __shared__ float sm[];  // [K*3] table + [RGB_TPB*3] staging
    float* tab = sm;
    float* stage = sm + K * 3;
    for (int e = threadIdx.x; e < K * 3; e += RGB_TPB) tab[e] = table[e];
    __syncthreads();
    const long long p0 = (long long)blockIdx.x * RGB_TPB, i = p0 + threadIdx.x;
    float r = 0.f, g = 0.f, bl = 0.f;
    if (i < n_pix) {
        const float* m = mask + i * K;
        if (make_hot) {
            float best = m[0]; int a = 0;
            for (int k = 1; k < K; ++k) { const float v = m[k]; if (v > best) { best = v; a = k; } }
            r = tab[a * 3]; g = tab[a * 3 + 1]; bl = tab[a * 3 + 2];
        } else {
            for (int k = 0; k < K; ++k) {
                const float v = m[k];
                r = fmaf(v, tab[k * 3], r); g = fmaf(v, tab[k * 3 + 1], g); bl = fmaf(v, tab[k * 3 + 2], bl);
            }
        }
    }
    stage[threadIdx.x * 3] = r; stage[threadIdx.x * 3 + 1] = g; stage[threadIdx.x * 3 + 2] = bl;
    __syncthreads();
    const long long n_out = n_pix * 3, o0 = p0 * 3;
    for (int e = threadIdx.x; e < RGB_TPB * 3; e += RGB_TPB)
        if (o0 + e < n_out) __stcs(out + o0 + e, stage[e]);
}

}  // namespace ups

using namespace ups;

static int prior_checks(const char* what, int B, int H, int W, int K) {
    UPS_REQUIRE(B >= 0 && B <= 65535, "%s: B=%d out of range", what, B);
    UPS_REQUIRE(H >= 1 && W >= 1 && (long long)H * W < (1ll << 31), "%s: bad H=%d W=%d", what, H, W);
    UPS_REQUIRE(K >= 1 && K <= PR_TPB, "%s: K=%d not in [1, %d]", what, K, PR_TPB);
    return UPS_OK;
}

static bool vec4_ok(int K, const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr,
                    const void* e = nullptr) {
    return K % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d) && aligned16(e);
}

extern "C" int ups_mumford_shah_fwd(const float* x, float alpha, float lambda, float* r, float* smooth, float* contour,
                                    float* edges, float* sums, int B, int H, int W, int K, void* ws, size_t ws_bytes,
                                    void* stream) {
    UPS_REQUIRE(x, "mumford_shah_fwd: null input");
    UPS_REQUIRE(r || smooth || contour || edges || sums, "mumford_shah_fwd: no output requested");
    UPS_REQUIRE(alpha != 0.f, "mumford_shah_fwd: alpha must be non-zero");
    if (int rc = prior_checks("mumford_shah_fwd", B, H, W, K)) return rc;
    if (B == 0) return UPS_OK;
    const int P = H * W;
    const int splits = priors_splits(B, P);
    const int per = (int)cdiv(P, splits);
    float* partial = nullptr;
    if (sums) {
        const size_t need = mumford_shah_ws_bytes(B, P, K);
        if (!ws || ws_bytes < need) { set_error("mumford_shah_fwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
        partial = static_cast<float*>(ws);
    }
    const float thr = (float)((double)lambda / (double)alpha);   // edge_set: Python-float division, one fp32 rounding
    const dim3 grid(splits, B);
    cudaStream_t s = as_stream(stream);
    if (vec4_ok(K, x, r, smooth, contour, edges)) {
        const size_t sm = sums ? (size_t)(PR_TPB / (K / 4)) * K * MS_NSUM * sizeof(float) : 0;
        mumford_shah_fwd_kernel<4><<<grid, PR_TPB, sm, s>>>(x, alpha, lambda, thr, r, smooth, contour, edges, partial, P, H, W, K, per);
    } else {
        const size_t sm = sums ? (size_t)(PR_TPB / K) * K * MS_NSUM * sizeof(float) : 0;
        mumford_shah_fwd_kernel<1><<<grid, PR_TPB, sm, s>>>(x, alpha, lambda, thr, r, smooth, contour, edges, partial, P, H, W, K, per);
    }
    if (int rc = after_launch("mumford_shah_fwd_kernel")) return rc;
    if (sums) {
        const long long n = (long long)B * K * MS_NSUM;
        mumford_shah_finalize_kernel<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(partial, sums, splits, K, n);
        return after_launch("mumford_shah_finalize_kernel");
    }
    return UPS_OK;
}

extern "C" int ups_mumford_shah_bwd(const float* x, float alpha, float lambda, const float* g_r, const float* g_smooth,
                                    const float* g_contour, const float* g_sums, float* dx, int B, int H, int W, int K,
                                    void* stream) {
    UPS_REQUIRE(x && dx, "mumford_shah_bwd: null pointer");
    if (int rc = prior_checks("mumford_shah_bwd", B, H, W, K)) return rc;
    if (B == 0) return UPS_OK;
    const int P = H * W;
    const int splits = priors_splits(B, P);
    const int per = (int)cdiv(P, splits);
    const dim3 grid(splits, B);
    cudaStream_t s = as_stream(stream);
    if (vec4_ok(K, x, dx, g_r, g_smooth, g_contour))
        mumford_shah_bwd_kernel<4><<<grid, PR_TPB, 0, s>>>(x, alpha, lambda, g_r, g_smooth, g_contour, g_sums, dx, P, H, W, K, per);
    else
        mumford_shah_bwd_kernel<1><<<grid, PR_TPB, 0, s>>>(x, alpha, lambda, g_r, g_smooth, g_contour, g_sums, dx, P, H, W, K, per);
    return after_launch("mumford_shah_bwd_kernel");
}

extern "C" int ups_logit_priors_fwd(const float* mean, float* out, int B, int H, int W, int K, void* ws, size_t ws_bytes,
                                    void* stream) {
    UPS_REQUIRE(mean && out, "logit_priors_fwd: null pointer");
    if (int rc = prior_checks("logit_priors_fwd", B, H, W, K)) return rc;
    UPS_REQUIRE(B > 0, "logit_priors_fwd: empty batch has no mean");
    const size_t need = priors_scalar_ws_bytes();
    if (!ws || ws_bytes < need) { set_error("logit_priors_fwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    const bool v4 = vec4_ok(K, mean);
    const long long n_vec = (long long)B * H * W * K / (v4 ? 4 : 1);
    long long blocks = cdiv(n_vec, PR_TPB);
    if (blocks > PR_BLOCKS) blocks = PR_BLOCKS;
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    if (v4) logit_priors_fwd_kernel<4><<<(unsigned)blocks, PR_TPB, 0, s>>>(mean, partial, n_vec, H, W, K);
    else logit_priors_fwd_kernel<1><<<(unsigned)blocks, PR_TPB, 0, s>>>(mean, partial, n_vec, H, W, K);
    if (int rc = after_launch("logit_priors_fwd_kernel")) return rc;
    logit_priors_finalize_kernel<<<1, 32, 0, s>>>(partial, out, (int)blocks, 1.0f / (float)B);
    return after_launch("logit_priors_finalize_kernel");
}

extern "C" int ups_logit_priors_bwd(const float* mean, const float* g_out, float* dmean, int B, int H, int W, int K,
                                    void* stream) {
    UPS_REQUIRE(mean && g_out && dmean, "logit_priors_bwd: null pointer");
    if (int rc = prior_checks("logit_priors_bwd", B, H, W, K)) return rc;
    UPS_REQUIRE(B > 0, "logit_priors_bwd: empty batch");
    const bool v4 = vec4_ok(K, mean, dmean);
    const long long n_vec = (long long)B * H * W * K / (v4 ? 4 : 1);
    const long long blocks = cdiv(n_vec, PR_TPB);
    UPS_REQUIRE(blocks < (1ll << 31), "logit_priors_bwd: tensor too large");
    cudaStream_t s = as_stream(stream);
    if (v4) logit_priors_bwd_kernel<4><<<(unsigned)blocks, PR_TPB, 0, s>>>(mean, g_out, dmean, n_vec, H, W, K, 1.0f / (float)B);
    else logit_priors_bwd_kernel<1><<<(unsigned)blocks, PR_TPB, 0, s>>>(mean, g_out, dmean, n_vec, H, W, K, 1.0f / (float)B);
    return after_launch("logit_priors_bwd_kernel");
}

extern "C" int ups_mean_field_sample_fwd(const float* mean, const float* eps, float noise_level, float* out, long long n,
                                         void* stream) {
    UPS_REQUIRE(mean && eps && out, "mean_field_sample_fwd: null pointer");
    UPS_REQUIRE(n >= 0, "mean_field_sample_fwd: n=%lld", n);
    UPS_REQUIRE(aligned16(mean) && aligned16(eps) && aligned16(out), "mean_field_sample_fwd: 16-byte alignment");
    if (n == 0) return UPS_OK;
    const long long blocks = cdiv((n >> 2) + 1, PR_TPB);
    UPS_REQUIRE(blocks < (1ll << 31), "mean_field_sample_fwd: tensor too large");
    mean_field_sample_kernel<<<(unsigned)blocks, PR_TPB, 0, as_stream(stream)>>>(mean, eps, noise_level, out, n);
    return after_launch("mean_field_sample_kernel");
}

static int xent_checks(const char* what, long long n_pix, int K, int mode) {
    UPS_REQUIRE(n_pix > 0 && K >= 1 && K <= KMAX, "%s: n_pix=%lld K=%d (K <= %d)", what, n_pix, K, KMAX);
    UPS_REQUIRE(mode == 0 || mode == 1, "%s: entropy_func mode %d (0 = cross_entropy, 1 = entropy)", what, mode);
    return UPS_OK;
}

extern "C" int ups_weak_xent_fwd(const float* logits, int mode, float* out, long long n_pix, int K, void* ws,
                                 size_t ws_bytes, void* stream) {
    UPS_REQUIRE(logits && out, "weak_xent_fwd: null pointer");
    if (int rc = xent_checks("weak_xent_fwd", n_pix, K, mode)) return rc;
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    const bool fast = (K == 4 || K == 8 || K == 16 || K == 32) && aligned16(logits);
    long long blocks;
    if (fast) {
        const long long n4 = n_pix * (K / 4);
        blocks = cdiv(n4, PR_TPB);
        if (blocks > PR_BLOCKS) blocks = PR_BLOCKS;
    } else {
        blocks = cdiv(n_pix, 128);
        UPS_REQUIRE(blocks < (1ll << 31), "weak_xent_fwd: tensor too large");
    }
    const size_t need = (size_t)blocks * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("weak_xent_fwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    if (fast) {
        const long long n4 = n_pix * (K / 4);
        if (K == 4) weak_xent_fwd_kernel<1><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, partial, n4, mode);
        else if (K == 8) weak_xent_fwd_kernel<2><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, partial, n4, mode);
        else if (K == 16) weak_xent_fwd_kernel<4><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, partial, n4, mode);
        else weak_xent_fwd_kernel<8><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, partial, n4, mode);
    } else {
        weak_xent_generic_kernel<false><<<(unsigned)blocks, 128, 0, s>>>(logits, nullptr, partial, n_pix, K, mode, 0.f);
    }
    if (int rc = after_launch("weak_xent_fwd_kernel")) return rc;
    scalar_finalize_kernel<<<1, 32, 0, s>>>(partial, out, blocks, 1.0f / (float)n_pix);
    return after_launch("scalar_finalize_kernel");
}

extern "C" int ups_weak_xent_bwd(const float* logits, int mode, const float* g_out, float* dlogits, long long n_pix, int K,
                                 void* stream) {
    UPS_REQUIRE(logits && g_out && dlogits, "weak_xent_bwd: null pointer");
    if (int rc = xent_checks("weak_xent_bwd", n_pix, K, mode)) return rc;
    cudaStream_t s = as_stream(stream);
    const float inv = 1.0f / (float)n_pix;
    if ((K == 4 || K == 8 || K == 16 || K == 32) && aligned16(logits) && aligned16(dlogits)) {
        const long long n4 = n_pix * (K / 4);
        const long long blocks = cdiv(n4, PR_TPB);
        UPS_REQUIRE(blocks < (1ll << 31), "weak_xent_bwd: tensor too large");
        if (K == 4) weak_xent_bwd_kernel<1><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, g_out, dlogits, n4, mode, inv);
        else if (K == 8) weak_xent_bwd_kernel<2><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, g_out, dlogits, n4, mode, inv);
        else if (K == 16) weak_xent_bwd_kernel<4><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, g_out, dlogits, n4, mode, inv);
        else weak_xent_bwd_kernel<8><<<(unsigned)blocks, PR_TPB, 0, s>>>(logits, g_out, dlogits, n4, mode, inv);
    } else {
        const long long blocks = cdiv(n_pix, 128);
        UPS_REQUIRE(blocks < (1ll << 31), "weak_xent_bwd: tensor too large");
        weak_xent_generic_kernel<true><<<(unsigned)blocks, 128, 0, s>>>(logits, g_out, dlogits, n_pix, K, mode, inv);
    }
    return after_launch("weak_xent_bwd_kernel");
}

extern "C" int ups_mask2rgb_fwd(const float* mask, const float* table, int make_hot, float* out, long long n_pix, int K,
                                void* stream) {
    UPS_REQUIRE(mask && table && out, "mask2rgb_fwd: null pointer");
    UPS_REQUIRE(n_pix >= 0 && K >= 1 && K <= 1024, "mask2rgb_fwd: n_pix=%lld K=%d", n_pix, K);
    if (n_pix == 0) return UPS_OK;
    const long long blocks = cdiv(n_pix, RGB_TPB);
    UPS_REQUIRE(blocks < (1ll << 31), "mask2rgb_fwd: tensor too large");
    const size_t sm = (size_t)(K * 3 + RGB_TPB * 3) * sizeof(float);
    mask2rgb_kernel<<<(unsigned)blocks, RGB_TPB, sm, as_stream(stream)>>>(mask, table, out, n_pix, K, make_hot);
    return after_launch("mask2rgb_kernel");
}
