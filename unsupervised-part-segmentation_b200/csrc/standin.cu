// Stand-in parameterised modules of the data-parallel step wrapper (SURVEY.md 8d "DP step", 8e).
//
// The path itself has no parameters: what a data-parallel training step all-reduces are the gradients of the
// CNNs around it.  Two tiny modules with the reference's shapes stand in for them and PRODUCE real, rank-dependent
// gradients at the two places where the real model's gradients appear relative to the path:
//
//   encoder tail  feat[b,k,:] = pooled[b,k,:] . Wlin + blin      cub/code/SB_model48i/model.py:50-52 (mean over HW, then the
//                                                                 1x1 conv to F of e_alpha; pooled comes from K2)
//       backward (after K4 produced dfeat):  dWlin[c,f] = sum_{b,k} pooled[b,k,c] dfeat[b,k,f] ; dblin[f] = sum dfeat
//   decoder head  recon[b,p,:] = concat(feat[b,label[b,p],:], one_hot(label[b,p])) . Whead + bhead
//                 a 1x1 conv F+K -> 3 on nn.unpool_features_gathered's injection (cub/code/nn.py:2469-2487; the first
//                 layer of `dd`, model.py:96,485, is such a conv on the injected map)
//       backward (before K4, from the reconstruction cotangent):  with R[b,k,c] = sum_{p: label=k} g_recon[b,p,c]
//                 dWhead[f,c] = sum_{b,k} feat[b,k,f] R[b,k,c] ; dWhead[F+k,c] = sum_b R[b,k,c] ; dbhead[c] = sum R
//
// All sums run in a fixed order (per-CTA partials, reduced by index): bit-reproducible, no atomics on data.
#include "common.cuh"

namespace ups {
namespace standin {

constexpr int TPB = 256;

// ---------------------------------------------------------------- encoder tail
__global__ void tail_fwd_kernel(const float* __restrict__ pooled, const float* __restrict__ Wlin,
                                const float* __restrict__ blin, float* __restrict__ feat, long long rows, int C, int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * F) return;
    const long long r = i / F;
    const int f = (int)(i - r * F);
    float s = blin[f];
    for (int c = 0; c < C; ++c) s = fmaf(pooled[r * C + c], Wlin[c * F + f], s);
    feat[i] = s;
}

// partial[cta][(c|C=bias), f] over the CTA's rows; finished by tail_bwd_finish_kernel
__global__ void __launch_bounds__(TPB) tail_bwd_partial_kernel(const float* __restrict__ pooled,
                                                                const float* __restrict__ dfeat,
                                                                float* __restrict__ partial, long long rows, int C, int F,
                                                                int rows_per_cta) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    const int n = (C + 1) * F;
    for (int t = threadIdx.x; t < n; t += TPB) {
        const int c = t / F, f = t - c * F;
        float s = 0.f;
        for (long long r = r0; r < r1; ++r) s = fmaf(c < C ? pooled[r * C + c] : 1.0f, dfeat[r * F + f], s);
        partial[(size_t)blockIdx.x * n + t] = s;
    }
}

// out[t] = sum_p partial[p][t]: one warp per output, lane l takes p = l, l+32, ... in ascending order, then a fixed
// shuffle tree: deterministic, and the `parts` loads of an output are 32-way parallel instead of one dependent chain
__global__ void __launch_bounds__(256) finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int n,
                                                     int parts) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= n) return;   // warp-uniform
    float s = 0.f;
    for (int p = lane; p < parts; p += 32) s += partial[(size_t)p * n + t];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[t] = s;
}

// ---------------------------------------------------------------- decoder head
__global__ void head_fwd_kernel(const long long* __restrict__ labels, const float* __restrict__ feat,
                                const float* __restrict__ Whead, const float* __restrict__ bhead,
                                float* __restrict__ recon, int P, int K, int F) {
    extern __shared__ float tab[];   // T[k][c] = bhead[c] + Whead[F+k,c] + sum_f feat[b,k,f] Whead[f,c]
    const int b = blockIdx.y;
    for (int t = threadIdx.x; t < K * 3; t += blockDim.x) {
        const int k = t / 3, c = t - 3 * k;
        float s = bhead[c] + Whead[(F + k) * 3 + c];
        for (int f = 0; f < F; ++f) s = fmaf(feat[((size_t)b * K + k) * F + f], Whead[f * 3 + c], s);
        tab[t] = s;
    }
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        int k = (int)labels[(size_t)b * P + p];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        float* o = recon + ((size_t)b * P + p) * 3;
        o[0] = tab[3 * k]; o[1] = tab[3 * k + 1]; o[2] = tab[3 * k + 2];
    }
}

// R partial [b][split][k][c] = sum over the CTA's pixels with label k of g_recon[b,p,c]; lane-private accumulators
__global__ void __launch_bounds__(TPB) head_pool_kernel(const float* __restrict__ g_recon,
                                                         const long long* __restrict__ labels,
                                                         float* __restrict__ partial, int P, int K, int pix_per_cta) {
    extern __shared__ float acc[];   // [warp][k*3+c][32 lanes]
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = TPB / 32;
    float* a = acc + (size_t)warp * K * 3 * 32;
    for (int i = lane; i < K * 3 * 32; i += 32) a[i] = 0.f;
    __syncwarp();
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(P, p0 + pix_per_cta);
    for (int p = p0 + threadIdx.x; p < p1; p += TPB) {
        int k = (int)__ldcs(labels + (size_t)b * P + p);
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        const float* g = g_recon + ((size_t)b * P + p) * 3;
        float* q = a + (k * 3) * 32 + lane;
        q[0] += __ldcs(g); q[32] += __ldcs(g + 1); q[64] += __ldcs(g + 2);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < K * 3; t += TPB) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w)
            for (int l = 0; l < 32; ++l) s += acc[((size_t)w * K * 3 + t) * 32 + ((l + t) & 31)];
        partial[((size_t)b * gridDim.x + blockIdx.x) * (K * 3) + t] = s;
    }
}

// one CTA per sample: that sample's contribution to dW / db; [b][(F+K)*3 + 3] (summed over b by finish_kernel)
__global__ void __launch_bounds__(TPB) head_grad_partial_kernel(const float* __restrict__ Rpart,
                                                                 const float* __restrict__ feat,
                                                                 float* __restrict__ partial, int K, int F, int splits) {
    extern __shared__ float R[];   // [K][3] of this sample
    const int n = (F + K) * 3 + 3;
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < K * 3; t += TPB) {
        float r = 0.f;
        for (int sp = 0; sp < splits; ++sp) r += Rpart[((size_t)b * splits + sp) * (K * 3) + t];
        R[t] = r;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += TPB) {
        float v = 0.f;
        if (t < F * 3) {
            const int f = t / 3, c = t - 3 * f;
            for (int k = 0; k < K; ++k) v = fmaf(feat[((size_t)b * K + k) * F + f], R[k * 3 + c], v);
        } else if (t < (F + K) * 3) {
            v = R[t - F * 3];
        } else {
            const int c = t - (F + K) * 3;
            for (int k = 0; k < K; ++k) v += R[k * 3 + c];
        }
        partial[(size_t)b * n + t] = v;
    }
}

constexpr int TAIL_CTAS = 256;
inline int head_splits(int P) { return (int)(cdiv(P, 4096) < 1 ? 1 : cdiv(P, 4096)); }

}  // namespace standin
}  // namespace ups

using namespace ups;

static void standin_carveouts() {
    static const bool once = []() {
        prefer_max_shared_carveout(standin::head_pool_kernel);
        prefer_max_shared_carveout(standin::head_grad_partial_kernel);
        prefer_max_shared_carveout(standin::tail_bwd_partial_kernel);
        prefer_max_shared_carveout(standin::finish_kernel);
        return true;
    }();
    (void)once;
}

extern "C" size_t ups_standin_workspace_bytes(int B, int P, int K, int F) {
    if (B <= 0 || P <= 0 || K <= 0 || F <= 0) return 0;
    const size_t tail = (size_t)standin::TAIL_CTAS * 4 * F * sizeof(float);
    const size_t head = ((size_t)B * standin::head_splits(P) * K * 3 +
                         (size_t)B * ((F + K) * 3 + 3)) * sizeof(float);
    return (tail > head ? tail : head) + 256;
}

extern "C" int ups_standin_tail_fwd(const float* pooled, const float* Wlin, const float* blin, float* feat, int B, int K,
                                    int C, int F, void* stream) {
    UPS_REQUIRE(pooled && Wlin && blin && feat, "standin_tail_fwd: null pointer");
    UPS_REQUIRE(B >= 0 && K > 0 && C > 0 && F > 0, "standin_tail_fwd: bad shape");
    const long long rows = (long long)B * K;
    if (rows == 0) return UPS_OK;
    standin::tail_fwd_kernel<<<(unsigned)cdiv(rows * F, 256), 256, 0, as_stream(stream)>>>(pooled, Wlin, blin, feat, rows, C, F);
    return after_launch("standin::tail_fwd_kernel");
}

extern "C" int ups_standin_tail_bwd(const float* pooled, const float* dfeat, float* dWlin_dblin, int B, int K, int C,
                                    int F, void* ws, size_t ws_bytes, void* stream) {
    UPS_REQUIRE(pooled && dfeat && dWlin_dblin, "standin_tail_bwd: null pointer");
    UPS_REQUIRE(B > 0 && K > 0 && C == 3 && F > 0, "standin_tail_bwd: bad shape (C must be 3)");
    const size_t need = (size_t)standin::TAIL_CTAS * (C + 1) * F * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("standin_tail_bwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    standin_carveouts();
    const long long rows = (long long)B * K;
    const int rpc = (int)cdiv(rows, standin::TAIL_CTAS);
    const int ctas = (int)cdiv(rows, rpc);
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    standin::tail_bwd_partial_kernel<<<ctas, standin::TPB, 0, s>>>(pooled, dfeat, partial, rows, C, F, rpc);
    if (int rc = after_launch("standin::tail_bwd_partial_kernel")) return rc;
    const int n = (C + 1) * F;
    standin::finish_kernel<<<(unsigned)cdiv(n, 8), 256, 0, s>>>(partial, dWlin_dblin, n, ctas);
    return after_launch("standin::finish_kernel");
}

extern "C" int ups_standin_head_fwd(const long long* labels, const float* feat, const float* Whead, const float* bhead,
                                    float* recon, int B, int P, int K, int F, void* stream) {
    UPS_REQUIRE(labels && feat && Whead && bhead && recon, "standin_head_fwd: null pointer");
    UPS_REQUIRE(B >= 0 && B <= 65535 && P > 0 && K > 0 && F > 0, "standin_head_fwd: bad shape");
    if (B == 0) return UPS_OK;
    dim3 grid((unsigned)(cdiv(P, 256 * 8) < 1 ? 1 : cdiv(P, 256 * 8)), B);
    standin::head_fwd_kernel<<<grid, 256, K * 3 * sizeof(float), as_stream(stream)>>>(labels, feat, Whead, bhead, recon, P, K, F);
    return after_launch("standin::head_fwd_kernel");
}

extern "C" int ups_standin_head_bwd(const float* g_recon, const long long* labels, const float* feat,
                                    float* dWhead_dbhead, int B, int P, int K, int F, void* ws, size_t ws_bytes,
                                    void* stream) {
    UPS_REQUIRE(g_recon && labels && feat && dWhead_dbhead, "standin_head_bwd: null pointer");
    UPS_REQUIRE(B > 0 && B <= 65535 && P > 0 && K > 0 && K <= 64 && F > 0, "standin_head_bwd: bad shape");
    const int n = (F + K) * 3 + 3;
    const int splits = standin::head_splits(P);
    const int rparts = splits;                             // partial sums of R per sample
    const int chunks = B;                                  // one CTA (and one partial dW) per sample
    const size_t need = ((size_t)B * rparts * K * 3 + (size_t)chunks * n) * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("standin_head_bwd: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    standin_carveouts();
    cudaStream_t s = as_stream(stream);
    float* Rpart = static_cast<float*>(ws);
    float* partial = Rpart + (size_t)B * rparts * K * 3;
    const int ppc = (int)cdiv(P, splits);
    const size_t sm = (size_t)(standin::TPB / 32) * K * 3 * 32 * sizeof(float);
    if (sm > 48 * 1024)
        UPS_CUDA(cudaFuncSetAttribute(standin::head_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    standin::head_pool_kernel<<<dim3(splits, B), standin::TPB, sm, s>>>(g_recon, labels, Rpart, P, K, ppc);
    if (int rc = after_launch("standin::head_pool_kernel")) return rc;
    standin::head_grad_partial_kernel<<<chunks, standin::TPB, K * 3 * sizeof(float), s>>>(Rpart, feat, partial, K, F, rparts);
    if (int rc = after_launch("standin::head_grad_partial_kernel")) return rc;
    standin::finish_kernel<<<(unsigned)cdiv(n, 8), 256, 0, s>>>(partial, dWhead_dbhead, n, chunks);
    return after_launch("standin::finish_kernel");
}
