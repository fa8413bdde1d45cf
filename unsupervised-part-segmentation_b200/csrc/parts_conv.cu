// SURVEY.md 8f N4 (encoder side): the appearance encoder's first convolution applied to the masked part
// images without materialising them.
//
//   view1_parts = mask_parts(view1', encoding_mask)                 cub/code/SB_model48i/model.py:176-187, :478
//   nn.apply_partwise(view1_parts, e_alpha)  -> [K*B,h,w,3]         cub/code/nn.py:81-113 (model.py:214-222)
//   e_alpha: h = nn.conv2d(x, config[0]) (3x3, stride 1, SAME, +b)  model.py:40, cub/code/nn.py:617-664
//
//   out[k*B+b, y,x, o] = b[o] + sum_t sum_c (img[b,q_t,c] * mask[b,q_t,k]) V[t,c,o],   q_t = (y+i-1, x+j-1), t = 3i+j
//
// The encoding mask is straight-through hard: a pixel belongs to one part, so of the K output planes only the
// (at most 9) planes named by the labels of the 3x3 neighbourhood differ from the bias.  The kernel reads the
// image and the mask once (76 B/pixel) and streams the K*Co outputs (the write is the roofline: K*Co*4 B/pixel);
// the [K*B,h,w,3] part images (805 MB at CUB B=256, written by K2 and read back by the conv) are never formed.
// Pixels with several non-zero mask entries (exact ties, soft masks) take a dense loop over k: exact for any mask.
#include "common.cuh"

namespace ups {
namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_ROWS = 8;

__device__ __forceinline__ float4 dot3(float4 e, float4 v0, float4 v1, float4 v2) {
    return make_float4(fmaf(e.z, v2.x, fmaf(e.y, v1.x, __fmul_rn(e.x, v0.x))), fmaf(e.z, v2.y, fmaf(e.y, v1.y, __fmul_rn(e.x, v0.y))),
                       fmaf(e.z, v2.z, fmaf(e.y, v1.z, __fmul_rn(e.x, v0.z))), fmaf(e.z, v2.w, fmaf(e.y, v1.w, __fmul_rn(e.x, v0.w))));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// grid (splits, B); a CTA walks `strips_per_cta` strips of 8 rows x W of one sample.  The strip's pixels (1-pixel
// halo) are staged as (c0*m, c1*m, c2*m, label); then one thread per (pixel, 4 output channels): the 9 tap
// contributions w_t = part_pixel . V[t], the K-|labels| planes that only see the bias get the bias, and each
// distinct label among the 9 taps gets bias + the sum of its taps.  Everything is per-thread predication; a
// pixel whose neighbourhood holds a several-non-zero mask pixel takes the dense loop over k.
__global__ void __launch_bounds__(PC_THREADS) parts_conv_fwd_kernel(const float* __restrict__ img,
                                                                    const float* __restrict__ mask,
                                                                    const float* __restrict__ V,
                                                                    const float* __restrict__ bias,
                                                                    float* __restrict__ out, int B, int H, int W, int K,
                                                                    int Co, int strips_per_cta) {
    extern __shared__ float4 smem4[];
    float* sV = reinterpret_cast<float*>(smem4);  // [9][3][Co]
    float* sB = sV + 27 * Co;                     // [Co]
    float4* sPx = reinterpret_cast<float4*>(sB + Co);  // [(8+2)][W+2] : (c0, c1, c2, label), premultiplied if label >= 0
    const int b = blockIdx.y, tid = threadIdx.x;
    const int Co4 = Co >> 2;
    for (int i = tid; i < 27 * Co; i += PC_THREADS) sV[i] = __ldg(V + i);
    for (int i = tid; i < Co; i += PC_THREADS) sB[i] = __ldg(bias + i);
    const int n_strips = (H + PC_ROWS - 1) / PC_ROWS;
    const int s_beg = blockIdx.x * strips_per_cta;
    const int s_end = min(n_strips, s_beg + strips_per_cta);
    const int Wp = W + 2;
    const size_t P = (size_t)H * W;
    const size_t plane = (size_t)B * P * Co;  // stride between part planes k -> k+1
    const float* mb = mask + (size_t)b * P * K;
    const float* ib = img + (size_t)b * P * 3;
    float* ob = out + (size_t)b * P * Co;
    for (int s = s_beg; s < s_end; ++s) {
        const int y0 = s * PC_ROWS;
        const int rows = min(PC_ROWS, H - y0);
        __syncthreads();
        for (int i = tid; i < (rows + 2) * Wp; i += PC_THREADS) {
            const int r = i / Wp, c = i - r * Wp;
            const int y = y0 - 1 + r, x = c - 1;
            float4 e = make_float4(0.f, 0.f, 0.f, __int_as_float(-2));
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const size_t q = (size_t)y * W + x;
                const float* m = mb + q * K;
                int lab = -2, nz = 0;
                float val = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float v = __ldg(m + k);
                    if (v != 0.f) {
                        if (nz == 0) { lab = k; val = v; }
                        ++nz;
                    }
                }
                if (nz > 1) { lab = -1; val = 1.f; }
                if (nz > 0)  // mask_parts: fl(image * mask) per channel (several-non-zero pixels keep the raw image)
                    e = make_float4(__fmul_rn(__ldg(ib + q * 3 + 0), val), __fmul_rn(__ldg(ib + q * 3 + 1), val),
                                    __fmul_rn(__ldg(ib + q * 3 + 2), val), __int_as_float(lab));
            }
            sPx[i] = e;
        }
        __syncthreads();
        // idx = (r*W + x)*Co4 + o4 advances by the block size: carry-propagate instead of dividing
        const int d_px = PC_THREADS / Co4, d_o4 = PC_THREADS - d_px * Co4;
        int px0 = tid / Co4, o4 = tid - px0 * Co4;
        int r = px0 / W, x = px0 - r * W;
        for (int idx = tid; idx < rows * W * Co4; idx += PC_THREADS, o4 += d_o4, x += d_px) {
            if (o4 >= Co4) { o4 -= Co4; ++x; }
            while (x >= W) { x -= W; ++r; }
            const int o = 4 * o4;
            const float4 bo = *reinterpret_cast<const float4*>(sB + o);
            const float4* e0 = sPx + r * Wp + x;
            int labs[9];
            float4 w[9];
            unsigned present = 0;
            bool dense = false;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3, dx = t - 3 * dy;
                const float4 e = e0[dy * Wp + dx];
                labs[t] = __float_as_int(e.w);
                const float* v = sV + t * 3 * Co + o;
                w[t] = dot3(e, *reinterpret_cast<const float4*>(v), *reinterpret_cast<const float4*>(v + Co),
                            *reinterpret_cast<const float4*>(v + 2 * Co));
                if (labs[t] >= 0) present |= 1u << labs[t];
                dense |= labs[t] == -1;
            }
            float* op = ob + ((size_t)(y0 + r) * W + x) * Co + o;
            if (!dense) {
                for (int k = 0; k < K; ++k)
                    if (!((present >> k) & 1u)) st4_stream(op + k * plane, bo);
                unsigned seen = 0;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int lab = labs[t];
                    if (lab >= 0 && !((seen >> lab) & 1u)) {
                        seen |= 1u << lab;
                        float4 sacc = add4(bo, w[t]);
#pragma unroll
                        for (int u = t + 1; u < 9; ++u)
                            if (labs[u] == lab) sacc = add4(sacc, w[u]);
                        st4_stream(op + lab * plane, sacc);
                    }
                }
            } else {
                for (int k = 0; k < K; ++k) {
                    float4 sacc = bo;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (labs[t] == k) {
                            sacc = add4(sacc, w[t]);
                        } else if (labs[t] == -1) {
                            const int dy = t / 3, dx = t - 3 * dy;
                            const float mv = __ldg(mb + ((size_t)(y0 + r + dy - 1) * W + (x + dx - 1)) * K + k);
                            if (mv != 0.f) {
                                float4 e = e0[dy * Wp + dx];  // raw image
                                e.x = __fmul_rn(e.x, mv); e.y = __fmul_rn(e.y, mv); e.z = __fmul_rn(e.z, mv);
                                const float* v = sV + t * 3 * Co + o;
                                sacc = add4(sacc, dot3(e, *reinterpret_cast<const float4*>(v),
                                                       *reinterpret_cast<const float4*>(v + Co),
                                                       *reinterpret_cast<const float4*>(v + 2 * Co)));
                            }
                        }
                    }
                    st4_stream(op + k * plane, sacc);
                }
            }
        }
    }
}

}  // namespace
}  // namespace ups

using namespace ups;

extern "C" int ups_parts_conv_fwd(const float* img, const float* mask, const float* V, const float* bias, float* out_pm,
                                  int B, int H, int W, int K, int C, int Co, void* stream) {
    UPS_REQUIRE(img && mask && V && bias && out_pm, "parts_conv_fwd: null pointer");
    UPS_REQUIRE(B >= 0 && H > 0 && W > 0, "parts_conv_fwd: bad sizes B=%d H=%d W=%d", B, H, W);
    UPS_REQUIRE(K >= 1 && K <= 32, "parts_conv_fwd: K=%d outside [1,32]", K);
    UPS_REQUIRE(C == 3, "parts_conv_fwd: C=%d (the part images are 3-channel, model.py:316-327)", C);
    UPS_REQUIRE(Co >= 4 && Co <= 128 && Co % 4 == 0, "parts_conv_fwd: Co=%d must be a multiple of 4 in [4,128]", Co);
    UPS_REQUIRE((long long)B * H * W * K < (1ll << 31), "parts_conv_fwd: K*B*H*W >= 2^31");
    UPS_REQUIRE(aligned16(V) && aligned16(bias) && aligned16(out_pm), "parts_conv_fwd: V, bias, out must be 16-byte aligned");
    if (B == 0) return UPS_OK;
    const size_t smem = (size_t)28 * Co * sizeof(float) + (size_t)(PC_ROWS + 2) * (W + 2) * sizeof(float4);
    UPS_REQUIRE(smem <= 200 * 1024, "parts_conv_fwd: W=%d Co=%d needs %zu bytes of shared memory", W, Co, smem);
    const int n_strips = (int)cdiv(H, PC_ROWS);
    long long want = cdiv(8ll * NUM_SMS, B);
    if (want < 1) want = 1;
    if (want > n_strips) want = n_strips;
    const int spc = (int)cdiv(n_strips, want);
    const dim3 grid((unsigned)cdiv(n_strips, spc), B);
    cudaStream_t st = as_stream(stream);
    UPS_CUDA(cudaFuncSetAttribute(parts_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    parts_conv_fwd_kernel<<<grid, PC_THREADS, smem, st>>>(img, mask, V, bias, out_pm, B, H, W, K, Co, spc);
    return after_launch("parts_conv_fwd_kernel");
}
