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

constexpr int PC_THREADS = 256;  // 8 warps = 8 rows of a strip; lane = output channel
constexpr int PC_ROWS = 8;

// grid (splits, B); a CTA walks `strips_per_cta` strips of 8 rows x W of one sample.
template <int CCH>
__global__ void __launch_bounds__(PC_THREADS) parts_conv_fwd_kernel(const float* __restrict__ img,
                                                                    const float* __restrict__ mask,
                                                                    const float* __restrict__ V,
                                                                    const float* __restrict__ bias,
                                                                    float* __restrict__ out, int B, int H, int W, int K,
                                                                    int Co, int strips_per_cta) {
    extern __shared__ float4 sPx[];  // [(8+2)][W+2] : (c0, c1, c2, label) — premultiplied by the mask value if label >= 0
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float Vr[9][3][CCH], bo[CCH];
#pragma unroll
    for (int c = 0; c < CCH; ++c) {
        const int o = lane + 32 * c;
        bo[c] = o < Co ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) Vr[t][ch][c] = o < Co ? __ldg(V + (t * 3 + ch) * Co + o) : 0.f;
    }
    const int n_strips = (H + PC_ROWS - 1) / PC_ROWS;
    const int s_beg = blockIdx.x * strips_per_cta;
    const int s_end = min(n_strips, s_beg + strips_per_cta);
    const int Wp = W + 2;
    const size_t P = (size_t)H * W;
    const float* mb = mask + (size_t)b * P * K;
    const float* ib = img + (size_t)b * P * 3;
    for (int s = s_beg; s < s_end; ++s) {
        const int y0 = s * PC_ROWS;
        __syncthreads();
        for (int i = tid; i < (PC_ROWS + 2) * Wp; i += PC_THREADS) {
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
                if (nz > 0)  // mask_parts: fl(image * mask) per channel (dense pixels keep the raw image)
                    e = make_float4(__fmul_rn(__ldg(ib + q * 3 + 0), val), __fmul_rn(__ldg(ib + q * 3 + 1), val),
                                    __fmul_rn(__ldg(ib + q * 3 + 2), val), __int_as_float(lab));
            }
            sPx[i] = e;
        }
        __syncthreads();
        const int y = y0 + warp;
        if (y >= H) continue;
        for (int x = 0; x < W; ++x) {
            int labs[9];
            float w[9][CCH];
            unsigned present = 0;
            bool any_dense = false;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3, dx = t - 3 * dy;
                const float4 e = sPx[(warp + dy) * Wp + x + dx];
                labs[t] = __float_as_int(e.w);
#pragma unroll
                for (int c = 0; c < CCH; ++c)
                    w[t][c] = fmaf(e.z, Vr[t][2][c], fmaf(e.y, Vr[t][1][c], __fmul_rn(e.x, Vr[t][0][c])));
                if (labs[t] >= 0) present |= 1u << labs[t];
                any_dense |= labs[t] == -1;
            }
            const size_t p = (size_t)y * W + x;
            for (int k = 0; k < K; ++k) {
                float acc[CCH];
#pragma unroll
                for (int c = 0; c < CCH; ++c) acc[c] = bo[c];
                if ((present >> k) & 1u) {
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        if (labs[t] == k) {
#pragma unroll
                            for (int c = 0; c < CCH; ++c) acc[c] += w[t][c];
                        }
                }
                if (any_dense) {
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        if (labs[t] == -1) {
                            const int dy = t / 3, dx = t - 3 * dy;
                            const float mv = __ldg(mb + ((size_t)(y + dy - 1) * W + (x + dx - 1)) * K + k);
                            if (mv != 0.f) {
                                const float4 e = sPx[(warp + dy) * Wp + x + dx];  // raw image
                                const float p0 = __fmul_rn(e.x, mv), p1 = __fmul_rn(e.y, mv), p2 = __fmul_rn(e.z, mv);
#pragma unroll
                                for (int c = 0; c < CCH; ++c)
                                    acc[c] += fmaf(p2, Vr[t][2][c], fmaf(p1, Vr[t][1][c], __fmul_rn(p0, Vr[t][0][c])));
                            }
                        }
                }
                float* orow = out + (((size_t)k * B + b) * P + p) * Co;
#pragma unroll
                for (int c = 0; c < CCH; ++c)
                    if (lane + 32 * c < Co) __stcs(orow + lane + 32 * c, acc[c]);
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
    UPS_REQUIRE(Co >= 1 && Co <= 128, "parts_conv_fwd: Co=%d outside [1,128]", Co);
    UPS_REQUIRE((long long)B * H * W * K < (1ll << 31), "parts_conv_fwd: K*B*H*W >= 2^31");
    if (B == 0) return UPS_OK;
    const size_t smem = (size_t)(PC_ROWS + 2) * (W + 2) * sizeof(float4);
    UPS_REQUIRE(smem <= 200 * 1024, "parts_conv_fwd: W=%d needs %zu bytes of shared memory", W, smem);
    const int n_strips = (int)cdiv(H, PC_ROWS);
    long long want = cdiv(8ll * NUM_SMS, B);
    if (want < 1) want = 1;
    if (want > n_strips) want = n_strips;
    const int spc = (int)cdiv(n_strips, want);
    const dim3 grid((unsigned)cdiv(n_strips, spc), B);
    cudaStream_t st = as_stream(stream);
#define UPS_PC_FWD(CCH)                                                                                              \
    do {                                                                                                             \
        UPS_CUDA(cudaFuncSetAttribute(parts_conv_fwd_kernel<CCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                      (int)smem));                                                                   \
        parts_conv_fwd_kernel<CCH><<<grid, PC_THREADS, smem, st>>>(img, mask, V, bias, out_pm, B, H, W, K, Co, spc); \
    } while (0)
    if (Co <= 32) UPS_PC_FWD(1);
    else if (Co <= 64) UPS_PC_FWD(2);
    else UPS_PC_FWD(4);
#undef UPS_PC_FWD
    return after_launch("parts_conv_fwd_kernel");
}
