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
#include <stdlib.h>

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


// ------------------------------------------------------------------ backward
// g_h [K*B,H,W,Co] (part-major) -> dmask [B,H,W,K], dV [9,3,Co], db [Co].
//   R_k[q,c]   = sum_{t,o} g_h[k*B+b, q - off_t, o] V[t,c,o]          (the cotangent of the part image: dense)
//   dmask[q,k] = sum_c img[q,c] R_k[q,c]
//   dV[t,c,o]  = sum_{k,b,q} img[q,c] mask[q,k] g_h[k*B+b, q - off_t, o]  (sparse: the pixels of part k only)
//   db[o]      = sum of g_h over everything
// Kernel 1 (dense, FMA-bound: 27*Co MAC per plane pixel): one CTA per plane (k,b) walks the rows top to bottom, two
// rows per step.  A thread owns a column x: it reads its two g_h pixels (Co floats each, straight to registers, no
// shared-memory staging), forms D[p][t][c] = sum_o g_h[p,o] V[t,c,o] (V from shared memory, broadcast) and parks
// the 27 values per pixel in a 4-row ring; R_k of the two rows just completed is then 9 ring reads per channel:
// R[q,c] = sum_t D[q - off_t][t][c].  No halo is recomputed (whole rows, whole plane).  db rides along as
// per-thread column sums.
// Kernel 2 (sparse gather): a warp owns a run of pixels of one sample; lane = output channel; for every pixel the
// 9 rows g_h[label-plane, q - off_t, :] are gathered (128-byte coalesced) into 27 register accumulators.
// Partials are reduced in a fixed order: deterministic.
constexpr int PCB_THREADS = 128;

template <int CO, int MINB>
__global__ void __launch_bounds__(PCB_THREADS, MINB) parts_conv_bwd_data_kernel(const float* __restrict__ g_h,
                                                                          const float* __restrict__ img,
                                                                          const float* __restrict__ V,
                                                                          float* __restrict__ dm_planes,
                                                                          float* __restrict__ ws_db, int B, int H, int W,
                                                                          int K) {
    constexpr int CO4 = CO / 4;
    extern __shared__ float4 smem4[];
    float* sVT = reinterpret_cast<float*>(smem4);  // [CO][28]: V transposed, (tap, channel) pair index fastest, [27] = 0
    float* sD = sVT + 28 * CO;                     // ring [4][W+2][27]; columns 0 and W+1 stay zero
    const int Wr = W + 2;
    float* sRed = sD + 4 * Wr * 27;                // [PCB_THREADS/32][CO] for the db reduction
    constexpr int GST = CO + 4;                    // staging stride per pixel (16-byte loads of 8 lanes: conflict-free)
    float* sStage = sRed + (PCB_THREADS / 32) * CO;  // [2 rows][PCB_THREADS][GST]: the thread's own two g_h pixels, one step ahead
    sStage += (4 - ((28 * CO + 4 * Wr * 27 + (PCB_THREADS / 32) * CO) & 3)) & 3;  // 16-byte alignment
    const int n = blockIdx.x;                      // plane k*B + b
    const int k = n / B, b = n - k * B;
    const int tid = threadIdx.x;
    for (int i = tid; i < 28 * CO; i += PCB_THREADS) {
        const int o = i / 28, tc = i - o * 28;
        sVT[i] = tc < 27 ? __ldg(V + tc * CO + o) : 0.f;
    }
    for (int i = tid; i < 4 * Wr * 27; i += PCB_THREADS) sD[i] = 0.f;
    const size_t P = (size_t)H * W;
    const float* gp = g_h + (size_t)n * P * CO;
    const float* ib = img + (size_t)b * P * 3;
    float* dm = dm_planes + ((size_t)b * K + k) * P;  // plane-major [B][K][P]: coalesced; transposed by the finish kernel
    float dbacc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) dbacc[o] = 0.f;
    __syncthreads();
    // step s: D rows 2s, 2s+1 -> ring slots (2s)&3, (2s+1)&3; barrier; R rows 2s-1, 2s (need D rows 2s-2 .. 2s+1);
    // barrier.  The g_h pixels of step s+1 are copied global -> shared (cp.async, the thread's own column, so no
    // barrier is involved) at the start of step s and picked up at the start of step s+1: a whole step in flight.
    // (Columns beyond the block size, W > 128, load directly.)
    const int n_steps = (H + 1) / 2 + 1;
    const bool one_col = W <= PCB_THREADS;
    float4 ga[CO4], gb[CO4];
    float* st_a = sStage + tid * GST;
    float* st_b = st_a + PCB_THREADS * GST;
    auto stage_rows = [&](int s) {  // asynchronous copy of rows 2s, 2s+1 at column tid; zeros outside the image
        const int ya = 2 * s, yb = 2 * s + 1;
        const bool oka = ya < H && tid < W, okb = yb < H && tid < W;
        const float* srca = oka ? gp + ((size_t)ya * W + tid) * CO : gp;
        const float* srcb = okb ? gp + ((size_t)yb * W + tid) * CO : gp;
#pragma unroll
        for (int j = 0; j < CO4; ++j) {
            const unsigned da = (unsigned)__cvta_generic_to_shared(st_a + 4 * j);
            const unsigned db_ = (unsigned)__cvta_generic_to_shared(st_b + 4 * j);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(srca + 4 * j), "r"(oka ? 16 : 0) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(db_), "l"(srcb + 4 * j), "r"(okb ? 16 : 0) : "memory");
        }
    };
    auto load_rows = [&](int s, int x) {
        const int ya = 2 * s, yb = 2 * s + 1;
        const float* srca = gp + ((size_t)ya * W + x) * CO;
#pragma unroll
        for (int j = 0; j < CO4; ++j) ga[j] = (ya < H && x < W) ? ld4_stream(srca + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < CO4; ++j)
            gb[j] = (yb < H && x < W) ? ld4_stream(srca + (size_t)W * CO + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (one_col) stage_rows(0);
    for (int s = 0; s < n_steps; ++s) {
        const int ya = 2 * s, yb = 2 * s + 1;
        const int slot_a = ya & 3;
        // the image pixels of the two R rows of this step: requested now, used after the barrier
        float im[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int y = 2 * s - 1 + h;
            const bool ok = one_col && tid < W && y >= 0 && y < H;
            const size_t q = (size_t)(ok ? y : 0) * W + (ok ? tid : 0);
#pragma unroll
            for (int c = 0; c < 3; ++c) im[h][c] = ok ? __ldg(ib + q * 3 + c) : 0.f;
        }
        if (one_col) {
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int j = 0; j < CO4; ++j) {
                ga[j] = *reinterpret_cast<const float4*>(st_a + 4 * j);
                gb[j] = *reinterpret_cast<const float4*>(st_b + 4 * j);
            }
        }
        for (int x = tid; x < W; x += PCB_THREADS) {
            float* dsta = sD + (slot_a * Wr + x + 1) * 27;
            float* dstb = dsta + Wr * 27;
            if (ya < H) {
                // both rows side by side: one 16-byte load of V feeds 4 packed FMAs of each row (row b = zeros past H)
                if (!one_col) load_rows(s, x);
#pragma unroll
                for (int j = 0; j < CO4; ++j) {
                    dbacc[4 * j] += ga[j].x + gb[j].x; dbacc[4 * j + 1] += ga[j].y + gb[j].y;
                    dbacc[4 * j + 2] += ga[j].z + gb[j].z; dbacc[4 * j + 3] += ga[j].w + gb[j].w;
                }
                if (one_col) stage_rows(s + 1);  // the staged pixels are in registers: refill for the next step
                // the 28 (tap, channel) columns in two halves of 16 and 12: 16 packed accumulators live at a time
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int q0 = 4 * half, nq = half ? 3 : 4;
                    pk::f2 acca[8], accb[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acca[i] = accb[i] = 0ull;  // +0.0f pairs
#pragma unroll
                    for (int j = 0; j < CO4; ++j) {
                        const float av[4] = {ga[j].x, ga[j].y, ga[j].z, ga[j].w};
                        const float bv[4] = {gb[j].x, gb[j].y, gb[j].z, gb[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const pk::f2 sa = pk::splat(av[e]), sb = pk::splat(bv[e]);
                            const ulonglong2* vr = reinterpret_cast<const ulonglong2*>(sVT + (4 * j + e) * 28) + q0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (q < nq) {
                                    const ulonglong2 v = vr[q];  // (V[tc], V[tc+1]), (V[tc+2], V[tc+3])
                                    acca[2 * q] = pk::fma2(sa, v.x, acca[2 * q]);
                                    accb[2 * q] = pk::fma2(sb, v.x, accb[2 * q]);
                                    acca[2 * q + 1] = pk::fma2(sa, v.y, acca[2 * q + 1]);
                                    accb[2 * q + 1] = pk::fma2(sb, v.y, accb[2 * q + 1]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < 2 * nq) {
                            const int tc = 2 * (2 * q0 + i);
                            float lo, hi;
                            pk::unpack(acca[i], lo, hi);
                            dsta[tc] = lo;
                            if (tc + 1 < 27) dsta[tc + 1] = hi;
                            pk::unpack(accb[i], lo, hi);
                            dstb[tc] = lo;
                            if (tc + 1 < 27) dstb[tc + 1] = hi;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int tc = 0; tc < 27; ++tc) dsta[tc] = dstb[tc] = 0.f;  // rows below the image contribute nothing
            }
        }
        __syncthreads();
        // slot of D row 2s-2+i, i = 0..3
        const int slots[4] = {(slot_a + 2) & 3, (slot_a + 3) & 3, slot_a, slot_a + 1};
        for (int x = tid; x < W; x += PCB_THREADS) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int y = 2 * s - 1 + h;
                if (y < 0 || y >= H) continue;
                float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    // D row y+1-dy = 2s-2 + (h+2-dy); row -1 is the zero-initialised slot 3, rows >= H hold zeros
                    const float* row = sD + (slots[h + 2 - dy] * Wr + x + 2) * 27 + 9 * dy;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float* d = row + 3 * dx - dx * 27;  // column x+1-dx (+1 for the zero border)
                        r0 += d[0]; r1 += d[1]; r2 += d[2];
                    }
                }
                const size_t q = (size_t)y * W + x;
                float c0 = im[h][0], c1 = im[h][1], c2 = im[h][2];
                if (!one_col) { c0 = __ldg(ib + q * 3); c1 = __ldg(ib + q * 3 + 1); c2 = __ldg(ib + q * 3 + 2); }
                dm[q] = fmaf(c2, r2, fmaf(c1, r1, c0 * r0));
            }
        }
        __syncthreads();
    }
    // db: thread partials -> warp (fixed xor tree) -> CTA (warp order) -> ws_db[n][CO]
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        float v = dbacc[o];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((tid & 31) == 0) sRed[(tid >> 5) * CO + o] = v;
    }
    __syncthreads();
    for (int o = tid; o < CO; o += PCB_THREADS) {
        float v = 0.f;
        for (int w = 0; w < PCB_THREADS / 32; ++w) v += sRed[w * CO + o];
        ws_db[(size_t)n * CO + o] = v;
    }
}

// dV partials: warp = run of `run` pixels of sample b (run % 32 == 0); lane = output channel (+32*c).
// Per pixel q with label k: 9 gathers g_h[k-plane][q - off_t][lane] and 27 FMAs; (y, x), the tap offsets and the
// tap validity come from incremental integer updates (no division, no 64-bit arithmetic in the loop).
template <int CCH>
__global__ void __launch_bounds__(256) parts_conv_bwd_filter_kernel(const float* __restrict__ g_h,
                                                                    const float* __restrict__ img,
                                                                    const float* __restrict__ mask,
                                                                    float* __restrict__ ws_dV, int B, int H, int W, int K,
                                                                    int Co, int run, int runs_per_sample) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // global warp = (b, run index)
    if (gw >= B * runs_per_sample) return;
    const int b = gw / runs_per_sample, ri = gw - b * runs_per_sample;
    const int P = H * W;
    const float* mb = mask + (size_t)b * P * K;
    const float* ib = img + (size_t)b * P * 3;
    const size_t plane = (size_t)P * Co;
    const float* gb = g_h + (size_t)b * plane;  // plane kk of this sample: gb + kk*B*plane
    float acc[27][CCH];
#pragma unroll
    for (int i = 0; i < 27; ++i)
#pragma unroll
        for (int c = 0; c < CCH; ++c) acc[i][c] = 0.f;
    int toff[9];  // element offset of the output pixel q - off_t relative to q
#pragma unroll
    for (int t = 0; t < 9; ++t) toff[t] = ((1 - t / 3) * W + (1 - t % 3)) * Co;
    const int q_beg = ri * run, q_end = min(P, q_beg + run);
    int y = q_beg / W, x = q_beg - y * W;
    for (int q0 = q_beg; q0 < q_end; q0 += 32) {
        const int qm = q0 + lane;
        int lab = -2;
        float sv = 0.f, i0 = 0.f, i1 = 0.f, i2 = 0.f;
        if (qm < q_end) {
            int nz = 0;
            for (int kk = 0; kk < K; ++kk) {
                const float v = __ldg(mb + (size_t)qm * K + kk);
                if (v != 0.f) {
                    if (nz == 0) { lab = kk; sv = v; }
                    ++nz;
                }
            }
            if (nz > 1) lab = -1;
            i0 = __ldg(ib + (size_t)qm * 3); i1 = __ldg(ib + (size_t)qm * 3 + 1); i2 = __ldg(ib + (size_t)qm * 3 + 2);
        }
        const int cnt = min(32, q_end - q0);
        for (int j = 0; j < cnt; ++j, ++x) {
            if (x == W) { x = 0; ++y; }
            const int lj = __shfl_sync(0xffffffffu, lab, j);
            if (lj == -2) continue;
            const float sj = __shfl_sync(0xffffffffu, sv, j);
            const float c0 = __shfl_sync(0xffffffffu, i0, j), c1 = __shfl_sync(0xffffffffu, i1, j),
                        c2 = __shfl_sync(0xffffffffu, i2, j);
            // output pixel of tap (dy,dx) is (y+1-dy, x+1-dx): row/column validity
            const bool vy[3] = {y + 1 < H, true, y > 0};
            const bool vx[3] = {x + 1 < W, true, x > 0};
            const int qoff = (q0 + j) * Co + lane;
            const int k_beg = lj >= 0 ? lj : 0, k_end = lj >= 0 ? lj + 1 : K;
            for (int kk = k_beg; kk < k_end; ++kk) {
                float m = sj;
                if (lj < 0) {
                    m = __ldg(mb + (size_t)(q0 + j) * K + kk);
                    if (m == 0.f) continue;
                }
                // mask_parts: fl(image * mask) per channel
                const float p0 = __fmul_rn(c0, m), p1 = __fmul_rn(c1, m), p2 = __fmul_rn(c2, m);
                const float* gq = gb + (size_t)kk * B * plane + qoff;
                float g[9][CCH];
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int c = 0; c < CCH; ++c)
                        g[t][c] = (vy[t / 3] && vx[t % 3] && lane + 32 * c < Co) ? __ldg(gq + toff[t] + 32 * c) : 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t)
#pragma unroll
                    for (int c = 0; c < CCH; ++c) {
                        acc[3 * t][c] = fmaf(p0, g[t][c], acc[3 * t][c]);
                        acc[3 * t + 1][c] = fmaf(p1, g[t][c], acc[3 * t + 1][c]);
                        acc[3 * t + 2][c] = fmaf(p2, g[t][c], acc[3 * t + 2][c]);
                    }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 27; ++i)
#pragma unroll
        for (int c = 0; c < CCH; ++c) {
            const int o = lane + 32 * c;
            if (o < Co) ws_dV[((size_t)gw * 27 + i) * Co + o] = acc[i][c];
        }
}

// dm_planes [B][K][P] -> dmask [B][P][K]; with probs: (+ g_extra) through the straight-through estimator (identity)
// and the softmax backward -> dlogits.  One thread per pixel; plane reads and the [P][K] writes are both coalesced
// per warp (32 pixels x 4 bytes per plane; K*4 contiguous bytes per thread).
template <int KP>
__global__ void __launch_bounds__(256) parts_conv_bwd_finish_kernel(const float* __restrict__ dm_planes,
                                                                    const float* __restrict__ probs,
                                                                    const float* __restrict__ g_extra,
                                                                    float* __restrict__ dmask, int B, int P, int K) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * P) return;
    const size_t b = i / P, q = i - b * P;
    const float* src = dm_planes + b * K * P + q;
    float d[KP], p[KP];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        d[k] = k < K ? __ldcs(src + (size_t)k * P) : 0.f;
        if (k < K && g_extra != nullptr) d[k] += __ldg(g_extra + i * K + k);
        p[k] = (k < K && probs != nullptr) ? __ldg(probs + i * K + k) : 0.f;
        dot = fmaf(d[k], p[k], dot);
    }
    if (probs != nullptr) {
#pragma unroll
        for (int k = 0; k < KP; ++k) d[k] = p[k] * (d[k] - dot);
    }
    float* dst = dmask + i * K;
    if ((K & 3) == 0) {
#pragma unroll
        for (int k = 0; k < KP; k += 4)
            if (k < K) st4_stream(dst + k, make_float4(d[k], d[k + 1], d[k + 2], d[k + 3]));
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) dst[k] = d[k];
    }
}

// out[i] = sum_r part[r][i] in a fixed order: thread j of a 256-thread CTA sums r = j, j+256, ..., then a tree
__global__ void __launch_bounds__(256) fixed_order_sum_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                              int n_rows, int n_cols) {
    __shared__ float s[256];
    const int i = blockIdx.x;
    float v = 0.f;
    for (int r = threadIdx.x; r < n_rows; r += 256) v += part[(size_t)r * n_cols + i];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int m = 128; m >= 1; m >>= 1) {
        if (threadIdx.x < m) s[threadIdx.x] += s[threadIdx.x + m];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[i] = s[0];
}

int pcb_run_len(int B, int P) {  // pixels per warp of the filter kernel: ~96 warps per SM over the chip (3-4 waves)
    long long warps = 96ll * NUM_SMS;
    long long per_sample = cdiv(warps, B > 0 ? B : 1);
    long long run = cdiv(P, per_sample);
    run = cdiv(run, 32) * 32;
    return (int)(run < 32 ? 32 : run);
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

extern "C" size_t ups_parts_conv_bwd_workspace_bytes(int B, int H, int W, int K, int Co) {
    if (B <= 0 || H <= 0 || W <= 0 || K < 1 || Co < 1) return 256;
    const int P = H * W;
    const int run = pcb_run_len(B, P);
    const size_t runs = (size_t)B * cdiv(P, run);
    return (runs * 27 * Co + (size_t)K * B * Co + (size_t)B * K * P) * sizeof(float) + 256;
}

extern "C" int ups_parts_conv_bwd(const float* g_out_pm, const float* img, const float* mask, const float* V,
                                  const float* probs, const float* g_extra, float* dmask, float* dV, float* db, int B,
                                  int H, int W, int K, int C, int Co, void* ws, size_t ws_bytes, void* stream) {
    UPS_REQUIRE(g_out_pm && img && mask && V && dmask, "parts_conv_bwd: null pointer");
    UPS_REQUIRE(B >= 0 && H > 0 && W > 0, "parts_conv_bwd: bad sizes B=%d H=%d W=%d", B, H, W);
    UPS_REQUIRE(K >= 1 && K <= 32, "parts_conv_bwd: K=%d outside [1,32]", K);
    UPS_REQUIRE(C == 3, "parts_conv_bwd: C=%d (the part images are 3-channel)", C);
    UPS_REQUIRE(Co == 8 || Co == 16 || Co == 32 || Co == 64, "parts_conv_bwd: Co=%d must be 8, 16, 32 or 64", Co);
    UPS_REQUIRE((long long)B * H * W * K < (1ll << 31), "parts_conv_bwd: K*B*H*W >= 2^31");
    UPS_REQUIRE(aligned16(g_out_pm) && aligned16(V) && aligned16(dmask), "parts_conv_bwd: g_out, V and dmask must be 16-byte aligned");
    if (B == 0) return UPS_OK;
    const size_t need = ups_parts_conv_bwd_workspace_bytes(B, H, W, K, Co);
    UPS_REQUIRE(ws != nullptr && ws_bytes >= need, "parts_conv_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    const int P = H * W;
    const int run = pcb_run_len(B, P);
    const int runs_per_sample = (int)cdiv(P, run);
    float* ws_dV = reinterpret_cast<float*>(ws);
    float* ws_db = ws_dV + (size_t)B * runs_per_sample * 27 * Co;
    float* dm_planes = ws_db + (size_t)K * B * Co;
    cudaStream_t st = as_stream(stream);
    const bool use_tc = parts_conv_bwd_tc_ok(B, H, W, K, Co);   // tcgen05 kernel (parts_conv_bwd_tc.cu): Co = 32, W in {128, 256}
    if (use_tc) {
        if (int rc = parts_conv_bwd_tc_launch(g_out_pm, img, V, dm_planes, ws_db, B, H, W, K, st)) return rc;
    } else {
    const size_t smem = (size_t)(28 * Co + 4 * (W + 2) * 27 + (PCB_THREADS / 32) * Co + 4 + 2 * PCB_THREADS * (Co + 4)) * sizeof(float);
    UPS_REQUIRE(smem <= 200 * 1024, "parts_conv_bwd: W=%d Co=%d needs %zu bytes of shared memory", W, Co, smem);
#define UPS_PCB(CO, MINB)                                                                                              \
    do {                                                                                                               \
        UPS_CUDA(cudaFuncSetAttribute(parts_conv_bwd_data_kernel<CO, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                     \
        parts_conv_bwd_data_kernel<CO, MINB><<<K * B, PCB_THREADS, smem, st>>>(g_out_pm, img, V, dm_planes, ws_db, B, H, \
                                                                               W, K);                                  \
    } while (0)
    // Co = 32: 2 CTAs per SM without spills (237 registers) measured 3 % faster than 3 CTAs with a 168-register cap
    const char* env_minb = getenv("UPS_PCB_MINB");  // tuning knob (profiles/r01_tuning.md)
    const int minb = env_minb ? atoi(env_minb) : 2;
    if (Co == 8) UPS_PCB(8, 3);
    else if (Co == 16) UPS_PCB(16, 3);
    else if (Co == 32 && minb == 2) UPS_PCB(32, 2);
    else if (Co == 32) UPS_PCB(32, 3);
    else UPS_PCB(64, 2);
#undef UPS_PCB
    if (int rc = after_launch("parts_conv_bwd_data_kernel")) return rc;
    }
    {
        const unsigned grid = (unsigned)cdiv((long long)B * P, 256);
        if (K <= 8) parts_conv_bwd_finish_kernel<8><<<grid, 256, 0, st>>>(dm_planes, probs, g_extra, dmask, B, P, K);
        else if (K <= 16) parts_conv_bwd_finish_kernel<16><<<grid, 256, 0, st>>>(dm_planes, probs, g_extra, dmask, B, P, K);
        else if (K <= 24) parts_conv_bwd_finish_kernel<24><<<grid, 256, 0, st>>>(dm_planes, probs, g_extra, dmask, B, P, K);
        else parts_conv_bwd_finish_kernel<32><<<grid, 256, 0, st>>>(dm_planes, probs, g_extra, dmask, B, P, K);
        if (int rc = after_launch("parts_conv_bwd_finish_kernel")) return rc;
    }
    if (dV != nullptr) {
        const int n_warps = B * runs_per_sample;
        const unsigned grid = (unsigned)cdiv(n_warps, 8);
        if (Co <= 32)
            parts_conv_bwd_filter_kernel<1><<<grid, 256, 0, st>>>(g_out_pm, img, mask, ws_dV, B, H, W, K, Co, run, runs_per_sample);
        else
            parts_conv_bwd_filter_kernel<2><<<grid, 256, 0, st>>>(g_out_pm, img, mask, ws_dV, B, H, W, K, Co, run, runs_per_sample);
        if (int rc = after_launch("parts_conv_bwd_filter_kernel")) return rc;
        fixed_order_sum_kernel<<<27 * Co, 256, 0, st>>>(ws_dV, dV, n_warps, 27 * Co);
        if (int rc = after_launch("fixed_order_sum_kernel")) return rc;
    }
    if (db != nullptr) {
        fixed_order_sum_kernel<<<Co, 256, 0, st>>>(ws_db, db, K * B, Co);
        if (int rc = after_launch("fixed_order_sum_kernel")) return rc;
    }
    return UPS_OK;
}
