// SURVEY.md 8f N4 (decoder side): the hourglass decoder's first convolution consuming the part
// assignment directly, so that the injected map [B,h,w,F+K] (1.34 GB at CUB B=256) never exists.
//
//   injected = concat(sum_k unpool_features(feat, mask), mask)       cub/code/SB_model48i/model.py:482-484
//   h        = conv2d(injected, V[3,3,F+K,Co]) + b                   model.py:96 (dd, :485), cub/code/nn.py:617-664
//
// Re-association: with the per-sample table  G[b,t,k,o] = sum_f feat[b,k,f] V[t,f,o] + V[t,F+k,o]  (t = 3i+j)
//   h[b,y,x,o] = b[o] + sum_t sum_k mask[b,y+i-1,x+j-1,k] G[b,t,k,o]
// and the decoding mask is straight-through hard (one non-zero per pixel, model.py:434-436,473), so the
// forward is 9 table-row gathers per pixel (exact for any mask: pixels with several non-zeros take a dense
// loop).  Backward: dmask is a dense (P x 9Co).(9Co x K) contraction on the g_out tile in shared memory,
// dG a label-sorted segmented sum (deterministic, no atomics), db a column sum.
#include "common.cuh"

namespace ups {
namespace {

constexpr int IC_FWD_THREADS = 256;  // 8 warps = 8 rows of a strip
constexpr int IC_FWD_ROWS = 8;
constexpr int IC_BWD_THREADS = 288;  // CUDA-core backward: 9 warps, one per filter tap in the dG phase
constexpr int IC_MMA_THREADS = 256;  // tensor-core backward: 8 warps, MTW m-tiles each
constexpr int IC_TW = 32;            // backward tile width (pixels); thread = pixels (x, x+16)
constexpr int IC_SMEM_BUDGET = 224 * 1024;   // per CTA (227 KB is the hardware limit)
constexpr int IC_SMEM_HALF = 112 * 1024 + 512;     // two CTAs per SM

// (label, value) of a mask pixel: label >= 0 one non-zero; -1 several non-zeros; -2 none
__device__ __forceinline__ int2 compact_pixel(const float* __restrict__ m, int K) {
    int lab = -2, nz = 0;
    float val = 0.f;
    if ((K & 3) == 0) {
        for (int k = 0; k < K; k += 4) {
            const float4 v = ld4(m + k);
#define UPS_IC_SEE(val_, j_)                             \
    if ((val_) != 0.f) {                                 \
        if (nz == 0) { lab = k + (j_); val = (val_); }   \
        ++nz;                                            \
    }
            UPS_IC_SEE(v.x, 0) UPS_IC_SEE(v.y, 1) UPS_IC_SEE(v.z, 2) UPS_IC_SEE(v.w, 3)
#undef UPS_IC_SEE
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float v = __ldg(m + k);
            if (v != 0.f) {
                if (nz == 0) { lab = k; val = v; }
                ++nz;
            }
        }
    }
    if (nz > 1) lab = -1;
    return make_int2(lab, __float_as_int(val));
}

// ------------------------------------------------------------------ table G = feat.V[:F] + V[F:]
__global__ void __launch_bounds__(256) inject_conv_table_fwd_kernel(const float* __restrict__ feat,
                                                                    const float* __restrict__ V, float* __restrict__ G,
                                                                    int K, int F, int Co) {
    extern __shared__ float sf[];  // feat[b] : K*F
    const int t = blockIdx.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < K * F; i += blockDim.x) sf[i] = __ldg(feat + (size_t)b * K * F + i);
    __syncthreads();
    const float* Vt = V + (size_t)t * (F + K) * Co;
    for (int idx = threadIdx.x; idx < K * Co; idx += blockDim.x) {
        const int k = idx / Co, o = idx - k * Co;
        float acc = 0.f;
        for (int f = 0; f < F; ++f) acc = fmaf(sf[k * F + f], __ldg(Vt + (size_t)f * Co + o), acc);
        acc += __ldg(Vt + (size_t)(F + k) * Co + o);
        G[(((size_t)b * 9 + t) * K + k) * Co + o] = acc;
    }
}

// dfeat[b,k,f] = sum_{t,o} dG[b,t,k,o] V[t,f,o]: one CTA per sample, one thread per (k,f), dG[b] in shared memory
__global__ void __launch_bounds__(256) inject_conv_table_bwd_feat_kernel(const float* __restrict__ dG,
                                                                         const float* __restrict__ V,
                                                                         float* __restrict__ dfeat, int K, int F, int Co) {
    extern __shared__ float4 sd4[];  // dG[b] : 9*K*Co
    float* sd = reinterpret_cast<float*>(sd4);
    const int b = blockIdx.x;
    const int n = 9 * K * Co;
    for (int i = 4 * threadIdx.x; i < n; i += 4 * blockDim.x) st4(sd + i, ld4(dG + (size_t)b * n + i));
    __syncthreads();
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < K * F; idx += gridDim.y * blockDim.x) {
        const int k = idx / F, f = idx - k * F;  // consecutive threads: consecutive f (distinct V rows), same dG row
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int t = 0; t < 9; ++t) {
            const float* Vr = V + ((size_t)t * (F + K) + f) * Co;
            const float* dr = sd + (t * K + k) * Co;
#pragma unroll 8
            for (int o = 0; o < Co; o += 4) {  // unrolled: the 8 loads of a tap are in flight together
                const float4 v = ld4(Vr + o);
                const float4 d = *reinterpret_cast<const float4*>(dr + o);
                a0 = fmaf(d.x, v.x, a0); a1 = fmaf(d.y, v.y, a1); a2 = fmaf(d.z, v.z, a2); a3 = fmaf(d.w, v.w, a3);
            }
        }
        dfeat[((size_t)b * K + k) * F + f] = (a0 + a1) + (a2 + a3);
    }
}

// dV[t,c,o] = sum_{b,k} feat[b,k,c] dG[b,t,k,o]  (c < F);   dV[t,F+k,o] = sum_b dG[b,t,k,o].
// grid (ceil((F+K)/4), 9): one CTA per 4 filter rows; warp w sums the samples b = w, w+8, ... (lane = o), then the
// 8 partial rows are added in warp order: deterministic.
__global__ void __launch_bounds__(256) inject_conv_table_bwd_filter_kernel(const float* __restrict__ dG,
                                                                           const float* __restrict__ feat,
                                                                           float* __restrict__ dV, int B, int K, int F,
                                                                           int Co) {
    extern __shared__ float sp[];  // [8 warps][4][Co]
    const int c0 = 4 * blockIdx.x, t = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int o0 = 0; o0 < Co; o0 += 32) {
        const int o = o0 + lane;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (o < Co) {
            const bool quad = (F & 3) == 0 && c0 + 3 < F;  // four feature rows: one 16-byte feat load per (b,k)
            for (int b = warp; b < B; b += 8) {
                const float* d = dG + (((size_t)b * 9 + t) * K) * Co + o;
                const float* fr = feat + (size_t)b * K * F;
                if (quad) {
#pragma unroll 8
                    for (int k = 0; k < K; ++k) {
                        const float dv = __ldg(d + (size_t)k * Co);
                        const float4 f4 = ld4(fr + (size_t)k * F + c0);
                        acc[0] = fmaf(f4.x, dv, acc[0]); acc[1] = fmaf(f4.y, dv, acc[1]);
                        acc[2] = fmaf(f4.z, dv, acc[2]); acc[3] = fmaf(f4.w, dv, acc[3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c0 + j;
                        if (c < F) {
#pragma unroll 8
                            for (int k = 0; k < K; ++k)
                                acc[j] = fmaf(__ldg(fr + (size_t)k * F + c), __ldg(d + (size_t)k * Co), acc[j]);
                        } else if (c < F + K) {
                            acc[j] += __ldg(d + (size_t)(c - F) * Co);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sp[(warp * 4 + j) * Co + o] = acc[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * Co; i += blockDim.x) {
        const int j = i / Co, o = i - j * Co;
        if (c0 + j >= F + K) continue;
        float s2 = 0.f;
        for (int w = 0; w < 8; ++w) s2 += sp[(w * 4 + j) * Co + o];
        dV[((size_t)t * (F + K) + c0 + j) * Co + o] = s2;
    }
}

// ------------------------------------------------------------------ forward
// grid (splits, B); a CTA walks `strips_per_cta` strips of 8 rows x W of one sample.  The strip's mask pixels
// (with a 1-pixel halo) are compacted to (table row offset, value) in shared memory; then one thread per
// (pixel, 4 output channels): 9 x (LDS.64 entry, LDS.128 table row, 4 FMA), one 16-byte streaming store.
template <bool DENSE>
__device__ __forceinline__ void inject_conv_fwd_rows(const float* __restrict__ sG, const float* __restrict__ sB,
                                                     const int2* __restrict__ sM, const float* __restrict__ mb,
                                                     float* __restrict__ ob, int y0, int rows, int H, int W, int K,
                                                     int Co, int Co4, int tid) {
    const int Wp = W + 2, KCo = K * Co;
    // idx = (r*W + x)*Co4 + o4 advances by the block size: carry-propagate instead of dividing
    const int d_px = IC_FWD_THREADS / Co4, d_o4 = IC_FWD_THREADS - d_px * Co4;
    int px0 = tid / Co4, o4 = tid - px0 * Co4;
    int r = px0 / W, x = px0 - r * W;
    for (int idx = tid; idx < rows * W * Co4; idx += IC_FWD_THREADS, o4 += d_o4, x += d_px) {
        if (o4 >= Co4) { o4 -= Co4; ++x; }
        while (x >= W) { x -= W; ++r; }
        const int o = 4 * o4;
        float4 acc = *reinterpret_cast<const float4*>(sB + o);
        const int2* e0 = sM + r * Wp + x;
        const float* g0 = sG + o;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t - 3 * dy;
            const int2 e = e0[dy * Wp + dx];
            if (!DENSE || e.x >= 0) {
                const float v = __int_as_float(e.y);
                const float4 g = *reinterpret_cast<const float4*>(g0 + t * KCo + e.x);
                acc.x = fmaf(v, g.x, acc.x); acc.y = fmaf(v, g.y, acc.y);
                acc.z = fmaf(v, g.z, acc.z); acc.w = fmaf(v, g.w, acc.w);
            } else {  // several non-zeros: dense over k
                const float* m = mb + ((size_t)(y0 + r + dy - 1) * W + (x + dx - 1)) * K;
                for (int k = 0; k < K; ++k) {
                    const float v = __ldg(m + k);
                    if (v != 0.f) {
                        const float4 g = *reinterpret_cast<const float4*>(g0 + t * KCo + k * Co);
                        acc.x = fmaf(v, g.x, acc.x); acc.y = fmaf(v, g.y, acc.y);
                        acc.z = fmaf(v, g.z, acc.z); acc.w = fmaf(v, g.w, acc.w);
                    }
                }
            }
        }
        st4_stream(ob + ((size_t)(y0 + r) * W + x) * Co + o, acc);
    }
}

__global__ void __launch_bounds__(IC_FWD_THREADS) inject_conv_fwd_kernel(const float* __restrict__ mask,
                                                                         const float* __restrict__ G,
                                                                         const float* __restrict__ bias,
                                                                         float* __restrict__ out, int H, int W, int K,
                                                                         int Co, int strips_per_cta) {
    extern __shared__ float4 smem4[];
    float* sG = reinterpret_cast<float*>(smem4);          // [9][K][Co]
    float* sB = sG + 9 * K * Co;                          // [Co]
    int2* sM = reinterpret_cast<int2*>(sB + Co);          // [(8+2)][W+2] (offset of the table row = label*Co, value)
    const int b = blockIdx.y, tid = threadIdx.x;
    const int nG = 9 * K * Co, Co4 = Co >> 2;
    const float* Gb = G + (size_t)b * nG;
    for (int i = 4 * tid; i < nG; i += 4 * IC_FWD_THREADS) st4(sG + i, ld4(Gb + i));
    for (int i = tid; i < Co; i += IC_FWD_THREADS) sB[i] = __ldg(bias + i);

    const int n_strips = (H + IC_FWD_ROWS - 1) / IC_FWD_ROWS;
    const int s_beg = blockIdx.x * strips_per_cta;
    const int s_end = min(n_strips, s_beg + strips_per_cta);
    const int Wp = W + 2;
    const float* mb = mask + (size_t)b * H * W * K;
    float* ob = out + (size_t)b * H * W * Co;
    for (int s = s_beg; s < s_end; ++s) {
        const int y0 = s * IC_FWD_ROWS;
        const int rows = min(IC_FWD_ROWS, H - y0);
        __syncthreads();  // sG visible / previous strip's readers done
        int dense = 0;
        for (int i = tid; i < (rows + 2) * Wp; i += IC_FWD_THREADS) {
            const int r = i / Wp, c = i - r * Wp;
            const int y = y0 - 1 + r, x = c - 1;
            int2 e = make_int2(0, 0);  // empty / outside: row 0 scaled by 0
            if (y >= 0 && y < H && x >= 0 && x < W) {
                e = compact_pixel(mb + ((size_t)y * W + x) * K, K);
                if (e.x == -1) dense = 1;
                else if (e.x < 0) e = make_int2(0, 0);
                else e.x *= Co;
            }
            sM[i] = e;
        }
        dense = __syncthreads_or(dense);
        if (dense)
            inject_conv_fwd_rows<true>(sG, sB, sM, mb, ob, y0, rows, H, W, K, Co, Co4, tid);
        else
            inject_conv_fwd_rows<false>(sG, sB, sM, mb, ob, y0, rows, H, W, K, Co, Co4, tid);
    }
}

// ------------------------------------------------------------------ backward
struct BwdSmem {
    int off_G, off_dG, off_db, off_g, off_list, off_tmp, off_dense, off_cnt, off_base, total;  // bytes
};
__host__ __device__ inline BwdSmem bwd_smem_layout(int KP, int K, int Co, int TH) {
    BwdSmem s;
    const int npx = TH * IC_TW, nch = npx / 32;
    int o = 0;
    s.off_G = o;     o += 4 * 9 * KP * Co;
    s.off_dG = o;    o += 4 * 9 * K * Co;
    s.off_db = o;    o += 4 * 9 * Co;
    s.off_g = o;     o += 4 * (TH + 2) * (IC_TW + 2) * (Co + 4);
    o = (o + 7) & ~7;
    s.off_list = o;  o += 8 * npx;
    s.off_tmp = o;   o += 8 * npx;
    s.off_dense = o; o += 4 * npx;
    s.off_cnt = o;   o += 4 * nch * (K + 1);   // per chunk: K label counts + 1 dense count
    s.off_base = o;  o += 4 * (2 * K + 2);     // bucket starts [K+1] + sizes [K+1]
    s.total = (o + 15) & ~15;
    return s;
}

// grid (splits, B).  Tile = TH x 32 mask pixels q of one sample; the g_out tile carries a 1-pixel halo.
//   dmask[q,k] = sum_t sum_o g_out[q - off_t, o] G[t,k,o]          (off_t = (i-1, j-1))
//   dG[t,k,o] += sum_q mask[q,k] g_out[q - off_t, o]
// If probs != NULL the dmask (+ g_extra) goes through the straight-through estimator (identity, nn.py:154-168)
// and the softmax backward, and dlogits is written instead.
template <int KP>
__global__ void __launch_bounds__(IC_BWD_THREADS, 1) inject_conv_bwd_kernel(
    const float* __restrict__ g_out, const float* __restrict__ mask, const float* __restrict__ G,
    const float* __restrict__ probs, const float* __restrict__ g_extra, float* __restrict__ dmask,
    float* __restrict__ ws_dG, float* __restrict__ ws_db, int H, int W, int K, int Co, int TH, int tiles_x, int n_tiles,
    int tiles_per_cta) {
    extern __shared__ float4 smem4[];
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem4);
    const BwdSmem L = bwd_smem_layout(KP, K, Co, TH);
    float* sG = reinterpret_cast<float*>(sm + L.off_G);      // [9][KP][Co], rows k >= K zero
    float* sdG = reinterpret_cast<float*>(sm + L.off_dG);    // [9][K][Co]
    float* sdb = reinterpret_cast<float*>(sm + L.off_db);    // [9 warps][Co]
    float* sg = reinterpret_cast<float*>(sm + L.off_g);      // [(TH+2)][34][Co+4]
    int2* sList = reinterpret_cast<int2*>(sm + L.off_list);  // label-sorted (q, value)
    int2* sTmp = reinterpret_cast<int2*>(sm + L.off_tmp);    // per pixel (label | rank<<8, value)
    int* sDense = reinterpret_cast<int*>(sm + L.off_dense);  // pixels with several non-zeros
    int* sCnt = reinterpret_cast<int*>(sm + L.off_cnt);      // [nch][K+1]
    int* sBase = reinterpret_cast<int*>(sm + L.off_base);    // [K+1] bucket starts
    int* sTot = sBase + K + 1;                               // [K+1] bucket sizes, [K] = #dense

    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int CoP = Co + 4, Co4 = Co >> 2, TWp = IC_TW + 2;
    const int npx = TH * IC_TW, nch = npx >> 5;
    const size_t img = (size_t)b * H * W;

    for (int i = tid; i < 9 * KP * Co; i += IC_BWD_THREADS) {
        const int o = i % Co, k = (i / Co) % KP, t = i / (Co * KP);
        sG[i] = k < K ? __ldg(G + (((size_t)b * 9 + t) * K + k) * Co + o) : 0.f;
    }
    for (int i = tid; i < 9 * K * Co; i += IC_BWD_THREADS) sdG[i] = 0.f;
    for (int i = tid; i < 9 * Co; i += IC_BWD_THREADS) sdb[i] = 0.f;

    const int t_beg = blockIdx.x * tiles_per_cta;
    const int t_end = min(n_tiles, t_beg + tiles_per_cta);
    for (int tile = t_beg; tile < t_end; ++tile) {
        const int y0 = (tile / tiles_x) * TH, x0 = (tile % tiles_x) * IC_TW;
        __syncthreads();  // previous tile fully consumed; initialisation visible
        // (1) g_out tile with halo, zero outside the image
        for (int i = tid; i < (TH + 2) * TWp * Co4; i += IC_BWD_THREADS) {
            const int px = i / Co4, o4 = i - px * Co4;
            const int r = px / TWp, c = px - r * TWp;
            const int y = y0 - 1 + r, x = x0 - 1 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H && x >= 0 && x < W) v = ld4_stream(g_out + ((img + (size_t)y * W + x) * Co + 4 * o4));
            st4(sg + px * CoP + 4 * o4, v);
        }
        // (2a) compact the tile's mask pixels, count labels per 32-pixel chunk
        for (int i = tid; i < nch * (K + 1); i += IC_BWD_THREADS) sCnt[i] = 0;
        __syncthreads();
        for (int ch = warp; ch < nch; ch += IC_BWD_THREADS / 32) {
            const int i = ch * 32 + lane;
            const int y = y0 + (i >> 5), x = x0 + (i & 31);
            int2 e = make_int2(-2, 0);
            if (y < H && x < W) e = compact_pixel(mask + (img + (size_t)y * W + x) * K, K);
            const unsigned peers = __match_any_sync(0xffffffffu, e.x);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0 && e.x >= -1) sCnt[ch * (K + 1) + (e.x >= 0 ? e.x : K)] = __popc(peers);
            sTmp[i] = make_int2((e.x & 0xff) | (rank << 8), e.y);
        }
        __syncthreads();
        // (2b) exclusive prefix of the counts: over chunks per label, then over labels
        if (tid <= K) {
            int run = 0;
            for (int ch = 0; ch < nch; ++ch) {
                const int c = sCnt[ch * (K + 1) + tid];
                sCnt[ch * (K + 1) + tid] = run;
                run += c;
            }
            sTot[tid] = run;  // [0..K) label totals, [K] pixels with several non-zeros
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int k = 0; k < K; ++k) {
                sBase[k] = run;
                run += sTot[k];
            }
            sBase[K] = run;
        }
        __syncthreads();
        // (2c) scatter into label-sorted order (stable: chunk order, then lane order)
        for (int i = tid; i < npx; i += IC_BWD_THREADS) {
            const int2 e = sTmp[i];
            const int lab = (int)(signed char)(e.x & 0xff), rank = e.x >> 8, ch = i >> 5;
            if (lab >= 0)
                sList[sBase[lab] + sCnt[ch * (K + 1) + lab] + rank] = make_int2(i, e.y);
            else if (lab == -1)
                sDense[sCnt[ch * (K + 1) + K] + rank] = i;
        }
        __syncthreads();

        // (3) dmask for two pixels (x, x+16) of one tile row per thread
        if (tid < TH * 16) {
            const int r = tid >> 4, xa = tid & 15;
            float acc[2][KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[0][k] = acc[1][k] = 0.f;
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3, dx = t - 3 * dy;
                const float* ga = sg + ((r + 2 - dy) * TWp + xa + 2 - dx) * CoP;
                const float* gb = ga + 16 * CoP;
                const float* Gt = sG + t * KP * Co;
#pragma unroll 2
                for (int o4 = 0; o4 < Co4; ++o4) {
                    const float4 va = *reinterpret_cast<const float4*>(ga + 4 * o4);
                    const float4 vb = *reinterpret_cast<const float4*>(gb + 4 * o4);
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const float4 w = *reinterpret_cast<const float4*>(Gt + k * Co + 4 * o4);
                        acc[0][k] = fmaf(va.x, w.x, fmaf(va.y, w.y, fmaf(va.z, w.z, fmaf(va.w, w.w, acc[0][k]))));
                        acc[1][k] = fmaf(vb.x, w.x, fmaf(vb.y, w.y, fmaf(vb.z, w.z, fmaf(vb.w, w.w, acc[1][k]))));
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int y = y0 + r, x = x0 + xa + 16 * h;
                if (y < H && x < W) {
                    const size_t base = (img + (size_t)y * W + x) * K;
                    if (g_extra != nullptr) {
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < K) acc[h][k] += __ldg(g_extra + base + k);
                    }
                    if (probs != nullptr) {
                        float p[KP];
                        float dot = 0.f;
#pragma unroll
                        for (int k = 0; k < KP; ++k) {
                            p[k] = k < K ? __ldg(probs + base + k) : 0.f;
                            dot = fmaf(acc[h][k], p[k], dot);
                        }
#pragma unroll
                        for (int k = 0; k < KP; ++k) acc[h][k] = p[k] * (acc[h][k] - dot);
                    }
                    if ((K & 3) == 0) {
#pragma unroll
                        for (int k = 0; k < KP; k += 4)
                            if (k < K)
                                st4_stream(dmask + base + k,
                                           make_float4(acc[h][k], acc[h][k + 1], acc[h][k + 2], acc[h][k + 3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < K) dmask[base + k] = acc[h][k];
                    }
                }
            }
        }

        // (4) dG: warp = tap, lane = o; per label a segmented sum over the sorted list (4 independent chains)
        {
            const int t = warp, dy = t / 3, dx = t - 3 * dy;
            const float* gt = sg + ((2 - dy) * TWp + 2 - dx) * CoP;  // + (ty*TWp + tx)*CoP + o
            for (int oc = 0; oc < Co; oc += 32) {
                const int o = oc + lane;
                if (o < Co) {
                    for (int k = 0; k < K; ++k) {
                        const int beg = sBase[k], end = sBase[k + 1];
                        if (beg == end) continue;
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        int j = beg;
                        for (; j + 4 <= end; j += 4) {
                            const int2 e0 = sList[j], e1 = sList[j + 1], e2 = sList[j + 2], e3 = sList[j + 3];
                            a0 = fmaf(__int_as_float(e0.y), gt[((e0.x >> 5) * TWp + (e0.x & 31)) * CoP + o], a0);
                            a1 = fmaf(__int_as_float(e1.y), gt[((e1.x >> 5) * TWp + (e1.x & 31)) * CoP + o], a1);
                            a2 = fmaf(__int_as_float(e2.y), gt[((e2.x >> 5) * TWp + (e2.x & 31)) * CoP + o], a2);
                            a3 = fmaf(__int_as_float(e3.y), gt[((e3.x >> 5) * TWp + (e3.x & 31)) * CoP + o], a3);
                        }
                        for (; j < end; ++j) {
                            const int2 e0 = sList[j];
                            a0 = fmaf(__int_as_float(e0.y), gt[((e0.x >> 5) * TWp + (e0.x & 31)) * CoP + o], a0);
                        }
                        sdG[(t * K + k) * Co + o] += (a0 + a1) + (a2 + a3);
                    }
                    const int nd = sTot[K];
                    for (int j = 0; j < nd; ++j) {  // pixels with several non-zeros (exact ties, soft masks)
                        const int i = sDense[j];
                        const float gv = gt[((i >> 5) * TWp + (i & 31)) * CoP + o];
                        const float* m = mask + (img + (size_t)(y0 + (i >> 5)) * W + (x0 + (i & 31))) * K;
                        for (int k = 0; k < K; ++k) {
                            const float v = __ldg(m + k);
                            if (v != 0.f) sdG[(t * K + k) * Co + o] = fmaf(v, gv, sdG[(t * K + k) * Co + o]);
                        }
                    }
                    // db: column sums of the tile interior (zero-filled outside the image), rows warp, warp+9, ...
                    float s = 0.f;
                    for (int ty = warp; ty < TH; ty += IC_BWD_THREADS / 32) {
                        const float* row = sg + ((ty + 1) * TWp + 1) * CoP + o;
                        for (int tx = 0; tx < IC_TW; ++tx) s += row[tx * CoP];
                    }
                    sdb[warp * Co + o] += s;
                }
            }
        }
    }
    __syncthreads();
    const size_t slot = (size_t)b * gridDim.x + blockIdx.x;
    for (int i = tid; i < 9 * K * Co; i += IC_BWD_THREADS) ws_dG[slot * 9 * K * Co + i] = sdG[i];
    for (int o = tid; o < Co; o += IC_BWD_THREADS) {
        float s = 0.f;
        for (int w = 0; w < IC_BWD_THREADS / 32; ++w) s += sdb[w * Co + o];
        ws_db[slot * Co + o] = s;
    }
}

// ---- tensor-core variant of the backward (Co in {8,16,32,64,128}) ---------------------------------
// Same tiling, sort and outputs as inject_conv_bwd_kernel; the dense contraction
//   dmask[q, k] = sum_{t,o} g_out[q - off_t, o] G[t,k,o]        (M = pixels, N = K, reduction = 9*Co)
// runs as mma.sync.m16n8k8 TF32 with the 3xTF32 split a = hi + lo (hi = the 19 bits the tensor core keeps):
// hi*hi' + lo*hi' + hi*lo'.  The tensor core adds into its accumulator with truncation, a bias that grows with
// the length of the accumulation chain (measured: 2.5e-5 after 216 chained MMAs), so the hi*hi' products are
// chained over at most 4 k-steps from zero and then added to an fp32 running sum by the CUDA cores (round to
// nearest); the two small terms, 2^-11 of the magnitude, keep one accumulator for the whole reduction.
// A fragments come straight from the padded g_out tile (row stride Co+4 floats: conflict-free), B fragments from
// the pre-split, XOR-swizzled table in shared memory.  Warps 0-7 own MTW m-tiles (16 pixels of one tile row)
// each; in the dG phase Co/4 lanes own a (tap, label) unit and walk its label-sorted pixel list, 4 channels per
// lane, 32/(Co/4) units per warp side by side.
struct BwdMmaSmem {
    int off_Ghi, off_Glo, off_dG, off_db, off_g, off_list, off_tmp, off_dense, off_cnt, off_base, total;
};
__host__ __device__ inline BwdMmaSmem bwd_mma_smem_layout(int KP, int K, int Co, int TH) {
    BwdMmaSmem s;
    const int npx = TH * IC_TW, nch = npx / 32;
    int o = 0;
    s.off_Ghi = o;   o += 4 * 9 * Co * KP;
    s.off_Glo = o;   o += 4 * 9 * Co * KP;
    s.off_dG = o;    o += 4 * 9 * K * Co;
    s.off_db = o;    o += 4 * 9 * Co;
    o = (o + 15) & ~15;
    s.off_g = o;     o += 4 * (TH + 2) * (IC_TW + 2) * (Co + 4);
    o = (o + 7) & ~7;
    s.off_list = o;  o += 8 * npx;
    s.off_tmp = o;   o += 8 * npx;
    s.off_dense = o; o += 4 * npx;
    s.off_cnt = o;   o += 4 * nch * (K + 1);
    s.off_base = o;  o += 4 * (2 * K + 2);
    s.total = (o + 15) & ~15;
    return s;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                         unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// a = hi + lo exactly; hi carries the bits the tensor core reads, lo the remaining 13 (the core keeps its top 11)
__device__ __forceinline__ void split_tf32(float a, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(a) & 0xffffe000u;
    lo = __float_as_uint(__fsub_rn(a, __uint_as_float(hi)));
}
// column swizzle of the [o][KP] table rows so that a B-fragment load (o = os+tig(+4), k = nt*8+gid) touches 32 banks
template <int NT>
__device__ __forceinline__ int b_swizzle(int o) {
    return NT == 2 ? 8 * ((o >> 1) & 1) : NT == 4 ? 8 * (o & 3) : 0;
}

// 16-byte global -> shared copy that bypasses registers; n_src = 0 writes zeros (outside the image)
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, int n_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(n_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// N tiles of 8 parts: KP = 8*NT >= K;  m-tiles per warp: tile rows TH = 4*MTW;  CO > 0: Co known at compile time
template <int NT, int MTW, int CO>
__global__ void __launch_bounds__(IC_MMA_THREADS, MTW <= 2 ? 2 : 1) inject_conv_bwd_mma_kernel(
    const float* __restrict__ g_out, const float* __restrict__ mask, const float* __restrict__ G,
    const float* __restrict__ probs, const float* __restrict__ g_extra, float* __restrict__ dmask,
    float* __restrict__ ws_dG, float* __restrict__ ws_db, int H, int W, int K, int Co_rt, int tiles_x, int n_tiles,
    int tiles_per_cta) {
    const int Co = CO > 0 ? CO : Co_rt;
    constexpr int KP = 8 * NT, TH = 4 * MTW;
    constexpr int TWp = IC_TW + 2, npx = TH * IC_TW, nch = npx >> 5;
    extern __shared__ float4 smem4[];
    unsigned char* sm = reinterpret_cast<unsigned char*>(smem4);
    const BwdMmaSmem L = bwd_mma_smem_layout(KP, K, Co, TH);
    float* sGhi = reinterpret_cast<float*>(sm + L.off_Ghi);  // [9][Co][KP] swizzled: the TF32 part of G
    float* sGlo = reinterpret_cast<float*>(sm + L.off_Glo);  // [9][Co][KP] swizzled: G - hi
    float* sdG = reinterpret_cast<float*>(sm + L.off_dG);    // [9][K][Co]
    float* sdb = reinterpret_cast<float*>(sm + L.off_db);    // [8 warps][Co] (sized for 9)
    float* sg = reinterpret_cast<float*>(sm + L.off_g);      // [(TH+2)][34][Co+4]
    int2* sList = reinterpret_cast<int2*>(sm + L.off_list);
    int2* sTmp = reinterpret_cast<int2*>(sm + L.off_tmp);
    int* sDense = reinterpret_cast<int*>(sm + L.off_dense);
    int* sCnt = reinterpret_cast<int*>(sm + L.off_cnt);
    int* sBase = reinterpret_cast<int*>(sm + L.off_base);
    int* sTot = sBase + K + 1;

    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int CoP = Co + 4, Co4 = Co >> 2;
    const size_t img = (size_t)b * H * W;

    for (int i = tid; i < 9 * KP * Co; i += IC_MMA_THREADS) {  // o fastest: coalesced global reads
        const int o = i % Co, tk = i / Co, k = tk % KP, t = tk / KP;
        const float v = k < K ? __ldg(G + (((size_t)b * 9 + t) * K + k) * Co + o) : 0.f;
        unsigned hi, lo;
        split_tf32(v, hi, lo);
        const int at = (t * Co + o) * KP + (k ^ b_swizzle<NT>(o));
        sGhi[at] = __uint_as_float(hi);
        sGlo[at] = __uint_as_float(lo);
    }
    for (int i = tid; i < 9 * K * Co; i += IC_MMA_THREADS) sdG[i] = 0.f;
    for (int i = tid; i < 9 * Co; i += IC_MMA_THREADS) sdb[i] = 0.f;

    // tile-load roles: Co4 (a power of two) divides the block size, so a thread keeps its channel quad
    const int ld_o4 = tid & (Co4 - 1), ld_px0 = tid / Co4, ld_step = IC_MMA_THREADS / Co4;
    // dG roles: Co4 lanes per (tap, label) unit, 32/Co4 units per warp
    const int lpu_shift = 31 - __clz(Co4);
    const int usub = lane >> lpu_shift, o4l = lane & (Co4 - 1), upw = 32 >> lpu_shift;
    int koff[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) koff[nt] = (nt * 8 + gid) ^ b_swizzle<NT>(tig);

    const int t_beg = blockIdx.x * tiles_per_cta;
    const int t_end = min(n_tiles, t_beg + tiles_per_cta);
    for (int tile = t_beg; tile < t_end; ++tile) {
        const int y0 = (tile / tiles_x) * TH, x0 = (tile % tiles_x) * IC_TW;
        __syncthreads();
        for (int px = ld_px0; px < (TH + 2) * TWp; px += ld_step) {  // asynchronous: lands during the mask sort
            const int r = px / TWp, c = px - r * TWp;
            const int y = y0 - 1 + r, x = x0 - 1 + c;
            const bool in = y >= 0 && y < H && x >= 0 && x < W;
            const float* src = in ? g_out + ((img + (size_t)y * W + x) * Co + 4 * ld_o4) : g_out;
            cp_async16(sg + px * CoP + 4 * ld_o4, src, in ? 16 : 0);
        }
        for (int i = tid; i < nch * (K + 1); i += IC_MMA_THREADS) sCnt[i] = 0;
        __syncthreads();
        for (int ch = warp; ch < nch; ch += IC_MMA_THREADS / 32) {
            const int i = ch * 32 + lane;
            const int y = y0 + (i >> 5), x = x0 + (i & 31);
            int2 e = make_int2(-2, 0);
            if (y < H && x < W) e = compact_pixel(mask + (img + (size_t)y * W + x) * K, K);
            const unsigned peers = __match_any_sync(0xffffffffu, e.x);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0 && e.x >= -1) sCnt[ch * (K + 1) + (e.x >= 0 ? e.x : K)] = __popc(peers);
            sTmp[i] = make_int2((e.x & 0xff) | (rank << 8), e.y);
        }
        __syncthreads();
        if (tid <= K) {
            int run = 0;
            for (int ch = 0; ch < nch; ++ch) {
                const int c = sCnt[ch * (K + 1) + tid];
                sCnt[ch * (K + 1) + tid] = run;
                run += c;
            }
            sTot[tid] = run;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int k = 0; k < K; ++k) {
                sBase[k] = run;
                run += sTot[k];
            }
            sBase[K] = run;
        }
        __syncthreads();
        for (int i = tid; i < npx; i += IC_MMA_THREADS) {
            const int2 e = sTmp[i];
            const int lab = (int)(signed char)(e.x & 0xff), rank = e.x >> 8, ch = i >> 5;
            if (lab >= 0)  // (offset of the pixel in the g_out tile, value)
                sList[sBase[lab] + sCnt[ch * (K + 1) + lab] + rank] = make_int2(((i >> 5) * TWp + (i & 31)) * CoP, e.y);
            else if (lab == -1)
                sDense[sCnt[ch * (K + 1) + K] + rank] = i;
        }
        cp_async_wait_all();
        __syncthreads();

        // (3) dmask on the tensor cores: warp w < 8 owns m-tiles m = w*MTW + mt: tile row m>>1, x half m&1
        if (warp < 8) {
            float run[MTW][NT][4];   // fp32 running sum, starts from g_extra (its loads fly during the MMA loop)
            float small[MTW][NT][4]; // lo*hi' + hi*lo'
            float pr[MTW][NT][4];    // probabilities of the same elements (0 if unused / outside)
            const int mt0 = warp * MTW;
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                const int m = mt0 + mt;
                const int y = y0 + (m >> 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int x = x0 + (m & 1) * 16 + gid + 8 * h;
                    const bool in = y < H && x < W;
                    const size_t base = (img + (size_t)y * W + x) * K;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int k = nt * 8 + 2 * tig + j;
                            run[mt][nt][2 * h + j] = (in && k < K && g_extra != nullptr) ? __ldg(g_extra + base + k) : 0.f;
                            pr[mt][nt][2 * h + j] = (in && k < K && probs != nullptr) ? __ldg(probs + base + k) : 0.f;
                            small[mt][nt][2 * h + j] = 0.f;
                        }
                }
            }
            for (int t = 0; t < 9; ++t) {
                const int dy = t / 3, dx = t - 3 * dy;
                for (int oc = 0; oc < Co; oc += 32) {  // chains of at most 4 k-steps
                    float main[MTW][NT][4];
#pragma unroll
                    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int j = 0; j < 4; ++j) main[mt][nt][j] = 0.f;
                    const int oe = min(Co, oc + 32);
                    for (int os = oc; os < oe; os += 8) {
                        unsigned bh[NT][2], bl[NT][2];
                        const int row = (t * Co + os + tig) * KP;
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) {
                            bh[nt][0] = __float_as_uint(sGhi[row + koff[nt]]);
                            bh[nt][1] = __float_as_uint(sGhi[row + 4 * KP + koff[nt]]);
                            bl[nt][0] = __float_as_uint(sGlo[row + koff[nt]]);
                            bl[nt][1] = __float_as_uint(sGlo[row + 4 * KP + koff[nt]]);
                        }
#pragma unroll
                        for (int mt = 0; mt < MTW; ++mt) {
                            const int m = mt0 + mt;
                            const int r = m >> 1, xb = (m & 1) * 16;
                            const float* a = sg + ((r + 2 - dy) * TWp + xb + gid + 2 - dx) * CoP + os + tig;
                            unsigned ah[4], al[4];
                            split_tf32(a[0], ah[0], al[0]);
                            split_tf32(a[8 * CoP], ah[1], al[1]);
                            split_tf32(a[4], ah[2], al[2]);
                            split_tf32(a[8 * CoP + 4], ah[3], al[3]);
#pragma unroll
                            for (int nt = 0; nt < NT; ++nt) {
                                mma_tf32(small[mt][nt], al[0], al[1], al[2], al[3], bh[nt][0], bh[nt][1]);
                                mma_tf32(small[mt][nt], ah[0], ah[1], ah[2], ah[3], bl[nt][0], bl[nt][1]);
                                mma_tf32(main[mt][nt], ah[0], ah[1], ah[2], ah[3], bh[nt][0], bh[nt][1]);
                            }
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                            for (int j = 0; j < 4; ++j) run[mt][nt][j] += main[mt][nt][j];
                }
            }
            // epilogue: thread holds, per m-tile and pixel half h (rows gid, gid+8), parts nt*8 + 2*tig + {0,1}
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt) {
                const int m = mt0 + mt;
                const int y = y0 + (m >> 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int x = x0 + (m & 1) * 16 + gid + 8 * h;
                    const bool in = y < H && x < W;
                    const size_t base = (img + (size_t)y * W + x) * K;
                    float d[NT][2], p[NT][2];
                    float dot = 0.f;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            d[nt][j] = run[mt][nt][2 * h + j] + small[mt][nt][2 * h + j];
                            p[nt][j] = pr[mt][nt][2 * h + j];
                            dot = fmaf(d[nt][j], p[nt][j], dot);
                        }
                    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int k = nt * 8 + 2 * tig;
                        float v0 = d[nt][0], v1 = d[nt][1];
                        if (probs != nullptr) {
                            v0 = p[nt][0] * (v0 - dot);
                            v1 = p[nt][1] * (v1 - dot);
                        }
                        if (in) {
                            if ((K & 1) == 0) {
                                if (k < K) __stcs(reinterpret_cast<float2*>(dmask + base + k), make_float2(v0, v1));
                            } else {
                                if (k < K) dmask[base + k] = v0;
                                if (k + 1 < K) dmask[base + k + 1] = v1;
                            }
                        }
                    }
                }
            }
        }

        // (4) dG: unit u = (filter row dy, label k): Co4 lanes walk the label's sorted pixel list once for the three
        // taps of the row (3 x float4 accumulators), 32/Co4 units per warp side by side
        {
            const int nd = sTot[K];
            for (int u0 = warp * upw; u0 < 3 * K; u0 += (IC_MMA_THREADS / 32) * upw) {
                const int u = u0 + usub;
                if (u < 3 * K) {
                    const int dy = u / K, k = u - dy * K;
                    const float* gt = sg + ((2 - dy) * TWp + 2) * CoP + 4 * o4l;  // tap dx reads gt - dx*CoP
                    const int beg = sBase[k], end = sBase[k + 1];
                    float4 a[3];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) a[dx] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = beg; j < end; ++j) {
                        const int2 e = sList[j];
                        const float v = __int_as_float(e.y);
                        const float* gp = gt + e.x;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float4 g = *reinterpret_cast<const float4*>(gp - dx * CoP);
                            a[dx].x = fmaf(v, g.x, a[dx].x); a[dx].y = fmaf(v, g.y, a[dx].y);
                            a[dx].z = fmaf(v, g.z, a[dx].z); a[dx].w = fmaf(v, g.w, a[dx].w);
                        }
                    }
                    for (int jd = 0; jd < nd; ++jd) {  // pixels with several non-zeros (exact ties, soft masks)
                        const int i = sDense[jd];
                        const float v = __ldg(mask + (img + (size_t)(y0 + (i >> 5)) * W + (x0 + (i & 31))) * K + k);
                        const float* gp = gt + ((i >> 5) * TWp + (i & 31)) * CoP;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float4 g = *reinterpret_cast<const float4*>(gp - dx * CoP);
                            a[dx].x = fmaf(v, g.x, a[dx].x); a[dx].y = fmaf(v, g.y, a[dx].y);
                            a[dx].z = fmaf(v, g.z, a[dx].z); a[dx].w = fmaf(v, g.w, a[dx].w);
                        }
                    }
                    if (beg != end || nd != 0) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            float4* dst = reinterpret_cast<float4*>(sdG + ((3 * dy + dx) * K + k) * Co + 4 * o4l);
                            float4 c = *dst;
                            c.x += a[dx].x; c.y += a[dx].y; c.z += a[dx].z; c.w += a[dx].w;
                            *dst = c;
                        }
                    }
                }
            }
        }
        // db: column sums of the tile interior (zero outside the image): warp w sums rows w, w+9, ...
        for (int oc = 0; oc < Co; oc += 32) {
            const int o = oc + lane;
            if (o < Co) {
                float s = 0.f;
                for (int ty = warp; ty < TH; ty += IC_MMA_THREADS / 32) {
                    const float* row = sg + ((ty + 1) * TWp + 1) * CoP + o;
#pragma unroll 8
                    for (int tx = 0; tx < IC_TW; ++tx) s += row[tx * CoP];
                }
                sdb[warp * Co + o] += s;
            }
        }
    }
    __syncthreads();
    const size_t slot_ws = (size_t)b * gridDim.x + blockIdx.x;
    for (int i = tid; i < 9 * K * Co; i += IC_MMA_THREADS) ws_dG[slot_ws * 9 * K * Co + i] = sdG[i];
    for (int o = tid; o < Co; o += IC_MMA_THREADS) {
        float s = 0.f;
        for (int w = 0; w < IC_MMA_THREADS / 32; ++w) s += sdb[w * Co + o];
        ws_db[slot_ws * Co + o] = s;
    }
}

// dG[b,i] = sum over the sample's splits (ascending); db[o] = sum over all (b, split) slots (ascending)
__global__ void inject_conv_bwd_finalize_kernel(const float* __restrict__ ws_dG, const float* __restrict__ ws_db,
                                                float* __restrict__ dG, float* __restrict__ db, int B, int splits,
                                                int n_per, int Co) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)B * n_per;
    if (i < total) {
        const size_t b = i / n_per, r = i - b * n_per;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += ws_dG[(b * splits + sp) * n_per + r];
        dG[i] = s;
    }
    if (db != nullptr && blockIdx.x == 0) {  // groups of Co threads sum interleaved slots, then the groups in order
        __shared__ float part[256];
        const int ngrp = Co <= 256 ? 256 / Co : 0;
        if (ngrp >= 1) {
            const int g = threadIdx.x / Co, o = threadIdx.x - g * Co;
            if (g < ngrp) {
                float s = 0.f;
                for (int sl = g; sl < B * splits; sl += ngrp) s += ws_db[(size_t)sl * Co + o];
                part[g * Co + o] = s;
            }
            __syncthreads();
            if (threadIdx.x < Co) {
                float s = 0.f;
                for (int gg = 0; gg < ngrp; ++gg) s += part[gg * Co + threadIdx.x];
                db[threadIdx.x] = s;
            }
        }
    }
}

int check_dims(const char* what, int B, int H, int W, int K, int Co) {
    UPS_REQUIRE(B >= 0 && H > 0 && W > 0, "%s: bad sizes B=%d H=%d W=%d", what, B, H, W);
    UPS_REQUIRE(K >= 1 && K <= 32, "%s: K=%d outside [1,32]", what, K);
    UPS_REQUIRE(Co >= 4 && Co <= 128 && Co % 4 == 0, "%s: Co=%d must be a multiple of 4 in [4,128]", what, Co);
    UPS_REQUIRE((long long)B * H * W < (1ll << 31), "%s: B*H*W >= 2^31", what);
    return UPS_OK;
}

int fwd_splits(int B, int H) {  // ~8 CTAs of 256 threads per SM
    const int n_strips = (int)cdiv(H, IC_FWD_ROWS);
    long long want = cdiv(8ll * NUM_SMS, B > 0 ? B : 1);
    if (want < 1) want = 1;
    if (want > n_strips) want = n_strips;
    return (int)want;
}

bool bwd_use_mma(int Co) { return Co >= 8 && (Co & (Co - 1)) == 0; }
int bwd_kp(int K) { return K <= 8 ? 8 : K <= 16 ? 16 : K <= 24 ? 24 : 32; }
int bwd_smem_bytes(bool mma, int KP, int K, int Co, int th) {
    return mma ? bwd_mma_smem_layout(KP, K, Co, th).total : bwd_smem_layout(KP, K, Co, th).total;
}
int bwd_tile_rows(bool mma, int KP, int K, int Co) {
    if (mma) {  // two CTAs per SM: 8-row tiles (two m-tiles per warp) up to K = 16, 4-row tiles beyond (registers)
        const int th = KP <= 16 ? 8 : 4;
        if (bwd_smem_bytes(true, KP, K, Co, th) <= IC_SMEM_HALF) return th;
        if (bwd_smem_bytes(true, KP, K, Co, 4) <= IC_SMEM_BUDGET) return 4;
        return 0;
    }
    for (int th = 16; th >= 4; th >>= 1)
        if (bwd_smem_bytes(false, KP, K, Co, th) <= IC_SMEM_BUDGET) return th;
    return 0;
}

struct BwdPlan {
    bool mma;
    int KP, TH, tiles_x, tiles_y, n_tiles, splits, tiles_per_cta, smem;
};
BwdPlan bwd_plan(int B, int H, int W, int K, int Co) {
    BwdPlan p;
    p.KP = bwd_kp(K);
    p.mma = bwd_use_mma(Co);
    p.TH = bwd_tile_rows(p.mma, p.KP, K, Co);
    if (p.TH == 0 && p.mma) {  // the pre-split table does not fit: CUDA-core variant
        p.mma = false;
        p.TH = bwd_tile_rows(false, p.KP, K, Co);
    }
    if (p.TH == 0) { p.splits = 0; return p; }
    p.smem = bwd_smem_bytes(p.mma, p.KP, K, Co, p.TH);
    p.tiles_x = (int)cdiv(W, IC_TW);
    p.tiles_y = (int)cdiv(H, p.TH);
    p.n_tiles = p.tiles_x * p.tiles_y;
    long long want = cdiv(16ll * NUM_SMS, B > 0 ? B : 1);  // one or two CTAs per SM resident; ~8 waves: a short tail
    if (want < 1) want = 1;
    if (want > p.n_tiles) want = p.n_tiles;
    p.tiles_per_cta = (int)cdiv(p.n_tiles, want);
    p.splits = (int)cdiv(p.n_tiles, p.tiles_per_cta);
    return p;
}

template <int KP>
int launch_bwd(const BwdPlan& p, const float* g_out, const float* mask, const float* G, const float* probs,
               const float* g_extra, float* dmask, float* ws_dG, float* ws_db, int B, int H, int W, int K, int Co,
               cudaStream_t st) {
    if (p.mma) {
#define UPS_IC_MMA(MTW, CO)                                                                                      \
    do {                                                                                                         \
        UPS_CUDA(cudaFuncSetAttribute(inject_conv_bwd_mma_kernel<KP / 8, MTW, CO>,                               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));                     \
        inject_conv_bwd_mma_kernel<KP / 8, MTW, CO><<<dim3(p.splits, B), IC_MMA_THREADS, p.smem, st>>>(          \
            g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, H, W, K, Co, p.tiles_x, p.n_tiles, p.tiles_per_cta); \
    } while (0)
        if (p.TH == 8 && Co == 32) UPS_IC_MMA(2, 32);  // the reference's first-layer width (final_hour.config[0])
        else if (p.TH == 8) UPS_IC_MMA(2, 0);
        else UPS_IC_MMA(1, 0);
#undef UPS_IC_MMA
        return after_launch("inject_conv_bwd_mma_kernel");
    }
    UPS_CUDA(cudaFuncSetAttribute(inject_conv_bwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));
    inject_conv_bwd_kernel<KP><<<dim3(p.splits, B), IC_BWD_THREADS, p.smem, st>>>(
        g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, H, W, K, Co, p.TH, p.tiles_x, p.n_tiles, p.tiles_per_cta);
    return after_launch("inject_conv_bwd_kernel");
}

}  // namespace
}  // namespace ups

using namespace ups;

// host-only: the tiling ups_inject_conv_bwd will use (tests check it against the shared-memory limits without a GPU)
extern "C" int ups_inject_conv_bwd_plan(int B, int H, int W, int K, int Co, int* out6) {
    UPS_REQUIRE(out6 != nullptr, "inject_conv_bwd_plan: null pointer");
    if (int rc = check_dims("inject_conv_bwd_plan", B, H, W, K, Co)) return rc;
    const BwdPlan p = bwd_plan(B, H, W, K, Co);
    out6[0] = p.splits > 0 ? (p.mma ? 2 : 1) : 0;  // 0 = does not fit, 1 = CUDA cores, 2 = mma.sync
    out6[1] = p.splits > 0 ? p.TH : 0;
    out6[2] = p.splits > 0 ? p.splits : 0;
    out6[3] = p.splits > 0 ? p.tiles_per_cta : 0;
    out6[4] = p.splits > 0 ? p.smem : 0;
    out6[5] = p.splits > 0 ? p.n_tiles : 0;
    return UPS_OK;
}

extern "C" size_t ups_inject_conv_workspace_bytes(int B, int H, int W, int K, int Co) {
    if (B <= 0 || H <= 0 || W <= 0 || K < 1 || K > 32 || Co < 4 || Co > 128 || (Co & 3)) return 256;
    const BwdPlan p = bwd_plan(B, H, W, K, Co);
    if (p.splits == 0) return 256;
    return (size_t)B * p.splits * (9 * K * Co + Co) * sizeof(float) + 256;
}

extern "C" int ups_inject_conv_table_fwd(const float* feat, const float* V, float* G, int B, int K, int F, int Co,
                                         void* stream) {
    UPS_REQUIRE(feat && V && G, "inject_conv_table_fwd: null pointer");
    UPS_REQUIRE(B >= 0 && K >= 1 && K <= 32 && F >= 1 && Co >= 1, "inject_conv_table_fwd: bad sizes");
    UPS_REQUIRE((size_t)K * F * sizeof(float) <= 48 * 1024, "inject_conv_table_fwd: K*F=%d too large", K * F);
    if (B == 0) return UPS_OK;
    inject_conv_table_fwd_kernel<<<dim3(9, B), 256, K * F * sizeof(float), as_stream(stream)>>>(feat, V, G, K, F, Co);
    return after_launch("inject_conv_table_fwd_kernel");
}

extern "C" int ups_inject_conv_table_bwd(const float* dG, const float* feat, const float* V, float* dfeat, float* dV,
                                         int B, int K, int F, int Co, void* stream) {
    UPS_REQUIRE(dG && feat && V, "inject_conv_table_bwd: null pointer");
    UPS_REQUIRE(B >= 0 && K >= 1 && K <= 32 && F >= 1 && Co >= 4 && Co % 4 == 0, "inject_conv_table_bwd: bad sizes");
    UPS_REQUIRE(aligned16(dG) && aligned16(V) && aligned16(feat), "inject_conv_table_bwd: dG, V and feat must be 16-byte aligned");
    const size_t smem = (size_t)9 * K * Co * sizeof(float);
    UPS_REQUIRE(smem <= 200 * 1024, "inject_conv_table_bwd: 9*K*Co=%d too large", 9 * K * Co);
    cudaStream_t st = as_stream(stream);
    if (dfeat != nullptr && B > 0) {
        UPS_CUDA(cudaFuncSetAttribute(inject_conv_table_bwd_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        const int parts = (int)cdiv((long long)K * F, 256) < 4 ? (int)cdiv((long long)K * F, 256) : 4;
        inject_conv_table_bwd_feat_kernel<<<dim3(B, parts), 256, smem, st>>>(dG, V, dfeat, K, F, Co);
        if (int rc = after_launch("inject_conv_table_bwd_feat_kernel")) return rc;
    }
    if (dV != nullptr) {
        inject_conv_table_bwd_filter_kernel<<<dim3((unsigned)cdiv(F + K, 4), 9), 256, 32 * Co * sizeof(float), st>>>(dG, feat, dV, B,
                                                                                                                        K, F, Co);
        if (int rc = after_launch("inject_conv_table_bwd_filter_kernel")) return rc;
    }
    return UPS_OK;
}

extern "C" int ups_inject_conv_fwd(const float* mask, const float* G, const float* bias, float* out, int B, int H, int W,
                                   int K, int Co, void* stream) {
    UPS_REQUIRE(mask && G && bias && out, "inject_conv_fwd: null pointer");
    if (int rc = check_dims("inject_conv_fwd", B, H, W, K, Co)) return rc;
    UPS_REQUIRE(aligned16(mask) && aligned16(G) && aligned16(out), "inject_conv_fwd: mask, G and out must be 16-byte aligned");
    if (B == 0) return UPS_OK;
    const size_t smem = (size_t)(9 * K * Co + Co) * sizeof(float) + (size_t)(IC_FWD_ROWS + 2) * (W + 2) * sizeof(int2);
    UPS_REQUIRE(smem <= (size_t)IC_SMEM_BUDGET, "inject_conv_fwd: K=%d Co=%d W=%d needs %zu bytes of shared memory", K, Co,
                W, smem);
    const int splits = fwd_splits(B, H);
    const int n_strips = (int)cdiv(H, IC_FWD_ROWS);
    const int spc = (int)cdiv(n_strips, splits);
    const dim3 grid((unsigned)cdiv(n_strips, spc), B);
    cudaStream_t st = as_stream(stream);
    UPS_CUDA(cudaFuncSetAttribute(inject_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    inject_conv_fwd_kernel<<<grid, IC_FWD_THREADS, smem, st>>>(mask, G, bias, out, H, W, K, Co, spc);
    return after_launch("inject_conv_fwd_kernel");
}

extern "C" int ups_inject_conv_bwd(const float* g_out, const float* mask, const float* G, const float* probs,
                                   const float* g_extra, float* dmask, float* dG, float* db, int B, int H, int W, int K,
                                   int Co, void* ws, size_t ws_bytes, void* stream) {
    UPS_REQUIRE(g_out && mask && G && dmask && dG, "inject_conv_bwd: null pointer");
    if (int rc = check_dims("inject_conv_bwd", B, H, W, K, Co)) return rc;
    UPS_REQUIRE(aligned16(g_out) && aligned16(mask) && aligned16(dmask), "inject_conv_bwd: buffers must be 16-byte aligned");
    if (B == 0) return UPS_OK;
    const BwdPlan p = bwd_plan(B, H, W, K, Co);
    UPS_REQUIRE(p.splits > 0, "inject_conv_bwd: K=%d Co=%d does not fit shared memory", K, Co);
    const size_t need = ups_inject_conv_workspace_bytes(B, H, W, K, Co);
    UPS_REQUIRE(ws != nullptr && ws_bytes >= need, "inject_conv_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    float* ws_dG = reinterpret_cast<float*>(ws);
    float* ws_db = ws_dG + (size_t)B * p.splits * 9 * K * Co;
    cudaStream_t st = as_stream(stream);
    int rc;
    switch (p.KP) {
        case 8: rc = launch_bwd<8>(p, g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, B, H, W, K, Co, st); break;
        case 16: rc = launch_bwd<16>(p, g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, B, H, W, K, Co, st); break;
        case 24: rc = launch_bwd<24>(p, g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, B, H, W, K, Co, st); break;
        default: rc = launch_bwd<32>(p, g_out, mask, G, probs, g_extra, dmask, ws_dG, ws_db, B, H, W, K, Co, st); break;
    }
    if (rc) return rc;
    const int n_per = 9 * K * Co;
    inject_conv_bwd_finalize_kernel<<<(unsigned)cdiv((long long)B * n_per, 256), 256, 0, st>>>(ws_dG, ws_db, dG, db, B,
                                                                                                 p.splits, n_per, Co);
    return after_launch("inject_conv_bwd_finalize_kernel");
}
