// SURVEY.md 8f N4, encoder side: the dense part of the backward of the first convolution on the part images, on
// tcgen05 (replaces parts_conv_bwd_data_kernel when Co == 32 and W is 128 or 256).
//
//   R_k[q,c]   = sum_{t,o} g_h[k*B+b, q - off_t, o] V[t,c,o]     cotangent of the part image of plane (k,b)
//   dmask[q,k] = sum_c img[q,c] R_k[q,c]                         (autodiff of model.py:176-187,478 + nn.py:617-664)
//
// Per pixel of a plane, D[p][tc] = sum_o g_h[p,o] V[tc,o] (tc = 3*tap + channel, 27 of 32 columns used) is a plain
// GEMM row: M = 128 pixels (one tile = 128 consecutive pixels of an image row), N = 32, reduction = Co = 32.  A pixel's
// 32 floats are exactly one 128-byte SWIZZLE_128B row, so the tile is TMA-loadable as it lies in HBM and is the UMMA
// A operand without any re-layout.  3xTF32 as in K4: the raw fp32 tile is the hi operand (the tensor core reads the top
// 19 bits), lo = g - trunc(g) goes to a second buffer;  D = G_hi.[V_hi | V_lo]^T (N = 64) and D[:, 0:32] += G_lo.V_hi^T,
// fp32 accumulators double-buffered in TMEM.  The 9-tap shifted sum R[q,c] = sum_t D[q - off_t][t][c] and the dot with
// the image run in the TMEM-drain epilogue through a 4-row shared-memory ring, as in the CUDA-core kernel.
//
// One persistent CTA per SM walks whole planes (rows top to bottom).  Warp roles:
//   warps 0-3   drain:  tcgen05.ld (thread = pixel) -> ring row y (27 floats per pixel)
//   warps 4-7   reduce: R and dmask of row y-1 from ring rows y-2, y-1, y -> dm_planes.  Drain and reduce are
//               two stages of a pipeline over the ring rows (mbarriers rowfull / rowfree per ring slot): as one
//               warp group they ran one warp per scheduler, latency-bound at 2150 cycles per tile (measured 4.0 ms)
//   warps 8-11  splitter: lo tile, and the column sums of g_h (db)
//   warp 12     TMA producer (one lane): 128 x 32 fp32 boxes into a 4-stage ring
//   warp 13     MMA issuer (one lane)
// Every cross-role hand-off is an mbarrier with a bounded wait (a protocol bug traps instead of hanging the GPU).
#include <stdlib.h>

#include "common.cuh"
#include "tc_helpers.cuh"

namespace ups {
namespace pctc {

using namespace tma;

constexpr int TILE = 128;   // pixels per tile = UMMA M
constexpr int CO = 32;      // reduction length = one SW128 row
constexpr int NTC = 32;     // (tap, channel) columns: 27 used
constexpr int MAX_NST = 4;  // TMA stages of 16 KB (8 stages measured slower: 4.77 vs 4.36 ms for the whole call, profiles/r02_tuning.md)
constexpr int G_BYTES = TILE * 128;          // 16 KB per tile
constexpr int B_BYTES = 2 * NTC * 128;       // [V_hi | V_lo] rows
constexpr int TPB = 576;
constexpr int W_RED = 8, W_SPLIT = 12, W_TMA = 16, W_MMA = 17;
constexpr uint32_t TMEM_COLS = 64;           // 2 accumulator buffers of 32 columns
constexpr unsigned FULLM = 0xffffffffu;

struct Layout {
    int stage0, lo0, b0, ring, red, bar, total;
};
__host__ __device__ inline Layout layout(int W, int NST) {
    Layout L;
    int o = 0;
    L.stage0 = o; o += NST * G_BYTES;
    L.lo0 = o;    o += 2 * G_BYTES;
    L.b0 = o;     o += B_BYTES;
    L.ring = o;   o += 4 * (W + 2) * 27 * 4;
    o = (o + 15) & ~15;
    L.red = o;    o += 128 * 16;
    L.bar = o;    o += 512;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(TPB, 1) parts_conv_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_g,
                                                                 const float* __restrict__ img,
                                                                 const float* __restrict__ V,
                                                                 float* __restrict__ dm_planes,
                                                                 float* __restrict__ ws_db, int B, int H, int W, int K,
                                                                 int n_planes, int NST) {
    constexpr uint32_t IDESC_32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sp = smem_raw + (sb - smem_u32(smem_raw));
    const Layout L = layout(W, NST);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = W / TILE, Wr = W + 2, P = H * W;
    float* sD = reinterpret_cast<float*>(sp + L.ring);     // ring [4][Wr][27]; columns 0 and Wr-1 stay zero
    float* sRed = reinterpret_cast<float*>(sp + L.red);    // [128 splitter threads][4]
    const uint32_t bar_full = sb + L.bar;                  // [MAX_NST]
    const uint32_t bar_empty = bar_full + 8 * MAX_NST;     // [MAX_NST]
    const uint32_t bar_lo_ready = bar_empty + 8 * MAX_NST; // [2]
    const uint32_t bar_lo_free = bar_lo_ready + 16;        // [2]
    const uint32_t bar_tm_full = bar_lo_free + 16;         // [2]
    const uint32_t bar_tm_empty = bar_tm_full + 16;        // [2]
    const uint32_t bar_rowfull = bar_tm_empty + 16;        // [4] ring slots: drain -> reduce
    const uint32_t bar_rowfree = bar_rowfull + 32;         // [4] reduce -> drain
    const uint32_t slot = bar_rowfree + 32;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_lo_ready + 8 * i, 128);
            mbar_init(bar_lo_free + 8 * i, 1);
            mbar_init(bar_tm_full + 8 * i, 1);
            mbar_init(bar_tm_empty + 8 * i, 256);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(bar_rowfull + 8 * i, 256); mbar_init(bar_rowfree + 8 * i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid == W_TMA * 32) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmap_g)) : "memory");
    // B operand, once per CTA: rows 0..31 = V_hi[tc][o], rows 32..63 = V_lo[tc][o] (rows 27..31 of each half zero),
    // K-major SWIZZLE_128B: 16-byte chunk cc of row r at r*128 + ((cc ^ (r & 7)) * 16)
    for (int i = tid; i < NTC * 8; i += TPB) {
        const int tc = i >> 3, cc = i & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tc < 27) v = ld4(V + tc * CO + 4 * cc);
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        sts4(sb + L.b0 + tc * 128 + ((cc ^ (tc & 7)) * 16), hi);
        sts4(sb + L.b0 + (NTC + tc) * 128 + ((cc ^ ((NTC + tc) & 7)) * 16), lo);
    }
    for (int i = tid; i < 4 * Wr * 27; i += TPB) sD[i] = 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(tmem_base) : "r"(slot));

    if (warp == W_TMA) {
        // ================================================================= TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int n = blockIdx.x; n < n_planes; n += gridDim.x)
                for (int y = 0; y < H; ++y)
                    for (int t = 0; t < NT; ++t, ++it) {
                        const uint32_t s = it % NST, ph = it / NST;
                        mbar_wait(bar_empty + 8 * s, (ph & 1) ^ 1);
                        mbar_expect_tx(bar_full + 8 * s, G_BYTES);
                        tma_load_2d(sb + L.stage0 + s * G_BYTES, &tmap_g, 0, n * P + y * W + t * TILE, bar_full + 8 * s);
                    }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            uint32_t it = 0;
            for (int n = blockIdx.x; n < n_planes; n += gridDim.x)
                for (int yt = 0; yt < H * NT; ++yt, ++it) {
                    const uint32_t s = it % NST, ph = it / NST, j = it & 1, u = (it >> 1) & 1;
                    mbar_wait(bar_full + 8 * s, ph & 1);
                    mbar_wait(bar_lo_ready + 8 * j, u);
                    mbar_wait(bar_tm_empty + 8 * j, u ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + j * 32;
                    const uint32_t a_hi = sb + L.stage0 + s * G_BYTES, a_lo = sb + L.lo0 + j * G_BYTES, bb = sb + L.b0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {   // the three 3xTF32 terms accumulate into the same 32 columns
                        const uint64_t b_hi = sw128_desc(bb + ks * 32), b_lo = sw128_desc(bb + NTC * 128 + ks * 32);
                        const uint64_t ah = sw128_desc(a_hi + ks * 32);
                        umma_tf32(d, ah, b_hi, IDESC_32, ks > 0 ? 1u : 0u);                    // hi.hi
                        umma_tf32(d, ah, b_lo, IDESC_32, 1u);                                  // hi.lo
                        umma_tf32(d, sw128_desc(a_lo + ks * 32), b_hi, IDESC_32, 1u);          // lo.hi
                    }
                    umma_commit(bar_empty + 8 * s);
                    umma_commit(bar_lo_free + 8 * j);
                    umma_commit(bar_tm_full + 8 * j);
                }
        }
        __syncwarp();
    } else if (warp >= W_SPLIT) {
        // ================================================================= splitter (+ db)
        const int st = tid - W_SPLIT * 32;        // 0..127
        const int cc = st & 7, rsub = st >> 3;    // 16-byte chunk of the row, row within a pass of 16 rows
        uint32_t it = 0;
        for (int n = blockIdx.x; n < n_planes; n += gridDim.x) {
            float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int yt = 0; yt < H * NT; ++yt, ++it) {
                const uint32_t s = it % NST, ph = it / NST, j = it & 1, u = (it >> 1) & 1;
                mbar_wait(bar_lo_free + 8 * j, u ^ 1);
                mbar_wait(bar_full + 8 * s, ph & 1);
                const uint32_t hi_base = sb + L.stage0 + s * G_BYTES, lo_base = sb + L.lo0 + j * G_BYTES;
                float4 g[TILE / 16];
                uint32_t off[TILE / 16];
#pragma unroll
                for (int p = 0; p < TILE / 16; ++p) {      // all loads first (the asm statements keep program order)
                    const int r = p * 16 + rsub;
                    off[p] = r * 128 + ((cc ^ (r & 7)) * 16);
                    g[p] = lds4(hi_base + off[p]);
                }
#pragma unroll
                for (int p = 0; p < TILE / 16; ++p) {
                    sts4(lo_base + off[p], make_float4(g[p].x - tf32_hi(g[p].x), g[p].y - tf32_hi(g[p].y),
                                                        g[p].z - tf32_hi(g[p].z), g[p].w - tf32_hi(g[p].w)));
                    dbacc.x += g[p].x; dbacc.y += g[p].y; dbacc.z += g[p].z; dbacc.w += g[p].w;
                }
                fence_proxy_async();
                mbar_arrive(bar_lo_ready + 8 * j);
            }
            // db of this plane: fixed-order sum over the 16 threads that own the same channels
            reinterpret_cast<float4*>(sRed)[st] = dbacc;
            asm volatile("bar.sync 2, 128;\n" ::: "memory");
            if (st < CO) {
                const int ch = st >> 2, e = st & 3;
                float v = 0.f;
                for (int r = 0; r < 16; ++r) v += sRed[(r * 8 + ch) * 4 + e];
                ws_db[(size_t)n * CO + st] = v;
            }
            asm volatile("bar.sync 2, 128;\n" ::: "memory");
        }
    } else if (warp < W_RED) {
        // ================================================================= drain (warps 0-7): thread = pixel of the tile,
        // warp w reads TMEM lane quadrant w & 3 and the column half w >> 2 (columns 0-15 or 16-31 of the 27 used)
        const int px = (warp & 3) * 32 + lane, half = warp >> 2;
        // Ring rows are numbered globally: plane i writes rows base-1 (zeros), base .. base+H-1 (D), base+H (zeros) with
        // base = i*(H+2) + 1; global row r lives in slot r & 3 and is the (r >> 2)-th occupant of that slot.
        uint32_t it = 0, base = 1;
        auto claim = [&](uint32_t r) { mbar_wait(bar_rowfree + 8 * (r & 3), ((r >> 2) & 1) ^ 1); };
        auto publish = [&](uint32_t r) { mbar_arrive(bar_rowfull + 8 * (r & 3)); };
        for (int n = blockIdx.x; n < n_planes; n += gridDim.x, base += H + 2) {
            for (int y = -1; y <= H; ++y) {
                const uint32_t r = base + y;
                claim(r);
                float* rowp = sD + ((r & 3) * Wr + 1 + px) * 27 + 16 * half;
                if (y >= 0 && y < H) {
                    for (int t = 0; t < NT; ++t, ++it) {
                        const uint32_t j = it & 1, u = (it >> 1) & 1;
                        mbar_wait(bar_tm_full + 8 * j, u);
                        tc_fence_after();
                        float d0[16];
                        tmem_ld16(tmem_base + j * 32 + 16 * half + ((uint32_t)((warp & 3) * 32) << 16), d0);
                        tmem_ld_wait();
                        tc_fence_before();
                        mbar_arrive(bar_tm_empty + 8 * j);
                        float* dst = rowp + t * TILE * 27;
#pragma unroll
                        for (int tc = 0; tc < 16; ++tc)
                            if (16 * half + tc < 27) dst[tc] = d0[tc];
                    }
                } else {
                    for (int t = 0; t < NT; ++t) {   // the rows above and below the image contribute nothing
                        float* dst = rowp + t * TILE * 27;
#pragma unroll
                        for (int tc = 0; tc < 16; ++tc)
                            if (16 * half + tc < 27) dst[tc] = 0.f;
                    }
                }
                publish(r);
            }
        }
    } else if (warp < W_SPLIT) {
        // ================================================================= reduce (warps 8-11): thread = column of the row
        const int ct = tid - W_RED * 32;
        uint32_t base = 1;
        for (int n = blockIdx.x; n < n_planes; n += gridDim.x, base += H + 2) {
            const int k = n / B, b = n - k * B;
            const float* ib = img + (size_t)b * P * 3;
            float* dm = dm_planes + ((size_t)b * K + k) * P;
            // image pixels of the row, fetched one row ahead and kept in registers (static indices: a dynamically
            // indexed array lands in local memory and the reduce warps then wait a DRAM latency per row)
            float im[2][3], imn[2][3];
            auto fetch = [&](int y, float (&dst)[2][3]) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const bool ok = t < NT && y < H;
                    const size_t q = ok ? (size_t)y * W + t * TILE + ct : 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) dst[t][c] = ok ? __ldg(ib + q * 3 + c) : 0.f;
                }
            };
            fetch(0, imn);
            for (int yy = 0; yy < H; ++yy) {
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < 3; ++c) im[t][c] = imn[t][c];
                fetch(yy + 1, imn);
                // R of row yy needs D rows yy+1 (tap dy = 0), yy, yy-1 (dy = 2); rows are published in order
                const uint32_t r1 = base + yy + 1;
                mbar_wait(bar_rowfull + 8 * (r1 & 3), (r1 >> 2) & 1);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (t >= NT) continue;
                    const int x = t * TILE + ct;
                    float r0 = 0.f, rr1 = 0.f, r2 = 0.f;
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t srow = (r1 - dy) & 3;
                        const float* row = sD + (srow * Wr + x + 2) * 27 + 9 * dy;
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const float* dd = row + 3 * dx - dx * 27;   // column x + 1 - dx (+1 for the zero border)
                            r0 += dd[0]; rr1 += dd[1]; r2 += dd[2];
                        }
                    }
                    dm[(size_t)yy * W + x] = fmaf(im[t][2], r2, fmaf(im[t][1], rr1, im[t][0] * r0));
                }
                // row yy-1 is not needed any more; after the last row of the plane neither are rows yy and yy+1
                mbar_arrive(bar_rowfree + 8 * ((r1 - 2) & 3));
                if (yy == H - 1) { mbar_arrive(bar_rowfree + 8 * ((r1 - 1) & 3)); mbar_arrive(bar_rowfree + 8 * (r1 & 3)); }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace pctc

// 1 if the tcgen05 kernel handles this shape (and is not disabled by UPS_PCB_TC=0)
bool parts_conv_bwd_tc_ok(int B, int H, int W, int K, int Co) {
    static const bool enabled = []() { const char* e = getenv("UPS_PCB_TC"); return !(e && atoi(e) == 0); }();
    return enabled && Co == pctc::CO && (W == 128 || W == 256) && H >= 1 && (long long)K * B * H * W < (1ll << 31);
}

// the dense part of ups_parts_conv_bwd: g_h -> dm_planes [B][K][P] and ws_db [K*B][Co]
int parts_conv_bwd_tc_launch(const float* g_h, const float* img, const float* V, float* dm_planes, float* ws_db, int B,
                             int H, int W, int K, cudaStream_t st) {
    pctc::EncodeTiledFn enc = pctc::encode_tiled_fn();
    UPS_REQUIRE(enc != nullptr, "parts_conv_bwd: cuTensorMapEncodeTiled not available from the driver");
    const long long rows = (long long)K * B * H * W;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)pctc::CO, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)pctc::CO * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)pctc::CO, (cuuint32_t)pctc::TILE};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g_h), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("parts_conv_bwd: cuTensorMapEncodeTiled failed (%d)", (int)cr); return UPS_E_CUDA; }
    int nst = pctc::MAX_NST;
    while (nst > 2 && (size_t)pctc::layout(W, nst).total + 1024 > 227 * 1024) --nst;
    const pctc::Layout L = pctc::layout(W, nst);
    const size_t sm = (size_t)L.total + 1024;
    UPS_REQUIRE(sm <= 227 * 1024, "parts_conv_bwd: W=%d needs %zu bytes of shared memory", W, sm);
    static const cudaError_t attr = cudaFuncSetAttribute(pctc::parts_conv_bwd_tc_kernel,
                                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    UPS_CUDA(attr);
    const int n_planes = K * B;
    const int grid = n_planes < NUM_SMS ? n_planes : NUM_SMS;
    pctc::parts_conv_bwd_tc_kernel<<<grid, pctc::TPB, sm, st>>>(tmap, img, V, dm_planes, ws_db, B, H, W, K, n_planes, nst);
    return after_launch("parts_conv_bwd_tc_kernel");
}

}  // namespace ups
