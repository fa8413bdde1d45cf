// K4 on the tensor cores, for K in {16, 32}, F = 64, P % 128 == 0.
//
//   GEMM1  dmask[64 px, K] = G[64 px, 64 f] . feat^T[64 f, K]       tcgen05.mma kind::tf32, M=64, N=K,
//                                                                   8 k-steps, fp32 accumulator in TMEM
//   GEMM2  dfeat[K, 64 f] += mh^T[K, 64 px] . G[64 px, 64 f]        warp-level mma.sync m16n8k8 tf32,
//                                                                   fp32 accumulators in registers
// GEMM2 contracts over the pixel axis, along which neither operand is contiguous in memory ("TN"
// weight-gradient shape).  tcgen05 accepts MN-major tf32 operands only in the 128B_BASE32B swizzled
// layout (a SWIZZLE_NONE MN-major descriptor is silently a no-op on B200: measured during bring-up,
// scripts/debug_tc.py), which would need a second 64 KB copy of the tile and halve the occupancy;
// mma.sync reads its fragments from the tile GEMM1 already uses.  GEMM2 is 1/3 of the kernel's flops.
//
// Why tensor cores here and nowhere else on the path: on CUDA cores this kernel needs 2*K*F = 2048
// FMAs per pixel for GEMM1 alone and every FMA operand comes through shared memory; ncu on the SIMT
// kernel (profiles/r01_ncu_v1_summary.md) shows 54 % shared-pipe and 68 % issue utilisation at 31 %
// DRAM, i.e. not memory-bound.  All other kernels of the path have < 100 flop per pixel.
//
// Precision: fp32 inputs are split a = hi + lo with hi = rna_tf32(a), lo = rna_tf32(a - hi) and the
// product is hi*hi' + lo*hi' + hi*lo' (3xTF32; the dropped lo*lo' term and the rounding of lo are
// <= 2^-22 relative), accumulated in fp32: same error class as an fp32 FMA chain of length 64.
//
// Shared-memory operand layout (no TMA descriptors needed): the canonical SWIZZLE_NONE "interleave"
// layout of 8x16-byte core matrices, written directly by 16-byte cp.async:
//   16-byte chunk j (4 floats along f) of pixel row r  ->  j*LBO + (r/8)*128 + (r%8)*16
// Read K-major (rows = pixels, K = f): LBO (next 4 f), SBO = 128 (next 8 pixels) -> GEMM1 A;
// the same bytes are read as mma.sync B fragments by GEMM2, so ONE copy of the g_inj tile feeds both.
#include "common.cuh"

namespace ups {
namespace tc {

constexpr int TILE = 64;       // pixels per tile = UMMA M of GEMM1
constexpr int F = 64;
constexpr int TPB = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
        if (it > (1u << 26)) __trap();
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
    // base_offset 0, lbo_mode 0, layout_type SWIZZLE_NONE (0) [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t lds1(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void mma_tf32(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int K>
struct Smem {
    static constexpr int NCH = K / 4;                 // 16-byte chunks per [.,K] row
    // G tile in the SWIZZLE_NONE core-matrix layout, chunk-block major with a 64-byte pad per block:
    //   16-byte chunk j (4 floats along f) of pixel row r -> j*A_LBO + (r/8)*128 + (r%8)*16
    // UMMA K-major descriptor: LBO = A_LBO (next 4 f), SBO = 128 (next 8 pixels).  The pad makes the
    // mma.sync fragment reads of GEMM2 (8 features x 4 pixels per load) hit 32 distinct banks.
    static constexpr int A_LBO = TILE * 16 + 64;      // 1088
    static constexpr int A_BYTES = 16 * A_LBO;        // 17 KB
    static constexpr int ROWT = TILE * K * 4;         // one [TILE][K] tile
    static constexpr int B_BYTES = K * F * 4;
    static constexpr int MHS = TILE + 4;              // row stride (floats) of the transposed hard-mask tile
    static constexpr int MH_BYTES = K * MHS * 4;
    static constexpr int A_HI = 0, A_LO = A_BYTES, PT = 2 * A_BYTES, GT = PT + ROWT, MH = GT + ROWT,
                         B_HI = MH + MH_BYTES, B_LO = B_HI + B_BYTES, BAR = B_LO + B_BYTES, TOTAL = BAR + 32;
};

// 128 threads, one 64-pixel tile at a time, 53 KB of shared memory (K=16): four CTAs per SM keep
// four tiles (4 x 26 KB of loads) in flight, which is what hides the HBM latency here - the phases
// of one tile (load, split, GEMM1, epilogue, GEMM2, store) are serial within a CTA.
template <int K>
__global__ void __launch_bounds__(TPB, (K == 16) ? 4 : 3) step_decode_bwd_tc_kernel(
    const float* __restrict__ g_inj, const float* __restrict__ m0, const float* __restrict__ g_m0,
    const float* __restrict__ feat, float* __restrict__ dl0, float* __restrict__ partial, int P, int pix_per_cta,
    float* __restrict__ dbg) {
    using L = Smem<K>;
    constexpr int NCH = L::NCH, FK = F + K;
    constexpr int SH = (NCH == 4) ? 1 : 0;            // row-major [.,K] tiles: chunk j of row r at j ^ ((r>>SH)&(NCH-1))
    constexpr uint32_t TMEM_COLS = 32;
    constexpr int MT = K / 16;                        // m-tiles (16 parts each) of GEMM2
    constexpr int RG = TILE / 8;                      // 8-row groups per tile
    constexpr int KS2 = TILE / 8;                     // k-steps of GEMM2
    constexpr uint32_t IDESC1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(K >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const uint32_t bar1 = sb + L::BAR, slot = sb + L::BAR + 16;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // B operand of GEMM1: feat[b] [K][64] split hi/lo, K-major interleave: chunk (k, j) -> (k/8)*2048 + j*128 + (k%8)*16
    for (int c = tid; c < K * 16; c += TPB) {
        const int k = c >> 4, j = c & 15;
        const float4 v = ld4(feat + ((size_t)b * K + k) * F + 4 * j);
        const float4 hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
        const float4 lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
        const uint32_t off = (k >> 3) * 2048 + j * 128 + (k & 7) * 16;
        sts4(sb + L::B_HI + off, hi);
        sts4(sb + L::B_LO + off, lo);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(tmem_base) : "r"(slot));
    const uint32_t D1 = tmem_base;
    // GEMM2 accumulators: warp w owns features [16w, 16w+16) as two n-tiles, all K parts as MT m-tiles
    float acc2[MT][2][4];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc2[a][n][i] = 0.f;

    const int p_begin = blockIdx.x * pix_per_cta, p_end = min(P, p_begin + pix_per_cta);
    // epilogue mapping: UMMA M=64 puts accumulator row r in TMEM lane (r/16)*32 + r%16, so lane l < 16 of
    // warp w owns pixel row 16w + l; the upper half-warp idles in the epilogue
    const bool epi = lane < 16;
    const int er = warp * 16 + (lane & 15);
    uint32_t phase = 0;
    for (int pt = p_begin; pt < p_end; pt += TILE) {
        const float* grow = g_inj + ((size_t)b * P + pt) * FK;
        // ---- (1) async loads: G (16 chunks/row -> UMMA layout), probabilities, external cotangent
#pragma unroll
        for (int q = 0; q < RG; ++q) {   // each warp: RG/4 row-groups x 4 passes of (8 rows x 4 chunks)
            const int rg = warp * (RG / 4) + (q >> 2), j = (q & 3) * 4 + (lane >> 3), r = rg * 8 + (lane & 7);
            cp_async16(sb + L::A_HI + j * L::A_LBO + rg * 128 + (lane & 7) * 16, grow + (size_t)r * FK + 4 * j);
        }
#pragma unroll
        for (int it = 0; it < (TILE * NCH) / TPB; ++it) {
            const int c = it * TPB + tid, r = c / NCH, j = c % NCH;
            const uint32_t off = r * (16 * NCH) + ((j ^ ((r >> SH) & (NCH - 1))) * 16);
            cp_async16(sb + L::PT + off, m0 + ((size_t)b * P + pt + r) * K + 4 * j);
            if (g_m0) cp_async16(sb + L::GT + off, g_m0 + ((size_t)b * P + pt + r) * K + 4 * j);
        }
        // the K tail columns g_inj[..., F:F+K] of the thread's own pixel row go straight to registers
        float4 tail[NCH];
        if (epi) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) tail[j] = ld4_stream(grow + (size_t)er * FK + F + 4 * j);
        }
        cp_async_commit_wait_all();
        __syncthreads();
        // ---- (2) 3xTF32 split of G, elementwise (layout-agnostic): hi in place, lo beside it
#pragma unroll 4
        for (int it = 0; it < (TILE * 16) / TPB; ++it) {
            const int c = it * TPB + tid;                                  // chunk id: block j = c / TILE, then row
            const uint32_t off = (uint32_t)(c / TILE) * L::A_LBO + (uint32_t)(c % TILE) * 16;
            const float4 v = lds4(sb + L::A_HI + off);
            const float4 hi = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
            const float4 lo = make_float4(rna_tf32(v.x - hi.x), rna_tf32(v.y - hi.y), rna_tf32(v.z - hi.z), rna_tf32(v.w - hi.w));
            sts4(sb + L::A_HI + off, hi);
            sts4(sb + L::A_LO + off, lo);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- (3) GEMM1: D1[TILE, K] = G . feat^T   (one thread issues; 8 k-steps x 3 split terms)
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const uint64_t a_hi = make_desc(sb + L::A_HI + s * 2 * L::A_LBO, L::A_LBO, 128);
                const uint64_t a_lo = make_desc(sb + L::A_LO + s * 2 * L::A_LBO, L::A_LBO, 128);
                const uint64_t b_hi = make_desc(sb + L::B_HI + s * 256, 128, 2048);
                const uint64_t b_lo = make_desc(sb + L::B_LO + s * 256, 128, 2048);
                umma_tf32(D1, a_hi, b_hi, IDESC1, s > 0 ? 1u : 0u);
                umma_tf32(D1, a_lo, b_hi, IDESC1, 1u);
                umma_tf32(D1, a_hi, b_lo, IDESC1, 1u);
            }
            umma_commit(bar1);
        }
        mbar_wait(bar1, phase);
        tc_fence_after();
        // ---- (4) epilogue, thread = pixel row: softmax backward + transposed hard mask for GEMM2
        {
            float dm[K];
#pragma unroll
            for (int c0 = 0; c0 < K; c0 += 16) tmem_ld16(D1 + ((uint32_t)(warp * 32) << 16) + c0, dm + c0);
            if (epi) {
                const int r = er;
                float pr[K], gp[K];
                const uint32_t rowoff = r * (16 * NCH);
                const int sw = (r >> SH) & (NCH - 1);
                float dot = 0.f, pmax = 0.f;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const uint32_t off = rowoff + ((j ^ sw) * 16);
                    const float4 p4 = lds4(sb + L::PT + off);
                    const float4 t4 = tail[j];
                    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (g_m0) g4 = lds4(sb + L::GT + off);
                    pr[4 * j] = p4.x; pr[4 * j + 1] = p4.y; pr[4 * j + 2] = p4.z; pr[4 * j + 3] = p4.w;
                    gp[4 * j] = dm[4 * j] + t4.x + g4.x;
                    gp[4 * j + 1] = dm[4 * j + 1] + t4.y + g4.y;
                    gp[4 * j + 2] = dm[4 * j + 2] + t4.z + g4.z;
                    gp[4 * j + 3] = dm[4 * j + 3] + t4.w + g4.w;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) { dot = fmaf(gp[k], pr[k], dot); pmax = fmaxf(pmax, pr[k]); }
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    float4 d;
                    d.x = pr[4 * j] * (gp[4 * j] - dot);
                    d.y = pr[4 * j + 1] * (gp[4 * j + 1] - dot);
                    d.z = pr[4 * j + 2] * (gp[4 * j + 2] - dot);
                    d.w = pr[4 * j + 3] * (gp[4 * j + 3] - dot);
                    sts4(sb + L::GT + rowoff + ((j ^ sw) * 16), d);
                    // transposed hard-mask tile mhT[k][pixel] (A operand of GEMM2)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float pv = pr[4 * j + i];
                        sts1(sb + L::MH + ((4 * j + i) * L::MHS + r) * 4, rna_tf32(st_value(pv == pmax ? 1.f : 0.f, pv)));
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- (5) GEMM2 (mma.sync): acc2[k, f] += sum_px mhT[k, px] * (G_hi + G_lo)[px, f]
        {
            const int g = lane >> 2, t = lane & 3;
#pragma unroll 4
            for (int s = 0; s < KS2; ++s) {
                uint32_t af[MT][4];
#pragma unroll
                for (int a = 0; a < MT; ++a) {
                    const uint32_t base = sb + L::MH + ((a * 16 + g) * L::MHS + 8 * s + t) * 4;
                    af[a][0] = lds1(base);
                    af[a][1] = lds1(base + 8 * L::MHS * 4);
                    af[a][2] = lds1(base + 16);
                    af[a][3] = lds1(base + 8 * L::MHS * 4 + 16);
                }
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    const int f = warp * 16 + n * 8 + g;
                    const uint32_t off = (f >> 2) * L::A_LBO + s * 128 + t * 16 + (f & 3) * 4;
                    const uint32_t bh0 = lds1(sb + L::A_HI + off), bh1 = lds1(sb + L::A_HI + off + 64);
                    const uint32_t bl0 = lds1(sb + L::A_LO + off), bl1 = lds1(sb + L::A_LO + off + 64);
#pragma unroll
                    for (int a = 0; a < MT; ++a) {
                        mma_tf32(acc2[a][n], af[a], bh0, bh1);
                        mma_tf32(acc2[a][n], af[a], bl0, bl1);
                    }
                }
            }
        }
        // ---- (6) coalesced store of the dl0 tile
#pragma unroll
        for (int it = 0; it < (TILE * NCH) / TPB; ++it) {
            const int c = it * TPB + tid, r = c / NCH, j = c % NCH;
            const float4 v = lds4(sb + L::GT + r * (16 * NCH) + ((j ^ ((r >> SH) & (NCH - 1))) * 16));
            st4_stream(dl0 + ((size_t)b * P + pt + r) * K + 4 * j, v);
        }
        phase ^= 1;
        __syncthreads();  // tile buffers are reused by the next iteration
    }
    // ---- dfeat partial of this CTA: mma.sync C fragment (row g / g+8 = part, cols 2t, 2t+1 = feature)
    {
        const int g = lane >> 2, t = lane & 3;
        float* dst = partial + ((size_t)b * gridDim.x + blockIdx.x) * (K * F);
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const int f = warp * 16 + n * 8 + 2 * t, k = a * 16 + g;
                dst[k * F + f] = acc2[a][n][0];
                dst[k * F + f + 1] = acc2[a][n][1];
                dst[(k + 8) * F + f] = acc2[a][n][2];
                dst[(k + 8) * F + f + 1] = acc2[a][n][3];
            }
    }
    if (dbg) {  // bring-up aid: raw TMEM [128 lanes][32 cols]
        float t[16];
        for (int c0 = 0; c0 < 32; c0 += 16) {
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, t);
            for (int i = 0; i < 16; ++i) dbg[(size_t)tid * 32 + c0 + i] = t[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

__global__ void split_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int n_per, int splits,
                                      long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / n_per;
    const int j = (int)(i % n_per);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)b * splits + sp) * n_per + j];
    out[i] = s;
}

}  // namespace tc
}  // namespace ups

using namespace ups;

static int decode_bwd_tc_impl(const float* g_inj, const float* m0, const float* g_m0, const float* feat, float* dl0,
                              float* dfeat, int B, int P, int K, int F, void* ws, size_t ws_bytes, void* stream,
                              float* dbg, int variant) {
    UPS_REQUIRE(g_inj && m0 && feat && dl0 && dfeat, "step_decode_bwd_tc: null pointer");
    UPS_REQUIRE(B >= 0 && B <= 65535, "step_decode_bwd_tc: B=%d out of range", B);
    UPS_REQUIRE(K == 16 || K == 32, "step_decode_bwd_tc: tensor-core path needs K in {16,32}, got %d", K);
    UPS_REQUIRE(F == 64, "step_decode_bwd_tc: tensor-core path needs F == 64, got %d", F);
    UPS_REQUIRE(P >= 128 && P % 128 == 0, "step_decode_bwd_tc: tensor-core path needs P %% 128 == 0, got %d", P);
    UPS_REQUIRE(aligned16(g_inj) && aligned16(m0) && aligned16(feat) && aligned16(dl0) && (!g_m0 || aligned16(g_m0)),
                "step_decode_bwd_tc: 16-byte alignment");
    if (B == 0) return UPS_OK;
    const int per = fused_pix_per_cta(B, P);
    const int splits = (int)cdiv(P, per);
    const size_t need = (size_t)B * splits * K * F * sizeof(float);
    if (!ws || ws_bytes < need) { set_error("step_decode_bwd_tc: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    dim3 grid(splits, B);
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    if (K == 16) {
        const size_t sm = tc::Smem<16>::TOTAL;
        UPS_CUDA(cudaFuncSetAttribute(tc::step_decode_bwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        tc::step_decode_bwd_tc_kernel<16><<<grid, tc::TPB, sm, s>>>(g_inj, m0, g_m0, feat, dl0, partial, P, per, dbg);
    } else {
        const size_t sm = tc::Smem<32>::TOTAL;
        UPS_CUDA(cudaFuncSetAttribute(tc::step_decode_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        tc::step_decode_bwd_tc_kernel<32><<<grid, tc::TPB, sm, s>>>(g_inj, m0, g_m0, feat, dl0, partial, P, per, dbg);
    }
    if (int rc = after_launch("step_decode_bwd_tc_kernel")) return rc;
    const long long n = (long long)B * K * F;
    tc::split_finalize_kernel<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(partial, dfeat, K * F, splits, n);
    return after_launch("split_finalize_kernel");
}

extern "C" int ups_step_decode_bwd_tc1(const float* g_inj, const float* m0, const float* g_m0, const float* feat,
                                      float* dl0, float* dfeat, int B, int P, int K, int F, void* ws, size_t ws_bytes,
                                      void* stream) {
    return decode_bwd_tc_impl(g_inj, m0, g_m0, feat, dl0, dfeat, B, P, K, F, ws, ws_bytes, stream, nullptr, 0);
}

// bring-up aid (not declared in include/ups_b200.h): dumps raw TMEM and the last hard-mask tile
extern "C" int ups_debug_decode_bwd_tc(const float* g_inj, const float* m0, const float* g_m0, const float* feat,
                                       float* dl0, float* dfeat, int B, int P, int K, int F, void* ws, size_t ws_bytes,
                                       void* stream, float* dbg, int variant) {
    return decode_bwd_tc_impl(g_inj, m0, g_m0, feat, dl0, dfeat, B, P, K, F, ws, ws_bytes, stream, dbg, variant);
}
