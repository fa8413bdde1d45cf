// K4 (decode backward) as a persistent, warp-specialised TMA + tcgen05 pipeline.
// K in {16, 32}, F = 64, P % 128 == 0.  Reference math: autodiff of
// cub/code/SB_model48i/model.py:426,434-436,482-484 (softmax -> ST(hard_max) -> unpool ++ mask).
//
//   dmask[px, k] = sum_f g_inj[px, f] * feat[k, f]            dense (P x F).(F x K): tcgen05 kind::tf32
//   dl0          = softmax_bwd(m0, dmask + g_inj[:, F:] + g_m0)
//   dfeat[k, f]  = sum_px mh[px, k] * g_inj[px, f]            mh is (almost always) one-hot: a scatter-add
//
// One CTA per SM loops over "chunks" (<= 16 tiles of 128 pixels of one sample).  Warp roles:
//   one lane           TMA producer: g_inj[128 px, 0:64] as two SWIZZLE_128B boxes of [128 x 32 floats]
//                      into a 3-stage ring -> lands directly in the UMMA K-major SW128 operand layout
//   warps 4-11         splitter: 3xTF32 split.  The raw fp32 tile IS the hi operand (the tensor core
//                      reads the top 19 bits); lo = g - trunc(g) goes to a second buffer.  The same
//                      pass scatters mon * g[px, :] into warp-private shared-memory accumulators of
//                      dfeat (lanes along f, the row's part label is warp-uniform): conflict-free
//                      8-byte read-modify-writes, no atomics, no branches on the label.
//   next warp, 1 lane  TMA producer (see above)
//   last warp, 1 lane  MMA issuer: D[128, 2K] = G_hi . [feat_hi | feat_lo]^T ; D[:, 0:K] += G_lo . feat_hi^T
//                      (M=128, N=2K / K, 8 k-steps of 8), fp32 accumulators double-buffered in TMEM
//   warps 0-3          epilogue: tcgen05.ld (thread = pixel row) -> warp-private smem transpose ->
//                      4 (8) lanes per pixel: + tail + g_m0, softmax backward, coalesced 16-byte stores.
//                      m0 / g_m0 / tail come straight from global (coalesced), prefetched one tile ahead.
// Hand-offs are mbarriers; tcgen05.commit releases the smem stage, the lo buffer and publishes the
// accumulator.  dfeat partials are reduced in a fixed order (bit-reproducible run to run).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_helpers.cuh"

namespace ups {
namespace tma {

constexpr int TILE = 128;  // pixels per tile = UMMA M
constexpr int F = 64;
constexpr int CHUNK_TILES = 16;
constexpr int SPL_WARP0 = 4;
constexpr unsigned FULLM = 0xffffffffu;

// DEEP = 1 (experiment, UPS_K4_DEEP=1): one more TMA stage and a single lo buffer instead of two.  Measured SLOWER
// (0.519 vs 0.440 ms at CUB B=256): letting the splitter run one tile ahead of the MMA matters more than a fourth tile in
// flight.  profiles/r02_tuning.md.
template <int K, int DEEP = 0>
struct Cfg {
    static constexpr int NST = ((K == 16) ? 3 : 2) + DEEP;
    static constexpr int NLO = DEEP ? 1 : 2;
    static constexpr int NSPL = (K == 16) ? 8 : 4;     // splitter warps
    static constexpr int W_TMA = SPL_WARP0 + NSPL, W_MMA = W_TMA + 1;
    static constexpr int TPB = (W_MMA + 1) * 32;
    static constexpr int BLK = TILE * 128;             // one SW128 block: [128 rows][32 floats]
    static constexpr int G_BYTES = 2 * BLK;            // 32 KB
    static constexpr int NB = 2 * K;                   // rows of [feat_hi | feat_lo]
    static constexpr int B_BLK = NB * 128;
    static constexpr int B_BYTES = 2 * B_BLK;
    static constexpr int DM_BYTES = TILE * K * 4;
    static constexpr int ACC_BYTES = NSPL * K * F * 4;  // warp-private dfeat accumulators
    static constexpr int STAGE0 = 0;
    static constexpr int LO0 = NST * G_BYTES;
    static constexpr int B0 = LO0 + NLO * G_BYTES;
    static constexpr int DM = B0 + 2 * B_BYTES;
    static constexpr int ACC = DM + DM_BYTES;
    static constexpr int RINFO = ACC + ACC_BYTES;       // 128 rows x (mask, mon)
    static constexpr int BAR = RINFO + 1024;
    static constexpr int TOTAL = BAR + 512;
    static constexpr uint32_t TMEM_COLS = 4 * K;       // 2 accumulator buffers of 2K columns (64 / 128)
};

// ------------------------------------------------------------------ schedule shared by all roles
// Work unit = chunk: ch -> sample b = ch / splits, tiles [sp*CHUNK_TILES, min(tps, +CHUNK_TILES)) of that
// sample.  Chunks are handed out dynamically: the TMA producer lane draws chunk ids from a global counter
// (atomicAdd) and publishes them through a small shared-memory ring guarded by mbarriers; every other role
// reads the same sequence from the ring.  A CTA that loses its SM to another kernel for a while (NCCL's
// all-reduce runs beside this kernel in the data-parallel step) simply draws fewer chunks.
constexpr int NSLOT = 8;

struct ChunkRing {
    uint32_t ring, bar_sfull, bar_sempty;   // shared-memory addresses
    uint32_t idx;                           // chunks drawn / consumed so far by this role
    int n_chunks;
    // producer lane only
    __device__ __forceinline__ int draw(unsigned int* counter) {
        const uint32_t slot = idx % NSLOT, ph = (idx / NSLOT) & 1;
        mbar_wait(bar_sempty + 8 * slot, ph ^ 1);
        int v = (int)atomicAdd(counter, 1u);
        if (v > n_chunks) v = n_chunks;
        asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(ring + 4 * slot), "r"(v) : "memory");
        mbar_arrive(bar_sfull + 8 * slot);
        ++idx;
        return v;
    }
    // consumers: `collective` = the whole warp calls this (lane 0 signals for the warp)
    __device__ __forceinline__ int take(bool collective, int lane) {
        const uint32_t slot = idx % NSLOT, ph = (idx / NSLOT) & 1;
        mbar_wait(bar_sfull + 8 * slot, ph);
        int v;
        asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(ring + 4 * slot) : "memory");
        if (collective) __syncwarp();
        if (!collective || lane == 0) mbar_arrive(bar_sempty + 8 * slot);
        ++idx;
        return v;
    }
};

struct TileIter {
    int ch, n_chunks, splits, tps;   // tps = tiles per sample
    int t, t_begin, t_end, b;
    __device__ __forceinline__ void set_chunk(int c, int n_chunks_, int splits_, int tps_) {
        ch = c; n_chunks = n_chunks_; splits = splits_; tps = tps_;
        t = t_begin = t_end = b = 0;
        if (ch < n_chunks) {
            b = ch / splits;
            const int sp = ch - b * splits;
            t_begin = sp * CHUNK_TILES;
            t_end = min(tps, t_begin + CHUNK_TILES);
            t = t_begin;
        }
    }
    __device__ __forceinline__ bool valid() const { return ch < n_chunks; }
    __device__ __forceinline__ long long row0() const { return ((long long)b * tps + t) * TILE; }
    __device__ __forceinline__ bool first_in_chunk() const { return t == t_begin; }
    __device__ __forceinline__ bool last_in_chunk() const { return t + 1 == t_end; }
    // consumer roles: step to the next tile, taking the next chunk from the ring at a chunk boundary
    __device__ __forceinline__ void next(ChunkRing& cr, bool collective, int lane) {
        if (++t == t_end) set_chunk(cr.take(collective, lane), n_chunks, splits, tps);
    }
};

template <int K, int DEEP>
__global__ void __launch_bounds__(Cfg<K, DEEP>::TPB, 1) step_decode_bwd_tma_kernel(
    const __grid_constant__ CUtensorMap tmap_g, const float* __restrict__ g_inj, const float* __restrict__ m0,
    const float* __restrict__ g_m0, const float* __restrict__ feat, float* __restrict__ dl0,
    float* __restrict__ partial, unsigned int* __restrict__ chunk_counter, int n_chunks, int splits, int tps, int gm_row) {
    using L = Cfg<K, DEEP>;
    constexpr int NLO = L::NLO;
    constexpr int NST = L::NST, FK = F + K, NSPL = L::NSPL, W_TMA = L::W_TMA, W_MMA = L::W_MMA;
    constexpr int LPP = K / 4;            // lanes per pixel in the coalesced [px][K] mapping
    constexpr int PW = 32 / LPP;          // pixels per warp pass
    constexpr int NPASS = LPP;            // passes to cover an epilogue warp's 32 rows
    constexpr int ROWS = TILE / NSPL;     // rows per splitter warp
    constexpr int SPASS = ROWS / PW;      // passes to cover them
    constexpr uint32_t IDESC_2K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * K) >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(K >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B operands need 1024-byte alignment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // barriers
    const uint32_t bar_full = sb + L::BAR;                 // [NST]  TMA -> splitter, MMA
    const uint32_t bar_empty = bar_full + 8 * NST;         // [NST]  MMA commit -> TMA
    const uint32_t bar_lo_ready = bar_empty + 8 * NST;     // [2]    splitter -> MMA
    const uint32_t bar_lo_free = bar_lo_ready + 16;        // [2]    MMA commit -> splitter
    const uint32_t bar_tm_full = bar_lo_free + 16;         // [2]    MMA commit -> epilogue
    const uint32_t bar_tm_empty = bar_tm_full + 16;        // [2]    epilogue -> MMA
    const uint32_t bar_sfull = bar_tm_empty + 16;          // [NSLOT] chunk ring: producer -> consumers
    const uint32_t bar_sempty = bar_sfull + 8 * NSLOT;     // [NSLOT] consumers -> producer
    const uint32_t ring = bar_sempty + 8 * NSLOT;          // [NSLOT] chunk ids
    const uint32_t slot = ring + 4 * NSLOT;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(slot), "r"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_lo_ready + 8 * i, NSPL * 32);
            mbar_init(bar_lo_free + 8 * i, 1);
            mbar_init(bar_tm_full + 8 * i, 1);
            mbar_init(bar_tm_empty + 8 * i, 128);
        }
        for (int i = 0; i < NSLOT; ++i) { mbar_init(bar_sfull + 8 * i, 1); mbar_init(bar_sempty + 8 * i, 4 + NSPL + 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid == W_TMA * 32) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmap_g)) : "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(tmem_base) : "r"(slot));

    ChunkRing cr;
    cr.ring = ring; cr.bar_sfull = bar_sfull; cr.bar_sempty = bar_sempty; cr.idx = 0; cr.n_chunks = n_chunks;
    TileIter ti;

    if (warp == W_TMA) {
        // ================================================================= TMA producer
        if (lane == 0) {
            // one chunk of look-ahead: the id of the next chunk is published before this chunk's loads are
            // issued, so consumers that prefetch across a chunk boundary never wait for the producer
            int cur = cr.draw(chunk_counter);
            int nxt = cur < n_chunks ? cr.draw(chunk_counter) : n_chunks;
            uint32_t it = 0;
            while (cur < n_chunks) {
                ti.set_chunk(cur, n_chunks, splits, tps);
                for (; ti.t < ti.t_end; ++ti.t, ++it) {
                    const uint32_t s = it % NST, n = it / NST;
                    mbar_wait(bar_empty + 8 * s, (n & 1) ^ 1);
                    mbar_expect_tx(bar_full + 8 * s, L::G_BYTES);
                    const int row = (int)ti.row0();
                    tma_load_2d(sb + L::STAGE0 + s * L::G_BYTES, &tmap_g, 0, row, bar_full + 8 * s);
                    tma_load_2d(sb + L::STAGE0 + s * L::G_BYTES + L::BLK, &tmap_g, 32, row, bar_full + 8 * s);
                }
                cur = nxt;
                if (cur < n_chunks) nxt = cr.draw(chunk_counter);
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ================================================================= MMA issuer
        if (lane == 0) {
            uint32_t chunk_n = 0;
            ti.set_chunk(cr.take(false, 0), n_chunks, splits, tps);
            for (uint32_t it = 0; ti.valid(); ti.next(cr, false, 0), ++it) {
                const uint32_t s = it % NST, n = it / NST, j = it & 1, u = (it >> 1) & 1;
                if (ti.first_in_chunk() && it > 0) ++chunk_n;
                const uint32_t bsel = chunk_n & 1;
                mbar_wait(bar_full + 8 * s, n & 1);
                const uint32_t jl = NLO == 2 ? j : 0u, ul = NLO == 2 ? u : (it & 1);   // lo buffer and its phase parity
                mbar_wait(bar_lo_ready + 8 * jl, ul);
                mbar_wait(bar_tm_empty + 8 * j, u ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + j * (2 * K);
                const uint32_t a_hi = sb + L::STAGE0 + s * L::G_BYTES, a_lo = sb + L::LO0 + jl * L::G_BYTES;
                const uint32_t bb = sb + L::B0 + bsel * L::B_BYTES;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t ao = (ks >> 2) * L::BLK + (ks & 3) * 32, bo = (ks >> 2) * L::B_BLK + (ks & 3) * 32;
                    const uint64_t bd = sw128_desc(bb + bo);
                    umma_tf32(d, sw128_desc(a_hi + ao), bd, IDESC_2K, ks > 0 ? 1u : 0u);   // hi.hi | hi.lo
                    umma_tf32(d, sw128_desc(a_lo + ao), bd, IDESC_K, 1u);                   // lo.hi
                }
                umma_commit(bar_empty + 8 * s);
                umma_commit(bar_lo_free + 8 * jl);
                umma_commit(bar_tm_full + 8 * j);
            }
        }
        __syncwarp();
    } else if (warp >= SPL_WARP0) {
        // ================================================================= splitter + dfeat scatter
        const int sw = warp - SPL_WARP0, st = tid - SPL_WARP0 * 32;      // st in [0, NSPL*32)
        const int c = lane & (LPP - 1), q = lane / LPP;
        // warp-private dfeat accumulators [K][64] in shared memory: the row's part label is warp-uniform,
        // lane = feature pair, so each update is a conflict-free 8-byte read-modify-write
        const uint32_t accw = sb + L::ACC + sw * (K * F * 4);
        auto zero_acc = [&]() {
#pragma unroll
            for (int i = 0; i < (K * F) / 128; ++i) sts4(accw + (i * 32 + lane) * 16, make_float4(0.f, 0.f, 0.f, 0.f));
        };
        zero_acc();
        ti.set_chunk(cr.take(true, lane), n_chunks, splits, tps);
        // this thread's slice of the B operand: part bn, 16-byte chunks [bc0, bc0 + BCH)
        constexpr int BCH = (K * 16) / (NSPL * 32);    // chunks per thread
        const int bn = st / (16 / BCH), bc0 = (st % (16 / BCH)) * BCH;
        float4 fpre[BCH];
        if (ti.valid()) {
#pragma unroll
            for (int i = 0; i < BCH; ++i) fpre[i] = ld4(feat + ((size_t)ti.b * K + bn) * F + 4 * (bc0 + i));
        }
        float4 pm[SPASS];
        if (ti.valid()) {
#pragma unroll
            for (int p = 0; p < SPASS; ++p)
                pm[p] = ld4(m0 + ((size_t)ti.row0() + sw * ROWS + p * PW + q) * K + 4 * c);
        }
        uint32_t chunk_n = 0;
        for (uint32_t it = 0; ti.valid(); ++it) {
            const uint32_t s = it % NST, n = it / NST, j = it & 1, u = (it >> 1) & 1;
            const bool first = ti.first_in_chunk(), last = ti.last_in_chunk();
            const int cur_chunk = ti.ch;
            if (first && it > 0) ++chunk_n;
            // per-row hard mask (bit k set where p_k == max) and the straight-through value of the maxima,
            // parked in a warp-private smem table so that the row loop below can stay rolled
            const uint32_t rinfo = sb + L::RINFO + sw * (ROWS * 8);
#pragma unroll
            for (int p = 0; p < SPASS; ++p) {
                const float4 v = pm[p];
                const float pmax = group_max<LPP>(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
                uint32_t mloc = (v.x == pmax ? 1u : 0u) | (v.y == pmax ? 2u : 0u) | (v.z == pmax ? 4u : 0u) | (v.w == pmax ? 8u : 0u);
                mloc <<= 4 * c;
#pragma unroll
                for (int o = 1; o < LPP; o <<= 1) mloc |= __shfl_xor_sync(FULLM, mloc, o);
                if (c == 0) sts2(rinfo + (p * PW + q) * 8, make_float2(__uint_as_float(mloc), st_value(1.0f, pmax)));
            }
            __syncwarp();
            const uint32_t jl = NLO == 2 ? j : 0u, ul = NLO == 2 ? u : (it & 1);
            mbar_wait(bar_lo_free + 8 * jl, ul ^ 1);    // the MMA of the previous user of this lo buffer has finished (and older B buffers)
            if (first) {
                // B operand of this chunk: [feat_hi (rows 0..K-1) | feat_lo (rows K..2K-1)], K-major SW128
                const uint32_t bb = sb + L::B0 + (chunk_n & 1) * L::B_BYTES;
#pragma unroll
                for (int i = 0; i < BCH; ++i) {
                    const int ch16 = bc0 + i, h = ch16 >> 3, cc = ch16 & 7;
                    const float4 v = fpre[i];
                    const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
                    sts4(bb + h * L::B_BLK + bn * 128 + ((cc ^ (bn & 7)) * 16), hi);
                    sts4(bb + h * L::B_BLK + (K + bn) * 128 + ((cc ^ ((K + bn) & 7)) * 16), lo);
                }
            }
            // advance the schedule now so that the next tile's m0 (and the next chunk's feat) are in flight
            ti.next(cr, true, lane);
            if (ti.valid()) {
#pragma unroll
                for (int p = 0; p < SPASS; ++p)
                    pm[p] = ld4(m0 + ((size_t)ti.row0() + sw * ROWS + p * PW + q) * K + 4 * c);
                if (ti.first_in_chunk()) {
#pragma unroll
                    for (int i = 0; i < BCH; ++i) fpre[i] = ld4(feat + ((size_t)ti.b * K + bn) * F + 4 * (bc0 + i));
                }
            }
            mbar_wait(bar_full + 8 * s, n & 1);
            const uint32_t hi_base = sb + L::STAGE0 + s * L::G_BYTES, lo_base = sb + L::LO0 + jl * L::G_BYTES;
            // lane -> floats (2*lane, 2*lane+1) of the row: block lane/16, chunk (lane%16)/2, half lane%2
            const uint32_t lane_off = (lane >> 4) * L::BLK + (lane & 1) * 8;
            const int lch = (lane & 15) >> 1;
            const uint32_t acc_lane = accw + lane * 8;
#pragma unroll 1
            for (int r8 = 0; r8 < ROWS / 8; ++r8) {
                // (a) 8 rows of the tile: all loads first (the asm statements keep program order), then lo
                float2 g[8], ri[8];
                uint32_t off[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = sw * ROWS + r8 * 8 + i;         // r & 7 == i
                    off[i] = r * 128 + ((lch ^ i) * 16) + lane_off;
                    g[i] = lds2(hi_base + off[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) ri[i] = lds2(rinfo + (r8 * 8 + i) * 8);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    sts2(lo_base + off[i], make_float2(g[i].x - tf32_hi(g[i].x), g[i].y - tf32_hi(g[i].y)));
                // (b) dfeat[label] += mon * g.  Rows of the batch that share a label are first combined in
                // registers (every pair compared once, warp-uniform), so that each label is read-modified-
                // written once and the 8 smem round trips overlap instead of chaining.
                int kk[8];
                bool simple = true;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t mb = __float_as_uint(ri[i].x);
                    kk[i] = __ffs(mb) - 1;
                    simple = simple && (mb != 0u) && ((mb & (mb - 1u)) == 0u);
                }
                if (simple) {
                    float2 cc[8];
                    uint32_t alive[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) cc[i] = make_float2(ri[i].y * g[i].x, ri[i].y * g[i].y);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        alive[i] = 1u;
#pragma unroll
                        for (int jj = 0; jj < i; ++jj)
                            if (kk[i] == kk[jj]) { cc[jj].x += cc[i].x; cc[jj].y += cc[i].y; alive[i] = 0u; }
                    }
                    float2 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { v[i] = make_float2(0.f, 0.f); lds2_if(v[i], acc_lane + kk[i] * (F * 4), alive[i]); }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        sts2_if(acc_lane + kk[i] * (F * 4), make_float2(v[i].x + cc[i].x, v[i].y + cc[i].y), alive[i]);
                } else {
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {          // tied maxima somewhere in these rows: plain sequential form
                        const float2 rii = lds2(rinfo + (r8 * 8 + i) * 8);
                        uint32_t mb = __float_as_uint(rii.x);
                        const float mo = rii.y;
                        const float2 gi = lds2(hi_base + (sw * ROWS + r8 * 8 + i) * 128 + ((lch ^ i) * 16) + lane_off);
                        while (mb) {
                            const int k = __ffs(mb) - 1;
                            mb &= mb - 1;
                            const uint32_t a = acc_lane + k * (F * 4);
                            float2 t = lds2(a);
                            t.x = fmaf(mo, gi.x, t.x);
                            t.y = fmaf(mo, gi.y, t.y);
                            sts2(a, t);
                        }
                    }
                }
            }
            __syncwarp();                              // rinfo is rewritten by the next tile
            fence_proxy_async();
            mbar_arrive(bar_lo_ready + 8 * jl);
            if (last) {
                // dfeat partial of this chunk: fixed-order sum over the splitter warps -> workspace
                asm volatile("bar.sync 1, %0;\n" ::"n"(NSPL * 32) : "memory");
                float* dst = partial + (size_t)cur_chunk * (K * F);
                const uint32_t acc0 = sb + L::ACC;
                for (int i = st; i < K * F; i += NSPL * 32) {
                    float v = lds1(acc0 + i * 4);
#pragma unroll
                    for (int w = 1; w < NSPL; ++w) v += lds1(acc0 + (w * K * F + i) * 4);
                    dst[i] = v;
                }
                asm volatile("bar.sync 1, %0;\n" ::"n"(NSPL * 32) : "memory");
                zero_acc();
            }
        }
    } else {
        // ================================================================= epilogue (warps 0-3)
        const int c = lane & (LPP - 1), q = lane / LPP;
        const uint32_t dmw = sb + L::DM;
        float4 p4[NPASS], gm4[NPASS], tl4[NPASS];
        auto prefetch = [&](long long row0) {
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
                const size_t r = (size_t)row0 + warp * 32 + p * PW + q;
                p4[p] = ld4_stream(m0 + r * K + 4 * c);
                gm4[p] = g_m0 ? ld_row4<LPP>(g_m0, r, c, gm_row, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
                tl4[p] = ld4_stream(g_inj + r * FK + F + 4 * c);
            }
        };
        ti.set_chunk(cr.take(true, lane), n_chunks, splits, tps);
        if (ti.valid()) prefetch(ti.row0());
        for (uint32_t it = 0; ti.valid(); ++it) {
            const uint32_t j = it & 1, u = (it >> 1) & 1;
            const long long row0 = ti.row0();
            ti.next(cr, true, lane);
            mbar_wait(bar_tm_full + 8 * j, u);
            tc_fence_after();
            // accumulator row (thread = pixel row warp*32 + lane): dm = D[:, 0:K] + D[:, K:2K]
            {
                float d[2 * K];
#pragma unroll
                for (int c0_ = 0; c0_ < 2 * K; c0_ += 32)
                    tmem_ld32(tmem_base + j * (2 * K) + c0_ + ((uint32_t)(warp * 32) << 16), d + c0_);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(bar_tm_empty + 8 * j);
                const int r = warp * 32 + lane;
                const int swz = (K == 16) ? ((r >> 1) & 3) : (r & 7);
#pragma unroll
                for (int jj = 0; jj < LPP; ++jj)
                    sts4(dmw + r * (K * 4) + ((jj ^ swz) * 16),
                         make_float4(d[4 * jj] + d[K + 4 * jj], d[4 * jj + 1] + d[K + 4 * jj + 1],
                                     d[4 * jj + 2] + d[K + 4 * jj + 2], d[4 * jj + 3] + d[K + 4 * jj + 3]));
            }
            __syncwarp();
            const bool more = ti.valid();
            const long long row0n = more ? ti.row0() : 0;
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
                const int r = warp * 32 + p * PW + q;
                const int swz = (K == 16) ? ((r >> 1) & 3) : (r & 7);
                const float4 dm = lds4(dmw + r * (K * 4) + ((c ^ swz) * 16));
                const float4 pr = p4[p];
                const float4 gp = make_float4(dm.x + tl4[p].x + gm4[p].x, dm.y + tl4[p].y + gm4[p].y,
                                              dm.z + tl4[p].z + gm4[p].z, dm.w + tl4[p].w + gm4[p].w);
                float dot = fmaf(gp.w, pr.w, fmaf(gp.z, pr.z, fmaf(gp.y, pr.y, gp.x * pr.x)));
                dot = group_sum<LPP>(dot);
                st4_stream(dl0 + ((size_t)row0 + r) * K + 4 * c,
                           make_float4(pr.x * (gp.x - dot), pr.y * (gp.y - dot), pr.z * (gp.z - dot), pr.w * (gp.w - dot)));
                if (more) {
                    const size_t rn = (size_t)row0n + r;
                    p4[p] = ld4_stream(m0 + rn * K + 4 * c);
                    gm4[p] = g_m0 ? ld_row4<LPP>(g_m0, rn, c, gm_row, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tl4[p] = ld4_stream(g_inj + rn * FK + F + 4 * c);
                }
            }
            __syncwarp();   // the warp's dm rows are rewritten by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(L::TMEM_COLS) : "memory");
}

__global__ void chunk_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int n_per, int splits,
                                      long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / n_per;
    const int j = (int)(i % n_per);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)b * splits + sp) * n_per + j];
    out[i] = s;
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

}  // namespace tma

size_t decode_bwd_tma_ws_bytes(int B, int P, int K, int F) {
    if (B <= 0 || P <= 0 || P % tma::TILE) return 0;
    const int tps = P / tma::TILE;
    const size_t splits = (size_t)cdiv(tps, tma::CHUNK_TILES);
    return (size_t)B * splits * K * F * sizeof(float) + 512;   // + the chunk counter (256-byte aligned)
}

}  // namespace ups

using namespace ups;

extern "C" int ups_step_decode_bwd_tc(const float* g_inj, const float* m0, const float* g_m0, const float* feat,
                                      float* dl0, float* dfeat, int B, int P, int K, int F, void* ws, size_t ws_bytes,
                                      void* stream) {
    return ups_step_decode_bwd_tc_rows(g_inj, m0, g_m0, K, feat, dl0, dfeat, B, P, K, F, ws, ws_bytes, stream);
}

extern "C" int ups_step_decode_bwd_tc_rows(const float* g_inj, const float* m0, const float* g_m0, int gm_row, const float* feat,
                                           float* dl0, float* dfeat, int B, int P, int K, int F, void* ws, size_t ws_bytes,
                                           void* stream) {
    UPS_REQUIRE(g_inj && m0 && feat && dl0 && dfeat, "step_decode_bwd_tc: null pointer");
    UPS_REQUIRE(gm_row >= 1 && gm_row <= K, "step_decode_bwd_tc: g_m0 rows of %d floats for K=%d", gm_row, K);
    UPS_REQUIRE(B >= 0 && B <= 65535, "step_decode_bwd_tc: B=%d out of range", B);
    UPS_REQUIRE(K == 16 || K == 32, "step_decode_bwd_tc: tensor-core path needs K in {16,32}, got %d", K);
    UPS_REQUIRE(F == 64, "step_decode_bwd_tc: tensor-core path needs F == 64, got %d", F);
    UPS_REQUIRE(P >= 128 && P % 128 == 0, "step_decode_bwd_tc: tensor-core path needs P %% 128 == 0, got %d", P);
    UPS_REQUIRE((long long)B * P < (1ll << 31), "step_decode_bwd_tc: B*P=%lld rows exceed the TMA coordinate range", (long long)B * P);
    UPS_REQUIRE(aligned16(g_inj) && aligned16(m0) && aligned16(feat) && aligned16(dl0) && (!g_m0 || gm_row < K || aligned16(g_m0)),
                "step_decode_bwd_tc: 16-byte alignment");
    if (B == 0) return UPS_OK;
    const int tps = P / tma::TILE;
    const int splits = (int)cdiv(tps, tma::CHUNK_TILES);
    const int n_chunks = B * splits;
    const size_t need = decode_bwd_tma_ws_bytes(B, P, K, F);
    if (!ws || ws_bytes < need) { set_error("step_decode_bwd_tc: workspace %zu < %zu bytes", ws_bytes, need); return UPS_E_WORKSPACE; }
    tma::EncodeTiledFn enc = tma::encode_tiled_fn();
    UPS_REQUIRE(enc != nullptr, "step_decode_bwd_tc: cuTensorMapEncodeTiled not available from the driver");
    // g_inj as a 2-D tensor [B*P rows][F+K floats]; box = [128 rows][32 floats], SWIZZLE_128B.  The descriptor depends
    // on (pointer, rows, row length) only: the last one encoded on this thread is reused (PartStep calls with the same
    // persistent buffer every step), so the steady state does no driver call here.
    struct MapCache { const float* ptr; long long rows; int fk; CUtensorMap map; };
    static thread_local MapCache cache = {nullptr, 0, 0, {}};
    if (cache.ptr != g_inj || cache.rows != (long long)B * P || cache.fk != F + K) {
        const cuuint64_t gdim[2] = {(cuuint64_t)(F + K), (cuuint64_t)B * (cuuint64_t)P};
        const cuuint64_t gstr[1] = {(cuuint64_t)(F + K) * sizeof(float)};
        const cuuint32_t box[2] = {32, (cuuint32_t)tma::TILE};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = enc(&cache.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g_inj), gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { cache.ptr = nullptr; set_error("step_decode_bwd_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr); return UPS_E_CUDA; }
        cache.ptr = g_inj; cache.rows = (long long)B * P; cache.fk = F + K;
    }
    const CUtensorMap& tmap = cache.map;
    const int grid = n_chunks < NUM_SMS ? n_chunks : NUM_SMS;
    cudaStream_t s = as_stream(stream);
    float* partial = static_cast<float*>(ws);
    const size_t counter_off = (((size_t)B * splits * K * F * sizeof(float)) + 255) & ~(size_t)255;
    unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + counter_off);
    UPS_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
    static const int deep = []() { const char* e = getenv("UPS_K4_DEEP"); return e ? atoi(e) : 0; }();
#define UPS_K4(KK, DD)                                                                                                       \
    {                                                                                                                        \
        const size_t sm = tma::Cfg<KK, DD>::TOTAL + 1024;                                                                    \
        static const cudaError_t attr = cudaFuncSetAttribute(tma::step_decode_bwd_tma_kernel<KK, DD>,                        \
                                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);          \
        UPS_CUDA(attr);   /* set once per process and kernel instance, not per call */                                       \
        tma::step_decode_bwd_tma_kernel<KK, DD><<<grid, tma::Cfg<KK, DD>::TPB, sm, s>>>(tmap, g_inj, m0, g_m0, feat, dl0,     \
                                                                                      partial, counter, n_chunks, splits,   \
                                                                                      tps, gm_row);                         \
    }
    if (K == 16) { if (deep) UPS_K4(16, 1) else UPS_K4(16, 0) } else { if (deep) UPS_K4(32, 1) else UPS_K4(32, 0) }
#undef UPS_K4
    if (int rc = after_launch("step_decode_bwd_tma_kernel")) return rc;
    const long long n = (long long)B * K * F;
    tma::chunk_finalize_kernel<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(partial, dfeat, K * F, splits, n);
    return after_launch("chunk_finalize_kernel");
}
