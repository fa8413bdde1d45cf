// K1: thin-plate-spline grid evaluation + bilinear gather (forward), as a device-side body so that it can run as its
// own kernel (tps.cu) and as one role of the fused forward kernel (step_fwd_fused.cu).
// Reference: baselines/unsupervised-disentangling/transformations.py:93-244 (_meshgrid, _transform, _interpolate).
#pragma once
#include "common.cuh"
#include "pk_math.cuh"

namespace ups {

// ------------------------------------------------------------------ a3(iii-v): grid + sample
constexpr int WARP_TPB = 128;  // threads per CTA
constexpr int WARP_PPT = 4;    // output pixels per thread (strided by the CTA width)

struct SampleConst {
    float T[22];
    float qx[8], qy[8];
    float sy, sx, my, mx;  // optional move/scal branch (transformations.py:202-208)
};

__device__ __forceinline__ void load_sample_const(SampleConst& k, float* sm, const float* coord, const float* T,
                                                  const float* move, const float* scal, int b) {
    // 22 T + 16 coord + 4 move/scal through shared memory, then into registers
    const int t = threadIdx.x;
    if (t < 22) sm[t] = T[b * 22 + t];
    else if (t < 38) sm[t] = coord[b * 16 + (t - 22)];
    else if (t < 40) sm[t] = scal ? scal[b * 2 + (t - 38)] : 1.0f;
    else if (t < 42) sm[t] = move ? move[b * 2 + (t - 40)] : 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 22; ++i) k.T[i] = sm[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { k.qx[i] = sm[22 + 2 * i + 1]; k.qy[i] = sm[22 + 2 * i + 0]; }  // ::-1 flip
    k.sy = sm[38]; k.sx = sm[39]; k.my = sm[40]; k.mx = sm[41];
}

__device__ __forceinline__ void sample_position(const SampleConst& k, bool has_move, int i, int j, float step_h,
                                                float step_w, float& x_s, float& y_s) {
    tps_coords(k.T, k.qx, k.qy, lin_at(j, step_w), lin_at(i, step_h), x_s, y_s);
    if (has_move) {
        y_s = __fadd_rn(__fmul_rn(y_s, k.sy), k.my);
        x_s = __fadd_rn(__fmul_rn(x_s, k.sx), k.mx);
    }
}

// ---- packed (fp32x2) evaluation of the sample position and the bilinear stencil: pk_math.cuh.
// Lane results are bit-identical to tps_coords / bilinear_stencil / bilinear_mix of canon_math.cuh;
// two control points (or two image channels) share every FFMA2.
struct SampleConstPk {
    pk::f2 T01[11];        // (T[0][i], T[1][i])
    pk::f2 qx2[4], qy2[4]; // control points (2n, 2n+1), after the ::-1 flip
    float sy, sx, my, mx;
};

__device__ __forceinline__ void load_sample_const_pk(SampleConstPk& k, float* sm, const float* coord, const float* T,
                                                     const float* move, const float* scal, int b) {
    const int t = threadIdx.x;
    if (t < 22) sm[t] = T[b * 22 + t];
    else if (t < 38) sm[t] = coord[b * 16 + (t - 22)];
    else if (t < 40) sm[t] = scal ? scal[b * 2 + (t - 38)] : 1.0f;
    else if (t < 42) sm[t] = move ? move[b * 2 + (t - 40)] : 0.0f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 11; ++i) k.T01[i] = pk::pack(sm[i], sm[11 + i]);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        k.qx2[n] = pk::pack(sm[22 + 4 * n + 1], sm[22 + 4 * n + 3]);
        k.qy2[n] = pk::pack(sm[22 + 4 * n + 0], sm[22 + 4 * n + 2]);
    }
    k.sy = sm[38]; k.sx = sm[39]; k.my = sm[40]; k.mx = sm[41];
}

// (x_s, y_s) packed; same operation order as canon_math.cuh::tps_coords
__device__ __forceinline__ pk::f2 sample_position_pk(const pk::Ops& o, const SampleConstPk& k, bool has_move, float x_t,
                                                     float y_t) {
    const pk::f2 xt = pk::splat(x_t), yt = pk::splat(y_t);
    pk::f2 a = o.add(k.T01[0], o.mul(k.T01[1], xt));
    a = o.add(a, o.mul(k.T01[2], yt));
    pk::f2 d2[4], arg[4], lg[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const pk::f2 dx = o.sub(xt, k.qx2[n]);
        const pk::f2 dy = o.sub(yt, k.qy2[n]);
        d2[n] = o.add(o.mul(dx, dx), o.mul(dy, dy));
        arg[n] = o.add(d2[n], pk::splat(1e-6f));
    }
    pk::log_canon2n<4>(o, arg, lg);   // the 8 radial-basis logs as 4 interleaved FFMA2 chains
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        float r0, r1;
        pk::unpack(o.mul(d2[n], lg[n]), r0, r1);
        a = o.add(a, o.mul(k.T01[3 + 2 * n], pk::splat(r0)));
        a = o.add(a, o.mul(k.T01[4 + 2 * n], pk::splat(r1)));
    }
    if (has_move) {
        float x_s, y_s;
        pk::unpack(a, x_s, y_s);
        y_s = __fadd_rn(__fmul_rn(y_s, k.sy), k.my);
        x_s = __fadd_rn(__fmul_rn(x_s, k.sx), k.mx);
        a = pk::pack(x_s, y_s);
    }
    return a;
}

struct BilinearPk {
    int oa, ob, oc, od;    // element offsets of the four clipped corners (times C)
    pk::f2 wab, wcd;       // (wa, wb), (wc, wd)
};

__device__ __forceinline__ BilinearPk bilinear_stencil_pk(const pk::Ops& o, pk::f2 xy_s, pk::f2 WHf, float wmax, float hmax,
                                                          int W, int Cc) {
    // (X, Y) = ((s + 1) * (W, H)) / 2   -- the division by 2 is the exact multiplication by 0.5
    const pk::f2 XY = o.mul(o.mul(o.add(xy_s, o.one), WHf), pk::splat(0.5f));
    float X, Y;
    pk::unpack(XY, X, Y);
    const float fx = floorf(X), fy = floorf(Y);
    float fx1, fy1;
    pk::unpack(o.add(pk::pack(fx, fy), o.one), fx1, fy1);
    const float x0f = fminf(fmaxf(fx, 0.0f), wmax), x1f = fminf(fmaxf(fx1, 0.0f), wmax);
    const float y0f = fminf(fmaxf(fy, 0.0f), hmax), y1f = fminf(fmaxf(fy1, 0.0f), hmax);
    const int x0 = (int)x0f, x1 = (int)x1f, y0 = (int)y0f, y1 = (int)y1f;
    float dx1, dy1, dx0, dy0;
    pk::unpack(o.sub(pk::pack(x1f, y1f), XY), dx1, dy1);
    pk::unpack(o.sub(XY, pk::pack(x0f, y0f)), dx0, dy0);
    const pk::f2 dy10 = pk::pack(dy1, dy0);
    BilinearPk s;
    s.wab = o.mul(pk::splat(dx1), dy10);
    s.wcd = o.mul(pk::splat(dx0), dy10);
    s.oa = (y0 * W + x0) * Cc; s.ob = (y1 * W + x0) * Cc;
    s.oc = (y0 * W + x1) * Cc; s.od = (y1 * W + x1) * Cc;
    return s;
}

// two channel values at once: lanes = (value A, value B) of the four corners
__device__ __forceinline__ pk::f2 bilinear_mix_pk(const pk::Ops& o, pk::f2 wa, pk::f2 wb, pk::f2 wc, pk::f2 wd, pk::f2 Ia,
                                                  pk::f2 Ib, pk::f2 Ic, pk::f2 Id) {
    pk::f2 r = o.add(o.mul(wa, Ia), o.mul(wb, Ib));
    r = o.add(r, o.mul(wc, Ic));
    return o.add(r, o.mul(wd, Id));
}

// One thread per output pixel.  A second image set U2 (first N2 samples) can ride on the same
// sample positions: CUB warps view0 and view0_target with the same parameters
// (cub/code/SB_model48i/model.py:306-309), so the 8 radial-basis evaluations are paid once.
// MINB = resident CTAs per SM asked of ptxas: besides capping registers, an explicit value makes ptxas
// keep the interleaving of the four log chains (without it, it re-serialises them to save registers).
template <int C>
__device__ __forceinline__ void tps_warp_fwd_body(const float* __restrict__ U, const float* __restrict__ U2,
                                                  const float* __restrict__ coord, const float* __restrict__ T,
                                                  const float* __restrict__ move, const float* __restrict__ scal,
                                                  float* __restrict__ out, float* __restrict__ out2,
                                                  float* __restrict__ mesh, int N2, int H, int W, int Crt, int oh, int ow,
                                                  int tile_x, int b, float* sm_const, float* sm_out) {
    // sm_const: 42 floats; sm_out: 2 * WARP_TPB * C floats, 16-byte aligned
    const int Cc = (C > 0) ? C : Crt;
    // float4 tile stores: C == 3 (a tile is 96 float4, tiles start at multiples of 128 pixels = 1536 B)
    const bool vec_store = (C == 3) && ((oh * ow) % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out2)) & 15u) == 0;
    const pk::Ops o;
    SampleConstPk k;
    load_sample_const_pk(k, sm_const, coord, T, move, scal, b);
    const bool has_move = (move != nullptr);
    const bool second = (U2 != nullptr) && (b < N2);
    const float step_w = lin_step(ow), step_h = lin_step(oh);
    const pk::f2 WHf = pk::pack((float)W, (float)H);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    const int OP = oh * ow;
    const float* Ub = U + (size_t)b * H * W * Cc;
    const float* Ub2 = second ? U2 + (size_t)b * H * W * Cc : nullptr;
    float* sm_out2 = sm_out + WARP_TPB * Cc;
    const int tile0 = tile_x * (WARP_TPB * WARP_PPT);
    int i = tile0 / ow, j = tile0 - i * ow + (int)threadIdx.x - WARP_TPB;
#pragma unroll 1
    for (int it = 0; it < WARP_PPT; ++it) {
        const int base = tile0 + it * WARP_TPB;
        if (base >= OP) break;
        const int pix = base + threadIdx.x;
        j += WARP_TPB;
        if (pix < OP) {
            // row / column of the pixel: one division per CTA (above), then carried forward
            while (j >= ow) { j -= ow; ++i; }
            // (float)i, (float)j exactly, off the conversion pipe (pk::small_int_as_magic)
            const float jf = __fsub_rn(pk::small_int_as_magic(j), pk::MAGIC), if_ = __fsub_rn(pk::small_int_as_magic(i), pk::MAGIC);
            const float x_t = __fadd_rn(-1.0f, __fmul_rn(jf, step_w)), y_t = __fadd_rn(-1.0f, __fmul_rn(if_, step_h));
            const pk::f2 xy_s = sample_position_pk(o, k, has_move, x_t, y_t);
            if (mesh) {
                float x_s, y_s;
                pk::unpack(xy_s, x_s, y_s);
                reinterpret_cast<float2*>(mesh)[(size_t)b * OP + pix] = make_float2(y_s, x_s);
            }
            const BilinearPk s = bilinear_stencil_pk(o, xy_s, WHf, wmax, hmax, W, Cc);
            float wa, wb, wc, wd;
            pk::unpack(s.wab, wa, wb);
            pk::unpack(s.wcd, wc, wd);
            const pk::f2 wa2 = pk::splat(wa), wb2 = pk::splat(wb), wc2 = pk::splat(wc), wd2 = pk::splat(wd);
            if (C == 3 && second) {
                // channel c of both image sets in one pair
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const pk::f2 r = bilinear_mix_pk(o, wa2, wb2, wc2, wd2,
                                                     pk::pack(__ldg(Ub + s.oa + c), __ldg(Ub2 + s.oa + c)),
                                                     pk::pack(__ldg(Ub + s.ob + c), __ldg(Ub2 + s.ob + c)),
                                                     pk::pack(__ldg(Ub + s.oc + c), __ldg(Ub2 + s.oc + c)),
                                                     pk::pack(__ldg(Ub + s.od + c), __ldg(Ub2 + s.od + c)));
                    pk::unpack(r, sm_out[threadIdx.x * 3 + c], sm_out2[threadIdx.x * 3 + c]);
                }
            } else {
                const int n_img = second ? 2 : 1;
                for (int im = 0; im < n_img; ++im) {
                    const float* Ui = im ? Ub2 : Ub;
                    float* so = (im ? sm_out2 : sm_out) + threadIdx.x * Cc;
                    int c = 0;
                    for (; c + 1 < Cc; c += 2) {
                        const pk::f2 r = bilinear_mix_pk(o, wa2, wb2, wc2, wd2,
                                                         pk::pack(__ldg(Ui + s.oa + c), __ldg(Ui + s.oa + c + 1)),
                                                         pk::pack(__ldg(Ui + s.ob + c), __ldg(Ui + s.ob + c + 1)),
                                                         pk::pack(__ldg(Ui + s.oc + c), __ldg(Ui + s.oc + c + 1)),
                                                         pk::pack(__ldg(Ui + s.od + c), __ldg(Ui + s.od + c + 1)));
                        pk::unpack(r, so[c], so[c + 1]);
                    }
                    if (c < Cc) {
                        float o1 = __fadd_rn(__fmul_rn(wa, __ldg(Ui + s.oa + c)), __fmul_rn(wb, __ldg(Ui + s.ob + c)));
                        o1 = __fadd_rn(o1, __fmul_rn(wc, __ldg(Ui + s.oc + c)));
                        so[c] = __fadd_rn(o1, __fmul_rn(wd, __ldg(Ui + s.od + c)));
                    }
                }
            }
        }
        // coalesced write of the tile: WARP_TPB*C contiguous floats
        const int n_live = min(WARP_TPB, OP - base) * Cc;
        float* ob_ = out + ((size_t)b * OP + base) * Cc;
        float* ob2 = second ? out2 + ((size_t)b * OP + base) * Cc : nullptr;
        if (vec_store && n_live == WARP_TPB * Cc) {
            // full tile of 16-byte aligned rows.  Each warp staged its own 32 pixels (96 floats = 24 float4,
            // 384 contiguous bytes of the output): a warp-level sync is enough, no CTA barrier.
            __syncwarp();
            const int lane = threadIdx.x & 31, w4 = (threadIdx.x >> 5) * 24;   // float4 index of the warp's slice
            if (lane < 24) {
                st4_stream(ob_ + 4 * (w4 + lane), reinterpret_cast<const float4*>(sm_out)[w4 + lane]);
                if (second) st4_stream(ob2 + 4 * (w4 + lane), reinterpret_cast<const float4*>(sm_out2)[w4 + lane]);
            }
            __syncwarp();
        } else {
            __syncthreads();
            for (int e = threadIdx.x; e < n_live; e += WARP_TPB) __stcs(ob_ + e, sm_out[e]);
            if (second)
                for (int e = threadIdx.x; e < n_live; e += WARP_TPB) __stcs(ob2 + e, sm_out2[e]);
            __syncthreads();
        }
    }
}

}  // namespace ups
