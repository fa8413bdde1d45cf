// Image ingest: the uint8 -> [-1, 1] fp32 normalisation the reference's data pipeline applies on the
// host right before feeding the views (cub/code/data/data.py:134,152 and pennaction/code/data/data.py:
// 134,152:  o.astype(np.float32) * 2.0 / 255.0 - 1.0), done on the device so that a step's views cross
// PCIe as bytes (1/4 of the fp32 traffic).  Elementwise, HBM-bound: 1 byte read + 4 bytes written.
#include "common.cuh"

namespace ups {

// numpy evaluates  fl(fl(fl((float)u * 2) / 255) - 1)  in fp32 (python scalars do not upcast): the 256
// possible results are built once per CTA with correctly rounded IEEE operations (no reciprocal tricks)
// and looked up per byte.
__device__ __forceinline__ float normalize_u8_canon(unsigned u) {
    return __fsub_rn(__fdiv_rn(__fmul_rn((float)u, 2.0f), 255.0f), 1.0f);
}

constexpr int ING_TPB = 256;

__global__ void __launch_bounds__(ING_TPB) views_u8_to_f32_kernel(const unsigned char* __restrict__ src,
                                                                  float* __restrict__ dst, long long n) {
    __shared__ float lut[256];
    lut[threadIdx.x] = normalize_u8_canon(threadIdx.x);
    __syncthreads();
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * ING_TPB;
    const uint4* src16 = reinterpret_cast<const uint4*>(src);
    for (long long i = (long long)blockIdx.x * ING_TPB + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldcs(src16 + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st4_stream(dst + 16 * i + 4 * q, make_float4(lut[w[q] & 0xFF], lut[(w[q] >> 8) & 0xFF],
                                                         lut[(w[q] >> 16) & 0xFF], lut[w[q] >> 24]));
    }
    // tail (n % 16 bytes), first CTA
    if (blockIdx.x == 0) {
        const long long t = (n16 << 4) + threadIdx.x;
        if (t < n) dst[t] = lut[src[t]];
    }
}

// int64 part labels (tf.argmax's dtype, model.py:447,470) -> one byte per pixel for the read-back of the label map:
// n_parts <= 255 in every shipped configuration, and eight bytes per pixel is 97 % of the step's device->host traffic.
__global__ void __launch_bounds__(ING_TPB) labels_i64_to_u8_kernel(const long long* __restrict__ src,
                                                                   unsigned char* __restrict__ dst, long long n) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * ING_TPB;
    for (long long i = (long long)blockIdx.x * ING_TPB + threadIdx.x; i < n4; i += stride) {
        const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(src) + 2 * i);
        const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(src) + 2 * i + 1);
        const unsigned w = (unsigned)(a.x & 0xFF) | ((unsigned)(a.y & 0xFF) << 8) | ((unsigned)(b.x & 0xFF) << 16) |
                           ((unsigned)(b.y & 0xFF) << 24);
        reinterpret_cast<unsigned*>(dst)[i] = w;
    }
    if (blockIdx.x == 0) {
        const long long t = (n4 << 2) + threadIdx.x;
        if (t < n) dst[t] = (unsigned char)(src[t] & 0xFF);
    }
}

}  // namespace ups

using namespace ups;

extern "C" int ups_labels_i64_to_u8(const long long* labels, unsigned char* out, long long n, void* stream) {
    UPS_REQUIRE(n >= 0, "labels_i64_to_u8: n=%lld", n);
    if (n == 0) return UPS_OK;
    UPS_REQUIRE(labels && out, "labels_i64_to_u8: null pointer");
    UPS_REQUIRE(aligned16(labels) && (reinterpret_cast<uintptr_t>(out) & 3u) == 0, "labels_i64_to_u8: alignment");
    long long ctas = cdiv((n >> 2) > 0 ? (n >> 2) : 1, ING_TPB);
    const long long cap = (long long)NUM_SMS * 16;
    if (ctas > cap) ctas = cap;
    labels_i64_to_u8_kernel<<<(unsigned)ctas, ING_TPB, 0, as_stream(stream)>>>(labels, out, n);
    return after_launch("labels_i64_to_u8_kernel");
}

extern "C" int ups_views_u8_to_f32(const unsigned char* src, float* dst, long long n, void* stream) {
    UPS_REQUIRE(n >= 0, "views_u8_to_f32: n=%lld", n);
    if (n == 0) return UPS_OK;
    UPS_REQUIRE(src && dst, "views_u8_to_f32: null pointer");
    UPS_REQUIRE(aligned16(src) && aligned16(dst), "views_u8_to_f32: 16-byte alignment");
    const long long n16 = n >> 4;
    long long ctas = cdiv(n16 > 0 ? n16 : 1, ING_TPB);
    const long long cap = (long long)NUM_SMS * 16;   // grid-stride above 16 CTAs per SM
    if (ctas > cap) ctas = cap;
    views_u8_to_f32_kernel<<<(unsigned)ctas, ING_TPB, 0, as_stream(stream)>>>(src, dst, n);
    return after_launch("views_u8_to_f32_kernel");
}
