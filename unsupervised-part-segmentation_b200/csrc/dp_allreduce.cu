// Gradient all-reduce of the data-parallel step wrapper over NVLink 5 / NVSwitch peer memory.
//
// The reference trains on one GPU and has no collective; its IMM baseline averages tower gradients on
// the host (baselines/imm/imm/train/cnn_train_multi.py:75-118: mean over towers of every (grad, var)).
// Here the same mean is ONE kernel per bucket that runs beside the path's backward kernels:
//
//   every rank owns the slice [rank*n/world, (rank+1)*n/world) of the bucket
//   * multicast variant (NVSwitch in-switch reduction, "NVLS"): one `multimem.ld_reduce.add.v4.f32` on the
//     multicast address fetches the SUM over all ranks of 16 bytes (the switch adds the 8 replies), the
//     1/world of the mean is applied in registers and one `multimem.st.v4.f32` broadcasts the result into
//     every rank's buffer: the gradient crosses each NVLink once per direction and no separate scale pass
//     touches HBM;
//   * peer variant (no multicast object, e.g. two GPUs without a switch): the owner loads its slice from
//     every rank's buffer through the peer mappings (rank 0..world-1, a fixed order: deterministic),
//     scales, and stores the result into every rank's buffer.
//
// Ranks synchronise with two flag barriers on a caller-owned, zero-initialised signal pad in peer memory
// (slot [cta][src rank] on the destination rank, compare-and-swap 0->1 by the sender with release, 1->0 by
// the receiver with acquire: self-resetting, no epoch counter, no host involvement).
//
// Co-residency is the design constraint: the kernel runs BESIDE the path's backward.  K4 is a persistent grid of one
// 448-thread CTA per SM holding 120 registers per thread (53 760 of the SM's 65 536) and ~200 KB of shared memory; a
// CTA of this kernel is 128 threads x <= 40 registers (5 120) and no shared memory, so it fits into what K4 leaves
// free on every SM and starts at once instead of waiting for K4's CTAs to retire (a 512-thread CTA did not fit:
// measured at N=2, its first bucket only started once K4 had drained and 0.2 ms of the reduction was exposed).
#include "common.cuh"

namespace ups {
namespace dp {

constexpr int MAX_WORLD = 16;
constexpr int TPB = 128;
constexpr int MIN_CTAS_PER_SM = 12;  // caps ptxas at 65536 / (128*12) = 42 registers per thread
constexpr int UNROLL = 4;
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;  // a dead peer must not hang the GPU

struct Peers {
    float* buf[MAX_WORLD];
    unsigned int* sig[MAX_WORLD];
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int cas_release_sys(unsigned int* addr, unsigned int cmp, unsigned int val) {
    unsigned int old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int cas_acquire_sys(unsigned int* addr, unsigned int cmp, unsigned int val) {
    unsigned int old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ void put_signal(unsigned int* addr) {
    const unsigned long long t0 = globaltimer_ns();
    while (cas_release_sys(addr, 0u, 1u) != 0u)
        if (globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) __trap();
}
__device__ __forceinline__ void wait_signal(unsigned int* addr) {
    const unsigned long long t0 = globaltimer_ns();
    while (cas_acquire_sys(addr, 1u, 0u) != 1u)
        if (globaltimer_ns() - t0 > SPIN_TIMEOUT_NS) __trap();
}

// Barrier of CTA `blockIdx.x` with the same CTA of every other rank.  Everything this CTA wrote before the call
// (to any rank) is visible to every rank's CTA after it.
__device__ __forceinline__ void rank_barrier(const Peers& pr, int rank, int world) {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int t = threadIdx.x;
        put_signal(pr.sig[t] + (size_t)blockIdx.x * world + rank);   // "rank has arrived" on rank t
        wait_signal(pr.sig[rank] + (size_t)blockIdx.x * world + t);  // rank t has arrived here
    }
    __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;   // peer memory is never cached in the local L2; do not allocate in L1 either (read once)
    asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(TPB, MIN_CTAS_PER_SM) dp_allreduce_kernel(const Peers pr, float* __restrict__ mc, int rank, int world,
                                                            long long n4, float scale) {
    rank_barrier(pr, rank, world);   // every rank's bucket is complete (its producers precede this kernel in stream order)
    const long long per = (n4 + world - 1) / world;
    const long long lo = (long long)rank * per;
    const long long hi = lo + per < n4 ? lo + per : n4;
    const long long stride = (long long)gridDim.x * TPB;
    for (long long i0 = lo + (long long)blockIdx.x * TPB + threadIdx.x; i0 < hi; i0 += stride * UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                if (MC) {
                    v[u] = mc_ld_reduce(mc + 4 * i);
                } else {
                    v[u] = ld_peer(pr.buf[0] + 4 * i);
                    for (int p = 1; p < world; ++p) {
                        const float4 t = ld_peer(pr.buf[p] + 4 * i);
                        v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                const float4 o = make_float4(v[u].x * scale, v[u].y * scale, v[u].z * scale, v[u].w * scale);
                if (MC) {
                    mc_st(mc + 4 * i, o);
                } else {
                    for (int p = 0; p < world; ++p) st_peer(pr.buf[p] + 4 * i, o);
                }
            }
        }
    }
    rank_barrier(pr, rank, world);   // every rank's slice has landed in this rank's bucket
}

}  // namespace dp
}  // namespace ups

using namespace ups;

extern "C" size_t ups_dp_allreduce_signal_bytes(int world, int n_ctas) {
    if (world <= 0 || n_ctas <= 0) return 0;
    return (size_t)world * (size_t)n_ctas * sizeof(unsigned int);
}

extern "C" int ups_dp_allreduce(void* const* peer_bufs, void* mc_buf, void* const* peer_signals, int rank, int world,
                                long long n_floats, float scale, int n_ctas, void* stream) {
    UPS_REQUIRE(peer_bufs && peer_signals, "dp_allreduce: null pointer table");
    UPS_REQUIRE(world >= 1 && world <= dp::MAX_WORLD, "dp_allreduce: world=%d out of range [1,%d]", world, dp::MAX_WORLD);
    UPS_REQUIRE(rank >= 0 && rank < world, "dp_allreduce: rank=%d not in [0,%d)", rank, world);
    UPS_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "dp_allreduce: n_floats=%lld must be a multiple of 4", n_floats);
    UPS_REQUIRE(n_ctas >= 1 && n_ctas <= 4 * NUM_SMS, "dp_allreduce: n_ctas=%d out of range", n_ctas);
    dp::Peers pr;
    for (int p = 0; p < dp::MAX_WORLD; ++p) { pr.buf[p] = nullptr; pr.sig[p] = nullptr; }
    for (int p = 0; p < world; ++p) {
        UPS_REQUIRE(peer_bufs[p] && peer_signals[p], "dp_allreduce: null mapping of rank %d", p);
        UPS_REQUIRE(aligned16(peer_bufs[p]), "dp_allreduce: 16-byte alignment of rank %d's bucket", p);
        pr.buf[p] = static_cast<float*>(peer_bufs[p]);
        pr.sig[p] = static_cast<unsigned int*>(peer_signals[p]);
    }
    UPS_REQUIRE(!mc_buf || aligned16(mc_buf), "dp_allreduce: 16-byte alignment of the multicast mapping");
    if (n_floats == 0) return UPS_OK;
    cudaStream_t s = as_stream(stream);
    static const bool carveout = []() {
        prefer_max_shared_carveout(dp::dp_allreduce_kernel<true>);
        prefer_max_shared_carveout(dp::dp_allreduce_kernel<false>);
        return true;
    }();
    (void)carveout;
    if (mc_buf)
        dp::dp_allreduce_kernel<true><<<n_ctas, dp::TPB, 0, s>>>(pr, static_cast<float*>(mc_buf), rank, world, n_floats / 4, scale);
    else
        dp::dp_allreduce_kernel<false><<<n_ctas, dp::TPB, 0, s>>>(pr, nullptr, rank, world, n_floats / 4, scale);
    return after_launch("dp_allreduce_kernel");
}
