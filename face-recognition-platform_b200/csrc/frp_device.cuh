// Device-side helpers: mbarrier / bulk-copy (TMA) PTX wrappers, warp sorting
// network, multi-value warp reduction.  sm_100a.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace frp {

#define FRP_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP), completion
// signalled on an mbarrier.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- programmatic dependent launch ------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while the
// previous kernel of the stream is still running; it must not touch that kernel's results before
// griddep_wait() (returns at once when the launch was an ordinary one).  griddep_launch() in the
// earlier kernel lets the next one be scheduled as soon as every CTA has issued it.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- warp-level bitonic sort, descending, 32*E u64 keys ----------------------
// element index i = j*32 + lane  (j < E).  E power of two.
template <int E>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&k)[E], int lane) {
    constexpr int N = 32 * E;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
            if (stride >= 32) {
                const int js = stride >> 5;
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    if ((j & js) == 0) {
                        const int i = j * 32 + lane;
                        const bool desc = ((i & size) == 0);
                        uint64_t a = k[j], b = k[j | js];
                        const bool swap = desc ? (a < b) : (a > b);
                        if (swap) { k[j] = b; k[j | js] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const int i = j * 32 + lane;
                    const bool desc = ((i & size) == 0);
                    const bool lower = ((lane & stride) == 0);
                    const uint64_t other = __shfl_xor_sync(FRP_FULL_MASK, k[j], stride);
                    const bool keep_max = (lower == desc);
                    const uint64_t mx = k[j] > other ? k[j] : other;
                    const uint64_t mn = k[j] > other ? other : k[j];
                    k[j] = keep_max ? mx : mn;
                }
            }
        }
    }
}

// ---- multi-value warp reduction ------------------------------------------------
// Sums NV (power of two <= 32) per-lane partial values across the warp; on
// return v[0] of lane l holds the total of value index (l & (NV-1)).
// Shuffles: NV*log2(32/NV) + NV-1  (instead of 5*NV).
// W (power of two <= 32, >= NV) = width of the lane group that is reduced: groups of W
// consecutive lanes are summed independently (sub-warp rows for short embeddings).
template <int NV, int W = 32>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[NV], int lane) {
    static_assert(NV <= W && W <= 32, "warp_multi_reduce: NV <= W <= 32");
    // full reductions over the lane bits above log2(NV)
#pragma unroll
    for (int s = W / 2; s >= NV; s >>= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(FRP_FULL_MASK, v[i], s);
    }
    // halving steps
#pragma unroll
    for (int s = (NV >> 1); s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = upper ? v[i] : v[i + s];
            const float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(FRP_FULL_MASK, send, s);
        }
    }
    return v[0];
}

__device__ __forceinline__ double warp_sum_f64(double x) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) x += __shfl_xor_sync(FRP_FULL_MASK, x, s);
    return x;
}

}  // namespace frp
