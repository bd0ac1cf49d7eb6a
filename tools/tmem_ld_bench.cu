// TMEM read-bandwidth probe (measurement aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/variants/tmem_ld_bench tools/tmem_ld_bench.cu
// Every warp of a CTA reads its own TMEM lane quadrant with tcgen05.ld.32x32b.x32 in a loop
// (DEPTH loads in flight per wait) and the CTA reports bytes per SM clock.  4 warps = one per
// quadrant (the 512-d epilogue), 8 warps = two per quadrant (the 128-d epilogue).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

template <int DEPTH>
__global__ void probe(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 1) * 256;
    uint32_t acc = 0;
    uint32_t v[DEPTH][32];
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) ld32(base + (uint32_t)(((i * DEPTH + d) & 7) * 32), v[d]);
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            ld_wait(v[d]);
            acc ^= v[d][0] ^ v[d][31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345u) sink[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}

template <int DEPTH>
static void run(int warps, int blocks, int iters) {
    unsigned long long* d_cycles; uint32_t* d_sink;
    cudaMalloc(&d_cycles, blocks * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 4);
    probe<DEPTH><<<blocks, warps * 32>>>(64, d_cycles, d_sink);
    probe<DEPTH><<<blocks, warps * 32>>>(iters, d_cycles, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    unsigned long long h[1024];
    cudaMemcpy(h, d_cycles, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int b = 0; b < blocks; ++b) mx = h[b] > mx ? h[b] : mx;
    const double bytes = (double)iters * DEPTH * 32 * 4 * 32 * warps;
    printf("{\"probe\": \"tcgen05.ld.32x32b.x32\", \"warps_per_cta\": %d, \"ctas\": %d, \"loads_in_flight\": %d, "
           "\"bytes_per_clk_per_sm\": %.1f, \"cycles_per_load\": %.1f}\n",
           warps, blocks, DEPTH, bytes / (double)mx, (double)mx / ((double)iters * DEPTH));
    cudaFree(d_cycles); cudaFree(d_sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    for (int blocks : {1, sms}) {
        for (int warps : {1, 4, 8}) {
            run<1>(warps, blocks, 20000);
            run<2>(warps, blocks, 10000);
            run<4>(warps, blocks, 5000);
        }
    }
    return 0;
}
