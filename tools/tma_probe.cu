// Probe: 3-D u8 tiled TMA loads (non-swizzled) with tensor maps passed as a
// __grid_constant__ array (dynamic index), vs a global-memory copy of the maps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

struct TmapSet { CUtensorMap m[4]; };
__constant__ int c_w[4] = {48, 80, 112, 208};
__constant__ int c_h[4] = {8, 14, 22, 48};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ TmapSet maps, const CUtensorMap* gmaps, int mode, int idx, int c0,
                      int c1, int c2, unsigned* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int bytes = c_w[idx] * c_h[idx];
    if (threadIdx.x == 0) {
        const CUtensorMap* mp = (mode == 0) ? &maps.m[idx] : (mode == 1 ? &maps.m[1] : &gmaps[idx]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(smem)), "l"(mp), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    unsigned s = 0;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += smem[i];
    atomicAdd(out, s);
    if (threadIdx.x == 0) { out[1] = smem[0]; out[2] = smem[bytes - 1]; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#include <stdlib.h>
int main(int argc, char** argv) {
    const int F = 2, H = 360, W = 640; const size_t stride = 3 * W, fb = H * stride;
    std::vector<uint8_t> h(F * fb);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int hw[4] = {48, 80, 112, 208}, hh[4] = {8, 14, 22, 48};
    TmapSet set; memset(&set, 0, sizeof(set));
    for (int i = 0; i < 4; ++i) {
        cuuint64_t gdim[3] = {(cuuint64_t)3 * W, (cuuint64_t)H, (cuuint64_t)F};
        cuuint64_t gstr[2] = {stride, fb};
        cuuint32_t box[3] = {(cuuint32_t)hw[i], (cuuint32_t)hh[i], 1}, es[3] = {1, 1, 1};
        CUresult r = fn(&set.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) printf("encode %d -> %d\n", i, (int)r);
    }
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(set)); cudaMemcpy(gm, &set, sizeof(set), cudaMemcpyHostToDevice);
    unsigned* out; cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    int t[5] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
    {
        cudaMemset(out, 0, 16);
        probe<<<1, 128, 32768>>>(set, gm, t[0], t[1], t[2], t[3], t[4], out);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned ho[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
        unsigned s = 0;
        for (int r = 0; r < hh[t[1]]; ++r) for (int c = 0; c < hw[t[1]]; ++c) {
            int y = t[3] + r, x = t[2] + c;
            if (y >= 0 && y < H && x >= 0 && x < 3 * W) s += h[(size_t)t[4] * fb + (size_t)y * stride + x];
        }
        printf("mode %d idx %d c=(%d,%d,%d): %s sum=%u expect=%u\n", t[0], t[1], t[2], t[3], t[4], cudaGetErrorString(e), ho[0], s);
    }
    return 0;
}
