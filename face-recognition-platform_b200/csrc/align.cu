// K1/K2 — batched 5-landmark alignment: similarity estimate (Umeyama onto
// arcface_dst), cv2.warpAffine-exact bilinear warp to 112x112 (OpenCV's fixed-point
// scheme: AB_BITS=10, INTER_BITS=5, 15-bit weights, BORDER_CONSTANT 0), fused
// (v-127.5)/127.5, BGR->RGB, HWC u8 -> NCHW fp16.
//
// Replaces insightface face_align.norm_crop + cv2.dnn.blobFromImages (third-party
// code reached from backend/app/utils/deepfake_utils.py:138).
//
// One warp per work item (half a crop = 49 output tiles of 16x8 px).  Per tile the
// warp stages the source bounding box of the tile into its private shared-memory
// buffer with 128-bit cp.async copies on 16-byte aligned row segments (zero-filled
// outside the frame = BORDER_CONSTANT 0), double buffered against the blend of the
// previous tile; each lane then produces 4 consecutive output pixels from shared
// memory with integer dp2a blends, maps them through a 256-entry fp16 table and
// writes 8-byte segments of the three output planes.
#include <math.h>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

namespace {

constexpr int kOut = 112;
constexpr int kTileW = 16, kTileH = 8;
constexpr int kTilesX = kOut / kTileW;              // 7
constexpr int kTilesY = kOut / kTileH;              // 14
constexpr int kTilesPerCrop = kTilesX * kTilesY;    // 98
constexpr int kItemsPerCrop = 2;
constexpr int kTilesPerItem = kTilesPerCrop / kItemsPerCrop;   // 49
constexpr int kWarps = 8;
constexpr int kTileBuf = 12 * 1024;                 // bytes per staging buffer
constexpr int kTabInts = 4 * kOut;                  // adelta, bdelta, X0, Y0
constexpr int kWarpSmem = kTabInts * 4 + 2 * kTileBuf + 32;

struct AlignArgs {
    const uint8_t* frames;
    int F, H, W;
    size_t stride, frame_bytes;
    const int32_t* frame_idx;
    const float* lmk;     // [n][5][2] or null
    const double* M;      // [n][6] or null
    int n;
    __half* out_blob;     // [n][3][112][112]
    uint8_t* out_aligned; // [n][112][112][3] or null
    int* counter;         // dynamic work counter
    int aligned16;        // frames base and stride are 16-byte aligned
};

__constant__ double c_dst[10] = {
    (double)38.2946f, (double)51.6963f, (double)73.5318f, (double)51.5014f, (double)56.0252f,
    (double)71.7366f, (double)41.5493f, (double)92.3655f, (double)70.7299f, (double)92.2041f};

// closed form of skimage _umeyama for 2-D similarity (proper rotation):
// scale*R = [[a,-b],[b,a]] / var(src),  a = sum(dst.src)/n, b = sum(dst x src)/n
__device__ void estimate_norm_dev(const float* lm, double* M) {
    double sx = 0, sy = 0, dx = 0, dy = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        sx += (double)lm[2 * i]; sy += (double)lm[2 * i + 1];
        dx += c_dst[2 * i];      dy += c_dst[2 * i + 1];
    }
    sx /= 5.0; sy /= 5.0; dx /= 5.0; dy /= 5.0;
    double a = 0, b = 0, var = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double px = (double)lm[2 * i] - sx, py = (double)lm[2 * i + 1] - sy;
        const double qx = c_dst[2 * i] - dx, qy = c_dst[2 * i + 1] - dy;
        a += qx * px + qy * py;
        b += qy * px - qx * py;
        var += px * px + py * py;
    }
    a /= 5.0; b /= 5.0; var /= 5.0;
    const double m00 = a / var, m10 = b / var;
    M[0] = m00; M[1] = -m10; M[2] = dx - (m00 * sx - m10 * sy);
    M[3] = m10; M[4] = m00;  M[5] = dy - (m10 * sx + m00 * sy);
}

// cv::warpAffine's inversion of the forward matrix (double, operation order kept,
// no fused multiply-add so the rounding matches the host library)
__device__ void invert_affine_dev(const double* M, double* iM) {
    double D = __dsub_rn(__dmul_rn(M[0], M[4]), __dmul_rn(M[1], M[3]));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(M[4], D), A22 = __dmul_rn(M[0], D);
    const double m1 = __dmul_rn(M[1], -D), m3 = __dmul_rn(M[3], -D);
    const double b1 = __dsub_rn(__dmul_rn(-A11, M[2]), __dmul_rn(m1, M[5]));
    const double b2 = __dsub_rn(__dmul_rn(-m3, M[2]), __dmul_rn(A22, M[5]));
    iM[0] = A11; iM[1] = m1; iM[2] = b1;
    iM[3] = m3;  iM[4] = A22; iM[5] = b2;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct TileMeta {
    int bo;        // byte offset (in the source row) of the first staged byte; multiple of 16
    int sy0;       // first staged source row
    int pitch;     // staged row pitch in bytes (multiple of 16)
    int direct;    // 1: footprint did not fit -> read taps from global memory
};

__device__ __forceinline__ int floor_div16(int v) { return v >> 4; }   // arithmetic shift = floor

// stage the source footprint of output tile (x0,y0) into buf
__device__ __forceinline__ TileMeta stage_tile(const AlignArgs& a, const uint8_t* frame, const int* tab,
                                               int x0, int y0, uint8_t* buf, int lane) {
    const int* adelta = tab;
    const int* bdelta = tab + kOut;
    const int* X0 = tab + 2 * kOut;
    const int* Y0 = tab + 3 * kOut;
    const int xa = adelta[x0], xb = adelta[x0 + kTileW - 1];
    const int ya = bdelta[x0], yb = bdelta[x0 + kTileW - 1];
    const int Xa = X0[y0], Xb = X0[y0 + kTileH - 1];
    const int Ya = Y0[y0], Yb = Y0[y0 + kTileH - 1];
    const int sx_min = (min(Xa, Xb) + min(xa, xb)) >> 10;
    const int sx_max = (max(Xa, Xb) + max(xa, xb)) >> 10;
    const int sy_min = (min(Ya, Yb) + min(ya, yb)) >> 10;
    const int sy_max = (max(Ya, Yb) + max(ya, yb)) >> 10;
    TileMeta m;
    m.bo = floor_div16(3 * sx_min) * 16;
    m.sy0 = sy_min;
    const int byte_hi = 3 * (sx_max + 2);
    const int nchunk = (byte_hi - m.bo + 15) >> 4;
    const int nrows = sy_max + 2 - sy_min;
    m.pitch = nchunk * 16;
    // coordinates near the int16 saturation of cv::remap's xy map are handled tap by tap
    const bool far = sx_min < -30000 || sx_max > 30000 || sy_min < -30000 || sy_max > 30000;
    m.direct = (!a.aligned16) || far || (m.pitch * nrows + 16 > kTileBuf) || nchunk > 64 || nrows <= 0;
    if (m.direct) return m;
    int sh = 0;
    while ((1 << sh) < nchunk) ++sh;
    const int total = nrows << sh;
    const int row_bytes = 3 * a.W;
    for (int i = lane; i < total; i += 32) {
        const int r = i >> sh;
        const int c = i & ((1 << sh) - 1);
        if (c < nchunk) {
            const int sy = sy_min + r;
            const int gb = m.bo + 16 * c;
            int valid = 0;
            if (sy >= 0 && sy < a.H && gb >= 0) valid = min(16, max(0, row_bytes - gb));
            const uint8_t* src = valid ? (frame + (size_t)sy * a.stride + gb) : frame;
            cp_async16(buf + r * m.pitch + 16 * c, src, valid);
        }
    }
    return m;
}

__device__ __forceinline__ unsigned tap_global(const AlignArgs& a, const uint8_t* frame, int sx, int sy) {
    // returns B | G<<8 | R<<16 of pixel (sx,sy), 0 outside the frame
    if (sx < 0 || sy < 0 || sx >= a.W || sy >= a.H) return 0u;
    const uint8_t* p = frame + (size_t)sy * a.stride + 3 * sx;
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
}

__global__ void __launch_bounds__(kWarps * 32, 1) align_kernel(const AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __half s_lut[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (v - 127.5) * float(1/127.5) in fp32, as cv2.dnn.blobFromImages computes it
    if (threadIdx.x < 256) {
        const float f = __fmul_rn(__fsub_rn((float)threadIdx.x, 127.5f), (float)(1.0 / 127.5));
        s_lut[threadIdx.x] = __float2half_rn(f);
    }
    __syncthreads();
    uint8_t* wsm = smem + (size_t)warp * kWarpSmem;
    int* tab = reinterpret_cast<int*>(wsm);
    uint8_t* bufs[2] = {wsm + kTabInts * 4, wsm + kTabInts * 4 + kTileBuf + 16};
    const unsigned short* lut = reinterpret_cast<const unsigned short*>(s_lut);

    const int nitems = a.n * kItemsPerCrop;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1);
        item = __shfl_sync(FRP_FULL_MASK, item, 0);
        if (item >= nitems) break;
        const int crop = item / kItemsPerCrop;
        const int tile_begin = (item % kItemsPerCrop) * kTilesPerItem;

        // ---- per-crop setup (every lane computes the same matrix) ----
        double M[6], iM[6];
        if (a.M) {
#pragma unroll
            for (int i = 0; i < 6; ++i) M[i] = a.M[(size_t)crop * 6 + i];
        } else {
            float lm[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) lm[i] = a.lmk[(size_t)crop * 10 + i];
            estimate_norm_dev(lm, M);
        }
        invert_affine_dev(M, iM);
        bool finite = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) finite = finite && (fabs(iM[i]) < 1e12);
        if (!finite) { iM[0] = iM[1] = iM[3] = iM[4] = 0.0; iM[2] = iM[5] = -1e6; }
        int fi = a.frame_idx[crop];
        fi = min(max(fi, 0), a.F - 1);
        const uint8_t* frame = a.frames + (size_t)fi * a.frame_bytes;
        __syncwarp();
        for (int x = lane; x < kOut; x += 32) {
            const double xd = (double)x;
            tab[x] = __double2int_rn(__dmul_rn(__dmul_rn(iM[0], xd), 1024.0));
            tab[kOut + x] = __double2int_rn(__dmul_rn(__dmul_rn(iM[3], xd), 1024.0));
            tab[2 * kOut + x] =
                __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[1], xd), iM[2]), 1024.0)) + 16;
            tab[3 * kOut + x] =
                __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[4], xd), iM[5]), 1024.0)) + 16;
        }
        __syncwarp();

        __half* oblob = a.out_blob + (size_t)crop * 3 * kOut * kOut;
        uint8_t* oal = a.out_aligned ? a.out_aligned + (size_t)crop * kOut * kOut * 3 : nullptr;

        TileMeta cur, nxt;
        {
            const int t = tile_begin;
            cur = stage_tile(a, frame, tab, (t % kTilesX) * kTileW, (t / kTilesX) * kTileH, bufs[0], lane);
            cp_async_commit();
        }
        for (int tt = 0; tt < kTilesPerItem; ++tt) {
            const int t = tile_begin + tt;
            if (tt + 1 < kTilesPerItem) {
                const int tn = t + 1;
                nxt = stage_tile(a, frame, tab, (tn % kTilesX) * kTileW, (tn / kTilesX) * kTileH,
                                 bufs[(tt + 1) & 1], lane);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();

            // ---- blend 4 consecutive output pixels per lane ----
            const uint8_t* buf = bufs[tt & 1];
            const int x0 = (t % kTilesX) * kTileW, y0 = (t / kTilesX) * kTileH;
            const int y = y0 + (lane >> 2);
            const int xs = x0 + (lane & 3) * 4;
            const int Xy = tab[2 * kOut + y], Yy = tab[3 * kOut + y];
            const int4 ad = *reinterpret_cast<const int4*>(tab + xs);
            const int4 bd = *reinterpret_cast<const int4*>(tab + kOut + xs);
            const int adv[4] = {ad.x, ad.y, ad.z, ad.w};
            const int bdv[4] = {bd.x, bd.y, bd.z, bd.w};
            unsigned vB[4], vG[4], vR[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int X = (Xy + adv[j]) >> 5, Y = (Yy + bdv[j]) >> 5;
                const int sx = max(-32768, min(32767, X >> 5));
                const int sy = max(-32768, min(32767, Y >> 5));
                const unsigned fx = X & 31, fy = Y & 31;
                unsigned p0lo, p0hi, p1lo, p1hi;   // row0: B0 G0 R0 B1 | G1 R1 . .  ; row1 likewise
                if (!cur.direct) {
                    const int addr = (sy - cur.sy0) * cur.pitch + (3 * sx - cur.bo);
                    const unsigned* w0 = reinterpret_cast<const unsigned*>(buf + (addr & ~3));
                    const unsigned* w1 = reinterpret_cast<const unsigned*>(buf + ((addr + cur.pitch) & ~3));
                    const unsigned shf = (addr & 3) * 8;
                    const unsigned a0 = w0[0], a1 = w0[1], a2 = w0[2];
                    const unsigned b0 = w1[0], b1 = w1[1], b2 = w1[2];
                    p0lo = __funnelshift_r(a0, a1, shf); p0hi = __funnelshift_r(a1, a2, shf);
                    p1lo = __funnelshift_r(b0, b1, shf); p1hi = __funnelshift_r(b1, b2, shf);
                } else {
                    const unsigned t00 = tap_global(a, frame, sx, sy), t01 = tap_global(a, frame, sx + 1, sy);
                    const unsigned t10 = tap_global(a, frame, sx, sy + 1), t11 = tap_global(a, frame, sx + 1, sy + 1);
                    p0lo = t00 | (t01 << 24); p0hi = t01 >> 8;
                    p1lo = t10 | (t11 << 24); p1hi = t11 >> 8;
                }
                // (B0,B1,G0,G1) and (R0,R1) byte pairs for the two 16-bit x 8-bit dot products
                const unsigned bg0 = __byte_perm(p0lo, p0hi, 0x4130), rr0 = __byte_perm(p0lo, p0hi, 0x0052);
                const unsigned bg1 = __byte_perm(p1lo, p1hi, 0x4130), rr1 = __byte_perm(p1lo, p1hi, 0x0052);
                // weights: w00=(32-fx)(32-fy)*32, w01=fx(32-fy)*32, w10=(32-fx)fy*32, w11=fx*fy*32
                const unsigned P = fx * 0xFFFFu + 32u;          // (32-fx) | fx<<16
                const unsigned W0 = ((32u - fy) * P) << 5;      // w00 | w01<<16
                const unsigned W1 = (fy * P) << 5;              // w10 | w11<<16
                unsigned accB = __dp2a_lo(W0, bg0, 16384u); accB = __dp2a_lo(W1, bg1, accB);
                unsigned accG = __dp2a_hi(W0, bg0, 16384u); accG = __dp2a_hi(W1, bg1, accG);
                unsigned accR = __dp2a_lo(W0, rr0, 16384u); accR = __dp2a_lo(W1, rr1, accR);
                vB[j] = accB >> 15; vG[j] = accG >> 15; vR[j] = accR >> 15;
            }
            // ---- normalise through the table, BGR->RGB planes, 8-byte stores ----
            const size_t o = (size_t)y * kOut + xs;
            uint2 r, g, b;
            r.x = (unsigned)lut[vR[0]] | ((unsigned)lut[vR[1]] << 16);
            r.y = (unsigned)lut[vR[2]] | ((unsigned)lut[vR[3]] << 16);
            g.x = (unsigned)lut[vG[0]] | ((unsigned)lut[vG[1]] << 16);
            g.y = (unsigned)lut[vG[2]] | ((unsigned)lut[vG[3]] << 16);
            b.x = (unsigned)lut[vB[0]] | ((unsigned)lut[vB[1]] << 16);
            b.y = (unsigned)lut[vB[2]] | ((unsigned)lut[vB[3]] << 16);
            *reinterpret_cast<uint2*>(oblob + o) = r;
            *reinterpret_cast<uint2*>(oblob + (size_t)kOut * kOut + o) = g;
            *reinterpret_cast<uint2*>(oblob + (size_t)2 * kOut * kOut + o) = b;
            if (oal) {
                uint8_t* p = oal + o * 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    p[3 * j] = (uint8_t)vB[j]; p[3 * j + 1] = (uint8_t)vG[j]; p[3 * j + 2] = (uint8_t)vR[j];
                }
            }
            __syncwarp();
            cur = nxt;
        }
    }
}

__global__ void estimate_norm_kernel(const float* __restrict__ lmk, int n, double* __restrict__ M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lm[10];
    for (int j = 0; j < 10; ++j) lm[j] = lmk[(size_t)i * 10 + j];
    double m[6];
    estimate_norm_dev(lm, m);
    for (int j = 0; j < 6; ++j) M[(size_t)i * 6 + j] = m[j];
}

std::atomic<int> g_align_profile{0};
thread_local cudaEvent_t t_align_ev[2] = {};
thread_local bool t_align_valid = false;

int align_device_impl(int device, const uint8_t* frames, int F, int H, int W, size_t stride,
                      const int32_t* frame_idx, const float* lmk, const double* M, int n,
                      uint16_t* out_blob, uint8_t* out_aligned, cudaStream_t st) {
    if (n == 0) return FRP_OK;
    // device attributes are cached: cudaGetDeviceProperties costs milliseconds per call
    static int s_major[16] = {}, s_sms[16] = {};
    const int di = (device >= 0 && device < 16) ? device : 0;
    if (!s_sms[di]) {
        int mj = 0, sms = 0;
        FRP_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, device));
        FRP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        s_major[di] = mj;
        s_sms[di] = sms;
    }
    if (s_major[di] != 10) {
        set_error("device %d is sm_%dx; built for sm_100a only (no fallback)", device, s_major[di]);
        return FRP_E_NODEVICE;
    }
    const bool prof = g_align_profile.load(std::memory_order_relaxed) != 0;
    int* counter = nullptr;
    FRP_CUDA(cudaMallocAsync((void**)&counter, sizeof(int), st));
    FRP_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    AlignArgs a;
    a.frames = frames; a.F = F; a.H = H; a.W = W; a.stride = stride;
    a.frame_bytes = (size_t)H * stride;
    a.frame_idx = frame_idx; a.lmk = lmk; a.M = M; a.n = n;
    a.out_blob = reinterpret_cast<__half*>(out_blob);
    a.out_aligned = out_aligned;
    a.counter = counter;
    a.aligned16 = (((uintptr_t)frames & 15) == 0 && (stride & 15) == 0 && (a.frame_bytes & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)kWarps * kWarpSmem;
    static thread_local bool configured = false;
    if (!configured) {
        FRP_CUDA(cudaFuncSetAttribute(align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int items = n * kItemsPerCrop;
    int grid = (items + kWarps - 1) / kWarps;
    if (grid > s_sms[di]) grid = s_sms[di];
    if (prof) {
        if (!t_align_ev[0]) { cudaEventCreate(&t_align_ev[0]); cudaEventCreate(&t_align_ev[1]); }
        cudaEventRecord(t_align_ev[0], st);
    }
    align_kernel<<<grid, kWarps * 32, smem, st>>>(a);
    FRP_CHECK_LAUNCH();
    if (prof) { cudaEventRecord(t_align_ev[1], st); t_align_valid = true; }
    FRP_CUDA(cudaFreeAsync(counter, st));
    return FRP_OK;
}

}  // namespace
}  // namespace frp

using namespace frp;

extern "C" {

int frp_profile_align_enable(int on) {
    g_align_profile.store(on ? 1 : 0);
    return FRP_OK;
}

int frp_profile_align_last(float* kernel_ms) {
    if (!t_align_valid || !kernel_ms) { set_error("frp_profile_align_last: no profiled call on this thread"); return FRP_E_INVALID; }
    FRP_CUDA(cudaEventSynchronize(t_align_ev[1]));
    FRP_CUDA(cudaEventElapsedTime(kernel_ms, t_align_ev[0], t_align_ev[1]));
    return FRP_OK;
}

int frp_align_device(int device, const uint8_t* frames_dev, int F, int H, int W, size_t stride,
                     const int32_t* frame_idx_dev, const float* lmk_dev, const double* M_dev, int n,
                     uint16_t* out_blob_dev, uint8_t* out_aligned_dev, void* stream) {
    if (n < 0 || F <= 0 || H <= 0 || W <= 0 || W > 32767 || H > 32767 || stride < (size_t)3 * W ||
        !frames_dev || !frame_idx_dev || (!lmk_dev && !M_dev) || !out_blob_dev) {
        set_error("frp_align_device: invalid argument");
        return FRP_E_INVALID;
    }
    int ndev = 0;
    int rc = frp_device_count(&ndev);
    if (rc) return rc;
    DeviceGuard dg(device);
    return align_device_impl(device, frames_dev, F, H, W, stride, frame_idx_dev, lmk_dev, M_dev, n,
                             out_blob_dev, out_aligned_dev, (cudaStream_t)stream);
}

int frp_align(int device, const uint8_t* frames, int F, int H, int W, size_t stride,
              const int32_t* frame_idx, const float* lmk, const double* M, int n, uint16_t* out_blob,
              uint8_t* out_aligned) {
    if (n < 0 || F <= 0 || H <= 0 || W <= 0 || stride < (size_t)3 * W || !frames || !frame_idx ||
        (!lmk && !M) || !out_blob) {
        set_error("frp_align: invalid argument");
        return FRP_E_INVALID;
    }
    if (n == 0) return FRP_OK;
    int ndev = 0;
    int rc = frp_device_count(&ndev);
    if (rc) return rc;
    DeviceGuard dg(device);
    // device copies keep a 16-byte aligned pitch so the staged path is always taken
    const size_t dstride = (stride + 15) & ~(size_t)15;
    uint8_t* d_frames = nullptr; int32_t* d_fi = nullptr; float* d_lmk = nullptr; double* d_M = nullptr;
    __half* d_blob = nullptr; uint8_t* d_al = nullptr;
    cudaStream_t st = 0;
    const size_t blob_elems = (size_t)n * 3 * kOut * kOut;
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(d_frames); cudaFree(d_fi); cudaFree(d_lmk); cudaFree(d_M); cudaFree(d_blob); cudaFree(d_al);
    };
#define FRP_TRY(x) if (e == cudaSuccess) e = (x)
    FRP_TRY(cudaMalloc((void**)&d_frames, (size_t)F * H * dstride));
    FRP_TRY(cudaMemcpy2D(d_frames, dstride, frames, stride, (size_t)3 * W, (size_t)F * H, cudaMemcpyHostToDevice));
    FRP_TRY(cudaMalloc((void**)&d_fi, (size_t)n * sizeof(int32_t)));
    FRP_TRY(cudaMemcpy(d_fi, frame_idx, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (M) {
        FRP_TRY(cudaMalloc((void**)&d_M, (size_t)n * 6 * sizeof(double)));
        FRP_TRY(cudaMemcpy(d_M, M, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        FRP_TRY(cudaMalloc((void**)&d_lmk, (size_t)n * 10 * sizeof(float)));
        FRP_TRY(cudaMemcpy(d_lmk, lmk, (size_t)n * 10 * sizeof(float), cudaMemcpyHostToDevice));
    }
    FRP_TRY(cudaMalloc((void**)&d_blob, blob_elems * sizeof(__half)));
    if (out_aligned) FRP_TRY(cudaMalloc((void**)&d_al, (size_t)n * kOut * kOut * 3));
    if (e != cudaSuccess) {
        set_error("frp_align: device staging failed: %s", cudaGetErrorString(e));
        cleanup();
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    rc = align_device_impl(device, d_frames, F, H, W, dstride, d_fi, d_lmk, d_M, n,
                           reinterpret_cast<uint16_t*>(d_blob), d_al, st);
    if (rc == FRP_OK) {
        FRP_TRY(cudaMemcpy(out_blob, d_blob, blob_elems * sizeof(__half), cudaMemcpyDeviceToHost));
        if (out_aligned) FRP_TRY(cudaMemcpy(out_aligned, d_al, (size_t)n * kOut * kOut * 3, cudaMemcpyDeviceToHost));
        if (e != cudaSuccess) {
            set_error("frp_align: kernel or D2H failed: %s", cudaGetErrorString(e));
            rc = FRP_E_CUDA;
            cudaGetLastError();
        }
    }
#undef FRP_TRY
    cleanup();
    return rc;
}

int frp_estimate_norm(const float* lmk, int n, double* M_out) {
    if (n < 0 || !lmk || !M_out) { set_error("frp_estimate_norm: invalid argument"); return FRP_E_INVALID; }
    if (n == 0) return FRP_OK;
    int ndev = 0;
    int rc = frp_device_count(&ndev);
    if (rc) return rc;
    float* d_l = nullptr; double* d_m = nullptr;
    FRP_CUDA(cudaMalloc((void**)&d_l, (size_t)n * 10 * sizeof(float)));
    cudaError_t e = cudaMalloc((void**)&d_m, (size_t)n * 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_l, lmk, (size_t)n * 10 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        estimate_norm_kernel<<<(n + 127) / 128, 128>>>(d_l, n, d_m);
        e = cudaGetLastError();
        count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpy(M_out, d_m, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_l); cudaFree(d_m);
    if (e != cudaSuccess) { set_error("frp_estimate_norm failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return FRP_E_CUDA; }
    return FRP_OK;
}

}  // extern "C"
