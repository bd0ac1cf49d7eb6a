// K1/K2 — batched 5-landmark alignment: similarity estimate (Umeyama onto
// arcface_dst), cv2.warpAffine-exact bilinear warp to 112x112 (OpenCV's fixed-point
// scheme: AB_BITS=10, INTER_BITS=5, 15-bit weights, BORDER_CONSTANT 0), fused
// (v-127.5)/127.5, BGR->RGB, HWC u8 -> NCHW fp16.
//
// Replaces insightface face_align.norm_crop + cv2.dnn.blobFromImages (third-party
// code reached from backend/app/utils/deepfake_utils.py:138).
//
// One warp per work item (half a crop).  The crop is cut into 16 x TH output tiles
// (TH = 16, or 8 for large faces); per tile ONE 3-D tiled TMA load
// (cp.async.bulk.tensor, SASS UTMALDG) stages the source bounding box of the tile —
// bytes x rows x frame, zero-filled outside the frame, which IS BORDER_CONSTANT 0 —
// into the warp's private shared-memory buffer and completes on the warp's mbarrier; 18
// warps per SM hide each other's TMA latency (FRP_ALIGN_NBUF=2 builds the double-buffered
// 10-warp variant, measured slower).  Box sizes come from a 5x5 ladder of tensor
// maps so small faces do not over-fetch (the inner TMA coordinate must be 16-byte
// aligned, so boxes start at floor16(3*sx_min)).  Lanes are mapped x-major (16 columns x
// 2 rows per warp instruction): each lane produces TH/2 output pixels from shared memory
// with integer dp2a blends (bit-exact with cv2), normalises with one FFMA (exhaustively
// equal to cv2's fp32 result after the fp16 cast) and a half-warp writes one 32-byte
// sector of each output plane.  Tiles whose footprint exceeds the largest box (or frames
// whose pitch is not 16-byte aligned) are sampled tap by tap from global memory.
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

namespace {

constexpr int kOut = 112;
constexpr int kTileW = 16;
constexpr int kTilesX = kOut / kTileW;              // 7
constexpr int kItemsPerCrop = 2;
#ifndef FRP_ALIGN_NBUF
#define FRP_ALIGN_NBUF 1     // measured: 18 single-buffered warps/SM 0.214 ms vs 10 double-buffered 0.227 ms (C4)
#endif
constexpr int kNBuf = FRP_ALIGN_NBUF;               // staging buffers per warp (2 = prefetch next tile)
constexpr int kWarps = (kNBuf == 2) ? 10 : 18;      // what fits in 227 KB of shared memory
constexpr int kNumW = 5, kNumH = 5;
__constant__ int c_box_w[kNumW] = {64, 96, 128, 176, 224};   // bytes (multiples of 16)
__constant__ int c_box_h[kNumH] = {8, 14, 22, 34, 46};       // rows
constexpr int kHostBoxW[kNumW] = {64, 96, 128, 176, 224};
constexpr int kHostBoxH[kNumH] = {8, 14, 22, 34, 46};
constexpr int kMaxBoxW = 224, kMaxBoxH = 46;
constexpr int kBufBytes = 10368;                    // >= 224*46 + 16 slack, multiple of 128
constexpr int kTabBytes = 4 * kOut * 4;             // adelta, bdelta, X0, Y0 (1792 = 14*128)
constexpr int kWarpSmem = (kNBuf == 2) ? 22656 : 12288;   // tab + buffers + mbarriers, multiple of 128

struct TmapSet {
    CUtensorMap m[kNumW * kNumH];
};

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
    int use_tma;          // tensor maps valid (16-byte aligned base / pitch)
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

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst_saddr, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_saddr), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

struct TileMeta {
    int c0;        // byte offset (in the source row) of the first staged byte
    int sy0;       // first staged source row
    int pitch;     // staged row pitch in bytes (box width)
    int direct;    // 1: footprint did not fit -> read taps from global memory
};

// choose the box for output tile (x0,y0,TH) and start its TMA load into buf
template <int TH>
__device__ __forceinline__ TileMeta stage_tile(const AlignArgs& a, const TmapSet& maps, const int* tab, int x0,
                                               int y0, int frame_index, uint32_t buf_saddr, uint64_t* bar,
                                               int lane) {
    // tables: adelta | bdelta | X0 | Y0, read with shared-space loads
    const uint32_t ts = smem_u32(tab);
    const int xa = (int)lds32(ts + 4 * x0), xb = (int)lds32(ts + 4 * (x0 + kTileW - 1));
    const int ya = (int)lds32(ts + 4 * (kOut + x0)), yb = (int)lds32(ts + 4 * (kOut + x0 + kTileW - 1));
    const int Xa = (int)lds32(ts + 4 * (2 * kOut + y0)), Xb = (int)lds32(ts + 4 * (2 * kOut + y0 + TH - 1));
    const int Ya = (int)lds32(ts + 4 * (3 * kOut + y0)), Yb = (int)lds32(ts + 4 * (3 * kOut + y0 + TH - 1));
    const int sx_min = (min(Xa, Xb) + min(xa, xb)) >> 10;
    const int sx_max = (max(Xa, Xb) + max(xa, xb)) >> 10;
    const int sy_min = (min(Ya, Yb) + min(ya, yb)) >> 10;
    const int sy_max = (max(Ya, Yb) + max(ya, yb)) >> 10;
    TileMeta m;
    m.c0 = (3 * sx_min) & ~15;                       // TMA needs the inner coordinate 16-byte aligned (floor, also for negatives)
    m.sy0 = sy_min;
    const int need_w = 3 * (sx_max + 2) - m.c0;       // bytes: taps sx .. sx_max+1 plus the alignment slack
    const int need_h = sy_max + 2 - sy_min;         // rows:  taps sy .. sy_max+1
    // coordinates near the int16 saturation of cv::remap's xy map are handled tap by tap
    const bool far = sx_min < -30000 || sx_max > 30000 || sy_min < -30000 || sy_max > 30000;
    int wi = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < kNumW - 1; ++i) wi += (need_w > c_box_w[i]) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < kNumH - 1; ++i) hi += (need_h > c_box_h[i]) ? 1 : 0;
    m.pitch = c_box_w[wi];
    m.direct = (!a.use_tma) || far || need_w > kMaxBoxW || need_h > kMaxBoxH || need_h <= 0;
    if (!m.direct && lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(c_box_w[wi] * c_box_h[hi]));
        tma_load_3d(&maps.m[hi * kNumW + wi], bar, buf_saddr, m.c0, m.sy0, frame_index);
    }
    return m;
}

__device__ __forceinline__ unsigned tap_global(const AlignArgs& a, const uint8_t* frame, int sx, int sy) {
    // returns B | G<<8 | R<<16 of pixel (sx,sy), 0 outside the frame
    if (sx < 0 || sy < 0 || sx >= a.W || sy >= a.H) return 0u;
    const uint8_t* p = frame + (size_t)sy * a.stride + 3 * sx;
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
}

// blend: 4 taps as (row0: B0 G0 R0 B1 | G1 R1 . .), (row1 likewise) -> B, G, R in [0,255]
__device__ __forceinline__ void blend_px(unsigned p0lo, unsigned p0hi, unsigned p1lo, unsigned p1hi, unsigned fx,
                                         unsigned fy, unsigned& vB, unsigned& vG, unsigned& vR) {
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
    vB = accB >> 15; vG = accG >> 15; vR = accR >> 15;
}

// (v - 127.5) * float(1/127.5) -> fp16 without a table: float(v) by the 2^23 trick, then
// ONE fused multiply-add; fp16(fma(v, c, -1)) == fp16((v - 127.5) * c) for all 256 values
// (checked exhaustively in tests/test_oracle_cpu.py::test_norm_fma_form_matches_cv2)
__device__ __forceinline__ unsigned short norm_half(unsigned v) {
    const float f = __uint_as_float(0x4B000000u | v) - 8388608.0f;
    const float n = __fmaf_rn(f, (float)(1.0 / 127.5), -1.0f);
    return __half_as_ushort(__float2half_rn(n));
}

// one output tile, x-major lane mapping: lane -> column x0 + (lane & 15) and rows
// y0 + 2j + (lane >> 4).  A warp-wide tap load then walks along ONE (rotated) source line
// per half-warp, which keeps shared-memory bank conflicts at <= ~2-way for any staged row
// pitch (the row-major mapping measured 4.7 wavefronts per load).
template <int TH>
__device__ __forceinline__ void blend_tile_x(const AlignArgs& a, const uint8_t* frame, const int* tab,
                                             const TileMeta& tm, uint32_t buf_saddr, int x0, int y0, int lane,
                                             __half* oblob, uint8_t* oal) {
    constexpr int NJ = TH / 2;
    const int x = x0 + (lane & 15);
    const int yb = y0 + (lane >> 4);
    const uint32_t ts = smem_u32(tab);
    const int ad = (int)lds32(ts + 4 * x), bd = (int)lds32(ts + 4 * (kOut + x));
    const uint32_t tsy = ts + 4 * (2 * kOut + yb);     // X0[yb]; Y0 is kOut ints further
    unsigned short* o16 = reinterpret_cast<unsigned short*>(oblob) + (size_t)yb * kOut + x;
    uint8_t* o8 = oal ? oal + ((size_t)yb * kOut + x) * 3 : nullptr;
    const int base = -tm.sy0 * tm.pitch - tm.c0;
    unsigned vB[NJ], vG[NJ], vR[NJ];
    if (!tm.direct) {      // warp-uniform: one branch per tile, not per pixel
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int y = yb + 2 * j;
            const int X = ((int)lds32(tsy + 8 * j) + ad) >> 5, Y = ((int)lds32(tsy + 8 * j + 4 * kOut) + bd) >> 5;
            const int off = (Y >> 5) * tm.pitch + 3 * (X >> 5) + base;
            const uint32_t w0 = buf_saddr + (off & ~3);
            const uint32_t w1 = w0 + tm.pitch;
            const unsigned shf = (off & 3) * 8;
            const unsigned a0 = lds32(w0), a1 = lds32(w0 + 4), a2 = lds32(w0 + 8);
            const unsigned b0 = lds32(w1), b1 = lds32(w1 + 4), b2 = lds32(w1 + 8);
            blend_px(__funnelshift_r(a0, a1, shf), __funnelshift_r(a1, a2, shf), __funnelshift_r(b0, b1, shf),
                     __funnelshift_r(b1, b2, shf), X & 31, Y & 31, vB[j], vG[j], vR[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int y = yb + 2 * j;
            const int X = ((int)lds32(tsy + 8 * j) + ad) >> 5, Y = ((int)lds32(tsy + 8 * j + 4 * kOut) + bd) >> 5;
            const int sx = max(-32768, min(32767, X >> 5));
            const int sy = max(-32768, min(32767, Y >> 5));
            const unsigned t00 = tap_global(a, frame, sx, sy), t01 = tap_global(a, frame, sx + 1, sy);
            const unsigned t10 = tap_global(a, frame, sx, sy + 1), t11 = tap_global(a, frame, sx + 1, sy + 1);
            blend_px(t00 | (t01 << 24), t01 >> 8, t10 | (t11 << 24), t11 >> 8, X & 31, Y & 31, vB[j], vG[j],
                     vR[j]);
        }
    }
    // BGR -> RGB planes; a half-warp writes 16 consecutive halfs (one 32-byte sector)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        o16[(size_t)j * 2 * kOut] = norm_half(vR[j]);
        o16[(size_t)kOut * kOut + (size_t)j * 2 * kOut] = norm_half(vG[j]);
        o16[(size_t)2 * kOut * kOut + (size_t)j * 2 * kOut] = norm_half(vB[j]);
    }
    if (o8) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            uint8_t* p = o8 + (size_t)j * 2 * kOut * 3;
            p[0] = (uint8_t)vB[j]; p[1] = (uint8_t)vG[j]; p[2] = (uint8_t)vR[j];
        }
    }
}

// all tiles [t_begin, t_end) of one crop with tile height TH, double buffered
template <int TH>
__device__ __forceinline__ void run_tiles(const AlignArgs& a, const TmapSet& maps, const uint8_t* frame, int fi,
                                          const int* tab, uint32_t buf_saddr0, uint64_t* bars, uint32_t& parity,
                                          int t_begin, int t_end, int lane, __half* oblob, uint8_t* oal) {
    if (kNBuf == 1) {
        // single buffer, more resident warps: the TMA latency of one warp is covered by the
        // blends of the others
        for (int t = t_begin; t < t_end; ++t) {
            const TileMeta tm = stage_tile<TH>(a, maps, tab, (t % kTilesX) * kTileW, (t / kTilesX) * TH, fi,
                                               buf_saddr0, &bars[0], lane);
            if (!tm.direct) {
                mbar_wait(&bars[0], parity & 1u);
                parity ^= 1u;
            }
            blend_tile_x<TH>(a, frame, tab, tm, buf_saddr0, (t % kTilesX) * kTileW, (t / kTilesX) * TH, lane,
                             oblob, oal);
            __syncwarp();
        }
        return;
    }
    TileMeta cur, nxt;
    cur = stage_tile<TH>(a, maps, tab, (t_begin % kTilesX) * kTileW, (t_begin / kTilesX) * TH, fi, buf_saddr0,
                         &bars[0], lane);
    nxt = cur;
    for (int t = t_begin; t < t_end; ++t) {
        const int b = (t - t_begin) & 1;
        if (t + 1 < t_end) {
            const int tn = t + 1;
            nxt = stage_tile<TH>(a, maps, tab, (tn % kTilesX) * kTileW, (tn / kTilesX) * TH, fi,
                                 buf_saddr0 + (uint32_t)((b ^ 1) * kBufBytes), &bars[b ^ 1], lane);
        }
        if (!cur.direct) {
            mbar_wait(&bars[b], (parity >> b) & 1u);
            parity ^= (1u << b);
        }
        blend_tile_x<TH>(a, frame, tab, cur, buf_saddr0 + (uint32_t)(b * kBufBytes), (t % kTilesX) * kTileW,
                         (t / kTilesX) * TH, lane, oblob, oal);
        __syncwarp();      // every lane is done with buffer b before it is refilled
        cur = nxt;
    }
}

__global__ void __launch_bounds__(kWarps * 32, 1)
align_kernel(const __grid_constant__ TmapSet maps, const AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t* wsm = smem + (size_t)warp * kWarpSmem;
    int* tab = reinterpret_cast<int*>(wsm);
    const uint32_t buf_saddr0 = smem_u32(wsm + kTabBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + kTabBytes + kNBuf * kBufBytes);
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    const int nitems = a.n * kItemsPerCrop;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1);
        item = __shfl_sync(FRP_FULL_MASK, item, 0);
        if (item >= nitems) break;
        const int crop = item / kItemsPerCrop;
        const int half = item % kItemsPerCrop;

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

        // 16x16 tiles when their source footprint fits the largest box, else 16x8
        const double ax = fabs(iM[0]), ay = fabs(iM[1]), bx = fabs(iM[3]), by = fabs(iM[4]);
        const bool big = 3.0 * (16.0 * ax + 16.0 * ay + 4.0) + 15.0 <= (double)kMaxBoxW &&
                         (16.0 * bx + 16.0 * by + 4.0) <= (double)kMaxBoxH;
        if (big) {
            const int nt = kTilesX * (kOut / 16);   // 49
            run_tiles<16>(a, maps, frame, fi, tab, buf_saddr0, bars, parity, half * nt / 2, (half + 1) * nt / 2,
                          lane, oblob, oal);
        } else {
            const int nt = kTilesX * (kOut / 8);    // 98
            run_tiles<8>(a, maps, frame, fi, tab, buf_saddr0, bars, parity, half * nt / 2, (half + 1) * nt / 2,
                         lane, oblob, oal);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// u8 tensor [F][H][3W bytes] with pitches (stride, frame_bytes); one map per box size
bool encode_frame_maps(TmapSet* set, const uint8_t* frames, int F, int H, int W, size_t stride,
                       size_t frame_bytes) {
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tensormap_encode_fn());
    if (!fn) return false;
    if (((uintptr_t)frames & 15) || (stride & 15) || (frame_bytes & 15)) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)3 * W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t gstr[2] = {(cuuint64_t)stride, (cuuint64_t)frame_bytes};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int hi = 0; hi < kNumH; ++hi)
        for (int wi = 0; wi < kNumW; ++wi) {
            cuuint32_t box[3] = {(cuuint32_t)kHostBoxW[wi], (cuuint32_t)kHostBoxH[hi], 1};
            CUresult r = fn(&set->m[hi * kNumW + wi], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(frames),
                            gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return false;
        }
    return true;
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
    TmapSet maps;
    memset(&maps, 0, sizeof(maps));
    a.use_tma = encode_frame_maps(&maps, frames, F, H, W, stride, a.frame_bytes) ? 1 : 0;
    const size_t smem = (size_t)kWarps * kWarpSmem + 128;
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
    align_kernel<<<grid, kWarps * 32, smem, st>>>(maps, a);
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
    // device copies keep a 16-byte aligned pitch so the TMA-staged path is always taken
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
