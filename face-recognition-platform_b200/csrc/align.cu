// K1/K2 — batched 5-landmark alignment: similarity estimate (Umeyama onto
// arcface_dst), cv2.warpAffine-exact bilinear warp to 112x112 (OpenCV's fixed-point
// scheme: AB_BITS=10, INTER_BITS=5, 15-bit weights, BORDER_CONSTANT 0), fused
// (v-127.5)/127.5, BGR->RGB, HWC u8 -> NCHW fp16.
//
// Replaces insightface face_align.norm_crop + cv2.dnn.blobFromImages (third-party
// code reached from backend/app/utils/deepfake_utils.py:138).
//
// align_plan_kernel (one thread per crop) computes the fp64 similarity estimate and cv2's
// inverse once; align_kernel then runs one warp per work item (a seventh of a crop's tiles,
// handed out by a dynamic counter one ticket ahead).  The crop is cut into 16 x TH output
// tiles (TH = 16, or 8 for large faces); per tile ONE 3-D tiled TMA load
// (cp.async.bulk.tensor, SASS UTMALDG) stages the source bounding box of the tile —
// bytes x rows x frame, zero-filled outside the frame, which IS BORDER_CONSTANT 0 —
// into the warp's private shared-memory buffer and completes on the warp's mbarrier; 18
// warps per SM hide each other's TMA latency, and crops whose box fits half the buffer
// also prefetch their next tile.  The box size comes from a 5x5 ladder of tensor maps and
// is fixed per crop (the map is affine: every tile of a crop has the same extents), the
// box origin is the sum of two table entries (the inner TMA coordinate must be 16-byte
// aligned, so boxes start at floor16(3*sx_min)).  Lanes are mapped x-major (16 columns x
// 2 rows per warp instruction): each lane produces TH/2 output pixels from shared memory
// with integer dp2a blends (bit-exact with cv2), normalises with one FFMA straight from
// the un-shifted blend sum (exhaustively equal to cv2's fp32 result after the fp16 cast)
// and a half-warp writes one 32-byte sector of each output plane.  Crops whose tiles
// exceed the largest box, whose coordinates approach cv::remap's int16 saturation, or
// whose frames are not 16-byte aligned are sampled tap by tap from global memory.
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
#ifndef FRP_ALIGN_WARPS
#define FRP_ALIGN_WARPS 18
#endif
constexpr int kWarps = FRP_ALIGN_WARPS;             // what fits in 227 KB of shared memory
#ifndef FRP_ALIGN_ITEMS
#define FRP_ALIGN_ITEMS 7     // measured (C4): 2 -> 0.183 ms, 4 -> 0.170 ms, 7 -> 0.165 ms
#endif
constexpr int kItemsPerCrop = FRP_ALIGN_ITEMS;      // work items (tile ranges) per crop
#ifndef FRP_ALIGN_PREF
#define FRP_ALIGN_PREF 1     // 0: prefer (TH=8, prefetch) over (TH=16, no prefetch); 1: the reverse
#endif
constexpr int kNumW = 5, kNumH = 5;
__constant__ int c_box_w[kNumW] = {64, 96, 128, 176, 224};   // bytes (multiples of 16)
__constant__ int c_box_h[kNumH] = {8, 14, 22, 34, 46};       // rows
constexpr int kHostBoxW[kNumW] = {64, 96, 128, 176, 224};
constexpr int kHostBoxH[kNumH] = {8, 14, 22, 34, 46};
constexpr int kMaxBoxW = 224, kMaxBoxH = 46;
#ifndef FRP_ALIGN_BUF
#define FRP_ALIGN_BUF 10368
#endif
constexpr int kBufBytes = FRP_ALIGN_BUF;            // >= 224*46 + 16 slack, multiple of 128
constexpr int kTabBytes = 4 * kOut * 4;             // adelta, bdelta, X0, Y0 (1792 = 14*128)
constexpr int kHalfBuf = (kBufBytes / 2) & ~127;     // boxes up to this size are double buffered inside kBufBytes
constexpr int kWarpSmem = kTabBytes + kBufBytes + 128;   // tab + buffer + mbarriers, multiple of 128

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
    double* iM;           // [n][6] inverse maps (workspace, written by align_plan_kernel)
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
__device__ __forceinline__ int2 lds64(uint32_t saddr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

struct TileMeta {
    int c0;        // byte offset (in the source row) of the first staged byte
    int sy0;       // first staged source row
    int pitch;     // staged row pitch in bytes (box width)
    int direct;    // 1: footprint did not fit -> read taps from global memory
};

// Per-crop staging plan.  The inverse map is affine and the fixed-point tables are monotone
// in x and in y, so (a) the source bounding box of EVERY 16 x TH tile of a crop is bounded by
// the same extents, and (b) its minimum corner is the sum of two table entries whose
// positions depend only on the signs of the matrix entries: one box size per crop, and
// four table reads per tile.
struct CropPlan {
    int map;            // tensor-map index (hi * kNumW + wi)
    int pitch;          // box width in bytes
    int box_bytes;
    int xsel_x, ysel_x; // tile-corner offsets (0 or kTileW-1, 0 or TH-1) of the minimum source x
    int xsel_y, ysel_y; // likewise for the minimum source y
    int direct;         // no TMA staging for this crop
    int dual;           // box fits half the buffer: prefetch the next tile while blending
};

// Box for 16 x TH tiles of a map with |iM| = (ax ay; bx by): false if it exceeds the ladder.
// Over a tile X = X0[y] + adelta[x] spans at most (15 ax + (TH-1) ay) * 1024 + 2 fixed-point
// units (each table rounds once), so sx_max - sx_min <= floor(ext_x + 2/1024) + 1; with
// c0 >= 3 sx_min - 15 the staged row needs 3 (sx_max + 2) - c0 <= 3 (floor(ext_x) + 3) + 15 bytes.
template <int TH>
__device__ __forceinline__ bool plan_box(double ax, double ay, double bx, double by, int& wi, int& hi) {
    const double ext_x = 15.0 * ax + (double)(TH - 1) * ay, ext_y = 15.0 * bx + (double)(TH - 1) * by;
    if (!(ext_x < 1000.0 && ext_y < 1000.0)) return false;      // also rejects NaN
    const int need_w = 3 * ((int)(ext_x + 0.01) + 3) + 15;
    const int need_h = (int)(ext_y + 0.01) + 3;
    if (need_w > kMaxBoxW || need_h > kMaxBoxH) return false;
    wi = 0; hi = 0;
#pragma unroll
    for (int i = 0; i < kNumW - 1; ++i) wi += (need_w > c_box_w[i]) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < kNumH - 1; ++i) hi += (need_h > c_box_h[i]) ? 1 : 0;
    return c_box_w[wi] * c_box_h[hi] <= kBufBytes;
}

// locate the box of output tile (x0,y0) and start its TMA load into buf
__device__ __forceinline__ TileMeta stage_tile(const CropPlan& pl, const TmapSet& maps, const int* tab, int x0,
                                               int y0, int frame_index, uint32_t buf_saddr, uint64_t* bar,
                                               int lane) {
    // tables: adelta | bdelta | (X0,Y0) pairs, read with shared-space loads
    const uint32_t ts = smem_u32(tab);
    const int Xmin = (int)lds32(ts + 8 * (kOut + y0 + pl.ysel_x)) + (int)lds32(ts + 4 * (x0 + pl.xsel_x));
    const int Ymin = (int)lds32(ts + 8 * (kOut + y0 + pl.ysel_y) + 4) + (int)lds32(ts + 4 * (kOut + x0 + pl.xsel_y));
    TileMeta m;
    m.c0 = (3 * (Xmin >> 10)) & ~15;   // TMA needs the inner coordinate 16-byte aligned (floor, also for negatives)
    m.sy0 = Ymin >> 10;
    m.pitch = pl.pitch;
    m.direct = pl.direct;
    if (!pl.direct && lane == 0) {
        mbar_expect_tx(bar, (uint32_t)pl.box_bytes);
        tma_load_3d(&maps.m[pl.map], bar, buf_saddr, m.c0, m.sy0, frame_index);
    }
    return m;
}

__device__ __forceinline__ unsigned tap_global(const AlignArgs& a, const uint8_t* frame, int sx, int sy) {
    // returns B | G<<8 | R<<16 of pixel (sx,sy), 0 outside the frame
    if (sx < 0 || sy < 0 || sx >= a.W || sy >= a.H) return 0u;
    const uint8_t* p = frame + (size_t)sy * a.stride + 3 * sx;
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
}

// blend: 4 taps as (row0: B0 G0 R0 B1 | G1 R1 . .), (row1 likewise) -> the three 23-bit sums
// acc = sum(w * p) + 2^14 (the pixel value is acc >> 15).  fx in [0,32), ty = fy << 5.
__device__ __forceinline__ void blend_px(unsigned p0lo, unsigned p0hi, unsigned p1lo, unsigned p1hi, unsigned fx,
                                         unsigned ty, unsigned& accB, unsigned& accG, unsigned& accR) {
    // (B0,B1,G0,G1) and (R0,R1) byte pairs for the two 16-bit x 8-bit dot products
    const unsigned bg0 = __byte_perm(p0lo, p0hi, 0x4130), rr0 = __byte_perm(p0lo, p0hi, 0x0052);
    const unsigned bg1 = __byte_perm(p1lo, p1hi, 0x4130), rr1 = __byte_perm(p1lo, p1hi, 0x0052);
    // weights: w00=(32-fx)(32-fy)*32, w01=fx(32-fy)*32, w10=(32-fx)fy*32, w11=fx*fy*32
    const unsigned P = fx * 0xFFFFu + 32u;          // (32-fx) | fx<<16
    const unsigned W1 = ty * P;                     // w10 | w11<<16
    const unsigned W0 = P * 1024u - W1;             // w00 | w01<<16 (no borrow between the fields)
    accB = __dp2a_lo(W0, bg0, 16384u); accB = __dp2a_lo(W1, bg1, accB);
    accG = __dp2a_hi(W0, bg0, 16384u); accG = __dp2a_hi(W1, bg1, accG);
    accR = __dp2a_lo(W0, rr0, 16384u); accR = __dp2a_lo(W1, rr1, accR);
}

// (v - 127.5) * float(1/127.5) -> fp16 for v = acc >> 15, straight from the un-shifted sum:
// 0x4B000000 | (acc & 0x7F8000) is the float 2^23 + 32768 v, and ONE fused multiply-add with
// c/32768 and the exactly representable constant -(1 + 256 c) yields round(v c - 1):
// fp16(fma(v, c, -1)) == fp16((v - 127.5) * c) for all 256 values (checked exhaustively in
// tests/test_oracle_cpu.py::test_norm_fma_form_matches_cv2)
__device__ __forceinline__ unsigned short norm_half(unsigned acc) {
    constexpr float c = (float)(1.0 / 127.5);
    const float f = __uint_as_float(0x4B000000u | (acc & 0x7F8000u));
    const float n = __fmaf_rn(f, c * (1.0f / 32768.0f), -(1.0f + 256.0f * c));
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
    const uint32_t tsy = ts + 8 * (kOut + yb);         // (X0,Y0)[yb]
    unsigned short* o16 = reinterpret_cast<unsigned short*>(oblob) + (size_t)yb * kOut + x;
    uint8_t* o8 = oal ? oal + ((size_t)yb * kOut + x) * 3 : nullptr;
    const int base = -tm.sy0 * tm.pitch - tm.c0;
    unsigned vB[NJ], vG[NJ], vR[NJ];      // un-shifted blend sums (pixel << 15 | fraction)
    if (!tm.direct) {      // warp-uniform: one branch per tile, not per pixel
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int2 xy = lds64(tsy + 16 * j);
            const int X = xy.x + ad, Y = xy.y + bd;          // 10 fractional bits
            const int off = (Y >> 10) * tm.pitch + 3 * (X >> 10) + base;
            const uint32_t w0 = buf_saddr + (off & ~3);
            const uint32_t w1 = w0 + tm.pitch;
            const unsigned shf = (unsigned)off << 3;          // funnel shifts use the low 5 bits
            const unsigned a0 = lds32(w0), a1 = lds32(w0 + 4), a2 = lds32(w0 + 8);
            const unsigned b0 = lds32(w1), b1 = lds32(w1 + 4), b2 = lds32(w1 + 8);
            blend_px(__funnelshift_r(a0, a1, shf), __funnelshift_r(a1, a2, shf), __funnelshift_r(b0, b1, shf),
                     __funnelshift_r(b1, b2, shf), (X >> 5) & 31, Y & 0x3E0, vB[j], vG[j], vR[j]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int2 xy = lds64(tsy + 16 * j);
            const int X = (xy.x + ad) >> 5, Y = (xy.y + bd) >> 5;
            const int sx = max(-32768, min(32767, X >> 5));
            const int sy = max(-32768, min(32767, Y >> 5));
            const unsigned t00 = tap_global(a, frame, sx, sy), t01 = tap_global(a, frame, sx + 1, sy);
            const unsigned t10 = tap_global(a, frame, sx, sy + 1), t11 = tap_global(a, frame, sx + 1, sy + 1);
            blend_px(t00 | (t01 << 24), t01 >> 8, t10 | (t11 << 24), t11 >> 8, X & 31, (Y & 31) << 5, vB[j], vG[j],
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
            p[0] = (uint8_t)(vB[j] >> 15); p[1] = (uint8_t)(vG[j] >> 15); p[2] = (uint8_t)(vR[j] >> 15);
        }
    }
}

// all tiles [t_begin, t_end) of one crop with tile height TH
template <int TH>
__device__ __forceinline__ void run_tiles(const AlignArgs& a, const TmapSet& maps, const CropPlan& pl,
                                          const uint8_t* frame, int fi, const int* tab, uint32_t buf_saddr0,
                                          uint64_t* bars, uint32_t& parity, int t_begin, int t_end, int lane,
                                          __half* oblob, uint8_t* oal) {
    if (!pl.dual) {
        // one buffer: the TMA latency of this warp is covered by the blends of the others
        for (int t = t_begin; t < t_end; ++t) {
            const TileMeta tm = stage_tile(pl, maps, tab, (t % kTilesX) * kTileW, (t / kTilesX) * TH, fi,
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
    // two half buffers: the next tile's box is in flight while this one is blended
    TileMeta cur, nxt;
    cur = stage_tile(pl, maps, tab, (t_begin % kTilesX) * kTileW, (t_begin / kTilesX) * TH, fi, buf_saddr0,
                     &bars[0], lane);
    nxt = cur;
    for (int t = t_begin; t < t_end; ++t) {
        const int b = (t - t_begin) & 1;
        if (t + 1 < t_end) {
            const int tn = t + 1;
            nxt = stage_tile(pl, maps, tab, (tn % kTilesX) * kTileW, (tn / kTilesX) * TH, fi,
                             buf_saddr0 + (uint32_t)((b ^ 1) * kHalfBuf), &bars[b ^ 1], lane);
        }
        mbar_wait(&bars[b], (parity >> b) & 1u);
        parity ^= (1u << b);
        blend_tile_x<TH>(a, frame, tab, cur, buf_saddr0 + (uint32_t)(b * kHalfBuf), (t % kTilesX) * kTileW,
                         (t / kTilesX) * TH, lane, oblob, oal);
        __syncwarp();      // every lane is done with buffer b before it is refilled
        cur = nxt;
    }
}

// One thread per crop: forward similarity (or the caller's matrix), cv2's inverse, and the
// degenerate-matrix guard.  Runs ahead of align_kernel so that the work items of a crop do
// not each repeat the fp64 estimate.
__global__ void align_plan_kernel(const AlignArgs a) {
    const int crop = blockIdx.x * blockDim.x + threadIdx.x;
    if (crop >= a.n) return;
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
#pragma unroll
    for (int i = 0; i < 6; ++i) a.iM[(size_t)crop * 6 + i] = iM[i];
}

__global__ void __launch_bounds__(kWarps * 32, 1)
align_kernel(const __grid_constant__ TmapSet maps, const AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t* wsm = smem + (size_t)warp * kWarpSmem;
    int* tab = reinterpret_cast<int*>(wsm);
    const uint32_t buf_saddr0 = smem_u32(wsm + kTabBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + kTabBytes + kBufBytes);
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    const int nitems = a.n * kItemsPerCrop;
    // the counter is bumped one item ahead: the atomic's round trip overlaps the current item
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.counter, 1);
    for (;;) {
        const int item = __shfl_sync(FRP_FULL_MASK, ticket, 0);
        if (item >= nitems) break;
        if (lane == 0) ticket = atomicAdd(a.counter, 1);
        const int crop = item / kItemsPerCrop;
        const int half = item % kItemsPerCrop;

        // ---- inverse map of this crop (align_plan_kernel computed it once per crop) ----
        double iM[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) iM[i] = a.iM[(size_t)crop * 6 + i];
        int fi = a.frame_idx[crop];
        fi = min(max(fi, 0), a.F - 1);
        const uint8_t* frame = a.frames + (size_t)fi * a.frame_bytes;
        __syncwarp();
        for (int x = lane; x < kOut; x += 32) {
            const double xd = (double)x;
            tab[x] = __double2int_rn(__dmul_rn(__dmul_rn(iM[0], xd), 1024.0));
            tab[kOut + x] = __double2int_rn(__dmul_rn(__dmul_rn(iM[3], xd), 1024.0));
            // (X0[y], Y0[y]) interleaved: the blend fetches both with one 64-bit load
            tab[2 * kOut + 2 * x] =
                __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[1], xd), iM[2]), 1024.0)) + 16;
            tab[2 * kOut + 2 * x + 1] =
                __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[4], xd), iM[5]), 1024.0)) + 16;
        }
        __syncwarp();

        __half* oblob = a.out_blob + (size_t)crop * 3 * kOut * kOut;
        uint8_t* oal = a.out_aligned ? a.out_aligned + (size_t)crop * kOut * kOut * 3 : nullptr;

        // ---- staging plan: tile height, box size, single or double buffering ----
        const double ax = fabs(iM[0]), ay = fabs(iM[1]), bx = fabs(iM[3]), by = fabs(iM[4]);
        int wi16 = 0, hi16 = 0, wi8 = 0, hi8 = 0;
        const bool fit16 = plan_box<16>(ax, ay, bx, by, wi16, hi16);
        const bool fit8 = plan_box<8>(ax, ay, bx, by, wi8, hi8);
        const int bytes16 = c_box_w[wi16] * c_box_h[hi16], bytes8 = c_box_w[wi8] * c_box_h[hi8];
        // coordinates near the int16 saturation of cv::remap's xy map are handled tap by tap:
        // the map is affine, so its extremes over the crop are at the four corners
        bool far = false;
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
                const double sx = iM[0] * (cx * (kOut - 1)) + iM[1] * (cy * (kOut - 1)) + iM[2];
                const double sy = iM[3] * (cx * (kOut - 1)) + iM[4] * (cy * (kOut - 1)) + iM[5];
                far = far || !(fabs(sx) < 29000.0 && fabs(sy) < 29000.0);
            }
        const bool tma_ok = a.use_tma && !far;
        bool th16;
        CropPlan pl;
        if (FRP_ALIGN_PREF == 0)
            th16 = fit16 && (bytes16 <= kHalfBuf || !(fit8 && bytes8 <= kHalfBuf));
        else
            th16 = fit16;
        pl.direct = !(tma_ok && (th16 ? fit16 : fit8));
        const int wi = th16 ? wi16 : wi8, hi = th16 ? hi16 : hi8;
        const int th1 = th16 ? 15 : 7;
        pl.map = hi * kNumW + wi;
        pl.pitch = c_box_w[wi];
        pl.box_bytes = th16 ? bytes16 : bytes8;
        pl.dual = !pl.direct && pl.box_bytes <= kHalfBuf;
        pl.xsel_x = (iM[0] >= 0.0) ? 0 : kTileW - 1;
        pl.ysel_x = (iM[1] >= 0.0) ? 0 : th1;
        pl.xsel_y = (iM[3] >= 0.0) ? 0 : kTileW - 1;
        pl.ysel_y = (iM[4] >= 0.0) ? 0 : th1;
        if (th16) {
            const int nt = kTilesX * (kOut / 16);   // 49
            run_tiles<16>(a, maps, pl, frame, fi, tab, buf_saddr0, bars, parity, half * nt / kItemsPerCrop,
                          (half + 1) * nt / kItemsPerCrop, lane, oblob, oal);
        } else {
            const int nt = kTilesX * (kOut / 8);    // 98
            run_tiles<8>(a, maps, pl, frame, fi, tab, buf_saddr0, bars, parity, half * nt / kItemsPerCrop,
                         (half + 1) * nt / kItemsPerCrop, lane, oblob, oal);
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
    // workspace: the work counter (first 256 bytes) and the per-crop inverse maps
    uint8_t* ws = nullptr;
    FRP_CUDA(cudaMallocAsync((void**)&ws, 256 + (size_t)n * 6 * sizeof(double), st));
    int* counter = reinterpret_cast<int*>(ws);
    FRP_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    AlignArgs a;
    a.frames = frames; a.F = F; a.H = H; a.W = W; a.stride = stride;
    a.frame_bytes = (size_t)H * stride;
    a.frame_idx = frame_idx; a.lmk = lmk; a.M = M; a.n = n;
    a.out_blob = reinterpret_cast<__half*>(out_blob);
    a.out_aligned = out_aligned;
    a.counter = counter;
    a.iM = reinterpret_cast<double*>(ws + 256);
    TmapSet maps;
    memset(&maps, 0, sizeof(maps));
    a.use_tma = encode_frame_maps(&maps, frames, F, H, W, stride, a.frame_bytes) ? 1 : 0;
    const size_t smem = (size_t)kWarps * kWarpSmem + 128;
    {
        int rc = FRP_OK;
        configure_func_smem((const void*)align_kernel, device, smem, &rc);
        if (rc) return rc;
    }
    const int items = n * kItemsPerCrop;
    int grid = (items + kWarps - 1) / kWarps;
    if (grid > s_sms[di]) grid = s_sms[di];
    if (prof) {
        if (!t_align_ev[0]) { cudaEventCreate(&t_align_ev[0]); cudaEventCreate(&t_align_ev[1]); }
        cudaEventRecord(t_align_ev[0], st);
    }
    align_plan_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
    FRP_CHECK_LAUNCH();
    align_kernel<<<grid, kWarps * 32, smem, st>>>(maps, a);
    FRP_CHECK_LAUNCH();
    if (prof) { cudaEventRecord(t_align_ev[1], st); t_align_valid = true; }
    FRP_CUDA(cudaFreeAsync(ws, st));
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
    // a buffer that lives on another GPU would fault inside the kernel: refuse it here
    for (const void* p : {(const void*)frames_dev, (const void*)out_blob_dev}) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != device) {
            set_error("frp_align_device: buffer lives on device %d, call names device %d", pa.device, device);
            return FRP_E_INVALID;
        }
        cudaGetLastError();
    }
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
