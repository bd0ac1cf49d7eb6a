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
#include <stdio.h>
#include <string.h>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

namespace {

constexpr int kOut = 112;
constexpr int kTileW = 16;
constexpr int kTilesX = kOut / kTileW;              // 7
constexpr int kBandRows = 16;                       // output rows per work item
constexpr int kBands = kOut / kBandRows;            // 7 work items per crop
#ifndef FRP_ALIGN_WARPS
#define FRP_ALIGN_WARPS 18
#endif
constexpr int kWarps = FRP_ALIGN_WARPS;             // what fits in 227 KB of shared memory
// Box ladder of the staged source tiles: widths in steps of 32 bytes, heights in steps of 4 rows
// (66 tensor maps, 8.25 KB of kernel parameters).  On the C4 workload the ladder stages
// 1.06 GB against 1.24 GB for the 5 x 5 ladder of round 1 (exact bounding boxes: 0.81 GB).
constexpr int kNumW = 6, kNumH = 11;
constexpr int kMaxBoxW = 224, kMaxBoxH = 48;
__host__ __device__ constexpr int box_w(int i) { return 64 + 32 * i; }   // bytes (multiples of 16)
__host__ __device__ constexpr int box_h(int i) { return 8 + 4 * i; }     // rows
#ifndef FRP_ALIGN_BUF
#define FRP_ALIGN_BUF 10240
#endif
constexpr int kBufBytes = FRP_ALIGN_BUF;            // staging bytes per warp, multiple of 128
constexpr int kHalfBuf = (kBufBytes / 2) & ~127;    // boxes up to this size are double buffered inside kBufBytes

// ---- per-crop record written by align_plan_kernel, read by align_kernel with bulk copies ----
//   [0,32)      CropHdr
//   [32,480)    adelta[112]  int32   round(iM00 * x * 1024)
//   [480,928)   bdelta[112]  int32   round(iM10 * x * 1024)
//   [928,...)   7 bands x 256 B: (X0,Y0)[16] int2 of the band's rows | tile origins (c0, sy0)[14] int2
// A work item (crop, band) needs the common block (928 B) and its band block (256 B).
struct CropHdr {
    int frame;       // clamped frame index
    int map;         // tensor-map index (hi * kNumW + wi)
    int pitch;       // staged row pitch = box width in bytes
    int box_bytes;
    int flags;       // bit 0 direct (no TMA staging), bit 1 dual (box fits half the buffer), bit 2 TH = 8
    int pad[3];
};
constexpr int kCommonBytes = 32 + 2 * kOut * 4;     // 928
constexpr int kBandBytes = 256;
constexpr int kRecBytes = ((kCommonBytes + kBands * kBandBytes + 63) / 64) * 64;   // 2752
constexpr int kItemBytes = kCommonBytes + kBandBytes;                             // 1184, per table slot
constexpr int kTabBytes = ((2 * kItemBytes + 127) / 128) * 128;                   // two table slots (2432)
constexpr int kWarpSmem = kTabBytes + kBufBytes + 128;                            // + mbarriers; multiple of 128
static_assert(kBufBytes % 128 == 0 && kWarpSmem % 128 == 0, "TMA destinations are 128-byte aligned");
static_assert((size_t)kWarps * kWarpSmem + 128 <= 227 * 1024, "per-warp staging exceeds the 227 KB of shared memory");

struct TmapSet {
    CUtensorMap m[kNumW * kNumH];
};
#ifdef FRP_ALIGN_DEBUG
__device__ int g_align_dbg[64];
#endif

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
    int* counter;         // [0] dynamic work counter, [1] number of direct crops
    int* direct_list;     // [n] crops the plan marked direct (compact, unordered)
    uint8_t* rec;         // [n][kRecBytes] plan records (workspace, written by align_plan_kernel)
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

// The same estimate as insightface actually evaluates it: it hands skimage float32 landmarks and
// the float32 template, so _umeyama's means, deviations, covariance A and variance are float32
// (numpy keeps the dtype) while the composition (d, T, scale) is float64.  The rotation is
// the closed form of the 2x2 SVD applied to the float32 A; LAPACK's float32 SVD rounds
// differently in the last bits, so matrices agree with skimage-on-float32 to ~1e-6 relative,
// not bit for bit (the float64 form above is what SURVEY 8c specifies and what frp_align uses
// when it is given landmarks).
__device__ void estimate_norm_f32_dev(const float* lm, double* M) {
    const float dstf[10] = {38.2946f, 51.6963f, 73.5318f, 51.5014f, 56.0252f, 71.7366f, 41.5493f, 92.3655f, 70.7299f, 92.2041f};
    float sx = 0.f, sy = 0.f, dx = 0.f, dy = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        sx = __fadd_rn(sx, lm[2 * i]); sy = __fadd_rn(sy, lm[2 * i + 1]);
        dx = __fadd_rn(dx, dstf[2 * i]); dy = __fadd_rn(dy, dstf[2 * i + 1]);
    }
    sx = __fdiv_rn(sx, 5.f); sy = __fdiv_rn(sy, 5.f); dx = __fdiv_rn(dx, 5.f); dy = __fdiv_rn(dy, 5.f);
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, px_[5], py_[5], mx = 0.f, my = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float px = __fsub_rn(lm[2 * i], sx), py = __fsub_rn(lm[2 * i + 1], sy);
        const float qx = __fsub_rn(dstf[2 * i], dx), qy = __fsub_rn(dstf[2 * i + 1], dy);
        a00 = __fadd_rn(a00, __fmul_rn(qx, px)); a01 = __fadd_rn(a01, __fmul_rn(qx, py));
        a10 = __fadd_rn(a10, __fmul_rn(qy, px)); a11 = __fadd_rn(a11, __fmul_rn(qy, py));
        px_[i] = px; py_[i] = py;
        mx = __fadd_rn(mx, px); my = __fadd_rn(my, py);
    }
    a00 = __fdiv_rn(a00, 5.f); a01 = __fdiv_rn(a01, 5.f); a10 = __fdiv_rn(a10, 5.f); a11 = __fdiv_rn(a11, 5.f);
    // numpy var: mean of squared deviations from the (float32) mean of the already demeaned data
    mx = __fdiv_rn(mx, 5.f); my = __fdiv_rn(my, 5.f);
    float vx = 0.f, vy = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float ex = __fsub_rn(px_[i], mx), ey = __fsub_rn(py_[i], my);
        vx = __fadd_rn(vx, __fmul_rn(ex, ex)); vy = __fadd_rn(vy, __fmul_rn(ey, ey));
    }
    const float var = __fadd_rn(__fdiv_rn(vx, 5.f), __fdiv_rn(vy, 5.f));
    // float64 composition: scale * R = [[a, -b], [b, a]] * (1 / var) with a = A00 + A11, b = A10 - A01
    const double a = (double)a00 + (double)a11, b = (double)a10 - (double)a01;
    const double inv = 1.0 / (double)var;
    const double m00 = a * inv, m10 = b * inv;
    M[0] = m00; M[1] = -m10; M[2] = (double)dx - (m00 * (double)sx - m10 * (double)sy);
    M[3] = m10; M[4] = m00;  M[5] = (double)dy - (m10 * (double)sx + m00 * (double)sy);
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
    wi = need_w <= 64 ? 0 : (need_w - 64 + 31) / 32;
    hi = need_h <= 8 ? 0 : (need_h - 8 + 3) / 4;
    return box_w(wi) * box_h(hi) <= kBufBytes;
}

// ---------------------------------------------------------------------------------------------
// align_plan_kernel: one warp per crop.  Forward similarity (or the caller's matrix), cv2's
// inverse, the fixed-point coordinate tables, the staging plan (tile height, box size, single or
// double buffering) and the source origin of every tile — everything a work item of
// align_kernel needs, so that kernel runs no fp64 and no per-tile box search.
// ---------------------------------------------------------------------------------------------
constexpr int kPlanWarps = 4;

__global__ void __launch_bounds__(kPlanWarps * 32) align_plan_kernel(const __grid_constant__ AlignArgs a) {
    __shared__ int s_tab[kPlanWarps][4 * kOut];      // adelta | bdelta | X0 | Y0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crop = blockIdx.x * kPlanWarps + warp;
    if (crop >= a.n) return;
    int* tab = s_tab[warp];
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

    uint8_t* rec = a.rec + (size_t)crop * kRecBytes;
    int* r_ad = reinterpret_cast<int*>(rec + 32);
    int* r_bd = reinterpret_cast<int*>(rec + 32 + kOut * 4);
    for (int x = lane; x < kOut; x += 32) {
        const double xd = (double)x;
        const int ad = __double2int_rn(__dmul_rn(__dmul_rn(iM[0], xd), 1024.0));
        const int bd = __double2int_rn(__dmul_rn(__dmul_rn(iM[3], xd), 1024.0));
        const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[1], xd), iM[2]), 1024.0)) + 16;
        const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iM[4], xd), iM[5]), 1024.0)) + 16;
        tab[x] = ad; tab[kOut + x] = bd; tab[2 * kOut + x] = X0; tab[3 * kOut + x] = Y0;
        r_ad[x] = ad;
        r_bd[x] = bd;
        // (X0, Y0)[row] interleaved inside the row's band block: one 64-bit load in the blend
        int* bandp = reinterpret_cast<int*>(rec + kCommonBytes + (x / kBandRows) * kBandBytes);
        bandp[2 * (x % kBandRows)] = X0;
        bandp[2 * (x % kBandRows) + 1] = Y0;
    }
    __syncwarp();

    // ---- staging plan ----
    const double ax = fabs(iM[0]), ay = fabs(iM[1]), bx = fabs(iM[3]), by = fabs(iM[4]);
    int wi16 = 0, hi16 = 0, wi8 = 0, hi8 = 0;
    const bool fit16 = plan_box<16>(ax, ay, bx, by, wi16, hi16);
    const bool fit8 = plan_box<8>(ax, ay, bx, by, wi8, hi8);
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
#ifdef FRP_ALIGN_FORCE_TH8
    const bool th16 = fit16 && !fit8;          // experiment: 8-row tiles wherever they fit
#else
    const bool th16 = fit16;
#endif
    const bool direct = !(a.use_tma && !far && (fit16 || fit8));
    const int wi = th16 ? wi16 : wi8, hi = th16 ? hi16 : hi8;
    const int th1 = th16 ? 15 : 7;
    const int xsel_x = (iM[0] >= 0.0) ? 0 : kTileW - 1, ysel_x = (iM[1] >= 0.0) ? 0 : th1;
    const int xsel_y = (iM[3] >= 0.0) ? 0 : kTileW - 1, ysel_y = (iM[4] >= 0.0) ? 0 : th1;
    if (lane == 0) {
        CropHdr h;
        int fi = a.frame_idx[crop];
        h.frame = min(max(fi, 0), a.F - 1);
        h.map = hi * kNumW + wi;
        h.pitch = box_w(wi);
        h.box_bytes = box_w(wi) * box_h(hi);
#ifdef FRP_ALIGN_NODUAL
        const bool dual = false;                   // A/B switch: single buffering only
#else
        const bool dual = !direct && h.box_bytes <= kHalfBuf;
#endif
        h.flags = (direct ? 1 : 0) | (dual ? 2 : 0) | (th16 ? 0 : 4);
        h.pad[0] = h.pad[1] = h.pad[2] = 0;
        *reinterpret_cast<CropHdr*>(rec) = h;
        if (direct) a.direct_list[atomicAdd(a.counter + 1, 1)] = crop;
    }
    // ---- source origin of every tile: the map is affine and the tables are monotone, so the
    // minimum corner of a tile's footprint is the sum of two table entries whose positions
    // depend only on the signs of the matrix entries ----
    const int TH = th16 ? 16 : 8;
    const int tiles_per_band = kTilesX * (kBandRows / TH);
    for (int t = lane; t < kBands * tiles_per_band; t += 32) {
        const int band = t / tiles_per_band, tb = t % tiles_per_band;
        const int x0 = (tb % kTilesX) * kTileW, y0 = band * kBandRows + (tb / kTilesX) * TH;
        const int Xmin = tab[2 * kOut + y0 + ysel_x] + tab[x0 + xsel_x];
        const int Ymin = tab[3 * kOut + y0 + ysel_y] + tab[kOut + x0 + xsel_y];
        int* org = reinterpret_cast<int*>(rec + kCommonBytes + band * kBandBytes + 128) + 2 * tb;
        org[0] = (3 * (Xmin >> 10)) & ~15;   // TMA needs the inner coordinate 16-byte aligned (floor, also for negatives)
        org[1] = Ymin >> 10;
    }
}

// ---------------------------------------------------------------------------------------------
// align_kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar_saddr, uint32_t dst_saddr, int c0,
                                            int c1, int c2) {
#ifdef FRP_ALIGN_ABL_NOTMA
    // ablation (timing only, wrong pixels): no staging; tile_expect completes the phase by itself
    (void)map; (void)bar_saddr; (void)dst_saddr; (void)c0; (void)c1; (void)c2;
    return;
#endif
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst_saddr), "l"(map), "r"(bar_saddr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_saddr), "r"(bytes) : "memory");
}
// arm a tile barrier for one TMA box
__device__ __forceinline__ void tile_expect(uint32_t bar_saddr, uint32_t bytes) {
#ifdef FRP_ALIGN_ABL_NOTMA
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
    (void)bytes;
#else
    mbar_expect_tx_a(bar_saddr, bytes);
#endif
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_saddr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}"
        ::"r"(bar_saddr), "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// third word of a tap row: only the start offset 3 (one pixel in four) reaches into it, so the
// load is predicated per lane — fewer shared-memory wavefronts (the bank conflicts of a warp
// load grow with the number of active lanes).  A lane that skips the load keeps `seed` (any
// defined value that is dead by now: it only feeds bytes the byte permutes never select), so
// the destination has no live range reaching back to an uninitialised register.
__device__ __forceinline__ uint32_t lds32_if3(uint32_t saddr, uint32_t off, uint32_t seed) {
#ifdef FRP_ALIGN_NOPRED
    return lds32(saddr);     // A/B switch: load the third word unconditionally
#endif
    uint32_t v = seed;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, 3;\n\t"
        "setp.eq.u32 p, t, 3;\n\t"
        "@p ld.shared.u32 %0, [%1];\n\t}"
        : "+r"(v) : "r"(saddr), "r"(off));
    return v;
}
__device__ __forceinline__ int2 lds64(uint32_t saddr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
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
__device__ __forceinline__ unsigned short norm_half(unsigned acc, unsigned mask, unsigned magic) {
    constexpr float c = (float)(1.0 / 127.5);
    // (acc & 0x7F8000) | 0x4B000000 as ONE three-input LOP3: mask and magic arrive in registers
    // (two immediates would split it into two ALU-pipe instructions, and the ALU pipe — LOP3 /
    // SHF / PRMT at half rate — is the busiest unit of this kernel)
    unsigned bits;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(bits) : "r"(acc), "r"(mask), "r"(magic));
    const float n = __fmaf_rn(__uint_as_float(bits), c * (1.0f / 32768.0f), -(1.0f + 256.0f * c));
    return __half_as_ushort(__float2half_rn(n));
}
// opaque register constants for norm_half (the compiler must not fold them back into immediates)
__device__ __forceinline__ void norm_consts(unsigned& mask, unsigned& magic) {
    asm volatile("mov.u32 %0, 0x7F8000;" : "=r"(mask));
    asm volatile("mov.u32 %0, 0x4B000000;" : "=r"(magic));
}

// BGR -> RGB planes; a half-warp writes 16 consecutive halfs (one 32-byte sector) per plane row
template <int NJ>
__device__ __forceinline__ void store_px(unsigned short* o16, uint8_t* o8, const unsigned (&vB)[NJ],
                                         const unsigned (&vG)[NJ], const unsigned (&vR)[NJ]) {
    unsigned mask, magic;
    norm_consts(mask, magic);
#ifdef FRP_ALIGN_ABL_NOSTORE
    // ablation (timing only): the values are computed, the plane stores are skipped unless a
    // sum takes a value it cannot have (keeps the blend alive)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
        if ((vR[j] | vG[j] | vB[j]) == 0xffffffffu)
            o16[j] = norm_half(vR[j], mask, magic) + norm_half(vG[j], mask, magic) + norm_half(vB[j], mask, magic);
#else
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        o16[(size_t)j * 2 * kOut] = norm_half(vR[j], mask, magic);
        o16[(size_t)kOut * kOut + (size_t)j * 2 * kOut] = norm_half(vG[j], mask, magic);
        o16[(size_t)2 * kOut * kOut + (size_t)j * 2 * kOut] = norm_half(vB[j], mask, magic);
    }
#endif
    if (o8) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            uint8_t* p = o8 + (size_t)j * 2 * kOut * 3;
            p[0] = (uint8_t)(vB[j] >> 15); p[1] = (uint8_t)(vG[j] >> 15); p[2] = (uint8_t)(vR[j] >> 15);
        }
    }
}

// One staged output tile, x-major lane mapping: lane -> column x0 + (lane & 15) and rows
// yr0 + 2j + (lane >> 4) of the band.  A warp-wide tap load then walks along ONE (rotated)
// source line per half-warp, which keeps shared-memory bank conflicts at <= ~2-way for any
// staged row pitch.  tabs: shared address of the item's table block.
template <int TH>
__device__ __forceinline__ void blend_tile_staged(uint32_t tabs, uint32_t buf_saddr, int c0, int sy0, int pitch,
                                                  int x0, int yr0, int lane, unsigned short* o16, uint8_t* o8) {
    constexpr int NJ = TH / 2;
    const int x = x0 + (lane & 15);
    const int yb = yr0 + (lane >> 4);
    const int ad = (int)lds32(tabs + 32 + 4 * x), bd = (int)lds32(tabs + 32 + 4 * (kOut + x));
    const uint32_t tsy = tabs + kCommonBytes + 8 * yb;            // (X0,Y0)[yb] of this band
    // shared address of source byte (0, 0) of the frame in the staged box coordinate system
    const int base = (int)buf_saddr - sy0 * pitch - c0;
    unsigned vB[NJ], vG[NJ], vR[NJ];      // un-shifted blend sums (pixel << 15 | fraction)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int2 xy = lds64(tsy + 16 * j);
        const int X = xy.x + ad, Y = xy.y + bd;          // 10 fractional bits
        const uint32_t off = (uint32_t)((Y >> 10) * pitch + 3 * (X >> 10) + base);   // byte address of the first tap
        uint32_t w0 = off & ~3u;
        uint32_t w1 = w0 + (uint32_t)pitch;
#ifdef FRP_ALIGN_DEBUG
        // address audit (debug builds only): record and clamp anything outside the staging buffer
        if (w0 < buf_saddr || w1 + 12 > buf_saddr + kBufBytes + 64) {
            int* dbg = g_align_dbg;
            const int slot = atomicAdd(dbg, 1);
            if (slot < 6) {
                int* d = dbg + 8 + slot * 8;
                d[0] = (int)(w0 - buf_saddr); d[1] = pitch; d[2] = c0; d[3] = sy0; d[4] = X; d[5] = Y; d[6] = x0 * 1000 + yr0 * 10 + TH / 8; d[7] = lane * 100 + j;
            }
            w0 = buf_saddr; w1 = buf_saddr;
        }
#endif
        const unsigned shf = off << 3;                    // funnel shifts use the low 5 bits
        const unsigned fx = (X >> 5) & 31, ty = Y & 0x3E0;
#ifdef FRP_ALIGN_ABL_NOTAPS
        // ablation (timing only, wrong pixels): no tap loads, operands from registers
        const unsigned a0 = w0, a1 = w1, a2 = (unsigned)X, b0 = (unsigned)Y, b1 = shf, b2 = w0 ^ w1;
#else
        const unsigned a0 = lds32(w0), a1 = lds32(w0 + 4), a2 = lds32_if3(w0 + 8, off, (unsigned)X);
        const unsigned b0 = lds32(w1), b1 = lds32(w1 + 4), b2 = lds32_if3(w1 + 8, off, (unsigned)Y);
#endif
        blend_px(__funnelshift_r(a0, a1, shf), __funnelshift_r(a1, a2, shf), __funnelshift_r(b0, b1, shf),
                 __funnelshift_r(b1, b2, shf), fx, ty, vB[j], vG[j], vR[j]);
    }
    store_px<NJ>(o16 + (size_t)yb * kOut + x, o8 ? o8 + ((size_t)yb * kOut + x) * 3 : nullptr, vB, vG, vR);
}

// The same tile sampled tap by tap from global memory (boxes beyond the ladder, coordinates near
// cv::remap's int16 saturation, frames that are not 16-byte aligned).  Rare: it runs in its own
// kernel (align_direct_kernel) so the staged kernel carries no call, no stack and no extra
// registers for it.  tab: adelta | bdelta of the crop, xy: (X0, Y0) of the band's 16 rows.
template <int TH>
__device__ __forceinline__ void blend_tile_direct(const AlignArgs& a, const uint8_t* frame, const int* tab,
                                                  const int2* xy, int x0, int yr0, int lane, unsigned short* o16,
                                                  uint8_t* o8) {
    constexpr int NJ = TH / 2;
    const int x = x0 + (lane & 15);
    const int yb = yr0 + (lane >> 4);
    const int ad = tab[x], bd = tab[kOut + x];
    unsigned vB[NJ], vG[NJ], vR[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int2 p = xy[yb + 2 * j];
        const int X = (p.x + ad) >> 5, Y = (p.y + bd) >> 5;
        const int sx = max(-32768, min(32767, X >> 5));
        const int sy = max(-32768, min(32767, Y >> 5));
        const unsigned t00 = tap_global(a, frame, sx, sy), t01 = tap_global(a, frame, sx + 1, sy);
        const unsigned t10 = tap_global(a, frame, sx, sy + 1), t11 = tap_global(a, frame, sx + 1, sy + 1);
        blend_px(t00 | (t01 << 24), t01 >> 8, t10 | (t11 << 24), t11 >> 8, X & 31, (Y & 31) << 5, vB[j], vG[j],
                 vR[j]);
    }
    store_px<NJ>(o16 + (size_t)yb * kOut + x, o8 ? o8 + ((size_t)yb * kOut + x) * 3 : nullptr, vB, vG, vR);
}

constexpr int kDirectWarps = 8;

// one warp per (direct crop, band), over the compact list the plan kernel wrote: a small fixed
// grid that exits at once when no crop needs it (the usual case)
constexpr int kDirectBlocks = 64;

__global__ void __launch_bounds__(kDirectWarps * 32) align_direct_kernel(const __grid_constant__ AlignArgs a) {
    __shared__ int s_tab[kDirectWarps][2 * kOut];
    __shared__ int2 s_xy[kDirectWarps][kBandRows];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = a.counter[1] * kBands;
    for (int item = blockIdx.x * kDirectWarps + warp; item < nitems; item += gridDim.x * kDirectWarps) {
    const int crop = a.direct_list[item / kBands], band = item % kBands;
    const uint8_t* rec = a.rec + (size_t)crop * kRecBytes;
    const CropHdr* h = reinterpret_cast<const CropHdr*>(rec);
    const int* src = reinterpret_cast<const int*>(rec + 32);
    __syncwarp();
    for (int i = lane; i < 2 * kOut; i += 32) s_tab[warp][i] = src[i];
    const int2* bxy = reinterpret_cast<const int2*>(rec + kCommonBytes + band * kBandBytes);
    if (lane < kBandRows) s_xy[warp][lane] = bxy[lane];
    __syncwarp();
    const uint8_t* frame = a.frames + (size_t)h->frame * a.frame_bytes;
    unsigned short* o16 = reinterpret_cast<unsigned short*>(a.out_blob) + (size_t)crop * 3 * kOut * kOut +
                          (size_t)band * kBandRows * kOut;
    uint8_t* o8 = a.out_aligned ? a.out_aligned + ((size_t)crop * kOut * kOut + (size_t)band * kBandRows * kOut) * 3
                                : nullptr;
    for (int t = 0; t < kTilesX; ++t)
        blend_tile_direct<16>(a, frame, s_tab[warp], s_xy[warp], t * kTileW, 0, lane, o16, o8);
    }
}

// all tiles of one band (16 output rows) of one crop with tile height TH
template <int TH>
__device__ __forceinline__ void run_band(const AlignArgs& a, const TmapSet& maps, const CropHdr& h, uint32_t tabs,
                                         uint32_t buf_saddr0, uint32_t bar_tile, uint32_t& parity, int lane,
                                         unsigned short* o16, uint8_t* o8) {
    constexpr int NT = kTilesX * (kBandRows / TH);       // 7 or 14 tiles
    const uint32_t orgs = tabs + kCommonBytes + 128;     // (c0, sy0) per tile
    if (h.flags & 1) return;          // direct crops are sampled by align_direct_kernel
    const CUtensorMap* map = &maps.m[h.map];
    const int pitch = h.pitch;
    if (!(h.flags & 2)) {
        // one buffer: the TMA latency of this warp is covered by the blends of the others
        for (int t = 0; t < NT; ++t) {
            const int2 org = lds64(orgs + 8 * t);
            if (lane == 0) {
                tile_expect(bar_tile, (uint32_t)h.box_bytes);
                tma_load_3d(map, bar_tile, buf_saddr0, org.x, org.y, h.frame);
            }
            mbar_wait_a(bar_tile, parity & 1u);
            parity ^= 1u;
            blend_tile_staged<TH>(tabs, buf_saddr0, org.x, org.y, pitch, (t % kTilesX) * kTileW, (t / kTilesX) * TH,
                                  lane, o16, o8);
            __syncwarp();
        }
        return;
    }
    // two half buffers: the next tile's box is in flight while this one is blended
    int2 cur = lds64(orgs);
    if (lane == 0) {
        tile_expect(bar_tile, (uint32_t)h.box_bytes);
        tma_load_3d(map, bar_tile, buf_saddr0, cur.x, cur.y, h.frame);
    }
#pragma unroll 1
    for (int t = 0; t < NT; ++t) {
        const int b = t & 1;
        // The table base is re-materialised through an opaque move every iteration.  Without it the
        // 8-row instantiation of this loop faulted ("illegal memory access") on every crop: its SASS
        // advances one register by 8 bytes per tile as the origin pointer (tabs + 8 t) AND uses
        // it as the base of the first row pair's (X0, Y0) load, while rows 1.. use a loop-invariant
        // copy — coordinates of row 0 came from the wrong table entries from tile 1 on
        // (tools/align_bisect.py found it; 16-row tiles compiled correctly and stayed bit-exact).
        uint32_t tabs_i = tabs;
        asm volatile("" : "+r"(tabs_i));
        const uint32_t orgs_i = tabs_i + kCommonBytes + 128;
        int2 nxt = cur;
        if (t + 1 < NT) {
            nxt = lds64(orgs_i + 8 * (t + 1));
            if (lane == 0) {
                tile_expect(bar_tile + 8 * (b ^ 1), (uint32_t)h.box_bytes);
                tma_load_3d(map, bar_tile + 8 * (b ^ 1), buf_saddr0 + (uint32_t)((b ^ 1) * kHalfBuf), nxt.x, nxt.y,
                            h.frame);
            }
        }
        mbar_wait_a(bar_tile + 8 * b, (parity >> b) & 1u);
        parity ^= (1u << b);
        blend_tile_staged<TH>(tabs_i, buf_saddr0 + (uint32_t)(b * kHalfBuf), cur.x, cur.y, pitch,
                              (t % kTilesX) * kTileW, (t / kTilesX) * TH, lane, o16, o8);
        __syncwarp();      // every lane is done with buffer b before it is refilled
        cur = nxt;
    }
}

// One warp per work item = (crop, band of 16 output rows), handed out by a dynamic counter two
// tickets ahead: while item i is blended, the plan record of item i+1 is already landing in the
// other table slot (two bulk copies: the crop's common block and the band block).
__global__ void __launch_bounds__(kWarps * 32, 1)
align_kernel(const __grid_constant__ TmapSet maps, const __grid_constant__ AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t wsm = ((smem_u32(smem_raw) + 127u) & ~127u) + (uint32_t)warp * kWarpSmem;
    const uint32_t tab0 = wsm;                              // two table slots of kItemBytes
    const uint32_t buf_al = wsm + kTabBytes;                // staging buffer (128-byte aligned)
    // [0,1] tile barriers, [2,3] item barriers.  64 bytes of slack separate them from the staging
    // buffer: the tap loads of a box that ends at the end of the buffer read up to 11 bytes
    // past it (never selected), and plain loads must not touch mbarrier objects
    const uint32_t bars = wsm + kTabBytes + kBufBytes + 64;
    if (lane == 0) {
        for (int i = 0; i < 4; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * i));
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t parity = 0;        // bits 0,1 tile barriers; bits 2,3 item barriers

    const int nitems = a.n * kBands;
    auto issue_item = [&](int item, int slot) {
        const uint8_t* rec = a.rec + (size_t)(item / kBands) * kRecBytes;
        const uint32_t bar = bars + 16 + 8 * slot;
        const uint32_t dst = tab0 + (uint32_t)slot * kItemBytes;
        mbar_expect_tx_a(bar, kItemBytes);
        bulk_g2s_a(dst, rec, kCommonBytes, bar);
        bulk_g2s_a(dst + kCommonBytes, rec + kCommonBytes + (item % kBands) * kBandBytes, kBandBytes, bar);
    };
    int t_cur = 0, t_next = 0;
    if (lane == 0) {
        t_cur = atomicAdd(a.counter, 1);
        t_next = atomicAdd(a.counter, 1);
        if (t_cur < nitems) issue_item(t_cur, 0);
    }
    t_cur = __shfl_sync(FRP_FULL_MASK, t_cur, 0);
    t_next = __shfl_sync(FRP_FULL_MASK, t_next, 0);
    int slot = 0;
    while (t_cur < nitems) {
        int t_next2 = 0;
        if (lane == 0) {
            if (t_next < nitems) issue_item(t_next, slot ^ 1);     // the other slot's item is finished
            t_next2 = atomicAdd(a.counter, 1);                   // round trip overlaps this item
        }
        const uint32_t tabs = tab0 + (uint32_t)slot * kItemBytes;
        mbar_wait_a(bars + 16 + 8 * slot, (parity >> (2 + slot)) & 1u);
        parity ^= (4u << slot);
        CropHdr h;
        {
            const int2 h0 = lds64(tabs), h1 = lds64(tabs + 8);
            h.frame = h0.x; h.map = h0.y; h.pitch = h1.x; h.box_bytes = h1.y; h.flags = (int)lds32(tabs + 16);
        }
        const int crop = t_cur / kBands, band = t_cur % kBands;
        unsigned short* o16 = reinterpret_cast<unsigned short*>(a.out_blob) + (size_t)crop * 3 * kOut * kOut +
                              (size_t)band * kBandRows * kOut;
        uint8_t* o8 = a.out_aligned ? a.out_aligned + ((size_t)crop * kOut * kOut + (size_t)band * kBandRows * kOut) * 3
                                    : nullptr;
        if (h.flags & 4) run_band<8>(a, maps, h, tabs, buf_al, bars, parity, lane, o16, o8);
        else run_band<16>(a, maps, h, tabs, buf_al, bars, parity, lane, o16, o8);
        __syncwarp();
        t_cur = t_next;
        t_next = __shfl_sync(FRP_FULL_MASK, t_next2, 0);
        slot ^= 1;
    }
}

__global__ void estimate_norm_kernel(const float* __restrict__ lmk, int n, double* __restrict__ M, int f32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lm[10];
    for (int j = 0; j < 10; ++j) lm[j] = lmk[(size_t)i * 10 + j];
    double m[6];
    if (f32) estimate_norm_f32_dev(lm, m);
    else estimate_norm_dev(lm, m);
    for (int j = 0; j < 6; ++j) M[(size_t)i * 6 + j] = m[j];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// u8 tensor [F][H][3W bytes] with pitches (stride, frame_bytes); one map per box size.  The 66
// descriptors of a frame buffer are cached per thread (encoding them costs ~30 us): a camera
// pipeline passes the same device buffer tick after tick.
struct MapCache {
    const uint8_t* frames = nullptr;
    int F = 0, H = 0, W = 0;
    size_t stride = 0;
    bool ok = false;
    TmapSet set;
};

bool encode_frame_maps(MapCache* mc, const uint8_t* frames, int F, int H, int W, size_t stride, size_t frame_bytes) {
    if (mc->frames == frames && mc->F == F && mc->H == H && mc->W == W && mc->stride == stride) return mc->ok;
    mc->frames = frames; mc->F = F; mc->H = H; mc->W = W; mc->stride = stride; mc->ok = false;
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tensormap_encode_fn());
    if (!fn) return false;
    if (((uintptr_t)frames & 15) || (stride & 15) || (frame_bytes & 15)) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)3 * W, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t gstr[2] = {(cuuint64_t)stride, (cuuint64_t)frame_bytes};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int hi = 0; hi < kNumH; ++hi)
        for (int wi = 0; wi < kNumW; ++wi) {
            cuuint32_t box[3] = {(cuuint32_t)box_w(wi), (cuuint32_t)box_h(hi), 1};
            CUresult r = fn(&mc->set.m[hi * kNumW + wi], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(frames),
                            gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return false;
        }
    mc->ok = true;
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
    // workspace: the work counter (first 256 bytes) and the per-crop plan records
    uint8_t* ws = nullptr;
    // persistent per-(thread, stream) block: the steady state of a camera pipeline allocates nothing
    const size_t list_bytes = (((size_t)n * sizeof(int)) + 255) & ~(size_t)255;
    Workspace* wsp = workspace_for(device, st);
    {
        const int rc = workspace_reserve(wsp, 256 + list_bytes + (size_t)n * kRecBytes, 0);
        if (rc) return rc;
    }
    ws = wsp->dev;
    int* counter = reinterpret_cast<int*>(ws);
    FRP_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));
    AlignArgs a;
    a.frames = frames; a.F = F; a.H = H; a.W = W; a.stride = stride;
    a.frame_bytes = (size_t)H * stride;
    a.frame_idx = frame_idx; a.lmk = lmk; a.M = M; a.n = n;
    a.out_blob = reinterpret_cast<__half*>(out_blob);
    a.out_aligned = out_aligned;
    a.counter = counter;
    a.direct_list = reinterpret_cast<int*>(ws + 256);
    a.rec = ws + 256 + list_bytes;
    static thread_local MapCache mc;
    if (!encode_frame_maps(&mc, frames, F, H, W, stride, a.frame_bytes)) memset(&mc.set, 0, sizeof(mc.set));
    a.use_tma = mc.ok ? 1 : 0;
    const size_t smem = (size_t)kWarps * kWarpSmem + 128;
    {
        int rc = FRP_OK;
        configure_func_smem((const void*)align_kernel, device, smem, &rc);
        if (rc) return rc;
    }
    const int items = n * kBands;
    int grid = (items + kWarps - 1) / kWarps;
    if (grid > s_sms[di]) grid = s_sms[di];
    if (prof) {
        if (!t_align_ev[0]) { cudaEventCreate(&t_align_ev[0]); cudaEventCreate(&t_align_ev[1]); }
        cudaEventRecord(t_align_ev[0], st);
    }
    align_plan_kernel<<<(n + kPlanWarps - 1) / kPlanWarps, kPlanWarps * 32, 0, st>>>(a);
    FRP_CHECK_LAUNCH();
    align_kernel<<<grid, kWarps * 32, smem, st>>>(mc.set, a);
    FRP_CHECK_LAUNCH();
    align_direct_kernel<<<kDirectBlocks, kDirectWarps * 32, 0, st>>>(a);
    FRP_CHECK_LAUNCH();
#ifdef FRP_ALIGN_DEBUG
    {
        int h_dbg[64];
        cudaError_t de = cudaStreamSynchronize(st);
        fprintf(stderr, "[align debug] sync: %s\n", cudaGetErrorString(de));
        if (cudaMemcpyFromSymbol(h_dbg, g_align_dbg, sizeof(h_dbg)) == cudaSuccess) {
            fprintf(stderr, "[align debug] out-of-buffer tap addresses: %d\n", h_dbg[0]);
            for (int i = 0; i < 6 && i < h_dbg[0]; ++i) {
                const int* d = h_dbg + 8 + i * 8;
                fprintf(stderr, "  rel=%d pitch=%d c0=%d sy0=%d X=%d Y=%d tile=%d lane_j=%d\n", d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
            }
            memset(h_dbg, 0, sizeof(h_dbg));
            cudaMemcpyToSymbol(g_align_dbg, h_dbg, sizeof(h_dbg));
        }
    }
#endif
    if (prof) { cudaEventRecord(t_align_ev[1], st); t_align_valid = true; }
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

static int estimate_norm_host(const float* lmk, int n, double* M_out, int f32) {
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
        estimate_norm_kernel<<<(n + 127) / 128, 128>>>(d_l, n, d_m, f32);
        e = cudaGetLastError();
        count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpy(M_out, d_m, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_l); cudaFree(d_m);
    if (e != cudaSuccess) { set_error("frp_estimate_norm failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return FRP_E_CUDA; }
    return FRP_OK;
}

int frp_estimate_norm(const float* lmk, int n, double* M_out) { return estimate_norm_host(lmk, n, M_out, 0); }
int frp_estimate_norm_f32(const float* lmk, int n, double* M_out) { return estimate_norm_host(lmk, n, M_out, 1); }

}  // extern "C"
