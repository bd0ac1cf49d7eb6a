// K5 — streaming 1:N scan matcher for small query batches (B <= 4 per pass).
//
// Replaces `face_recognition.face_distance(stored, q)` + `<= tolerance`
// (backend/app/services/face_service.py:410-411) and the cosine 1:N of
// insightface compute_sim for the memory-bound regime: the gallery is read from
// HBM exactly once per pass (algorithmic bytes = rows * dimp * sizeof(T)).
//
// Structure (one persistent CTA per SM):
//   warp 8 (producer): 1-D bulk async copies (TMA engine, cp.async.bulk) of
//       16-32 KB gallery chunks into a shared-memory ring, mbarrier-signalled.
//   warps 0..7 (consumers): a row is spread over LPR = min(32, row bytes / 16) lanes,
//       so short rows (128-d fp16 = 16 lanes) are processed 32/LPR at a time; a lane
//       reads its 16-byte slices from shared memory, accumulates QB dot products in
//       fp32 for RW row slots, a transposed reduction inside each LPR-lane group leaves
//       the (row, query) totals on distinct lanes; score -> cut test -> append to a
//       per-warp candidate buffer in shared memory, compacted to the running top-KP by
//       a warp bitonic sort when it fills.  Distances never go to HBM.
//   At the end the 8 warps' lists are merged to one top-KP list per query per
//   CTA: cand[B][nlists][KP] (u64 keys), consumed by the re-rank kernel.
#pragma once
#include "frp_device.cuh"
#include "frp_internal.h"
#include "match_scan.h"

namespace frp {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;

template <int KP> struct CapOf { static constexpr int value = (KP <= 32) ? 64 : 2 * KP; };

template <typename T> struct Vec;
template <> struct Vec<__half> { static constexpr int n = 8; };
template <> struct Vec<float> { static constexpr int n = 4; };

// compile-time geometry of one instantiation
template <typename T, int DIMP, int QB>
struct Geo {
    static constexpr int VEC = Vec<T>::n;
    static constexpr int UNITS = DIMP / VEC;                       // 16-byte units per row
    static constexpr int LPR = UNITS < 32 ? UNITS : 32;            // lanes per row
    static constexpr int RPS = 32 / LPR;                           // rows side by side in a warp
    static constexpr int UPL = UNITS / LPR;                        // units per lane
    static constexpr int ROWBYTES = DIMP * (int)sizeof(T);
    // row slots per warp per chunk: 8 for short rows (<= 512 B) so a stage stays ~32 KB and the
    // per-chunk barrier/reduction overhead is spread over more bytes, 4 for long rows
    static constexpr int RWMAX = (ROWBYTES <= 512) ? 8 : 4;
    static constexpr int RW = (RWMAX * QB <= LPR) ? RWMAX : (LPR / QB);   // keeps NV <= LPR
    static constexpr int NV = RW * QB;                             // (slot, query) values per lane group
    static constexpr int WARP_ROWS = RW * RPS;                     // rows per warp per chunk
    static constexpr int CHUNK_ROWS = kConsumerWarps * WARP_ROWS;  // rows per shared-memory stage
    static constexpr int STAGE_BYTES = CHUNK_ROWS * ROWBYTES;
    static_assert(RW >= 1 && NV <= LPR && UNITS == LPR * UPL, "unsupported scan geometry");
};

template <typename T, int VEC>
__device__ __forceinline__ void unpack16(const uint4& v, float (&f)[VEC]);
template <>
__device__ __forceinline__ void unpack16<__half, 8>(const uint4& v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
template <>
__device__ __forceinline__ void unpack16<float, 4>(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}

// Sort the first n (<= CAP) keys of buf descending, keep KP, return new count.
template <int KP, int CAP>
__device__ __forceinline__ int warp_refine(uint64_t* buf, int n, int lane, float* new_cut) {
    constexpr int E = CAP / 32;
    uint64_t k[E];
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        k[j] = (i < n) ? buf[i] : 0ull;
    }
    warp_sort_desc<E>(k, lane);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < E; ++j) {
        const int i = j * 32 + lane;
        if (i < KP) buf[i] = k[j];
    }
    // key at position KP-1 lives in register (KP-1)/32 of lane (KP-1)%32
    const uint64_t kth = __shfl_sync(FRP_FULL_MASK, k[(KP - 1) / 32], (KP - 1) % 32);
    if (n >= KP) *new_cut = key_score(kth);
    __syncwarp();
    return n < KP ? n : KP;
}

template <typename T, int DIMP, int QB, int KP>
__global__ void __launch_bounds__(kThreads, 1) match_scan_kernel(const ScanArgs a) {
    using G = Geo<T, DIMP, QB>;
    constexpr int VEC = G::VEC, LPR = G::LPR, RPS = G::RPS, UPL = G::UPL, RW = G::RW, NV = G::NV;
    constexpr int WARP_ROWS = G::WARP_ROWS, CHUNK_ROWS = G::CHUNK_ROWS;
    constexpr int ROWBYTES = G::ROWBYTES, STAGE_BYTES = G::STAGE_BYTES;
    constexpr int CAP = CapOf<KP>::value;
    static_assert(WARP_ROWS <= CAP - KP, "candidate buffer too small for the chunk");

    extern __shared__ __align__(128) uint8_t smem[];
    const int stages = a.stages;
    uint8_t* ring = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* wbuf = empty_bar + kMaxStages;                 // [8 warps][QB][CAP]
    int* wcnt = reinterpret_cast<int*>(wbuf + kConsumerWarps * QB * CAP);  // [8][QB]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qbase = blockIdx.y * QB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    if (threadIdx.x < kConsumerWarps * QB) wcnt[threadIdx.x] = 0;
    __syncthreads();

    const uint32_t nchunks = a.nchunks;
    if (warp == kConsumerWarps) {
        // ===== producer: stream this CTA's chunks through the ring =====
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
                mbar_wait(&empty_bar[s], phase ^ 1);
                const uint32_t row0 = c * CHUNK_ROWS;
                const uint32_t nr = min((uint32_t)CHUNK_ROWS, a.rows - row0);
                const uint32_t bytes = nr * ROWBYTES;
                mbar_expect_tx(&full_bar[s], bytes);
                bulk_g2s(ring + (size_t)s * STAGE_BYTES,
                         reinterpret_cast<const uint8_t*>(a.g) + (size_t)row0 * ROWBYTES, bytes,
                         &full_bar[s]);
                if (++s == stages) { s = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===== consumers =====
    const int sub = lane / LPR;          // which of the RPS side-by-side rows
    const int lu = lane % LPR;           // unit index inside the row (first of UPL)
    // query slices in registers: element t of unit (lu + LPR j) is q[(lu + LPR j)*VEC + t]
    float qreg[QB][UPL * VEC];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        const int qg = min(qbase + qi, a.B - 1);
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            const int u = lu + LPR * j;
#pragma unroll
            for (int t = 0; t < VEC; ++t) qreg[qi][j * VEC + t] = a.q[(size_t)qg * DIMP + u * VEC + t];
        }
    }
    // inside its lane group, lane lu < NV finishes value index lu = slot*QB + qi
    const bool res_lane = lu < NV;
    const int my_slot = (lu & (NV - 1)) / QB;
    const int my_qi = (lu & (NV - 1)) % QB;
    const int my_row_in_warp = my_slot * RPS + sub;
    const bool my_valid_q = (qbase + my_qi) < a.B;
    float my_cut = my_valid_q ? a.qp[qbase + my_qi].cut0 : __int_as_float(0x7f800000);
    uint64_t* my_buf = wbuf + ((size_t)warp * QB + my_qi) * CAP;
    int* my_cnt = wcnt + warp * QB + my_qi;
    const bool is_l2 = (a.metric == FRP_METRIC_L2);

    // |g|^2 of this lane's result row, fetched one chunk ahead so the HBM latency of the
    // norm load never sits on a chunk's critical path
    const uint32_t row_in_chunk = warp * WARP_ROWS + my_row_in_warp;
    auto load_n2 = [&](uint32_t chunk) -> float {
        const uint64_t r = (uint64_t)chunk * CHUNK_ROWS + row_in_chunk;
        return (is_l2 && res_lane && chunk < nchunks && r < a.rows) ? __ldg(a.norm2 + r) : 0.f;
    };
    float n2_next = load_n2(blockIdx.x);

    int s = 0;
    uint32_t phase = 0;
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint32_t row_l = c * CHUNK_ROWS + row_in_chunk;   // this lane's result row
        const float n2 = n2_next;
        n2_next = load_n2(c + gridDim.x);

        mbar_wait(&full_bar[s], phase);
        const uint8_t* base = ring + (size_t)s * STAGE_BYTES + (size_t)(warp * WARP_ROWS + sub) * ROWBYTES;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            const int u = lu + LPR * j;
            uint4 v[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r)
                v[r] = *reinterpret_cast<const uint4*>(base + (size_t)(r * RPS) * ROWBYTES + u * 16);
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                float f[VEC];
                unpack16<T, VEC>(v[r], f);
#pragma unroll
                for (int qi = 0; qi < QB; ++qi) {
#pragma unroll
                    for (int t = 0; t < VEC; ++t)
                        acc[r * QB + qi] = fmaf(f[t], qreg[qi][j * VEC + t], acc[r * QB + qi]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // stage consumed (data is in registers)
        if (++s == stages) { s = 0; phase ^= 1; }

        float sc = warp_multi_reduce<NV, LPR>(acc, lane);
        if (is_l2) sc = 2.f * sc - n2;
        const bool pass = res_lane && (row_l < a.rows) && (sc >= my_cut);
        const unsigned m = __ballot_sync(FRP_FULL_MASK, pass);
        if (m) {
            if (pass && a.alive[row_l]) {
                const int pos = atomicAdd(my_cnt, 1);
                my_buf[pos] = make_key(sc, row_l);
            }
            __syncwarp();
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                int* cnt = wcnt + warp * QB + qi;
                const int n = *cnt;
                if (n > CAP - WARP_ROWS) {   // warp-uniform: the next chunk could overflow the buffer
                    float nc = 0.f;
                    __syncwarp();
                    const int kept = warp_refine<KP, CAP>(wbuf + ((size_t)warp * QB + qi) * CAP, n, lane, &nc);
                    if (lane == 0) *cnt = kept;
                    if (n >= KP && my_qi == qi && nc > my_cut) my_cut = nc;
                    __syncwarp();
                }
            }
        }
    }

    // ===== final: per-warp top-KP, then merge the 8 warps' lists per query =====
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        int* cnt = wcnt + warp * QB + qi;
        const int n = *cnt;
        float nc;
        __syncwarp();
        const int kept = warp_refine<KP, CAP>(wbuf + ((size_t)warp * QB + qi) * CAP, n, lane, &nc);
        // zero the tail so the merge can read KP slots unconditionally
        for (int i = kept + lane; i < KP; i += 32) wbuf[((size_t)warp * QB + qi) * CAP + i] = 0ull;
        if (lane == 0) *cnt = kept;
    }
    named_bar_sync(1, kConsumerWarps * 32);
    if (warp < QB && (qbase + warp) < a.B) {
        constexpr int E = (kConsumerWarps * KP) / 32;
        uint64_t k[E];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int i = j * 32 + lane;           // i in [0, 8*KP)
            k[j] = wbuf[((size_t)(i / KP) * QB + warp) * CAP + (i % KP)];
        }
        warp_sort_desc<E>(k, lane);
        uint64_t* out = a.cand + ((size_t)(qbase + warp) * gridDim.x + blockIdx.x) * KP;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int i = j * 32 + lane;
            if (i < KP) out[i] = k[j];
        }
    }
}

template <typename T, int DIMP, int QB, int KP>
int launch_one(const ScanArgs& a, int nlists, int qgroups, cudaStream_t st) {
    using G = Geo<T, DIMP, QB>;
    constexpr int CAP = CapOf<KP>::value;
    constexpr int SB = G::STAGE_BYTES;
    int stages = kSmemBudget / SB;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    const size_t smem = (size_t)stages * SB + 2 * kMaxStages * sizeof(uint64_t) +
                        (size_t)kConsumerWarps * QB * CAP * sizeof(uint64_t) +
                        kConsumerWarps * QB * sizeof(int) + 16;
    auto kern = match_scan_kernel<T, DIMP, QB, KP>;
    {
        int rc = FRP_OK;
        configure_func_smem((const void*)kern, a.device, smem, &rc);
        if (rc) return rc;
    }
    ScanArgs b = a;
    b.stages = stages;
    // chunk size depends on the instantiation; CTAs beyond the last chunk just emit empty lists
    // (gridDim.x must stay `nlists`: it is the candidate stride the re-rank kernel reads)
    b.nchunks = (uint32_t)((a.rows + G::CHUNK_ROWS - 1) / G::CHUNK_ROWS);
    dim3 grid(nlists, qgroups);
    kern<<<grid, kThreads, smem, st>>>(b);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

template <typename T, int DIMP>
int dispatch_qb_kp(const ScanArgs& a, int B, int KP, int nlists, cudaStream_t st) {
    // QB: queries per pass.  4 keeps the FMA work under the HBM time for fp16 rows.
    const int QB = (B >= 4) ? 4 : (B >= 2 ? 2 : 1);
    const int qgroups = (B + QB - 1) / QB;
#define FRP_SCAN_CASE(QB_, KP_)                                                   \
    if (QB == QB_ && KP == KP_) return launch_one<T, DIMP, QB_, KP_>(a, nlists, qgroups, st);
    FRP_SCAN_CASE(1, 16) FRP_SCAN_CASE(2, 16) FRP_SCAN_CASE(4, 16)
    FRP_SCAN_CASE(1, 32) FRP_SCAN_CASE(2, 32) FRP_SCAN_CASE(4, 32)
    FRP_SCAN_CASE(1, 64) FRP_SCAN_CASE(2, 64) FRP_SCAN_CASE(4, 64)
#undef FRP_SCAN_CASE
    set_error("match_scan: unsupported KP=%d", KP);
    return FRP_E_UNSUPPORTED;
}

}  // namespace
}  // namespace frp

// one translation unit per (element type, padded dim) keeps the build parallel
#define FRP_SCAN_INSTANTIATE(NAME, T, DIMP)                                              \
    namespace frp {                                                                      \
    int NAME(const ScanArgs& a, int B, int KP, int nlists, cudaStream_t st) {            \
        return dispatch_qb_kp<T, DIMP>(a, B, KP, nlists, st);                            \
    }                                                                                    \
    }
