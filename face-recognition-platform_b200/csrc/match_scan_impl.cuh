// K5 — streaming 1:N scan matcher for small query batches (B <= 4 per pass).
//
// Replaces `face_recognition.face_distance(stored, q)` + `<= tolerance`
// (backend/app/services/face_service.py:410-411) and the cosine 1:N of
// insightface compute_sim for the memory-bound regime: the gallery is read from
// HBM exactly once per pass (algorithmic bytes = rows * dimp * sizeof(T)).
//
// Structure (one persistent CTA per SM):
//   warp 8 (producer): 1-D bulk async copies (TMA engine, cp.async.bulk) of
//       32 KB gallery chunks into a shared-memory ring, mbarrier-signalled.
//   warps 0..7 (consumers): each owns RW=4 rows of a chunk; a lane reads its
//       16-byte slices of the row from shared memory, accumulates QB dot
//       products in fp32, a transposed warp reduction leaves (row,query) totals
//       on distinct lanes; score -> cut test -> append to a per-warp candidate
//       buffer in shared memory, compacted to the running top-KP by a warp
//       bitonic sort when it fills.  Distances never go to HBM.
//   At the end the 8 warps' lists are merged to one top-KP list per query per
//   CTA: cand[B][nlists][KP] (u64 keys), consumed by the re-rank kernel.
#pragma once
#include "frp_device.cuh"
#include "frp_internal.h"
#include "match_scan.h"

namespace frp {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kRW = 4;                             // rows per warp per chunk
constexpr int kChunkRows = kConsumerWarps * kRW;   // 32 rows per stage
static_assert(kChunkRows == kScanChunkRows, "chunk size mismatch");
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;

template <int KP> struct CapOf { static constexpr int value = (KP <= 32) ? 64 : 2 * KP; };

template <typename T> struct Vec;
template <> struct Vec<__half> { static constexpr int n = 8; };
template <> struct Vec<float> { static constexpr int n = 4; };

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
    constexpr int VEC = Vec<T>::n;
    constexpr int UNITS = DIMP / VEC;                       // 16-byte units per row
    constexpr int UPL = (UNITS + 31) / 32;                  // units per lane
    constexpr int ROWBYTES = DIMP * (int)sizeof(T);
    constexpr int STAGE_BYTES = kChunkRows * ROWBYTES;
    constexpr int CAP = CapOf<KP>::value;
    constexpr int NV = kRW * QB;

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
                const uint32_t row0 = c * kChunkRows;
                const uint32_t nr = min((uint32_t)kChunkRows, a.rows - row0);
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
    // query slices in registers: element e of unit (lane + 32 j) is q[(lane+32j)*VEC + t]
    float qreg[QB][UPL * VEC];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        const int qg = min(qbase + qi, a.B - 1);
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            const int u = lane + 32 * j;
#pragma unroll
            for (int t = 0; t < VEC; ++t)
                qreg[qi][j * VEC + t] = (u < UNITS) ? a.q[(size_t)qg * DIMP + u * VEC + t] : 0.f;
        }
    }
    // lane l < NV finishes value index l = r*QB + qi
    const int my_r = (lane & (NV - 1)) / QB;
    const int my_qi = (lane & (NV - 1)) % QB;
    const bool my_valid_q = (qbase + my_qi) < a.B;
    float my_cut = my_valid_q ? a.qp[qbase + my_qi].cut0 : __int_as_float(0x7f800000);
    uint64_t* my_buf = wbuf + ((size_t)warp * QB + my_qi) * CAP;
    int* my_cnt = wcnt + warp * QB + my_qi;
    const bool is_l2 = (a.metric == FRP_METRIC_L2);

    int s = 0;
    uint32_t phase = 0;
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint32_t row_l = c * kChunkRows + warp * kRW + my_r;   // this lane's result row
        float n2 = 0.f;
        if (is_l2 && lane < NV && row_l < a.rows) n2 = __ldg(a.norm2 + row_l);

        mbar_wait(&full_bar[s], phase);
        const uint8_t* base = ring + (size_t)s * STAGE_BYTES + (size_t)(warp * kRW) * ROWBYTES;
        float acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.f;
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            const int u = lane + 32 * j;
            if (UNITS >= 32 * (j + 1) || u < UNITS) {
                uint4 v[kRW];
#pragma unroll
                for (int r = 0; r < kRW; ++r)
                    v[r] = *reinterpret_cast<const uint4*>(base + r * ROWBYTES + u * 16);
#pragma unroll
                for (int r = 0; r < kRW; ++r) {
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
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // stage consumed (data is in registers)
        if (++s == stages) { s = 0; phase ^= 1; }

        float sc = warp_multi_reduce<NV>(acc, lane);
        if (is_l2) sc = 2.f * sc - n2;
        const bool pass = (lane < NV) && (row_l < a.rows) && (sc >= my_cut);
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
                if (n > CAP - kRW) {   // warp-uniform
                    float cut = 0.f;
                    bool upd = false;
                    float nc = 0.f;
                    __syncwarp();
                    const int kept = warp_refine<KP, CAP>(wbuf + ((size_t)warp * QB + qi) * CAP, n,
                                                          lane, &nc);
                    upd = (n >= KP);
                    cut = nc;
                    if (lane == 0) *cnt = kept;
                    if (upd && my_qi == qi && cut > my_cut) my_cut = cut;
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

template <typename T, int DIMP>
constexpr int stage_bytes() { return kChunkRows * DIMP * (int)sizeof(T); }

template <typename T, int DIMP, int QB, int KP>
int launch_one(const ScanArgs& a, int nlists, int qgroups, cudaStream_t st) {
    constexpr int CAP = CapOf<KP>::value;
    constexpr int SB = stage_bytes<T, DIMP>();
    int stages = kSmemBudget / SB;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    const size_t smem = (size_t)stages * SB + 2 * kMaxStages * sizeof(uint64_t) +
                        (size_t)kConsumerWarps * QB * CAP * sizeof(uint64_t) +
                        kConsumerWarps * QB * sizeof(int) + 16;
    auto kern = match_scan_kernel<T, DIMP, QB, KP>;
    static thread_local size_t configured = 0;   // per instantiation & thread; cheap to repeat
    if (configured < smem) {
        FRP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    ScanArgs b = a;
    b.stages = stages;
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
