// Match -> tracking hand-off on the device.
//
// backend/app/services/tracking_service.py:118-134 drops a detection of (person, camera)
// that arrives within `cooldown` seconds of the last RECORDED one (timestamp stored at :179);
// backend/app/routes/camera.py:250 keeps a match only if `distance <= confidence_threshold`.
// Both filters run here on the match results while they are still in HBM: an open-addressing
// hash table keyed by (global gallery row, camera id) holds the last recorded time; hits that
// pass are compacted in (query, rank) order into frp_hit_t records, so cooled-down or
// below-confidence hits are never copied to the host or turned into Python objects.
#include <string.h>

#include <mutex>
#include <shared_mutex>

#include "frp_device.cuh"
#include "frp_internal.h"

struct frp_cooldown {
    int device = 0;
    uint64_t cap = 0;            // slots, power of two
    unsigned long long* keys = nullptr;   // 0 = empty, else (row << 24 | camera) + 1
    double* last = nullptr;      // last recorded time of the slot's key
    unsigned int* winner = nullptr;  // per-call arbitration: lowest (b*k+j) claiming the slot
    std::shared_mutex mu;        // filter calls mutate the table: unique
};

namespace frp {
namespace {

constexpr int kFilterThreads = 1024;
constexpr int kMaxPer = 16;                         // entries per thread: B*k <= 16384
constexpr uint32_t kNoSlot = 0xffffffffu;
constexpr int kMaxProbe = 128;
constexpr double kNever = -1.0e300;

__global__ void cooldown_init_kernel(unsigned long long* keys, double* last, unsigned int* winner, uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        keys[i] = 0ull;
        last[i] = kNever;
        winner[i] = 0xffffffffu;
    }
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// ONE block: phase 1 claims slots (first hit in (b, j) order wins a key), phase 2 decides and
// stamps the table, phase 3 compacts the recorded hits in order.  Entry e = b*k + j; thread t
// owns the contiguous entries [t*per, (t+1)*per), so a block-wide exclusive scan of the
// per-thread keep counts gives the ordered output position.
__global__ void __launch_bounds__(kFilterThreads)
cooldown_filter_kernel(unsigned long long* __restrict__ keys, double* __restrict__ last,
                       unsigned int* __restrict__ winner, uint64_t cap, int metric,
                       const int64_t* __restrict__ idx, const float* __restrict__ score,
                       const int32_t* __restrict__ count, const int32_t* __restrict__ status,
                       const int32_t* __restrict__ camera, const float* __restrict__ limit, int B, int k,
                       double now, double cooldown, frp_hit_t* __restrict__ hits, int32_t* __restrict__ n_hits,
                       int32_t* __restrict__ overflow) {
    __shared__ int s_warp[kFilterThreads / 32];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int total = B * k;
    const int per = (total + kFilterThreads - 1) / kFilterThreads;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // uncertified queries must go through the caller's fallback ladder first: report and
    // leave the table untouched
    if (status) {
        for (int b = tid; b < B; b += kFilterThreads)
            if (status[b] != FRP_STATUS_OK) atomicOr(&s_bad, 1);
        __syncthreads();
        if (s_bad) {
            if (tid == 0) { *n_hits = -1; if (overflow) *overflow = 0; }
            return;
        }
    }
    uint32_t slot[kMaxPer];
    uint8_t state[kMaxPer];      // 0 = not a hit, 1 = candidate with slot, 2 = record without slot (table full)
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        slot[i] = kNoSlot;
        state[i] = 0;
        const int e = tid * per + i;
        if (i >= per || e >= total) continue;
        const int b = e / k, j = e % k;
        if (j >= count[b]) continue;
        const float v = score[e];
        if (limit) {
            const float lim = limit[b];
            if (lim == lim) {    // NaN = no second filter for this query
                const bool pass = (metric == FRP_METRIC_COS) ? (v >= lim) : (v <= lim);
                if (!pass) continue;
            }
        }
        const unsigned long long key = (((unsigned long long)idx[e] << 24) | ((unsigned long long)camera[b] & 0xffffffull)) + 1ull;
        uint64_t h = mix64(key) & (cap - 1);
        bool found = false;
        for (int p = 0; p < kMaxProbe; ++p) {
            const unsigned long long cur = atomicCAS(&keys[h], 0ull, key);
            if (cur == 0ull || cur == key) { found = true; break; }
            h = (h + 1) & (cap - 1);
        }
        if (!found) { state[i] = 2; if (overflow) atomicOr(overflow, 1); continue; }
        slot[i] = (uint32_t)h;
        // `last` is only written in phase 2, after the barrier: this read sees the previous calls
        const double t_last = last[h];
        if (now - t_last < cooldown) continue;            // duplicate within the cooldown: dropped
        state[i] = 1;
        atomicMin(&winner[h], (unsigned int)e);
    }
    __threadfence();
    __syncthreads();
    int nkeep = 0;
    uint32_t keepmask = 0;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        const int e = tid * per + i;
        bool keep = state[i] == 2;
        // cooldown <= 0: `now - last < cooldown` is never true, every hit is recorded
        if (state[i] == 1 && (winner[slot[i]] == (unsigned int)e || cooldown <= 0.0)) keep = true;
        if (keep) { keepmask |= 1u << i; ++nkeep; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        if (slot[i] != kNoSlot) {
            if ((keepmask >> i) & 1u) last[slot[i]] = now;
            winner[slot[i]] = 0xffffffffu;                // ready for the next call
        }
    }
    // ordered compaction
    int incl = nkeep;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int o = __shfl_up_sync(FRP_FULL_MASK, incl, s);
        if (lane >= s) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int o = __shfl_up_sync(FRP_FULL_MASK, w, s);
            if (lane >= s) w += o;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    int pos = incl - nkeep + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        if ((keepmask >> i) & 1u) {
            const int e = tid * per + i;
            frp_hit_t hrec;
            hrec.query = e / k; hrec.rank = e % k; hrec.row = idx[e]; hrec.score = score[e];
            hrec.camera = camera[e / k];
            hits[pos++] = hrec;
        }
    }
    if (tid == kFilterThreads - 1) *n_hits = s_warp[kFilterThreads / 32 - 1];
}

int filter_launch(frp_cooldown* c, int metric, const int64_t* idx, const float* score, const int32_t* count,
                  const int32_t* status, const int32_t* camera, const float* limit, int B, int k, double now,
                  double cooldown, frp_hit_t* hits, int32_t* n_hits, int32_t* overflow, cudaStream_t st) {
    if (overflow) FRP_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int32_t), st));
    cooldown_filter_kernel<<<1, kFilterThreads, 0, st>>>(c->keys, c->last, c->winner, c->cap, metric, idx, score,
                                                         count, status, camera, limit, B, k, now, cooldown, hits,
                                                         n_hits, overflow);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

}  // namespace
}  // namespace frp

using namespace frp;

extern "C" {

int frp_cooldown_create(int device, uint64_t capacity, frp_cooldown_t** out) {
    if (!out) { set_error("frp_cooldown_create: null argument"); return FRP_E_INVALID; }
    int ndev = 0;
    int rc = frp_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) { set_error("frp_cooldown_create: device %d out of range", device); return FRP_E_INVALID; }
    uint64_t cap = 1024;
    while (cap < capacity && cap < (1ull << 32)) cap <<= 1;
    DeviceGuard dg(device);
    frp_cooldown* c = new frp_cooldown();
    c->device = device;
    c->cap = cap;
    cudaError_t e = cudaMalloc((void**)&c->keys, cap * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->last, cap * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&c->winner, cap * sizeof(unsigned int));
    if (e != cudaSuccess) {
        set_error("frp_cooldown_create: allocation of %llu slots failed: %s", (unsigned long long)cap, cudaGetErrorString(e));
        frp_cooldown_destroy(c);
        cudaGetLastError();
        return FRP_E_NOMEM;
    }
    *out = c;
    rc = frp_cooldown_clear(c, nullptr);
    if (rc == FRP_OK && cudaStreamSynchronize(0) != cudaSuccess) rc = FRP_E_CUDA;
    if (rc) { frp_cooldown_destroy(c); *out = nullptr; }
    return rc;
}

int frp_cooldown_destroy(frp_cooldown_t* c) {
    if (!c) return FRP_OK;
    {
        std::unique_lock<std::shared_mutex> lk(c->mu);
        DeviceGuard dg(c->device);
        cudaDeviceSynchronize();
        if (c->keys) cudaFree(c->keys);
        if (c->last) cudaFree(c->last);
        if (c->winner) cudaFree(c->winner);
        cudaGetLastError();
    }
    delete c;
    return FRP_OK;
}

int frp_cooldown_clear(frp_cooldown_t* c, void* stream) {
    if (!c) { set_error("frp_cooldown_clear: null argument"); return FRP_E_INVALID; }
    std::unique_lock<std::shared_mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    cooldown_init_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(c->keys, c->last, c->winner, c->cap);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

int frp_cooldown_filter_device(frp_cooldown_t* c, int metric, const int64_t* idx_dev, const float* score_dev,
                               const int32_t* count_dev, const int32_t* camera_dev, const float* limit_dev, int B,
                               int k, double now, double cooldown, frp_hit_t* hits_dev, int32_t* n_hits_dev,
                               int32_t* overflow_dev, void* stream) {
    if (!c || !idx_dev || !score_dev || !count_dev || !camera_dev || !hits_dev || !n_hits_dev || B <= 0 || k <= 0 ||
        (long long)B * k > (long long)kFilterThreads * kMaxPer) {
        set_error("frp_cooldown_filter_device: invalid argument (B=%d k=%d, B*k <= %d)", B, k, kFilterThreads * kMaxPer);
        return FRP_E_INVALID;
    }
    std::unique_lock<std::shared_mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    return filter_launch(c, metric, idx_dev, score_dev, count_dev, nullptr, camera_dev, limit_dev, B, k, now, cooldown,
                         hits_dev, n_hits_dev, overflow_dev, (cudaStream_t)stream);
}

int frp_match_hits(frp_gallery_t* g, frp_cooldown_t* c, const float* queries, int B, double tau, int k, int flags,
                   const int32_t* camera, const float* limit, double now, double cooldown, frp_hit_t* hits,
                   int32_t* n_hits) {
    if (!g || !c || !queries || !camera || !hits || !n_hits || B <= 0 || k <= 0 || k > 64 ||
        (long long)B * k > (long long)kFilterThreads * kMaxPer) {
        set_error("frp_match_hits: invalid argument (B=%d k=%d, B*k <= %d)", B, k, kFilterThreads * kMaxPer);
        return FRP_E_INVALID;
    }
    if (c->device != g->device) { set_error("frp_match_hits: table on device %d, gallery on %d", c->device, g->device); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    std::unique_lock<std::shared_mutex> lkc(c->mu);
    DeviceGuard dg(g->device);
    cudaStream_t st = thread_stream(g->device);
    auto al = [](size_t n) { return (n + 255) & ~(size_t)255; };
    // staging (host pinned and device share the layout):
    //   queries | camera | limit | idx | score | count | status | n_hits, overflow | hits
    const size_t o_q = 0, o_cam = al((size_t)B * g->dim * 4), o_lim = o_cam + al((size_t)B * 4);
    const size_t o_idx = o_lim + al((size_t)B * 4), o_sc = o_idx + al((size_t)B * k * 8);
    const size_t o_cnt = o_sc + al((size_t)B * k * 4), o_st = o_cnt + al((size_t)B * 4);
    const size_t o_n = o_st + al((size_t)B * 4), o_hits = o_n + 256;
    const size_t total = o_hits + al((size_t)B * k * sizeof(frp_hit_t));
    // scratch bound as in frp_match (covers every prefilter configuration)
    const size_t cand = (size_t)g->sm_count * 64 + 64;
    const size_t scratch = al(al((size_t)B * g->dimp * 4) + al((size_t)(((B + 127) / 128) * 128) * g->dimp * 2) +
                              al((size_t)B * sizeof(QParams)) + al((size_t)B * cand * 8) + 1024);
    Workspace* w = workspace_for(g->device, st);
    int rc = workspace_reserve(w, scratch + total, total);
    if (rc) return rc;
    uint8_t* d = w->dev + scratch;
    uint8_t* h = w->pinned;
    memcpy(h + o_q, queries, (size_t)B * g->dim * 4);
    memcpy(h + o_cam, camera, (size_t)B * 4);
    if (limit) memcpy(h + o_lim, limit, (size_t)B * 4);
    // one H2D of queries | camera | limit
    cudaError_t e = cudaMemcpyAsync(d, h, o_idx, cudaMemcpyHostToDevice, st);
    int32_t* hn = reinterpret_cast<int32_t*>(h + o_n);
    for (int level = 0; level < 3 && e == cudaSuccess && rc == FRP_OK; ++level) {
        int f = flags;
        int KP = pick_kp(k);
        if (level >= 1) f = (flags & ~(FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_SCAN)) | FRP_MATCH_FORCE_EXACT;
        if (level == 2) { f |= FRP_MATCH_WIDE; KP = 64; }
        rc = match_device_locked(g, reinterpret_cast<const float*>(d + o_q), (size_t)g->dim, nullptr, B, tau, k, f, KP,
                                 reinterpret_cast<int64_t*>(d + o_idx), reinterpret_cast<float*>(d + o_sc), nullptr,
                                 reinterpret_cast<int32_t*>(d + o_cnt), reinterpret_cast<int32_t*>(d + o_st), st);
        if (rc) break;
        // the last level records whatever the widest exact pass produced (status ignored)
        rc = filter_launch(c, g->metric, reinterpret_cast<const int64_t*>(d + o_idx),
                           reinterpret_cast<const float*>(d + o_sc), reinterpret_cast<const int32_t*>(d + o_cnt),
                           level < 2 ? reinterpret_cast<const int32_t*>(d + o_st) : nullptr,
                           reinterpret_cast<const int32_t*>(d + o_cam),
                           limit ? reinterpret_cast<const float*>(d + o_lim) : nullptr, B, k, now, cooldown,
                           reinterpret_cast<frp_hit_t*>(d + o_hits), reinterpret_cast<int32_t*>(d + o_n),
                           reinterpret_cast<int32_t*>(d + o_n) + 1, st);
        if (rc) break;
        // one D2H: n_hits | overflow | hits
        e = cudaMemcpyAsync(h + o_n, d + o_n, total - o_n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess || hn[0] >= 0) break;
    }
    if (e != cudaSuccess) {
        set_error("frp_match_hits: CUDA error: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    if (rc) return rc;
    const int n = hn[0] < 0 ? 0 : hn[0];
    memcpy(hits, h + o_hits, (size_t)n * sizeof(frp_hit_t));
    *n_hits = n;
    return FRP_OK;
}

}  // extern "C"
