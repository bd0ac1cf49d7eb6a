// Gallery against itself on the device: cluster_faces (backend/app/services/face_service.py:552-585).
//
// The reference walks the faces in dict order; an unassigned face seeds a cluster and absorbs
// every unassigned face j with dist(i, j) <= threshold — an O(N^2) Python loop of single
// face_distance calls.  Here the gallery is matched against itself in slabs of up to 256
// not-yet-assigned rows: the slab's master rows are the queries (gathered on the device, no
// host round trip of vectors), they run through the SAME prefilter (tcgen05 kernel for
// slabs of more than 4 rows) + exact fp64 re-rank as frp_match with threshold tau and k = 64,
// and only the thresholded neighbour lists (<= 64 rows per seed candidate) come back.  The
// greedy pass itself is sequential by definition and runs on the host over those sparse
// lists.  A seed whose list is full (or uncertified) gets its complete neighbourhood from a
// thresholded dense pass that returns row indices, never the N scores.
#include <string.h>

#include <vector>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {
namespace {

constexpr int kSlab = 256;    // seed candidates per device pass
constexpr int kNbr = 64;      // neighbours returned per candidate (frp_match's k limit)

// one warp per gallery row: exact fp64 score of `qrow` vs every live row from the fp32
// master, rows passing the threshold appended (unordered) to out[]
__global__ void dense_neighbours_kernel(const float* __restrict__ g32, const uint8_t* __restrict__ alive,
                                        uint64_t rows, int dimp, int metric, uint64_t qrow, double tau,
                                        int strict, uint32_t* __restrict__ out, unsigned int* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows || !alive[row]) return;
    const float4* gr = reinterpret_cast<const float4*>(g32 + row * dimp);
    const float4* qv = reinterpret_cast<const float4*>(g32 + qrow * dimp);
    double acc = 0.0;
    for (int i = lane; i < (dimp >> 2); i += 32) {
        const float4 g = __ldg(gr + i);
        const float4 q = __ldg(qv + i);
        if (metric == FRP_METRIC_COS) {
            acc += (double)g.x * (double)q.x + (double)g.y * (double)q.y + (double)g.z * (double)q.z +
                   (double)g.w * (double)q.w;
        } else {
            const double dx = (double)g.x - (double)q.x, dy = (double)g.y - (double)q.y;
            const double dz = (double)g.z - (double)q.z, dw = (double)g.w - (double)q.w;
            acc += dx * dx + dy * dy + dz * dz + dw * dw;
        }
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) {
        bool p;
        if (metric == FRP_METRIC_COS) p = strict ? (acc > tau) : (acc >= tau);
        else { const double d = sqrt(acc); p = strict ? (d < tau) : (d <= tau); }
        if (p) out[atomicAdd(count, 1u)] = (uint32_t)row;
    }
}

}  // namespace
}  // namespace frp

using namespace frp;

extern "C" int frp_cluster(frp_gallery_t* g, double tau, int flags, int32_t* labels, uint64_t labels_len,
                           int32_t* n_clusters, uint64_t* stats4) {
    if (!g || !labels || !n_clusters) { set_error("frp_cluster: null argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    const uint64_t rows = g->rows;
    if (labels_len < rows) {
        set_error("frp_cluster: gallery has %llu rows, labels holds %llu", (unsigned long long)rows,
                  (unsigned long long)labels_len);
        return FRP_E_RANGE;
    }
    if (rows >= (1ull << 32)) { set_error("frp_cluster: more than 2^32 rows per shard"); return FRP_E_UNSUPPORTED; }
    if (g->seq) { set_error("frp_cluster: galleries with explicit row sequences (mutable shards) are not supported"); return FRP_E_UNSUPPORTED; }
    DeviceGuard dg(g->device);
    cudaStream_t st = thread_stream(g->device);
    const uint64_t launches0 = g_launches.load();
    const int mflags = (flags & FRP_MATCH_STRICT) | (flags & (FRP_MATCH_FORCE_SCAN | FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_EXACT));
    const bool strict = (flags & FRP_MATCH_STRICT) != 0;

    std::vector<uint8_t> assigned((size_t)rows);
    for (uint64_t r = 0; r < rows; ++r) { assigned[(size_t)r] = g->h_alive[(size_t)r] ? 0 : 1; labels[r] = -1; }

    // device / pinned staging of one slab: rows u64[256] | idx i64[256][64] | count i32[256] | status i32[256]
    const size_t off_rows = 0, off_idx = kSlab * sizeof(uint64_t);
    const size_t off_cnt = off_idx + (size_t)kSlab * kNbr * sizeof(int64_t);
    const size_t off_st = off_cnt + kSlab * sizeof(int32_t);
    const size_t slab_bytes = off_st + kSlab * sizeof(int32_t);
    // slab staging lives in the calling thread's workspace (device block behind the match scratch,
    // pinned block): no allocation per call
    uint8_t* d_slab = nullptr; uint8_t* h_slab = nullptr;
    uint32_t* d_dense = nullptr; unsigned int* d_dense_cnt = nullptr;
    std::vector<uint32_t> h_dense;
    int rc = FRP_OK;
    {
        auto al = [](size_t n) { return (n + 255) & ~(size_t)255; };
        const size_t cand = (size_t)g->sm_count * 64 + 64;
        const size_t scratch = al(al((size_t)kSlab * g->dimp * 4) + al((size_t)(((kSlab + 127) / 128) * 128) * g->dimp * 2) +
                                  al((size_t)kSlab * sizeof(QParams)) + al((size_t)kSlab * cand * 8) + 1024);
        Workspace* w = workspace_for(g->device, st);
        rc = workspace_reserve(w, scratch + al(slab_bytes), al(slab_bytes));
        if (rc) return rc;
        d_slab = w->dev + scratch;
        h_slab = w->pinned;
    }
    cudaError_t e = cudaSuccess;
    auto cleanup = [&]() {
        if (d_dense) cudaFree(d_dense);
        if (d_dense_cnt) cudaFree(d_dense_cnt);
    };
    uint64_t* h_rows = reinterpret_cast<uint64_t*>(h_slab + off_rows);
    const int64_t* h_idx = reinterpret_cast<const int64_t*>(h_slab + off_idx);
    const int32_t* h_cnt = reinterpret_cast<const int32_t*>(h_slab + off_cnt);
    const int32_t* h_st = reinterpret_cast<const int32_t*>(h_slab + off_st);

    int32_t cid = 0;
    uint64_t n_slabs = 0, n_matched = 0, n_dense = 0;
    uint64_t cursor = 0;
    while (cursor < rows && rc == FRP_OK) {
        int nb = 0;
        while (cursor < rows && nb < kSlab) {
            if (!assigned[(size_t)cursor]) h_rows[nb++] = cursor;
            ++cursor;
        }
        if (nb == 0) break;
        ++n_slabs;
        n_matched += (uint64_t)nb;
        e = cudaMemcpyAsync(d_slab + off_rows, h_rows, (size_t)nb * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        rc = match_device_locked(g, g->g32, (size_t)g->dimp, reinterpret_cast<const uint64_t*>(d_slab + off_rows), nb,
                                 tau, kNbr, mflags, pick_kp(kNbr), reinterpret_cast<int64_t*>(d_slab + off_idx), nullptr,
                                 nullptr, reinterpret_cast<int32_t*>(d_slab + off_cnt),
                                 reinterpret_cast<int32_t*>(d_slab + off_st), st);
        if (rc) break;
        e = cudaMemcpyAsync(h_slab + off_idx, d_slab + off_idx, slab_bytes - off_idx, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) break;
        // greedy pass over this slab, in row order (face_service.py:564-583)
        for (int q = 0; q < nb && rc == FRP_OK; ++q) {
            const uint64_t i = h_rows[q];
            if (assigned[(size_t)i]) continue;          // absorbed by an earlier seed of this slab
            assigned[(size_t)i] = 1;
            labels[i] = cid;
            if (h_st[q] == FRP_STATUS_OK && h_cnt[q] < kNbr) {
                for (int j = 0; j < h_cnt[q]; ++j) {
                    const uint64_t r = (uint64_t)(h_idx[(size_t)q * kNbr + j] - (int64_t)g->base_row);
                    if (r < rows && !assigned[(size_t)r]) { assigned[(size_t)r] = 1; labels[r] = cid; }
                }
            } else {
                // a full (or uncertified) list: the complete thresholded neighbourhood of this seed
                ++n_dense;
                if (!d_dense) {
                    e = cudaMalloc((void**)&d_dense, (size_t)rows * sizeof(uint32_t));
                    if (e == cudaSuccess) e = cudaMalloc((void**)&d_dense_cnt, sizeof(unsigned int));
                    if (e != cudaSuccess) break;
                    h_dense.resize((size_t)rows);
                }
                e = cudaMemsetAsync(d_dense_cnt, 0, sizeof(unsigned int), st);
                if (e != cudaSuccess) break;
                const int wpb = 8;
                dense_neighbours_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
                    g->g32, g->alive, rows, g->dimp, g->metric, i, tau, strict ? 1 : 0, d_dense, d_dense_cnt);
                e = cudaGetLastError();
                count_launch();
                unsigned int nn = 0;
                if (e == cudaSuccess) e = cudaMemcpyAsync(&nn, d_dense_cnt, sizeof(nn), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e == cudaSuccess && nn)
                    e = cudaMemcpyAsync(h_dense.data(), d_dense, (size_t)nn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) break;
                for (unsigned int j = 0; j < nn; ++j) {
                    const uint32_t r = h_dense[j];
                    if (!assigned[r]) { assigned[r] = 1; labels[r] = cid; }
                }
            }
            ++cid;
        }
        if (e != cudaSuccess) break;
    }
    cleanup();
    if (e != cudaSuccess) {
        set_error("frp_cluster: CUDA error: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    if (rc) return rc;
    *n_clusters = cid;
    if (stats4) {
        stats4[0] = n_slabs; stats4[1] = n_matched; stats4[2] = n_dense;
        stats4[3] = g_launches.load() - launches0;
    }
    return FRP_OK;
}
