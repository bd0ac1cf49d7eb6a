// frp_match / frp_scores orchestration: query prep -> prefilter (streaming scan or
// tcgen05 GEMM) -> exact re-rank + certificate -> (host API) fp32-master fallback
// for uncertified queries.  Replaces face_service.compare_faces /
// batch_compare_faces / find_k_nearest arithmetic (face_service.py:395-481,590-612).
#include <string.h>

#include <mutex>
#include <vector>

#include "frp_internal.h"

namespace frp {

namespace {

std::once_flag g_pool_once[16];

void configure_pool(int device) {
    if (device < 0 || device >= 16) return;
    std::call_once(g_pool_once[device], [device]() {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thr = ~0ull;   // keep freed workspace cached in the pool
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    });
}

int pick_kp(int k) { return k <= 10 ? 16 : (k <= 24 ? 32 : 64); }

// optional per-kernel timing of the last match of the calling thread (bench roofline)
std::atomic<int> g_profile{0};
thread_local bool t_profile_valid = false;
cudaEvent_t* profile_events() {
    static thread_local cudaEvent_t ev[4] = {};
    if (!ev[0]) {
        for (int i = 0; i < 4; ++i)
            if (cudaEventCreate(&ev[i]) != cudaSuccess) { cudaGetLastError(); ev[0] = nullptr; return nullptr; }
    }
    return ev;
}

__global__ void fill_empty_kernel(int64_t* idx, float* score, int32_t* count, int32_t* status, int B, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * k) { idx[i] = -1; score[i] = __int_as_float(0x7fc00000); }
    if (i < B) { count[i] = 0; if (status) status[i] = FRP_STATUS_OK; }
}

// caller holds the shared lock and has set the device
int match_device_locked(frp_gallery* g, const float* q_dev, int B, double tau, int k, int flags, int KP,
                        int64_t* out_idx, float* out_score, int32_t* out_count, int32_t* out_status,
                        cudaStream_t st) {
    configure_pool(g->device);
    if (g->rows == 0) {
        const int n = B * k;
        fill_empty_kernel<<<(n + 255) / 256, 256, 0, st>>>(out_idx, out_score, out_count, out_status, B, k);
        FRP_CHECK_LAUNCH();
        set_match_path("empty");
        return FRP_OK;
    }
    const bool exact = (flags & FRP_MATCH_FORCE_EXACT) != 0;
    bool use_gemm = false;
    if (!exact && KP <= 32 && gemm_supported(g)) {
        // B > 4: tensor cores.  B <= 4 is HBM-bound on either kernel; on large galleries the
        // tcgen05 kernel's 2-D TMA ring streams faster than the scan kernel's FMA loop keeps up
        // with (4 M x 512: 0.566 vs 0.60-0.63 ms; 4 M x 128, B = 4: 0.158 vs 0.243 ms), on small
        // ones the scan kernel has the shorter ramp (200 k x 128: 14 vs 30 us)
        const size_t gbytes = (size_t)g->rows * g->dimp * sizeof(__half);
        const bool big = gbytes >= ((size_t)256 << 20) && (g->dimp >= 256 || B >= 3);
        if (flags & FRP_MATCH_FORCE_GEMM) use_gemm = true;
        else if (!(flags & FRP_MATCH_FORCE_SCAN) && (B > 4 || big)) use_gemm = true;
    }
    if ((flags & FRP_MATCH_FORCE_GEMM) && !use_gemm) {
        set_error("frp_match: tcgen05 path not available for this gallery (dimp=%d)", g->dimp);
        return FRP_E_UNSUPPORTED;
    }
    const int nlists = use_gemm ? gemm_num_lists(g, B) : scan_num_lists(g, exact);
    const int ec = exact ? FRP_EPS_SCAN_F32 : (use_gemm ? FRP_EPS_GEMM_F16 : FRP_EPS_SCAN_F16);

    const size_t qprep_bytes = (size_t)B * g->dimp * sizeof(float);
    const size_t q16_bytes = (size_t)(((B + 127) / 128) * 128) * g->dimp * sizeof(__half);
    const size_t qp_bytes = (((size_t)B * sizeof(QParams)) + 255) & ~(size_t)255;
    const size_t cand_bytes = (size_t)B * nlists * KP * sizeof(uint64_t);
    uint8_t* ws = nullptr;
    const size_t total = qprep_bytes + q16_bytes + qp_bytes + cand_bytes + 1024;
    FRP_CUDA(cudaMallocAsync((void**)&ws, total, st));
    float* qprep = reinterpret_cast<float*>(ws);
    __half* q16 = reinterpret_cast<__half*>(ws + qprep_bytes);
    QParams* qp = reinterpret_cast<QParams*>(ws + qprep_bytes + q16_bytes);
    uint64_t* cand = reinterpret_cast<uint64_t*>(ws + qprep_bytes + q16_bytes + qp_bytes);

    int rc = FRP_OK;
    const bool prof = g_profile.load(std::memory_order_relaxed) != 0;
    cudaEvent_t* ev = prof ? profile_events() : nullptr;
    if (use_gemm) {
        // padded query rows (B..ceil128) must be zero for the MMA tile
        cudaError_t e = cudaMemsetAsync(q16, 0, q16_bytes, st);
        if (e != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e)); rc = FRP_E_CUDA; }
    }
    if (ev) cudaEventRecord(ev[0], st);
    if (!rc) rc = launch_prep_queries(g, q_dev, B, tau, flags, ec, qprep, use_gemm ? q16 : nullptr, qp, st);
    if (ev) cudaEventRecord(ev[1], st);
    if (!rc) {
        if (use_gemm) {
            rc = launch_match_gemm(g, q16, qp, B, KP, cand, nlists, st);
            set_match_path("gemm_tcgen05");
        } else {
            rc = launch_match_scan(g, exact, qprep, qp, B, KP, cand, nlists, st);
            set_match_path(exact ? "scan_f32" : "scan_f16");
        }
    }
    if (ev) cudaEventRecord(ev[2], st);
    if (!rc) rc = launch_rerank(g, qprep, qp, B, k, KP, flags, tau, cand, nlists, out_idx, out_score,
                                out_count, out_status, st);
    if (ev) { cudaEventRecord(ev[3], st); t_profile_valid = true; }
    cudaFreeAsync(ws, st);
    return rc;
}

int check_match_args(const frp_gallery* g, const void* q, int B, int k, const void* oi, const void* os,
                     const void* oc) {
    if (!g || !q || !oi || !os || !oc || B <= 0 || k <= 0) {
        set_error("frp_match: invalid argument (B=%d k=%d)", B, k);
        return FRP_E_INVALID;
    }
    if (k > 64) {
        set_error("frp_match: k=%d exceeds the supported maximum of 64", k);
        return FRP_E_UNSUPPORTED;
    }
    return FRP_OK;
}

cudaStream_t thread_stream(int device) {
    static thread_local cudaStream_t st[16] = {};
    if (device < 0 || device >= 16) return 0;
    if (!st[device]) {
        if (cudaStreamCreateWithFlags(&st[device], cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            st[device] = 0;
        }
    }
    return st[device];
}

}  // namespace
}  // namespace frp

using namespace frp;

extern "C" {

int frp_profile_enable(int on) {
    g_profile.store(on ? 1 : 0);
    return FRP_OK;
}

int frp_profile_last(float* prep_ms, float* prefilter_ms, float* rerank_ms) {
    if (!t_profile_valid) { set_error("frp_profile_last: no profiled match on this thread"); return FRP_E_INVALID; }
    cudaEvent_t* ev = profile_events();
    if (!ev) { set_error("frp_profile_last: events unavailable"); return FRP_E_CUDA; }
    FRP_CUDA(cudaEventSynchronize(ev[3]));
    float a = 0, b = 0, c = 0;
    FRP_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
    FRP_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
    FRP_CUDA(cudaEventElapsedTime(&c, ev[2], ev[3]));
    if (prep_ms) *prep_ms = a;
    if (prefilter_ms) *prefilter_ms = b;
    if (rerank_ms) *rerank_ms = c;
    return FRP_OK;
}

int frp_match_device(frp_gallery_t* g, const float* queries_dev, int B, double tau, int k, int flags,
                     int64_t* out_idx_dev, float* out_score_dev, int32_t* out_count_dev,
                     int32_t* out_status_dev, void* stream) {
    int rc = check_match_args(g, queries_dev, B, k, out_idx_dev, out_score_dev, out_count_dev);
    if (rc) return rc;
    // buffers of another GPU would fault inside the kernels: refuse them here
    for (const void* p : {(const void*)queries_dev, (const void*)out_idx_dev}) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
            pa.device != g->device) {
            set_error("frp_match_device: buffer lives on device %d, the gallery on device %d", pa.device, g->device);
            return FRP_E_INVALID;
        }
        cudaGetLastError();
    }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    return match_device_locked(g, queries_dev, B, tau, k, flags, (flags & FRP_MATCH_WIDE) ? 64 : pick_kp(k),
                               out_idx_dev, out_score_dev,
                               out_count_dev, out_status_dev, (cudaStream_t)stream);
}

int frp_match(frp_gallery_t* g, const float* queries, int B, double tau, int k, int flags,
              int64_t* out_idx, float* out_score, int32_t* out_count, int32_t* out_status) {
    int rc = check_match_args(g, queries, B, k, out_idx, out_score, out_count);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    configure_pool(g->device);
    cudaStream_t st = thread_stream(g->device);
    const size_t qb = (size_t)B * g->dim * sizeof(float);
    const size_t ib = (size_t)B * k * sizeof(int64_t), sb = (size_t)B * k * sizeof(float);
    const size_t cb = (size_t)B * sizeof(int32_t);
    uint8_t* d = nullptr;
    const size_t off_i = (qb + 255) & ~(size_t)255;
    const size_t off_s = off_i + ((ib + 255) & ~(size_t)255);
    const size_t off_c = off_s + ((sb + 255) & ~(size_t)255);
    const size_t off_t = off_c + ((cb + 255) & ~(size_t)255);
    const size_t total = off_t + ((cb + 255) & ~(size_t)255);
    FRP_CUDA(cudaMallocAsync((void**)&d, total, st));
    float* dq = reinterpret_cast<float*>(d);
    int64_t* di = reinterpret_cast<int64_t*>(d + off_i);
    float* ds = reinterpret_cast<float*>(d + off_s);
    int32_t* dc = reinterpret_cast<int32_t*>(d + off_c);
    int32_t* dt = reinterpret_cast<int32_t*>(d + off_t);
    std::vector<int32_t> status(B, 0);
    cudaError_t e = cudaMemcpyAsync(dq, queries, qb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = match_device_locked(g, dq, B, tau, k, flags, pick_kp(k), di, ds, dc, dt, st);
    }
    if (e == cudaSuccess && rc == FRP_OK) {
        e = cudaMemcpyAsync(out_idx, di, ib, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_score, ds, sb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_count, dc, cb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(status.data(), dt, cb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    // uncertified queries: re-run against the fp32 master (wider lists on the second try)
    if (e == cudaSuccess && rc == FRP_OK) {
        for (int attempt = 0; attempt < 2; ++attempt) {
            // attempt 0: fp32 master with the same list width (skipped when the call was
            // already exact); attempt 1: fp32 master with the widest lists (KP = 64)
            if (attempt == 0 && (flags & FRP_MATCH_FORCE_EXACT)) continue;
            if (attempt == 1 && pick_kp(k) == 64 && (flags & FRP_MATCH_FORCE_EXACT)) break;
            std::vector<int> redo;
            for (int b = 0; b < B; ++b)
                if (status[b] == FRP_STATUS_FALLBACK) redo.push_back(b);
            if (redo.empty()) break;
            const int KP2 = (attempt == 0) ? pick_kp(k) : 64;
            const int nb = (int)redo.size();
            std::vector<float> q2((size_t)nb * g->dim);
            for (int j = 0; j < nb; ++j)
                memcpy(&q2[(size_t)j * g->dim], queries + (size_t)redo[j] * g->dim, g->dim * sizeof(float));
            std::vector<int64_t> i2((size_t)nb * k);
            std::vector<float> s2((size_t)nb * k);
            std::vector<int32_t> c2(nb), t2(nb);
            e = cudaMemcpyAsync(dq, q2.data(), q2.size() * sizeof(float), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) break;
            rc = match_device_locked(g, dq, nb, tau, k, (flags & ~(FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_SCAN)) | FRP_MATCH_FORCE_EXACT,
                                     KP2, di, ds, dc, dt, st);
            if (rc) break;
            e = cudaMemcpyAsync(i2.data(), di, i2.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s2.data(), ds, s2.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c2.data(), dc, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(t2.data(), dt, nb * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) break;
            for (int j = 0; j < nb; ++j) {
                const int b = redo[j];
                memcpy(out_idx + (size_t)b * k, &i2[(size_t)j * k], k * sizeof(int64_t));
                memcpy(out_score + (size_t)b * k, &s2[(size_t)j * k], k * sizeof(float));
                out_count[b] = c2[j];
                // 1 = answered by the fp32 fallback (certified); stays FALLBACK for the next attempt otherwise
                status[b] = (t2[j] == FRP_STATUS_OK) ? -1 : FRP_STATUS_FALLBACK;
            }
        }
        for (int b = 0; b < B; ++b) {
            // -1: certified by the fallback -> report FRP_STATUS_FALLBACK (informational);
            // still FALLBACK after both attempts -> 2 (uncertified: > 64 exact ties at the cut)
            if (status[b] == -1) status[b] = FRP_STATUS_FALLBACK;
            else if (status[b] == FRP_STATUS_FALLBACK) status[b] = 2;
        }
    }
    cudaFreeAsync(d, st);
    cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("frp_match: CUDA error: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    if (rc == FRP_OK && out_status) memcpy(out_status, status.data(), cb);
    return rc;
}

int frp_scores_device(frp_gallery_t* g, const float* queries_dev, int B, double* out_dev, void* stream) {
    if (!g || !queries_dev || !out_dev || B <= 0) { set_error("frp_scores_device: invalid argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    configure_pool(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (g->rows == 0) return FRP_OK;
    uint8_t* ws = nullptr;
    const size_t qprep_bytes = (size_t)B * g->dimp * sizeof(float);
    FRP_CUDA(cudaMallocAsync((void**)&ws, qprep_bytes + (size_t)B * sizeof(QParams) + 256, st));
    float* qprep = reinterpret_cast<float*>(ws);
    QParams* qp = reinterpret_cast<QParams*>(ws + ((qprep_bytes + 255) & ~(size_t)255));
    int rc = launch_prep_queries(g, queries_dev, B, 0.0, FRP_MATCH_NO_THRESHOLD, FRP_EPS_SCAN_F32, qprep, nullptr,
                                 qp, st);
    if (!rc) rc = launch_scores(g, qprep, B, out_dev, st);
    cudaFreeAsync(ws, st);
    return rc;
}

int frp_scores(frp_gallery_t* g, const float* queries, int B, double* out) {
    if (!g || !queries || !out || B <= 0) { set_error("frp_scores: invalid argument"); return FRP_E_INVALID; }
    uint64_t rows = 0;
    frp_gallery_count(g, &rows, nullptr);
    if (rows == 0) return FRP_OK;
    DeviceGuard dg(g->device);
    configure_pool(g->device);
    cudaStream_t st = thread_stream(g->device);
    float* dq = nullptr; double* dout = nullptr;
    FRP_CUDA(cudaMallocAsync((void**)&dq, (size_t)B * g->dim * sizeof(float), st));
    cudaError_t e = cudaMallocAsync((void**)&dout, (size_t)B * rows * sizeof(double), st);
    int rc = FRP_OK;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dq, queries, (size_t)B * g->dim * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = frp_scores_device(g, dq, B, dout, st);
    if (e == cudaSuccess && rc == FRP_OK) e = cudaMemcpyAsync(out, dout, (size_t)B * rows * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (dq) cudaFreeAsync(dq, st);
    if (dout) cudaFreeAsync(dout, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) { set_error("frp_scores: CUDA error: %s", cudaGetErrorString(e)); cudaGetLastError(); return FRP_E_CUDA; }
    return rc;
}

}  // extern "C"
