// frp_match / frp_scores orchestration: query prep -> prefilter (streaming scan or
// tcgen05 GEMM) -> exact re-rank + certificate -> (host API) fp32-master fallback
// for uncertified queries.  Replaces face_service.compare_faces /
// batch_compare_faces / find_k_nearest arithmetic (face_service.py:395-481,590-612).
//
// Hot-call hygiene: the steady state of every entry point allocates nothing and encodes no
// gallery tensor map — scratch lives in a per-(thread, stream) Workspace that only grows,
// the gallery's TMA descriptors are cached in the gallery (gallery.cu / match_gemm.cu), the
// query descriptor in the workspace, and the host API moves its results with ONE D2H copy
// through pinned staging.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "frp_internal.h"

namespace frp {

int pick_kp(int k) { return k <= 10 ? 16 : (k <= 24 ? 32 : 64); }
// the tcgen05 epilogue keeps at most 32 candidates per thread in registers; lists narrower
// than k are sound because the certificate compares the k-th exact score with the weakest
// key of every FULL list (a query with more than 32 of its top-k in one CTA's rows falls back)
static int pick_kp_gemm(int k) { return k <= 10 ? 16 : 32; }

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

bool pdl_enabled() {
    static const bool on = []() { const char* e = getenv("FRP_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

Workspace* workspace_for(int device, cudaStream_t st) {
    static thread_local std::map<std::pair<int, cudaStream_t>, Workspace> t_ws;
    Workspace& w = t_ws[std::make_pair(device, st)];
    w.device = device;
    return &w;
}

int workspace_reserve(Workspace* w, size_t dev_bytes, size_t pinned_bytes) {
    if (dev_bytes > w->dev_bytes) {
        // cudaFree synchronises the device, so work still reading the old block has finished
        if (w->dev) cudaFree(w->dev);
        w->dev = nullptr; w->dev_bytes = 0; w->q16_ptr = nullptr;
        size_t want = dev_bytes + dev_bytes / 2;
        want = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        cudaError_t e = cudaMalloc((void**)&w->dev, want);
        if (e != cudaSuccess) {
            set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return FRP_E_NOMEM;
        }
        w->dev_bytes = want;
    }
    if (pinned_bytes > w->pinned_bytes) {
        if (w->pinned) cudaFreeHost(w->pinned);
        w->pinned = nullptr; w->pinned_bytes = 0;
        size_t want = pinned_bytes + pinned_bytes / 2;
        want = (want + 65535) & ~(size_t)65535;
        cudaError_t e = cudaMallocHost((void**)&w->pinned, want);
        if (e != cudaSuccess) {
            set_error("pinned staging allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return FRP_E_NOMEM;
        }
        w->pinned_bytes = want;
    }
    return FRP_OK;
}

namespace {

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

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

__global__ void fill_empty_kernel(int64_t* idx, float* score, double* score64, int32_t* count, int32_t* status,
                                  int B, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * k) {
        idx[i] = -1;
        if (score) score[i] = __int_as_float(0x7fc00000);
        if (score64) score64[i] = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (i < B) { count[i] = 0; if (status) status[i] = FRP_STATUS_OK; }
}

// bytes of match scratch for (B, widest nlists * KP)
size_t match_scratch_bytes(const frp_gallery* g, int B, size_t cand_slots) {
    return align256((size_t)B * g->dimp * sizeof(float)) +
           align256((size_t)(((B + 127) / 128) * 128) * g->dimp * sizeof(__half)) +
           align256((size_t)B * sizeof(QParams)) + align256((size_t)B * cand_slots * sizeof(uint64_t)) + 1024;
}

}  // namespace

// caller holds the shared lock and has set the device
int match_device_locked(frp_gallery* g, const float* q_dev, size_t qstride, const uint64_t* q_rows, int B,
                        double tau, int k, int flags, int KP, int64_t* out_idx, float* out_score,
                        double* out_score64, int32_t* out_count, int32_t* out_status, cudaStream_t st) {
    if (g->rows == 0) {
        const int n = B * k;
        fill_empty_kernel<<<(n + 255) / 256, 256, 0, st>>>(out_idx, out_score, out_score64, out_count, out_status,
                                                           B, k);
        FRP_CHECK_LAUNCH();
        set_match_path("empty");
        return FRP_OK;
    }
    // the reference's own scale (a few thousand faces): one fused exact launch, no prefilter
    if (!q_rows && !(flags & (FRP_MATCH_FORCE_SCAN | FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_EXACT | FRP_MATCH_WIDE)) &&
        match_small_supported(g, B)) {
        const bool prof = g_profile.load(std::memory_order_relaxed) != 0;
        cudaEvent_t* ev = prof ? profile_events() : nullptr;
        if (ev) { cudaEventRecord(ev[0], st); cudaEventRecord(ev[1], st); }
        int rc = launch_match_small(g, q_dev, qstride, B, tau, k, flags, out_idx, out_score, out_score64, out_count,
                                    out_status, st);
        if (ev) { cudaEventRecord(ev[2], st); cudaEventRecord(ev[3], st); t_profile_valid = true; }
        set_match_path("small_exact");
        return rc;
    }
    const bool exact = (flags & FRP_MATCH_FORCE_EXACT) != 0;
    const bool wide = (flags & FRP_MATCH_WIDE) != 0;      // the ladder's last rung: 64-key lists on the scan kernel
    bool use_gemm = false;
    if (!exact && !wide && gemm_supported(g)) {
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
    if (use_gemm) KP = pick_kp_gemm(k);
    const int nlists = use_gemm ? gemm_num_lists(g, B) : scan_num_lists(g, exact);
    const int ec = exact ? FRP_EPS_SCAN_F32 : (use_gemm ? FRP_EPS_GEMM_F16 : FRP_EPS_SCAN_F16);

    Workspace* w = workspace_for(g->device, st);
    int rc = workspace_reserve(w, match_scratch_bytes(g, B, (size_t)nlists * KP), 0);
    if (rc) return rc;
    const int groups = (B + 127) / 128;
    const size_t qprep_bytes = align256((size_t)B * g->dimp * sizeof(float));
    const size_t q16_bytes = align256((size_t)groups * 128 * g->dimp * sizeof(__half));
    const size_t qp_bytes = align256((size_t)B * sizeof(QParams));
    float* qprep = reinterpret_cast<float*>(w->dev);
    __half* q16 = reinterpret_cast<__half*>(w->dev + qprep_bytes);
    QParams* qp = reinterpret_cast<QParams*>(w->dev + qprep_bytes + q16_bytes);
    uint64_t* cand = reinterpret_cast<uint64_t*>(w->dev + qprep_bytes + q16_bytes + qp_bytes);

    const bool prof = g_profile.load(std::memory_order_relaxed) != 0;
    cudaEvent_t* ev = prof ? profile_events() : nullptr;
    if (use_gemm) {
        // padded query rows (B .. groups*128) must be zero for the MMA tile
        const int pad_rows = groups * 128 - B;
        if (pad_rows) {
            cudaError_t e = cudaMemsetAsync(q16 + (size_t)B * g->dimp, 0, (size_t)pad_rows * g->dimp * sizeof(__half), st);
            if (e != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e)); return FRP_E_CUDA; }
        }
        if (w->q16_ptr != q16 || w->q_groups != groups || w->q_dimp != g->dimp) {
            rc = encode_query_map(w->tmap_q, q16, groups, g->dimp);
            if (rc) return rc;
            w->q16_ptr = q16; w->q_groups = groups; w->q_dimp = g->dimp;
        }
    }
    if (ev) cudaEventRecord(ev[0], st);
    rc = launch_prep_queries(g, q_dev, qstride, q_rows, B, tau, flags, ec, qprep, use_gemm ? q16 : nullptr, qp, st);
    if (ev) cudaEventRecord(ev[1], st);
    if (!rc) {
        if (use_gemm) {
            // pure top-k over a large gallery: a sampling pass over the first ~1.5 % of the rows
            // lifts every query's cut from -inf to a sound bound (see raise_cut_kernel), so the
            // main pass inserts a handful of candidates per list instead of ~KP ln(rows / KP)
            const uint64_t sample = (uint64_t)nlists * 4 * 256;
            if ((flags & FRP_MATCH_NO_THRESHOLD) && !(flags & FRP_MATCH_NO_SAMPLE) && g->rows >= 32 * sample) {
                rc = launch_match_gemm(g, w->tmap_q, qp, B, KP, cand, nlists, st, sample);
                if (!rc) rc = launch_raise_cut(qp, B, k, cand, nlists, KP, st);
            }
            if (!rc) rc = launch_match_gemm(g, w->tmap_q, qp, B, KP, cand, nlists, st);
            set_match_path("gemm_tcgen05");
        } else {
            rc = launch_match_scan(g, exact, qprep, qp, B, KP, cand, nlists, st);
            set_match_path(exact ? "scan_f32" : "scan_f16");
        }
    }
    if (ev) cudaEventRecord(ev[2], st);
    if (!rc) rc = launch_rerank(g, qprep, qp, B, k, KP, flags, tau, cand, nlists, out_idx, out_score, out_score64,
                                out_count, out_status, st);
    if (ev) { cudaEventRecord(ev[3], st); t_profile_valid = true; }
    return rc;
}

namespace {

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

// buffers of another GPU would fault inside the kernels: refuse them here
int check_same_device(const frp_gallery* g, const void* a, const void* b, const char* who) {
    for (const void* p : {a, b}) {
        cudaPointerAttributes pa;
        if (p && cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
            pa.device != g->device) {
            set_error("%s: buffer lives on device %d, the gallery on device %d", who, pa.device, g->device);
            return FRP_E_INVALID;
        }
        cudaGetLastError();
    }
    return FRP_OK;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes pa;
    const bool ok = cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return ok;
}

struct HostLayout {
    size_t off_q, off_i, off_s, off_c, off_t, total;   // queries | idx | score | count | status
    HostLayout(int B, int dim, int k) {
        off_q = 0;
        off_i = align256((size_t)B * dim * sizeof(float));
        off_s = off_i + (size_t)B * k * sizeof(int64_t);
        off_c = off_s + (size_t)B * k * sizeof(float);
        off_t = off_c + (size_t)B * sizeof(int32_t);
        total = align256(off_t + (size_t)B * sizeof(int32_t));
    }
    size_t result_bytes() const { return total - off_i; }
};

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
    rc = check_same_device(g, queries_dev, out_idx_dev, "frp_match_device");
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    return match_device_locked(g, queries_dev, (size_t)g->dim, nullptr, B, tau, k, flags,
                               (flags & FRP_MATCH_WIDE) ? 64 : pick_kp(k), out_idx_dev, out_score_dev, nullptr,
                               out_count_dev, out_status_dev, (cudaStream_t)stream);
}

int frp_match_record_device(frp_gallery_t* g, const float* queries_dev, int B, double tau, int k, int flags,
                            void* record_dev, void* stream) {
    int rc = check_match_args(g, queries_dev, B, k, record_dev, record_dev, record_dev);
    if (rc) return rc;
    rc = check_same_device(g, queries_dev, record_dev, "frp_match_record_device");
    if (rc) return rc;
    uint8_t* rec = reinterpret_cast<uint8_t*>(record_dev);
    const size_t o1 = (size_t)B * k * 8, o2 = o1 + (size_t)B * k * 8, o3 = o2 + (size_t)B * 4;
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    return match_device_locked(g, queries_dev, (size_t)g->dim, nullptr, B, tau, k, flags,
                               (flags & FRP_MATCH_WIDE) ? 64 : pick_kp(k), reinterpret_cast<int64_t*>(rec), nullptr,
                               reinterpret_cast<double*>(rec + o1), reinterpret_cast<int32_t*>(rec + o2),
                               reinterpret_cast<int32_t*>(rec + o3), (cudaStream_t)stream);
}

int frp_match(frp_gallery_t* g, const float* queries, int B, double tau, int k, int flags,
              int64_t* out_idx, float* out_score, int32_t* out_count, int32_t* out_status) {
    int rc = check_match_args(g, queries, B, k, out_idx, out_score, out_count);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    cudaStream_t st = thread_stream(g->device);
    const HostLayout L(B, g->dim, k);
    Workspace* w = workspace_for(g->device, st);
    // device block: [match scratch (sized for the widest configuration) | queries | results];
    // pinned block: queries | results.  The scratch bound covers every prefilter (tcgen05: at most
    // 2 lists per SM x 32 keys; scan: 1 list per SM x 64 keys), so the nested reserve never regrows.
    const size_t scratch = align256(match_scratch_bytes(g, B, (size_t)g->sm_count * 64 + 64));
    rc = workspace_reserve(w, scratch + L.total, L.total);
    if (rc) return rc;
    uint8_t* d = w->dev + scratch;
    float* dq = reinterpret_cast<float*>(d + L.off_q);
    int64_t* di = reinterpret_cast<int64_t*>(d + L.off_i);
    float* ds = reinterpret_cast<float*>(d + L.off_s);
    int32_t* dc = reinterpret_cast<int32_t*>(d + L.off_c);
    int32_t* dt = reinterpret_cast<int32_t*>(d + L.off_t);
    uint8_t* h = w->pinned;
    const size_t qb = (size_t)B * g->dim * sizeof(float);
    const size_t ib = (size_t)B * k * sizeof(int64_t), sb = (size_t)B * k * sizeof(float);
    const size_t cb = (size_t)B * sizeof(int32_t);

    // The reference's own scale (<= 4 096 faces): the fused exact kernel reads the queries from and
    // writes the results to the pinned staging block directly (mapped host memory under UVA): one
    // launch and one stream sync, no copy operations at all.
    if (!(flags & (FRP_MATCH_FORCE_SCAN | FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_EXACT | FRP_MATCH_WIDE)) &&
        match_small_supported(g, B)) {
        memcpy(h + L.off_q, queries, qb);
        // completion: the kernel publishes a sequence number in mapped host memory and the host
        // polls that word (a stream synchronisation costs several microseconds more than the
        // kernel itself at this size); the stream is queried now and then so that a failed kernel
        // still ends the wait.  FRP_SMALL_POLL=0 falls back to cudaStreamSynchronize.
        static const bool poll_enabled = []() { const char* e = getenv("FRP_SMALL_POLL"); return !(e && e[0] == '0'); }();
        if (poll_enabled && !w->done_flag) {
            unsigned int* hf = nullptr;
            unsigned int* dc0 = nullptr;
            if (cudaMallocHost((void**)&hf, 64) == cudaSuccess && cudaMalloc((void**)&dc0, 4) == cudaSuccess &&
                cudaMemset(dc0, 0, 4) == cudaSuccess) {
                *hf = 0;
                w->done_flag = hf; w->done_count = dc0;
            } else {
                cudaGetLastError();
                if (hf) cudaFreeHost(hf);
                if (dc0) cudaFree(dc0);
            }
        }
        const bool poll = poll_enabled && w->done_flag;
        const unsigned int seq = poll ? ++w->done_seq : 0;
        rc = launch_match_small(g, reinterpret_cast<const float*>(h + L.off_q), (size_t)g->dim, B, tau, k, flags,
                                reinterpret_cast<int64_t*>(h + L.off_i), reinterpret_cast<float*>(h + L.off_s), nullptr,
                                reinterpret_cast<int32_t*>(h + L.off_c), reinterpret_cast<int32_t*>(h + L.off_t), st,
                                poll ? w->done_count : nullptr, poll ? w->done_flag : nullptr, seq, queries);
        set_match_path("small_exact");
        cudaError_t es = cudaSuccess;
        if (poll && rc == FRP_OK) {
            unsigned int spins = 0;
            while (*w->done_flag != seq) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                if ((++spins & 0xfff) == 0) {
                    const cudaError_t qe = cudaStreamQuery(st);
                    if (qe == cudaErrorNotReady) continue;
                    es = qe;                      // finished (the flag is set by now) or failed
                    if (qe != cudaSuccess || *w->done_flag == seq) break;
                    es = cudaStreamSynchronize(st);
                    break;
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        } else {
            es = cudaStreamSynchronize(st);
        }
        if (es != cudaSuccess) {
            set_error("frp_match: CUDA error: %s", cudaGetErrorString(es));
            cudaGetLastError();
            return FRP_E_CUDA;
        }
        if (rc) return rc;
        memcpy(out_idx, h + L.off_i, ib);
        memcpy(out_score, h + L.off_s, sb);
        memcpy(out_count, h + L.off_c, cb);
        if (out_status) memcpy(out_status, h + L.off_t, cb);
        return FRP_OK;
    }
    // small batches are staged (a 2 KB memcpy is cheaper than asking the driver whether the
    // caller's buffer is pinned); large pinned batches are copied from where they are
    const void* qsrc = queries;
    if (qb <= 65536 || !is_pinned_host(queries)) { memcpy(h + L.off_q, queries, qb); qsrc = h + L.off_q; }
    const uint8_t* block0 = w->dev;
    cudaError_t e = cudaMemcpyAsync(dq, qsrc, qb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        rc = match_device_locked(g, dq, (size_t)g->dim, nullptr, B, tau, k, flags, pick_kp(k), di, ds, nullptr, dc, dt, st);
    if (rc == FRP_OK && w->dev != block0) {      // cannot happen while the scratch bound above holds
        set_error("frp_match: workspace moved under the staging area");
        return FRP_E_NOMEM;
    }
    // ONE device->host copy of idx | score | count | status
    if (e == cudaSuccess && rc == FRP_OK)
        e = cudaMemcpyAsync(h + L.off_i, d + L.off_i, L.result_bytes(), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == FRP_OK) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("frp_match: CUDA error: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    if (rc) return rc;
    memcpy(out_idx, h + L.off_i, ib);
    memcpy(out_score, h + L.off_s, sb);
    memcpy(out_count, h + L.off_c, cb);
    int32_t* status = reinterpret_cast<int32_t*>(h + L.off_t);
    bool any = false;
    for (int b = 0; b < B; ++b) any = any || status[b] != FRP_STATUS_OK;
    if (any) {
        // uncertified queries: re-run against the fp32 master (wider lists on the second try)
        std::vector<int32_t> stt(status, status + B);
        for (int attempt = 0; attempt < 2; ++attempt) {
            // attempt 0: fp32 master with the same list width (skipped when the call was
            // already exact); attempt 1: fp32 master with the widest lists (KP = 64)
            if (attempt == 0 && (flags & FRP_MATCH_FORCE_EXACT)) continue;
            if (attempt == 1 && pick_kp(k) == 64 && (flags & FRP_MATCH_FORCE_EXACT)) break;
            std::vector<int> redo;
            for (int b = 0; b < B; ++b)
                if (stt[b] == FRP_STATUS_FALLBACK) redo.push_back(b);
            if (redo.empty()) break;
            const int KP2 = (attempt == 0) ? pick_kp(k) : 64;
            const int nb = (int)redo.size();
            float* hq2 = reinterpret_cast<float*>(h + L.off_q);
            for (int j = 0; j < nb; ++j)
                memcpy(hq2 + (size_t)j * g->dim, queries + (size_t)redo[j] * g->dim, g->dim * sizeof(float));
            e = cudaMemcpyAsync(dq, hq2, (size_t)nb * g->dim * sizeof(float), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) break;
            const HostLayout L2(nb, g->dim, k);
            // the redo batch reuses the head of the same result block with its own (smaller) layout
            int64_t* di2 = reinterpret_cast<int64_t*>(d + L.off_i);
            float* ds2 = reinterpret_cast<float*>(d + L.off_i + (size_t)nb * k * 8);
            int32_t* dc2 = reinterpret_cast<int32_t*>(d + L.off_i + (size_t)nb * k * 12);
            int32_t* dt2 = dc2 + nb;
            rc = match_device_locked(g, dq, (size_t)g->dim, nullptr, nb, tau, k,
                                     ((flags & ~(FRP_MATCH_FORCE_GEMM | FRP_MATCH_FORCE_SCAN)) | FRP_MATCH_FORCE_EXACT) |
                                         (attempt == 1 ? FRP_MATCH_WIDE : 0),
                                     KP2, di2, ds2, nullptr, dc2, dt2, st);
            if (rc) break;
            const size_t rb = (size_t)nb * k * 12 + (size_t)nb * 8;
            e = cudaMemcpyAsync(h + L.off_i, d + L.off_i, rb, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) break;
            const int64_t* i2 = reinterpret_cast<const int64_t*>(h + L.off_i);
            const float* s2 = reinterpret_cast<const float*>(h + L.off_i + (size_t)nb * k * 8);
            const int32_t* c2 = reinterpret_cast<const int32_t*>(h + L.off_i + (size_t)nb * k * 12);
            const int32_t* t2 = c2 + nb;
            for (int j = 0; j < nb; ++j) {
                const int b = redo[j];
                memcpy(out_idx + (size_t)b * k, i2 + (size_t)j * k, k * sizeof(int64_t));
                memcpy(out_score + (size_t)b * k, s2 + (size_t)j * k, k * sizeof(float));
                out_count[b] = c2[j];
                // -1 = answered by the fp32 fallback (certified); stays FALLBACK for the next attempt otherwise
                stt[b] = (t2[j] == FRP_STATUS_OK) ? -1 : FRP_STATUS_FALLBACK;
            }
            (void)L2;
        }
        if (e != cudaSuccess) {
            set_error("frp_match: CUDA error in the fallback pass: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return FRP_E_CUDA;
        }
        if (rc) return rc;
        for (int b = 0; b < B; ++b) {
            // -1: certified by the fallback -> FRP_STATUS_FALLBACK (informational); still FALLBACK
            // after both attempts -> FRP_STATUS_UNCERTIFIED (> 64 exact ties at the cut)
            if (stt[b] == -1) stt[b] = FRP_STATUS_FALLBACK;
            else if (stt[b] == FRP_STATUS_FALLBACK) stt[b] = FRP_STATUS_UNCERTIFIED;
        }
        if (out_status) memcpy(out_status, stt.data(), cb);
    } else if (out_status) {
        memcpy(out_status, status, cb);
    }
    return FRP_OK;
}

// caller holds the shared lock and has set the device; rows <= out_stride has been checked
static int scores_device_locked(frp_gallery_t* g, const float* queries_dev, int B, double* out_dev,
                                uint64_t out_stride, cudaStream_t st) {
    if (g->rows == 0) return FRP_OK;
    Workspace* w = workspace_for(g->device, st);
    const size_t qprep_bytes = align256((size_t)B * g->dimp * sizeof(float));
    int rc = workspace_reserve(w, qprep_bytes + align256((size_t)B * sizeof(QParams)), 0);
    if (rc) return rc;
    float* qprep = reinterpret_cast<float*>(w->dev);
    QParams* qp = reinterpret_cast<QParams*>(w->dev + qprep_bytes);
    rc = launch_prep_queries(g, queries_dev, (size_t)g->dim, nullptr, B, 0.0, FRP_MATCH_NO_THRESHOLD, FRP_EPS_SCAN_F32,
                             qprep, nullptr, qp, st);
    if (!rc) rc = launch_scores(g, qprep, B, out_dev, out_stride, st);
    return rc;
}

int frp_scores_device(frp_gallery_t* g, const float* queries_dev, int B, double* out_dev, uint64_t out_stride,
                      uint64_t* rows_out, void* stream) {
    if (!g || !queries_dev || !out_dev || B <= 0) { set_error("frp_scores_device: invalid argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    if (rows_out) *rows_out = g->rows;
    if (g->rows > out_stride) {
        set_error("frp_scores_device: gallery has %llu rows, the output stride is %llu",
                  (unsigned long long)g->rows, (unsigned long long)out_stride);
        return FRP_E_RANGE;
    }
    DeviceGuard dg(g->device);
    return scores_device_locked(g, queries_dev, B, out_dev, out_stride, (cudaStream_t)stream);
}

int frp_scores(frp_gallery_t* g, const float* queries, int B, double* out, uint64_t out_stride, uint64_t* rows_out) {
    if (!g || !queries || !out || B <= 0) { set_error("frp_scores: invalid argument"); return FRP_E_INVALID; }
    // ONE shared lock across the row-count read, the kernel and the copy: an append between
    // "size the buffer" and "fill it" cannot happen
    std::shared_lock<std::shared_mutex> lk(g->mu);
    const uint64_t rows = g->rows;
    if (rows_out) *rows_out = rows;
    if (rows > out_stride) {
        set_error("frp_scores: gallery has %llu rows, the output stride is %llu", (unsigned long long)rows,
                  (unsigned long long)out_stride);
        return FRP_E_RANGE;
    }
    if (rows == 0) return FRP_OK;
    DeviceGuard dg(g->device);
    cudaStream_t st = thread_stream(g->device);
    // dense dumps are large (B x rows doubles): a call-scoped stream-ordered allocation, not
    // workspace; the pool keeps freed blocks so repeated calls do not go back to the driver
    {
        static std::once_flag once[16];
        const int dv = g->device;
        if (dv >= 0 && dv < 16)
            std::call_once(once[dv], [dv]() {
                cudaMemPool_t pool;
                if (cudaDeviceGetDefaultMemPool(&pool, dv) == cudaSuccess) {
                    uint64_t thr = ~0ull;
                    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
                }
                cudaGetLastError();
            });
    }
    float* dq = nullptr; double* dout = nullptr;
    FRP_CUDA(cudaMallocAsync((void**)&dq, (size_t)B * g->dim * sizeof(float), st));
    cudaError_t e = cudaMallocAsync((void**)&dout, (size_t)B * rows * sizeof(double), st);
    int rc = FRP_OK;
    if (e == cudaSuccess) e = cudaMemcpyAsync(dq, queries, (size_t)B * g->dim * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = scores_device_locked(g, dq, B, dout, rows, st);
    if (e == cudaSuccess && rc == FRP_OK)
        e = cudaMemcpy2DAsync(out, out_stride * sizeof(double), dout, rows * sizeof(double), rows * sizeof(double),
                              (size_t)B, cudaMemcpyDeviceToHost, st);
    if (dq) cudaFreeAsync(dq, st);
    if (dout) cudaFreeAsync(dout, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) { set_error("frp_scores: CUDA error: %s", cudaGetErrorString(e)); cudaGetLastError(); return FRP_E_CUDA; }
    return rc;
}

}  // extern "C"
