// Internal declarations shared by the translation units of libfrp_b200.so.
// sm_100a only; no CPU fallback anywhere.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <shared_mutex>
#include <vector>

#include "frp_b200.h"

#define FRP_NUM_SM_DEFAULT 148

struct frp_gallery {
    int device = 0;
    int dim = 0;    // caller-visible embedding length
    int dimp = 0;   // padded length: power of two >= 64 (zero padded)
    int metric = FRP_METRIC_L2;
    uint64_t capacity = 0;   // allocated rows
    uint64_t rows = 0;       // appended rows incl. tombstones
    uint64_t dead = 0;       // tombstones
    uint64_t base_row = 0;   // global index of local row 0 (row sharding)
    __half* g16 = nullptr;   // [capacity][dimp] scan copy (COS: normalised)
    float* g32 = nullptr;    // [capacity][dimp] fp32 master (COS: normalised)
    float* norm2 = nullptr;  // [capacity] |g|^2 of the master row
    uint8_t* alive = nullptr;  // [capacity] 1 = live, 0 = tombstone
    // optional [capacity] global insertion sequence of each row (mutable row-sharded galleries:
    // appends go to the least-full shard, so global order != base_row + local row); when set,
    // results carry seq[row] and the cross-shard merge breaks ties by it
    uint64_t* seq = nullptr;
    float* d_gmax = nullptr;   // device [2]: max |g| seen, max |g - fp16(g)| seen (monotone upper bounds)
    float gmax = 0.f;          // host copies
    float gerr = 0.f;
    int sm_count = FRP_NUM_SM_DEFAULT;
    // host mirror of `alive` (mutations check it instead of a blocking device read)
    std::vector<uint8_t> h_alive;
    // row mutations: private stream + pinned staging row, so an update is one small H2D + one
    // kernel + an 8-byte read-back on that stream (no cudaMalloc, no device-wide sync)
    cudaStream_t mut_stream = nullptr;
    float* h_stage = nullptr;     // pinned [dim] + 2 floats (gmax read-back)
    float* d_stage = nullptr;     // device [dim]
    // TMA descriptors of the fp16 scan copy ([capacity][dimp], boxes of 256 and 128 rows x 64
    // halfs), encoded when the buffers are (re)allocated, i.e. under the unique lock
    alignas(64) unsigned char tmap_g[128];
    alignas(64) unsigned char tmap_gh[128];
    bool tmap_valid = false;
    mutable std::shared_mutex mu;  // match = shared, mutate = unique
};

namespace frp {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void set_match_path(const char* name);

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// ---- per-query parameters produced by prep_queries -------------------------
struct QParams {
    float cut0;       // initial prefilter cut in s-space (rounded down), -inf when unthresholded
    float eps;        // bound on |s_approx - s_exact| for the scan copy in use
    double s_thresh;  // exact threshold in s-space (COS: tau; L2: |q|^2 - tau^2)
    double qnorm2;    // |q|^2 of the (fp32) query  (L2)
};

// s-space: larger is better.  COS: s = g.q ; L2: s = 2 g.q - |g|^2  (d^2 = |q|^2 - s)
// candidate key: (orderable(score) << 32) | ~row  -> descending u64 order ==
// score descending, row ascending.  0 is the empty sentinel.
__host__ __device__ inline uint32_t f32_orderable(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float f32_from_orderable(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ inline uint64_t make_key(float s, uint32_t row) {
    return ((uint64_t)f32_orderable(s) << 32) | (uint64_t)(~row);
}
__host__ __device__ inline float key_score(uint64_t k) { return f32_from_orderable((uint32_t)(k >> 32)); }
__host__ __device__ inline uint32_t key_row(uint64_t k) { return ~(uint32_t)(k & 0xffffffffu); }

// prefilter error-bound modes (see DESIGN.md "exactness certificate")
#define FRP_EPS_SCAN_F16 0   /* fp16 gallery copy x fp32 query */
#define FRP_EPS_GEMM_F16 1   /* fp16 gallery copy x fp16 query (tcgen05) */
#define FRP_EPS_SCAN_F32 2   /* fp32 master x fp32 query */

// ---- kernels / launchers (defined in the .cu files) --------------------------
// prep: normalise (COS), pad to dimp, make fp16 copy, fill QParams
// q_dev rows are `qstride` floats apart (dim for caller queries, dimp for gallery master rows);
// q_rows (optional, device) gathers the queries from q_dev by row index (self-match)
int launch_prep_queries(const frp_gallery* g, const float* q_dev, size_t qstride, const uint64_t* q_rows,
                        int B, double tau, int flags, int eps_mode, float* qprep, __half* q16,
                        QParams* qp, cudaStream_t st);

// streaming scan prefilter: writes cand[B][nlists][KP] keys. use_master: scan fp32 master.
int scan_num_lists(const frp_gallery* g, bool use_master);
int launch_match_scan(const frp_gallery* g, bool use_master, const float* qprep, const QParams* qp,
                      int B, int KP, uint64_t* cand, int nlists, cudaStream_t st);

// tcgen05 prefilter (B > 4): same candidate-list contract.
bool gemm_supported(const frp_gallery* g);
int gemm_num_lists(const frp_gallery* g, int B);
// tmap_q: the caller's cached descriptor of q16 ([groups*128][dimp] fp16), see Workspace
int launch_match_gemm(frp_gallery* g, const void* tmap_q, const QParams* qp, int B, int KP,
                      uint64_t* cand, int nlists, cudaStream_t st, uint64_t rows_limit = 0);
// galleries of a few thousand rows: everything (prep, exact scores, threshold, top-k) in one launch
bool match_small_supported(const frp_gallery* g, int B);
// q_host (optional): host-readable copy of the same queries; small blocks ride in the kernel parameters
int launch_match_small(const frp_gallery* g, const float* q_dev, size_t qstride, int B, double tau, int k, int flags,
                       int64_t* out_idx, float* out_score, double* out_score64, int32_t* out_count,
                       int32_t* out_status, cudaStream_t st, unsigned int* done_count = nullptr,
                       volatile unsigned int* done_flag = nullptr, unsigned int done_seq = 0,
                       const float* q_host = nullptr);
// pure top-k: raise every query's prefilter cut to (k-th best key of `cand`) - 2 eps
int launch_raise_cut(QParams* qp, int B, int k, const uint64_t* cand, int nlists, int KP, cudaStream_t st);
int encode_query_map(void* tmap_out, const __half* q16, int groups, int dimp);

// (re)encode the gallery's cached tensor maps; caller holds the unique lock (or owns g exclusively)
void gemm_refresh_maps(frp_gallery* g);

// cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint (no libcuda link); null if unavailable
void* tensormap_encode_fn();

// exact fp64 re-rank of the survivors + certificate + output
// out_score (f32) and out_score64 (f64) are both optional; at least one is written
int launch_rerank(const frp_gallery* g, const float* qprep, const QParams* qp, int B, int k, int KP,
                  int flags, double tau, const uint64_t* cand, int nlists, int64_t* out_idx,
                  float* out_score, double* out_score64, int32_t* out_count, int32_t* out_status,
                  cudaStream_t st);

int launch_scores(const frp_gallery* g, const float* qprep, int B, double* out, uint64_t out_stride,
                  cudaStream_t st);

// Per-(thread, stream) scratch of the match pipeline: grows on demand, never shrinks, so the
// steady state issues no allocation.  Work on one stream is ordered, hence reuse is safe; a
// thread that drives several streams gets one workspace per stream.
struct Workspace {
    int device = -1;
    uint8_t* dev = nullptr;       // device scratch
    size_t dev_bytes = 0;
    uint8_t* pinned = nullptr;    // pinned host staging (host API only)
    size_t pinned_bytes = 0;
    // cached query tensor map (valid for q16_ptr / groups / dimp)
    alignas(64) unsigned char tmap_q[128];
    const void* q16_ptr = nullptr;
    int q_groups = 0, q_dimp = 0;
    // completion word of the small-gallery kernel: device counter + mapped host flag + sequence
    unsigned int* done_count = nullptr;
    volatile unsigned int* done_flag = nullptr;
    unsigned int done_seq = 0;
};
Workspace* workspace_for(int device, cudaStream_t st);
int workspace_reserve(Workspace* w, size_t dev_bytes, size_t pinned_bytes);

// caller holds (at least) the shared lock and has set the device.  q_rows: see launch_prep_queries.
int match_device_locked(frp_gallery* g, const float* q_dev, size_t qstride, const uint64_t* q_rows, int B,
                        double tau, int k, int flags, int KP, int64_t* out_idx, float* out_score,
                        double* out_score64, int32_t* out_count, int32_t* out_status, cudaStream_t st);
int pick_kp(int k);
cudaStream_t thread_stream(int device);
void configure_func_smem(const void* func, int device, size_t bytes, int* rc);
void configure_func_carveout(const void* func, int device);
// programmatic dependent launches between the kernels of one match (FRP_PDL=0 turns them off)
bool pdl_enabled();

}  // namespace frp

#define FRP_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            frp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                       \
            return FRP_E_CUDA;                                                              \
        }                                                                                   \
    } while (0)

#define FRP_CHECK_LAUNCH()                                                                  \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            frp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),     \
                           __FILE__, __LINE__);                                             \
            return FRP_E_CUDA;                                                              \
        }                                                                                   \
        frp::count_launch();                                                                \
    } while (0)
