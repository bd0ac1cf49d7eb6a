// Internal declarations shared by the translation units of libfrp_b200.so.
// sm_100a only; no CPU fallback anywhere.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <shared_mutex>

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
    float* d_gmax = nullptr;   // device [2]: max |g| seen, max |g - fp16(g)| seen (monotone upper bounds)
    float gmax = 0.f;          // host copies
    float gerr = 0.f;
    int sm_count = FRP_NUM_SM_DEFAULT;
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
int launch_prep_queries(const frp_gallery* g, const float* q_dev, int B, double tau, int flags,
                        int eps_mode, float* qprep, __half* q16, QParams* qp, cudaStream_t st);

// streaming scan prefilter: writes cand[B][nlists][KP] keys. use_master: scan fp32 master.
int scan_num_lists(const frp_gallery* g, bool use_master);
int launch_match_scan(const frp_gallery* g, bool use_master, const float* qprep, const QParams* qp,
                      int B, int KP, uint64_t* cand, int nlists, cudaStream_t st);

// tcgen05 prefilter (B > 4): same candidate-list contract.
bool gemm_supported(const frp_gallery* g);
int gemm_num_lists(const frp_gallery* g, int B);
int launch_match_gemm(frp_gallery* g, const __half* q16, const QParams* qp, int B, int KP,
                      uint64_t* cand, int nlists, cudaStream_t st);

// cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint (no libcuda link); null if unavailable
void* tensormap_encode_fn();

// exact fp64 re-rank of the survivors + certificate + output
int launch_rerank(const frp_gallery* g, const float* qprep, const QParams* qp, int B, int k, int KP,
                  int flags, double tau, const uint64_t* cand, int nlists, int64_t* out_idx,
                  float* out_score, int32_t* out_count, int32_t* out_status, cudaStream_t st);

int launch_scores(const frp_gallery* g, const float* qprep, int B, double* out, cudaStream_t st);

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
