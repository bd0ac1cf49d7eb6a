// Query preparation, exact re-rank + certificate (K6), dense score dump, and the
// cross-shard top-k merge.
//
// Replaces the tail of face_service.compare_faces (face_service.py:411-432):
// threshold, stable sort by distance, and the routes' `[:top_k]`
// (routes/face.py:692-693); find_k_nearest's argpartition/argsort
// (face_service.py:601-603).
#include <math.h>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

namespace {

constexpr int kRCap = 256;          // max survivors re-ranked per query (k <= 64 plus the 2-eps band)
constexpr int kRerankThreads = 256;

__device__ __forceinline__ float float_round_down(double x) {
    return __double2float_rd(x);
}

// ---------------------------------------------------------------------------
// prep_queries: one block per query.
// ---------------------------------------------------------------------------
// eps_mode: 0 = fp16 gallery x fp32 query (scan), 1 = fp16 x fp16 (tcgen05), 2 = fp32 master
// scan.  The prefilter error bound is rigorous at any scale (Cauchy-Schwarz on the MEASURED
// rounding errors): |s~ - s| <= |g16-g|*|q16| + |g|*|q16-q| + fp32 accumulation.
__global__ void prep_queries_kernel(const float* __restrict__ q, size_t qstride,
                                    const uint64_t* __restrict__ q_rows, int dim, int dimp, int metric,
                                    double tau, int flags, int eps_mode, float gmax, float gerr,
                                    float* __restrict__ qprep, __half* __restrict__ q16,
                                    QParams* __restrict__ qp) {
    __shared__ double red[32];
    __shared__ double red2[32];
    griddep_launch();      // the prefilter may set itself up (barriers, TMEM, first gallery tiles) meanwhile
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float* src = q + (q_rows ? (size_t)q_rows[b] : (size_t)b) * qstride;
    double acc = 0.0;
    for (int i = tid; i < dim; i += blockDim.x) {
        const double x = (double)src[i];
        acc += x * x;
    }
    acc = warp_sum_f64(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {
        double t = (tid < (blockDim.x >> 5)) ? red[tid] : 0.0;
        t = warp_sum_f64(t);
        if (tid == 0) red[0] = t;
    }
    __syncthreads();
    const double n2 = red[0];
    const double nrm = sqrt(n2);
    double scale = 1.0;
    if (metric == FRP_METRIC_COS) scale = n2 > 0.0 ? 1.0 / nrm : 0.0;
    double qe2 = 0.0, qp2 = 0.0;      // |q16 - q|^2 and |q|^2 of the prepared query
    for (int i = tid; i < dimp; i += blockDim.x) {
        const float v = (i < dim) ? (float)((double)src[i] * scale) : 0.f;
        qprep[(size_t)b * dimp + i] = v;
        qp2 += (double)v * (double)v;
        if (q16) {
            const __half h = __float2half_rn(v);
            q16[(size_t)b * dimp + i] = h;
            const double d = (double)v - (double)__half2float(h);
            qe2 += d * d;
        }
    }
    qe2 = warp_sum_f64(qe2);
    qp2 = warp_sum_f64(qp2);
    __syncthreads();
    if ((tid & 31) == 0) { red[tid >> 5] = qe2; red2[tid >> 5] = qp2; }
    __syncthreads();
    if (tid == 0) {
        double e2 = 0.0, p2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { e2 += red[w]; p2 += red2[w]; }
        QParams p;
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const float qn = __double2float_ru(sqrt(p2)) * 1.000001f;               // |q| (prepared)
        float qerr = (eps_mode == 1) ? __double2float_ru(sqrt(e2)) : 0.f;     // |q16 - q|
        if (!(qerr == qerr)) qerr = __int_as_float(0x7f800000);
        const float ge = (eps_mode == 2) ? 0.f : gerr;
        // operand rounding (exact bound) + fp32 accumulation (2^-23 per term, truncating adders)
        const float eps_dot = ge * (qn + qerr) + gmax * qerr +
                              (float)dimp * 1.2e-7f * (gmax + ge) * (qn + qerr);
        if (metric == FRP_METRIC_COS) {
            p.eps = eps_dot * 1.0001f + 1e-7f;
            p.s_thresh = tau;
            p.qnorm2 = 1.0;
        } else {
            // s = 2 g.q - |g|^2: twice the dot error + rounding of the stored norm and of the fma
            p.eps = 2.f * eps_dot * 1.0001f + 2.4e-7f * (gmax * gmax + 2.f * gmax * qn) + 1e-12f;
            p.s_thresh = (tau >= 0.0) ? (n2 - tau * tau) : inf;
            p.qnorm2 = n2;
        }
        if (flags & FRP_MATCH_NO_THRESHOLD) {
            p.cut0 = __int_as_float(0xff800000);   // -inf
            p.s_thresh = -inf;
        } else {
            p.cut0 = float_round_down(p.s_thresh - (double)p.eps);
        }
        qp[b] = p;
    }
}

// ---------------------------------------------------------------------------
// rerank: one block per query.
// ---------------------------------------------------------------------------
struct RerankArgs {
    const float* g32;
    const float* qprep;
    const QParams* qp;
    const uint64_t* cand;
    int nlists, KP, k, dimp, metric, flags, npad;
    double tau;
    uint64_t base_row;
    const uint64_t* seq;    // optional per-row global sequence (overrides base_row + row)
    int64_t* out_idx;
    float* out_score;       // optional
    double* out_score64;    // optional (shard record: the merge ranks on the exact value)
    int32_t* out_count;
    int32_t* out_status;
};

__device__ __forceinline__ bool better(int metric, double va, uint32_t ra, double vb, uint32_t rb) {
    if (metric == FRP_METRIC_COS) return va > vb || (va == vb && ra < rb);
    return va < vb || (va == vb && ra < rb);
}

__global__ void __launch_bounds__(kRerankThreads) rerank_kernel(const RerankArgs a) {
    extern __shared__ __align__(16) uint64_t keys[];   // [npad]
    __shared__ int s_cnt, s_m, s_npass;
    __shared__ unsigned int s_lfull;
    __shared__ double s_kth;
    __shared__ double ex_s[kRCap], ex_v[kRCap];
    __shared__ uint32_t ex_row[kRCap];
    __shared__ uint8_t ex_pass[kRCap];

    griddep_wait();        // launched as a programmatic dependent of the prefilter
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const QParams qp = a.qp[b];
    if (tid == 0) { s_cnt = 0; s_lfull = 0u; s_m = 0; s_npass = 0; s_kth = 0.0; }
    __syncthreads();

    // 1. compact non-empty candidate keys; track the weakest key of every FULL list
    const int total = a.nlists * a.KP;
    const uint64_t* cand = a.cand + (size_t)b * total;
    for (int i = tid; i < total; i += blockDim.x) {
        const uint64_t key = cand[i];
        if (key) {
            const int pos = atomicAdd(&s_cnt, 1);
            keys[pos] = key;
            if ((i % a.KP) == a.KP - 1) atomicMax(&s_lfull, (unsigned int)(key >> 32));
        }
    }
    __syncthreads();
    const int n = s_cnt;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + tid; i < np2; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();

    // 2. block bitonic sort, descending (score desc, row asc)
    for (int size = 2; size <= np2; size <<= 1) {
        for (int stride = size >> 1; stride >= 1; stride >>= 1) {
            for (int i = tid; i < (np2 >> 1); i += blockDim.x) {
                const int lo = ((i / stride) * (stride << 1)) + (i % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const uint64_t x = keys[lo], y = keys[hi];
                if (desc ? (x < y) : (x > y)) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }

    // 3. survivors: everything within 2*eps of the k-th best approximate score
    const int k = a.k;
    if (n >= k && n > 0) {
        const float thr = key_score(keys[k - 1]) - 2.f * qp.eps;
        for (int i = tid; i < n; i += blockDim.x) {
            const bool p = key_score(keys[i]) >= thr;
            const bool pn = (i + 1 < n) ? (key_score(keys[i + 1]) >= thr) : false;
            if (p && !pn) s_m = i + 1;
        }
    } else if (tid == 0) {
        s_m = n;
    }
    __syncthreads();
    const bool overflow = s_m > kRCap;
    const int m = overflow ? kRCap : s_m;

    // 4. exact fp64 re-rank from the fp32 master (one warp per survivor)
    const float4* qv = reinterpret_cast<const float4*>(a.qprep + (size_t)b * a.dimp);
    const int nvec = a.dimp >> 2;
    for (int c = warp; c < m; c += (blockDim.x >> 5)) {
        const uint32_t row = key_row(keys[c]);
        const float4* gr = reinterpret_cast<const float4*>(a.g32 + (size_t)row * a.dimp);
        double acc = 0.0;
        if (a.metric == FRP_METRIC_COS) {
            for (int i = lane; i < nvec; i += 32) {
                const float4 g = __ldg(gr + i);
                const float4 q = __ldg(qv + i);
                acc += (double)g.x * (double)q.x + (double)g.y * (double)q.y +
                       (double)g.z * (double)q.z + (double)g.w * (double)q.w;
            }
        } else {
            for (int i = lane; i < nvec; i += 32) {
                const float4 g = __ldg(gr + i);
                const float4 q = __ldg(qv + i);
                const double dx = (double)g.x - (double)q.x, dy = (double)g.y - (double)q.y;
                const double dz = (double)g.z - (double)q.z, dw = (double)g.w - (double)q.w;
                acc += dx * dx + dy * dy + dz * dz + dw * dw;
            }
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) {
            ex_row[c] = row;
            if (a.metric == FRP_METRIC_COS) {
                ex_v[c] = acc;
                ex_s[c] = acc;
            } else {
                ex_v[c] = sqrt(acc);
                ex_s[c] = qp.qnorm2 - acc;
            }
        }
    }
    __syncthreads();

    // 5. exact threshold (inclusive like face_service.py:411 unless STRICT), rank, emit
    const bool thresholded = !(a.flags & FRP_MATCH_NO_THRESHOLD);
    const bool strict = (a.flags & FRP_MATCH_STRICT) != 0;
    if (tid < m) {
        const double v = ex_v[tid];
        bool p = true;
        if (thresholded) {
            if (a.metric == FRP_METRIC_COS) p = strict ? (v > a.tau) : (v >= a.tau);
            else p = strict ? (v < a.tau) : (v <= a.tau);
        }
        if (!(v == v)) p = false;
        ex_pass[tid] = p ? 1 : 0;
        if (p) atomicAdd(&s_npass, 1);
    }
    __syncthreads();
    const int npass = s_npass;
    if (tid < m && ex_pass[tid]) {
        int rank = 0;
        const double v = ex_v[tid];
        const uint32_t r = ex_row[tid];
        for (int c = 0; c < m; ++c)
            if (c != tid && ex_pass[c] && better(a.metric, ex_v[c], ex_row[c], v, r)) ++rank;
        if (rank < k) {
            a.out_idx[(size_t)b * k + rank] = a.seq ? (int64_t)a.seq[r] : (int64_t)(a.base_row + r);
            if (a.out_score) a.out_score[(size_t)b * k + rank] = (float)v;
            if (a.out_score64) a.out_score64[(size_t)b * k + rank] = v;
            if (rank == k - 1) s_kth = ex_s[tid];
        }
    }
    const int cnt = npass < k ? npass : k;
    for (int i = cnt + tid; i < k; i += blockDim.x) {
        a.out_idx[(size_t)b * k + i] = -1;
        if (a.out_score) a.out_score[(size_t)b * k + i] = __int_as_float(0x7fc00000);
        if (a.out_score64) a.out_score64[(size_t)b * k + i] = __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncthreads();

    // 6. certificate: rows outside the candidate lists have s_approx <= Lfull
    //    (or < cut0), hence s_exact <= Lfull + eps (or < s_thresh).
    if (tid == 0) {
        a.out_count[b] = cnt;
        // a non-finite bound (fp16 overflow of the gallery copy or of the query) certifies nothing:
        // NaN/inf approximate scores never enter the lists
        bool ok = !overflow && (qp.eps < 3.0e38f);
        const bool lf_valid = s_lfull != 0u;
        const double bound = (double)f32_from_orderable(s_lfull) + (double)qp.eps;
        if (lf_valid) {
            if (npass >= k) ok = ok && (s_kth > bound);
            else ok = ok && thresholded && (bound < qp.s_thresh);
        }
        if (a.out_status) a.out_status[b] = ok ? FRP_STATUS_OK : FRP_STATUS_FALLBACK;
    }
}

// ---------------------------------------------------------------------------
// raise_cut: pure top-k (no threshold) keeps a live top-KP per prefilter thread, and the insert
// rate of a list that starts from -inf only decays like KP/rows_seen.  A sampling pass over the
// first ~1.5 % of the gallery gives every query k real candidates; every row of the final top-k
// must score at least as well as the k-th of them, so the main pass can run thresholded at
//     cut = (k-th best approximate score of the sample) - 2 eps
// (a row of the true top-k has exact score >= the exact score of that sample row >= its
// approximate score - eps, and its own approximate score is within eps of the exact one).
// One block per query: compact the sample's keys, find the k-th largest, lift cut0.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raise_cut_kernel(QParams* __restrict__ qp, int k,
                                                        const uint64_t* __restrict__ cand, int total) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_prefix, s_need;
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint64_t* c = cand + (size_t)b * total;
    // radix select on the 32-bit orderable score (4 passes of 8 bits, most significant first):
    // the k-th largest key's score without sorting the ~5 000 keys
    unsigned int prefix = 0, need = (unsigned int)k;
    bool enough = true;
    for (int pass = 0; pass < 4 && enough; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < total; i += blockDim.x) {
            const uint64_t key = c[i];
            if (!key) continue;
            const unsigned int o = (unsigned int)(key >> 32);
            if (pass == 0 || (o >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(o >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int acc = 0;
            int d = 255;
            for (; d >= 0; --d) {
                if (acc + s_hist[d] >= need) break;
                acc += s_hist[d];
            }
            if (d < 0) { s_need = 0xffffffffu; }              // fewer than k keys in the sample
            else { s_prefix = prefix | ((unsigned int)d << shift); s_need = need - acc; }
        }
        __syncthreads();
        if (s_need == 0xffffffffu) enough = false;
        else { prefix = s_prefix; need = s_need; }
        __syncthreads();
    }
    if (tid == 0 && enough) {
        QParams p = qp[b];
        const float kth = f32_from_orderable(prefix);
        const float lifted = __fadd_rd(kth, -2.f * p.eps);
        if (lifted == lifted && lifted > p.cut0) { p.cut0 = lifted; qp[b] = p; }
    }
}

// ---------------------------------------------------------------------------
// match_small: the reference's own scale (BASELINE configs[0]: 1 query vs 1 000 x 128-d) in ONE
// launch.  A gallery of a few thousand rows needs no prefilter and no certificate: one block per
// query normalises the query (COS), evaluates EVERY live row exactly (fp64 from the fp32 master,
// one warp per row, the same expression and summation order as rerank_kernel, so the values are
// bit-identical to the large-gallery path) and thresholds.  Top-k in (score, then row) order, stable
// like the reference's sort: up to 32 candidates (the usual thresholded query) are ranked by one
// warp; more (pure top-k) go through per-warp k-best lists and a ranking of their union.
// ---------------------------------------------------------------------------
// The rows of one query are split over a cluster of kSmallCluster CTAs (8 SMs pull the gallery
// instead of one); every CTA stores its exact values into CTA 0's shared memory (DSMEM), and CTA 0
// extracts the top-k after one cluster barrier.  `done_flag` (mapped host memory, optional): the
// last block to finish publishes `done_seq` there after a system fence, so the host can poll one
// word instead of paying a stream synchronisation.
constexpr int kSmallRows = 4096;
constexpr int kSmallCluster = 8;
constexpr int kSmallInline = 512;    // floats of query data carried in the kernel parameters
constexpr int kSmallRank = 32;       // up to this many candidates are ranked by one warp
constexpr int kSmallUnroll = 4;      // gallery rows in flight per warp (the block is latency-bound)
constexpr int small_threads(int cs) { return cs == 1 ? 1024 : 512; }

struct SmallArgs {
    const float* g32;
    const uint8_t* alive;
    const uint64_t* seq;
    const float* q;
    size_t qstride;
    int rows, dim, dimp, metric, flags, k;
    double tau;
    uint64_t base_row;
    int64_t* out_idx;
    float* out_score;
    double* out_score64;
    int32_t* out_count;
    int32_t* out_status;
    unsigned int* done_count;        // device counter of finished queries (zero between launches)
    volatile unsigned int* done_flag;   // mapped host word, or null
    unsigned int done_seq;
    // a query block of up to kSmallInline floats travels in the kernel parameters themselves (the
    // single-query compare_faces call): no trip over PCIe before the first gallery row is read
    int inline_q;
    float qinline[kSmallInline];
};

__device__ __forceinline__ uint32_t small_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void small_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void small_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void small_store_rank0(double* local_ptr, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(0));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(small_threads(CS)) match_small_kernel(const __grid_constant__ SmallArgs a) {
    constexpr int kSmallThreads = small_threads(CS);
    extern __shared__ __align__(16) uint8_t small_smem[];
    double* sc = reinterpret_cast<double*>(small_smem);                       // [rows] exact value (NaN = not a candidate)
    float* qv = reinterpret_cast<float*>(small_smem + (((size_t)a.rows * 8 + 15) & ~(size_t)15));   // [dimp] prepared query
    __shared__ double red[32];
    __shared__ int s_ncand, s_nunion;
    __shared__ int s_cand[kSmallRank];
    __shared__ int s_union[(small_threads(CS) / 32) * 64];     // every warp's k <= 64 best rows
    const int b = (CS == 1) ? blockIdx.x : blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kSmallThreads >> 5;
    const int rank = (CS == 1) ? 0 : (int)small_cluster_rank();
    if (CS > 1) small_cluster_arrive();      // CTA 0's shared memory exists once every CTA has arrived
    // this CTA's slice of the rows (whole groups of kSmallUnroll)
    const int per = (((a.rows + CS - 1) / CS + kSmallUnroll - 1) / kSmallUnroll) * kSmallUnroll;
    const int lo = min(rank * per, a.rows), hi = min(lo + per, a.rows);
    const float* src = a.q + (size_t)b * a.qstride;
    // ---- query prep.  The raw query is staged once (it may live in mapped host memory: one trip
    // over PCIe); |q|^2 is summed in prep_queries_kernel's order (128 strided partial sums, warp
    // butterflies, then 4 -> 1), so the prepared query is bit-identical to the large-gallery path ----
    if (tid == 0) { s_ncand = 0; s_nunion = 0; }
    if (a.inline_q) {
        for (int i = tid; i < a.dimp; i += blockDim.x) qv[i] = (i < a.dim) ? a.qinline[b * a.dim + i] : 0.f;
    } else {
        for (int i = tid; i < a.dimp; i += blockDim.x) qv[i] = (i < a.dim) ? src[i] : 0.f;
    }
    __syncthreads();
    if (a.metric == FRP_METRIC_COS) {
        double acc = 0.0;
        if (tid < 128)
            for (int i = tid; i < a.dim; i += 128) { const double x = (double)qv[i]; acc += x * x; }
        acc = warp_sum_f64(acc);
        if (lane == 0 && warp < 4) red[warp] = acc;
        __syncthreads();
        if (tid < 32) {
            double t = (tid < 4) ? red[tid] : 0.0;
            t = warp_sum_f64(t);
            if (tid == 0) red[0] = t;
        }
        __syncthreads();
        const double n2 = red[0];
        const double scale = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
        for (int i = tid; i < a.dim; i += blockDim.x) qv[i] = (float)((double)qv[i] * scale);
        __syncthreads();
    }
    // ---- every live row, exactly: kSmallUnroll rows per warp in flight; each row's partial sums
    // are accumulated in the order rerank_kernel uses, so the values are bit-identical ----
    const bool thresholded = !(a.flags & FRP_MATCH_NO_THRESHOLD);
    const bool strict = (a.flags & FRP_MATCH_STRICT) != 0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const float4* q4 = reinterpret_cast<const float4*>(qv);
    const int nvec = a.dimp >> 2;
    if (CS > 1) small_cluster_wait();
    for (int row0 = lo + warp * kSmallUnroll; row0 < hi; row0 += NW * kSmallUnroll) {
        double s[kSmallUnroll];
#pragma unroll
        for (int u = 0; u < kSmallUnroll; ++u) s[u] = 0.0;
        for (int i = lane; i < nvec; i += 32) {
            float4 g[kSmallUnroll];
#pragma unroll
            for (int u = 0; u < kSmallUnroll; ++u) {
                const int row = min(row0 + u, a.rows - 1);
                g[u] = __ldg(reinterpret_cast<const float4*>(a.g32 + (size_t)row * a.dimp) + i);
            }
            const float4 q = q4[i];
#pragma unroll
            for (int u = 0; u < kSmallUnroll; ++u) {
                if (a.metric == FRP_METRIC_COS) {
                    s[u] += (double)g[u].x * (double)q.x + (double)g[u].y * (double)q.y + (double)g[u].z * (double)q.z +
                            (double)g[u].w * (double)q.w;
                } else {
                    const double dx = (double)g[u].x - (double)q.x, dy = (double)g[u].y - (double)q.y;
                    const double dz = (double)g[u].z - (double)q.z, dw = (double)g[u].w - (double)q.w;
                    s[u] += dx * dx + dy * dy + dz * dz + dw * dw;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kSmallUnroll; ++u) {
            const double t = warp_sum_f64(s[u]);
            const int row = row0 + u;
            if (lane == 0 && row < hi) {
                double v = (a.metric == FRP_METRIC_COS) ? t : sqrt(t);
                bool p = a.alive[row] != 0 && (v == v);
                if (p && thresholded) {
                    if (a.metric == FRP_METRIC_COS) p = strict ? (v > a.tau) : (v >= a.tau);
                    else p = strict ? (v < a.tau) : (v <= a.tau);
                }
                if (CS == 1) sc[row] = p ? v : nan;
                else small_store_rank0(&sc[row], p ? v : nan);
            }
        }
    }
    if (CS > 1) {
        small_cluster_arrive();
        small_cluster_wait();
        if (rank != 0) return;
    } else {
        __syncthreads();
    }
    // ---- top-k.  The usual thresholded query has a handful of candidates: list them, and when
    // there are at most kSmallRank one warp ranks them (rank = number of candidates that are
    // better in (value, row) order: a strict total order, so the ranks are a permutation) ----
    double last_v = 0.0;
    int last_r = -1;
    int cnt = 0;
    for (int row0 = 0; row0 < a.rows; row0 += blockDim.x) {      // one shared atomic per warp, not per row
        const int row = row0 + tid;
        const bool is_cand = row < a.rows && sc[row] == sc[row];
        const unsigned int m = __ballot_sync(FRP_FULL_MASK, is_cand);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_ncand, __popc(m));
            base = __shfl_sync(FRP_FULL_MASK, base, 0);
            const int p = base + __popc(m & ((1u << lane) - 1u));
            if (is_cand && p < kSmallRank) s_cand[p] = row;
        }
    }
    __syncthreads();
    const int ncand = s_ncand;
    if (ncand <= kSmallRank) {
        if (warp == 0) {
            const int mr = (lane < ncand) ? s_cand[lane] : -1;
            const double mv = (mr >= 0) ? sc[mr] : nan;
            int rk = 0;
            for (int j = 0; j < ncand; ++j) {
                const double ov = __shfl_sync(FRP_FULL_MASK, mv, j);
                const int orr = __shfl_sync(FRP_FULL_MASK, mr, j);
                if (mr >= 0 && better(a.metric, ov, (uint32_t)orr, mv, (uint32_t)mr)) ++rk;
            }
            if (mr >= 0 && rk < a.k) {
                a.out_idx[(size_t)b * a.k + rk] = a.seq ? (int64_t)a.seq[mr] : (int64_t)(a.base_row + (uint64_t)mr);
                if (a.out_score) a.out_score[(size_t)b * a.k + rk] = (float)mv;
                if (a.out_score64) a.out_score64[(size_t)b * a.k + rk] = mv;
            }
        }
        cnt = min(ncand, a.k);
    }
    // More candidates than one warp ranks (pure top-k, wide thresholds).  Stage 1: every warp
    // takes a contiguous slice of the rows and extracts ITS k best by k rounds of a warp-wide
    // arg-best (shuffles only, no block barrier).  Stage 2: the union (at most NW * k rows, it
    // contains the global top-k) is ranked by counting, one thread per listed row (one warp per
    // listed row with the lanes striding the union measured no faster).
    if (ncand > kSmallRank) {
        const int slice = (a.rows + NW - 1) / NW;
        const int w_lo = min(warp * slice, a.rows), w_hi = min(w_lo + slice, a.rows);
        for (int r = 0; r < a.k; ++r) {
            double bv = nan;
            int br = -1;
            for (int row = w_lo + lane; row < w_hi; row += 32) {
                const double v = sc[row];
                if (!(v == v)) continue;
                if (last_r >= 0 && !better(a.metric, last_v, (uint32_t)last_r, v, (uint32_t)row)) continue;   // already listed
                if (br < 0 || better(a.metric, v, (uint32_t)row, bv, (uint32_t)br)) { bv = v; br = row; }
            }
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                const double ov = __shfl_xor_sync(FRP_FULL_MASK, bv, sft);
                const int orr = __shfl_xor_sync(FRP_FULL_MASK, br, sft);
                if (orr >= 0 && (br < 0 || better(a.metric, ov, (uint32_t)orr, bv, (uint32_t)br))) { bv = ov; br = orr; }
            }
            if (br < 0) break;                       // warp-uniform after the butterfly
            if (lane == 0) s_union[atomicAdd(&s_nunion, 1)] = br;
            last_v = bv;
            last_r = br;
        }
        __syncthreads();
        const int nu = s_nunion;
        for (int e = tid; e < nu; e += blockDim.x) {
            const int mr = s_union[e];
            const double mv = sc[mr];
            int rk = 0;
            for (int j2 = 0; j2 < nu; ++j2) {
                const int orr = s_union[j2];
                if (better(a.metric, sc[orr], (uint32_t)orr, mv, (uint32_t)mr)) ++rk;
            }
            if (rk < a.k) {
                a.out_idx[(size_t)b * a.k + rk] = a.seq ? (int64_t)a.seq[mr] : (int64_t)(a.base_row + (uint64_t)mr);
                if (a.out_score) a.out_score[(size_t)b * a.k + rk] = (float)mv;
                if (a.out_score64) a.out_score64[(size_t)b * a.k + rk] = mv;
            }
        }
        cnt = min(ncand, a.k);
    }
    for (int i = cnt + tid; i < a.k; i += blockDim.x) {
        a.out_idx[(size_t)b * a.k + i] = -1;
        if (a.out_score) a.out_score[(size_t)b * a.k + i] = __int_as_float(0x7fc00000);
        if (a.out_score64) a.out_score64[(size_t)b * a.k + i] = nan;
    }
    if (tid == 0) {
        a.out_count[b] = cnt;
        if (a.out_status) a.out_status[b] = FRP_STATUS_OK;     // exact by construction
    }
    if (a.done_flag) {
        // every result word of this query was written by this block: barrier, then a system-scope
        // fence (cumulative over the barrier), then count; the last query publishes the sequence
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            const unsigned int nq = (CS == 1) ? gridDim.x : gridDim.y;
            if (atomicAdd(a.done_count, 1u) + 1u == nq) {
                *a.done_count = 0;
                __threadfence_system();
                *a.done_flag = a.done_seq;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// dense scores: one warp per gallery row, fp64 from the fp32 master.
// ---------------------------------------------------------------------------
__global__ void scores_kernel(const float* __restrict__ g32, const uint8_t* __restrict__ alive,
                              const float* __restrict__ qprep, uint64_t rows, int dimp, int B,
                              int metric, double* __restrict__ out, uint64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* gr = reinterpret_cast<const float4*>(g32 + row * dimp);
    const int nvec = dimp >> 2;
    const bool live = alive[row] != 0;
    for (int b = 0; b < B; ++b) {
        const float4* qv = reinterpret_cast<const float4*>(qprep + (size_t)b * dimp);
        double acc = 0.0;
        for (int i = lane; i < nvec; i += 32) {
            const float4 g = __ldg(gr + i);
            const float4 q = __ldg(qv + i);
            if (metric == FRP_METRIC_COS) {
                acc += (double)g.x * (double)q.x + (double)g.y * (double)q.y +
                       (double)g.z * (double)q.z + (double)g.w * (double)q.w;
            } else {
                const double dx = (double)g.x - (double)q.x, dy = (double)g.y - (double)q.y;
                const double dz = (double)g.z - (double)q.z, dw = (double)g.w - (double)q.w;
                acc += dx * dx + dy * dy + dz * dz + dw * dw;
            }
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) {
            double v = (metric == FRP_METRIC_COS) ? acc : sqrt(acc);
            if (!live) v = __longlong_as_double(0x7ff8000000000000ll);   // NaN marks tombstones
            out[(size_t)b * out_stride + row] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// cross-shard merge: one block per query over R*k gathered entries.
// ---------------------------------------------------------------------------
constexpr int kMergeMax = 1024;

// in_* planes of rank r start at base + r * rank_stride (bytes); status may be null.
__global__ void merge_topk_kernel(int metric, int R, int B, int k, const uint8_t* __restrict__ in_idx_b,
                                  const uint8_t* __restrict__ in_score_b,
                                  const uint8_t* __restrict__ in_count_b,
                                  const uint8_t* __restrict__ in_status_b, size_t rank_stride,
                                  int64_t* __restrict__ out_idx, float* __restrict__ out_score,
                                  double* __restrict__ out_score64, int32_t* __restrict__ out_count,
                                  int32_t* __restrict__ out_any_status) {
    __shared__ int64_t s_idx[kMergeMax];
    __shared__ double s_sc[kMergeMax];
    __shared__ int s_n;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int total = R * k;
    for (int e = tid; e < total; e += blockDim.x) {
        const int r = e / k, j = e % k;
        const int32_t* cnt = reinterpret_cast<const int32_t*>(in_count_b + (size_t)r * rank_stride);
        if (j < cnt[b]) {
            const int64_t* idx = reinterpret_cast<const int64_t*>(in_idx_b + (size_t)r * rank_stride);
            const double* sc = reinterpret_cast<const double*>(in_score_b + (size_t)r * rank_stride);
            const int pos = atomicAdd(&s_n, 1);
            s_idx[pos] = idx[(size_t)b * k + j];
            s_sc[pos] = sc[(size_t)b * k + j];
        }
    }
    if (in_status_b && out_any_status && tid < R) {
        const int32_t* stt = reinterpret_cast<const int32_t*>(in_status_b + (size_t)tid * rank_stride);
        if (stt[b] != 0) atomicOr(out_any_status, 1);
    }
    __syncthreads();
    const int n = s_n;
    // same comparator as rerank_kernel (float64 score, then global row): the merged order of R
    // shards is exactly the order one gallery holding all rows would produce
    for (int e = tid; e < n; e += blockDim.x) {
        const double v = s_sc[e];
        const int64_t id = s_idx[e];
        int rank = 0;
        for (int c = 0; c < n; ++c) {
            const double vc = s_sc[c];
            const int64_t ic = s_idx[c];
            const bool bt = (metric == FRP_METRIC_COS) ? (vc > v || (vc == v && ic < id))
                                                       : (vc < v || (vc == v && ic < id));
            if (bt) ++rank;
        }
        if (rank < k) {
            out_idx[(size_t)b * k + rank] = id;
            if (out_score) out_score[(size_t)b * k + rank] = (float)v;
            if (out_score64) out_score64[(size_t)b * k + rank] = v;
        }
    }
    const int cnt = n < k ? n : k;
    for (int i = cnt + tid; i < k; i += blockDim.x) {
        out_idx[(size_t)b * k + i] = -1;
        if (out_score) out_score[(size_t)b * k + i] = __int_as_float(0x7fc00000);
        if (out_score64) out_score64[(size_t)b * k + i] = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (tid == 0) out_count[b] = cnt;
}

// ---------------------------------------------------------------------------
// fused exchange + merge over peer memory (NVLink / NVSwitch), no NCCL call:
//   phase 1  every block copies a slice of this rank's packed record straight into
//            slot [epoch&1][rank] of EVERY rank's exchange buffer (peer stores);
//            the last block to finish publishes flag[epoch&1][rank] = epoch on every
//            rank (st.release.sys after __threadfence_system)
//   phase 2  the blocks wait (ld.acquire.sys) until all R flags of THIS rank's buffer
//            carry the epoch, then merge the R records like merge_topk_kernel.
// Exchange buffer of a rank: [0,256) u32 flag[2][32]; then recv[2][R][rec_bytes].
// Double buffering by epoch parity is sufficient: a rank can only publish epoch e+1
// after its own merge of epoch e has completed (stream order), so slot parity e&1 is
// never overwritten while a peer still reads it.
// ---------------------------------------------------------------------------
constexpr int kExchHeader = 256;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(128)
push_merge_kernel(int metric, int R, int rank, int B, int k, const uint8_t* __restrict__ send,
                  uint8_t* const* __restrict__ peers, size_t rec_bytes, uint32_t epoch,
                  unsigned int* __restrict__ block_counter, int64_t* __restrict__ out_idx,
                  float* __restrict__ out_score, int32_t* __restrict__ out_count,
                  int32_t* __restrict__ out_any_status) {
    __shared__ int64_t s_idx[kMergeMax];
    __shared__ double s_sc[kMergeMax];
    __shared__ int s_n;
    __shared__ bool s_last;
    const int tid = threadIdx.x;
    const uint32_t par = epoch & 1u;

    // ---- phase 1: push (a few blocks only; ONE system-scope fence per pushing block: the
    // block barrier orders the other threads' peer stores before thread 0's fence) ----
    const unsigned npush = min((unsigned)gridDim.x, (unsigned)(2 * R));
    if (blockIdx.x < npush) {
        const uint4* src = reinterpret_cast<const uint4*>(send);
        const size_t nvec = rec_bytes / 16;
        const size_t total = nvec * (size_t)R;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < total; i += (size_t)npush * blockDim.x) {
            const int r = (int)(i / nvec);
            const size_t v = i % nvec;
            uint4* dst = reinterpret_cast<uint4*>(peers[r] + kExchHeader + ((size_t)par * R + rank) * rec_bytes);
            dst[v] = src[v];
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            s_last = (atomicAdd(block_counter, 1u) == npush - 1);
            if (s_last) {
                __threadfence_system();
                for (int r = 0; r < R; ++r)
                    st_release_sys(reinterpret_cast<uint32_t*>(peers[r]) + par * 32 + rank, epoch);
                *block_counter = 0u;           // ready for the next launch
            }
        }
    }

    // ---- phase 2: wait for every rank's record in this rank's buffer ----
    const uint8_t* mine = peers[rank];
    if (tid < R) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(mine) + par * 32 + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != epoch) {
            if (clock64() - t0 > 6000000000ll) {      // ~3 s: a peer died; flag the result, do not hang
                if (out_any_status) atomicOr(out_any_status, 0x100);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();

    // ---- merge ----
    const size_t o1 = (size_t)B * k * 8, o2 = o1 + (size_t)B * k * 8, o3 = o2 + (size_t)B * 4;
    const uint8_t* recv = mine + kExchHeader + (size_t)par * R * rec_bytes;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        const int total = R * k;
        for (int e = tid; e < total; e += blockDim.x) {
            const int r = e / k, j = e % k;
            const uint8_t* rec = recv + (size_t)r * rec_bytes;
            const int32_t* cnt = reinterpret_cast<const int32_t*>(rec + o2);
            if (j < cnt[b]) {
                const int pos = atomicAdd(&s_n, 1);
                s_idx[pos] = reinterpret_cast<const int64_t*>(rec)[(size_t)b * k + j];
                s_sc[pos] = reinterpret_cast<const double*>(rec + o1)[(size_t)b * k + j];
            }
        }
        if (out_any_status && tid < R) {
            const int32_t* stt = reinterpret_cast<const int32_t*>(recv + (size_t)tid * rec_bytes + o3);
            if (stt[b] != 0) atomicOr(out_any_status, 1);
        }
        __syncthreads();
        const int n = s_n;
        for (int e = tid; e < n; e += blockDim.x) {
            const double v = s_sc[e];
            const int64_t id = s_idx[e];
            int rk = 0;
            for (int c = 0; c < n; ++c) {
                const double vc = s_sc[c];
                const int64_t ic = s_idx[c];
                const bool bt = (metric == FRP_METRIC_COS) ? (vc > v || (vc == v && ic < id))
                                                           : (vc < v || (vc == v && ic < id));
                if (bt) ++rk;
            }
            if (rk < k) {
                out_idx[(size_t)b * k + rk] = id;
                out_score[(size_t)b * k + rk] = (float)v;
            }
        }
        const int cntb = n < k ? n : k;
        for (int i = cntb + tid; i < k; i += blockDim.x) {
            out_idx[(size_t)b * k + i] = -1;
            out_score[(size_t)b * k + i] = __int_as_float(0x7fc00000);
        }
        if (tid == 0) out_count[b] = cntb;
        __syncthreads();
    }
}

}  // namespace

int launch_prep_queries(const frp_gallery* g, const float* q_dev, size_t qstride, const uint64_t* q_rows,
                        int B, double tau, int flags, int eps_mode, float* qprep, __half* q16,
                        QParams* qp, cudaStream_t st) {
    configure_func_carveout((const void*)prep_queries_kernel, g->device);
    prep_queries_kernel<<<B, 128, 0, st>>>(q_dev, qstride, q_rows, g->dim, g->dimp, g->metric, tau, flags,
                                           eps_mode, g->gmax, g->gerr, qprep, q16, qp);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

bool match_small_supported(const frp_gallery* g, int B) { return g->rows > 0 && g->rows <= (uint64_t)kSmallRows && B <= 64; }

// FRP_SMALL_CLUSTER=0: one CTA per query instead of a cluster of kSmallCluster (measurement aid)
static const bool g_small_cluster = []() {
    const char* e = getenv("FRP_SMALL_CLUSTER");
    return !(e && e[0] == '0');
}();

int launch_match_small(const frp_gallery* g, const float* q_dev, size_t qstride, int B, double tau, int k, int flags,
                       int64_t* out_idx, float* out_score, double* out_score64, int32_t* out_count,
                       int32_t* out_status, cudaStream_t st, unsigned int* done_count,
                       volatile unsigned int* done_flag, unsigned int done_seq, const float* q_host) {
    SmallArgs a;
    a.g32 = g->g32; a.alive = g->alive; a.seq = g->seq; a.q = q_dev; a.qstride = qstride;
    a.rows = (int)g->rows; a.dim = g->dim; a.dimp = g->dimp; a.metric = g->metric; a.flags = flags; a.k = k;
    a.tau = tau; a.base_row = g->base_row;
    a.out_idx = out_idx; a.out_score = out_score; a.out_score64 = out_score64; a.out_count = out_count;
    a.out_status = out_status;
    a.done_count = done_count; a.done_flag = done_count ? done_flag : nullptr; a.done_seq = done_seq;
    a.inline_q = 0;
    if (q_host && (size_t)B * g->dim <= (size_t)kSmallInline && qstride == (size_t)g->dim) {
        memcpy(a.qinline, q_host, (size_t)B * g->dim * sizeof(float));
        a.inline_q = 1;
    }
    const size_t smem = (((size_t)a.rows * 8 + 15) & ~(size_t)15) + (size_t)g->dimp * 4;
    // a handful of rows per CTA is not worth the cluster barrier
    if (g_small_cluster && a.rows >= 64) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kSmallCluster, (unsigned)B, 1);
        cfg.blockDim = dim3(small_threads(kSmallCluster), 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kSmallCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, match_small_kernel<kSmallCluster>, a);
        if (e == cudaSuccess) {
            count_launch();
            return FRP_OK;
        }
        // a device partition that cannot co-schedule 8 CTAs refuses the cluster launch (not a
        // sticky error): the one-CTA form of the same kernel below computes the same values
        cudaGetLastError();
    }
    match_small_kernel<1><<<B, small_threads(1), smem, st>>>(a);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

int launch_raise_cut(QParams* qp, int B, int k, const uint64_t* cand, int nlists, int KP, cudaStream_t st) {
    {
        int dev = 0;
        cudaGetDevice(&dev);
        configure_func_carveout((const void*)raise_cut_kernel, dev);
    }
    raise_cut_kernel<<<B, 256, 0, st>>>(qp, k, cand, nlists * KP);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

int launch_scores(const frp_gallery* g, const float* qprep, int B, double* out, uint64_t out_stride,
                  cudaStream_t st) {
    if (g->rows == 0) return FRP_OK;
    const int wpb = 8;
    const uint64_t blocks = (g->rows + wpb - 1) / wpb;
    scores_kernel<<<(unsigned)blocks, wpb * 32, 0, st>>>(g->g32, g->alive, qprep, g->rows, g->dimp, B,
                                                         g->metric, out, out_stride);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

// tau travels separately from QParams so the exact comparison uses the caller's double
int launch_rerank(const frp_gallery* g, const float* qprep, const QParams* qp, int B, int k, int KP,
                  int flags, double tau, const uint64_t* cand, int nlists, int64_t* out_idx,
                  float* out_score, double* out_score64, int32_t* out_count, int32_t* out_status,
                  cudaStream_t st) {
    RerankArgs a;
    a.g32 = g->g32;
    a.qprep = qprep;
    a.qp = qp;
    a.cand = cand;
    a.nlists = nlists;
    a.KP = KP;
    a.k = k;
    a.dimp = g->dimp;
    a.metric = g->metric;
    a.flags = flags;
    a.tau = tau;
    a.base_row = g->base_row;
    a.seq = g->seq;
    a.out_idx = out_idx;
    a.out_score = out_score;
    a.out_score64 = out_score64;
    a.out_count = out_count;
    a.out_status = out_status;
    int npad = 1;
    while (npad < nlists * KP) npad <<= 1;
    a.npad = npad;
    const size_t smem = (size_t)npad * sizeof(uint64_t);
    if (smem > 160 * 1024) {
        set_error("rerank: %d candidate slots exceed shared memory", nlists * KP);
        return FRP_E_UNSUPPORTED;
    }
    if (smem > 32 * 1024) {
        int rc = FRP_OK;
        configure_func_smem((const void*)rerank_kernel, g->device, smem, &rc);
        if (rc) return rc;
    }
    configure_func_carveout((const void*)rerank_kernel, g->device);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)B);
        cfg.blockDim = dim3(kRerankThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, rerank_kernel, a);
        if (le != cudaSuccess) {
            set_error("rerank launch failed: %s", cudaGetErrorString(le));
            cudaGetLastError();
            return FRP_E_CUDA;
        }
    }
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

}  // namespace frp

extern "C" {

size_t frp_exchange_buffer_bytes(int R, int B, int k) {
    return (size_t)frp::kExchHeader + 2 * (size_t)R * frp_shard_record_bytes(B, k);
}

int frp_push_merge_device(int metric, int R, int rank, int B, int k, const void* send_record,
                          void* const* peers_dev, uint32_t epoch, unsigned int* counter_dev,
                          int64_t* out_idx, float* out_score, int32_t* out_count,
                          int32_t* out_any_status, void* stream) {
    using namespace frp;
    if (R <= 0 || R > 32 || rank < 0 || rank >= R || B <= 0 || k <= 0 || R * k > kMergeMax || !send_record ||
        !peers_dev || !counter_dev || !out_idx || !out_score || !out_count) {
        set_error("frp_push_merge_device: invalid argument (R=%d rank=%d B=%d k=%d)", R, rank, B, k);
        return FRP_E_INVALID;
    }
    const size_t rec = frp_shard_record_bytes(B, k);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_any_status) FRP_CUDA(cudaMemsetAsync(out_any_status, 0, sizeof(int32_t), st));
    // Blocks spin on flags written by OTHER GPUs; the local pushes they depend on are done by the
    // lowest-numbered blocks, which never wait before pushing.  The grid is kept below one
    // block per SM (148) so every block is resident even when another stream (alignment) holds
    // part of each SM, and a peer that never arrives ends in the 3 s timeout flag, not a hang.
    int grid = B < 128 ? B : 128;
    push_merge_kernel<<<grid, 128, 0, st>>>(metric, R, rank, B, k, (const uint8_t*)send_record,
                                            (uint8_t* const*)peers_dev, rec, epoch, counter_dev, out_idx,
                                            out_score, out_count, out_any_status);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

size_t frp_shard_record_bytes(int B, int k) {
    // idx i64 | score f64 | count i32 | status i32
    size_t n = (size_t)B * k * 8 + (size_t)B * k * 8 + (size_t)B * 4 + (size_t)B * 4;
    return (n + 15) / 16 * 16;
}

int frp_merge_packed_device(int metric, int R, int B, int k, const void* records, int64_t* out_idx,
                            float* out_score, double* out_score64, int32_t* out_count,
                            int32_t* out_any_status, void* stream) {
    using namespace frp;
    if (R <= 0 || B <= 0 || k <= 0 || R * k > kMergeMax || R > 128 || !records || !out_idx ||
        (!out_score && !out_score64) || !out_count) {
        set_error("frp_merge_packed_device: invalid argument (R=%d B=%d k=%d, R*k <= %d)", R, B, k, kMergeMax);
        return FRP_E_INVALID;
    }
    const size_t rec = frp_shard_record_bytes(B, k);
    const uint8_t* base = (const uint8_t*)records;
    const size_t o1 = (size_t)B * k * 8, o2 = o1 + (size_t)B * k * 8, o3 = o2 + (size_t)B * 4;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_any_status) FRP_CUDA(cudaMemsetAsync(out_any_status, 0, sizeof(int32_t), st));
    {
        int dev = 0;
        cudaGetDevice(&dev);
        configure_func_carveout((const void*)merge_topk_kernel, dev);
    }
    merge_topk_kernel<<<B, 128, 0, st>>>(metric, R, B, k, base, base + o1, base + o2, base + o3, rec,
                                         out_idx, out_score, out_score64, out_count, out_any_status);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

}  // extern "C"
