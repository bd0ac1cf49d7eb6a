// HBM-resident watchlist gallery.  Replaces state.ENCODINGS (backend/app/state.py:78)
// and the per-call `np.array([ENCODINGS[t] ...])` rebuild (face_service.py:409).
//
// Layout per gallery (row = insertion sequence = tie order):
//   g16   [cap][dimp] fp16   streaming scan copy (COS: L2-normalised rows)
//   g32   [cap][dimp] fp32   master for the exact re-rank (COS: normalised)
//   norm2 [cap]       fp32   |g|^2 of the master row (L2 epilogue term)
//   alive [cap]       u8     tombstones (del ENCODINGS[t], face_service.py:521-523)
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "frp_device.cuh"
#include "frp_internal.h"

namespace frp {

static thread_local std::string t_error;
static thread_local const char* t_match_path = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}
void set_match_path(const char* name) { t_match_path = name; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: the
// opt-in is remembered per (kernel, device), not per thread.
void configure_func_smem(const void* func, int device, size_t bytes, int* rc) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = done[std::make_pair(func, device)];
    if (cur >= bytes) return;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(%zu bytes of dynamic shared memory) failed: %s", bytes,
                  cudaGetErrorString(e));
        cudaGetLastError();
        if (rc) *rc = FRP_E_CUDA;
        return;
    }
    cur = bytes;
}

// The small kernels around the prefilter (query prep, re-rank, merge) ask for the same
// shared-memory carve-out as the prefilter itself (all of it): when consecutive kernels on a
// stream want different L1 / shared splits the SMs drain and reconfigure between them, which
// showed up as tens of microseconds per step that no kernel accounted for.
void configure_func_carveout(const void* func, int device) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, bool> done;
    std::lock_guard<std::mutex> lk(mu);
    bool& d = done[std::make_pair(func, device)];
    if (d) return;
    cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    d = true;
}

// One warp per row: fp64 norm, optional normalisation, fp32 master + fp16 copy.
// src rows are `src_stride` floats apart; normalise = 0 ingests COS rows verbatim (snapshot load:
// the stored master is already normalised and must stay bit-identical).
__global__ void ingest_rows_kernel(const float* __restrict__ src, uint64_t n, int dim, int dimp,
                                   uint64_t src_stride, int normalise, uint64_t first_row,
                                   __half* __restrict__ g16,
                                   float* __restrict__ g32, float* __restrict__ norm2,
                                   uint8_t* __restrict__ alive, float* __restrict__ gmax) {
    const int lane = threadIdx.x & 31;
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n) return;
    const float* s = src + w * src_stride;
    double acc = 0.0;
    for (int i = lane; i < dim; i += 32) {
        const double x = (double)s[i];
        acc += x * x;
    }
    acc = warp_sum_f64(acc);
    double scale = 1.0;
    if (normalise) scale = acc > 0.0 ? 1.0 / sqrt(acc) : 0.0;
    const uint64_t row = first_row + w;
    double acc2 = 0.0, err2 = 0.0;
    for (int i = lane; i < dimp; i += 32) {
        float v = 0.f;
        if (i < dim) v = (float)((double)s[i] * scale);
        const __half h = __float2half_rn(v);
        g32[row * dimp + i] = v;
        g16[row * dimp + i] = h;
        acc2 += (double)v * (double)v;
        const double d = (double)v - (double)__half2float(h);      // inf - x = inf on fp16 overflow
        err2 += d * d;
    }
    acc2 = warp_sum_f64(acc2);
    err2 = warp_sum_f64(err2);
    if (lane == 0) {
        norm2[row] = (float)acc2;
        alive[row] = 1;
        // gmax[0]: monotone upper bound of |g|; gmax[1]: of |g - fp16(g)|, the ACTUAL rounding
        // error of the scan copy (rigorous prefilter bound by Cauchy-Schwarz, any scale).
        // Rounded up; non-negative floats order like ints.
        const float nr = __double2float_ru(sqrt(acc2));
        float er = __double2float_ru(sqrt(err2));
        if (!(er == er)) er = __int_as_float(0x7f800000);
        atomicMax((int*)gmax, __float_as_int(nr));
        atomicMax((int*)gmax + 1, __float_as_int(er));
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ src32, const __half* __restrict__ src16,
                                   const float* __restrict__ srcn, const uint64_t* __restrict__ idx,
                                   uint64_t n, int dimp, float* __restrict__ dst32,
                                   __half* __restrict__ dst16, float* __restrict__ dstn,
                                   uint8_t* __restrict__ dalive, const uint64_t* __restrict__ srcseq,
                                   uint64_t* __restrict__ dstseq) {
    const int lane = threadIdx.x & 31;
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n) return;
    const uint64_t r = idx[w];
    for (int i = lane; i < dimp; i += 32) {
        dst32[w * dimp + i] = src32[r * dimp + i];
        dst16[w * dimp + i] = src16[r * dimp + i];
    }
    if (lane == 0) {
        dstn[w] = srcn[r];
        dalive[w] = 1;
        if (srcseq) dstseq[w] = srcseq[r];
    }
}

static int pad_dim(int dim) {
    int p = 64;
    while (p < dim) p <<= 1;
    return p;
}

static void free_buffers(frp_gallery* g) {
    if (g->g16) cudaFree(g->g16);
    if (g->g32) cudaFree(g->g32);
    if (g->norm2) cudaFree(g->norm2);
    if (g->alive) cudaFree(g->alive);
    if (g->seq) cudaFree(g->seq);
    g->g16 = nullptr; g->g32 = nullptr; g->norm2 = nullptr; g->alive = nullptr; g->seq = nullptr;
    g->tmap_valid = false;
}

static int alloc_buffers(int dimp, uint64_t cap, __half** g16, float** g32, float** norm2,
                         uint8_t** alive) {
    *g16 = nullptr; *g32 = nullptr; *norm2 = nullptr; *alive = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc((void**)g16, cap * dimp * sizeof(__half))) != cudaSuccess ||
        (e = cudaMalloc((void**)g32, cap * dimp * sizeof(float))) != cudaSuccess ||
        // one tile (256 rows) of slack: the tcgen05 matcher bulk-copies whole 256-row norm tiles
        (e = cudaMalloc((void**)norm2, (cap + 256) * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc((void**)alive, cap)) != cudaSuccess) {
        set_error("gallery allocation of %llu rows x %d failed: %s", (unsigned long long)cap, dimp,
                  cudaGetErrorString(e));
        if (*g16) cudaFree(*g16);
        if (*g32) cudaFree(*g32);
        if (*norm2) cudaFree(*norm2);
        if (*alive) cudaFree(*alive);
        cudaGetLastError();
        return FRP_E_NOMEM;
    }
    // rows beyond `rows` are streamed by the last tcgen05 tile (its TMA box is bounded by the
    // capacity, not by the row count): keep them finite
    if (cudaMemset(*g16, 0, cap * dimp * sizeof(__half)) != cudaSuccess ||
        cudaMemset(*norm2, 0, (cap + 256) * sizeof(float)) != cudaSuccess) {
        set_error("gallery clear failed");
        cudaFree(*g16); cudaFree(*g32); cudaFree(*norm2); cudaFree(*alive);
        *g16 = nullptr; *g32 = nullptr; *norm2 = nullptr; *alive = nullptr;
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    return FRP_OK;
}

// caller holds the unique lock
static int ensure_capacity(frp_gallery* g, uint64_t need) {
    if (need <= g->capacity) return FRP_OK;
    uint64_t cap = g->capacity ? g->capacity : 1024;
    while (cap < need) cap = cap + cap / 2 + 1024;
    __half* n16; float* n32; float* nn; uint8_t* na;
    int rc = alloc_buffers(g->dimp, cap, &n16, &n32, &nn, &na);
    if (rc) return rc;
    uint64_t* nseq = nullptr;
    if (g->seq) {
        if (cudaMalloc((void**)&nseq, cap * sizeof(uint64_t)) != cudaSuccess) {
            cudaFree(n16); cudaFree(n32); cudaFree(nn); cudaFree(na);
            set_error("gallery sequence allocation failed");
            cudaGetLastError();
            return FRP_E_NOMEM;
        }
        if (g->rows) FRP_CUDA(cudaMemcpy(nseq, g->seq, g->rows * sizeof(uint64_t), cudaMemcpyDeviceToDevice));
    }
    if (g->rows) {
        FRP_CUDA(cudaMemcpy(n16, g->g16, g->rows * g->dimp * sizeof(__half), cudaMemcpyDeviceToDevice));
        FRP_CUDA(cudaMemcpy(n32, g->g32, g->rows * g->dimp * sizeof(float), cudaMemcpyDeviceToDevice));
        FRP_CUDA(cudaMemcpy(nn, g->norm2, g->rows * sizeof(float), cudaMemcpyDeviceToDevice));
        FRP_CUDA(cudaMemcpy(na, g->alive, g->rows, cudaMemcpyDeviceToDevice));
    }
    free_buffers(g);
    g->g16 = n16; g->g32 = n32; g->norm2 = nn; g->alive = na; g->seq = nseq;
    g->capacity = cap;
    gemm_refresh_maps(g);
    return FRP_OK;
}

// caller holds the unique lock; src is device fp32 [n][dim]
static int ingest_device(frp_gallery* g, const float* src, uint64_t n, uint64_t first_row,
                         cudaStream_t st, uint64_t src_stride = 0, bool verbatim = false) {
    if (n == 0) return FRP_OK;
    const int wpb = 8;
    const uint64_t blocks = (n + wpb - 1) / wpb;
    const int normalise = (g->metric == FRP_METRIC_COS && !verbatim) ? 1 : 0;
    ingest_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, st>>>(src, n, g->dim, g->dimp,
                                                              src_stride ? src_stride : (uint64_t)g->dim,
                                                              normalise, first_row, g->g16, g->g32,
                                                              g->norm2, g->alive, g->d_gmax);
    FRP_CHECK_LAUNCH();
    return FRP_OK;
}

static int refresh_gmax(frp_gallery* g, cudaStream_t st) {
    // pinned landing zone behind the staging row: the copy is truly asynchronous and only
    // `st` is waited for
    float* both = g->h_stage + g->dim;
    FRP_CUDA(cudaMemcpyAsync(both, g->d_gmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    FRP_CUDA(cudaStreamSynchronize(st));
    g->gmax = both[0];
    g->gerr = both[1];
    return FRP_OK;
}

static int append_host_f32(frp_gallery* g, const float* rows, uint64_t n, uint64_t* first_row) {
    std::unique_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    int rc = ensure_capacity(g, g->rows + n);
    if (rc) return rc;
    const uint64_t first = g->rows;
    const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / (sizeof(float) * g->dim));
    float* stage = nullptr;
    if (n) FRP_CUDA(cudaMalloc((void**)&stage, std::min(n, chunk) * g->dim * sizeof(float)));
    for (uint64_t lo = 0; lo < n; lo += chunk) {
        const uint64_t m = std::min(chunk, n - lo);
        cudaError_t e = cudaMemcpy(stage, rows + lo * g->dim, m * g->dim * sizeof(float),
                                   cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(stage);
            set_error("H2D copy of gallery rows failed: %s", cudaGetErrorString(e));
            return FRP_E_CUDA;
        }
        rc = ingest_device(g, stage, m, first + lo, 0);
        if (rc) { cudaFree(stage); return rc; }
        FRP_CUDA(cudaStreamSynchronize(0));
    }
    if (stage) cudaFree(stage);
    g->rows += n;
    g->h_alive.resize((size_t)g->rows, 1);
    if (first_row) *first_row = first;
    return refresh_gmax(g, 0);
}

// In-place update of one live row: pinned staging row -> one 4*dim-byte H2D, the ingest kernel
// and the 8-byte bound read-back, all on the gallery's private stream.  No allocation, no
// device-wide synchronisation; matchers that hold the shared lock have returned (host API) or
// are ordered by their caller (device API) before the unique lock is granted.
template <typename T>
static int update_row(frp_gallery* g, uint64_t row, const T* vec) {
    std::unique_lock<std::shared_mutex> lk(g->mu);
    if (row >= g->rows) { set_error("row %llu out of range", (unsigned long long)row); return FRP_E_RANGE; }
    if (!g->h_alive[(size_t)row]) { set_error("row %llu is deleted", (unsigned long long)row); return FRP_E_RANGE; }
    DeviceGuard dg(g->device);
    for (int i = 0; i < g->dim; ++i) g->h_stage[i] = (float)vec[i];
    FRP_CUDA(cudaMemcpyAsync(g->d_stage, g->h_stage, (size_t)g->dim * sizeof(float), cudaMemcpyHostToDevice,
                             g->mut_stream));
    int rc = ingest_device(g, g->d_stage, 1, row, g->mut_stream);
    if (rc) return rc;
    return refresh_gmax(g, g->mut_stream);
}

}  // namespace frp

using namespace frp;

extern "C" {

int frp_abi_version(void) { return FRP_ABI_VERSION; }
const char* frp_last_error(void) { return t_error.c_str(); }
uint64_t frp_kernel_launches(void) { return g_launches.load(); }
const char* frp_last_match_path(void) { return t_match_path; }

int frp_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return FRP_E_NODEVICE;
    }
    return FRP_OK;
}

int frp_device_info(int device, char* name, size_t name_len, int* sm_count, uint64_t* mem_bytes,
                    int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    FRP_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_len) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (mem_bytes) *mem_bytes = (uint64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return FRP_OK;
}

int frp_gallery_create(int device, int dim, int metric, uint64_t capacity, frp_gallery_t** out) {
    if (!out || dim <= 0 || dim > 512 || (metric != FRP_METRIC_L2 && metric != FRP_METRIC_COS)) {
        set_error("frp_gallery_create: invalid argument (dim=%d metric=%d)", dim, metric);
        return FRP_E_INVALID;
    }
    int ndev = 0;
    int rc = frp_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) {
        set_error("frp_gallery_create: device %d out of range (%d visible)", device, ndev);
        return FRP_E_INVALID;
    }
    cudaDeviceProp p;
    FRP_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", device,
                  p.major, p.minor);
        return FRP_E_NODEVICE;
    }
    DeviceGuard dg(device);
    frp_gallery* g = new frp_gallery();
    g->device = device;
    g->dim = dim;
    g->dimp = pad_dim(dim);
    g->metric = metric;
    g->sm_count = p.multiProcessorCount;
    if (capacity < 1024) capacity = 1024;
    rc = alloc_buffers(g->dimp, capacity, &g->g16, &g->g32, &g->norm2, &g->alive);
    if (rc) { delete g; return rc; }
    g->capacity = capacity;
    cudaError_t e = cudaMalloc((void**)&g->d_gmax, 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(g->d_gmax, 0, 2 * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->mut_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&g->h_stage, (size_t)(dim + 2) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_stage, (size_t)dim * sizeof(float));
    if (e != cudaSuccess) {
        set_error("gallery scalar allocation failed: %s", cudaGetErrorString(e));
        free_buffers(g);
        if (g->d_gmax) cudaFree(g->d_gmax);
        if (g->mut_stream) cudaStreamDestroy(g->mut_stream);
        if (g->h_stage) cudaFreeHost(g->h_stage);
        if (g->d_stage) cudaFree(g->d_stage);
        delete g;
        cudaGetLastError();
        return FRP_E_CUDA;
    }
    gemm_refresh_maps(g);
    *out = g;
    return FRP_OK;
}

int frp_gallery_destroy(frp_gallery_t* g) {
    if (!g) return FRP_OK;
    {
        // wait for every matcher that still holds the shared lock, then for the device work
        // they queued; the caller must not use the handle afterwards
        std::unique_lock<std::shared_mutex> lk(g->mu);
        DeviceGuard dg(g->device);
        cudaDeviceSynchronize();
        free_buffers(g);
        if (g->d_gmax) cudaFree(g->d_gmax);
        if (g->d_stage) cudaFree(g->d_stage);
        if (g->h_stage) cudaFreeHost(g->h_stage);
        if (g->mut_stream) cudaStreamDestroy(g->mut_stream);
        g->d_gmax = nullptr; g->d_stage = nullptr; g->h_stage = nullptr; g->mut_stream = nullptr;
        cudaGetLastError();
    }
    delete g;
    return FRP_OK;
}

int frp_gallery_append_f32(frp_gallery_t* g, const float* rows, uint64_t n, uint64_t* first_row) {
    if (!g || (!rows && n)) { set_error("frp_gallery_append_f32: null argument"); return FRP_E_INVALID; }
    return append_host_f32(g, rows, n, first_row);
}

int frp_gallery_append_f64(frp_gallery_t* g, const double* rows, uint64_t n, uint64_t* first_row) {
    if (!g || (!rows && n)) { set_error("frp_gallery_append_f64: null argument"); return FRP_E_INVALID; }
    // the reference keeps Python floats (f64, face_service.py:374); the HBM master is fp32
    std::vector<float> tmp((size_t)n * g->dim);
    for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = (float)rows[i];
    return append_host_f32(g, tmp.data(), n, first_row);
}

int frp_gallery_append_device(frp_gallery_t* g, const float* rows_dev, uint64_t n, uint64_t* first_row,
                              void* stream) {
    if (!g || (!rows_dev && n)) { set_error("frp_gallery_append_device: null argument"); return FRP_E_INVALID; }
    std::unique_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (g->rows + n > g->capacity) {
        FRP_CUDA(cudaStreamSynchronize(st));
        int rc = ensure_capacity(g, g->rows + n);
        if (rc) return rc;
    }
    const uint64_t first = g->rows;
    int rc = ingest_device(g, rows_dev, n, first, st);
    if (rc) return rc;
    g->rows += n;
    g->h_alive.resize((size_t)g->rows, 1);
    if (first_row) *first_row = first;
    return refresh_gmax(g, st);
}

int frp_gallery_update_f32(frp_gallery_t* g, uint64_t row, const float* vec) {
    if (!g || !vec) { set_error("frp_gallery_update_f32: null argument"); return FRP_E_INVALID; }
    return update_row(g, row, vec);
}
int frp_gallery_update_f64(frp_gallery_t* g, uint64_t row, const double* vec) {
    if (!g || !vec) { set_error("frp_gallery_update_f64: null argument"); return FRP_E_INVALID; }
    return update_row(g, row, vec);
}

int frp_gallery_remove(frp_gallery_t* g, uint64_t row) {
    if (!g) { set_error("frp_gallery_remove: null gallery"); return FRP_E_INVALID; }
    std::unique_lock<std::shared_mutex> lk(g->mu);
    if (row >= g->rows) { set_error("row %llu out of range", (unsigned long long)row); return FRP_E_RANGE; }
    if (!g->h_alive[(size_t)row]) { set_error("row %llu already deleted", (unsigned long long)row); return FRP_E_RANGE; }
    DeviceGuard dg(g->device);
    FRP_CUDA(cudaMemsetAsync(g->alive + row, 0, 1, g->mut_stream));
    FRP_CUDA(cudaStreamSynchronize(g->mut_stream));
    g->h_alive[(size_t)row] = 0;
    g->dead += 1;
    return FRP_OK;
}

int frp_gallery_count(const frp_gallery_t* g, uint64_t* rows, uint64_t* live) {
    if (!g) { set_error("frp_gallery_count: null gallery"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    if (rows) *rows = g->rows;
    if (live) *live = g->rows - g->dead;
    return FRP_OK;
}

int frp_gallery_info(const frp_gallery_t* g, int* device, int* dim, int* dimp, int* metric,
                     uint64_t* capacity) {
    if (!g) { set_error("frp_gallery_info: null gallery"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    if (device) *device = g->device;
    if (dim) *dim = g->dim;
    if (dimp) *dimp = g->dimp;
    if (metric) *metric = g->metric;
    if (capacity) *capacity = g->capacity;
    return FRP_OK;
}

int frp_gallery_set_base(frp_gallery_t* g, uint64_t base_row) {
    if (!g) { set_error("frp_gallery_set_base: null gallery"); return FRP_E_INVALID; }
    std::unique_lock<std::shared_mutex> lk(g->mu);
    g->base_row = base_row;
    return FRP_OK;
}

int frp_gallery_set_seq(frp_gallery_t* g, uint64_t first_row, uint64_t n, const uint64_t* seq) {
    if (!g || (!seq && n)) { set_error("frp_gallery_set_seq: null argument"); return FRP_E_INVALID; }
    std::unique_lock<std::shared_mutex> lk(g->mu);
    if (first_row + n > g->rows) { set_error("frp_gallery_set_seq: rows [%llu, %llu) out of range", (unsigned long long)first_row, (unsigned long long)(first_row + n)); return FRP_E_RANGE; }
    DeviceGuard dg(g->device);
    if (!g->seq) {
        // first use: every existing row gets its default sequence base_row + row
        FRP_CUDA(cudaMalloc((void**)&g->seq, g->capacity * sizeof(uint64_t)));
        std::vector<uint64_t> init((size_t)g->rows);
        for (uint64_t r = 0; r < g->rows; ++r) init[(size_t)r] = g->base_row + r;
        if (g->rows) FRP_CUDA(cudaMemcpy(g->seq, init.data(), g->rows * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    if (n) FRP_CUDA(cudaMemcpy(g->seq + first_row, seq, n * sizeof(uint64_t), cudaMemcpyHostToDevice));
    return FRP_OK;
}

int frp_gallery_gather(const frp_gallery_t* g, const uint64_t* idx, uint64_t n, frp_gallery_t** out) {
    if (!g || !out || (!idx && n)) { set_error("frp_gallery_gather: null argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    for (uint64_t i = 0; i < n; ++i)
        if (idx[i] >= g->rows) { set_error("gather index %llu out of range", (unsigned long long)idx[i]); return FRP_E_RANGE; }
    frp_gallery_t* ng = nullptr;
    int rc = frp_gallery_create(g->device, g->dim, g->metric, n, &ng);
    if (rc) return rc;
    DeviceGuard dg(g->device);
    if (g->seq) {
        if (cudaMalloc((void**)&ng->seq, ng->capacity * sizeof(uint64_t)) != cudaSuccess) {
            set_error("gather: sequence allocation failed");
            frp_gallery_destroy(ng);
            cudaGetLastError();
            return FRP_E_NOMEM;
        }
    }
    ng->base_row = g->base_row;
    if (n) {
        uint64_t* didx = nullptr;
        FRP_CUDA(cudaMalloc((void**)&didx, n * sizeof(uint64_t)));
        cudaError_t e = cudaMemcpy(didx, idx, n * sizeof(uint64_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            const int wpb = 8;
            gather_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32>>>(
                g->g32, g->g16, g->norm2, didx, n, g->dimp, ng->g32, ng->g16, ng->norm2, ng->alive, g->seq,
                ng->seq);
            e = cudaGetLastError();
            count_launch();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
        }
        cudaFree(didx);
        if (e != cudaSuccess) {
            set_error("gather failed: %s", cudaGetErrorString(e));
            frp_gallery_destroy(ng);
            return FRP_E_CUDA;
        }
    }
    ng->rows = n;
    ng->h_alive.assign((size_t)n, 1);
    ng->gmax = g->gmax;
    ng->gerr = g->gerr;
    const float both[2] = {ng->gmax, ng->gerr};
    cudaMemcpy(ng->d_gmax, both, 2 * sizeof(float), cudaMemcpyHostToDevice);
    *out = ng;
    return FRP_OK;
}

int frp_gallery_get_row(const frp_gallery_t* g, uint64_t row, float* out) {
    if (!g || !out) { set_error("frp_gallery_get_row: null argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    if (row >= g->rows) { set_error("row %llu out of range", (unsigned long long)row); return FRP_E_RANGE; }
    DeviceGuard dg(g->device);
    FRP_CUDA(cudaMemcpy2D(out, g->dim * sizeof(float), g->g32 + row * g->dimp, g->dimp * sizeof(float),
                          g->dim * sizeof(float), 1, cudaMemcpyDeviceToHost));
    return FRP_OK;
}

struct SnapshotHeader {
    char magic[8];  // "FRPGAL1\0"
    int32_t dim, dimp, metric, reserved;
    uint64_t rows, dead, base_row;
    float gmax;
    float pad;
};

int frp_gallery_save(const frp_gallery_t* g, const char* path) {
    if (!g || !path) { set_error("frp_gallery_save: null argument"); return FRP_E_INVALID; }
    std::shared_lock<std::shared_mutex> lk(g->mu);
    DeviceGuard dg(g->device);
    // written next to the target and renamed over it: a crash mid-save never leaves a
    // truncated snapshot under the final name
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("cannot open %s for writing", tmp.c_str()); return FRP_E_IO; }
    SnapshotHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "FRPGAL1", 8);
    h.dim = g->dim; h.dimp = g->dimp; h.metric = g->metric;
    h.rows = g->rows; h.dead = g->dead; h.base_row = g->base_row; h.gmax = g->gmax;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / (sizeof(float) * g->dimp));
    std::vector<float> buf((size_t)std::min<uint64_t>(chunk, std::max<uint64_t>(g->rows, 1)) * g->dimp);
    for (uint64_t lo = 0; ok && lo < g->rows; lo += chunk) {
        const uint64_t m = std::min(chunk, g->rows - lo);
        if (cudaMemcpy(buf.data(), g->g32 + lo * g->dimp, m * g->dimp * sizeof(float),
                       cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
        ok = fwrite(buf.data(), sizeof(float), (size_t)m * g->dimp, f) == (size_t)m * g->dimp;
    }
    if (ok && g->rows) ok = fwrite(g->h_alive.data(), 1, (size_t)g->rows, f) == (size_t)g->rows;
    ok = (fflush(f) == 0) && ok;
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), path) == 0;
    if (!ok) { remove(tmp.c_str()); set_error("write to %s failed", path); cudaGetLastError(); return FRP_E_IO; }
    return FRP_OK;
}

int frp_gallery_load(int device, const char* path, frp_gallery_t** out) {
    if (!path || !out) { set_error("frp_gallery_load: null argument"); return FRP_E_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return FRP_E_IO; }
    SnapshotHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "FRPGAL1", 8) != 0) {
        fclose(f);
        set_error("%s is not a gallery snapshot", path);
        return FRP_E_IO;
    }
    // the header is untrusted input: every field is checked against what this library would
    // have written, and the payload size against the file, before anything is allocated
    bool sane = h.dim > 0 && h.dim <= 512 && h.dimp == pad_dim(h.dim) &&
                (h.metric == FRP_METRIC_L2 || h.metric == FRP_METRIC_COS) && h.dead <= h.rows &&
                h.rows < (1ull << 40);
    if (sane) {
        const uint64_t expect = sizeof(h) + h.rows * (uint64_t)h.dimp * sizeof(float) + h.rows;
        sane = fseek(f, 0, SEEK_END) == 0 && (uint64_t)ftell(f) == expect &&
               fseek(f, (long)sizeof(h), SEEK_SET) == 0;
    }
    if (!sane) {
        fclose(f);
        set_error("%s: corrupt or truncated snapshot header (dim=%d dimp=%d metric=%d rows=%llu)", path, h.dim,
                  h.dimp, h.metric, (unsigned long long)h.rows);
        return FRP_E_IO;
    }
    frp_gallery_t* g = nullptr;
    int rc = frp_gallery_create(device, h.dim, h.metric, h.rows, &g);
    if (rc) { fclose(f); return rc; }
    DeviceGuard dg(device);
    // rows are stored as the fp32 master padded to dimp (COS: already normalised).  They are
    // ingested VERBATIM: the fp16 copy, |g|^2 and the measured rounding error of the
    // certificate are all derived from exactly the master that is stored.
    const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / (sizeof(float) * h.dimp));
    const uint64_t nbuf = std::min<uint64_t>(chunk, std::max<uint64_t>(h.rows, 1));
    std::vector<float> buf((size_t)nbuf * h.dimp);
    float* stage = nullptr;
    bool ok = cudaMalloc((void**)&stage, (size_t)nbuf * h.dimp * sizeof(float)) == cudaSuccess;
    for (uint64_t lo = 0; ok && lo < h.rows; lo += chunk) {
        const uint64_t m = std::min(chunk, h.rows - lo);
        ok = fread(buf.data(), sizeof(float), (size_t)m * h.dimp, f) == (size_t)m * h.dimp;
        if (!ok) break;
        ok = cudaMemcpy(stage, buf.data(), m * h.dimp * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok) ok = ingest_device(g, stage, m, lo, 0, (uint64_t)h.dimp, true) == FRP_OK;
        if (ok) ok = cudaStreamSynchronize(0) == cudaSuccess;
    }
    if (stage) cudaFree(stage);
    g->rows = h.rows;
    g->h_alive.assign((size_t)h.rows, 1);
    uint64_t dead = 0;
    if (ok && h.rows) {
        ok = fread(g->h_alive.data(), 1, (size_t)h.rows, f) == (size_t)h.rows &&
             cudaMemcpy(g->alive, g->h_alive.data(), h.rows, cudaMemcpyHostToDevice) == cudaSuccess;
        for (uint64_t r = 0; ok && r < h.rows; ++r) {
            g->h_alive[(size_t)r] = g->h_alive[(size_t)r] ? 1 : 0;
            dead += g->h_alive[(size_t)r] ? 0 : 1;
        }
    }
    fclose(f);
    if (ok && dead != h.dead) { ok = false; }
    if (ok) ok = refresh_gmax(g, 0) == FRP_OK;
    if (!ok) { set_error("read of %s failed (I/O error or inconsistent tombstones)", path); frp_gallery_destroy(g); cudaGetLastError(); return FRP_E_IO; }
    g->dead = h.dead;
    g->base_row = h.base_row;
    *out = g;
    return FRP_OK;
}

}  // extern "C"
