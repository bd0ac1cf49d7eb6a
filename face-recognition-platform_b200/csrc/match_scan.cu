// Dispatcher of the streaming scan matcher (K5); kernels live in match_scan_impl.cuh.
#include "match_scan.h"

namespace frp {

int scan_num_lists(const frp_gallery* g, bool /*use_master*/) {
    const uint64_t nchunks = (g->rows + kScanChunkRows - 1) / kScanChunkRows;
    uint64_t n = (uint64_t)g->sm_count;
    if (nchunks < n) n = nchunks;
    if (n < 1) n = 1;
    return (int)n;
}

int launch_match_scan(const frp_gallery* g, bool use_master, const float* qprep, const QParams* qp,
                      int B, int KP, uint64_t* cand, int nlists, cudaStream_t st) {
    if (g->rows >= (1ull << 32)) {
        set_error("match_scan: %llu rows exceed the 2^32 rows-per-shard limit", (unsigned long long)g->rows);
        return FRP_E_UNSUPPORTED;
    }
    ScanArgs a;
    a.g = use_master ? (const void*)g->g32 : (const void*)g->g16;
    a.norm2 = g->norm2;
    a.alive = g->alive;
    a.q = qprep;
    a.qp = qp;
    a.cand = cand;
    a.rows = (uint32_t)g->rows;
    a.nchunks = (uint32_t)((g->rows + kScanChunkRows - 1) / kScanChunkRows);
    a.B = B;
    a.metric = g->metric;
    a.stages = 0;
    a.device = g->device;
    switch (g->dimp) {
        case 64: return use_master ? scan_f32_64(a, B, KP, nlists, st) : scan_f16_64(a, B, KP, nlists, st);
        case 128: return use_master ? scan_f32_128(a, B, KP, nlists, st) : scan_f16_128(a, B, KP, nlists, st);
        case 256: return use_master ? scan_f32_256(a, B, KP, nlists, st) : scan_f16_256(a, B, KP, nlists, st);
        case 512: return use_master ? scan_f32_512(a, B, KP, nlists, st) : scan_f16_512(a, B, KP, nlists, st);
        default:
            set_error("match_scan: unsupported padded dim %d (supported: 64, 128, 256, 512)", g->dimp);
            return FRP_E_UNSUPPORTED;
    }
}

}  // namespace frp
