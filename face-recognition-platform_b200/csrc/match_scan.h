// Shared declarations of the streaming scan matcher (see match_scan_impl.cuh).
#pragma once
#include "frp_internal.h"

namespace frp {

constexpr int kScanChunkRows = 32;   // rows per shared-memory stage

struct ScanArgs {
    const void* g;          // [rows][DIMP] T
    const float* norm2;     // [rows]
    const uint8_t* alive;   // [rows]
    const float* q;         // [B][DIMP] prepared fp32 queries
    const QParams* qp;      // [B]
    uint64_t* cand;         // [B][nlists][KP]
    uint32_t rows;
    uint32_t nchunks;
    int B;
    int metric;
    int stages;
    int device;
};

#define FRP_SCAN_DECL(NAME) int NAME(const ScanArgs& a, int B, int KP, int nlists, cudaStream_t st);
FRP_SCAN_DECL(scan_f16_64) FRP_SCAN_DECL(scan_f16_128) FRP_SCAN_DECL(scan_f16_256) FRP_SCAN_DECL(scan_f16_512)
FRP_SCAN_DECL(scan_f32_64) FRP_SCAN_DECL(scan_f32_128) FRP_SCAN_DECL(scan_f32_256) FRP_SCAN_DECL(scan_f32_512)
#undef FRP_SCAN_DECL

}  // namespace frp
