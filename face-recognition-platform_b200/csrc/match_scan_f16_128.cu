#include "match_scan_impl.cuh"
FRP_SCAN_INSTANTIATE(scan_f16_128, __half, 128)
