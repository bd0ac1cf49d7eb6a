#include "match_scan_impl.cuh"
FRP_SCAN_INSTANTIATE(scan_f16_64, __half, 64)
