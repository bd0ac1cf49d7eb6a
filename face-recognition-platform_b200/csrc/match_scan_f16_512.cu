#include "match_scan_impl.cuh"
FRP_SCAN_INSTANTIATE(scan_f16_512, __half, 512)
