#include "match_scan_impl.cuh"
FRP_SCAN_INSTANTIATE(scan_f32_256, float, 256)
