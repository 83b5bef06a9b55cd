// Instantiations of the fused distance + top-k scan for FP32 form 1 (see scan_impl.cuh).
#include "scan_impl.cuh"

int gmd_scan_launch_form1(int DP, const gmd_scan::ScanParams& p, cudaStream_t st) {
    switch (DP) {
#define X(d) \
    case d:  \
        return gmd_scan::scan_dispatch<d, 1>(p, st);
        GMD_SCAN_DIMS(X)
#undef X
        default:
            break;
    }
    gmd_set_error("scan: no kernel for padded dimension %d", DP);
    return GMD_ERR_UNSUPPORTED;
}
