// Instantiations of the k = 1 (best-matching unit) scan (see scan_impl.cuh).
#include "scan_impl.cuh"

int gmd_bmu_launch(int DP, const gmd_scan::ScanParams& p, cudaStream_t st) {
    switch (DP) {
#define X(d) \
    case d:  \
        return gmd_scan::bmu_launch<d>(p, st);
        GMD_BMU_DIMS(X)
#undef X
        default:
            break;
    }
    gmd_set_error("bmu scan: no kernel for padded dimension %d", DP);
    return GMD_ERR_UNSUPPORTED;
}
