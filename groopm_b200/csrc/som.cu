// Batched SOM training step on the device: the weight update of
// ONE iteration of SOM.train (som.py:328-384) for all T training vectors of the iteration,
// given their best-matching units on the frozen weights.
//
// The reference scatters: for every vector it adds  -(w - x) * stamp  onto a 3 x 3 tiling of
// the mesh around the BMU (som.py:355-364), then folds the 9 sheets columns first, rows
// second (:366-370) and clips (:374).  Here every (node, dimension) GATHERS instead: it walks
// the T vectors in the reference's order and keeps the nine per-sheet partial sums in
// registers, then folds them in the same order -- the float64 result is bit-identical and
// deterministic (no atomics), and nothing of size 9 x mesh is ever built.
#include "common.cuh"

namespace {

constexpr int SOM_CHUNK = 512;  // training vectors staged in shared memory at a time

__global__ void __launch_bounds__(128) som_update_kernel(const double* __restrict__ W, double* __restrict__ Wout,
                                                         int rows, int cols, int D, const double* __restrict__ X,
                                                         const int32_t* __restrict__ index, const int32_t* __restrict__ loc,
                                                         int T, const double* __restrict__ stamp, int sr,
                                                         const unsigned char* __restrict__ mask) {
    __shared__ int s_r[SOM_CHUNK], s_c[SOM_CHUNK], s_j[SOM_CHUNK];
    const int M = rows * cols;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < (long long)M * D;
    const int node = live ? (int)(e / D) : 0, dim = live ? (int)(e % D) : 0;
    const int nr = node / cols, nc = node - nr * cols;
    const double w = live ? W[e] : 0.0;
    const int sw = 2 * sr + 1;
    double P[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) P[a][b] = 0.0;

    for (int t0 = 0; t0 < T; t0 += SOM_CHUNK) {
        const int n = min(SOM_CHUNK, T - t0);
        __syncthreads();
        for (int x = threadIdx.x; x < n; x += blockDim.x) {
            const int l = loc[t0 + x];
            const int r = l / cols;  // som.py:352-353: row = int(loc/cols), col = loc - row*cols
            s_r[x] = r;
            s_c[x] = l - r * cols;
            s_j[x] = index[t0 + x];
        }
        __syncthreads();
        if (!live) continue;
        for (int x = 0; x < n; ++x) {
            const int dr0 = nr - s_r[x], dc0 = nc - s_c[x];
            // cheap reject: the node is outside the stamp on every sheet
            if (min(abs(dr0), rows - abs(dr0)) > sr || min(abs(dc0), cols - abs(dc0)) > sr) continue;
            // -1 * (worksheet - trainVector[j])   (som.py:358)
            const double tneg = -__dsub_rn(w, X[(size_t)s_j[x] * D + dim]);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int dr = dr0 + (a - 1) * rows;
                if (dr < -sr || dr > sr) continue;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const int dc = dc0 + (b - 1) * cols;
                    if (dc < -sr || dc > sr) continue;
                    // som.py:360-363 multiplies dimensions 0..3 by the stamp, by name; any further dimension
                    // of the patch keeps the whole difference
                    const double s = dim < 4 ? stamp[(dr + sr) * sw + (dc + sr)] : 1.0;
                    P[a][b] = __dadd_rn(P[a][b], __dmul_rn(tneg, s));  // *= stamp ; deltasheet += patch
                }
            }
        }
    }
    if (!live) return;
    // fold: columns of every row of sheets (som.py:366-367), then rows (:368-369)
    double mid[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) mid[a] = __dadd_rn(__dadd_rn(P[a][1], P[a][0]), P[a][2]);
    const double fold = __dadd_rn(__dadd_rn(mid[1], mid[0]), mid[2]);
    double v = w;
    if (!mask || mask[node]) v = fmin(fmax(__dadd_rn(w, fold), 0.0), 1.0);  // np.clip(., 0, 1)
    Wout[e] = v;
}

}  // namespace

extern "C" int gmd_som_update_f64(const double* W, double* Wout, int rows, int cols, int D, const double* X,
                                  int64_t n_train, const int32_t* index, const int32_t* loc, int T,
                                  const double* stamp, int stamp_radius, const unsigned char* mask, void* stream) {
    GMD_REQUIRE(W && Wout && X && index && loc && stamp, GMD_ERR_BAD_ARG, "som_update: null pointer");
    GMD_REQUIRE(rows > 0 && cols > 0 && D > 0 && T >= 0 && n_train > 0 && stamp_radius >= 0, GMD_ERR_BAD_ARG,
                "som_update: bad sizes (rows=%d cols=%d D=%d T=%d)", rows, cols, D, T);
    // a stamp wider than the 3 x 3 tiling would index outside the reference's worksheet
    GMD_REQUIRE(stamp_radius <= rows && stamp_radius <= cols, GMD_ERR_BAD_ARG,
                "som_update: stamp radius %d exceeds the mesh (%d x %d)", stamp_radius, rows, cols);
    const long long n = (long long)rows * cols * D;
    som_update_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        W, Wout, rows, cols, D, X, index, loc, T, stamp, stamp_radius, mask);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
