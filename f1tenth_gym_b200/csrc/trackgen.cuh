// Track rasteriser for generated tracks (reference unittest/random_trackgen.py:156-165, 167-178).
// The reference offsets the closed centerline by +-WIDTH with shapely and draws the two offset curves with
// matplotlib; an offset curve is the level set {p : dist(p, centerline) = WIDTH}, so a wall pixel is one whose centre
// lies within half a line width of that level set.  One thread per pixel, distance to every segment of the closed
// polyline (a few hundred) from shared memory; all fp64, no contraction, so a numpy restatement is bit-identical.
#pragma once
#include <stdint.h>

namespace f110 {

// seg = [M][5]: ax, ay, bx-ax, by-ay, 1/|b-a|^2 (0 for a degenerate segment), pixel units.
__global__ void __launch_bounds__(256) k_rasterize_track(const double *__restrict__ seg, int M, double r_in2, double r_out2,
                                                         int H, int W, uint8_t *__restrict__ occupied,
                                                         double *__restrict__ dist2_out) {
    extern __shared__ double s_seg[];
    for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < 5 * M; i += blockDim.x * blockDim.y) s_seg[i] = seg[i];
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= W || r >= H) return;
    const double px = (double)c + 0.5, py = (double)r + 0.5;
    double best = 1.0e300;
    for (int i = 0; i < M; i++) {
        const double *s = s_seg + 5 * i;
        const double dx = px - s[0], dy = py - s[1];
        double t = (dx * s[2] + dy * s[3]) * s[4];
        t = fmin(fmax(t, 0.0), 1.0);
        const double qx = dx - t * s[2], qy = dy - t * s[3];
        best = fmin(best, qx * qx + qy * qy);
    }
    const size_t o = (size_t)r * W + c;
    occupied[o] = (best >= r_in2 && best <= r_out2) ? 1 : 0;
    if (dist2_out) dist2_out[o] = best;
}

}  // namespace f110
