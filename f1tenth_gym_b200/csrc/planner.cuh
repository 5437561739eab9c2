// planner.cuh — batched pure-pursuit planner, one warp per agent, fp64.
//
// Behavioural spec: reference examples/waypoint_follow.py
//   nearest_point_on_trajectory :15-47, first_point_on_trajectory_intersecting_circle :49-131 (wrap=True),
//   get_actuation :133-144, PurePursuitPlanner._get_current_waypoint :183-202 and .plan :204-217
// including its quirks: the look-ahead POINT used for steering is the waypoint at the start of the
// intersected segment (wpts[i2]), the speed is taken from the NEAREST segment's row, `end = next + 1e-6`.
#pragma once
#include <math.h>
#include <stdint.h>

namespace f110 {

// does segment i (start (sx,sy), end (ex,ey) + 1e-6) intersect the look-ahead circle with a valid parameter?
__device__ __forceinline__ bool pp_segment_hits(double sx, double sy, double ex, double ey, double px, double py,
                                                double radius, bool is_start, double start_t) {
    const double Vx = (ex + 1e-6) - sx, Vy = (ey + 1e-6) - sy;
    const double a = Vx * Vx + Vy * Vy;
    const double b = 2.0 * (Vx * (sx - px) + Vy * (sy - py));
    const double c = (sx * sx + sy * sy) + (px * px + py * py) - 2.0 * (sx * px + sy * py) - radius * radius;
    double disc = b * b - 4 * a * c;
    if (disc < 0) return false;
    disc = sqrt(disc);
    const double t1 = (-b - disc) / (2.0 * a), t2 = (-b + disc) / (2.0 * a);
    if (is_start)
        return (t1 >= 0.0 && t1 <= 1.0 && t1 >= start_t) || (t2 >= 0.0 && t2 <= 1.0 && t2 >= start_t);
    return (t1 >= 0.0 && t1 <= 1.0) || (t2 >= 0.0 && t2 <= 1.0);
}

// actions_out[a] = (steering angle, speed): the layout f110_step consumes.
// Several waypoint tables (one per track of a multi-map batch) are concatenated in wx/wy/wv: table t occupies
// [table_start[t], table_start[t+1]) and pose a follows table pose_table[a]; table_start == NULL: one table of n rows.
__global__ void __launch_bounds__(128) k_pure_pursuit(const double *__restrict__ wx, const double *__restrict__ wy,
                                                      const double *__restrict__ wv, int n,
                                                      const int32_t *__restrict__ table_start,
                                                      const int32_t *__restrict__ pose_table,
                                                      const double *__restrict__ pose_x, const double *__restrict__ pose_y,
                                                      const double *__restrict__ pose_theta, int M, double lookahead,
                                                      double vgain, double wheelbase, double max_reacquire,
                                                      double *__restrict__ actions_out) {
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (a >= M) return;
    if (table_start) {
        const int t = pose_table[a];
        const int off = table_start[t];
        n = table_start[t + 1] - off;
        wx += off; wy += off; wv += off;
    }
    const double px = pose_x[a], py = pose_y[a], th = pose_theta[a];
    // nearest_point_on_trajectory: np.argmin = first minimum
    double bd = INFINITY, bt = 0.0;
    int bi = 0x7fffffff;
    for (int i = lane; i < n - 1; i += 32) {
        const double dx = wx[i + 1] - wx[i], dy = wy[i + 1] - wy[i];
        const double l2 = dx * dx + dy * dy;
        const double dot = (px - wx[i]) * dx + (py - wy[i]) * dy;
        double t = dot / l2;
        if (t < 0.0) t = 0.0;
        if (t > 1.0) t = 1.0;
        const double qx = wx[i] + t * dx, qy = wy[i] + t * dy;
        const double ex = px - qx, ey = py - qy;
        const double d = sqrt(ex * ex + ey * ey);
        if (d < bd) { bd = d; bi = i; bt = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const double ot = __shfl_xor_sync(0xffffffffu, bt, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bt = ot; }
    }
    if (bi == 0x7fffffff) bi = 0;      // all distances NaN: argmin of NaNs is 0
    bool have = false;
    int i2 = 0;
    if (bd < lookahead) {
        const double tt = (double)bi + bt;
        const int start_i = (int)tt;
        const double start_t = fmod(tt, 1.0);
        int first = -1000000;
        for (int base = start_i; base < n - 1 && first == -1000000; base += 32) {
            const int i = base + lane;
            bool hit = false;
            if (i < n - 1) hit = pp_segment_hits(wx[i], wy[i], wx[i + 1], wy[i + 1], px, py, lookahead, i == start_i, start_t);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) first = base + (__ffs(m) - 1);
        }
        if (first == -1000000) {       // wrap=True: segments -1 .. start_i-1 with Python's modulo indexing
            for (int base = -1; base < start_i && first == -1000000; base += 32) {
                const int i = base + lane;
                bool hit = false;
                if (i < start_i) {
                    const int i0 = ((i % n) + n) % n, i1 = (((i + 1) % n) + n) % n;
                    hit = pp_segment_hits(wx[i0], wy[i0], wx[i1], wy[i1], px, py, lookahead, false, 0.0);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) first = base + (__ffs(m) - 1);
            }
        }
        if (first != -1000000) { have = true; i2 = first < 0 ? first + n : first; }
    } else if (bd < max_reacquire) {
        have = true; i2 = bi;
    }
    if (lane == 0) {
        double speed = 4.0, steer = 0.0;        // plan(): no waypoint -> (4.0, 0.0)
        if (have) {
            const double waypoint_y = sin(-th) * (wx[i2] - px) + cos(-th) * (wy[i2] - py);
            if (fabs(waypoint_y) < 1e-6) steer = 0.;
            else {
                const double radius = 1 / (2.0 * waypoint_y / (lookahead * lookahead));
                steer = atan(wheelbase / radius);
            }
            speed = vgain * wv[bi];
        }
        actions_out[2 * (size_t)a] = steer;
        actions_out[2 * (size_t)a + 1] = speed;
    }
}

}  // namespace f110
