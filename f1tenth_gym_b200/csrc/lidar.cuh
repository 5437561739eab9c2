// lidar.cuh — sphere-tracing ray-march on the distance-transform grid, iTTC predicate and the
// opponent ray-cast, fp64.
//
// Behavioural spec: reference gym/f110_gym/envs/laser_models.py:55-346.  Compiled with -fmad=false
// (`x += d*c` is two roundings in the reference; SURVEY.md 7.1 shows fp32 or FMA-contracted
// positions flip `int(x/res)` cells and break 1e-4 parity).
#pragma once
#include <math.h>
#include <stdint.h>

namespace f110 {

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream is still running; it must call pdl_wait() before it touches anything the
// predecessor writes.  pdl_launch_dependents() lets the NEXT kernel of the stream start its launch early.  Both are no-ops
// for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct MapView {
    const double *__restrict__ dt;
    const double *__restrict__ dt_cells;   // dt / resolution (cell units); fast path only, else NULL
    const uint8_t *__restrict__ dt_codes;  // rank code of dt_cells per cell (255 = escape), or NULL
    const double *__restrict__ dt_lut;     // [256] code -> dt_cells value (exact), entry 255 unused
    const double *__restrict__ sines;
    const double *__restrict__ cosines;
    double orig_x, orig_y, orig_c, orig_s, resolution, inv_resolution, x_max, y_max;
    double eps, max_range, dt_oob, theta_dis_f;
    int32_t height, width, theta_dis;
};

// ---- fast path in CELL UNITS --------------------------------------------------------------------
// When resolution = 2^-k and the map origin is unrotated, scaling every length by 1/res is exact in
// fp64 and commutes with rounding: X = x/res, D = d/res satisfy fl(X + fl(D*c)) = fl(x + fl(d*c))/res.
// Marching in cell units removes the two multiplies by 1/res per lookup; floor() is taken with one
// round-down add of 2^52+2^51 on the fp64 pipe (F2I.F64 would go through the quarter-rate XU pipe),
// the four fp64 bounds tests become two unsigned compares, and an off-map lookup is redirected to the
// last cell, which holds exactly the value the reference reads through dt[-1,-1].
struct CellConsts {
    double ox, oy;        // orig / res
    double eps, tmax;     // eps / res, max_range / res
    unsigned width, height, last;
};

__device__ __forceinline__ unsigned cell_index(double X, double Y, const CellConsts &k) {
    const double MAGIC = 6755399441055744.0;   // 2^52 + 2^51
    int c = __double2loint(__dadd_rd(X - k.ox, MAGIC));
    int r = __double2loint(__dadd_rd(Y - k.oy, MAGIC));
    bool inb = ((unsigned)c < k.width) && ((unsigned)r < k.height);
    return inb ? (unsigned)r * k.width + (unsigned)c : k.last;
}

// trace_ray (laser_models.py:106-146) in cell units; returns total (cell units, unclamped > tmax possible)
template <bool COUNT>
__device__ __forceinline__ double trace_ray_cells(const double *__restrict__ dtc, double X, double Y, double s,
                                                  double c, const CellConsts &k, int &nlook) {
    double D = __ldg(dtc + cell_index(X, Y, k));
    double T = D;
    int n = 1;
    while (D > k.eps && T <= k.tmax) {
        X = X + D * c;
        Y = Y + D * s;
        D = __ldg(dtc + cell_index(X, Y, k));
        T = T + D;
        if (COUNT) n++;
    }
    nlook = n;
    return T;
}

// Coded table: one BYTE per cell holding the rank of the cell's DT value among the 255 smallest
// distinct values of the map (on example_map: every distance below 27.3 cells = 1.7 m, i.e. every
// cell a ray from the track can touch), 255 = "escape: read the fp64 table".  Lossless by
// construction.  A 32-byte sector now holds 32 cells instead of 4, so the per-scan footprint drops
// from ~1300 to ~480 sectors and the march's dependent load becomes an L1 hit almost always; the
// code -> fp64 lookup is a second, always-L1-resident 2 KB table.
__device__ __forceinline__ double coded_lookup(const MapView &m, unsigned idx) {
    const unsigned code = __ldg(m.dt_codes + idx);
    if (code == 255u) return __ldg(m.dt_cells + idx);
    return __ldg(m.dt_lut + code);
}

// laser_models.py:55-104 xy_2_rc + distance_transform.
// FAST: resolution is a power of two and the origin is unrotated, so x_rot == x - orig_x exactly,
// x_rot/res == x_rot*inv_res exactly, and the four fp64 bounds tests collapse to two unsigned integer
// compares on floor() of the scaled coordinate.  Off-map reads dt[-1,-1] (numba negative-index wrap).
template <bool FAST>
__device__ __forceinline__ double dt_lookup(const MapView &m, double x, double y) {
    if (FAST) {
        double tx = (x - m.orig_x) * m.inv_resolution;
        double ty = (y - m.orig_y) * m.inv_resolution;
        int c = __double2int_rd(tx);
        int r = __double2int_rd(ty);
        bool inb = ((unsigned)c < (unsigned)m.width) && ((unsigned)r < (unsigned)m.height);
        return inb ? __ldg(m.dt + (size_t)r * (size_t)m.width + (size_t)c) : m.dt_oob;
    } else {
        double x_trans = x - m.orig_x;
        double y_trans = y - m.orig_y;
        double x_rot = x_trans * m.orig_c + y_trans * m.orig_s;
        double y_rot = -x_trans * m.orig_s + y_trans * m.orig_c;
        if (x_rot < 0 || x_rot >= m.x_max || y_rot < 0 || y_rot >= m.y_max) return m.dt_oob;
        int c = (int)(x_rot / m.resolution);
        int r = (int)(y_rot / m.resolution);
        return __ldg(m.dt + (size_t)r * (size_t)m.width + (size_t)c);
    }
}

// laser_models.py:106-146 trace_ray.  Returns the clamped range; nlook gets the number of DT lookups.
template <bool FAST>
__device__ __forceinline__ double trace_ray(const MapView &m, double x, double y, double s, double c,
                                            int &nlook) {
    double d = dt_lookup<FAST>(m, x, y);
    double total = d;
    int n = 1;
    while (d > m.eps && total <= m.max_range) {
        x = x + d * c;
        y = y + d * s;
        d = dt_lookup<FAST>(m, x, y);
        total = total + d;
        n++;
    }
    nlook = n;
    return (total > m.max_range) ? m.max_range : total;
}

// laser_models.py:167-172: first beam's LUT index from the scan pose yaw.
__device__ __forceinline__ double theta_index0(double yaw, double fov, double theta_dis_f) {
    double ti = theta_dis_f * (yaw - fov / 2.) / (2. * M_PI);
    if (!(fabs(ti) < theta_dis_f)) ti = fmod(ti, theta_dis_f);      // fmod(x, y) == x for |x| < |y|: the usual case
    while (ti < 0) ti += theta_dis_f;
    return ti;
}

// laser_models.py:175-184: the reference walks theta_index sequentially (`+= inc`, wrap at theta_dis).
// Per-lane closed form ti0 + i*inc (mod theta_dis).  Every sequential add rounds by at most half an ulp
// of a value < 2048 (1.14e-13) and the wrap subtraction is exact, so after i <= num_beams adds the
// sequential value is within num_beams * 1.14e-13 of the closed form (2.5e-10 for 2160 beams): int() can
// only disagree when the fractional part is within `guard` (host: 4 * that bound) of an integer, and in
// that rare case (~1e-9 per beam; the replay costs up to ~20 us on one lane, so it must stay rare) the
// exact sequential recurrence is replayed for this beam.
__device__ __forceinline__ int beam_theta_index(double ti0, int i, double inc, double theta_dis_f,
                                                double guard = 1e-6) {
    double v = ti0 + (double)i * inc;
    while (v >= theta_dis_f) v -= theta_dis_f;
    int iv = (int)v;
    double fr = v - (double)iv;
    if (fr < guard || fr > 1.0 - guard) {
        double t = ti0;
        for (int k = 0; k < i; k++) {
            t += inc;
            while (t >= theta_dis_f) t -= theta_dis_f;
        }
        return (int)t;
    }
    return iv;
}

// laser_models.py:188-217 check_ttc_jit, one beam (error_model='numpy': x/0 -> inf/nan, compares false)
__device__ __forceinline__ bool ttc_hit(double range, double vel, double cos_i, double side_i, double thresh) {
    if (vel == 0.0) return false;
    double proj_vel = vel * cos_i;
    double a = range - side_i;
    // necessary condition for 0 <= fl(a/p) < thresh, with margin; the exact IEEE division (a ~30-instruction
    // sequence) then only runs for beams that are actually near a collision
    if (!(fabs(a) <= thresh * fabs(proj_vel) * 1.000001)) return false;
    double ttc = a / proj_vel;
    return (ttc < thresh) && (ttc >= 0.0);
}

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// laser_models.py:249-280 get_range, with (v3x, v3y) = (cos, sin)(beam_theta + pi/2) hoisted per beam
__device__ __forceinline__ double get_range(double ox, double oy, double v3x, double v3y, double vax,
                                            double vay, double vbx, double vby) {
    double v1x = ox - vax, v1y = oy - vay;
    double v2x = vbx - vax, v2y = vby - vay;
    double denom = v2x * v3x + v2y * v3y;
    double distance = INFINITY;
    if (fabs(denom) > 0.0) {
        double d1 = cross2(v2x, v2y, v1x, v1y) / denom;
        double d2 = (v1x * v3x + v1y * v3y) / denom;
        if (d1 >= 0.0 && d2 >= 0.0 && d2 <= 1.0) distance = d1;
    } else {
        // are_collinear(o, va, vb) :232-247
        double bax = vax - ox, bay = vay - oy, cax = ox - vbx, cay = oy - vby;
        if (fabs(cross2(bax, bay, cax, cay)) < 1e-8) {
            double da = sqrt((vax - ox) * (vax - ox) + (vay - oy) * (vay - oy));
            double db = sqrt((vbx - ox) * (vbx - ox) + (vby - oy) * (vby - oy));
            distance = da < db ? da : db;
        }
    }
    return distance;
}

// get_range for the opponent ray-cast, where the only use of the result is `scan[i] = min(scan[i], range)`: identical
// outcome, but the two fp64 divisions only run for an edge that can actually shorten the beam.  With d1 = c / denom and
// d2 = n2 / denom (both IEEE quotients), the reference's test `d1 >= 0 and 0 <= d2 <= 1` fails for certain when
//   |n2| > |denom| (1 + 1e-12)                      (d2 > 1: rounding cannot bring it back to 1),
//   n2, denom (or c, denom) have opposite signs and the quotient cannot underflow to -0.0   (d2 < 0, d1 < 0),
// and an edge with |c| >= cur |denom| (1 + 1e-12) has d1 >= cur: it cannot lower a beam that currently reads `cur`.
__device__ __forceinline__ double get_range_below(double ox, double oy, double v3x, double v3y, double vax, double vay,
                                                  double vbx, double vby, double cur) {
    const double v1x = ox - vax, v1y = oy - vay;
    const double v2x = vbx - vax, v2y = vby - vay;
    const double denom = v2x * v3x + v2y * v3y;
    double distance = INFINITY;
    if (fabs(denom) > 0.0) {
        const double c = cross2(v2x, v2y, v1x, v1y);
        const double n2 = v1x * v3x + v1y * v3y;
        const double ad = fabs(denom);
        if (fabs(n2) > ad * (1.0 + 1e-12)) return distance;
        if (((n2 < 0.0) != (denom < 0.0)) && fabs(n2) > 1e-290 * ad) return distance;
        if (((c < 0.0) != (denom < 0.0)) && fabs(c) > 1e-290 * ad) return distance;
        if (fabs(c) >= cur * ad * (1.0 + 1e-12)) return distance;
        const double d1 = c / denom;
        const double d2 = n2 / denom;
        if (d1 >= 0.0 && d2 >= 0.0 && d2 <= 1.0) distance = d1;
    } else {
        double bax = vax - ox, bay = vay - oy, cax = ox - vbx, cay = oy - vby;
        if (fabs(cross2(bax, bay, cax, cay)) < 1e-8) {
            double da = sqrt((vax - ox) * (vax - ox) + (vay - oy) * (vay - oy));
            double db = sqrt((vbx - ox) * (vbx - ox) + (vby - oy) * (vby - oy));
            distance = da < db ? da : db;
        }
    }
    return distance;
}

// np.argmin(np.abs(scan_angles - a)) (first minimum).  scan_angles is strictly increasing, so
// |scan_angles[i] - a| is convex in i: evaluate the table around the analytic estimate.
__device__ __forceinline__ int nearest_beam(const double *__restrict__ scan_angles, int num_beams,
                                            double fov, double angle_increment, double a) {
    double est = (a + fov / 2.) / angle_increment;
    int i0;
    if (!(est > 0.0)) i0 = 0;
    else if (est >= (double)(num_beams - 1)) i0 = num_beams - 1;
    else i0 = (int)est;
    int lo = max(i0 - 2, 0), hi = min(i0 + 3, num_beams - 1);
    int best = lo;
    double bv = fabs(scan_angles[lo] - a);
    for (int i = lo + 1; i <= hi; i++) {
        double v = fabs(scan_angles[i] - a);
        if (v < bv) { bv = v; best = i; }
    }
    return best;
}

// laser_models.py:282-315 get_blocked_view_indices, warp-cooperative: lanes 0-3 take one vertex each and lane 4
// the ego heading, so the five atan2 evaluations run as ONE pass through the atan2 code instead of five.
// (cos_yaw, sin_yaw) = (cos, sin)(pose yaw).  All 32 lanes must call; every lane gets the result.
__device__ __forceinline__ void blocked_view_indices_warp(double px, double py, double cos_yaw, double sin_yaw,
                                                          const double v[8], const double *__restrict__ scan_angles,
                                                          int num_beams, double fov, double angle_increment, int lane,
                                                          double cx, double cy, int &min_ind, int &max_ind,
                                                          double &centre_dir) {
    double ay, ax;
    const int k = lane & 3;
    if (lane == 4) { ay = sin_yaw; ax = cos_yaw; }
    else if (lane == 5) { ay = cy - py; ax = cx - px; }      // world direction to the opponent's centre
    else {
        // (select with compile-time indices: a dynamically indexed v[] would live in local memory)
        const double vkx = (k == 0) ? v[0] : (k == 1) ? v[2] : (k == 2) ? v[4] : v[6];
        const double vky = (k == 0) ? v[1] : (k == 1) ? v[3] : (k == 2) ? v[5] : v[7];
        const double vx = vkx - px, vy = vky - py;
        const double norm = sqrt(vx * vx + vy * vy);
        ax = vx / norm; ay = vy / norm;
    }
    const double at = atan2(ay, ax);
    const double ego_a = __shfl_sync(0xffffffffu, at, 4);
    centre_dir = __shfl_sync(0xffffffffu, at, 5);
    double angle = ego_a - at;
    if (angle > M_PI) angle = angle - 2 * M_PI;
    else if (angle < -M_PI) angle = angle + 2 * M_PI;
    int ind = nearest_beam(scan_angles, num_beams, fov, angle_increment, -angle);
    int lo = ind, hi = ind;       // lanes 0-3 hold the four vertex indices (lanes >= 6 mirror lane & 3)
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, 1)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 1));
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, 2)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, 2));
    min_ind = __shfl_sync(0xffffffffu, lo, 0);
    max_ind = __shfl_sync(0xffffffffu, hi, 0);
}

// laser_models.py:282-315 get_blocked_view_indices, one thread, with the heading given as (cos, sin) like the warp-cooperative
// version above (bit-identical to it: the same five atan2 arguments, the same min / max); also returns the world direction to
// the opponent's centre (cx, cy)
__device__ __forceinline__ void blocked_view_indices_cs(double px, double py, double cos_yaw, double sin_yaw, const double v[8],
                                                        const double *__restrict__ scan_angles, int num_beams, double fov,
                                                        double angle_increment, double cx, double cy, int &min_ind, int &max_ind,
                                                        double &centre_dir) {
    const double ego_a = atan2(sin_yaw, cos_yaw);
    centre_dir = atan2(cy - py, cx - px);
    int lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double vx = v[2 * i] - px, vy = v[2 * i + 1] - py;
        const double norm = sqrt(vx * vx + vy * vy);
        const double ax = vx / norm, ay = vy / norm;
        double angle = ego_a - atan2(ay, ax);
        if (angle > M_PI) angle = angle - 2 * M_PI;
        else if (angle < -M_PI) angle = angle + 2 * M_PI;
        const int ind = nearest_beam(scan_angles, num_beams, fov, angle_increment, -angle);
        if (i == 0) { lo = hi = ind; }
        else { lo = min(lo, ind); hi = max(hi, ind); }
    }
    min_ind = lo;
    max_ind = hi;
}

// laser_models.py:282-315 get_blocked_view_indices
__device__ __forceinline__ void blocked_view_indices(double px, double py, double yaw, const double v[8],
                                                     const double *__restrict__ scan_angles, int num_beams,
                                                     double fov, double angle_increment, int &min_ind,
                                                     int &max_ind) {
    double ego_a = atan2(sin(yaw), cos(yaw));
    int lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double vx = v[2 * i] - px, vy = v[2 * i + 1] - py;
        double norm = sqrt(vx * vx + vy * vy);
        double ux = vx / norm, uy = vy / norm;
        double angle = ego_a - atan2(uy, ux);
        if (angle > M_PI) angle = angle - 2 * M_PI;
        else if (angle < -M_PI) angle = angle + 2 * M_PI;
        int ind = nearest_beam(scan_angles, num_beams, fov, angle_increment, -angle);
        if (i == 0) { lo = hi = ind; }
        else { lo = min(lo, ind); hi = max(hi, ind); }
    }
    min_ind = lo;
    max_ind = hi;
}

// ---- counter-based RNG for the optional scan noise (laser_models.py:450-452) -------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// one standard normal per (seed, stream, index) via Box-Muller on two 32-bit uniforms
__device__ __forceinline__ double normal_sample(uint64_t seed, uint64_t stream, uint64_t index) {
    uint32_t c[4] = { (uint32_t)index, (uint32_t)(index >> 32), (uint32_t)stream, (uint32_t)(stream >> 32) };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = ((double)c[0] + 1.0) * (1.0 / 4294967296.0);   // (0, 1]
    double u2 = (double)c[1] * (1.0 / 4294967296.0);           // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

}  // namespace f110
