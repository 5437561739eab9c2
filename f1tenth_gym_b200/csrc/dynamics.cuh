// dynamics.cuh — pid + single-track vehicle dynamics + RK4/Euler tick, fp64, one thread per agent.
//
// Behavioural spec: reference gym/f110_gym/envs/dynamic_models.py:29-221 and
// base_classes.py:256-409 (RaceCar.update_pose).  The file is compiled with -fmad=false: every
// a*b+c below is two roundings, exactly like the reference's numba code (no FMA contraction), and the
// operation order follows Python operator precedence so that results match bit for bit up to the
// last-ulp differences of CUDA's sin/cos/tan versus the host libm.
#pragma once
#include <math.h>

namespace f110 {

enum { P_MU = 0, P_CSF, P_CSR, P_LF, P_LR, P_H, P_M, P_I, P_SMIN, P_SMAX, P_SVMIN, P_SVMAX,
       P_VSWITCH, P_AMAX, P_VMIN, P_VMAX, P_WIDTH, P_LENGTH };

// dynamic_models.py:29-60
__device__ __forceinline__ double accl_constraints(double vel, double accl, double v_switch, double a_max,
                                                   double v_min, double v_max) {
    double pos_limit = (vel > v_switch) ? a_max * v_switch / vel : a_max;
    if ((vel <= v_min && accl <= 0) || (vel >= v_max && accl >= 0)) accl = 0.;
    else if (accl <= -a_max) accl = -a_max;
    else if (accl >= pos_limit) accl = pos_limit;
    return accl;
}

// dynamic_models.py:62-87
__device__ __forceinline__ double steering_constraint(double angle, double sv, double s_min, double s_max,
                                                      double sv_min, double sv_max) {
    if ((angle <= s_min && sv <= 0) || (angle >= s_max && sv >= 0)) sv = 0.;
    else if (sv <= sv_min) sv = sv_min;
    else if (sv >= sv_max) sv = sv_max;
    return sv;
}

// dynamic_models.py:90-121 (vehicle_dynamics_ks): x has 5 entries (x, y, steer, v, yaw), f has 5
__device__ __forceinline__ void vehicle_dynamics_ks(const double x[5], double u_sv, double u_accl,
                                                    const double *__restrict__ p, double f[5]) {
    const double lwb = p[P_LF] + p[P_LR];
    const double u0 = steering_constraint(x[2], u_sv, p[P_SMIN], p[P_SMAX], p[P_SVMIN], p[P_SVMAX]);
    const double u1 = accl_constraints(x[3], u_accl, p[P_VSWITCH], p[P_AMAX], p[P_VMIN], p[P_VMAX]);
    f[0] = x[3] * cos(x[4]);
    f[1] = x[3] * sin(x[4]);
    f[2] = u0;
    f[3] = u1;
    f[4] = x[3] / lwb * tan(x[2]);
}

// dynamic_models.py:123-176 (vehicle_dynamics_st, with the :90-121 kinematic model for |v| < 0.5)
__device__ __forceinline__ void vehicle_dynamics_st(const double x[7], double u_sv, double u_accl,
                                                    const double *__restrict__ p, double f[7]) {
    const double g = 9.81;
    const double mu = p[P_MU], C_Sf = p[P_CSF], C_Sr = p[P_CSR], lf = p[P_LF], lr = p[P_LR], h = p[P_H],
                 m = p[P_M], I = p[P_I];
    double u0 = steering_constraint(x[2], u_sv, p[P_SMIN], p[P_SMAX], p[P_SVMIN], p[P_SVMAX]);
    double u1 = accl_constraints(x[3], u_accl, p[P_VSWITCH], p[P_AMAX], p[P_VMIN], p[P_VMAX]);
    if (fabs(x[3]) < 0.5) {
        // kinematic branch: the constraints are applied a second time on the constrained inputs
        double lwb = lf + lr;
        double v0 = steering_constraint(x[2], u0, p[P_SMIN], p[P_SMAX], p[P_SVMIN], p[P_SVMAX]);
        double v1 = accl_constraints(x[3], u1, p[P_VSWITCH], p[P_AMAX], p[P_VMIN], p[P_VMAX]);
        double tan_d = tan(x[2]);
        double cos_d = cos(x[2]);
        double s4, c4;
        sincos(x[4], &s4, &c4);         // one range reduction; same values as sin() / cos() (tests pin the state)
        f[0] = x[3] * c4;
        f[1] = x[3] * s4;
        f[2] = v0;
        f[3] = v1;
        f[4] = x[3] / lwb * tan_d;
        f[5] = u1 / lwb * tan_d + x[3] / (lwb * (cos_d * cos_d)) * u0;
        f[6] = 0.;
    } else {
        double ang = x[6] + x[4];
        double sa, ca;
        sincos(ang, &sa, &ca);
        f[0] = x[3] * ca;
        f[1] = x[3] * sa;
        f[2] = u0;
        f[3] = u1;
        f[4] = x[5];
        f[5] = -mu * m / (x[3] * I * (lr + lf)) *
                   (lf * lf * C_Sf * (g * lr - u1 * h) + lr * lr * C_Sr * (g * lf + u1 * h)) * x[5]
               + mu * m / (I * (lr + lf)) * (lr * C_Sr * (g * lf + u1 * h) - lf * C_Sf * (g * lr - u1 * h)) * x[6]
               + mu * m / (I * (lr + lf)) * lf * C_Sf * (g * lr - u1 * h) * x[2];
        f[6] = (mu / (x[3] * x[3] * (lr + lf)) * (C_Sr * (g * lf + u1 * h) * lr - C_Sf * (g * lr - u1 * h) * lf) - 1) * x[5]
               - mu / (x[3] * (lr + lf)) * (C_Sr * (g * lf + u1 * h) + C_Sf * (g * lr - u1 * h)) * x[6]
               + mu / (x[3] * (lr + lf)) * (C_Sf * (g * lr - u1 * h)) * x[2];
    }
}

// dynamic_models.py:178-221
__device__ __forceinline__ void pid(double speed, double steer, double current_speed, double current_steer,
                                    double max_sv, double max_a, double max_v, double min_v, double &accl,
                                    double &sv) {
    double steer_diff = steer - current_steer;
    // steer_diff / |steer_diff| is exactly +-1 for a finite non-zero value: no division needed
    sv = (fabs(steer_diff) > 1e-4) ? copysign(1.0, steer_diff) * max_sv : 0.0;
    double vel_diff = speed - current_speed;
    // the four kp cases (:198-213) are (10|2) * max_a / (max_v | -min_v): selecting the operands first gives the
    // same two roundings with one division and no divergence
    const double num = ((current_speed > 0.) ? 10.0 : 2.0) * max_a;
    const double den = (vel_diff > 0) ? max_v : -min_v;
    const double kp = num / den;
    accl = kp * vel_diff;
}

// base_classes.py:282-404: pid -> RK4 (or Euler) with the control held constant -> single-shot yaw wrap
__device__ __forceinline__ void integrate_tick(double st[7], double steer_cmd, double speed_cmd,
                                               const double *__restrict__ p, double dt, int integrator) {
    double accl, sv;
    pid(speed_cmd, steer_cmd, st[3], st[2], p[P_SVMAX], p[P_AMAX], p[P_VMAX], p[P_VMIN], accl, sv);
    if (integrator == 1) {
        double k1[7], k2[7], k3[7], k4[7], xs[7];
        vehicle_dynamics_st(st, sv, accl, p, k1);
#pragma unroll
        for (int k = 0; k < 7; k++) xs[k] = st[k] + dt * (k1[k] / 2);
        vehicle_dynamics_st(xs, sv, accl, p, k2);
#pragma unroll
        for (int k = 0; k < 7; k++) xs[k] = st[k] + dt * (k2[k] / 2);
        vehicle_dynamics_st(xs, sv, accl, p, k3);
#pragma unroll
        for (int k = 0; k < 7; k++) xs[k] = st[k] + dt * k3[k];
        vehicle_dynamics_st(xs, sv, accl, p, k4);
        double w = dt * (1. / 6);   // `self.time_step*(1/6)` is a scalar product evaluated first (:373)
#pragma unroll
        for (int k = 0; k < 7; k++) st[k] = st[k] + w * (((k1[k] + 2 * k2[k]) + 2 * k3[k]) + k4[k]);
    } else {
        double f[7];
        vehicle_dynamics_st(st, sv, accl, p, f);
#pragma unroll
        for (int k = 0; k < 7; k++) st[k] = st[k] + dt * f[k];
    }
    const double two_pi = 2 * M_PI;
    if (st[4] > two_pi) st[4] = st[4] - two_pi;
    else if (st[4] < 0) st[4] = st[4] + two_pi;
}

}  // namespace f110
