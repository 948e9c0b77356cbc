// Env hooks on the device: check_done / check_violation / get_constraint_values for one state row.
// The reference evaluates these in numpy: fp64 arithmetic on fp32 states (point-robot, BoundedConstraint) or
// fp32 arithmetic with an fp64 last step (tracking).  Every operation below uses an explicit round-to-nearest
// intrinsic so that nvcc cannot contract a*b+c into an FMA: results are bit-identical to numpy's.
#pragma once
#include "common.cuh"

namespace drpo {

struct HookOut {
  bool done, viol;
  float cv[DRPO_MAX_CON];
};

__device__ __forceinline__ double np_minimum(double a, double b) {   // np.minimum propagates NaN
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}
__device__ __forceinline__ float np_minimum_f(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : (a < b ? a : b));
}
__device__ __forceinline__ double norm2_d(double dx, double dy) {    // np.linalg.norm(axis=1) on a 2-vector
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
__device__ __forceinline__ float norm2_f(float dx, float dy) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// s: pointer to the row's state_dim floats, element stride `st` (1 for row-major rows)
template <typename Load>
__device__ __forceinline__ void eval_hooks(const drpo_env_params& p, Load s, HookOut& o) {
  if (p.kind == DRPO_ENV_POINT_ROBOT) {
    // src/env/point_robot.py:96-130
    const float x32 = s(0), y32 = s(1);
    const double x = (double)x32, y = (double)y32;
    double mind = INFINITY;
    for (int h = 0; h < p.n_hazards; ++h) {
      double d = norm2_d(__dsub_rn(p.hazard_xy[h][0], x), __dsub_rn(p.hazard_xy[h][1], y));
      mind = np_minimum(d, mind);
    }
    const double cv = __dsub_rn(p.hazard_size, mind);
    o.viol = cv > 0.0;
    const float b = p.xy_bound;
    const bool oob = (x32 < -b) || (x32 > b) || (y32 < -b) || (y32 > b);
    const bool goal = norm2_d(__dsub_rn(x, p.goal_xy[0]), __dsub_rn(y, p.goal_xy[1])) <= p.goal_size;
    o.done = oob || goal;
    o.cv[0] = __double2float_rn(cv);
  } else if (p.kind == DRPO_ENV_BOUNDED) {
    // src/env/poles/constraints.py:203-204: x @ filter.T @ A.T - b with fp64 filter/A/b.  The 0/1 filter multiplies
    // EVERY state dim, so a non-finite value in another dim turns y into NaN (0*inf); we reproduce that.
    // (loops are unrolled to their compile-time maxima with predicates so that yv[] / o.cv[] stay in registers)
    int n_nonfinite = 0;
    for (int d = 0; d < p.state_dim; ++d) n_nonfinite += !isfinite(s(d));
    double yv[DRPO_MAX_ACTIVE];
    const int na = p.n_active;
#pragma unroll
    for (int a = 0; a < DRPO_MAX_ACTIVE; ++a) {
      yv[a] = 0.0;
      if (a < na) {
        const float xa = s(p.active_dims[a]);
        const int others = n_nonfinite - (!isfinite(xa) ? 1 : 0);
        yv[a] = others > 0 ? (double)NAN : (double)xa;
      }
    }
    bool viol = false;
#pragma unroll
    for (int c = 0; c < 2 * DRPO_MAX_ACTIVE; ++c) {
      if (c < 2 * na) {
        // row c of A = [-I; I]; literal sum_a y[a]*A[c][a] so NaN/inf propagate exactly as in the matmul
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < DRPO_MAX_ACTIVE; ++a) {
          if (a < na) {
            const double coef = (c < na) ? ((a == c) ? -1.0 : 0.0) : ((a == c - na) ? 1.0 : 0.0);
            acc = __dadd_rn(acc, __dmul_rn(yv[a], coef));
          }
        }
        double bb = 0.0;
#pragma unroll
        for (int a = 0; a < DRPO_MAX_ACTIVE; ++a) { if (c < na && a == c) bb = -p.lower[a]; if (c >= na && a == c - na) bb = p.upper[a]; }
        const double cv = __dsub_rn(acc, bb);
        viol = viol || (cv > 0.0);
        o.cv[c] = __double2float_rn(cv);
      }
    }
    bool done = viol;                         // inverted_pendulum.py:79-82 ; quadrotor.py:112-114
    for (int j = 0; j < p.n_done_dims; ++j) {
      const float v = s(p.done_dims[j]), t = p.done_thr[j];
      done = done || (v < -t) || (v > t);     // fp32 compare against the fp32-rounded threshold (NEP 50 weak scalar)
    }
    o.viol = viol; o.done = done;
  } else {
    // src/env/tracking/pyth_veh3dofconti_surrcstr_data.py:253-338 — fp32 until the final 2r - min_dist
    o.done = (fabsf(s(0)) > 5.f) || (fabsf(s(1)) > 2.f) || (fabsf(s(2)) > 3.14159274101257324f);
    const float d = (float)((p.veh_length - p.veh_width) / 2.0);
    const double two_r = 2.0 * (sqrt(2.0) / 2.0 * p.veh_width);
    const float phi = s(6);
    const float c = cosf(phi), sn = sinf(phi);
    float mind = INFINITY;
    for (int v = 0; v < p.surr_veh_num; ++v) {
      const int b = p.surr_start + 4 * v;
      const float xs = s(b), ys = s(b + 1), ph = s(b + 2);
      const float xe = __fadd_rn(__fmul_rn(xs, c), __fmul_rn(ys, sn));
      const float ye = __fadd_rn(__fmul_rn(-xs, sn), __fmul_rn(ys, c));
      const float dc = __fmul_rn(d, cosf(ph)), ds = __fmul_rn(d, sinf(ph));
      const float cx0 = __fadd_rn(xe, dc), cy0 = __fadd_rn(ye, ds);
      const float cx1 = __fsub_rn(xe, dc), cy1 = __fsub_rn(ye, ds);
      float m = norm2_f(__fsub_rn(d, cx0), __fsub_rn(0.f, cy0));            // d1
      m = np_minimum_f(m, norm2_f(__fsub_rn(d, cx1), __fsub_rn(0.f, cy1)));  // d2
      m = np_minimum_f(m, norm2_f(__fsub_rn(-d, cx0), __fsub_rn(0.f, cy0))); // d3
      m = np_minimum_f(m, norm2_f(__fsub_rn(-d, cx1), __fsub_rn(0.f, cy1))); // d4
      mind = np_minimum_f(mind, m);
    }
    const double cv = __dsub_rn(two_r, (double)mind);
    o.viol = cv > 0.0;
    o.cv[0] = __double2float_rn(cv);
  }
}

}  // namespace drpo
