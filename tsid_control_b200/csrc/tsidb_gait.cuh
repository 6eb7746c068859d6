/* tsidb_gait.cuh — per-env gait phase machine and reference generator on the device (one thread per env).
 *
 * The reference keeps this logic on the host and never wires it into its loop (`update_tasks` is commented
 * out at ref:main.py:117); the pieces it is made of are restated here so that closed-loop batched rollouts
 * need no host round trip:
 *   - contact switching with the legacy semantics: on lift-off the foot-task reference becomes the current
 *     foot placement, on touch-down the contact reference becomes the current placement
 *     (ref:legacy/biped.py:168-212; the WalkController versions at ref:ctrl/WalkController.py:215-253 raise);
 *   - swing foot: x linear in time, z the parabola through (0,0), (T/2,h), (T,0) — the 2-knot and 3-knot
 *     CubicSplines of ref:ctrl/Foot_Trajectory.py:8-19 with rise_ratio 0.5 — with first and second derivatives
 *     as velocity and acceleration references;
 *   - CoM: one semi-implicit Euler step of the linear inverted pendulum per tick,
 *     acc = (zmp - pos) w^2; vel += acc dt; pos += vel dt  (ref:ctrl/LIPM.py:44-47, w^2 = 9.80665/h0 at :15),
 *     with the ZMP at the centre of the support (both contact references in double support, the stance one
 *     in single support);
 *   - gait cycle as in the benchmark workload: phase in [0,1): [0,0.2) double support,
 *     [0.2,0.6) left foot in contact / right foot swinging, [0.6,1) right in contact / left swinging; one
 *     swing lasts step_duration, so the phase advances by 0.4 dt / step_duration per tick.
 * tests/gait_ref.py is the numpy restatement these functions are checked against.
 * Compiles for the host under tests/emu (TSIDB_EMU). */
#ifndef TSIDB_GAIT_CUH_
#define TSIDB_GAIT_CUH_
#include <stdint.h>

#ifndef TSIDB_DEV
#define TSIDB_DEV __device__ __forceinline__
#endif

struct GaitConf {
  double dt, step_duration, step_length, step_height, w2, com_z;
};

/* device arrays, [N][k] row-major */
struct GaitState {
  double* phi;        /* [N]       gait phase                                   */
  uint8_t* mask;      /* [N]       contact mask of the tick about to run        */
  double* vcmd;       /* [N][2]    velocity command                             */
  double* lipm;       /* [N][4]    pendulum position xy, velocity xy            */
  double* origin;     /* [N][2][12] placement of each foot at its last lift-off */
  double* com;        /* [N][9]    references handed to the tick ...            */
  double* foot[2];    /* [N][24]                                                */
  double* contact[2]; /* [N][12]                                                */
  int32_t* fails;     /* [N]       ticks whose QP did not reach status 0        */
};

TSIDB_DEV int gait_mask_of(double phi) { return phi < 0.2 ? 3 : (phi < 0.6 ? 1 : 2); }

/* state of env e at reset: standing references, pendulum at the standing CoM moving with the command */
TSIDB_DEV void gait_reset_env(const GaitConf& G, const GaitState& S, const double* com9, const double* foot_lf24,
                              const double* foot_rf24, const double* contact_lf12, const double* contact_rf12,
                              const double* phase0, const double* vcmd, int e) {
  const double phi = phase0 ? phase0[e] : 0.0;
  S.phi[e] = phi;
  S.mask[e] = (uint8_t)gait_mask_of(phi);
  const double vx = vcmd ? vcmd[2 * e] : 0.0, vy = vcmd ? vcmd[2 * e + 1] : 0.0;
  S.vcmd[2 * e] = vx; S.vcmd[2 * e + 1] = vy;
  S.lipm[4 * e] = com9[0]; S.lipm[4 * e + 1] = com9[1]; S.lipm[4 * e + 2] = vx; S.lipm[4 * e + 3] = vy;
  for (int k = 0; k < 9; k++) S.com[9 * e + k] = com9[k];
  S.com[9 * e + 3] = vx; S.com[9 * e + 4] = vy;
  for (int k = 0; k < 24; k++) { S.foot[0][24 * e + k] = foot_lf24[k]; S.foot[1][24 * e + k] = foot_rf24[k]; }
  for (int k = 0; k < 12; k++) {
    S.contact[0][12 * e + k] = contact_lf12[k]; S.contact[1][12 * e + k] = contact_rf12[k];
    S.origin[24 * e + k] = foot_lf24[k]; S.origin[24 * e + 12 + k] = foot_rf24[k];
  }
  if (S.fails) S.fails[e] = 0;
}

/* one tick of the phase machine for env e; foot_now[f] = placements (p, R column-major) measured by the tick
 * that just ran, status = its QP status (may be null) */
TSIDB_DEV void gait_step_env(const GaitConf& G, const GaitState& S, const double* foot_now_lf, const double* foot_now_rf,
                             const int32_t* status, int e) {
  double phi = S.phi[e] + 0.4 * G.dt / G.step_duration;
  if (phi >= 1.0) phi -= 1.0;
  const int old = S.mask[e];
  const int nm = gait_mask_of(phi);
  const double T = G.step_duration, h = G.step_height;
  const double L = (S.vcmd[2 * e] >= 0.0) ? G.step_length : -G.step_length;
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const int bit = 1 << f;
    const double* now = (f == 0 ? foot_now_lf : foot_now_rf) + 12 * (size_t)e;
    double* org = S.origin + 24 * (size_t)e + 12 * f;
    double* fr = S.foot[f] + 24 * (size_t)e;
    double* cr = S.contact[f] + 12 * (size_t)e;
    if ((old & bit) && !(nm & bit)) {
      /* lift-off [Biped.remove*FootContact]: the swing starts from where the foot is */
      for (int k = 0; k < 12; k++) org[k] = now[k];
    }
    if (!(old & bit) && (nm & bit)) {
      /* touch-down [Biped.add*FootContact]: contact (and foot task) reference := current placement */
      for (int k = 0; k < 12; k++) { cr[k] = now[k]; fr[k] = now[k]; }
      for (int k = 12; k < 24; k++) fr[k] = 0.0;
    }
    if (!(nm & bit)) {
      /* swinging [FootTrajectory]: right foot swings in [0.2,0.6), left foot in [0.6,1) */
      const double s = (phi - (f == 1 ? 0.2 : 0.6)) / 0.4;
      fr[0] = org[0] + L * s;
      fr[1] = org[1];
      fr[2] = org[2] + 4.0 * h * s * (1.0 - s);
      for (int k = 3; k < 12; k++) fr[k] = org[k];
      for (int k = 12; k < 24; k++) fr[k] = 0.0;
      fr[12] = L / T;
      fr[14] = 4.0 * h * (1.0 - 2.0 * s) / T;
      fr[20] = -8.0 * h / (T * T);
    }
  }
  /* LIPM step towards the centre of the support */
  const double* cl = S.contact[0] + 12 * (size_t)e;
  const double* cr_ = S.contact[1] + 12 * (size_t)e;
  double zx, zy;
  if (nm == 3) { zx = 0.5 * (cl[0] + cr_[0]); zy = 0.5 * (cl[1] + cr_[1]); }
  else if (nm == 1) { zx = cl[0]; zy = cl[1]; }
  else { zx = cr_[0]; zy = cr_[1]; }
  double* lp = S.lipm + 4 * (size_t)e;
  const double ax = (zx - lp[0]) * G.w2, ay = (zy - lp[1]) * G.w2;
  lp[2] += ax * G.dt; lp[3] += ay * G.dt;
  lp[0] += lp[2] * G.dt; lp[1] += lp[3] * G.dt;
  double* c = S.com + 9 * (size_t)e;
  c[0] = lp[0]; c[1] = lp[1]; c[2] = G.com_z;
  c[3] = lp[2]; c[4] = lp[3]; c[5] = 0.0;
  c[6] = ax; c[7] = ay; c[8] = 0.0;
  S.phi[e] = phi;
  S.mask[e] = (uint8_t)nm;
  if (S.fails && status && status[e] != 0) S.fails[e] += 1;
}

/* ---- the reference's planners, one thread per env -------------------------------------------
 * Both are pinned to the reference's own Python through tests/golden/planners.npz (generated by importing
 * ref:ctrl/Foot_Trajectory.py and ref:ctrl/Footstep_Planner.py, tests/golden/make_planner_golden.py). */
#ifdef TSIDB_EMU
#define TS_MUL(a, b) ((a) * (b))
#define TS_ADD(a, b) ((a) + (b))
#else
#define TS_MUL(a, b) __dmul_rn(a, b) /* no FMA contraction where a threshold decides (numpy rounds every product) */
#define TS_ADD(a, b) __dadd_rn(a, b)
#endif

/* FootTrajectory (ref:ctrl/Foot_Trajectory.py:6-43) at time t: x, y, yaw are 2-knot CubicSplines = straight lines;
 * z is the spline through (t0, z0), (t0 + T/2, z0 + h), (t1, z1) when rise_ratio == 0.5 and through
 * (t0, z0), (t0 + r T, z0 + h), (t1 - r T, z1 + h), (t1, z1) otherwise; with scipy's not-a-knot ends a spline through
 * <= 4 knots is ONE polynomial, evaluated here from Newton divided differences.  start/target = (x, y, z, yaw).
 * out[16] = value, 1st, 2nd, 3rd derivative of (x, y, z, yaw).  (The reference's get_velocity / get_acceleration
 * return the 2nd and 3rd derivative, :35,:43; a TSID foot reference needs the 1st and 2nd — all four are here.) */
TSIDB_DEV void foot_trajectory_eval(double t0, double t1, const double* start, const double* target, double h, double rr,
                                    double t, double* out) {
  const double T = t1 - t0, u = t - t0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (k == 2) continue;
    const double c1 = (target[k] - start[k]) / T;
    out[k] = start[k] + c1 * u; out[4 + k] = c1; out[8 + k] = 0.0; out[12 + k] = 0.0;
  }
  double a0, a1, a2, a3;
  if (rr != 0.5) {
    const double u1 = T * rr, u2 = T - T * rr, u3 = T;
    const double y0 = start[2], y1 = start[2] + h, y2 = target[2] + h, y3 = target[2];
    const double d01 = (y1 - y0) / u1, d12 = (y2 - y1) / (u2 - u1), d23 = (y3 - y2) / (u3 - u2);
    const double d012 = (d12 - d01) / u2, d123 = (d23 - d12) / (u3 - u1);
    const double d0123 = (d123 - d012) / u3;
    /* c0 + c1 u + c2 u (u - u1) + c3 u (u - u1) (u - u2) in powers of u */
    a0 = y0; a1 = d01 - d012 * u1 + d0123 * u1 * u2; a2 = d012 - d0123 * (u1 + u2); a3 = d0123;
  } else {
    const double u1 = T * rr, u2 = T;
    const double y0 = start[2], y1 = start[2] + h, y2 = target[2];
    const double d01 = (y1 - y0) / u1, d12 = (y2 - y1) / (u2 - u1);
    const double d012 = (d12 - d01) / u2;
    a0 = y0; a1 = d01 - d012 * u1; a2 = d012; a3 = 0.0;
  }
  out[2] = ((a3 * u + a2) * u + a1) * u + a0;
  out[6] = (3.0 * a3 * u + 2.0 * a2) * u + a1;
  out[10] = 6.0 * a3 * u + 2.0 * a2;
  out[14] = 6.0 * a3;
}

/* FootstepPlanner.add_step (ref:ctrl/Footstep_Planner.py:74-90): out = (x, y, yaw, side) */
TSIDB_DEV void footstep_add(double dx, double dy, int side, double px, double py, double L, double W, double* out) {
  const double nrm = sqrt(TS_ADD(TS_MUL(dx, dx), TS_MUL(dy, dy)));
  const double tx = dx / nrm, ty = dy / nrm;
  const double sw = W / 2 * (side == 0 ? 1.0 : -1.0);
  out[0] = TS_ADD(TS_ADD(px, TS_MUL(tx, L / 2)), TS_MUL(-ty, sw));
  out[1] = TS_ADD(TS_ADD(py, TS_MUL(ty, L / 2)), TS_MUL(tx, sw));
  out[2] = atan2(dy, dx);
  out[3] = (double)side;
}
/* FootstepPlanner.plan (ref:ctrl/Footstep_Planner.py:92-125) for one env: path [n_pts][2], init [2][4] the two initial
 * supports (x, y, yaw, side); steps [max_steps][4]; returns the number of footsteps (the initial supports included),
 * or -1 when max_steps is too small. */
TSIDB_DEV int footstep_plan_env(const double* path, int n_pts, const double* init, double L, double W, double* steps, int max_steps) {
  if (max_steps < 2 || n_pts < 2) return -1;
  for (int k = 0; k < 8; k++) steps[k] = init[k];
  int ns = 2;
  int side = (int)init[7];
  double distance = 0.0, dx = 0.0, dy = 0.0;
  for (int i = 0; i < n_pts - 1; i++) {
    dx = path[2 * (i + 1)] - path[2 * i];
    dy = path[2 * (i + 1) + 1] - path[2 * i + 1];
    distance = TS_ADD(distance, sqrt(TS_ADD(TS_MUL(dx, dx), TS_MUL(dy, dy))));
    if (distance >= L) {
      side = !side;
      if (ns >= max_steps) return -1;
      footstep_add(dx, dy, side, path[2 * i], path[2 * i + 1], L, W, steps + 4 * ns);
      ns++;
      distance = 0.0;
    }
  }
  side = !side;
  if (ns >= max_steps) return -1;
  footstep_add(dx, dy, side, path[2 * (n_pts - 1)], path[2 * (n_pts - 1) + 1], L, W, steps + 4 * ns);
  ns++;
  if (distance > 0) {
    side = !side;
    if (ns >= max_steps) return -1;
    footstep_add(dx, dy, side, path[2 * (n_pts - 1)], path[2 * (n_pts - 1) + 1], L, W, steps + 4 * ns);
    ns++;
  }
  return ns;
}

/* per-env diagnostics of one tick from its auxiliary outputs:
 *   cop[3]   centre of pressure, ref:ctrl/WalkController.py:255-289: per foot in contact with f_z > 1e-3 the local
 *            CoP (w[4]/w[2], w[3]/w[2], 0) from the wrench w = T f, mapped to the world by the sole placement,
 *            then the f_z-weighted mean over the feet in contact.  (The reference returns None unless both feet
 *            are in contact because it reads f_rf unconditionally; a single contact gives that foot's CoP here,
 *            no contact gives zeros.)
 *   cp[3]    capture point com + vcom / w with z = 0, ref:legacy/biped.py:224-227
 *   poly[4]  support "polygon" (lf.xy, rf.xy) of ref:legacy/biped.py:229-234 */
TSIDB_DEV void diagnostics_env(const double* com9, const double* foot_lf12, const double* foot_rf12, const double* wrench12,
                               const uint8_t* mask, double w, double* cop, double* cp, double* poly, int e) {
  const int m = mask ? (mask[e] & 3) : 3;
  double nx = 0.0, ny = 0.0, den = 0.0;
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const double* W = wrench12 + 12 * (size_t)e + 6 * f;
    const double* P = (f == 0 ? foot_lf12 : foot_rf12) + 12 * (size_t)e; /* p, R column-major */
    const double fz = W[2];
    double lx = 0.0, ly = 0.0;
    if (fz > 1e-3) { lx = W[4] / fz; ly = W[3] / fz; }
    /* world = R (lx, ly, 0) + p; R column-major: R[r][c] = P[3 + 3 c + r] */
    const double wx = P[3] * lx + P[6] * ly + P[0];
    const double wy = P[4] * lx + P[7] * ly + P[1];
    if ((m >> f) & 1) { nx += wx * fz; ny += wy * fz; den += fz; }
  }
  if (cop) {
    cop[3 * (size_t)e] = den != 0.0 ? nx / den : 0.0;
    cop[3 * (size_t)e + 1] = den != 0.0 ? ny / den : 0.0;
    cop[3 * (size_t)e + 2] = 0.0;
  }
  if (cp) {
    const double* c = com9 + 9 * (size_t)e;
    cp[3 * (size_t)e] = c[0] + c[3] / w;
    cp[3 * (size_t)e + 1] = c[1] + c[4] / w;
    cp[3 * (size_t)e + 2] = 0.0;
  }
  if (poly) {
    poly[4 * (size_t)e] = foot_lf12[12 * (size_t)e]; poly[4 * (size_t)e + 1] = foot_lf12[12 * (size_t)e + 1];
    poly[4 * (size_t)e + 2] = foot_rf12[12 * (size_t)e]; poly[4 * (size_t)e + 3] = foot_rf12[12 * (size_t)e + 1];
  }
}

#ifndef TSIDB_EMU
__global__ void tsidb_diagnostics_kernel(int n, const double* com9, const double* foot_lf12, const double* foot_rf12,
                                         const double* wrench12, const uint8_t* mask, double w, double* cop, double* cp, double* poly) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  diagnostics_env(com9, foot_lf12, foot_rf12, wrench12, mask, w, cop, cp, poly, e);
}
__global__ void tsidb_foot_trajectory_kernel(int n, double t0, double t1, const double* start4, const double* target4, double h,
                                            double rr, const double* t, double* out16) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  foot_trajectory_eval(t0, t1, start4 + 4 * (size_t)e, target4 + 4 * (size_t)e, h, rr, t[e], out16 + 16 * (size_t)e);
}
__global__ void tsidb_footstep_plan_kernel(int n, const double* path, const int32_t* n_pts, int max_pts, const double* init8,
                                           double L, double W, double* steps, int32_t* n_steps, int max_steps) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  n_steps[e] = footstep_plan_env(path + 2 * (size_t)max_pts * e, n_pts ? n_pts[e] : max_pts, init8 + 8 * (size_t)e, L, W,
                                 steps + 4 * (size_t)max_steps * e, max_steps);
}
__global__ void tsidb_gait_reset_kernel(int n, GaitConf G, GaitState S, const double* defaults /* 9+24+24+12+12 */,
                                        const double* phase0, const double* vcmd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  gait_reset_env(G, S, defaults, defaults + 9, defaults + 33, defaults + 57, defaults + 69, phase0, vcmd, e);
}
__global__ void tsidb_gait_step_kernel(int n, GaitConf G, GaitState S, const double* foot_now_lf, const double* foot_now_rf,
                                       const int32_t* status) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  gait_step_env(G, S, foot_now_lf, foot_now_rf, status, e);
}
#endif
#endif
