/* emu_main.cpp — runs tick_env() of the CUDA kernel header on the host, one emulated warp per env. */
#include "emu_cuda.h"

#include <string>

#include "../../tsid_control_b200/csrc/tsidb_host_const.h"
static DevConst g_const[TSIDB_MAX_SLOTS];
#include "../../tsid_control_b200/csrc/tsidb_kernels.cuh"
#include "../../tsid_control_b200/csrc/tsidb_gait.cuh"

namespace emu { Warp W; }

struct Job { const DevConst* C; double* sm; const TickArgs* a; int env; int stage; };
static double g_mdl[MDL_SIZE];   /* the CTA-shared copies the kernels keep in shared memory */
static double g_lfinv[144];
static Job g_job;

static void lane_entry(int lane) {
  if (g_job.stage == 0) {
    if (g_job.C->nv == 26) dynamics_env<26>(*g_job.C, g_mdl, g_job.sm, *g_job.a, g_job.env, g_job.env, lane);
    else dynamics_env<24>(*g_job.C, g_mdl, g_job.sm, *g_job.a, g_job.env, g_job.env, lane);
  } else if (g_job.stage == 1) {
    unsigned parity = 0;
    const int m = g_job.a->mask ? (((const uint8_t*)g_job.a->mask)[g_job.env] & 3) : 3;
    const int nc = (m & 1) + ((m >> 1) & 1);
    if (g_job.C->nv == 26) {
      if (nc == 2) eliminate_env<26, 2>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
      else if (nc == 1) eliminate_env<26, 1>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
      else eliminate_env<26, 0>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
    } else {
      if (nc == 2) eliminate_env<24, 2>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
      else if (nc == 1) eliminate_env<24, 1>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
      else eliminate_env<24, 0>(*g_job.C, g_lfinv, g_job.sm, *g_job.a, g_job.env, lane, parity);
    }
  } else {
    unsigned parity = 0;
    LaneConst K;
    lane_const_init(*g_job.C, K, lane);
    const int m = g_job.a->mask ? (((const uint8_t*)g_job.a->mask)[g_job.env] & 3) : 3;
    const int nc = (m & 1) + ((m >> 1) & 1);
    if (g_job.C->nv == 26) {
      if (nc == 2) activeset_env<26, 2>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
      else if (nc == 1) activeset_env<26, 1>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
      else activeset_env<26, 0>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
    } else {
      if (nc == 2) activeset_env<24, 2>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
      else if (nc == 1) activeset_env<24, 1>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
      else activeset_env<24, 0>(*g_job.C, K, g_job.sm, *g_job.a, g_job.env, g_job.env, lane, parity);
    }
  }
  emu::W.done[lane] = true;
  swapcontext(&emu::W.ctx[lane], &emu::W.sched);
}

extern "C" int emu_fill_const(const tsidb_model* m, const tsidb_conf* c, const double* refs /*9+24+24+12+12+23*/) {
  std::string err;
  if (!tsidb_fill_devconst(m, c, &g_const[0], &err)) { fprintf(stderr, "emu: %s\n", err.c_str()); return -1; }
  if (refs) {
    DevConst& D = g_const[0];
    memcpy(D.ref_com, refs, 9 * 8);
    memcpy(D.ref_foot[0], refs + 9, 24 * 8);
    memcpy(D.ref_foot[1], refs + 33, 24 * 8);
    memcpy(D.ref_contact[0], refs + 57, 12 * 8);
    memcpy(D.ref_contact[1], refs + 69, 12 * 8);
    memcpy(D.ref_posture, refs + 81, 23 * 8);
  }
  return 0;
}

/* run envs [0, n_envs) sequentially; also returns a copy of the shared-memory image of the LAST env */
static int run_warp(int env) {
  const size_t STK = 1 << 20;
  emu::Warp& W = emu::W;
  W.arrived = 0;
  for (int l = 0; l < 32; l++) {
    W.done[l] = false;
    getcontext(&W.ctx[l]);
    W.ctx[l].uc_stack.ss_sp = W.stacks + l * STK;
    W.ctx[l].uc_stack.ss_size = STK;
    W.ctx[l].uc_link = &W.sched;
    makecontext(&W.ctx[l], (void (*)())lane_entry, 1, l);
  }
  long spins = 0;
  for (;;) {
    bool all = true;
    for (int l = 0; l < 32; l++) {
      if (W.done[l]) continue;
      all = false;
      W.cur = l;
      swapcontext(&W.sched, &W.ctx[l]);
    }
    if (all) break;
    if (++spins > 3000000L) { fprintf(stderr, "emu: env %d: lanes diverged at a collective (deadlock)\n", env); return -2; }
  }
  if (W.arrived != 0) { fprintf(stderr, "emu: env %d: %d lanes left waiting at a collective\n", env, W.arrived); return -3; }
  return 0;
}

/* run envs [0, n_envs) sequentially through both kernels' per-env bodies (prepare, then active set);
 * also returns a copy of the active-set shared-memory image of the LAST env */
extern "C" int emu_tick(const TickArgs* a_in, double* sm_out) {
  const size_t STK = 1 << 20;
  emu::Warp& W = emu::W;
  if (!W.stacks) W.stacks = (char*)malloc(32 * STK);
  TickArgs a = *a_in;
  stage_model(g_const[0], g_mdl, 0, 1);
  for (int k = 0; k < 144; k++) g_lfinv[k] = g_const[0].Lfinv[k / 12][k % 12];
  const int smn = a_layout(TSIDB_NVX, 2).per_env > e_per_env(TSIDB_NVX, 2) ? a_layout(TSIDB_NVX, 2).per_env : e_per_env(TSIDB_NVX, 2);
  double* sm = (double*)calloc(smn, sizeof(double));
  double* ws = (double*)calloc((size_t)a.n_envs * SA_IMAGE, sizeof(double));
  a.ws = ws;
  double* ws3 = (double*)calloc((size_t)a.n_envs * SE_IMAGE, sizeof(double));
  a.ws3 = ws3;
  a.perm = nullptr;
  int rc = 0;
  for (int env = 0; env < a.n_envs && rc == 0; env++) {
    for (int stage = 0; stage < (a.kin_only ? 1 : 3) && rc == 0; stage++) {
      for (int k = 0; k < smn; k++) sm[k] = NAN; /* poison: catches reads of unwritten smem */
      g_job = Job{&g_const[0], sm, &a, env, stage};
      rc = run_warp(env);
    }
  }
  if (sm_out) memcpy(sm_out, sm, smn * sizeof(double));
  free(sm);
  free(ws);
  free(ws3);
  return rc;
}
extern "C" int emu_sm_per_env() { return a_layout(TSIDB_NVX, 2).per_env > e_per_env(TSIDB_NVX, 2) ? a_layout(TSIDB_NVX, 2).per_env : e_per_env(TSIDB_NVX, 2); }

/* the gait phase machine of tsidb_gait.cuh, one env after the other: reset when defaults81 is given, else one step */
extern "C" int emu_gait(int n, const double* gconf6, double* phi, uint8_t* mask, double* vcmd, double* lipm, double* origin,
                        double* com, double* foot_lf, double* foot_rf, double* contact_lf, double* contact_rf, int32_t* fails,
                        const double* defaults81, const double* phase0, const double* vcmd0, const double* foot_now_lf,
                        const double* foot_now_rf, const int32_t* status, const double* steps, const int32_t* n_steps,
                        int32_t* step_idx, double* swing, int max_steps, double rise_ratio) {
  GaitConf G{gconf6[0], gconf6[1], gconf6[2], gconf6[3], gconf6[4], gconf6[5], rise_ratio, max_steps, 0};
  GaitState S{phi, mask, vcmd, lipm, origin, com, {foot_lf, foot_rf}, {contact_lf, contact_rf}, fails, steps, n_steps, step_idx, swing};
  for (int e = 0; e < n; e++) {
    if (defaults81) {
      gait_reset_env(G, S, defaults81, defaults81 + 9, defaults81 + 33, defaults81 + 57, defaults81 + 69, phase0, vcmd0, e);
      if (steps) gait_plan_init_env(G, S, e); /* tsidb_gait_set_plan right after the reset */
    }
    else gait_step_env(G, S, foot_now_lf, foot_now_rf, status, e);
  }
  return 0;
}

/* the planners of tsidb_gait.cuh on the host, env after env (same functions the device kernels call) */
extern "C" int emu_foot_trajectory(int n, double t0, double t1, const double* start4, const double* target4, double h, double rr,
                                   const double* t, double* out16) {
  for (int e = 0; e < n; e++) foot_trajectory_eval(t0, t1, start4 + 4 * e, target4 + 4 * e, h, rr, t[e], out16 + 16 * e);
  return 0;
}
extern "C" int emu_footstep_plan(int n, const double* path, const int32_t* n_pts, int max_pts, const double* init8, double L, double W,
                                 double* steps, int32_t* n_steps, int max_steps) {
  for (int e = 0; e < n; e++)
    n_steps[e] = footstep_plan_env(path + 2 * (size_t)max_pts * e, n_pts ? n_pts[e] : max_pts, init8 + 8 * (size_t)e, L, W,
                                   steps + 4 * (size_t)max_steps * e, max_steps);
  return 0;
}
