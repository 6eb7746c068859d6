/*
 * tsid_oracle.c — CPU restatement of the reference's per-tick hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * libtsidb.so (the product) never does.
 *
 * PARITY UNPINNED.  The arithmetic of the reference's tick
 * (ref:main.py:119-127) lives in three third-party libraries that are not
 * vendored, not pinned (the reference has no requirements/lock file) and not
 * installable in this image: `tsid` (stack-of-tasks/tsid, API level >= 1.2),
 * `pinocchio` (2.x semantics restated here) and `eiquadprog`
 * (eiquadprog-fast.hpp).  The reference holds no tests or golden vectors for
 * this path.  This file restates the published
 * algorithms of those libraries, following the reference's own call sites for
 * problem structure and constants.  It is validated by physical identities
 * and an independent KKT check in tests/, not by reference outputs.
 *
 * What follows what:
 *   ot_dynamics()     pinocchio::computeAllTerms + updateFramePlacements +
 *                     centerOfMass(q,v,0) + ccrba, as sequenced by
 *                     tsid::RobotWrapper::computeAllTerms — implicit in
 *                     formulation.computeProblemData, ref:main.py:119
 *   ot_assemble()     tsid task/contact compute() + InverseDynamicsFormulationAccForce
 *                     ::computeProblemData stacking — problem defined at
 *                     ref:ctrl/WalkController.py:55-187, ref:legacy/biped.py:31-130
 *   ot_solve()        tsid::SolverHQuadProgFast::solve + eiquadprog-fast
 *                     solve_quadprog — ref:main.py:121, ref:ctrl/WalkController.py:186-187
 *   ot_decode()       getAccelerations / getActuatorForces / getContactForce —
 *                     ref:main.py:126-127, ref:ctrl/WalkController.py:263,273
 *   ot_integrate()    integrate_dv — ref:ctrl/WalkController.py:291-295
 *
 * Build: see oracle/Makefile.  -DREAL="long double" gives the 80-bit "truth"
 * variant used for conditioning studies.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stdio.h>

#include "../include/tsidb.h"
#include "tsid_oracle.h"

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

#ifdef ORACLE_LONG_DOUBLE
#define R_SQRT sqrtl
#define R_SIN sinl
#define R_COS cosl
#define R_ACOS acosl
#define R_FABS fabsl
#define R_EPS LDBL_EPSILON
#else
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_ACOS acos
#define R_FABS fabs
#define R_EPS DBL_EPSILON
#endif
/* decisions of the solver (tolerances) always use the fp64 epsilon so that the
 * long-double build follows the same pivot rules */
#define QP_EPS DBL_EPSILON

#define NBMAX TSIDB_MAX_BODIES
#define NVMAX TSIDB_MAX_NV
#define NMAX (TSIDB_MAX_NV + 24)
#define NEQMAX 18
#define NINMAX (2 * (34 + TSIDB_MAX_NA + TSIDB_MAX_NV))

/* ------------------------------------------------------------------ spatial algebra */
typedef struct { real R[9]; real p[3]; } se3;           /* R row-major */
typedef struct { real lin[3]; real ang[3]; } sv;        /* Motion or Force, [linear; angular] */
typedef struct { real m; real c[3]; real I[9]; } sinertia;

static inline void cross3(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matvec3(const real* R, const real* v, real* o) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matTvec3(const real* R, const real* v, real* o) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  real y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  real z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matmul3(const real* A, const real* B, real* C) {
  real t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, t, sizeof t);
}
static inline void matmulT3(const real* A, const real* B, real* C) { /* A * B^T */
  real t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      t[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
  memcpy(C, t, sizeof t);
}
/* SE3 composition a*b */
static inline void se3_mul(const se3* a, const se3* b, se3* o) {
  se3 t;
  matmul3(a->R, b->R, t.R);
  matvec3(a->R, b->p, t.p);
  for (int k = 0; k < 3; k++) t.p[k] += a->p[k];
  *o = t;
}
/* a^-1 * b */
static inline void se3_inv_mul(const se3* a, const se3* b, se3* o) {
  se3 t;
  real Rt[9] = {a->R[0], a->R[3], a->R[6], a->R[1], a->R[4], a->R[7], a->R[2], a->R[5], a->R[8]};
  matmul3(Rt, b->R, t.R);
  real d[3] = {b->p[0] - a->p[0], b->p[1] - a->p[1], b->p[2] - a->p[2]};
  matvec3(Rt, d, t.p);
  *o = t;
}
/* Motion: M.act(v) */
static inline void se3_act_motion(const se3* M, const sv* v, sv* o) {
  sv t;
  matvec3(M->R, v->ang, t.ang);
  matvec3(M->R, v->lin, t.lin);
  real c[3];
  cross3(M->p, t.ang, c);
  for (int k = 0; k < 3; k++) t.lin[k] += c[k];
  *o = t;
}
/* Motion: M.actInv(v) */
static inline void se3_actinv_motion(const se3* M, const sv* v, sv* o) {
  sv t;
  real c[3], d[3];
  cross3(M->p, v->ang, c);
  for (int k = 0; k < 3; k++) d[k] = v->lin[k] - c[k];
  matTvec3(M->R, d, t.lin);
  matTvec3(M->R, v->ang, t.ang);
  *o = t;
}
/* Force: M.act(f) */
static inline void se3_act_force(const se3* M, const sv* f, sv* o) {
  sv t;
  matvec3(M->R, f->lin, t.lin);
  matvec3(M->R, f->ang, t.ang);
  real c[3];
  cross3(M->p, t.lin, c);
  for (int k = 0; k < 3; k++) t.ang[k] += c[k];
  *o = t;
}
/* Motion x Motion */
static inline void motion_cross(const sv* a, const sv* b, sv* o) {
  sv t;
  real c1[3], c2[3];
  cross3(a->lin, b->ang, c1);
  cross3(a->ang, b->lin, c2);
  for (int k = 0; k < 3; k++) t.lin[k] = c1[k] + c2[k];
  cross3(a->ang, b->ang, t.ang);
  *o = t;
}
/* Motion x* Force */
static inline void motion_cross_force(const sv* v, const sv* f, sv* o) {
  sv t;
  real c1[3], c2[3];
  cross3(v->ang, f->lin, t.lin);
  cross3(v->ang, f->ang, c1);
  cross3(v->lin, f->lin, c2);
  for (int k = 0; k < 3; k++) t.ang[k] = c1[k] + c2[k];
  *o = t;
}
/* Inertia * Motion (pinocchio InertiaTpl::__mult__) */
static inline void inertia_mul(const sinertia* Y, const sv* v, sv* f) {
  sv t;
  real c[3];
  cross3(Y->c, v->ang, c);
  for (int k = 0; k < 3; k++) t.lin[k] = Y->m * (v->lin[k] - c[k]);
  matvec3(Y->I, v->ang, t.ang);
  cross3(Y->c, t.lin, c);
  for (int k = 0; k < 3; k++) t.ang[k] += c[k];
  *f = t;
}
/* M.act(Y) */
static inline void inertia_se3_act(const se3* M, const sinertia* Y, sinertia* o) {
  sinertia t;
  t.m = Y->m;
  matvec3(M->R, Y->c, t.c);
  for (int k = 0; k < 3; k++) t.c[k] += M->p[k];
  real RI[9];
  matmul3(M->R, Y->I, RI);
  matmulT3(RI, M->R, t.I);
  *o = t;
}
/* pinocchio InertiaTpl::__pequ__ */
static inline void inertia_add(sinertia* a, const sinertia* b) {
  real mab = a->m + b->m;
  if (mab == 0) return;
  real ab[3] = {a->c[0] - b->c[0], a->c[1] - b->c[1], a->c[2] - b->c[2]};
  real k = a->m * b->m / mab;
  /* skew(ab)^2 = ab ab^T - |ab|^2 I */
  real n2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      real s2 = ab[i] * ab[j] - (i == j ? n2 : 0);
      a->I[3 * i + j] = a->I[3 * i + j] + b->I[3 * i + j] - k * s2;
    }
  for (int i = 0; i < 3; i++) a->c[i] = (a->m * a->c[i] + b->m * b->c[i]) / mab;
  a->m = mab;
}

/* Eigen::Quaternion::toRotationMatrix, q = (x,y,z,w) as stored in the configuration vector */
static void quat_to_R(const real* q, real* R) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  real tx = 2 * x, ty = 2 * y, tz = 2 * z;
  real twx = tx * w, twy = ty * w, twz = tz * w;
  real txx = tx * x, txy = ty * x, txz = tz * x;
  real tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

/* pinocchio log3 (2.x, acos based) */
static void log3_(const real* R, real* w, real* theta_out) {
  const real PI = (real)3.14159265358979323846264338327950288L;
  real tr = R[0] + R[4] + R[8];
  real theta;
  if (tr >= 3) theta = 0;
  else if (tr <= -1) theta = PI;
  else theta = R_ACOS((tr - 1) / 2);
  if (theta >= PI - (real)1e-2) {
    real cphi = -(tr - 1) / 2;
    real beta = theta * theta / (1 + cphi);
    real tmp[3] = {(R[0] + cphi) * beta, (R[4] + cphi) * beta, (R[8] + cphi) * beta};
    w[0] = (R[7] > R[5] ? 1 : -1) * (tmp[0] > 0 ? R_SQRT(tmp[0]) : 0);
    w[1] = (R[2] > R[6] ? 1 : -1) * (tmp[1] > 0 ? R_SQRT(tmp[1]) : 0);
    w[2] = (R[3] > R[1] ? 1 : -1) * (tmp[2] > 0 ? R_SQRT(tmp[2]) : 0);
  } else {
    const real prec3 = (real)1.220703125e-4; /* TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4) */
    real t = ((theta > prec3) ? theta / R_SIN(theta) : (real)1) / 2;
    w[0] = t * (R[7] - R[5]);
    w[1] = t * (R[2] - R[6]);
    w[2] = t * (R[3] - R[1]);
  }
  *theta_out = theta;
}
/* pinocchio log6 */
static void log6_(const se3* M, sv* out) {
  real w[3], t;
  log3_(M->R, w, &t);
  real t2 = t * t, alpha, beta;
  const real prec3 = (real)1.220703125e-4;
  if (t < prec3) {
    alpha = 1 - t2 / 12 - t2 * t2 / 720;
    beta = (real)1 / 12 + t2 / 720;
  } else {
    real st = R_SIN(t), ct = R_COS(t);
    alpha = t * st / (2 * (1 - ct));
    beta = 1 / t2 - st / (2 * t * (1 - ct));
  }
  real wxp[3];
  cross3(w, M->p, wxp);
  real wdp = w[0] * M->p[0] + w[1] * M->p[1] + w[2] * M->p[2];
  for (int k = 0; k < 3; k++) {
    out->lin[k] = alpha * M->p[k] - (real)0.5 * wxp[k] + (beta * wdp) * w[k];
    out->ang[k] = w[k];
  }
}

/* ------------------------------------------------------------------ dynamics terms */
typedef struct {
  int nb, nv, na;
  se3 liMi[NBMAX], oMi[NBMAX];
  sv v[NBMAX], a[NBMAX]; /* a: zero-joint-acceleration, no gravity ("drift") */
  real M[NVMAX * NVMAX];
  real nle[NVMAX];
  real J[6 * NVMAX];     /* world joint Jacobian, rows 0..2 linear, 3..5 angular; column-major 6 x nv */
  real com[3], vcom[3], acom[3], mass;
  real Jcom[3 * NVMAX];  /* row-major 3 x nv */
  real Ag[6 * NVMAX];    /* row-major 6 x nv */
  real dAg_v_ang[3];     /* angular part of d(hg)/dt at zero joint acceleration */
  se3 oMf[2];
  sv vF[2], aF[2];       /* LOCAL frame velocity / classic acceleration drift */
  real JF[2][6 * NVMAX]; /* LOCAL frame Jacobian, row-major 6 x nv */
} dyn_t;

/* ---- switches for the [UPSTREAM] details recalled with the least certainty.  Default 0 = the
 * restatement's reading of tsid / pinocchio / eiquadprog.  tests/test_assumptions.py flips them one at a time to
 * measure how much each assumption moves the answer.  Test infrastructure only. */
int g_assume[ORACLE_A_COUNT];
int oracle_set_assumption(int which, int alt) {
  if (which < 0 || which >= ORACLE_A_COUNT) return -1;
  g_assume[which] = alt;
  return 0;
}

static void load_inertia(const tsidb_model* m, int i, sinertia* Y) {
  Y->m = m->mass[i];
  for (int k = 0; k < 3; k++) Y->c[k] = m->com[i][k];
  for (int k = 0; k < 9; k++) Y->I[k] = m->inertia[i][k];
}

/* columns of body i in v-space */
static inline int dof0(int i) { return i == 0 ? 0 : 5 + i; }
static inline int ndof(int i) { return i == 0 ? 6 : 1; }

/* motion subspace column k of body i in the LOCAL joint frame */
static inline void S_col(int i, int k, sv* s) {
  memset(s, 0, sizeof *s);
  if (i == 0) { if (k < 3) s->lin[k] = 1; else s->ang[k - 3] = 1; }
  else s->ang[2] = 1;
}

static void ot_dynamics(const tsidb_model* m, const double* q, const double* v, dyn_t* d) {
  const int nb = m->nb, nv = nb + 5;
  d->nb = nb; d->nv = nv; d->na = nb - 1;
  sv a_gf[NBMAX], f[NBMAX];
  sinertia Ycrb[NBMAX];
  real mcom[NBMAX][3], mvcom[NBMAX][3], macom[NBMAX][3], mass[NBMAX];
  memset(d->M, 0, sizeof d->M);
  memset(d->J, 0, sizeof d->J);

  /* ---- forward pass: pinocchio CATForwardStep + forwardKinematics(q,v,0) ---- */
  for (int i = 0; i < nb; i++) {
    sv vJ;
    memset(&vJ, 0, sizeof vJ);
    if (i == 0) {
      real qq[4] = {q[3], q[4], q[5], q[6]};
      quat_to_R(qq, d->liMi[0].R);
      for (int k = 0; k < 3; k++) { d->liMi[0].p[k] = q[k]; vJ.lin[k] = v[k]; vJ.ang[k] = v[3 + k]; }
      d->oMi[0] = d->liMi[0];
      d->v[0] = vJ;
    } else {
      real ca = R_COS((real)q[6 + i]), sa = R_SIN((real)q[6 + i]);
      se3 jM, pl;
      real Rz[9] = {ca, -sa, 0, sa, ca, 0, 0, 0, 1};
      memcpy(jM.R, Rz, sizeof Rz);
      jM.p[0] = jM.p[1] = jM.p[2] = 0;
      for (int k = 0; k < 9; k++) pl.R[k] = m->jR[i][k];
      for (int k = 0; k < 3; k++) pl.p[k] = m->jp[i][k];
      se3_mul(&pl, &jM, &d->liMi[i]);
      vJ.ang[2] = v[5 + i];
      int par = m->parent[i];
      se3_mul(&d->oMi[par], &d->liMi[i], &d->oMi[i]);
      sv vp;
      se3_actinv_motion(&d->liMi[i], &d->v[par], &vp);
      for (int k = 0; k < 3; k++) { d->v[i].lin[k] = vJ.lin[k] + vp.lin[k]; d->v[i].ang[k] = vJ.ang[k] + vp.ang[k]; }
    }
    /* world joint Jacobian columns: oMi.act(S) */
    for (int k = 0; k < ndof(i); k++) {
      sv s, sw;
      S_col(i, k, &s);
      se3_act_motion(&d->oMi[i], &s, &sw);
      real* col = &d->J[6 * (dof0(i) + k)];
      for (int r = 0; r < 3; r++) { col[r] = sw.lin[r]; col[3 + r] = sw.ang[r]; }
    }
    /* a_gf = c + v x vJ + liMi.actInv(a_gf[parent]); same without gravity for the drift */
    sv vxvj;
    motion_cross(&d->v[i], &vJ, &vxvj);
    sv ap, agp;
    if (i == 0) {
      sv g0, z0;
      memset(&g0, 0, sizeof g0);
      memset(&z0, 0, sizeof z0);
      for (int k = 0; k < 3; k++) g0.lin[k] = -(real)m->gravity[k];
      se3_actinv_motion(&d->liMi[0], &g0, &agp);
      ap = z0;
    } else {
      se3_actinv_motion(&d->liMi[i], &a_gf[m->parent[i]], &agp);
      se3_actinv_motion(&d->liMi[i], &d->a[m->parent[i]], &ap);
    }
    for (int k = 0; k < 3; k++) {
      a_gf[i].lin[k] = vxvj.lin[k] + agp.lin[k]; a_gf[i].ang[k] = vxvj.ang[k] + agp.ang[k];
      d->a[i].lin[k] = vxvj.lin[k] + ap.lin[k];  d->a[i].ang[k] = vxvj.ang[k] + ap.ang[k];
    }
    sinertia Y;
    load_inertia(m, i, &Y);
    Ycrb[i] = Y;
    sv Ya, h, vxh;
    inertia_mul(&Y, &a_gf[i], &Ya);
    inertia_mul(&Y, &d->v[i], &h);
    motion_cross_force(&d->v[i], &h, &vxh);
    for (int k = 0; k < 3; k++) { f[i].lin[k] = Ya.lin[k] + vxh.lin[k]; f[i].ang[k] = Ya.ang[k] + vxh.ang[k]; }
    /* CoM terms in the local frame, mass weighted (pinocchio centerOfMass with acceleration) */
    mass[i] = Y.m;
    real wxc[3], wxv[3], axc[3];
    cross3(d->v[i].ang, Y.c, wxc);
    for (int k = 0; k < 3; k++) { mcom[i][k] = Y.m * Y.c[k]; mvcom[i][k] = Y.m * (wxc[k] + d->v[i].lin[k]); }
    cross3(d->a[i].ang, Y.c, axc);
    cross3(d->v[i].ang, mvcom[i], wxv);
    for (int k = 0; k < 3; k++) macom[i][k] = Y.m * (axc[k] + d->a[i].lin[k]) + wxv[k];
  }

  /* ---- backward pass: CRBA, nle, CoM, Jcom (CATBackwardStep) ---- */
  for (int i = nb - 1; i >= 0; i--) {
    /* CRBA: F = Ycrb[i]*S in frame i, carried up the ancestors */
    for (int k = 0; k < ndof(i); k++) {
      sv s, F;
      S_col(i, k, &s);
      inertia_mul(&Ycrb[i], &s, &F);
      int col = dof0(i) + k;
      int j = i;
      for (;;) {
        for (int kk = 0; kk < ndof(j); kk++) {
          sv sj;
          S_col(j, kk, &sj);
          real val = 0;
          for (int r = 0; r < 3; r++) val += sj.lin[r] * F.lin[r] + sj.ang[r] * F.ang[r];
          int row = dof0(j) + kk;
          if (row <= col) d->M[row * nv + col] = val;
        }
        if (j == 0) break;
        sv Fp;
        se3_act_force(&d->liMi[j], &F, &Fp);
        F = Fp;
        j = m->parent[j];
      }
    }
    for (int k = 0; k < ndof(i); k++) {
      sv s;
      S_col(i, k, &s);
      real val = 0;
      for (int r = 0; r < 3; r++) val += s.lin[r] * f[i].lin[r] + s.ang[r] * f[i].ang[r];
      d->nle[dof0(i) + k] = val;
    }
    /* Jcom columns of body i (needs the subtree mass/com of i, complete at this point) */
    {
      real cw[3];
      matvec3(d->oMi[i].R, mcom[i], cw);
      for (int r = 0; r < 3; r++) cw[r] += mass[i] * d->oMi[i].p[r];
      for (int k = 0; k < ndof(i); k++) {
        const real* col = &d->J[6 * (dof0(i) + k)];
        real cx[3];
        cross3(cw, col + 3, cx);
        for (int r = 0; r < 3; r++) d->Jcom[r * nv + dof0(i) + k] = mass[i] * col[r] - cx[r];
      }
    }
    if (i > 0) {
      int par = m->parent[i];
      sinertia Yp;
      inertia_se3_act(&d->liMi[i], &Ycrb[i], &Yp);
      inertia_add(&Ycrb[par], &Yp);
      sv fp;
      se3_act_force(&d->liMi[i], &f[i], &fp);
      for (int k = 0; k < 3; k++) { f[par].lin[k] += fp.lin[k]; f[par].ang[k] += fp.ang[k]; }
      real t[3];
      matvec3(d->liMi[i].R, mcom[i], t);
      for (int k = 0; k < 3; k++) mcom[par][k] += t[k] + mass[i] * d->liMi[i].p[k];
      matvec3(d->liMi[i].R, mvcom[i], t);
      for (int k = 0; k < 3; k++) mvcom[par][k] += t[k];
      matvec3(d->liMi[i].R, macom[i], t);
      for (int k = 0; k < 3; k++) macom[par][k] += t[k];
      mass[par] += mass[i];
    }
  }
  /* tsid::RobotWrapper::computeAllTerms: copy the upper triangle of M to the lower one */
  for (int r = 0; r < nv; r++)
    for (int c = 0; c < r; c++) d->M[r * nv + c] = d->M[c * nv + r];
  /* body 0 -> universe */
  {
    real t[3];
    d->mass = mass[0];
    matvec3(d->liMi[0].R, mcom[0], t);
    for (int k = 0; k < 3; k++) d->com[k] = (t[k] + mass[0] * d->liMi[0].p[k]) / mass[0];
    matvec3(d->liMi[0].R, mvcom[0], t);
    for (int k = 0; k < 3; k++) d->vcom[k] = t[k] / mass[0];
    matvec3(d->liMi[0].R, macom[0], t);
    for (int k = 0; k < 3; k++) d->acom[k] = t[k] / mass[0];
    for (int k = 0; k < 3 * nv; k++) d->Jcom[k] /= mass[0];
  }

  /* ---- ccrba: Ag = X_com^* sum oYcrb*J ; and d(hg)/dt drift (pinocchio
   * computeCentroidalMomentumTimeVariation(model,data) with data.a = drift) ---- */
  {
    sinertia oY[NBMAX];
    for (int i = 0; i < nb; i++) {
      sinertia Y;
      load_inertia(m, i, &Y);
      inertia_se3_act(&d->oMi[i], &Y, &oY[i]);
    }
    for (int i = nb - 1; i >= 0; i--) {
      for (int k = 0; k < ndof(i); k++) {
        const real* col = &d->J[6 * (dof0(i) + k)];
        sv s, F;
        for (int r = 0; r < 3; r++) { s.lin[r] = col[r]; s.ang[r] = col[3 + r]; }
        inertia_mul(&oY[i], &s, &F);
        real cx[3];
        cross3(F.lin, d->com, cx);
        for (int r = 0; r < 3; r++) {
          d->Ag[r * nv + dof0(i) + k] = F.lin[r];
          d->Ag[(3 + r) * nv + dof0(i) + k] = F.ang[r] + cx[r];
        }
      }
      if (i > 0) inertia_add(&oY[m->parent[i]], &oY[i]);
    }
    sv fa[NBMAX];
    for (int i = 0; i < nb; i++) {
      sinertia Y;
      load_inertia(m, i, &Y);
      sv Ya, h, vxh;
      inertia_mul(&Y, &d->a[i], &Ya);
      inertia_mul(&Y, &d->v[i], &h);
      motion_cross_force(&d->v[i], &h, &vxh);
      for (int k = 0; k < 3; k++) { fa[i].lin[k] = Ya.lin[k] + vxh.lin[k]; fa[i].ang[k] = Ya.ang[k] + vxh.ang[k]; }
    }
    for (int i = nb - 1; i > 0; i--) {
      sv fp;
      se3_act_force(&d->liMi[i], &fa[i], &fp);
      for (int k = 0; k < 3; k++) { fa[m->parent[i]].lin[k] += fp.lin[k]; fa[m->parent[i]].ang[k] += fp.ang[k]; }
    }
    sv f0;
    se3_act_force(&d->liMi[0], &fa[0], &f0);
    real cx[3];
    cross3(f0.lin, d->com, cx);
    for (int k = 0; k < 3; k++) d->dAg_v_ang[k] = f0.ang[k] + cx[k];
  }

  /* ---- operational frames (updateFramePlacements, frameVelocity, frameClassicAcceleration,
   * frameJacobianLocal) ---- */
  for (int s = 0; s < 2; s++) {
    int b = m->foot_body[s];
    se3 pl;
    for (int k = 0; k < 9; k++) pl.R[k] = m->fR[s][k];
    for (int k = 0; k < 3; k++) pl.p[k] = m->fp[s][k];
    se3_mul(&d->oMi[b], &pl, &d->oMf[s]);
    se3_actinv_motion(&pl, &d->v[b], &d->vF[s]);
    se3_actinv_motion(&pl, &d->a[b], &d->aF[s]);
    real wxv[3];
    cross3(d->vF[s].ang, d->vF[s].lin, wxv);
    if (!g_assume[ORACLE_A_SPATIAL_FRAME_ACC]) /* A11.6: frameClassicAcceleration = spatial + w x v */
      for (int k = 0; k < 3; k++) d->aF[s].lin[k] += wxv[k];
    memset(d->JF[s], 0, sizeof d->JF[s]);
    for (int j = b; j >= 0; j = m->parent[j]) {
      for (int k = 0; k < ndof(j); k++) {
        int c = dof0(j) + k;
        sv sw, sl;
        for (int r = 0; r < 3; r++) { sw.lin[r] = d->J[6 * c + r]; sw.ang[r] = d->J[6 * c + 3 + r]; }
        se3_actinv_motion(&d->oMf[s], &sw, &sl);
        for (int r = 0; r < 3; r++) { d->JF[s][r * nv + c] = sl.lin[r]; d->JF[s][(3 + r) * nv + c] = sl.ang[r]; }
      }
      if (j == 0) break;
    }
  }
}

/* ------------------------------------------------------------------ problem assembly */
typedef struct {
  int n, nv, na, nc, neq, nin; /* nin = one-sided rows = 2 * reference nIn */
  int foot_of_slot[2];         /* contact slot (x order) -> foot (0 LF, 1 RF) */
  int slot_of_foot[2];         /* foot -> slot or -1 */
  real H[NMAX * NMAX], g[NMAX];
  real CE[NEQMAX * NMAX], ce0[NEQMAX];
  real CI[NINMAX * NMAX], ci0[NINMAX];
  int ci_canon[NINMAX];  /* stacked CI row -> block-wise numbering (identity unless ORACLE_A_CI_INTERLEAVED) */
} qp_t;

static void se3_from_vec12(const double* r, se3* M) {
  /* tsid vectorToSE3: p, then R column-major */
  for (int k = 0; k < 3; k++) M->p[k] = r[k];
  for (int c = 0; c < 3; c++)
    for (int rr = 0; rr < 3; rr++) M->R[3 * rr + c] = r[3 + 3 * c + rr];
}

/* tsid::TaskSE3Equality::compute, local frame.  ref: pos(12) vel(6) acc(6); vel/acc may be NULL (zero) */
static void se3_task(const dyn_t* d, int foot, const double* kp, const double* kd, const double* ref12,
                     const double* vref, const double* aref, real* b6) {
  se3 Mref, err;
  se3_from_vec12(ref12, &Mref);
  sv pe;
  if (!g_assume[ORACLE_A_LOG6_OLD_SIGN]) {
    se3_inv_mul(&d->oMf[foot], &Mref, &err); /* errorInSE3: log6(oMi^-1 * Mref) */
    log6_(&err, &pe);
  } else { /* A11.7, the older tsid convention: a_des = -Kp log6(Mref^-1 oMi) + ... */
    se3_inv_mul(&Mref, &d->oMf[foot], &err);
    log6_(&err, &pe);
    for (int k = 0; k < 3; k++) { pe.lin[k] = -pe.lin[k]; pe.ang[k] = -pe.ang[k]; }
  }
  sv vr, ar, vrl, arl;
  memset(&vr, 0, sizeof vr);
  memset(&ar, 0, sizeof ar);
  if (vref) for (int k = 0; k < 3; k++) { vr.lin[k] = vref[k]; vr.ang[k] = vref[3 + k]; }
  if (aref) for (int k = 0; k < 3; k++) { ar.lin[k] = aref[k]; ar.ang[k] = aref[3 + k]; }
  /* m_wMl has the frame rotation and zero translation: actInv = R^T on both parts */
  matTvec3(d->oMf[foot].R, vr.lin, vrl.lin);
  matTvec3(d->oMf[foot].R, vr.ang, vrl.ang);
  matTvec3(d->oMf[foot].R, ar.lin, arl.lin);
  matTvec3(d->oMf[foot].R, ar.ang, arl.ang);
  for (int k = 0; k < 3; k++) {
    real ades_l = kp[k] * pe.lin[k] + kd[k] * (vrl.lin[k] - d->vF[foot].lin[k]) + arl.lin[k];
    real ades_a = kp[3 + k] * pe.ang[k] + kd[3 + k] * (vrl.ang[k] - d->vF[foot].ang[k]) + arl.ang[k];
    b6[k] = ades_l - d->aF[foot].lin[k];
    b6[3 + k] = ades_a - d->aF[foot].ang[k];
  }
}

/* tsid::Contact6d force generator T (6x12), inequality matrix B (17x12) */
static void contact_mats(const tsidb_conf* c, real* T, real* B, real* lb, real* ub) {
  memset(T, 0, sizeof(real) * 72);
  for (int i = 0; i < 4; i++) {
    real p[3] = {c->contact_points[0][i], c->contact_points[1][i], c->contact_points[2][i]};
    for (int k = 0; k < 3; k++) T[k * 12 + 3 * i + k] = 1;
    /* skew(p) */
    T[3 * 12 + 3 * i + 1] = -p[2]; T[3 * 12 + 3 * i + 2] = p[1];
    T[4 * 12 + 3 * i + 0] = p[2];  T[4 * 12 + 3 * i + 2] = -p[0];
    T[5 * 12 + 3 * i + 0] = -p[1]; T[5 * 12 + 3 * i + 1] = p[0];
  }
  real n[3] = {c->contact_normal[0], c->contact_normal[1], c->contact_normal[2]};
  real ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, t1[3], t2[3];
  cross3(n, ex, t1);
  if (R_SQRT(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]) < (real)1e-5) cross3(n, ey, t1);
  cross3(n, t1, t2);
  real n1 = R_SQRT(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
  real n2 = R_SQRT(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
  for (int k = 0; k < 3; k++) { t1[k] /= n1; t2[k] /= n2; }
  memset(B, 0, sizeof(real) * 17 * 12);
  real mu = c->mu;
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 3; k++) {
      B[(4 * i + 0) * 12 + 3 * i + k] = -t1[k] - mu * n[k];
      B[(4 * i + 1) * 12 + 3 * i + k] = t1[k] - mu * n[k];
      B[(4 * i + 2) * 12 + 3 * i + k] = -t2[k] - mu * n[k];
      B[(4 * i + 3) * 12 + 3 * i + k] = t2[k] - mu * n[k];
      B[16 * 12 + 3 * i + k] = n[k];
    }
  for (int i = 0; i < 16; i++) { lb[i] = (real)-1e10; ub[i] = 0; }
  lb[16] = c->fmin;
  ub[16] = c->fmax;
}

static void add_cost(qp_t* P, real w, const real* A, const real* b, int rows) {
  /* H += w A^T A ; g -= w A^T b   (SolverHQuadProgFast::solve) */
  const int n = P->n;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      real s = 0;
      for (int r = 0; r < rows; r++) s += A[r * n + i] * A[r * n + j];
      P->H[i * n + j] += w * s;
    }
  for (int i = 0; i < n; i++) {
    real s = 0;
    for (int r = 0; r < rows; r++) s += A[r * n + i] * b[r];
    P->g[i] -= w * s;
  }
}

static void add_ineq(qp_t* P, const real* A, const real* lb, const real* ub, int rows) {
  const int n = P->n;
  for (int r = 0; r < rows; r++) {
    /* A11.4: block-wise (all lower sides of the constraint, then all upper sides) or interleaved (lb_r, ub_r, ...);
     * ci_canon maps a stacked row back to the block-wise numbering every comparison uses */
    const int lo = g_assume[ORACLE_A_CI_INTERLEAVED] ? P->nin + 2 * r : P->nin + r;
    const int hi = g_assume[ORACLE_A_CI_INTERLEAVED] ? P->nin + 2 * r + 1 : P->nin + rows + r;
    for (int j = 0; j < n; j++) {
      P->CI[lo * n + j] = A[r * n + j];
      P->CI[hi * n + j] = -A[r * n + j];
    }
    P->ci0[lo] = -lb[r];
    P->ci0[hi] = ub[r];
    P->ci_canon[lo] = P->nin + r;
    P->ci_canon[hi] = P->nin + rows + r;
  }
  P->nin += 2 * rows;
}

static void ot_assemble(const tsidb_model* m, const tsidb_conf* c, const oracle_problem* pb,
                        const dyn_t* d, qp_t* P, real* Jc_out /* 24 x nv */) {
  const int nv = d->nv, na = d->na, nc = pb->nc;
  const int n = nv + 12 * nc;
  memset(P, 0, sizeof *P);
  P->n = n; P->nv = nv; P->na = na; P->nc = nc;
  P->slot_of_foot[0] = P->slot_of_foot[1] = -1;
  for (int s = 0; s < nc; s++) { P->foot_of_slot[s] = pb->contact_order[s]; P->slot_of_foot[pb->contact_order[s]] = s; }

  real T[72], B[17 * 12], flb[17], fub[17];
  contact_mats(c, T, B, flb, fub);

  /* contact motion constraints and Jc = T^T J (12 x nv per contact, x order) */
  real Jc[24 * NVMAX];
  real bmot[2][6];
  memset(Jc, 0, sizeof Jc);
  for (int s = 0; s < nc; s++) {
    int foot = P->foot_of_slot[s];
    se3_task(d, foot, c->kp_contact, c->kd_contact, pb->ref_contact[foot], NULL, NULL, bmot[s]);
    for (int r = 0; r < 12; r++)
      for (int j = 0; j < nv; j++) {
        real acc = 0;
        for (int k = 0; k < 6; k++) acc += T[k * 12 + r] * d->JF[foot][k * nv + j];
        Jc[(12 * s + r) * nv + j] = acc;
      }
  }
  if (Jc_out) memcpy(Jc_out, Jc, sizeof(real) * 24 * nv);

  /* ---- level 0 equalities: base dynamics, then contact motion in contact order ---- */
  for (int r = 0; r < 6; r++) {
    for (int j = 0; j < nv; j++) P->CE[r * n + j] = d->M[r * nv + j];
    for (int j = 0; j < 12 * nc; j++) P->CE[r * n + nv + j] = -Jc[j * nv + r];
    P->ce0[r] = -(-d->nle[r]); /* ce0 = -vector, vector = -h_u */
  }
  P->neq = 6;
  for (int s = 0; s < nc; s++) {
    int foot = P->foot_of_slot[s];
    for (int r = 0; r < 6; r++) {
      for (int j = 0; j < nv; j++) P->CE[(P->neq + r) * n + j] = d->JF[foot][r * nv + j];
      P->ce0[P->neq + r] = -bmot[s][r];
    }
    P->neq += 6;
  }

  /* ---- level 0 inequalities in the order given ---- */
  static const real BIG = (real)1e10;
  for (int bi = 0; bi < pb->n_ci_blocks; bi++) {
    int blk = pb->ci_order[bi];
    real A[NVMAX * NMAX], lb[NVMAX], ub[NVMAX];
    if (blk == ORACLE_CI_FORCE_LF || blk == ORACLE_CI_FORCE_RF) {
      int foot = blk == ORACLE_CI_FORCE_LF ? 0 : 1;
      int s = P->slot_of_foot[foot];
      if (s < 0) continue;
      memset(A, 0, sizeof(real) * 17 * n);
      for (int r = 0; r < 17; r++)
        for (int j = 0; j < 12; j++) A[r * n + nv + 12 * s + j] = B[r * 12 + j];
      add_ineq(P, A, flb, fub, 17);
    } else if (blk == ORACLE_CI_ACTUATION) {
      /* [M_a | -J_a^T] x in [tau_min - h_a, tau_max - h_a] */
      for (int r = 0; r < na; r++) {
        for (int j = 0; j < nv; j++) A[r * n + j] = d->M[(6 + r) * nv + j];
        for (int j = 0; j < 12 * nc; j++) A[r * n + nv + j] = -Jc[j * nv + 6 + r];
        lb[r] = c->tau_min[r] - d->nle[6 + r];
        ub[r] = c->tau_max[r] - d->nle[6 + r];
      }
      add_ineq(P, A, lb, ub, na);
    } else if (blk == ORACLE_CI_JOINT_BOUNDS) {
      /* tsid::TaskJointBounds::compute: identity rows over dv, base rows +-1e10 */
      memset(A, 0, sizeof(real) * nv * n);
      for (int r = 0; r < nv; r++) A[r * n + r] = 1;
      for (int r = 0; r < 6; r++) { lb[r] = -BIG; ub[r] = BIG; }
      for (int i = 0; i < na; i++) {
        real hi = ((real)c->v_max[i] - (real)pb->v[6 + i]) / (real)c->joint_bounds_dt;
        real lo = ((real)c->v_min[i] - (real)pb->v[6 + i]) / (real)c->joint_bounds_dt;
        ub[6 + i] = hi < BIG ? hi : BIG;
        lb[6 + i] = lo > -BIG ? lo : -BIG;
      }
      add_ineq(P, A, lb, ub, nv);
    }
  }

  /* ---- level 1 cost in the order given ---- */
  for (int ti = 0; ti < pb->n_cost; ti++) {
    int task = pb->cost_order[ti];
    real A[NVMAX * NMAX], b[NVMAX];
    memset(A, 0, sizeof A);
    if (task == ORACLE_T_FORCEREG_LF || task == ORACLE_T_FORCEREG_RF) {
      int foot = task == ORACLE_T_FORCEREG_LF ? 0 : 1;
      int s = P->slot_of_foot[foot];
      if (s < 0) continue;
      if (!g_assume[ORACLE_A_FORCEREG_12x12]) {
        for (int r = 0; r < 6; r++) {
          for (int j = 0; j < 12; j++) A[r * n + nv + 12 * s + j] = (real)c->force_reg_weights[r] * T[r * 12 + j];
          b[r] = 0; /* A * fRef, fRef = 0 */
        }
        add_cost(P, c->w_force_reg, A, b, 6);
      } else { /* A11.1 alternative: the regularisation acts on the 12 corner-force components themselves */
        for (int r = 0; r < 12; r++) { A[r * n + nv + 12 * s + r] = 1; b[r] = 0; }
        add_cost(P, c->w_force_reg, A, b, 12);
      }
    } else if (task == ORACLE_T_FOOT_LF || task == ORACLE_T_FOOT_RF) {
      int foot = task == ORACLE_T_FOOT_LF ? 0 : 1;
      real b6[6];
      se3_task(d, foot, c->kp_foot, c->kd_foot, pb->ref_foot[foot], pb->ref_foot[foot] + 12, pb->ref_foot[foot] + 18, b6);
      for (int r = 0; r < 6; r++) {
        for (int j = 0; j < nv; j++) A[r * n + j] = d->JF[foot][r * nv + j];
        b[r] = b6[r];
      }
      add_cost(P, c->w_foot, A, b, 6);
    } else if (task == ORACLE_T_COM) {
      for (int r = 0; r < 3; r++) {
        for (int j = 0; j < nv; j++) A[r * n + j] = d->Jcom[r * nv + j];
        real pe = d->com[r] - (real)pb->ref_com[r];
        real ve = d->vcom[r] - (real)pb->ref_com[3 + r];
        real ades = -(real)c->kp_com[r] * pe - (real)c->kd_com[r] * ve + (real)pb->ref_com[6 + r];
        b[r] = ades - d->acom[r];
      }
      add_cost(P, c->w_com, A, b, 3);
    } else if (task == ORACLE_T_POSTURE) {
      for (int r = 0; r < na; r++) {
        A[r * n + 6 + r] = 1;
        real pe = (real)pb->q[7 + r] - (real)pb->ref_posture[r];
        real ve = (real)pb->v[6 + r];
        b[r] = -(real)c->kp_posture[r] * pe - (real)c->kd_posture[r] * ve;
      }
      add_cost(P, c->w_posture, A, b, na);
    } else if (task == ORACLE_T_AM) {
      /* tsid::TaskAMEquality: A = Ag_ang, b = -Kp (L - Lref) + dLref - drift, Lref = dLref = 0 */
      for (int r = 0; r < 3; r++) {
        real L = 0;
        for (int j = 0; j < nv; j++) {
          A[r * n + j] = d->Ag[(3 + r) * nv + j];
          L += d->Ag[(3 + r) * nv + j] * (real)pb->v[j];
        }
        b[r] = -(real)c->kp_am[r] * L - d->dAg_v_ang[r];
      }
      add_cost(P, c->w_am, A, b, 3);
    }
  }
  for (int i = 0; i < n; i++) P->H[i * n + i] += (real)c->hessian_reg;
}

/* ------------------------------------------------------------------ eiquadprog-fast */
typedef struct {
  real J[NMAX * NMAX], R[NMAX * NMAX], L[NMAX * NMAX];
  real d[NMAX], z[NMAX], r[NMAX + NEQMAX], np[NMAX], u[NMAX + NEQMAX + 1], x_old[NMAX], u_old[NMAX + NEQMAX + 1];
  real s[NINMAX];
  int A[NMAX + NEQMAX + 1], A_old[NMAX + NEQMAX + 1], iai[NINMAX], iaexcl[NINMAX];
} eq_ws;

static real eq_distance(real a, real b) {
  real a1 = R_FABS(a), b1 = R_FABS(b);
  if (a1 > b1) { real t = b1 / a1; return a1 * R_SQRT(1 + t * t); }
  else if (b1 > a1) { real t = a1 / b1; return b1 * R_SQRT(1 + t * t); }
  return a1 * R_SQRT((real)2);
}
static void compute_d(int n, real* d, const real* J, const real* np) { /* d = J^T np */
  for (int i = 0; i < n; i++) {
    real s = 0;
    for (int k = 0; k < n; k++) s += J[k * n + i] * np[k];
    d[i] = s;
  }
}
static void update_z(int n, real* z, const real* J, const real* d, int iq) { /* z = J[:, iq:] d[iq:] */
  for (int i = 0; i < n; i++) {
    real s = 0;
    for (int k = iq; k < n; k++) s += J[i * n + k] * d[k];
    z[i] = s;
  }
}
static void update_r(int n, const real* R, real* r, const real* d, int iq) { /* R[:iq,:iq] r = d[:iq] */
  for (int i = iq - 1; i >= 0; i--) {
    real s = d[i];
    for (int k = i + 1; k < iq; k++) s -= R[i * n + k] * r[k];
    r[i] = s / R[i * n + i];
  }
}
static int add_constraint(int n, real* R, real* J, real* d, int* iq, real* R_norm) {
  for (int j = n - 1; j >= *iq + 1; j--) {
    real cc = d[j - 1], ss = d[j];
    real h = eq_distance(cc, ss);
    if (h == 0) continue;
    d[j] = 0;
    ss = ss / h;
    cc = cc / h;
    if (cc < 0) { cc = -cc; ss = -ss; d[j - 1] = -h; }
    else d[j - 1] = h;
    real xny = ss / (1 + cc);
    for (int k = 0; k < n; k++) {
      real t1 = J[k * n + j - 1], t2 = J[k * n + j];
      J[k * n + j - 1] = t1 * cc + t2 * ss;
      J[k * n + j] = xny * (t1 + J[k * n + j - 1]) - t2;
    }
  }
  (*iq)++;
  for (int i = 0; i < *iq; i++) R[i * n + (*iq - 1)] = d[i];
  if (R_FABS(d[*iq - 1]) <= QP_EPS * *R_norm) return 0;
  if (R_FABS(d[*iq - 1]) > *R_norm) *R_norm = R_FABS(d[*iq - 1]);
  return 1;
}
static void delete_constraint(int n, real* R, real* J, int* A, real* u, int neq, int* iq, int l) {
  int qq = 0;
  for (int i = neq; i < *iq; i++)
    if (A[i] == l) { qq = i; break; }
  for (int i = qq; i < *iq - 1; i++) {
    A[i] = A[i + 1];
    u[i] = u[i + 1];
    for (int k = 0; k < n; k++) R[k * n + i] = R[k * n + i + 1];
  }
  A[*iq - 1] = A[*iq];
  u[*iq - 1] = u[*iq];
  A[*iq] = 0;
  u[*iq] = 0;
  for (int j = 0; j < *iq; j++) R[j * n + *iq - 1] = 0;
  (*iq)--;
  if (*iq == 0) return;
  for (int j = qq; j < *iq; j++) {
    real cc = R[j * n + j], ss = R[(j + 1) * n + j];
    real h = eq_distance(cc, ss);
    if (h == 0) continue;
    cc = cc / h;
    ss = ss / h;
    R[(j + 1) * n + j] = 0;
    if (cc < 0) { R[j * n + j] = -h; cc = -cc; ss = -ss; }
    else R[j * n + j] = h;
    real xny = ss / (1 + cc);
    for (int k = j + 1; k < *iq; k++) {
      real t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
      R[j * n + k] = t1 * cc + t2 * ss;
      R[(j + 1) * n + k] = xny * (t1 + R[j * n + k]) - t2;
    }
    for (int k = 0; k < n; k++) {
      real t1 = J[k * n + j], t2 = J[k * n + j + 1];
      J[k * n + j] = t1 * cc + t2 * ss;
      J[k * n + j + 1] = xny * (J[k * n + j] + t1) - t2;
    }
  }
}

enum { EQ_OPTIMAL = 0, EQ_INFEASIBLE = 1, EQ_UNBOUNDED = 2, EQ_MAX_ITER = 3, EQ_REDUNDANT_EQ = 4 };

/* eiquadprog-fast solve_quadprog.  Returns the EiquadprogFast status. */
static int eq_solve(const qp_t* P, int max_iter, eq_ws* w, real* x, int* iters_out, int* q_out, real* u_out, int* A_out) {
  const int n = P->n, neq = P->neq, nin = P->nin;
  const real inf = (real)DBL_MAX;
  real* J = w->J; real* R = w->R; real* L = w->L;
  real *d = w->d, *z = w->z, *r = w->r, *np = w->np, *u = w->u, *s = w->s;
  int* A = w->A; int* iai = w->iai; int* iaexcl = w->iaexcl;
  int iter = 0, iq, ip = 0, l = 0;
  real c1 = 0, c2 = 0, R_norm = 1, t, t1, t2, ss, psi;
  *iters_out = 0; *q_out = 0;

  for (int i = 0; i < n; i++) c1 += P->H[i * n + i];
  /* Eigen::LLT (lower) */
  memset(L, 0, sizeof(real) * n * n);
  for (int j = 0; j < n; j++) {
    real sdiag = P->H[j * n + j];
    for (int k = 0; k < j; k++) sdiag -= L[j * n + k] * L[j * n + k];
    if (!(sdiag > 0)) return EQ_UNBOUNDED;
    real ljj = R_SQRT(sdiag);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; i++) {
      real sacc = P->H[i * n + j];
      for (int k = 0; k < j; k++) sacc -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = sacc / ljj;
    }
  }
  memset(R, 0, sizeof(real) * n * n);
  memset(d, 0, sizeof(real) * n);
  /* J = L^-T : solve L^T J = I column by column */
  for (int c = 0; c < n; c++) {
    for (int i = n - 1; i >= 0; i--) {
      real sacc = (i == c) ? 1 : 0;
      for (int k = i + 1; k < n; k++) sacc -= L[k * n + i] * J[k * n + c];
      J[i * n + c] = sacc / L[i * n + i];
    }
  }
  for (int i = 0; i < n; i++) c2 += J[i * n + i];
  /* x = -H^-1 g */
  {
    real y[NMAX];
    for (int i = 0; i < n; i++) {
      real sacc = P->g[i];
      for (int k = 0; k < i; k++) sacc -= L[i * n + k] * y[k];
      y[i] = sacc / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      real sacc = y[i];
      for (int k = i + 1; k < n; k++) sacc -= L[k * n + i] * x[k];
      x[i] = sacc / L[i * n + i];
    }
    for (int i = 0; i < n; i++) x[i] = -x[i];
  }
  /* equality constraints */
  iq = 0;
  for (int i = 0; i < neq; i++) {
    for (int k = 0; k < n; k++) np[k] = P->CE[i * n + k];
    compute_d(n, d, J, np);
    update_z(n, z, J, d, iq);
    update_r(n, R, r, d, iq);
    real zz = 0, znp = 0, npx = 0;
    for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; npx += np[k] * x[k]; }
    t2 = 0;
    if (R_FABS(zz) > QP_EPS) t2 = (-npx - P->ce0[i]) / znp;
    for (int k = 0; k < n; k++) x[k] += t2 * z[k];
    u[iq] = t2;
    for (int k = 0; k < iq; k++) u[k] -= t2 * r[k];
    A[i] = -i - 1;
    if (!add_constraint(n, R, J, d, &iq, &R_norm)) return EQ_REDUNDANT_EQ;
  }
  for (int i = 0; i < nin; i++) iai[i] = i;

l1:
  iter++;
  *iters_out = iter;
  if (iter >= max_iter) { *q_out = iq; goto finish_maxiter; }
  for (int i = neq; i < iq; i++) { ip = A[i]; iai[ip] = -1; }
  ss = 0;
  ip = 0;
  psi = 0;
  for (int i = 0; i < nin; i++) {
    real sacc = 0;
    for (int k = 0; k < n; k++) sacc += P->CI[i * n + k] * x[k];
    s[i] = sacc + P->ci0[i];
    iaexcl[i] = 1;
    psi += s[i] < 0 ? s[i] : 0;
  }
  if (R_FABS(psi) <= nin * QP_EPS * c1 * c2 * (real)100.0) { *q_out = iq; goto finish_optimal; }
  for (int i = 0; i < iq; i++) { w->u_old[i] = u[i]; w->A_old[i] = A[i]; }
  for (int k = 0; k < n; k++) w->x_old[k] = x[k];

l2:
  for (int i = 0; i < nin; i++)
    if (s[i] < ss && iai[i] != -1 && iaexcl[i]) { ss = s[i]; ip = i; }
  if (ss >= 0) { *q_out = iq; goto finish_optimal; }
  for (int k = 0; k < n; k++) np[k] = P->CI[ip * n + k];
  u[iq] = 0;
  A[iq] = ip;
#ifdef ORACLE_TRACE
  fprintf(stderr, "[orc] iter %d pick %d s=%.17g iq=%d\n", iter, ip, (double)ss, iq - neq);
#endif

l2a:
  compute_d(n, d, J, np);
  if (iq >= n) memset(z, 0, sizeof(real) * n);
  else update_z(n, z, J, d, iq);
  update_r(n, R, r, d, iq);
  l = 0;
  t1 = inf;
  for (int k = neq; k < iq; k++) {
    real tmp;
    if (r[k] > 0 && ((tmp = u[k] / r[k]) < t1)) { t1 = tmp; l = A[k]; }
  }
  {
    real zz = 0, znp = 0;
    for (int k = 0; k < n; k++) { zz += z[k] * z[k]; znp += z[k] * np[k]; }
    if (R_FABS(zz) > QP_EPS) t2 = -s[ip] / znp;
    else t2 = inf;
    t = t1 < t2 ? t1 : t2;
#ifdef ORACLE_TRACE
    fprintf(stderr, "[orc]   t1=%.17g (l %d) t2=%.17g zz=%.6g znp=%.6g\n", (double)t1, l, (double)t2, (double)zz, (double)znp);
#endif
    if (t >= inf) { *q_out = iq; return EQ_UNBOUNDED; }
    if (t2 >= inf) {
      for (int k = 0; k < iq; k++) u[k] -= t * r[k];
      u[iq] += t;
      iai[l] = l;
      delete_constraint(n, R, J, A, u, neq, &iq, l);
      goto l2a;
    }
    for (int k = 0; k < n; k++) x[k] += t * z[k];
    for (int k = 0; k < iq; k++) u[k] -= t * r[k];
    u[iq] += t;
  }
  if (t == t2) {
    if (!add_constraint(n, R, J, d, &iq, &R_norm)) {
      iaexcl[ip] = 0;
      delete_constraint(n, R, J, A, u, neq, &iq, ip);
      for (int i = 0; i < nin; i++) iai[i] = i;
      for (int i = 0; i < iq; i++) {
        A[i] = w->A_old[i];
        if (A[i] >= 0) iai[A[i]] = -1;
        u[i] = w->u_old[i];
      }
      for (int k = 0; k < n; k++) x[k] = w->x_old[k];
      goto l2;
    } else iai[ip] = -1;
    goto l1;
  }
  iai[l] = l;
  delete_constraint(n, R, J, A, u, neq, &iq, l);
  {
    real sacc = 0;
    for (int k = 0; k < n; k++) sacc += P->CI[ip * n + k] * x[k];
    s[ip] = sacc + P->ci0[ip];
  }
  goto l2a;

finish_optimal:
  for (int i = 0; i < iq; i++) { u_out[i] = u[i]; A_out[i] = A[i]; }
  return EQ_OPTIMAL;
finish_maxiter:
  for (int i = 0; i < iq; i++) { u_out[i] = u[i]; A_out[i] = A[i]; }
  return EQ_MAX_ITER;
}

/* ------------------------------------------------------------------ the tick */
int oracle_tick(const tsidb_model* m, const tsidb_conf* c, const oracle_problem* pb, oracle_result* out,
                oracle_dump* dump) {
  dyn_t* d = (dyn_t*)malloc(sizeof(dyn_t));
  qp_t* P = (qp_t*)malloc(sizeof(qp_t));
  eq_ws* w = (eq_ws*)malloc(sizeof(eq_ws));
  real Jc[24 * NVMAX];
  if (!d || !P || !w) { free(d); free(P); free(w); return -1; }
  ot_dynamics(m, pb->q, pb->v, d);
  ot_assemble(m, c, pb, d, P, Jc);
  const int n = P->n, nv = P->nv, na = P->na, nc = P->nc;
  real x[NMAX], u[NMAX + NEQMAX + 1];
  int A[NMAX + NEQMAX + 1];
  int iters = 0, q = 0;
  int st = eq_solve(P, c->max_iter, w, x, &iters, &q, u, A);
  /* SolverHQuadProgFast status mapping */
  int status;
  switch (st) {
    case EQ_OPTIMAL: status = TSIDB_STATUS_OPTIMAL; break;
    case EQ_UNBOUNDED: status = TSIDB_STATUS_INFEASIBLE; break;
    case EQ_INFEASIBLE: status = TSIDB_STATUS_INFEASIBLE; break;
    case EQ_MAX_ITER: status = TSIDB_STATUS_MAX_ITER_REACHED; break;
    default: status = TSIDB_STATUS_ERROR; break;
  }
  out->status = status;
  out->iters = iters;
  out->n = n;
  out->n_active = 0;
  memset(out->f, 0, sizeof out->f);
  memset(out->tau, 0, sizeof out->tau);
  memset(out->dv, 0, sizeof out->dv);
  memset(out->x, 0, sizeof out->x);
  memset(out->lambda, 0, sizeof out->lambda);
  if (st == EQ_OPTIMAL || st == EQ_MAX_ITER) {
    for (int i = 0; i < n; i++) out->x[i] = (double)x[i];
    for (int i = 0; i < q; i++) out->lambda[i] = (double)u[i];
    for (int i = P->neq; i < q; i++) out->active[out->n_active++] = P->ci_canon[A[i]];
    /* decodeSolution: dv = x[:nv], f = x[nv:], tau = h_a + M_a dv - J_a^T f */
    for (int i = 0; i < nv; i++) out->dv[i] = (double)x[i];
    for (int s = 0; s < nc; s++)
      for (int k = 0; k < 12; k++) out->f[12 * P->foot_of_slot[s] + k] = (double)x[nv + 12 * s + k];
    for (int r = 0; r < na; r++) {
      real acc = d->nle[6 + r];
      real Mdv = 0, Jf = 0;
      for (int j = 0; j < nv; j++) Mdv += d->M[(6 + r) * nv + j] * x[j];
      for (int j = 0; j < 12 * nc; j++) Jf += Jc[j * nv + 6 + r] * x[nv + j];
      out->tau[r] = (double)(acc + Mdv - Jf);
    }
  }
  /* diagnostics the reference reads from formulation.data() after the solve (ref:main.py:135-142) */
  for (int k = 0; k < 3; k++) { out->com[k] = (double)d->com[k]; out->com[3 + k] = (double)d->vcom[k]; out->com[6 + k] = (double)d->acom[k]; }
  for (int s = 0; s < 2; s++) {
    for (int k = 0; k < 3; k++) out->foot[s][k] = (double)d->oMf[s].p[k];
    for (int cc = 0; cc < 3; cc++)
      for (int rr = 0; rr < 3; rr++) out->foot[s][3 + 3 * cc + rr] = (double)d->oMf[s].R[3 * rr + cc];
  }
  if (dump) {
    dump->n = n; dump->neq = P->neq; dump->nin = P->nin; dump->nv = nv;
    for (int i = 0; i < nv * nv; i++) dump->M[i] = (double)d->M[i];
    for (int i = 0; i < nv; i++) dump->nle[i] = (double)d->nle[i];
    for (int s = 0; s < 2; s++) {
      for (int i = 0; i < 6 * nv; i++) dump->JF[s][i] = (double)d->JF[s][i];
      for (int k = 0; k < 3; k++) {
        dump->vF[s][k] = (double)d->vF[s].lin[k]; dump->vF[s][3 + k] = (double)d->vF[s].ang[k];
        dump->aF[s][k] = (double)d->aF[s].lin[k]; dump->aF[s][3 + k] = (double)d->aF[s].ang[k];
      }
    }
    for (int i = 0; i < 3 * nv; i++) dump->Jcom[i] = (double)d->Jcom[i];
    for (int i = 0; i < 6 * nv; i++) dump->Ag[i] = (double)d->Ag[i];
    for (int k = 0; k < 3; k++) dump->dAg_v_ang[k] = (double)d->dAg_v_ang[k];
    for (int i = 0; i < n * n; i++) dump->H[i] = (double)P->H[i];
    for (int i = 0; i < n; i++) dump->g[i] = (double)P->g[i];
    for (int i = 0; i < P->neq * n; i++) dump->CE[i] = (double)P->CE[i];
    for (int i = 0; i < P->neq; i++) dump->ce0[i] = (double)P->ce0[i];
    for (int i = 0; i < P->nin * n; i++) dump->CI[i] = (double)P->CI[i];
    for (int i = 0; i < P->nin; i++) dump->ci0[i] = (double)P->ci0[i];
    for (int b = 0; b < d->nb; b++) {
      for (int k = 0; k < 9; k++) dump->oMi_R[b][k] = (double)d->oMi[b].R[k];
      for (int k = 0; k < 3; k++) dump->oMi_p[b][k] = (double)d->oMi[b].p[k];
    }
  }
  free(d); free(P); free(w);
  return 0;
}

/* ------------------------------------------------------------------ integrate_dv */
/* ref:ctrl/WalkController.py:291-295: v_mean = v + dt/2 dv ; v += dt dv ; q = pin.integrate(q, dt v_mean)
 * pinocchio: free-flyer = SE3 (+) exp6, quaternion first-order renormalised; revolute joints add. */
static void exp3_(const real* w, real* R) {
  real t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  real t = R_SQRT(t2);
  const real prec3 = (real)1.220703125e-4;
  real ct = R_COS(t), st = R_SIN(t);
  real alpha_vxvx = (t > prec3) ? (1 - ct) / t2 : (real)0.5 - t2 / 24;
  real alpha_vx = (t > prec3) ? st / t : 1 - t2 / 6;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[3 * i + j] = alpha_vxvx * w[i] * w[j];
  R[1] -= alpha_vx * w[2]; R[2] += alpha_vx * w[1];
  R[3] += alpha_vx * w[2]; R[5] -= alpha_vx * w[0];
  R[6] -= alpha_vx * w[1]; R[7] += alpha_vx * w[0];
  real dg = (t > prec3) ? ct : 1 - t2 / 2;
  R[0] += dg; R[4] += dg; R[8] += dg;
}
static void exp6_(const real* v, const real* w, se3* M) {
  real t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  real t = R_SQRT(t2);
  const real prec3 = (real)1.220703125e-4;
  real ct = R_COS(t), st = R_SIN(t);
  real alpha_wxv = (t > prec3) ? (1 - ct) / t2 : (real)0.5 - t2 / 24;
  real alpha_v = (t > prec3) ? st / t : 1 - t2 / 6;
  real alpha_w = (t > prec3) ? (1 - alpha_v) / t2 : (real)1 / 6 - t2 / 120;
  real dg = (t > prec3) ? ct : 1 - t2 / 2;
  real wxv[3];
  cross3(w, v, wxv);
  real wdv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
  for (int k = 0; k < 3; k++) M->p[k] = alpha_v * v[k] + (alpha_w * wdv) * w[k] + alpha_wxv * wxv[k];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) M->R[3 * i + j] = alpha_wxv * w[i] * w[j];
  M->R[1] -= alpha_v * w[2]; M->R[2] += alpha_v * w[1];
  M->R[3] += alpha_v * w[2]; M->R[5] -= alpha_v * w[0];
  M->R[6] -= alpha_v * w[1]; M->R[7] += alpha_v * w[0];
  M->R[0] += dg; M->R[4] += dg; M->R[8] += dg;
  (void)exp3_;
}
/* Eigen::Quaternion = rotation matrix (Shepperd, as Eigen implements it); out (x,y,z,w) */
static void R_to_quat(const real* R, real* q) {
  real t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = R_SQRT(t + 1);
    q[3] = (real)0.5 * t;
    t = (real)0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = R_SQRT(R[4 * i] - R[4 * j] - R[4 * k] + 1);
    q[i] = (real)0.5 * t;
    t = (real)0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
int oracle_integrate(const tsidb_model* m, double* q, double* v, const double* dv, double dt) {
  const int nb = m->nb, nv = nb + 5, na = nb - 1;
  real vm[NVMAX];
  for (int i = 0; i < nv; i++) { vm[i] = dt * ((real)v[i] + (real)0.5 * dt * (real)dv[i]); v[i] = (double)((real)v[i] + (real)dt * (real)dv[i]); }
  /* SpecialEuclideanOperationTpl<3>::integrate_impl: M1 = exp6(v); p' = p + R p1; quat' = quat * quat(R1);
   * sign continuity, then firstOrderNormalize */
  se3 M1;
  exp6_(vm, vm + 3, &M1);
  real R0[9], qq[4] = {q[3], q[4], q[5], q[6]};
  quat_to_R(qq, R0);
  real t[3];
  matvec3(R0, M1.p, t);
  for (int k = 0; k < 3; k++) q[k] = (double)((real)q[k] + t[k]);
  real q1[4];
  R_to_quat(M1.R, q1);
  /* quaternion product a*b, (x,y,z,w) */
  real ax = qq[0], ay = qq[1], az = qq[2], aw = qq[3], bx = q1[0], by = q1[1], bz = q1[2], bw = q1[3];
  real r[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
               aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
  real dotp = r[0] * qq[0] + r[1] * qq[1] + r[2] * qq[2] + r[3] * qq[3];
  if (dotp < 0) for (int k = 0; k < 4; k++) r[k] = -r[k];
  /* quaternion::firstOrderNormalize: q *= (3 - |q|^2)/2 */
  real N2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
  real al = ((real)3 - N2) / 2;
  for (int k = 0; k < 4; k++) q[3 + k] = (double)(r[k] * al);
  for (int i = 0; i < na; i++) q[7 + i] = (double)((real)q[7 + i] + vm[6 + i]);
  return 0;
}

/* ------------------------------------------------------------------ batch driver (CPU baseline) */
typedef struct {
  const tsidb_model* m; const tsidb_conf* c; const oracle_problem* pbs; oracle_result* outs;
  int lo, hi;
} job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int i = j->lo; i < j->hi; i++) oracle_tick(j->m, j->c, &j->pbs[i], &j->outs[i], NULL);
  return NULL;
}
int oracle_tick_batch(const tsidb_model* m, const tsidb_conf* c, const oracle_problem* pbs, oracle_result* outs,
                      int n_envs, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  job_t jobs[256];
  int per = (n_envs + n_threads - 1) / n_threads;
  int nt = 0;
  for (int t = 0; t < n_threads; t++) {
    int lo = t * per, hi = lo + per > n_envs ? n_envs : lo + per;
    if (lo >= hi) break;
    jobs[t] = (job_t){m, c, pbs, outs, lo, hi};
    if (pthread_create(&th[t], NULL, worker, &jobs[t]) != 0) return -1;
    nt++;
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  return 0;
}

int oracle_real_bytes(void) { return (int)sizeof(real); }
