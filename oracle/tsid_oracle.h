/* tsid_oracle.h — interface of the CPU oracle (test infrastructure; see tsid_oracle.c). */
#ifndef TSID_ORACLE_H_
#define TSID_ORACLE_H_
#include <stdint.h>
#include "../include/tsidb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* level-0 inequality blocks, listed in formulation insertion order by the caller
 * (ref:ctrl/WalkController.py:83-85,124-126,172-176,184; re-added contacts go last) */
enum { ORACLE_CI_FORCE_LF = 0, ORACLE_CI_FORCE_RF = 1, ORACLE_CI_ACTUATION = 2, ORACLE_CI_JOINT_BOUNDS = 3 };
/* level-1 cost tasks, listed in insertion order by the caller */
enum {
  ORACLE_T_FORCEREG_LF = 0, ORACLE_T_FORCEREG_RF = 1, ORACLE_T_FOOT_LF = 2, ORACLE_T_FOOT_RF = 3,
  ORACLE_T_COM = 4, ORACLE_T_POSTURE = 5, ORACLE_T_AM = 6
};

#define ORACLE_NMAX (TSIDB_MAX_NV + 24)
#define ORACLE_NINMAX (2 * (34 + TSIDB_MAX_NA + TSIDB_MAX_NV))

typedef struct oracle_problem {
  const double* q;            /* nq */
  const double* v;            /* nv */
  int32_t nc;                 /* active contacts */
  int32_t contact_order[2];   /* x order: foot ids (0 LF, 1 RF) */
  int32_t n_ci_blocks;
  int32_t ci_order[4];
  int32_t n_cost;
  int32_t cost_order[8];
  const double* ref_com;      /* 9 */
  const double* ref_foot[2];  /* 24 each: pos 12, vel 6, acc 6 */
  const double* ref_contact[2]; /* 12 each */
  const double* ref_posture;  /* na */
} oracle_problem;

typedef struct oracle_result {
  int32_t status, iters, n, n_active;
  int32_t active[ORACLE_NMAX + 1]; /* indices into the reference's stacked CI */
  double tau[TSIDB_MAX_NA], dv[TSIDB_MAX_NV], f[24], x[ORACLE_NMAX], lambda[ORACLE_NMAX + 19];
  double com[9];
  double foot[2][12];
} oracle_result;

typedef struct oracle_dump {
  int32_t n, neq, nin, nv;
  double M[TSIDB_MAX_NV * TSIDB_MAX_NV], nle[TSIDB_MAX_NV];
  double JF[2][6 * TSIDB_MAX_NV], vF[2][6], aF[2][6];
  double Jcom[3 * TSIDB_MAX_NV], Ag[6 * TSIDB_MAX_NV], dAg_v_ang[3];
  double H[ORACLE_NMAX * ORACLE_NMAX], g[ORACLE_NMAX];
  double CE[18 * ORACLE_NMAX], ce0[18];
  double CI[ORACLE_NINMAX * ORACLE_NMAX], ci0[ORACLE_NINMAX];
  double oMi_R[TSIDB_MAX_BODIES][9], oMi_p[TSIDB_MAX_BODIES][3];
} oracle_dump;

int oracle_tick(const tsidb_model* m, const tsidb_conf* c, const oracle_problem* pb, oracle_result* out,
                oracle_dump* dump /* may be NULL */);
int oracle_tick_batch(const tsidb_model* m, const tsidb_conf* c, const oracle_problem* pbs, oracle_result* outs,
                      int n_envs, int n_threads);
int oracle_integrate(const tsidb_model* m, double* q, double* v, const double* dv, double dt);
int oracle_real_bytes(void);

/* the upstream details this restatement is least sure of, switchable one at a time (default 0 = the
 * restatement's reading) so that tests/test_assumptions.py can measure how much each one moves the answer.  The
 * joint-bounds time step (dt vs 2 dt) and the Hessian regulariser are plain tsidb_conf fields and need no switch. */
enum {
  ORACLE_A_FORCEREG_12x12 = 0,   /* A11.1: force regularisation on the 12 corner-force components, not on diag(w) T f (6x12) */
  ORACLE_A_CI_INTERLEAVED = 1,   /* A11.4: two-sided rows stacked (lb_i, ub_i) pairwise instead of block-wise */
  ORACLE_A_SPATIAL_FRAME_ACC = 2,/* A11.6: frame drift = spatial acceleration, without the w x v of frameClassicAcceleration */
  ORACLE_A_LOG6_OLD_SIGN = 3,    /* A11.7: a_des = -Kp log6(Mref^-1 M) (older tsid) instead of +Kp log6(M^-1 Mref) */
  ORACLE_A_COUNT = 4
};
int oracle_set_assumption(int which, int alt);

#ifdef __cplusplus
}
#endif
#endif
