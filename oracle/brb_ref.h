/*
 * brb_ref.h — fp64 CPU ORACLE for the balance-robot env step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (balance_robot_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic restated here lives in the third-party dependency
 * mujoco==3.2.0 (reference conda-environment.yaml:7), whose sources are not under /root/reference
 * and which cannot be installed in this image; the reference ships no golden vectors or tests.
 * The restatement follows SURVEY.md Appendix A (mj_step for this model class) and is pinned only by
 * closed-form physics checks (tests/test_oracle_physics.py).  The env logic restated in
 * brb_ref_env.c follows reference sources that ARE present (cited per function).
 *
 * Reference call sites this oracle stands in for:
 *   mujoco.mj_step(model, data, nstep=250)     envs/env01_v1.py:24, envs/env01_v2.py:37, envs/env03_v1.py:34
 *   MujocoEnv.set_state -> mj_forward          envs/env01_v1.py:57, envs/env01_v2.py:70
 *   MjModel.from_xml_path (derived constants)  envs/RobotBaseEnv.py:56-65
 */
#ifndef BRB_REF_H
#define BRB_REF_H

#ifdef __cplusplus
extern "C" {
#endif

#define BRB_MAXBODY 6
#define BRB_MAXJNT 6
#define BRB_MAXNQ 20
#define BRB_MAXNV 16
#define BRB_MAXGEOM 8
#define BRB_MAXPAIR 16
#define BRB_MAXU 4
#define BRB_MAXCON 32
#define BRB_MAXEFC 128

#define BRB_GEOM_PLANE 0
#define BRB_GEOM_CYLINDER 5
#define BRB_GEOM_BOX 6
#define BRB_JNT_FREE 0
#define BRB_JNT_HINGE 3

/* flags */
#define BRB_FLAG_ACTDERIV_SKIP_CLAMPED 1 /* A.9: no actuator velocity derivative while on forcerange */
#define BRB_FLAG_RPY_FROM_FIRST_ROW 2    /* A.7: R_py = 2 mu^2 R(first pyramid row) */
#define BRB_FLAG_CYLINDER_BOX 4          /* wheel-block contacts through the own analytic cylinder-box collider (brb_ref.c) */

typedef struct BrbRefModel {
  int nq, nv, nu, nbody, njnt, ngeom, npair, flags;
  double timestep, gravity[3];
  /* bodies (0 = world) */
  int body_parent[BRB_MAXBODY], body_jnt[BRB_MAXBODY];
  double body_pos[BRB_MAXBODY][3], body_quat[BRB_MAXBODY][4];
  double body_mass[BRB_MAXBODY], body_ipos[BRB_MAXBODY][3], body_inertia[BRB_MAXBODY][9];
  /* joints */
  int jnt_type[BRB_MAXJNT], jnt_body[BRB_MAXJNT], jnt_qposadr[BRB_MAXJNT], jnt_dofadr[BRB_MAXJNT];
  double jnt_axis[BRB_MAXJNT][3], jnt_pos[BRB_MAXJNT][3], jnt_damping[BRB_MAXJNT];
  /* geoms */
  int geom_type[BRB_MAXGEOM], geom_body[BRB_MAXGEOM];
  double geom_size[BRB_MAXGEOM][3], geom_pos[BRB_MAXGEOM][3], geom_quat[BRB_MAXGEOM][4];
  /* candidate contact pairs (explicit + dynamic, parameters already mixed) */
  int pair_geom1[BRB_MAXPAIR], pair_geom2[BRB_MAXPAIR], pair_condim[BRB_MAXPAIR];
  double pair_friction[BRB_MAXPAIR][5], pair_solref[BRB_MAXPAIR][2], pair_solimp[BRB_MAXPAIR][5];
  double pair_margin[BRB_MAXPAIR], pair_gap[BRB_MAXPAIR];
  /* velocity actuators */
  int act_jnt[BRB_MAXU], act_ctrllimited[BRB_MAXU], act_forcelimited[BRB_MAXU];
  double act_kv[BRB_MAXU], act_gear[BRB_MAXU], act_ctrlrange[BRB_MAXU][2], act_forcerange[BRB_MAXU][2];
  /* derived by brb_ref_model_finalize (mj_setConst restatement) */
  double qpos0[BRB_MAXNQ], body_invweight0[BRB_MAXBODY][2], meaninertia;
  double solver_tolerance; /* scaled-gradient stop; oracle converges far tighter than MuJoCo's 1e-8 */
} BrbRefModel;

typedef struct BrbRefContact {
  double dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5];
  int pair, dim, body1, body2, efc_address, exclude;
} BrbRefContact;

typedef struct BrbRefData {
  /* state */
  double qpos[BRB_MAXNQ], qvel[BRB_MAXNV], qacc_warmstart[BRB_MAXNV], ctrl[BRB_MAXU], time;
  /* position-dependent */
  double xpos[BRB_MAXBODY][3], xquat[BRB_MAXBODY][4], xmat[BRB_MAXBODY][9], xipos[BRB_MAXBODY][3];
  double xanchor[BRB_MAXJNT][3], xaxis[BRB_MAXJNT][3];
  double geom_xpos[BRB_MAXGEOM][3], geom_xmat[BRB_MAXGEOM][9];
  double qM[BRB_MAXNV * BRB_MAXNV];
  /* forces / accelerations */
  double qfrc_bias[BRB_MAXNV], qfrc_passive[BRB_MAXNV], qfrc_actuator[BRB_MAXNV], actuator_force[BRB_MAXU];
  double qfrc_smooth[BRB_MAXNV], qacc_smooth[BRB_MAXNV], qacc[BRB_MAXNV], qfrc_constraint[BRB_MAXNV];
  /* constraints */
  int ncon, nefc, solver_niter;
  BrbRefContact contact[BRB_MAXCON];
  double efc_J[BRB_MAXEFC * BRB_MAXNV], efc_pos[BRB_MAXEFC], efc_margin[BRB_MAXEFC], efc_aref[BRB_MAXEFC];
  double efc_R[BRB_MAXEFC], efc_D[BRB_MAXEFC], efc_force[BRB_MAXEFC], efc_vel[BRB_MAXEFC];
  /* statistics accumulated by brb_ref_step */
  long long stat_substeps, stat_contact_substeps, stat_newton_iters, stat_efc_rows;
} BrbRefData;

int brb_ref_sizeof_model(void);
int brb_ref_sizeof_data(void);
int brb_ref_sizeof_contact(void);

/* mj_setConst restatement: qpos0, body_invweight0, meaninertia.  Returns 0 or a negative errno. */
int brb_ref_model_finalize(BrbRefModel *m);
/* mj_resetData */
void brb_ref_reset_data(const BrbRefModel *m, BrbRefData *d);
/* mj_forward (no integration): kinematics .. constraint solve; updates qacc_warmstart */
void brb_ref_forward(const BrbRefModel *m, BrbRefData *d);
/* mj_step x nstep */
void brb_ref_step(const BrbRefModel *m, BrbRefData *d, int nstep);
/* pieces, exposed for tests */
void brb_ref_kinematics(const BrbRefModel *m, BrbRefData *d);
void brb_ref_mass_matrix(const BrbRefModel *m, BrbRefData *d);
void brb_ref_bias(const BrbRefModel *m, BrbRefData *d);
double brb_ref_energy(const BrbRefModel *m, BrbRefData *d, double *kinetic, double *potential);
/* own analytic cylinder-box collider (see brb_ref.c): 1 + dist / normal (cylinder -> box) / pos when within margin */
int brb_ref_cylinder_box(const double c[3], const double axis[3], double R, double L, const double b[3], const double E_rows[9],
                         const double h[3], double margin, double *dist, double normal[3], double pos[3]);
/* point Jacobian of body b at world point p: jacp, jacr are 3 x nv row-major (either may be NULL) */
void brb_ref_jac(const BrbRefModel *m, const BrbRefData *d, int body, const double p[3], double *jacp, double *jacr);

#ifdef __cplusplus
}
#endif
#endif
