/* TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU float64 restatement of EgoPose's rollout path.
 * Never linked or loaded by the product path (egopose_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * PARITY UNPINNED for the physics step: the arithmetic of sim.step()/sim.forward()/mj_fullM lives in
 * MuJoCo (third-party C library behind mujoco-py; version not pinned by the reference, README.md:21),
 * which is not present in /root/reference nor installable here.  The physics below restates MuJoCo's
 * documented smooth-dynamics pipeline; env logic, reward, filter and the PPO
 * half ARE pinned against the reference's own Python (tests/golden/).
 */
#ifndef EGOPOSE_ORACLE_H
#define EGOPOSE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define EO_MAXB 32      /* bodies */
#define EO_MAXV 64      /* dofs */
#define EO_NEE 5

/* packed expert row (doubles); same packing the product uses (include/egopose_b200.h) */
#define EO_X_QPOS 0
#define EO_X_QVEL 59
#define EO_X_RLINV_LOCAL 117
#define EO_X_RANGV 120
#define EO_X_RQ_RMH 123
#define EO_X_EE_POS 127
#define EO_X_BQUAT 142
#define EO_X_BANGVEL 226
#define EO_X_STRIDE 292

typedef struct {
    int nq, nv, nu, nbody;
    double timestep;
    double gravity[3];
    const int *body_parent, *body_dofadr, *body_dofnum, *body_qposadr;
    const double *body_pos, *body_mass, *body_ipos, *body_inertia;  /* [nb][3],[nb],[nb][3],[nb][6] */
    const int *dof_body, *dof_parent;
    const double *dof_armature, *dof_axis, *dof_anchor;             /* [nv],[nv][3],[nv][3] */
    int ee_body[EO_NEE];
    int head_body;
    /* joint limits as MuJoCo soft constraints (humanoid_1205_v1.xml:10 limited="true", :28-30 range=...; MuJoCo defaults
     * solref 0.02 1, solimp 0.9 0.95 0.001 0.5 2, margin 0).  dof_range NULL = smooth dynamics only (the north star).
     * UNPINNED like the rest of the MuJoCo arithmetic (no MuJoCo binary here): restated from MuJoCo's documented
     * constraint model (engine_core_constraint.c: mj_instantiateLimit, mj_makeImpedance, mj_referenceConstraint; the
     * solver's optimum instead of its iterates). */
    const double *dof_range;        /* [nv][2] lower, upper in radians; lower >= upper: no limit on that dof */
    const double *dof_invweight0;   /* [nv] diag(M^-1) at qpos0 (mjModel.dof_invweight0 of a hinge) */
    double solref[2], solimp[5];
    /* floor contact (xml:21 plane z = 0, condim 3, friction 1; xml:11 body geoms: margin 0.001; pyramidal cones, the
     * MuJoCo default).  One collision geom per body (0 sphere | 1 capsule | 2 box; p0 / p1 in the body frame).  Only
     * geom-floor pairs: contacts between body geoms are NOT modelled.  geom_type NULL = no contacts.  UNPINNED, restated
     * from engine_collision_primitive.c (mjc_PlaneSphere / PlaneCapsule / PlaneBox) and engine_core_constraint.c
     * (mj_instantiateContact, mj_diagApprox, mj_makeImpedance). */
    const int *geom_type;
    const double *geom_size, *geom_p0, *geom_p1;    /* [nb][3] each */
    const double *body_invweight0;                  /* [nb][2] translational, rotational (mjModel.body_invweight0) */
    double contact_margin, contact_mu;
} EoModel;

typedef struct {            /* mjData subset the reference reads */
    double qpos[EO_MAXV + 1], qvel[EO_MAXV], ctrl[EO_MAXV], qacc[EO_MAXV];
    double xpos[EO_MAXB][3], xquat[EO_MAXB][4], xipos[EO_MAXB][3];
    double qM[EO_MAXV * EO_MAXV];       /* dense, row-major (what mj_fullM returns) */
    double qfrc_bias[EO_MAXV];
    double cdof[EO_MAXV][6];
    double subtree_com[3];
    int n_efc, solver_iter;             /* constraint rows of the last eo_forward and active-set iterations it took */
} EoData;

typedef struct {            /* cfg subset (ego_pose/utils/egomimic_config.py:94-122) */
    int frame_skip, episode_len, fr_margin;
    const double *jkp, *jkd, *a_ref, *a_scale, *torque_lim;   /* [nu] */
    const double *b_diffw;                                   /* [nbody-1] */
    double w_p, w_v, w_e, w_rp, w_rv, k_p, k_v, k_e, k_rh, k_rq, k_rl, k_ra;
    int v_ord, decay;
    double end_reward;
    double fix_head_lb;     /* NaN = use expert head_height_lb - 0.1 (humanoid_v1.py:193-196) */
} EoCfg;

typedef struct {
    int n_takes;
    const int *take_off;            /* [n_takes+1] frame offsets into rows */
    const double *rows;             /* [total_frames][EO_X_STRIDE] */
    const double *head_height_lb;   /* [n_takes] */
    const double *ctx;              /* [total_frames][ctx_dim] per-frame video context, or NULL */
    int ctx_dim;
} EoExpert;

typedef struct {            /* PolicyGaussian(MLP) weights, torch nn.Linear layout [out][in] */
    int in_dim, h1, h2, out_dim;
    const double *W1, *b1, *W2, *b2, *W3, *b3, *log_std;
} EoPolicy;

typedef struct {            /* one environment (HumanoidEnv instance state) */
    EoData d;
    int cur_t, take, start_ind;
    double prev_qpos[EO_MAXV + 1], prev_qvel[EO_MAXV];
    double bquat[4 * EO_MAXB], prev_bquat[4 * EO_MAXB];
} EoEnv;

/* physics (MuJoCo restatement) */
void eo_forward(const EoModel *m, EoData *d);                   /* mj_forward (smooth part) */
void eo_step(const EoModel *m, EoData *d);                      /* mj_step: forward then Euler */
void eo_kinematics(const EoModel *m, EoData *d, double *dof_axis_w, double *dof_anchor_w);
int eo_chol_solve(int n, double *A, double *b);                 /* in-place dense Cholesky solve */
int eo_constraint_solve(const EoModel *m, EoData *d, const double *qfrc_smooth);   /* rows + solve -> d->qacc */
void eo_solver_stats(long *out3, int reset);
void eo_limit_row(const EoModel *m, double dist, double vel, double invweight, double *D, double *aref);

/* env (ego_pose/envs/humanoid_v1.py) */
void eo_compute_torque(const EoModel *m, const EoCfg *c, const EoData *d, const double *ctrl, double *torque);
void eo_env_set_state(const EoModel *m, EoEnv *e, const double *qpos, const double *qvel);
void eo_env_reset(const EoModel *m, const EoCfg *c, const EoExpert *x, EoEnv *e, int take, int start);
void eo_env_step(const EoModel *m, const EoCfg *c, const EoExpert *x, EoEnv *e, const double *action,
                 int *fail, int *end);
void eo_env_obs(const EoModel *m, const EoEnv *e, double *obs);
void eo_body_quat(const EoModel *m, const double *qpos, double *bquat);
void eo_ee_pos(const EoModel *m, const EoData *d, int heading, double *ee);
double eo_reward(const EoModel *m, const EoCfg *c, const EoExpert *x, const EoEnv *e, int end, double *info5);
void eo_expert_features(const EoModel *m, int L, const double *qpos, double dt, double *rows,
                        double *head_z_min);

/* math helpers (utils/math.py, utils/transformation.py) exposed for golden checks */
void eo_quat_mul(const double *q1, const double *q0, double *out);
void eo_quat_inv(const double *q, double *out);
void eo_quat_from_euler(double ai, double aj, double ak, double *q);
void eo_de_heading(const double *q, double *out);
void eo_transform_vec(const double *v, const double *q, int heading, double *out);
void eo_rotation_from_quat(const double *q, double *axis, double *angle);
void eo_qvel_fd6(const double *cur_qpos, const double *next_qpos, double dt, int heading, double *out6);
void eo_qvel_fd(int nq, const double *cur_qpos, const double *next_qpos, double dt, int transform, double *out);
void eo_angvel_fd(int nquat, const double *prev, const double *cur, double dt, double *out);

/* policy + rollout driver (agents/agent.py:29-76 batched over independent envs) */
void eo_policy_mean(const EoPolicy *p, const double *x, double *mean, double *scratch);
int eo_rollout(const EoModel *m, const EoCfg *c, const EoExpert *x, const EoPolicy *p,
               int n_env, int T, int max_resets, const int *reset_take, const int *reset_start,
               const double *eps, const unsigned char *mean_flag,
               const double *zf_mean, const double *zf_std, double zf_clip,
               double *states, double *actions, double *rewards, double *masks, double *next_states,
               double *exps, int *v_metas, double *c_info, double *raw_obs, double *final_qpos,
               double *final_qvel, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
