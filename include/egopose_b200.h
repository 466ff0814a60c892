/* egopose_b200 C ABI - the drop-in boundary of the B200-native PPO rollout-and-update hot path.
 *
 * The reference (Khrylx/EgoPose, pure Python) has no FFI of its own; its native seams on this path
 * are the third-party C ABIs it reaches through mujoco_py / scipy / torch:
 *   mj_step / mj_forward / mj_fullM    envs/common/mujoco_env.py:22-23,100-101, ego_pose/envs/humanoid_v1.py:134,174
 *   LAPACK dpotrf/dpotrs               ego_pose/envs/humanoid_v1.py:143 (scipy cho_factor/cho_solve)
 *   ATen addmm/relu/normal_            agents/agent.py:44-47 (policy forward + Gaussian sample, batch 1)
 *   python GAE loop                    core/common.py:5-25
 *   autograd + torch.optim.Adam        agents/agent_pg.py:19-26, agents/agent_ppo.py:44-65
 * Each entry point below names the reference code it replaces.  The Python host mirror of the
 * Agent/Policy/Value/TrajBatch API (egopose_b200/) binds these with ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions:
 *   - every pointer argument named d_* is a CUDA device pointer owned by the caller; the library
 *     allocates only inside *_create / *_upload and frees only in *_destroy
 *   - all compute entry points are asynchronous on `stream` (a cudaStream_t passed as void*)
 *   - return 0 on success, negative EGP_E* on failure; egp_last_error_string() describes the last
 *     failure of the calling thread; nothing throws or aborts across the boundary
 *   - arithmetic is float64 ("_f64"), the reference dtype (ego_pose/ego_mimic.py:31-32)
 */
#ifndef EGOPOSE_B200_H
#define EGOPOSE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGP_OK 0
#define EGP_EINVAL (-1)
#define EGP_ECUDA (-2)
#define EGP_ENOMEM (-3)
#define EGP_ESIZE (-4)

#define EGP_MAX_BODY 24
#define EGP_MAX_DOF 64
#define EGP_MAX_CHAIN 12
#define EGP_NEE 5

/* packed expert row, in doubles (derived from the dict written by ego_pose/data_process/gen_expert.py:28-83) */
#define EGP_X_QPOS 0
#define EGP_X_QVEL 59
#define EGP_X_RLINV_LOCAL 117
#define EGP_X_RANGV 120
#define EGP_X_RQ_RMH 123
#define EGP_X_EE_POS 127
#define EGP_X_BQUAT 142
#define EGP_X_BANGVEL 226
#define EGP_X_STRIDE 292

typedef struct EgpModel EgpModel;       /* opaque: model + PD/reward constants + expert tables on one device */

/* What mujoco_py.load_model_from_path gives the reference (envs/common/mujoco_env.py:22) plus the cfg
 * constants of ego_pose/utils/egomimic_config.py:94-122.  Host pointers, copied during create. */
typedef struct {
    int32_t nq, nv, nu, nbody;
    double timestep;
    double gravity[3];
    const int32_t *body_parent, *body_dofadr, *body_dofnum, *body_qposadr;      /* [nbody] */
    const double *body_pos, *body_mass, *body_ipos, *body_inertia;              /* [nbody][3],[nbody],[nbody][3],[nbody][6] */
    const int32_t *dof_body, *dof_parent;                                      /* [nv] */
    const double *dof_armature, *dof_axis, *dof_anchor;                        /* [nv],[nv][3],[nv][3] */
    int32_t ee_body[EGP_NEE];           /* LeftFoot RightFoot LeftHand RightHand Head (humanoid_v1.py:100) */
    int32_t head_body;
    /* cfg */
    int32_t frame_skip;                 /* 15, humanoid_v1.py:16 */
    const double *jkp, *jkd, *a_ref, *a_scale, *torque_lim;                    /* [nu] */
    const double *b_diffw;                                                     /* [nbody-1] */
    double w_p, w_v, w_e, w_rp, w_rv, k_p, k_v, k_e, k_rh, k_rq, k_rl, k_ra;   /* reward_function.py:8-12 */
    int32_t v_ord, decay;
} EgpModelDesc;

/* Per-rollout settings (the mutable parts of HumanoidEnv / Agent state). */
typedef struct {
    int32_t n_env, horizon;             /* E environments x T recorded steps each, rows n = e*T + t */
    int32_t episode_len;                /* cfg.env_episode_len, humanoid_v1.py:197 */
    int32_t fr_margin;                  /* cfg.fr_margin, humanoid_v1.py:214 */
    double end_reward;                  /* env.end_reward, ego_mimic.py:112 */
    double fix_head_lb;                 /* NaN: expert head_height_lb - 0.1 (humanoid_v1.py:193-196) */
    double noise_rate;                  /* Agent.noise_rate, agents/agent.py:46 */
    int32_t mean_action;                /* Agent.mean_action */
    double zf_clip;                     /* ZFilter clip (5), <=0 disables; stats frozen during a rollout */
    uint64_t seed;                      /* Philox key for perf-mode noise / reset draws */
    uint64_t iteration;                 /* Philox stream offset so successive rollouts differ */
    int32_t max_resets;                 /* parity mode: columns of d_reset_take/start */
    /* Evaluation roll-out (ego_pose/ego_mimic_eval.py:102-177): a failing environment is NOT restarted; its state is
     * replaced in place by the d_state_pred row of the next frame, aligned to the simulated root position / heading
     * (reset_env_state :93-99, utils/tools.py:71-75), and the episode clock keeps running ('naivefs' fail-safe :167-173
     * with fix_head_lb = 0.3, :52-53).  The first state of every episode is placed the same way (:127).
     * 1 = fail rule of the environment ('naivefs'), 2 = value rule ('valuefs', see EgpRolloutIn.value_net). */
    int32_t eval_mode;
} EgpRolloutCfg;

/* PolicyGaussian(MLP) weights on the device, torch nn.Linear layout [out][in] (core/policy_gaussian.py:8-24,
 * models/mlp.py:5-25); two hidden relu layers. */
typedef struct {
    int32_t in_dim, h1, h2, out_dim;
    const double *d_W1, *d_b1, *d_W2, *d_b2, *d_W3, *d_b3, *d_log_std;
} EgpPolicyWeights;

/* Optional inputs; NULL selects the in-kernel Philox path. */
typedef struct {
    const double *d_eps;                /* [E*T][nu] standard-normal noise (parity mode) */
    const int32_t *d_reset_take;        /* [E][max_resets] pre-drawn expert_ind (humanoid_v1.py:210) */
    const int32_t *d_reset_start;       /* [E][max_resets] pre-drawn start_ind (humanoid_v1.py:214) */
    const uint8_t *d_mean_flag;         /* [E*T] 1 = mean action this step (agents/agent.py:46) */
    const double *d_zf_mean, *d_zf_std; /* [S] frozen ZFilter statistics or NULL (identity) */
    /* Optional per-rollout video context table on the device, replacing the per-frame table uploaded with the
     * experts.  ctx_mode 0: row = take_off[take] + start + t (per frame).  ctx_mode 1: row = (win_off[take] +
     * start - fr_margin) * ctx_T + t: one test-mode VideoStateNet output v_out[t] per (take, start) episode window
     * (models/video_state_net.py:36-39,61-64), ctx_T = env_episode_len rows per window. */
    const double *d_ctx;
    const int32_t *d_win_off;           /* [n_takes + 1], ctx_mode 1 / 2 */
    int32_t ctx_dim, ctx_mode, ctx_T;
    /* ctx_mode 2: ONE row per (take, start) window, row = win_off[take] + start - fr_margin, constant over the
     * episode: test-mode VideoForecastNet v_out (models/video_forecast_net.py:58-59, causal LSTM over the fr_margin
     * frames before the episode start).
     * Optional state LSTM stepped once per env step on the filtered state (VideoForecastNet.s_net in 'step' mode,
     * models/video_forecast_net.py:60-61,89-93 + models/rnn.py:22-26,36-43): h, c = 0 at every episode start, the
     * policy input becomes cat(ctx row, h_t) instead of cat(ctx row, state).  snet_hdim = 0 disables. */
    const double *d_snet_W;             /* [4H][S + H]: row 4u + g = cat(weight_ih[g*H + u], weight_hh[g*H + u]), g = i,f,g,o */
    const double *d_snet_b;             /* [4H] bias_ih + bias_hh in the same row order */
    double *d_snet_state;               /* caller-owned scratch, ceil(E / 32) * 2H * 32 doubles */
    int32_t snet_hdim;                  /* H (even) */
    const int32_t *d_fix_len;           /* [E] per-environment episode length (env.set_fix_sampling(len=), humanoid_v1.py:197) or NULL */
    const double *d_state_pred;         /* eval_mode: [total_frames][S] predicted observations (qpos[2:] | qvel in the
                                         * heading frame, humanoid_v1.py:73-96), row = take_off[take] + start + cur_t */
    /* eval_mode 2 ('valuefs', ego_mimic_eval.py:156-159,167): the value net is evaluated on cat(value context row, state)
     * before every step, its output is pushed into a running mean (utils/zfilter.py:18-27) and the state is replaced when
     * value < 0.6 * mean.  value_net: HOST pointer to the Value(MLP) weights (same in_dim / hidden sizes as the policy,
     * out_dim 1); d_vctx: value_vs_net's context table (same shape / indexing as the policy's, NULL = the policy's);
     * d_value_stat: [E][2] (n, mean) read at launch and written back, so successive launches continue one statistic. */
    const EgpPolicyWeights *value_net;
    const double *d_vctx;
    double *d_value_stat;
    const double *d_init_qpos, *d_init_qvel;    /* [E][nq], [E][nv] or NULL: simulator state set right after the first reset of
                                         * every environment (env.set_state of the ego-mimic prediction, ego_forecast_eval.py:119-121) */
} EgpRolloutIn;

/* TrajBatchEgo layout (core/trajbatch.py:6-16, ego_pose/core/trajbatch_ego.py:7-9), all device, row-major. */
typedef struct {
    double *d_states;                   /* [N][S] */
    double *d_actions;                  /* [N][nu] */
    double *d_masks;                    /* [N] */
    double *d_next_states;              /* [N][S] or NULL (never read by the update, agent_ego.py:37-42) */
    double *d_rewards;                  /* [N] */
    double *d_exps;                     /* [N] */
    int32_t *d_v_metas;                 /* [N][2] (expert_ind, start_ind) */
    double *d_c_info;                   /* [N][5] or NULL */
    double *d_raw_obs;                  /* [N][S] unfiltered observations or NULL */
    double *d_final_qpos, *d_final_qvel;/* [E][nq], [E][nv] or NULL */
    double *d_logger;                   /* [EGP_LOG_SIZE] reductions for core/logger_rl.py, or NULL */
    double *d_values;                   /* [N] value-net output per step (value_net given) or NULL */
    double *d_qpos_traj, *d_qvel_traj;  /* [N][nq], [N][nv] simulator state BEFORE step t (traj_pred / vel_pred of
                                         * ego_mimic_eval.py:136-138) or NULL */
} EgpTrajOut;

/* d_logger slots */
#define EGP_LOG_NUM_STEPS 0
#define EGP_LOG_NUM_EPISODES 1
#define EGP_LOG_TOTAL_REWARD 2      /* env reward is the constant 1.0 -> episode lengths (humanoid_v1.py:192) */
#define EGP_LOG_TOTAL_C_REWARD 3
#define EGP_LOG_MIN_C_REWARD 4
#define EGP_LOG_MAX_C_REWARD 5
#define EGP_LOG_C_INFO 6            /* 5 slots */
#define EGP_LOG_MIN_EPISODE_REWARD 11
#define EGP_LOG_MAX_EPISODE_REWARD 12
#define EGP_LOG_NUM_NAN_RESETS 13
#define EGP_LOG_NUM_FAILSAFE_RESETS 14  /* eval_mode: in-place state replacements (num_reset, ego_mimic_eval.py:170) */
#define EGP_LOG_SIZE 16

const char *egp_last_error_string(void);
int egp_version(void);

/* --- model / expert tables ------------------------------------------------------------------- */
int egp_model_create(const EgpModelDesc *desc, int device, EgpModel **out);
void egp_model_destroy(EgpModel *m);

/* Joint limits (humanoid_1205_v1.xml:10 limited="true", :28-30 range=...) as MuJoCo soft constraints inside mj_step
 * (what sim.step() adds to the smooth dynamics when a hinge leaves its range): range [nv][2] radians per dof (lower >=
 * upper: none; the free root is ignored), invweight0 [nv] = diag(M^-1) at qpos0 (mjModel.dof_invweight0), solref
 * (timeconst, dampratio) and solimp (d0, dwidth, width, midpoint, power), NULL = MuJoCo's defaults 0.02 1 /
 * 0.9 0.95 0.001 0.5 2.  Host pointers.  range = NULL switches the limits off again (smooth dynamics: the default).
 * The block-sweep rollout kernel carries the rows for the plain policy plan (training and evaluation roll-outs); the
 * state LSTM, the value rule and the chunked wide-policy plan fall back to the one-warp kernel or return EGP_ESIZE.
 * Floor contact: egp_model_set_contacts. */
int egp_model_set_joint_limits(EgpModel *m, const double *range, const double *invweight0, const double *solref,
                               const double *solimp);

/* Floor contact (humanoid_1205_v1.xml:21 plane z = 0 with condim 3, friction 1; :11 geom margin 0.001) inside mj_step,
 * MuJoCo's soft-constraint model with pyramidal friction cones: one collision geom per body, geom_type [nbody] (0 sphere |
 * 1 capsule | 2 box), geom_size [nbody][3] (radius | radius | half extents), geom_p0 / geom_p1 [nbody][3] (centre, or the
 * capsule's two end points, in the body frame), body_invweight0 [nbody][2] (mjModel.body_invweight0: translational,
 * rotational), margin, sliding friction; solref / solimp as above (one set is shared with the limit rows).  Host pointers;
 * geom_type = NULL switches contacts off (the default).  Only geom-floor pairs are generated: contacts BETWEEN body geoms
 * are not modelled.  Same kernel restriction as the joint limits. */
/* diagnostics: constrained solves (one environment, one sub-step) since the last reset that left the active-set loop at its
 * cap of 100 passes without a fixed point; synchronises the device */
int64_t egp_cons_cap_hits(int reset);
/* diagnostics of the block-sweep kernel: out2[0] = sub-steps, out2[1] = backward + forward passes of the active-set iteration
 * they took (counted by the first CTA of every launch) */
int egp_cons_passes(int64_t *out2, int reset);
int egp_model_set_contacts(EgpModel *m, const int32_t *geom_type, const double *geom_size, const double *geom_p0,
                           const double *geom_p1, const double *body_invweight0, double margin, double friction,
                           const double *solref, const double *solimp);

/* replaces HumanoidEnv.load_experts (humanoid_v1.py:45-54): packed rows [total_frames][EGP_X_STRIDE],
 * take offsets [n_takes+1], per-take head_height_lb, optional per-frame context rows (host pointers) */
int egp_expert_upload(EgpModel *m, int n_takes, const int32_t *take_off, const double *rows,
                      const double *head_height_lb, const double *ctx, int ctx_dim);

/* replaces ego_pose/data_process/gen_expert.py:28-83 for one take: d_qpos [L][nq] -> d_rows [L][EGP_X_STRIDE],
 * d_head_z [L] (head height per frame; min over frames = head_height_lb) */
int egp_expert_features_f64(EgpModel *m, int L, const double *d_qpos, double *d_rows, double *d_head_z,
                            void *stream);
/* same, plus the expert-dict keys that are not rollout inputs (gen_expert.py:44,49-50,79): d_extras [L][EGP_XE_STRIDE]
 * = head_pos 3 | com 3 | ee_wpos 15 (end-effector body positions in the world frame, get_ee_pos(None)); NULL to skip */
#define EGP_XE_HEAD_POS 0
#define EGP_XE_COM 3
#define EGP_XE_EE_WPOS 6
#define EGP_XE_STRIDE 21
int egp_expert_features_ex_f64(EgpModel *m, int L, const double *d_qpos, double *d_rows, double *d_head_z,
                               double *d_extras, void *stream);

/* --- physics, single calls (parity/debug; replaces sim.forward()/sim.step()/compute_torque) ---- */
/* one mj_forward at (qpos, qvel, ctrl): qfrc_bias [n][nv], xpos [n][nbody][3], qacc = M^-1 (ctrl - bias) [n][nv]
 * (d_ctrl [n][nu] may be NULL = zero actuation) */
int egp_forward_debug_f64(EgpModel *m, int n, const double *d_qpos, const double *d_qvel, const double *d_ctrl,
                          double *d_bias, double *d_xpos, double *d_qacc, void *stream);
/* env.step(action) for n independent states that have just been reset-forwarded (humanoid_v1.py:158-199):
 * d_qpos/d_qvel in-out [n][nq]/[n][nv]; d_action [n][nu]; outputs obs [n][S], head_z [n] (stale xpos) */
int egp_env_step_debug_f64(EgpModel *m, int n, double *d_qpos, double *d_qvel, const double *d_action,
                           double *d_obs, double *d_head_z, double *d_torque0, void *stream);

/* --- fused rollout (K1+K2+K3): replaces Agent.sample / sample_worker (agents/agent.py:29-111) ---- */
int egp_rollout_f64(EgpModel *m, const EgpPolicyWeights *pol, const EgpRolloutCfg *cfg, const EgpRolloutIn *in,
                    const EgpTrajOut *out, void *stream);

/* --- PPO update kernels ---------------------------------------------------------------------- */
/* K4: core/common.py:5-21 estimate_advantages reverse scan over the flat batch.  d_adv gets the
 * UN-normalised advantages, d_ret = values + adv, d_stats[0..2] = (n, mean, M2) of adv so that
 * (adv - mean) / sqrt(M2 / (n - 1)) is the reference's standardisation (:22).
 * d_work: egp_gae_work_bytes(n) bytes of scratch, 16-byte aligned. */
int64_t egp_gae_work_bytes(int64_t n);
int egp_gae_f64(const double *d_rewards, const double *d_masks, const double *d_values, double gamma, double tau,
                int64_t n, double *d_adv, double *d_ret, double *d_stats, void *d_work, void *stream);
/* Batches of at least this many samples take the one-pass scan (tiles handed out by ticket from the end of the batch,
 * each folds the published maps of the tiles after it: 5 doubles of HBM traffic per sample); smaller ones the two-pass
 * scan whose second read is an L2 hit.  Results are bit-identical.  Returns the previous value; n < 0 only queries. */
int64_t egp_gae_set_onepass_min(int64_t n);
/* (x - stats.mean) / std in place (materialises the reference's normalised advantages) */
int egp_standardize_f64(double *d_x, int64_t n, const double *d_stats, void *stream);

/* core/distributions.py:21-22: logp[n] = sum_j log N(a | mu, exp(log_std)) */
int egp_gauss_logp_f64(const double *d_mu, const double *d_actions, const double *d_log_std, int64_t n, int adim,
                       double *d_logp, void *stream);

/* K5: agents/agent_ppo.py:58-65 ppo_loss forward + backward wrt mu (and log_std) in one pass.
 * d_stats as written by egp_gae (advantages are standardised on the fly); inv_count = 1 / #(exps != 0)
 * over the GLOBAL batch.  d_dmu [n][adim] = dL/dmu; d_dlogstd [adim] (may be NULL) and d_loss[0] are
 * ACCUMULATED with atomics - zero them first. */
int egp_ppo_loss_grad_f64(const double *d_mu, const double *d_actions, const double *d_log_std,
                          const double *d_adv, const double *d_stats, const double *d_logp0, const double *d_exps,
                          double clip_eps, double inv_count, int64_t n, int adim, double *d_dmu, double *d_dlogstd,
                          double *d_loss, void *stream);

/* agents/agent_pg.py:22-23: L = mean((V - R)^2); d_dv = 2 (V - R) * inv_n ; d_loss[0] accumulated */
int egp_value_loss_grad_f64(const double *d_v, const double *d_ret, double inv_n, int64_t n, double *d_dv,
                            double *d_loss, void *stream);

/* fused bias + relu forward: y = relu(y + b) in place, y [n][dim]; and backward mask dy *= (y > 0) */
int egp_bias_relu_f64(double *d_y, const double *d_b, int64_t n, int dim, void *stream);
int egp_relu_bwd_f64(double *d_dy, const double *d_y, int64_t n, int dim, void *stream);
/* row gather out[i][:] = in[perm[i]][:] (the mini-batch shuffle of agents/agent_ppo.py:26-32) */
int egp_gather_rows_f64(const double *d_in, const int64_t *d_perm, int64_t n, int dim, double *d_out, void *stream);
/* fused: dy *= (y > 0) in place and out[dim] = column sums of the masked dy (bias gradient), one pass */
int egp_relu_bwd_colsum_f64(double *d_dy, const double *d_y, int64_t n, int dim, double *d_out, void *stream);
/* column sums (bias gradients): out[dim] = sum_n x[n][dim] */
int egp_colsum_f64(const double *d_x, int64_t n, int dim, double *d_out, void *stream);
/* shifted column moments for the batched ZFilter update (utils/zfilter.py:18-27 merged per rollout):
 * out[0..dim) = sum_n (x - shift), out[dim..2dim) = sum_n (x - shift)^2 ; d_shift [dim] may be NULL (0) */
int egp_col_moments_f64(const double *d_x, int64_t n, int dim, const double *d_shift, double *d_out, void *stream);
/* x[n][S] with per-row context gather: out[n][ctx_dim + S] = cat(ctx[frame(n)], states[n])
 * (models/video_state_net.py:62-64 'cat(v_out[t], state)'); frame(n) = take_off[v_meta[n][0]] + v_meta[n][1] + t(n) */
int egp_build_input_f64(EgpModel *m, const double *d_states, const int32_t *d_v_metas, const double *d_masks,
                        int64_t n, int horizon, double *d_x, void *stream);

/* K6: torch.nn.utils.clip_grad_norm_ (agents/agent_ppo.py:53-56) + torch.optim.Adam step
 * (ego_pose/ego_mimic.py:70-77) over one flat parameter buffer.  d_norm2: 1 double of scratch. */
int egp_sumsq_f64(const double *d_g, int64_t n, double *d_norm2, void *stream);
int egp_adam_step_f64(double *d_p, const double *d_g, double *d_m, double *d_v, int64_t n, double lr, double beta1,
                      double beta2, double eps, int64_t step, double max_norm, const double *d_norm2, void *stream);

/* --- float64 dense layers on the int8 tensor cores (Ozaki scheme, egopose_b200/csrc/ozaki.cu) --------------
 * Replace the cuBLAS DGEMMs behind models/mlp.py:22-25 / core/policy_gaussian.py:19-24 / core/critic.py:15-18 forward
 * and backward in agents/agent_pg.py:19-26 and agents/agent_ppo.py:44-51.
 * Slices: X = round(x 2^(7S-1-e)) (e constant along the contraction) as S signed base-128 digits in [-64, 64], most
 * significant first: x = 2^(e+1-7S) * sum_{t=1..S} q_t 128^(S-t).  (egp_oz_radix_bits() returns 7; a -DOZ_RADIX_BITS=8
 * build stores two's-complement bytes instead - an experiment that turned out less accurate, see csrc/ozaki.cuh.) */
int egp_oz_radix_bits(void);
/* d_x [m][k] (leading dimension ldx) -> d_out [S][m][kp] (kp >= k, multiple of 16, zero padded), d_exps [m];
 * d_colmax (optional, [k] doubles zeroed by the caller) receives the column abs-max of x (for the colsT slicing) */
int egp_oz_slice_rows_f64(const double *d_x, int64_t m, int k, int64_t ldx, int n_slices, int8_t *d_out, int kp,
                          int32_t *d_exps, double *d_colmax, void *stream);
/* column abs-max only: d_colmax [f] (zeroed by the caller) */
int egp_oz_colmax_f64(const double *d_x, int64_t n, int f, int64_t ldx, double *d_colmax, void *stream);
/* d_x [n][f] -> transposed column-scaled slices d_out [S][f + ones_row][np] (np >= n, multiple of 16), d_exps [f + ones_row]
 * from d_colmax [f]; ones_row = 1 appends the constant feature 1.0 (its products are column sums = bias gradients) */
int egp_oz_slice_cols_t_f64(const double *d_x, int64_t n, int f, int64_t ldx, int n_slices, const double *d_colmax,
                           int8_t *d_out, int64_t np, int32_t *d_exps, int ones_row, void *stream);
/* C [m][n] (ldc) = A B^T (+ bias[n], relu; zeroed where d_mask [m][ldm] <= 0) from slices A [S][m][kp], B [S][n][kp] and
 * their exponents.  Long contractions with few output tiles (weight gradients) run split-K through d_work
 * (egp_oz_gemm_work_bytes; bias / relu / mask are not available there). */
int64_t egp_oz_gemm_work_bytes(int64_t m, int n, int64_t kp, int n_slices);
int egp_oz_gemm_f64(const int8_t *d_a, const int32_t *d_ea, int64_t m, const int8_t *d_b, const int32_t *d_eb, int n,
                    int64_t kp, int n_slices, const double *d_bias, int relu, const double *d_mask, int64_t ldm,
                    double *d_c, int64_t ldc, void *d_work, int64_t work_bytes, void *stream);
/* The same product; additionally the epilogue records the abs-maxima of the FINAL output (after bias / relu / mask) for
 * the slicers of C: d_rowmax [m] = high 32 bits of max_n |C[m][n]|, d_colmax [n] = bit pattern (high word << 32) of
 * max_m |C[m][n]| - both optional, combined with atomicMax, so the caller zeroes them; not available with split-K. */
int egp_oz_gemm_max_f64(const int8_t *d_a, const int32_t *d_ea, int64_t m, const int8_t *d_b, const int32_t *d_eb, int n,
                        int64_t kp, int n_slices, const double *d_bias, int relu, const double *d_mask, int64_t ldm,
                        double *d_c, int64_t ldc, uint32_t *d_rowmax, double *d_colmax, void *d_work, int64_t work_bytes,
                        void *stream);
/* BOTH orientations from one read of d_x [n][f], given its recorded maxima (egp_oz_gemm_max_f64, or d_colmax from
 * egp_oz_slice_rows_f64 / egp_oz_colmax_f64 and d_rowmax = high words of the row maxima): row slices d_out_rows [S][n][kp]
 * (kp >= f, multiple of 32) + d_exps_rows [n] exactly as egp_oz_slice_rows_f64 writes them, transposed slices
 * d_out_t [S][f + ones_row][np] + d_exps_t [f + ones_row] exactly as egp_oz_slice_cols_t_f64 writes them. */
int egp_oz_slice_both_f64(const double *d_x, int64_t n, int f, int64_t ldx, int n_slices, const uint32_t *d_rowmax,
                          const double *d_colmax, int8_t *d_out_rows, int kp, int32_t *d_exps_rows, int8_t *d_out_t, int64_t np,
                          int32_t *d_exps_t, int ones_row, void *stream);

/* --- chunked MLP forward / loss / backward on the int8 tensor cores (egopose_b200/csrc/oz_mlp.cu) ------------
 * One optimisation step's forward + loss + backward of a two-hidden-layer relu MLP (models/mlp.py:22-25 trunk,
 * core/policy_gaussian.py:19-24 or core/critic.py:15-18 head), i.e. the autograd graph of agents/agent_pg.py:19-26
 * (value) and agents/agent_ppo.py:44-51,58-65 (policy), in row chunks whose intermediates stay L2-resident. */
typedef struct {
    int32_t in_dim, h1, h2, out_dim;
    const double *d_W1, *d_b1, *d_W2, *d_b2, *d_W3, *d_b3;      /* torch nn.Linear layout [out][in] */
    double *d_gW1, *d_gb1, *d_gW2, *d_gb2, *d_gW3, *d_gb3;      /* gradients, same layouts (NULL for forward only) */
    /* optional: dL/dx[:, :dx_cols] -> d_dx [n][dx_cols] (the gradient reaching a learned video-context net whose output
     * occupies the first input columns: models/video_state_net.py:62-69, agent_ego.py:28-32); NULL / 0 to skip */
    double *d_dx;
    int32_t dx_cols;
} EgpMlpNet;

typedef struct {
    int32_t kind;                       /* 0 forward only, 1 PPO clipped surrogate (agent_ppo.py:58-65), 2 value MSE (agent_pg.py:22-23) */
    /* kind 1: per-sample arrays over the whole batch, see egp_ppo_loss_grad_f64 */
    const double *d_actions, *d_log_std, *d_adv, *d_stats, *d_logp0, *d_exps;
    double clip_eps, inv_count;
    double *d_dlogstd;                  /* [out_dim] accumulated, or NULL */
    /* kind 2 */
    const double *d_returns;
    double inv_n;
    double *d_loss;                     /* [1] accumulated */
    /* kind 1: 1 = this pass runs at the parameters that define fixed_log_probs (agent_ppo.py:18-20, first epoch): logp is
     * WRITTEN to d_logp0 and the ratio is 1, so no separate forward pass for the fixed log-probs is needed */
    int32_t init_logp0;
} EgpMlpLoss;

int64_t egp_oz_mlp_chunk_rows(void);   /* 128 rows x number of SMs */
/* on = 1 (default): the GEMM epilogues / the value head's backward kernel record the row and column abs-maxima of their
 * outputs, and every intermediate of egp_oz_mlp_step_f64 is sliced in BOTH orientations from one read; 0: separate row,
 * column-maximum and transposed passes (bit-identical results); negative: query.  Returns the setting in force. */
int egp_oz_mlp_set_fused_slicing(int on);
int64_t egp_oz_mlp_work_bytes(int in_dim, int h1, int h2, int out_dim, int64_t chunk_rows, int n_slices);
int64_t egp_oz_mlp_xcache_bytes(int in_dim, int64_t n, int64_t chunk_rows, int n_slices);
/* d_x [n][in_dim] (leading dimension ldx).  kind 0 writes y [n][out_dim] to d_y; kinds 1 / 2 write the six gradients
 * (overwritten) and optionally y.  d_xcache / xcache_state: 0 none, 1 fill while running, 2 reuse (same x, n, chunk_rows,
 * n_slices as when it was filled): the int8 slices of the constant input are then read instead of recomputed. */
int egp_oz_mlp_step_f64(const EgpMlpNet *net, const double *d_x, int64_t ldx, int64_t n, const EgpMlpLoss *loss, double *d_y,
                        int n_slices, int64_t chunk_rows, void *d_xcache, int xcache_state, void *d_work, int64_t work_bytes,
                        void *stream);

/* --- fused LSTM sequence recurrence (csrc/lstm.cu) ------------------------------------------------------------------
 * Replaces the per-step LSTMCell loops of models/rnn.py:45-61 (RNN.batch_forward: both directions of the BiLSTM of
 * models/video_state_net.py:36-70) and the state-LSTM unroll of models/video_forecast_net.py:95-111 in the PPO update.  The
 * non-recurrent GEMMs (input projection, weight gradients) stay with the caller.  Rows are "time-major packed": step s owns
 * rows [off[s], off[s+1]) = the batch elements alive at step s (counts non-increasing; dense [L, B]: off[s] = s B); d_off is a
 * DEVICE array of L + 1 int64.  Gate order i | f | g | o (torch.nn.LSTMCell).  hidden size H = 64 or 128. */
int64_t egp_lstm_wfrag_elems(int H);    /* doubles per packed W_hh operand (4 H H) */
/* d_Whh [4H][H] (torch weight_hh) -> tensor-core fragment order for the forward (h W_hh^T) and backward (dG W_hh) products */
int egp_lstm_pack_whh_f64(const double *d_Whh, int H, double *d_Wf_fwd, double *d_Wf_bwd, void *stream);
/* d_xi [Np][4H] = x W_ih^T + b_ih + b_hh  ->  d_h [Np][H], d_gates [Np][4H] (post-activation, saved for backward), d_c [Np][H] */
int egp_lstm_seq_fwd_f64(const double *d_xi, const int64_t *d_off, int L, int64_t B, int H, const double *d_Wf_fwd, double *d_h,
                         double *d_gates, double *d_c, void *stream);
/* d_dh [Np][H] upstream gradient of every h  ->  d_dxi [Np][4H] gradient of the gate pre-activations (dW_hh = d_dxi^T H_prev,
 * dW_ih = d_dxi^T X, db = column sums of d_dxi are GEMMs of the caller) */
int egp_lstm_seq_bwd_f64(const double *d_dh, const double *d_gates, const double *d_c, const int64_t *d_off, int L, int64_t B, int H,
                         const double *d_Wf_bwd, double *d_dxi, void *stream);

/* --- gradient exchange over NVLink peer memory ------------------------------------------------------
 * The data-parallel update differentiates one global batch (agents/agent_ppo.py:44-56): rank-local gradients of the flat
 * [value | policy] buffer are summed once per PPO epoch.  One process per GPU of a node; every rank owns an exchange block
 * (cudaMalloc, shared through CUDA IPC): src [n] (write the local gradient here), out [n] (the sum, bit-identical on every
 * rank: every rank adds the peers' src in rank order, reading them over NVLink).  egp_allreduce_grads_f64 is ONE kernel
 * launch per rank and call: cross-GPU flag barrier, peer-to-peer sum, flag barrier (after completion nobody reads this
 * rank's src any more).  All ranks must make the same sequence of calls.
 *   egp_comm_create  : allocates the block, returns the IPC handle (egp_comm_handle_bytes() bytes) to hand to the peers
 *   egp_comm_connect : all_handles = the world handles in rank order (exchanged by the caller, e.g. all_gather)
 *   egp_comm_error   : 0, or which barriers timed out (bit 0: a peer's gradient never became ready, bit 1: a peer never
 *                      finished reading; a rank did not arrive within ~4 s); synchronises the device */
typedef struct EgpComm EgpComm;
int64_t egp_comm_handle_bytes(void);
int egp_comm_create(int rank, int world, int device, int64_t n, EgpComm **out, void *handle_out);
int egp_comm_connect(EgpComm *c, const void *all_handles);
int egp_comm_connect_local(EgpComm *c, EgpComm *const *all);   /* ranks that live in this process: all = the world communicators in rank order */
int64_t egp_comm_set_timeout_cycles(int64_t cycles);            /* barrier timeout (SM cycles, default 8e9); returns the previous value; <= 0 only queries; per device */
double *egp_comm_src(EgpComm *c);
double *egp_comm_out(EgpComm *c);
int egp_allreduce_grads_f64(EgpComm *c, int64_t n, void *stream);
int egp_comm_error(EgpComm *c);
void egp_comm_destroy(EgpComm *c);

#ifdef __cplusplus
}
#endif
#endif
