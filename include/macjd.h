/* macjd.h -- C ABI of the B200-native MA-CJD hot path (libmacjd_b200.so).
 *
 * The reference (mfathulkr/MA-CJD-Cooperative-Jamming-Decision-Making-via-MARL) has no
 * native / FFI layer: its seam is the duck-typed Python API that
 * runners/episode_runner.py and main.py call.  Each entry point below is what a
 * binding for that seam would call; the reference interface it replaces is cited as
 * file:line (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name ends in _host; buffers are owned by the caller, kernels never allocate
 *   - all work is enqueued on ctx->stream of device ctx->device and returns
 *     immediately (stream-ordered); no global mutable state, re-entrant for
 *     disjoint buffers
 *   - return value: 0 = MACJD_OK, negative = error (macjd_status_string()); the
 *     library never throws, aborts or prints
 *   - optional outputs may be NULL
 */
#ifndef MACJD_H_
#define MACJD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MACJD_ABI_VERSION 1

#if defined(__GNUC__)
#define MACJD_API __attribute__((visibility("default")))
#else
#define MACJD_API
#endif

enum macjd_status {
  MACJD_OK = 0,
  MACJD_ERR_INVALID_ARG = -1,  /* NULL / out-of-range argument                         */
  MACJD_ERR_UNSUPPORTED = -2,  /* dims outside what the kernels were built for          */
  MACJD_ERR_CUDA = -3,         /* a CUDA runtime call failed; see macjd_last_cuda_error */
  MACJD_ERR_WORKSPACE = -4     /* caller-provided workspace too small                   */
};

typedef struct macjd_ctx {
  int32_t device;    /* CUDA device ordinal the buffers live on */
  int32_t reserved;
  void* stream;      /* cudaStream_t (NULL = legacy default stream) */
} macjd_ctx;

MACJD_API const char* macjd_status_string(int status);
MACJD_API const char* macjd_last_cuda_error(void);   /* thread-local text of the last CUDA failure */
MACJD_API int macjd_abi_version(void);
/* sizeof() of every struct below, for binding self-checks: index = order of declaration */
MACJD_API size_t macjd_abi_sizeof(int which);

/* ===================================================================== environment
 * Replaces simulation/environment.py:29-573 (ElectromagneticEnvironment) together with
 * core/radar.py, core/jammer.py, utils/math_utils.py, utils/state_utils.py for a batch
 * of n_envs independent episodes.
 */
typedef struct macjd_env_tables {
  int32_t n_envs;
  int32_t n_jammers;        /* J = n_agents                                        */
  int32_t n_radars;         /* R (<= 64); n_actions = 2R + 1                       */
  int32_t n_targets;        /* K protected targets (reference: 1)                  */
  int32_t n_types;          /* max_radar_types of the one-hot block                */
  int32_t episode_limit;
  /* scenario tables, float64: element (row, env) at data[row*row_stride + env*env_stride].
   * Per-env tables: row_stride = n_envs, env_stride = 1.  One scenario shared by all
   * envs: row_stride = 1, env_stride = 0.
   * rows: radar r  -> 16*r + {0 pt,1 gt,2 gr,3 wavelength,4 loss,5 latm,6 pn_watts,7 Ga,
   *                          8 D,9 threat,10 x,11 y,12 theta_m,13 theta_a,14 t_s,15 type_id}
   *       jammer j -> 16*R + 8*j + {0 gj,1 loss,2 latm,3 bj,4 x,5 y,6 power_min,7 power_max}
   *       target k -> 16*R + 8*J + 3*k + {0 x,1 y,2 rcs}          (gains/losses linear) */
  const double* data;
  int64_t row_stride;
  int64_t env_stride;
  double rd_min, rd_max;    /* clip range of the tracking penalty  (environment.py:365) */
  double rp_min, rp_max;    /* power penalty at P=1 / P=0          (environment.py:377) */
  double alb_a;             /* ln(0.62/prfa)                       (core/radar.py:70)   */
  double alb_zoff;          /* 5 log10(m)/(6.2+4.54/sqrt(m)+0.44)  (core/radar.py:73)   */
  double alb_den;           /* 1.7 + 0.12 A                        (core/radar.py:74)   */
} macjd_env_tables;

typedef struct macjd_env_io {
  /* inputs of step() -- environment.py:221 `actions` as two arrays */
  const int32_t* act_d;     /* [n_envs][J] discrete T_i                              */
  const float* act_p;       /* [n_envs][J] normalised power P_i                      */
  const float* noise;       /* [n_envs][R*K + J] injected uniforms, or NULL -> Philox */
  uint64_t seed;            /* Philox key when noise == NULL                          */
  int32_t auto_reset;       /* != 0: step_count <- 0 after a terminated step (opt-in) */
  int32_t reserved;
  /* per-episode state */
  int32_t* step_count;      /* [n_envs] in/out                   (environment.py:235) */
  /* outputs of step() */
  float* reward;            /* [n_envs]                          (environment.py:457) */
  float* r_d;               /* [n_envs]                          (environment.py:359) */
  float* r_p;               /* [n_envs]                          (environment.py:372) */
  float* r_j;               /* [n_envs]                          (environment.py:385) */
  double* reward64;         /* [n_envs] optional float64 reward                       */
  uint8_t* terminated;      /* [n_envs]                          (environment.py:460) */
  float* pd;                /* [R*K][n_envs] Pd per radar-target (environment.py:337) */
  uint8_t* detected;        /* [R*K][n_envs] u <= pd             (environment.py:341) */
  uint8_t* tracking;        /* [R][n_envs] radar state == TRACK  (core/radar.py:90)   */
  float* snr0;              /* [R*K][n_envs] SNR without jamming (environment.py:327) */
  float* snr1;              /* [R*K][n_envs] SNR with jamming    (environment.py:333) */
  float* jsr_db;            /* [R*K][n_envs] 10log10(D*Prjs/(Ga*Ps))     (extension)  */
  float* pd_net;            /* [K][n_envs] 1 - prod_r(1 - pd[r][k])      (extension)  */
  float* jam_power;         /* [J][n_envs] jammer.power in W     (environment.py:272) */
  /* outputs of step() and reset(): the static views */
  float* state;             /* [n_envs][S]                       (environment.py:479) */
  float* obs;               /* [n_envs][J][S]                    (environment.py:512) */
  uint8_t* avail;           /* [n_envs][J][A] all ones           (environment.py:539) */
} macjd_env_io;

/* environment.py:221-477  step(actions) for all envs. */
MACJD_API int macjd_env_step(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io);
/* environment.py:208-219  reset(): zero step_count, write state / obs / avail. */
MACJD_API int macjd_env_reset(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io);

/* ===================================================================== agent step
 * Replaces core/mac.py:59-198 (BasicMAC.select_actions / forward / init_hidden),
 * core/networks.py:16-180 (RNNAgent: fc1 -> GRUCell, actor MLP, MP-DQN Q-head evaluated
 * for every discrete action), utils/action_selectors.py:15-63 (masked epsilon-greedy) and
 * the time-unrolled use of the same network by core/qmix.py:217-280, as ONE fused kernel
 * per launch (csrc/agent_act.cuh).
 *
 * Weights are read from a packed, K-major copy of the RNNAgent state_dict
 * (core/networks.py: pack_agent_weights builds it; every matrix is stored transposed,
 * [in][out] with `out` contiguous, K padded with zero rows to a multiple of 32):
 */
typedef struct macjd_agent_weights {
  int32_t obs_dim;        /* O                                                         */
  int32_t obs_pad;        /* O rounded up to a multiple of 32                          */
  int32_t hidden;         /* H  = rnn_hidden_dim,   multiple of 64, <= 256             */
  int32_t actor_hidden;   /* AH = actor_hidden_dim, multiple of 64, <= 256             */
  int32_t n_actions;      /* A <= 64                                                   */
  int32_t reserved;
  const float* wa1t;      /* [obs_pad][AH]  actor.0.weight^T                           */
  const float* ba1;       /* [AH]                                                      */
  const float* wa2t;      /* [AH][AH]       actor.2.weight^T                           */
  const float* ba2;       /* [AH]                                                      */
  const float* wa3t;      /* [AH][A]        actor.4.weight^T                           */
  const float* ba3;       /* [A]                                                       */
  const float* wfc1t;     /* [obs_pad][H]   fc1.weight^T                               */
  const float* bfc1;      /* [H]                                                       */
  const float* wrzt;      /* [2H][2H] rows 0..H-1: rnn.weight_ih[0:2H]^T, rows H..2H-1:
                             rnn.weight_hh[0:2H]^T; columns: r gate then z gate        */
  const float* brz;       /* [2H]  (bias_ih + bias_hh)[0:2H]                           */
  const float* wint;      /* [H][H] rnn.weight_ih[2H:3H]^T                             */
  const float* bin;       /* [H]    rnn.bias_ih[2H:3H]                                 */
  const float* whnt;      /* [H][H] rnn.weight_hh[2H:3H]^T                             */
  const float* bhn;       /* [H]    rnn.bias_hh[2H:3H]                                 */
  const float* wqt;       /* [H][H] fc2_q_head.0.weight[:, 0:H]^T                      */
  const float* bq1;       /* [H]    fc2_q_head.0.bias                                  */
  const float* w1a;       /* [A][H] fc2_q_head.0.weight[:, H+a]  (one-hot columns)     */
  const float* w1p;       /* [H]    fc2_q_head.0.weight[:, H+A]  (parameter column)    */
  const float* w2;        /* [H]    fc2_q_head.2.weight                                */
  const float* bq2;       /* [1]    fc2_q_head.2.bias                                  */
} macjd_agent_weights;

typedef struct macjd_agent_io {
  int32_t n_rows;             /* M = batch * n_agents                                   */
  int32_t n_steps;            /* T timesteps unrolled inside this launch (acting: 1)    */
  const float* obs;           /* [T][M][O]                                              */
  float* hidden;              /* [M][H] recurrent state, updated in place; may be NULL  */
  int32_t hidden_zero_init;   /* != 0: start from zeros (mac.py:189-198 init_hidden)    */
  int32_t test_mode;          /* != 0: greedy only (action_selectors.py:59-61)          */
  int32_t tile_rows;          /* rows per CTA: 0 = auto, or 32 / 64 (tuning knob)        */
  int32_t reserved;
  float* hidden_seq;          /* [T][M][H] h_t after every step, optional               */
  float* q_all;               /* [T][M][A] Q(s, a, P_a) for every action, optional      */
  float* params_all;          /* [T][M][A] actor outputs P_a, optional                  */
  int32_t* greedy;            /* [T][M] argmax_a Q without availability mask, optional
                                 (double-DQN action of qmix.py:143)                     */
  const int32_t* sel_actions; /* [T][M] optional: actions to gather ...                 */
  float* q_sel;               /* [T][M] ... Q(s, sel_actions) (qmix.py:147)             */
  const uint8_t* avail;       /* [T][M][A] availability mask, NULL = all available      */
  const float* u_eps;         /* [T][M] injected uniforms for the epsilon test, or NULL */
  const int32_t* rand_actions;/* [T][M] injected random (available) actions, or NULL    */
  float epsilon;              /* action_selectors.py:30-32, evaluated by the caller     */
  uint32_t rng_step;          /* Philox step counter when draws are not injected        */
  uint64_t seed;
  int32_t* actions;           /* [T][M] chosen discrete action; NULL = no selection     */
  float* power;               /* [T][M] P of the chosen action (mac.py:151-164)         */
  float* q_chosen;            /* [T][M] Q of the chosen action, optional                */
} macjd_agent_io;

/* One launch: for t in 0..T-1: h <- GRU(relu(fc1 obs_t), h); P <- actor(obs_t);
 * Q_a <- Qhead(h, a, P_a) for all a; masked epsilon-greedy / argmax / gathers. */
MACJD_API int macjd_agent_forward(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io);

#ifdef __cplusplus
}
#endif
#endif /* MACJD_H_ */
