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
 *     immediately (stream-ordered); the calling thread's current CUDA device is
 *     the same on return as on entry; entry points may be called from several
 *     threads on disjoint buffers (process-wide state is limited to lock-free /
 *     mutex-protected caches of idempotent per-device driver settings)
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

#define MACJD_ABI_VERSION 2

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
  /* optional: everything of a step that depends on the scenario alone, evaluated once by
   * macjd_env_prepare() into a caller-owned buffer of macjd_env_derived_bytes() bytes (same
   * per-env / shared convention as `data`).  NULL: the step kernel derives it from `data` on
   * every step (round 1's kernel); results are the same. */
  const double* derived;
} macjd_env_tables;

/* macjd_env_io.flags.  MACJD_ENV_FOLLOWS_AGENT: the caller launches this step directly behind the kernel
 * that produces its actions (same stream, normally macjd_agent_forward).  The step kernel is scheduled
 * early in any case (programmatic dependent launch); with the flag it uses the time it waits for the
 * agent kernel to copy the scenario columns into shared memory.  Results are identical; launched on its
 * own the copy would only add latency, so leave the flag unset then. */
#define MACJD_ENV_FOLLOWS_AGENT 1

typedef struct macjd_env_io {
  /* inputs of step() -- environment.py:221 `actions` as two arrays */
  const int32_t* act_d;     /* [n_envs][J] discrete T_i                              */
  const float* act_p;       /* [n_envs][J] normalised power P_i                      */
  const float* noise;       /* [n_envs][R*K + J] injected uniforms, or NULL -> Philox */
  uint64_t seed;            /* Philox key when noise == NULL                          */
  int32_t auto_reset;       /* != 0: step_count <- 0 after a terminated step (opt-in) */
  int32_t flags;            /* MACJD_ENV_FOLLOWS_AGENT or 0                             */
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
  int32_t env_begin;        /* step only envs [env_begin, env_begin + env_count): every pointer above still */
  int32_t env_count;        /* names the whole batch (0, 0 = all; groups of one batch on separate streams)  */
} macjd_env_io;

/* Scenario-only terms of the step (environment.py:316-349 echo power and no-jamming Pd per radar-target pair,
 * core/jammer.py:73-98 link denominators, environment.py:365 clipped tracking penalties, environment.py:479-510
 * state rows), which the reference recomputes every step although the scenario never changes within a run:
 * computed once here.  `tab->derived` is ignored; point it at `derived` afterwards.  Stream-ordered. */
MACJD_API size_t macjd_env_derived_bytes(const macjd_env_tables* tab);
MACJD_API int macjd_env_prepare(const macjd_ctx* ctx, const macjd_env_tables* tab, void* derived);

/* environment.py:221-477  step(actions) for all envs. */
MACJD_API int macjd_env_step(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io);
/* environment.py:208-219  reset(): zero step_count, write state / obs / avail. */
MACJD_API int macjd_env_reset(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io);

/* Host-buffer form of step(): what the reference's Python loop hands over and gets back
 * (environment.py:221 `actions` list in; environment.py:462-477 reward, terminated and the next
 * observations out).  `io` names DEVICE staging buffers exactly as for macjd_env_step (act_d,
 * act_p, reward, terminated, obs ... must be set); the call copies the host actions in on
 * ctx->stream, launches the step, copies the requested outputs back and returns after the
 * stream has drained.  Host pointers may be pageable or page-locked; page-locked buffers are
 * handed to the kernel in place when that is faster than a copy-engine transfer (outputs up to
 * 8 MB, inputs up to 1 MB; csrc/macjd_api.cu: direct_host_limit), so the io staging buffers
 * may go unused.  NULL outputs are skipped. */
/* flags of the host-buffer structs.  MACJD_HOST_PINNED: the caller vouches that every host pointer in
 * the struct is page-locked memory of the unified address space (cudaHostAlloc / torch pin_memory), so
 * the library uses them as device addresses without asking the driver about each one on every call
 * (about 1 us per pointer).  Without the flag each pointer is queried; pageable memory is copied. */
#define MACJD_HOST_PINNED 1u

typedef struct macjd_env_host {
  const int32_t* act_d;     /* host [n_envs][J]                                       */
  const float* act_p;       /* host [n_envs][J]                                       */
  float* reward;            /* host [n_envs], optional                                */
  uint8_t* terminated;      /* host [n_envs], optional                                */
  float* obs;               /* host [n_envs][J][S], optional                          */
  float* state;             /* host [n_envs][S], optional                             */
  uint32_t flags;           /* MACJD_HOST_PINNED: see below                           */
  uint32_t reserved;
} macjd_env_host;
MACJD_API int macjd_env_step_host(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io,
                                  const macjd_env_host* host);

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
  int32_t tc_format;      /* layout of tc_chunks: 0 = weight chunks only (path 2), 1 = weight
                             chunks followed by the per-layer constant block (needed by path 3) */
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
  const float* tc_chunks; /* optional: the dense layers again, packed for the tcgen05 path
                             (csrc/agent_act_tc2.cuh): 128 x 32 weight chunks in consumption order,
                             UMMA K-major layout, TF32 hi part then lo part (16 KB per chunk);
                             with tc_format 1 the 13 856-byte constant block (biases, actor
                             output layer, Q-head vectors; csrc/agent_tc_common.cuh: TcConst) follows
                             the last chunk so that the CTA-pair kernel fetches it with one bulk
                             copy, and for A > 8 the per-action tables follow that block:
                             [H][A8] actor.4.weight^T, [H][A8] fc2_q_head.0.weight[:, H + a],
                             [A8] actor.4.bias (A8 = A rounded up to 8, zero padded);
                             NULL = FP32 SIMT kernel only                              */
  const float* wiht;      /* optional [H][3H] rnn.weight_ih^T (gates r | z | n side by side) and ...           */
  const float* whht;      /* ... [H][3H] rnn.weight_hh^T: let macjd_agent_unroll form a GRU side with ONE
                             product instead of two (NULL: it uses wrzt / wint / whnt)                  */
  const float* rec_chunks;/* optional, H = 128 or 256: rnn.weight_hh packed for the tcgen05 recurrence launch of
                             macjd_agent_unroll (csrc/gru_rec_tc2.cuh): for every 128-unit block b of the hidden
                             state, for every gate g (r, z, n), the H / 32 chunks of rows [g H + 128 b, + 128) x
                             32 k in the tc_chunks chunk format (UMMA K-major, TF32 hi part then lo part, 32 KB
                             per chunk); NULL: the recurrence runs as one GEMM + one gate launch per timestep  */
  const float* bgx;       /* [3H] (brz | bin): the biases of the GRU's input side as one vector, added by the
                             input-product GEMM when the recurrence launch is used (needed with rec_chunks)   */
} macjd_agent_weights;

typedef struct macjd_agent_io {
  int32_t n_rows;             /* M = batch * n_agents                                   */
  int32_t n_steps;            /* T timesteps unrolled inside this launch (acting: 1)    */
  const float* obs;           /* [T][M][O]                                              */
  float* hidden;              /* [M][H] recurrent state, updated in place; may be NULL  */
  int32_t hidden_zero_init;   /* != 0: start from zeros (mac.py:189-198 init_hidden)    */
  int32_t test_mode;          /* != 0: greedy only (action_selectors.py:59-61)          */
  int32_t tile_rows;          /* rows per CTA: 0 = auto, or 8/16/32/64 (tuning knob)     */
  int32_t path;               /* 0 = auto (tensor cores when tc_chunks is given and the dims
                                 allow), 1 = FP32 SIMT, 3 = tcgen05 3xTF32 with CTA pairs
                                 (cta_group::2, 128 rows per pair; MACJD_ERR_UNSUPPORTED if the
                                 dims do not fit).  2 named round 1's single-CTA tensor-core
                                 kernel, since removed: MACJD_ERR_UNSUPPORTED */
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
  int32_t part;               /* 0 = the whole step (default); the CTA-pair kernel also runs the two
                                 halves of a step separately, for time-unrolled use where only
                                 the recurrence is sequential (core/qmix.py:217-280):
                                 1 = recurrence only: h <- GRU(relu(fc1 obs_t), h) for T steps,
                                     writes hidden / hidden_seq, nothing else;
                                 2 = heads only (n_steps = 1): `hidden` is the post-GRU state of
                                     every row (read, not updated); actor, Q-head, selection.
                                 3 = input pre-pass (n_steps = 1, rows = all T x M): the three GRU
                                     input products W_ir xf, W_iz xf, W_in xf (xf = relu(fc1 obs),
                                     no biases) -> gate_x;
                                 4 = recurrence on gate_x: per step only W_h* h and the gates
                                     (obs is not read); writes hidden / hidden_seq.  1 and 3+4
                                     differ by FP32 rounding (x and h products are summed in the
                                     epilogue instead of in the accumulator).
                                 Other kernels return MACJD_ERR_UNSUPPORTED for part != 0.   */
  int32_t rng_row_offset;     /* added to the row index in the Philox counters: a launch over rows
                                 [k, k + n_rows) of a larger batch (pointers offset by the caller) draws
                                 what the whole-batch launch draws for those rows                    */
  float* gate_x;              /* [T][M][3][H] part 3 output / part 4 input, else unused  */
  const float* epsilon_dev;   /* optional DEVICE values that override `epsilon` (n_steps floats, one per step) and
                                 `rng_step` (one scalar; step t uses it + t): kernel                       */
  const uint32_t* rng_step_dev; /* parameters are frozen when a launch is replayed from a CUDA graph, these
                                 are not (BatchedEpisodeRunner replays a whole episode as one graph) */
  int32_t* actions_mirror;    /* [T][M] optional second destinations of actions / power (same values):   */
  float* power_mirror;        /* the host-step call points them at the caller's page-locked buffers while
                                 the primary copies stay in HBM for the env kernel                       */
  const float* hidden_in;     /* optional [M][H]: the initial recurrent state is read from here instead of
                                 `hidden`, which is then only written (or NULL).  A rollout that records
                                 h_t per step (hidden_seq) chains the steps through those records and
                                 saves the second 4 MB store per step (runners/episode_runner.py)     */
  int32_t obs_group;          /* 0 / 1: `obs` holds one row per agent row ([T][M][O]).  G > 1: rows [k G, (k + 1) G)
                                 share observation row k and `obs` is [T][M / G][O] -- with G = n_agents, the
                                 env's global state once per env instead of once per jammer (in this environment
                                 every agent observes the state, environment.py:512-522).  M must be a multiple of G */
  int32_t reserved2;
} macjd_agent_io;

/* One launch: for t in 0..T-1: h <- GRU(relu(fc1 obs_t), h); P <- actor(obs_t);
 * Q_a <- Qhead(h, a, P_a) for all a; masked epsilon-greedy / argmax / gathers. */
MACJD_API int macjd_agent_forward(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io);
/* The time-unrolled forward of core/qmix.py:217-280 (the learner's eval / target unrolls: T steps from `hidden`
 * or zeros, h_t / Q for all actions / arg-max / gathers out, no action selection) as batched layers: every layer
 * but the recurrence h_t -> h_t+1 runs as ONE dense product over all T x M rows, all on the tensor cores (3xTF32)
 * where they fill its tiles.  The recurrence itself is ONE launch when w->rec_chunks (and w->bgx, w->wiht) is given
 * and H is 128 or 256 (csrc/gru_rec_tc2.cuh: CTA pairs keep their 128 rows of h in shared memory for all T steps
 * and stream W_hh per step; the environment variable MACJD_REC_KERNEL=0 disables it), else one product and one gate
 * launch per timestep.  For network widths the fused CTA-pair
 * kernel does not take (macjd_agent_pair_supported == 0, e.g. rnn_hidden_dim 256); same outputs as
 * macjd_agent_forward with n_steps = T up to FP32 summation order.  `io` as for macjd_agent_forward (actions /
 * power / obs_group must be unset); workspace: macjd_agent_unroll_workspace_floats() floats, caller-owned. */
MACJD_API size_t macjd_agent_unroll_workspace_floats(const macjd_agent_weights* w, int32_t n_rows, int32_t n_steps);
MACJD_API int macjd_agent_unroll(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io, float* workspace,
                                 size_t workspace_floats);

/* 1 if the CTA-pair tensor-core kernel (io->path 3 / auto, io->part 1 and 2) can run these weights. */
MACJD_API int macjd_agent_pair_supported(const macjd_agent_weights* w);
/* k-extent of one packed weight chunk in macjd_agent_weights.tc_chunks as this build of the library
 * expects it (core/networks.py packs accordingly); 0 = built without the tensor-core kernel. */
MACJD_API int macjd_agent_tc_chunk_k(void);

/* One timestep of the reference's rollout loop on the device (runners/episode_runner.py:119-165:
 * mac.select_actions(obs, avail) then env.step(actions)): the fused agent step (aio, n_steps = 1, part 0,
 * actions / power set) followed by the env step on the actions it chose (eio->act_d / act_p are ignored: the
 * step reads aio->actions / aio->power).  When the CTA-pair tensor-core kernel takes the agent dims, a CTA's 64
 * rows are whole envs (64 % n_jammers == 0) and `tab->derived` is set, this is ONE launch: the CTA that
 * selected an env's actions also evaluates that env's step (actions handed over in shared memory) and copies
 * its next state / obs / avail.  Otherwise the two kernels run back to back.  Results are identical to
 * macjd_agent_forward followed by macjd_env_step.  Requires aio->n_rows == n_envs * n_jammers (env-major). */
MACJD_API int macjd_rollout_step(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                                 const macjd_env_tables* tab, const macjd_env_io* eio);
/* aio->n_steps = T timesteps of that loop (runners/episode_runner.py:49-119).  The buffers are the time-major
 * rollout trajectory: aio's obs / avail / actions / power / hidden_seq ... are [T][M]... as documented in
 * macjd_agent_io, and the env outputs of step t go t steps behind eio's addresses (reward, r_d, r_p, r_j,
 * terminated: [T][n]; state [T][n][S], obs [T][n][J][S], avail [T][n][J][A]); because the agent of step t + 1
 * reads what the env of step t wrote, eio->obs must be aio->obs + M * O (the next slot of the same trajectory)
 * and eio->avail likewise (MACJD_ERR_INVALID_ARG otherwise).  aio->epsilon_dev, when given, holds T values (one
 * per step); the Philox step counter of step t is aio->rng_step + t.  The info-only env outputs (pd, detected,
 * ...) keep the last step's values.  Fused (macjd_rollout_fused_supported) this is ONE launch: each CTA pair
 * keeps its rows' recurrent state in shared memory across the steps and runs its own envs' steps; otherwise the
 * steps are issued one by one.  Results are identical to T calls of macjd_rollout_step.  Injected draws
 * (u_eps, rand_actions, noise) are not supported here: MACJD_ERR_UNSUPPORTED. */
MACJD_API int macjd_rollout_steps(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                                  const macjd_env_tables* tab, const macjd_env_io* eio);
/* 1 if macjd_rollout_step runs these weights / tables as one launch (it can, and the env work a CTA would take
 * on is small next to its agent step: at most 16 KB of next-step views per CTA). */
MACJD_API int macjd_rollout_fused_supported(const macjd_agent_weights* w, const macjd_env_tables* tab);

/* Host-buffer form of BasicMAC.select_actions (core/mac.py:59-187: numpy observations and
 * availability masks in, chosen discrete actions and their power levels out).  `io` is a
 * single-step (n_steps = 1) macjd_agent_io whose obs / avail / actions / power name DEVICE
 * staging buffers; the call copies host obs (and avail) in on ctx->stream, launches the
 * fused step, copies actions / power (and q_chosen when both sides give it) back and returns
 * after the stream has drained (small page-locked buffers are read / written in place by the
 * kernel, see macjd_env_step_host).  The recurrent state stays on the device (io->hidden). */
typedef struct macjd_act_host {
  const float* obs;         /* host [M][O]; with io->obs_group = G > 1: [M / G][O] (one row per env: the
                               global state instead of the reference's per-jammer copies -- half the
                               bytes over PCIe at two jammers; ask macjd_env_host for `state`, not `obs`) */
  const uint8_t* avail;     /* host [M][A], optional (NULL: io->avail is used as is)  */
  int32_t* actions;         /* host [M]                                               */
  float* power;             /* host [M]                                               */
  float* q_chosen;          /* host [M], optional                                     */
  uint32_t flags;           /* MACJD_HOST_PINNED                                      */
  uint32_t reserved;
} macjd_act_host;
MACJD_API int macjd_agent_act_host(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io,
                                   const macjd_act_host* host);

/* One iteration of the reference's rollout loop for a host-side caller
 * (runners/episode_runner.py:119-165: mac.select_actions(obs, avail) followed by env.step(actions)) as ONE
 * call with ONE stream drain: host observations / masks in, the fused agent step, the fused env step on
 * the actions just chosen (they never leave the device: the env kernel reads aio->actions / aio->power,
 * eio->act_d / act_p and ehost->act_d / act_p are ignored and may be NULL), then the chosen actions
 * (ahost->actions / power) and the env outputs (ehost->reward / terminated / obs / state) back to the
 * host.  The env kernel is launched as a programmatic dependent of the agent kernel.  Requirements and
 * buffer rules are those of macjd_agent_act_host and macjd_env_step_host, plus
 * aio->n_rows == n_envs * n_jammers (rows ordered env-major, as BasicMAC flattens them).
 * Results are bit-identical to the two calls made one after the other. */
MACJD_API int macjd_rollout_step_host(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                                      const macjd_act_host* ahost, const macjd_env_tables* tab,
                                      const macjd_env_io* eio, const macjd_env_host* ehost);

/* ===================================================================== replay ring
 * Replaces the data movement of utils/replay_buffer.py:78-214 (store_episode / sample) and
 * runners/episode_runner.py:187-277 (EpisodeBatch staging): index-driven episode copies
 * between rollout buffers, ring slots and training batches, all in HBM.
 */
#define MACJD_MAX_COPY_KEYS 12
typedef struct macjd_copy_desc {
  const void* src;
  void* dst;
  int64_t src_ep_stride;   /* bytes between consecutive episodes on the source side      */
  int64_t src_t_stride;    /* bytes between consecutive timesteps on the source side     */
  int64_t dst_ep_stride;
  int64_t dst_t_stride;
  int32_t n_t;             /* timesteps to copy (T or T+1)                               */
  int32_t inner_bytes;     /* contiguous bytes per timestep (on the SOURCE side when `convert` is set) */
  int32_t vec_bytes;       /* filled in by the library                                   */
  int32_t convert;         /* 0: plain bytes.  1: source float32 -> destination bfloat16 (round to nearest even),
                              2: source bfloat16 -> destination float32: the optional BF16 storage of
                              `hidden_state` in the ring (72 % of an episode, utils/replay_buffer.py:66) */
} macjd_copy_desc;

/* For b in [0, n_eps): slot = idx ? idx[b] : b.  index_on_src != 0: dst episode b <- src
 * episode slot (sample);  index_on_src == 0: dst episode slot <- src episode b (store).
 * descs_host is read on the host at call time. */
MACJD_API int macjd_replay_copy(const macjd_ctx* ctx, const macjd_copy_desc* descs_host, int32_t n_keys,
                                const int32_t* idx, int32_t n_eps, int32_t index_on_src);

/* ===================================================================== learner
 * Replaces core/qmix.py:76-215 (QMixLearner.train) below the time-unrolled agent passes
 * (which are macjd_agent_forward with n_steps = T): QMixer forward / backward
 * (core/networks.py:250-316), the Q-head on stored hidden states with its backward
 * (core/networks.py:131-180, core/qmix.py:161-184), the double-DQN TD target and masked
 * loss (core/qmix.py:155,191-194), gradient-norm clipping and Adam (core/qmix.py:197-200).
 * Rows are time-major: r = t * batch + b.
 */
typedef struct macjd_mixer_dims {
  int32_t n_rows;        /* R = batch * (T - 1)   */
  int32_t state_dim;     /* S                     */
  int32_t n_agents;      /* N                     */
  int32_t embed_dim;     /* E = mixing_embed_dim  */
  int32_t hyper_hidden;  /* HH = hyper_hidden_dim */
  int32_t reserved;
} macjd_mixer_dims;

/* QMixer tensors in PyTorch layout ([out][in] row-major); the same struct addresses the
 * weights and the gradient buffers. */
typedef struct macjd_mixer_params {
  float* ln_w;  float* ln_b;     /* state_norm.{weight,bias}      [S]               */
  float* w1a_w; float* w1a_b;    /* hyper_w_1.0      [HH][S], [HH]                   */
  float* w1b_w; float* w1b_b;    /* hyper_w_1.2      [N*E][HH], [N*E]                */
  float* wfa_w; float* wfa_b;    /* hyper_w_final.0  [HH][S], [HH]                   */
  float* wfb_w; float* wfb_b;    /* hyper_w_final.2  [E][HH], [E]                    */
  float* b1_w;  float* b1_b;     /* hyper_b_1        [E][S], [E]                     */
  float* va_w;  float* va_b;     /* V.0              [E][S], [E]                     */
  float* vb_w;  float* vb_b;     /* V.2              [1][E], [1]                     */
} macjd_mixer_params;

MACJD_API size_t macjd_mixer_workspace_floats(const macjd_mixer_dims* dims);
/* q [R][N], states [R][S] -> q_tot [R]; leaves the intermediates in `workspace`. */
MACJD_API int macjd_mixer_forward(const macjd_ctx* ctx, const macjd_mixer_dims* dims, const macjd_mixer_params* w,
                                  const float* q, const float* states, float* q_tot, float* workspace,
                                  size_t workspace_floats);
/* dq_tot [R] -> gradients of every QMixer tensor (overwritten) and dq [R][N] (may be NULL).
 * Needs the workspace left by macjd_mixer_forward on the same inputs. */
MACJD_API int macjd_mixer_backward(const macjd_ctx* ctx, const macjd_mixer_dims* dims, const macjd_mixer_params* w,
                                   const float* q, const float* dq_tot, float* workspace, size_t workspace_floats,
                                   const macjd_mixer_params* grads, float* dq);

typedef struct macjd_qhead_dims {
  int32_t n_rows;     /* batch * (T - 1) * n_agents */
  int32_t hidden;     /* H                          */
  int32_t n_actions;  /* A                          */
  int32_t reserved;
} macjd_qhead_dims;

MACJD_API size_t macjd_qhead_scratch_floats(const macjd_qhead_dims* dims);
/* q[r] = Qhead(hidden[r], a_d[r], a_c[r]) with the packed weights of macjd_agent_weights
 * (wqt, bq1, w1a, w1p, w2, bq2); hid [R][H] keeps the ReLU activations for the backward. */
MACJD_API int macjd_qhead_forward(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                                  const float* hidden, const int32_t* a_d, const float* a_c, float* q, float* hid);
/* The same with caller-owned scratch (any size; macjd_qhead_scratch_floats() is enough): room for the tensor-core GEMM to
 * pre-split the weight matrix once per call instead of once per CTA (tall batches).  Same results. */
MACJD_API int macjd_qhead_forward_ws(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                                     const float* hidden, const int32_t* a_d, const float* a_c, float* q, float* hid,
                                     float* scratch, size_t scratch_floats);
/* dq [R] -> gradients in PyTorch layout: g_w1 [H][H+A+1], g_b1 [H], g_w2 [H], g_b2 [1].
 * `hid` (from the forward) is consumed. */
MACJD_API int macjd_qhead_backward(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                                   const float* hidden, const int32_t* a_d, const float* a_c, float* hid,
                                   const float* dq, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                   float* scratch, size_t scratch_floats);

/* After the optimiser step (core/qmix.py:197-200; macjd_clip_adam updates the nn.Parameters in place) the packed copy
 * of the one trained agent layer, fc2_q_head (core/qmix.py:178), is refreshed by ONE launch: w1 [H][H + A + 1], b1 [H],
 * w2 [H], b2 [1] (PyTorch layout) -> w's fields wqt, bq1, w1a, w1p, w2, bq2 (written through their pointers) and,
 * optionally (H = 128, pointers may be NULL): tc_chunks = the last H / tc_kc chunks of w->tc_chunks (q.0[:, :H];
 * [chunk][hi | lo][128 tc_kc]), tc_q_c = the constant block's [128][4] (bq1, w1p, w2, -) rows, tc_w1a = its
 * [128][tc_w1a_stride] one-hot-column table.  Bit-identical to re-packing everything on the host. */
MACJD_API int macjd_qhead_repack(const macjd_ctx* ctx, const macjd_agent_weights* w, const float* w1, const float* b1,
                                 const float* w2, const float* b2, float* tc_chunks, int32_t tc_kc, float* tc_q_c,
                                 float* tc_w1a, int32_t tc_w1a_stride);

/* out[i] = q_all[i][idx[i]] for i < n: the target network's Q at the eval network's greedy
 * action (core/qmix.py:147 torch.gather). */
MACJD_API int macjd_gather_q(const macjd_ctx* ctx, int32_t n, int32_t n_actions, const float* q_all,
                             const int32_t* idx, float* out);

MACJD_API size_t macjd_td_scratch_floats(int32_t n_rows);
/* targets = reward + gamma (1 - terminated) tq_tot;  td = (q_tot - targets) mask;
 * dq_tot = 2 td mask (NOT divided by sum(mask): macjd_clip_adam applies 1/sums[1]);
 * sums[0..3] = { sum td^2, sum mask, sum q_tot, sum targets }. */
MACJD_API int macjd_td_loss(const macjd_ctx* ctx, int32_t n_rows, const float* q_tot, const float* tq_tot,
                            const float* reward, const uint8_t* terminated, const uint8_t* filled, float gamma,
                            float* dq_tot, float* targets, float* sums, float* scratch, size_t scratch_floats);

#define MACJD_MAX_OPT_TENSORS 32
typedef struct macjd_opt_tensors {
  int32_t count;
  int32_t reserved;
  float* param[MACJD_MAX_OPT_TENSORS];    /* parameter tensors, updated in place           */
  int64_t numel[MACJD_MAX_OPT_TENSORS];   /* their sizes; gradients / moments are the flat
                                             concatenation in the same order               */
} macjd_opt_tensors;

MACJD_API size_t macjd_opt_scratch_floats(void);
/* norm = ||grad|| / sums[1];  coef = min(1, max_norm / (norm + 1e-6))  (clip_grad_norm_);
 * Adam step `step` (1-based) on every tensor with g = grad * coef / sums[1].
 * scal[0] = norm (pre-clip), scal[1] = coef / sums[1], scal[2] = loss = sums[0] / sums[1]. */
MACJD_API int macjd_clip_adam(const macjd_ctx* ctx, const macjd_opt_tensors* tensors, const float* grad, float* m,
                              float* v, const float* sums, float max_norm, float lr, float beta1, float beta2,
                              float eps, int64_t step, float* scal, float* scratch, size_t scratch_floats);

/* The same step with its two bias corrections read from DEVICE memory at run time -- bias_corr[0] = 1 - beta1^step,
 * bias_corr[1] = sqrt(1 - beta2^step) -- so that ONE captured launch sequence (a CUDA graph of the whole train step,
 * QMixLearner.train_sampled) serves every step; the caller refreshes the two floats before each replay. */
MACJD_API int macjd_clip_adam_dev(const macjd_ctx* ctx, const macjd_opt_tensors* tensors, const float* grad, float* m,
                                  float* v, const float* sums, float max_norm, float lr, float beta1, float beta2,
                                  float eps, const float* bias_corr, float* scal, float* scratch, size_t scratch_floats);

/* A tensor-core kernel waits on its pipeline barriers with a bound; if a wait ever runs out (a peer CTA
 * that never arrives), the launch finishes with invalid outputs instead of trapping, and every LATER entry
 * point returns MACJD_ERR_CUDA ("tcgen05 pipeline wait timed out ...", macjd_last_cuda_error) until
 * macjd_clear_pipeline_fault() is called. */
MACJD_API int macjd_clear_pipeline_fault(void);

/* The dense-layer primitive every learner entry point above is built from (nn.Linear of core/networks.py:54-79,
 * 215-248 and its two backward products):  C[m][n] (ldc) = act(sum_k A(m,k) B(k,n) + bias[n]) (+ C if accumulate),
 *   A(m,k) = ta ? A[k lda + m] : A[m lda + k],   B(k,n) = tb ? B[n ldb + k] : B[k ldb + n]  (tb = 1: a PyTorch weight)
 *   act: 0 none, 1 ReLU, 3 sigmoid.  splitk_ws (optional, >= splits x M x N floats): lets a product with few output
 * tiles and a long contraction (weight gradients: K = batch rows) be split over K, reduced in a fixed order; for
 * K < 2048 and M >= 1024 the same workspace (>= ceil(N / 128) x ceil(K / 16) x 4096 floats) lets the tensor-core kernel
 * split and lay out B once per call instead of once per CTA (bit-identical results).
 * Runs on the tcgen05 tensor cores (3xTF32 operand split, FP32-level accuracy; csrc/tc_gemm.cuh) when the problem
 * fills its 128 x 128 tiles, else on the FP32 SIMT kernel (csrc/sgemm.cuh). */
MACJD_API int macjd_gemm(const macjd_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t ta,
                         const float* B, int32_t ldb, int32_t tb, float* C, int32_t ldc, const float* bias, int32_t act,
                         int32_t accumulate, float* splitk_ws, size_t splitk_ws_floats);

/* ===================================================================== tensor-core self-test
 * D[M][N] = A[M][K] B[N][K]^T on the tcgen05 TF32 pipe with the 3xTF32 operand split
 * (csrc/tc05.cuh).  M in {64, 128}, N multiple of 16 <= 256, K multiple of 8.  Exercises the
 * shared-memory descriptors, TMEM allocation, tcgen05.mma / commit / ld used by the agent
 * kernel's tensor-core path. */
MACJD_API int macjd_tc_gemm_selftest(const macjd_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A,
                                     const float* B, float* D);

#ifdef __cplusplus
}
#endif
#endif /* MACJD_H_ */
