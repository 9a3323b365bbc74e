/* hzb200 — C ABI of the B200-native HanabiZero self-play hot path.
 *
 * One shared library (hanabizero_b200/csrc/libhzb200.so, sm_100a) replaces the two native
 * libraries the reference binds on this path:
 *   (1) the Cython extension `core.ctree.cytree` over core/ctree/{cnode,cminimax}.{h,cpp}
 *       (/root/reference/core/ctree/cytree.pyx:17-101, ctree.pxd:9-77), and
 *   (2) `libpyhanabi.so`, the cffi-loaded C API over envs/hanabi/hanabi_lib
 *       (/root/reference/envs/hanabi/pyhanabi.h:24-195), as driven by HanabiEnv.reset/step
 *       (/root/reference/envs/hanabi/rl_env.py:148-267, 292-442).
 *
 * Conventions: every function returns 0 on success and a negative hz_status on failure
 * (hz_last_error() then holds a message, thread-local).  All `dev` pointers are device memory on
 * the handle's device; `stream` is a cudaStream_t passed as void*.  No call synchronises the
 * stream or allocates unless stated ("sync").  Batches are structure-of-arrays over trees/games;
 * tree i / game i of a batch is the reference's roots[i] / envs[i].
 */
#ifndef HZB200_H
#define HZB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hz_trees hz_trees;
typedef struct hz_envs hz_envs;

typedef enum {
  HZ_OK = 0,
  HZ_ERR_ARG = -1,      /* bad argument (NULL handle, size mismatch, x out of capacity ...) */
  HZ_ERR_CUDA = -2,     /* a CUDA runtime call or launch failed */
  HZ_ERR_STATE = -3,    /* call order violated (backprop without traverse, step before reset ...) */
  HZ_ERR_ILLEGAL = -4   /* an illegal Hanabi move was submitted (reference: REQUIRE -> abort,
                           hanabi_state.cc:222) */
} hz_status;

const char* hz_last_error(void);
int hz_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t hz_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Tree batch.  Replaces cytree.Roots + the per-tree node pools (CRoots, cnode.h:43-60;
 * Roots.__cinit__, cytree.pyx:42-45) and the search paths held by ResultsWrapper/CSearchResults
 * (cnode.h:62-73).  Capacity: `max_sims` expansions per tree beyond the root.
 * ------------------------------------------------------------------------------------------ */
int hz_trees_create(hz_trees** out, int device, int num_trees, int num_actions, int max_sims); /* sync */
int hz_trees_destroy(hz_trees* t);                                                             /* sync */
int hz_trees_num(const hz_trees* t);
int hz_trees_actions(const hz_trees* t);
int hz_trees_capacity(const hz_trees* t);

/* CRoots::prepare / prepare_no_noise (cnode.cpp:247-259; cytree.pyx:47-51): masked softmax
 * expansion of every root with hidden index (0, i), then the Dirichlet-noise mix.
 * noises == NULL selects prepare_no_noise.  All inputs dev: noises/logits float[N][A],
 * rewards float[N], masks int32[N][A] (0 = illegal).  Also discards any previous search. */
int hz_trees_prepare(hz_trees* t, void* stream, float exploration_fraction, const float* noises,
                     const float* rewards, const float* logits, const int32_t* masks);

/* cmulti_traverse (cnode.cpp:407-441; cytree.multi_traverse, cytree.pyx:97-101) with the tie
 * rule rand()==0.  minmax: dev float[N][2] = {minimum, maximum} per tree — the storage of
 * CMinMaxStatsList (cminimax.h:13-36), owned by the caller like the reference's separate
 * MinMaxStatsList object; value_delta_max is CMinMaxStats::value_delta_max.
 * Outputs (dev, any may be NULL): out_ix/out_iy int32[N] = the leaf's PARENT hidden-state
 * index, out_action int32[N] = last action, out_action64 int64[N] (same, for torch .long()).
 * If pool != NULL, also gathers the parent hidden state rows (mcts.py:31-35 done on device):
 *   out_hidden[i, :] = pool[(ix*N + iy) * row_bytes ...], row_bytes % 16 == 0. */
int hz_trees_traverse(hz_trees* t, void* stream, int pb_c_base, float pb_c_init, float discount,
                      const float* minmax, float value_delta_max, int32_t* out_ix, int32_t* out_iy,
                      int32_t* out_action, int64_t* out_action64, const void* pool,
                      void* out_hidden, int row_bytes);

/* cmulti_back_propagate (cnode.cpp:337-344; cytree.multi_back_propagate, cytree.pyx:87-94):
 * expand every leaf found by the last traverse with hidden index (hidden_state_index_x, i) and
 * all-legal softmax priors, back-propagate `values` along the path, then refresh the per-tree
 * min/max over all expanded non-root nodes (cback_propagate + update_tree_q, cnode.cpp:296-335).
 * rewards/values float[N], logits float[N][A] (dev).  sanitize_nan != 0 zeroes NaN logits first
 * (what core/mcts.py:48-49 does on the host).  minmax is written. */
int hz_trees_backprop(hz_trees* t, void* stream, int hidden_state_index_x, float discount,
                      const float* rewards, const float* values, const float* logits,
                      int sanitize_nan, float* minmax);

/* Fused step for the device-resident search loop: backprop of simulation k immediately followed
 * by the traverse (+ gather) of simulation k+1 in ONE launch.  Same results as calling
 * hz_trees_backprop then hz_trees_traverse. */
int hz_trees_backprop_traverse(hz_trees* t, void* stream, int hidden_state_index_x, float discount,
                               const float* rewards, const float* values, const float* logits,
                               int sanitize_nan, float* minmax, float value_delta_max,
                               int pb_c_base, float pb_c_init, int32_t* out_ix, int32_t* out_iy,
                               int32_t* out_action, int64_t* out_action64, const void* pool,
                               void* out_hidden, int row_bytes);

/* One launch per simulation for the device-resident loop, consuming the network's RAW outputs:
 *   x == 0 : traverse only (first simulation after prepare);
 *   x >= 1 : decode value/reward of simulation x from their categorical logits
 *            (core/config.py:210-232), expand + back-propagate (as hz_trees_backprop), store the new
 *            node's hidden state into pool[x], then (do_traverse != 0) traverse for simulation
 *            x+1 and write the next network input batch: the parent's hidden state followed by
 *            one-hot(last action) (the concat of dynamics(), config/hanabi_control/model.py:199-203). */
typedef struct hz_search_io {
  /* network outputs of simulation x, row-major, `elem_bytes` per element (2 = half, 4 = float) */
  const void* value_logits;   int64_t ld_value;    /* [N][support_width] */
  const void* reward_logits;  int64_t ld_reward;   /* [N][support_width] */
  const void* policy_logits;  int64_t ld_policy;   /* [N][A] */
  const void* next_state;     int64_t ld_state;    /* [N][state_cols], or NULL when the network wrote the new
                                                    * hidden state straight into pool[x] (no copy in the kernel) */
  const float* support;       /* dev float[support_width] */
  int32_t support_width;      /* <= 256 */
  float support_delta;
  int32_t elem_bytes;
  int32_t sanitize_nan;       /* zero NaN policy logits (core/mcts.py:48-49) */
  /* hidden-state pool and the batch handed to the network */
  void* pool;                 /* [capacity+1][N][state_cols], elem_bytes each */
  int32_t state_cols;         /* state_cols * elem_bytes must be a multiple of 16 */
  void* out_batch;            /* [N][ld_batch]: state_cols of hidden state, then onehot_cols of one-hot */
  int64_t ld_batch;           /* elements; ld_batch * elem_bytes multiple of 16 */
  int32_t onehot_cols;        /* 0..32; columns >= num_actions stay zero */
  int32_t* out_ix;            /* optional int32[N]: parent hidden-state index x */
  int32_t* out_action;        /* optional int32[N]: last action */
  /* search constants */
  float* minmax;              /* dev float[N][2], read and written */
  float value_delta_max;
  float discount;
  int32_t pb_c_base;
  float pb_c_init;
  /* != 0: launch with programmatic stream serialization — the kernel fetches its tree state while the preceding
   * kernel of the stream drains and synchronises on it before reading the network outputs.  The caller guarantees
   * that the preceding kernel is not the previous tree step (there is at least one network kernel in between). */
  int32_t programmatic_launch;
  /* > 0: stage at most this many expanded nodes' records in shared memory (and none of q).  0 = as much of the tree as
   * fits, the fastest choice for a search that runs alone.  With several searches in flight on different streams a
   * small limit (4) is better: the launch then needs ~6 KB of shared memory per CTA and shares SMs with the library
   * GEMM CTAs of the other searches instead of waiting for SMs without one. */
  int32_t stage_limit;
} hz_search_io;
int hz_trees_search_step(hz_trees* t, void* stream, int x, int do_traverse, const hz_search_io* io);

/* After replaying a CUDA graph that contains the calls above, the handle's host-side progress
 * (number of completed back-propagations since prepare) must be restored by hand: a replay runs
 * the kernels but not the host code.  expansions in [0, capacity]. */
int hz_trees_set_progress(hz_trees* t, int expansions);

/* Tie rule of cselect_child (cnode.cpp:346-374).  The reference draws rand() % ties over the children whose score is
 * within 1e-6 of the best (after srand(gettimeofday) per traverse, cnode.cpp:409-411).  mode 0 (default, the parity
 * contract with the rand() == 0 build of the reference): the first element of that list.  mode 1: a uniform draw
 * from a counter-based generator keyed by (seed, tree_offset + tree index, simulations completed, depth) — the same
 * tie LIST as the reference, a reproducible draw instead of the clock-seeded libc stream.  tree_offset is the global
 * index of this batch's tree 0, so a root batch sharded over ranks draws exactly what the unsharded batch draws.
 * Takes effect at the next hz_trees_prepare (which stores the rule in device memory, so a captured CUDA graph of the
 * simulation loop follows the rule of the search it is replayed for); pass a fresh seed per search. */
int hz_trees_set_tie_break(hz_trees* t, int mode, uint64_t seed, int tree_offset);

/* CRoots::get_distributions / get_values (cnode.cpp:276-292): out_visits int32[N][A] (dev),
 * out_values float[N] (dev). */
int hz_trees_root_stats(hz_trees* t, void* stream, int32_t* out_visits, float* out_values);

/* CRoots::get_trajectories (cnode.cpp:266-274): out int32[N][max_len] (dev), -1 padded. */
int hz_trees_trajectories(hz_trees* t, void* stream, int32_t* out, int max_len);

/* Test/inspection export: per tree and expansion index x in [1, cap]: reward, value_sum (dev
 * float[N][cap]) and visit_count (dev int32[N][cap]); entries past the last expansion are 0.
 * out_path_len int32[N] = length of the last search path (nodes incl. root). NULLs allowed. */
int hz_trees_export(hz_trees* t, void* stream, int cap, float* out_reward, float* out_value_sum,
                    int32_t* out_visits, float* out_root_priors, int32_t* out_path_len);

/* mcts.py:31-35 as a standalone op: out[i,:] = pool[(ix[i]*num + iy[i]) * row_bytes ...]. */
int hz_gather_hidden(void* stream, const void* pool, const int32_t* ix, const int32_t* iy,
                     void* out, int num, int row_bytes);

/* ------------------------------------------------------------------------------------------
 * Hanabi game batch.  Replaces, per game i, one HanabiGame + HanabiState + ObservationEncoder
 * (pyhanabi.h: NewGame, NewState, StateApplyMove, StateDealRandomCard, NewObservation,
 * EncodeObservation, EncodeOwnHandObservation, ObsGetLegalMove, StateScore,
 * StateEndOfGameStatus) as sequenced by HanabiEnv.reset/step.
 * preset: 0 = "Hanabi-Full", 1 = "Hanabi-Small" (rl_env.py:110-131).  seeds: HOST int32[N],
 * game i seeds its own std::mt19937 with seeds[i] (hanabi_game.cc:43-51); the stream persists
 * across resets like the reference's game object.
 * ------------------------------------------------------------------------------------------ */
int hz_envs_create(hz_envs** out, int device, int num_games, int preset, const int32_t* seeds); /* sync */
int hz_envs_destroy(hz_envs* e);                                                                /* sync */
/* dims: [enc_len, own_len, players, actions, colors, ranks, hand_size, max_info, max_life,
 *        local_dim, global_dim, state_dump_len] */
int hz_envs_dims(const hz_envs* e, int32_t* out12);

/* HanabiEnv.reset (rl_env.py:148-267) for the games with reset_mask[i] != 0 (dev uint8[N];
 * NULL = all games). Observations are produced by hz_envs_observe. */
int hz_envs_reset(hz_envs* e, void* stream, const uint8_t* reset_mask);

/* HanabiEnv.step (rl_env.py:292-442) for every game with active[i] != 0 (dev uint8[N]; NULL =
 * all): apply move uid actions[i] (dev int32[N]), deal while chance, then reward = score delta,
 * done, score.  out_reward/out_score int32[N], out_done uint8[N] (dev).  An illegal action
 * leaves that game untouched and raises the sticky error flag read by hz_envs_check. */
int hz_envs_step(hz_envs* e, void* stream, const int32_t* actions, const uint8_t* active,
                 int32_t* out_reward, uint8_t* out_done, int32_t* out_score);

/* The observation tuple HanabiEnv returns for the current player (rl_env.py:254-263, 426-434):
 * out_global float[N][global_dim] = own-hand ‖ canonical encoding ‖ turn one-hot,
 * out_local float[N][local_dim], out_legal float[N][actions] (dev; any may be NULL).
 * ld_global/ld_local = row stride in elements (>= dim; lets the caller write straight into a
 * frame-stack buffer). */
int hz_envs_observe(hz_envs* e, void* stream, float* out_global, int64_t ld_global,
                    float* out_local, int64_t ld_local, float* out_legal);

/* step + auto-reset of finished games + observe in one launch (the self-play inner loop). */
int hz_envs_step_observe(hz_envs* e, void* stream, const int32_t* actions, const uint8_t* active,
                         int auto_reset, int32_t* out_reward, uint8_t* out_done, int32_t* out_score,
                         float* out_global, int64_t ld_global, float* out_local, int64_t ld_local,
                         float* out_legal);

/* The same two calls with byte-valued outputs: every element is 0 or 1 as uint8 — the value type
 * the reference encoder itself emits (std::vector<int> of 0/1, canonical_encoders.cc:441-486,
 * serialised by pyhanabi.cc:855-895) at a quarter of the float32 traffic.  Row strides in bytes. */
int hz_envs_observe_u8(hz_envs* e, void* stream, uint8_t* out_global, int64_t ld_global, uint8_t* out_local,
                       int64_t ld_local, uint8_t* out_legal);
int hz_envs_step_observe_u8(hz_envs* e, void* stream, const int32_t* actions, const uint8_t* active,
                            int auto_reset, int32_t* out_reward, uint8_t* out_done, int32_t* out_score,
                            uint8_t* out_global, int64_t ld_global, uint8_t* out_local, int64_t ld_local,
                            uint8_t* out_legal);

/* The host-facing form of the same launch: ONE bit-packed row of uint32 words per game (dev, row stride ld_words >=
 * W + 4 with W = ceil(global_dim / 32): 25 for Hanabi-Full, 7 for Hanabi-Small):
 *   words [0, W)   the global observation, bit j = (row[j >> 5] >> (j & 31)) & 1 — the encoder emits 0/1 values
 *                  (canonical_encoders.cc:441-486), so 785 of them are 100 bytes, not 785 or 3140; the local
 *                  observation is bits [own_len, global_dim) of the same string (rl_env.py:261-262)
 *   word  W        legal-move mask, bit a = move uid a is legal (rl_env.py:263)
 *   words W+1..3   reward (int32), done (0/1), score of the step (rl_env.py:436-442)
 * actions == NULL observes without stepping (reward 0).  A host caller copies N * (W + 4) * 4 bytes per step.
 * out_meta (optional, dev uint32[N][4], 16-byte aligned): the four trailing words {legal mask, reward, done, score} go
 * there instead (rows then need only W words) — a host policy that looks at the legal mask reads 16 contiguous bytes
 * per game instead of one word out of every 116-byte row. */
int hz_envs_step_observe_bits(hz_envs* e, void* stream, const int32_t* actions, const uint8_t* active, int auto_reset,
                              uint32_t* out_bits, int64_t ld_words, uint32_t* out_meta);
/* Host-facing step for callers that keep actions and results in page-locked HOST memory (a vectorised stand-in for
 * the per-step pyhanabi.h calls of HanabiEnv.step, rl_env.py:292-442): h_actions (int32[N], NULL = observe only),
 * h_bits / h_meta as in hz_envs_step_observe_bits but HOST pointers.  No staging copies: pinned memory is
 * device-addressable, the kernel reads and writes it directly, so a step is one launch + one event record on
 * `stream`.  hz_envs_host_wait blocks the calling thread until the last hz_envs_host_step of this handle has landed in
 * host memory.  Neither call holds any lock: one host thread per game batch may drive its own handle. */
int hz_envs_host_step(hz_envs* e, void* stream, const int32_t* h_actions, int auto_reset, uint32_t* h_bits,
                      int64_t ld_words, uint32_t* h_meta);
int hz_envs_host_wait(hz_envs* e);
/* Fused random policy for device-resident random play (rollouts, env throughput runs): after this call every
 * observing launch of the handle (hz_envs_observe*, hz_envs_step_observe*, hz_envs_host_step) also writes, for every
 * game it observes, a uniformly random legal move of the observed position into next_actions (dev int32[N], caller
 * owned; 0 for a finished game) — pass the same buffer as `actions` of the next step and a whole random-play step is
 * ONE launch.  Draw d of game i is the k-th legal move with k from the generator of hz_host_random_legal keyed
 * (seed, i, step = d): the host twin reproduces any pick.  The per-game draw counters live on the device (CUDA-graph
 * replays keep advancing them) and are reset to 0 by this call; next_actions == NULL switches the policy off. */
int hz_envs_set_random_policy(hz_envs* e, void* stream, int32_t* next_actions, uint64_t seed);
/* HOST helper for callers that drive the games from the CPU through the packed rows (no device work): a uniformly
 * random legal move per game from word `legal_word` (= W) of each host row, counter-based (seed, game, step). */
int hz_host_random_legal(const uint32_t* rows, int64_t ld_words, int legal_word, int num_games, int num_actions,
                         uint64_t seed, uint32_t step, int32_t* out_actions);

/* sync: returns HZ_ERR_ILLEGAL if any game saw an illegal action since the last check
 * (out_game = first offending game index), clearing the flag. */
int hz_envs_check(hz_envs* e, void* stream, int32_t* out_game);

/* Full hidden state of every game for parity tests, int32[N][state_dump_len] (dev), layout of
 * oracle/hanabi_oracle.c:ohanabi_dump. */
int hz_envs_dump(hz_envs* e, void* stream, int32_t* out);

/* ------------------------------------------------------------------------------------------
 * Glue around the PyTorch network (the GEMMs stay cuBLAS): keeps value/reward decoding and the
 * layer epilogues on the device in one launch each.
 * ------------------------------------------------------------------------------------------ */
/* inverse_scalar_transform (/root/reference/core/config.py:210-232): per row softmax over `width`
 * support bins (logits dev, elem_bytes 4 = float / 2 = half, row stride ld elements), expectation
 * against support[width] (dev float), / delta, inverse of h(x)=sign(x)(sqrt(|x|+1)-1)+0.001x,
 * NaN -> 0.  out: dev float[rows]. */
int hz_support_decode(void* stream, const void* logits, int elem_bytes, const float* support,
                      float* out, int rows, int width, int64_t ld, float delta);
/* ------------------------------------------------------------------------------------------
 * "Next" rows: the hops between a search and the next env step.
 * ------------------------------------------------------------------------------------------ */
/* Root exploration noise (np.random.dirichlet([alpha] * A) per root, /root/reference/core/selfplay_worker.py:279,
 * reanalyze_worker.py:343-344) from counter-based streams: Philox4x32-10 keyed by `seed`, counter (block, action,
 * root_offset + root, step); Gamma(alpha) by Marsaglia-Tsang, normalised in ascending order, rounded to float32.
 * out float[N][A] (dev).  legal_mask (optional, float[N][A]): entries with mask 0 are zeroed AFTER normalisation, as
 * the reanalyze caller does.  hz_host_dirichlet_noise is the host twin (host pointers, no device work): the same
 * arithmetic built from IEEE float64 operations only, bit-identical output — a search can be replayed on the CPU
 * (e.g. through the reference's own cytree) with exactly the noise the device drew. */
int hz_dirichlet_noise(void* stream, float* out, int num_roots, int num_actions, double alpha, uint64_t seed,
                       uint32_t step, uint32_t root_offset, const float* legal_mask);
int hz_host_dirichlet_noise(float* out, int num_roots, int num_actions, double alpha, uint64_t seed, uint32_t step,
                            uint32_t root_offset, const float* legal_mask);

/* select_action (/root/reference/core/utils.py:280-295): visits int32[N][A] (dev; counts of illegal
 * actions are zeroed in place like the reference does), legal float[N][A], temperature float[N] or
 * NULL (= 1), uniforms double[N] in [0,1) or NULL (deterministic: first arg-max).  Sampling follows
 * numpy.random.choice (cdf = cumsum(p)/sum, searchsorted side='right') for the given uniform.
 * out_action int32[N], out_entropy float[N] or NULL (base-2 entropy of the visit distribution). */
int hz_select_action(void* stream, int32_t* visits, const float* legal, const float* temperature,
                     const double* uniforms, int num, int num_actions, int32_t* out_action, float* out_entropy);
/* Frame stack on the device (core/game.py:169-174): stack float[N][depth][dim] shifts left by one
 * frame and appends obs[i] (row stride ld_obs); where done[i] != 0 every slot is filled with obs[i]
 * (first frame of the next episode replicated, selfplay_worker.py:137). */
int hz_stack_push(void* stream, float* stack, const float* obs, int64_t ld_obs, const uint8_t* done, int num,
                  int stack_depth, int dim);

/* The same frame stack as a RING of 0/1 bytes: ring uint8[N][stack_depth][padded_dim] (dev; padded_dim = frame length
 * rounded up to 16).  The env kernel writes every new observation straight into the slot holding the oldest frame
 * (hz_envs_step_observe_u8 with that slot as its output row), so a move costs no memmove at all.
 * hz_ring_gather: the network input in frame order, oldest first: out[i][j * frame_stride + d] =
 *   ring[i][(head + j) % stack_depth][d] as half (elem_bytes 2) or float (4); frame_stride >= padded_dim, both
 *   multiples of 16 (the first layer's weight columns are laid out with the same stride).
 * hz_ring_refill: games with done[i] != 0 (NULL = all) get slot src_slot copied into every other slot — the first
 *   observation of a new episode fills the whole stack (core/selfplay_worker.py:137). */
int hz_ring_gather(void* stream, const uint8_t* ring, int head, void* out, int64_t ld_out, int frame_stride, int num,
                   int stack_depth, int padded_dim, int elem_bytes);
int hz_ring_refill(void* stream, uint8_t* ring, int src_slot, const uint8_t* done, int num, int stack_depth, int padded_dim);

/* Trajectory record of N self-play games in HBM: GameHistory.init /
 * store_search_stats / append / game_over (/root/reference/core/game.py:73-93,143-148,176-204) for a
 * whole batch.  The caller owns the buffers (device memory, zero-initialised) and passes this view.
 * `banks` >= 2 banks per game (B below), used round-robin: an episode closed by done[i] waits in its bank for
 * hz_traj_pack while the game's next episodes are recorded into the following banks.  Observations are 0/1
 * valued and kept as bytes. */
typedef struct hz_traj_view {
  uint8_t* obs;        /* [N][B][stack + max_len][obs_dim]: `stack` copies of the first frame, then one per move */
  uint8_t* legal;      /* [N][B][max_len + 1][actions]: legal mask before move t (row 0 from begin) */
  int32_t* action;     /* [N][B][max_len] */
  int32_t* reward;     /* [N][B][max_len] raw env rewards (the turn-reward reshaping of
                          selfplay_worker.py:29-39 is applied when the episode is handed over) */
  int32_t* visits;     /* [N][B][max_len][actions] root child visit counts */
  float* root_value;   /* [N][B][max_len] */
  int32_t* len;        /* [N][B] moves recorded in each bank */
  uint8_t* bank;       /* [N] bank currently written */
  uint8_t* finished;   /* [N][B] 1 = complete episode awaiting hz_traj_pack */
  int32_t* overflow;   /* [1] sticky: 1 + first game that ran out of room (episode longer than max_len,
                          or every bank full because the host did not pack in time) */
  int32_t num, obs_dim, actions, stack, max_len, banks;
} hz_traj_view;
/* GameHistory.init for the games with mask[i] != 0 (dev uint8[N], NULL = all): the first observation
 * (dev float[N][ld_obs]) fills the `stack` leading frames, the first legal mask (dev float[N][actions])
 * row 0. */
int hz_traj_begin(void* stream, const hz_traj_view* v, const float* obs, int64_t ld_obs, const float* legal,
                  const uint8_t* mask);
/* One self-play move of every game with active[i] != 0 (NULL = all): store_search_stats(visits, root_value)
 * + append(action, obs, reward, legal); done[i] != 0 closes the episode (game_over) and moves on to the next bank —
 * call hz_traj_begin with mask = done afterwards, once the env has been reset.  All pointers dev. */
int hz_traj_append(void* stream, const hz_traj_view* v, const int32_t* actions, const float* obs, int64_t ld_obs,
                   const float* legal, const int32_t* reward, const int32_t* visits, const float* root_values,
                   const uint8_t* done, const uint8_t* active);
/* Hand over finished episodes: episode e = (ep_game[e], ep_bank[e]) (dev int32) with T_e moves is copied to
 * rows [step_off[e], step_off[e] + T_e) of out_action/out_reward/out_root/out_visits, to rows
 * [step_off[e] + e*stack, +stack + T_e) of out_obs and [step_off[e] + e, +T_e + 1) of out_legal (dev
 * int64 step_off = exclusive prefix sum of the T_e); the bank is freed. */
int hz_traj_pack(void* stream, const hz_traj_view* v, int num_episodes, const int32_t* ep_game, const int32_t* ep_bank,
                 const int64_t* step_off, uint8_t* out_obs, uint8_t* out_legal, int32_t* out_action,
                 int32_t* out_reward, int32_t* out_visits, float* out_root);
/* Policy targets from visit counts (store_search_stats, game.py:194-197; the reanalyze caller,
 * /root/reference/core/reanalyze_worker.py:352-367): out[i][a] = visits[i][a] / sum_a visits[i]
 * as a correctly rounded double quotient (Python's int / int), rows with mask[i] == 0 (dev uint8[N], NULL = keep
 * all) all zero.  out64 double[N][A] and/or out32 float[N][A] (dev). */
int hz_visit_policy(void* stream, const int32_t* visits, const uint8_t* mask, int num, int num_actions, double* out64,
                    float* out32);

/* A fixed chain of nn.Linear-shaped GEMMs executed with cuBLASLt, one launch each:
 *   D[m][n] = act( A[m][k] . W[n][k]^T + bias[n] + C[m][n] ),  all row-major, strided batches allowed.
 * Pointers are captured at creation (static buffers: the chain is CUDA-graph friendly); hz_gemm_plan_set_operand
 * re-points single operands. */
typedef struct hz_gemm_step {
  const void* a;    int64_t lda; int64_t stride_a;     /* activations */
  const void* w;    int64_t ldw; int64_t stride_w;     /* weights in nn.Linear layout [n][k] */
  const void* bias; int64_t stride_bias;               /* [n] or NULL */
  const void* c;    int64_t ldc; int64_t stride_c;     /* residual or NULL */
  void* d;          int64_t ldd; int64_t stride_d;     /* output */
  int32_t m, n, k, batch, relu;
} hz_gemm_step;
typedef struct hz_gemm_plan hz_gemm_plan;
int hz_gemm_plan_create(hz_gemm_plan** out, int device, int elem_bytes, const hz_gemm_step* steps, int n_steps); /* sync */
int hz_gemm_plan_destroy(hz_gemm_plan* p);                                                                     /* sync */
int hz_gemm_plan_steps(const hz_gemm_plan* p);
/* Re-point one operand of one step (which: 0 = A, 1 = C, 2 = D) at another buffer of the same shape and leading
 * dimension, 256-byte aligned.  Takes effect for the following hz_gemm_plan_run calls (and is what a CUDA graph
 * captures): the search loop makes the dynamics network write each new hidden state straight into pool[x]. */
int hz_gemm_plan_set_operand(hz_gemm_plan* p, int step, int which, void* ptr);
/* Size every step's library kernel for `sm_count` SMs instead of the whole device (CUBLASLT_MATMUL_DESC_SM_COUNT_TARGET;
 * 0 = whole device, the default): the heuristic picks fewer, fatter CTAs, so the GEMMs of several plans running on
 * different streams — independent searches in flight — share the GPU instead of queueing behind one another's
 * grid-filling launches.  Re-selects the kernels at once; call before capturing the plan in a CUDA graph. */
int hz_gemm_plan_set_sm_target(hz_gemm_plan* p, int sm_count);
/* cuBLASLt launches issued through hz_gemm_plan_run in this process (library GEMMs, counted apart from
 * hz_launch_count, which counts this library's own kernels) */
int64_t hz_gemm_launch_count(void);
int hz_gemm_plan_run(hz_gemm_plan* p, void* stream, int first, int count);

/* Row-block resident executor of the Hanabi-Full recurrent_inference chain (fp16): the same function as the
 * seven-step hz_gemm_plan that hanabizero_b200/plan.py builds for MuZeroNetFull
 * (/root/reference/core/model.py:74-84, config/hanabi_control/model.py:199-216,301-318, eval mode, BatchNorm folded),
 * as ONE launch in which each CTA keeps 128 rows of the batch in shared memory / TMEM through all layers and only the
 * weights stream in (TMA + tcgen05, hanabizero_b200/csrc/hz_rowchain.cu).  A 4096-row batch occupies 32 SMs, which is
 * what several searches in flight want.  All pointers dev, fp16, 16-byte aligned, captured at creation. */
typedef struct hz_rowchain_weights {
  const void* w1;  int64_t ld_w1;     /* [512][ld_w1]: fc1 (+bn1) over [state | one-hot action] */
  const void* w1a_t;                  /* [32][512]: the action columns of w1, transposed */
  const void* b1;
  const void* w2;  const void* b2;    /* [512][512] */
  const void* w3;  const void* b3;    /* [512][512]; the residual is the input state */
  const void* wh1; const void* bh1;   /* [768][512]: first layers of policy | value | reward heads */
  const void* wb2; const void* bb2;   /* [3][256][256]: second layers, same order */
  const void* wa2; const void* ba2;   /* [256][256]: policy residual block, second layer */
  const void* wb3; const void* bb3;   /* [3][logit_cols][256]: output layers value | reward | policy */
  int32_t state_cols, head_cols, onehot_cols, logit_cols;   /* 512, 256, 32, <= 256 (multiple of 16) */
} hz_rowchain_weights;
typedef struct hz_rowchain hz_rowchain;
/* x0 [rows][ld_x0] = state | one-hot(action) (what hz_trees_search_step hands over), state [rows][512] receives the
 * next hidden state, out_logits [3][rows][logit_cols] = value | reward | policy logits. */
int hz_rowchain_create(hz_rowchain** out, int device, const hz_rowchain_weights* w, int rows, const void* x0,
                       int64_t ld_x0, void* state, void* out_logits);                                   /* sync */
int hz_rowchain_destroy(hz_rowchain* e);
/* Re-point the next-state output (the search loop passes pool[x]); takes effect for the following runs. */
int hz_rowchain_set_state(hz_rowchain* e, void* state);
int hz_rowchain_run(hz_rowchain* e, void* stream);
int hz_rowchain_grid(const hz_rowchain* e);                     /* CTAs per launch */
/* Debug: per-CTA globaltimer stamps of the first row block, uint64[grid][10 phases][4] =
 * {MMA may start, MMAs issued, accumulators complete, tile written back}. */
int hz_rowchain_set_trace(hz_rowchain* e, int enable);
int hz_rowchain_read_trace(hz_rowchain* e, uint64_t* host_out, int64_t count);                          /* sync */

#ifdef __cplusplus
}
#endif
#endif /* HZB200_H */
