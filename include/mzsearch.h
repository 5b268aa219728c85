/*
 * mzsearch.h — C ABI of libmzsearch.so: the B200-native batched MuZero search that drops in behind
 * muax.MuZero.act / muax.policy.Policy.
 *
 * The reference has no native interface on this path: `MuZero._plan` (muax/model.py:222-243) hands a
 * `RootFnOutput` and the `_recurrent_inference` callback (muax/model.py:265-282) to
 * `mctx.muzero_policy` / `mctx.gumbel_muzero_policy` (muax/policy.py:13-47) and XLA compiles the lot.
 * These entry points are what an FFI for that call would bind; each one cites the reference
 * interface it replaces.  Plain pointers and sizes only — no torch / C++ types cross this boundary.
 *
 * Conventions
 *   - every `*_dev` pointer is device memory on `mz_config.device`, caller-owned, borrowed for the
 *     duration of the stream-ordered call; `stream` is a cudaStream_t passed as void* (NULL = default);
 *   - a handle is NOT thread-safe; use one handle per (device, stream);
 *   - every function returns MZ_OK (0) on success; on failure it returns MZ_ERR_INVALID_ARGUMENT (the caller passed
 *     something the reference would reject with a ValueError / chex assertion) or MZ_ERR_RUNTIME (CUDA failure,
 *     unsupported configuration, call out of sequence) and a message is available from mz_last_error();
 *   - the library never falls back to a CPU path: without a usable CUDA device mz_create fails.
 */
#ifndef MZSEARCH_H_
#define MZSEARCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MZ_OK 0
#define MZ_ERR_RUNTIME 1
#define MZ_ERR_INVALID_ARGUMENT 2

#define MZ_MAX_LAYERS 8
#define MZ_MAX_ACTIONS 32

#define MZ_POLICY_MUZERO 0 /* mctx.muzero_policy        — muax/policy.py:13-30 */
#define MZ_POLICY_GUMBEL 1 /* mctx.gumbel_muzero_policy — muax/policy.py:33-47 */

#define MZ_QTRANSFORM_BY_PARENT_AND_SIBLINGS 0 /* what MuZero.act forces: muax/model.py:230-231 */
#define MZ_QTRANSFORM_COMPLETED_BY_MIX_VALUE 1 /* mctx's default for the Gumbel policy: muax/policy.py:44 */

#define MZ_PRNG_THREEFRY_LEGACY 0        /* jax_threefry_partitionable = False (JAX < 0.5.0) */
#define MZ_PRNG_THREEFRY_PARTITIONABLE 1 /* jax_threefry_partitionable = True */

#define MZ_ACT_ELU 0  /* jax.nn.elu — muax/nn.py:78-104 */
#define MZ_ACT_RELU 1

#define MZ_ENGINE_AUTO 0
#define MZ_ENGINE_STEPWISE 1 /* select / recurrent / backup kernels per simulation, trees in HBM (also the callback mode) */
#define MZ_ENGINE_FUSED 2    /* the best one-launch engine for the configuration: warp, else tree-warp, else resident */
/* 3..6 were the CTA-phased / group / lane / lane2 shared-memory engines of round 1 (retired) */
#define MZ_ENGINE_RESIDENT 7    /* one launch per act, a CTA owns T trees kept in HBM/L2 (any shapes, weights streamed) */
#define MZ_ENGINE_FUSED_WARP 8  /* warp-autonomous, trees in shared memory, compile-time shapes of the stock MLP family */
#define MZ_ENGINE_TREEWARP 9    /* warp-autonomous, trees as records in L1/L2, weights in shared memory, any shapes */

#define MZ_PRECISION_FP32 0 /* parity mode: every engine, bit-identical to the CPU checkers */
#define MZ_PRECISION_BF16 1 /* throughput mode: recurrent_fn on tcgen05 tensor cores, bf16 operands, fp32 accumulate */

#define MZ_FLAG_WANT_TREE 1u /* keep / produce the mctx.Tree view of the search (mz_get_tree); muax never reads it */

/* One hk.Sequential of hk.Linear layers with an activation between layers (none after the last):
 * muax/nn.py:63-65, 77-84, 97-104.  Offsets are in floats into the weight blob; W is [in][out]
 * row-major (haiku's layout, y = x @ W + b), b is [out]. */
typedef struct mz_stack {
  int32_t n_layers;
  int32_t in_dim[MZ_MAX_LAYERS];
  int32_t out_dim[MZ_MAX_LAYERS];
  int64_t w_off[MZ_MAX_LAYERS];
  int64_t b_off[MZ_MAX_LAYERS];
} mz_stack;

/* Static shape of one search engine instance: replaces the static arguments of the jitted
 * `MuZero._plan` (muax/model.py:222) plus the network definitions of muax/nn.py:59-115. */
typedef struct mz_config {
  int32_t batch;               /* rows (environments / trees) owned by this handle */
  int32_t num_actions;         /* A, 1..MZ_MAX_ACTIONS */
  int32_t embed_dim;           /* E */
  int32_t obs_dim;             /* flat observation size; 0 if roots are always supplied by the caller */
  int32_t support_size;        /* S; value/reward heads have 2S+1 logits (muax/model.py:48,260) */
  int32_t max_num_simulations; /* capacity: the tree workspace holds max_num_simulations + 1 nodes */
  int32_t activation;          /* MZ_ACT_* */
  int32_t repr_minmax;         /* min_max_normalize after Representation (muax/nn.py:69) */
  int32_t dyn_minmax;          /* min_max_normalize on the next state (muax/nn.py:114) */
  int32_t prng_mode;           /* MZ_PRNG_* */
  int32_t device;              /* CUDA device ordinal */
  float discount;              /* muax/model.py:47,275 */
  mz_stack repr;               /* obs -> s                     (muax/nn.py:59-70)  */
  mz_stack pred_v;             /* s -> value logits [2S+1]      (muax/nn.py:77-80)  */
  mz_stack pred_pi;            /* s -> policy logits [A]        (muax/nn.py:81-84)  */
  mz_stack dyn_ns;             /* [s, onehot(a)] -> next s      (muax/nn.py:97-100) */
  mz_stack dyn_r;              /* [s, onehot(a)] -> reward logits (muax/nn.py:101-104) */
} mz_config;

/* Per-call arguments: the keyword arguments of MuZero.act / Policy.__call__
 * (muax/model.py:82-96, muax/policy.py:17-30, 37-47). */
typedef struct mz_search_args {
  int32_t policy;         /* MZ_POLICY_* */
  int32_t qtransform;     /* MZ_QTRANSFORM_* */
  int32_t num_simulations;
  int32_t max_depth;      /* <= 0 means None */
  int32_t max_considered; /* max_num_considered_actions (Gumbel) */
  int32_t global_batch;   /* rows of the whole (multi-GPU) batch; <= 0 means == batch */
  int32_t batch_offset;   /* global index of this handle's row 0 (PRNG draws are indexed globally) */
  int32_t engine;         /* MZ_ENGINE_* */
  float temperature;
  float dirichlet_fraction;
  float dirichlet_alpha;
  float pb_c_init;
  float pb_c_base;
  float gumbel_scale;
  float value_scale;      /* qtransform_completed_by_mix_value, default 0.1 */
  float maxvisit_init;    /* qtransform_completed_by_mix_value, default 50 */
  uint32_t key0, key1;    /* the jax PRNGKey words */
  uint32_t flags;         /* MZ_FLAG_* */
  int32_t precision;      /* MZ_PRECISION_* */
  /* mctx.stochastic_muzero_policy (muax/policy.py:50-67), callback mode only (mz_begin .. mz_finish): the handle's
   * num_actions is A' = A + C (decision actions then chance outcomes); num_decision_actions = A (> 0 switches the
   * mode on).  Nodes at even depth are decision nodes (pUCT over A' with -inf priors on the chance slots), nodes at
   * odd depth are afterstates (argmax prior / (visits + 1)); root noise, invalid-action mask, visit summary and the
   * final draw see the A decision actions only.  The caller's recurrent_fn plays mctx's stochastic_recurrent_fn. */
  int32_t num_decision_actions;
} mz_search_args;

/* Device views of the search tree after a search (mctx.Tree field names).
 * [B,N], [B,N,A] and [B,N,E] row-major; valid until the next call on the handle. */
typedef struct mz_tree_view {
  int32_t batch, num_nodes, num_actions, embed_dim;
  const int32_t *node_visits, *parents, *action_from_parent, *children_index, *children_visits;
  const float *raw_values, *node_values, *children_prior_logits, *children_values, *children_rewards,
      *children_discounts, *embeddings;
  const float* root_noise;   /* [B,A] dirichlet noise / root gumbel actually used */
  const int32_t* sim_depth;  /* [B,num_simulations] selected path length per simulation */
} mz_tree_view;

typedef struct mz_handle mz_handle;

const char* mz_last_error(void);

/* Fills `args` with the reference defaults (muax/model.py:82-96, muax/policy.py:17-47). */
void mz_default_args(mz_search_args* args);

/* Replaces `MuZero.__init__` + the first (compiling) `_plan` call: allocates the tree workspace. */
int mz_create(mz_handle** out, const mz_config* cfg);
int mz_destroy(mz_handle* h);

/* Replaces passing `params` into `_plan` (muax/model.py:162): copies the fp32 weight blob (layout given
 * by the mz_stack offsets in mz_config).  `on_device` != 0: `blob` is device memory. */
int mz_set_weights(mz_handle* h, const float* blob, size_t n_floats, int on_device, void* stream);

/* Replaces `MuZero._plan` (muax/model.py:222-243): root inference + policy + search + action.
 * Root: give `obs_dev` [B,obs_dim] (the library runs Representation + Prediction, model.py:251-263); or
 * obs_dev == NULL and root_emb_dev [B,E] from the caller's own Representation (e.g. a conv torso) — the
 * library then runs Prediction on it; or additionally root_logits_dev [B,A] + root_value_dev [B] = a
 * complete RootFnOutput computed by the caller.
 * invalid_dev: [B,A] uint8, 1 = invalid, or NULL.  noise_dev: [B,A] injected Dirichlet noise (MuZero) or
 * root Gumbel (Gumbel policy), or NULL to draw on device.
 * Outputs: action [B] int32, action_weights [B,A], root_value [B] (raw network value, model.py:243). */
int mz_search(mz_handle* h, const float* obs_dev, const float* root_logits_dev, const float* root_value_dev,
              const float* root_emb_dev, const uint8_t* invalid_dev, const float* noise_dev,
              const mz_search_args* args, int32_t* action_out_dev, float* action_weights_out_dev,
              float* root_value_out_dev, void* stream);

/* Multi-GPU env sharding (the reference has no counterpart — its act is single-process): the next
 * mz_search calls store action / action_weights / root_value not only at the given output pointers but also at
 * (pointer + byte_deltas[i]) for i < n — the same slots of the peer GPUs' gather buffers, mapped into this process
 * (CUDA IPC / torch symmetric memory) — so the per-act all-gather is done by the search kernel's own NVLink stores.
 * Only the warp engine honours it (mz_search fails otherwise); n = 0 switches it off.  The caller orders the peers'
 * reads after the stores (a cross-rank barrier after the kernel). */
int mz_set_peer_outputs(mz_handle* h, int32_t n, const int64_t* byte_deltas);

/* Completion flags for that exchange, so that no barrier kernel sits between two acts: flags_dev = W int32 words in
 * this rank's (symmetric) gather buffer, one per source rank.  With flags set, the last CTA of the next mz_search's
 * kernel — after every output store, local and peer, has been fenced system-wide — stores `step` at flags_dev[rank]
 * and at the same word of every peer (flags_dev + rank, shifted by the byte_deltas of mz_set_peer_outputs).  Steps
 * must increase.  flags_dev = NULL switches it off.  mz_peer_wait enqueues a one-CTA kernel on `stream` that returns
 * when flags_dev[q] >= step for all q < world (every rank's rows of act `step` have arrived here); it gives up after
 * a few seconds of polling instead of hanging the device. */
int mz_set_peer_flags(mz_handle* h, int32_t* flags_dev, int32_t rank, int32_t step);
int mz_peer_wait(mz_handle* h, const int32_t* flags_dev, int32_t world, int32_t step, void* stream);

/* Same call with HOST buffers (what `MuZero.act` sees: numpy in, numpy out — model.py:160-174): stages
 * through mapped pinned memory (read / written in place by the kernels; copy-engine H2D for large observation
 * batches), searches and synchronises the stream. */
int mz_search_host(mz_handle* h, const float* obs_host, const uint8_t* invalid_host, const float* noise_host,
                   const mz_search_args* args, int32_t* action_out_host, float* action_weights_out_host,
                   float* root_value_out_host, void* stream);

/* Replaces `MuZero._recurrent_inference` (muax/model.py:265-282) for a batch on its own: Dynamic -> Prediction -> the
 * two support transforms.  action_dev [B] int32, embedding_dev [B,E] -> reward [B], value [B], prior logits [B,A],
 * next embedding [B,E] (discount is the handle's constant).  precision: MZ_PRECISION_FP32 (the kernel every fp32 engine
 * is checked against) or MZ_PRECISION_BF16 (the tcgen05 kernel of the throughput mode). */
int mz_recurrent(mz_handle* h, const int32_t* action_dev, const float* embedding_dev, int32_t precision,
                 float* reward_out_dev, float* value_out_dev, float* prior_logits_out_dev, float* next_embedding_out_dev,
                 void* stream);

/* Callback mode — for networks the declarative stacks cannot express.  The caller plays mctx's
 * `recurrent_fn` (muax/model.py:265-282) between mz_select and mz_expand_backup:
 *   mz_begin(root, ...); for sim in range(num_simulations): mz_select -> recurrent_fn -> mz_expand_backup;
 *   mz_finish(...). */
int mz_begin(mz_handle* h, const float* root_logits_dev, const float* root_value_dev, const float* root_emb_dev,
             const uint8_t* invalid_dev, const float* noise_dev, const mz_search_args* args, void* stream);
/* parent_emb_out_dev [B,E] = embeddings[b, parent[b]]; action_out_dev [B]. */
int mz_select(mz_handle* h, int32_t sim, int32_t* action_out_dev, float* parent_emb_out_dev, void* stream);
int mz_expand_backup(mz_handle* h, int32_t sim, const float* reward_dev, const float* discount_dev,
                     const float* prior_logits_dev, const float* value_dev, const float* next_emb_dev, void* stream);
int mz_finish(mz_handle* h, int32_t* action_out_dev, float* action_weights_out_dev, void* stream);

/* The tree of the last search; fails unless that search ran with MZ_FLAG_WANT_TREE (muax never reads
 * PolicyOutput.search_tree, so the default act does not pay for the mctx SoA view). */
int mz_get_tree(mz_handle* h, mz_tree_view* view);

/* Number of kernels this library launched on behalf of the handle since creation. */
int mz_launch_count(mz_handle* h, int64_t* count);
/* Device time (ms, CUDA events on `stream`) of the search kernels of the last mz_search* call. */
int mz_last_kernel_ms(mz_handle* h, float* ms);

/* Evaluates include/mz_math.h device functions elementwise (bit-parity tests against the host build).
 * kind: 0 expf, 1 logf, 2 expm1f, 3 inv_scaling, 4 gumbel-from-bits (x reinterpreted as uint32),
 * 5 the kernels' batched fast division: x holds n (a, b) pairs, y[i] = a/b, or a marker NaN (0x7fc00001) where the
 * kernels would fall back to the IEEE slow path. */
int mz_math_probe(int32_t kind, const float* x_dev, float* y_dev, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MZSEARCH_H_ */
