/*
 * azb.h — C ABI of azdopt-b200: the B200-native replacement for ONE hot path of
 * ariasanovsky/azdopt, the batched search step of the c21 example
 * (`NablaOptimizer::par_roll_out_episodes` over `ROTModifyParentsOnce`).
 *
 * The reference has no FFI: its boundary is Rust traits plus the public methods
 * of `NablaOptimizer` (az-discrete-opt/src/nabla/optimizer/mod.rs:29-364).  Each
 * entry point below names the reference item it replaces (file:line relative to
 * the reference repository).  The space's four closures and `n_as_tol` are Rust
 * fn pointers / closures (graph-state/src/rooted_tree/space.rs:13-19,
 * graph-state/examples/04-c21-tree.rs:96-105,136-138); they cannot cross a C ABI,
 * so they travel as data in `azb_config`.
 *
 * Conventions: every function returns an `int` status (AZB_OK = 0); nothing
 * aborts.  Buffers are caller-owned host memory unless a name ends in `_dev`.
 * A handle is bound to one CUDA device and is NOT thread-safe (like the
 * reference's `&mut self`).  Row-major everywhere.
 *
 *   A  = ACTION_DIM = (N-1)(N-2)/2 - 1     (rooted_tree/space.rs:48)
 *   S  = STATE_DIM  = 2A                    (rooted_tree/space.rs:46)
 *   W  = ceil(A/32) words of an action bit mask (bit a of word a/32)
 */
#ifndef AZB_H
#define AZB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZB_VERSION 100
#define AZB_NONE 0xFFFFFFFFu
#define AZB_MAX_VERTICES 64
#define AZB_MAX_TOL 8
#define AZB_ASYNC_AUTO 0xFFFFFFFFu  /* azb_config.async_workers: let the library pick the model SMs from the root count */
#define AZB_ASYNC_SHARED 0xFFFFFFFEu /* azb_config.async_workers: no SM is taken from the trees — every SM walks trees with 28 warps and
                                       its last warpgroup is one of 8 members of a model group (N <= 46) */

enum {
    AZB_OK = 0,
    AZB_ERR_INVALID = 1,       /* bad argument / bad config */
    AZB_ERR_CUDA = 2,          /* a CUDA call failed; see azb_last_error */
    AZB_ERR_CAPACITY = 3,      /* a per-tree arena (nodes, predictions, parents, hash, frontier) overflowed */
    AZB_ERR_NAN = 4,           /* NaN reached a comparison (the reference panics: next_action.rs:51,74,85) */
    AZB_ERR_LAMBDA = 5,        /* lambda_1 < 1.4 (the reference asserts: rooted_tree/ordered_edge.rs:79) */
    AZB_ERR_UNREACHABLE = 6,   /* inactive non-root walker (the reference's unreachable!(): tree/mod.rs:227) */
    AZB_ERR_STATE = 7          /* call order violated (e.g. step before init_trees) */
};

/* where h_theta comes from in the fused device loop */
enum {
    AZB_PRIOR_MLP = 0,      /* the built-in MLP (replaces ActionModel::write_predictions, nabla/model/dfdx.rs:69-84) */
    AZB_PRIOR_HASH = 1,     /* counter hash of (prior_seed, global root, step, action) in [0,1): parity runs */
    AZB_PRIOR_INJECTED = 2  /* rows uploaded by azb_set_priors / azb_add_actions_host */
};

/* MLP arithmetic */
enum {
    AZB_MLP_FP32 = 0,     /* fp32 CUDA-core GEMM: what cuBLAS sgemm gives the reference */
    AZB_MLP_TC = 1,       /* tcgen05 tensor cores, bf16 weights and hidden activations, fp32 accumulate */
    AZB_MLP_TC3 = 2       /* tcgen05 tensor cores at f32 accuracy: every operand as bf16 hi + lo, three products per dot
                             product (hi.hi + hi.lo + lo.hi), fp32 accumulate — within ~1e-6 of the f32 forward the
                             reference runs (nabla/model/dfdx.rs:69-84), at about three times the tensor work */
};

typedef struct azb_config {
    uint32_t struct_size;       /* sizeof(azb_config), for ABI growth */
    uint32_t n_vertices;        /* N, 5..64                                (04-c21-tree.rs:33) */
    uint32_t n_roots;           /* B on this device                        (04-c21-tree.rs:54) */
    int32_t device;             /* CUDA ordinal */
    uint64_t first_root;        /* global index of local root 0 (sharding; feeds the hash generators) */
    float c_lower;              /* C_LOWER_BOUND = 2                       (04-c21-tree.rs:58) */
    float c_upper;              /* C_UPPER_BOUND                           (04-c21-tree.rs:59-68); 0 = derive from N */
    uint32_t n_as_tol[AZB_MAX_TOL]; /* revisit budget by depth             (04-c21-tree.rs:136-138) */
    uint32_t n_as_tol_len;
    uint32_t n_as_tol_default;
    uint32_t mlp_hidden[3];     /* 512, 1024, 512                          (04-c21-tree.rs:42-44) */
    uint32_t mlp_mode;          /* AZB_MLP_* */
    uint32_t prior_mode;        /* AZB_PRIOR_* */
    uint64_t prior_seed;
    uint32_t max_steps;         /* steps between azb_init_trees calls the arenas are sized for (800: 04-c21-tree.rs:134) */
    uint32_t cap_nodes;         /* per-tree capacities; 0 = derive from max_steps */
    uint32_t cap_preds;
    uint32_t cap_parents;
    uint32_t max_episodes;      /* 0: every launch runs each tree to the end of its step (the reference's lock step).
                                   k > 0: a tree that keeps hitting terminal nodes / transpositions yields after k
                                   episodes and finishes the step in a later launch; same results, shorter launches */
    uint32_t n_groups;          /* >1: the trees of this handle advance as that many independent groups on their own
                                   CUDA streams (needs max_episodes = 0); results are identical, launches overlap */
    uint32_t async_workers;     /* > 0 (needs AZB_PRIOR_MLP + AZB_MLP_TC): azb_step(h, n >= 2) runs as ONE persistent cooperative
                                   kernel in which trees never wait for each other — the CTAs of that many SMs become
                                   tensor-core model workers that answer state vectors in 128-row tiles as they fill, all
                                   other CTAs walk trees: a CTA owns a block of the batch's trees, each warp its own while there
                                   is a warp for every tree, any free warp any runnable tree of its CTA beyond that.  Same
                                   results as the lock step (trees are independent).  0 = lock step.  AZB_ASYNC_AUTO = the
                                   measured best layout for the root count (20 model SMs up to 4096 roots — in
                                   pairs per tile below 4096 —, 36 beyond; 16 for N >= 47), or the lock step where the
                                   asynchronous kernel does not apply (no tensor-core model, fewer than 1024 roots,
                                   max_episodes or n_groups set); azb_get_config shows the choice.  Environment, read when the
                                   handle first runs this way: AZB_ASYNC_GROUP=g (worker SMs per tile), AZB_ASYNC_PAIR=1 (model
                                   CTAs as CTA pairs of one cluster: tcgen05.mma.cta_group::2, even async_workers),
                                   AZB_ASYNC_STEAL=0/1 (tree scheduling inside a CTA), AZB_ASYNC_FLUSH_NS,
                                   AZB_ASYNC_TIMEOUT_MS (watchdog floor, default 2000). */
    uint32_t reserved[5];
} azb_config;

typedef struct azb_counters {   /* workload counters; definitions in oracle/azb_oracle.h and DESIGN.md */
    uint64_t n_sel, d_sel, n_cur, n_cand, n_probe, n_ins, n_term, n_hit, n_arc, n_pred, n_cn, d_cn, n_reset, n_live,
        n_noop, n_visit;
} azb_counters;

typedef struct azb_improvement { /* one ArgminImprovement::Improved (optimizer/mod.rs:24-27) */
    uint32_t step;               /* step index since azb_init_trees (0-based) */
    uint32_t tree;               /* local root index */
    uint32_t node;               /* node index inside that tree */
    float eval;
} azb_improvement;

typedef struct azb_handle azb_handle;

/* ---- library ---- */
int azb_version(void);
const char *azb_strerror(int code);
/* last CUDA / validation message of this handle (never NULL) */
const char *azb_last_error(const azb_handle *h);

/* ---- construction: NablaOptimizer::par_new (optimizer/mod.rs:39-118) is create + set_roots + init_trees ---- */
int azb_config_default(azb_config *cfg, uint32_t n_vertices, uint32_t n_roots);
int azb_create(const azb_config *cfg, azb_handle **out);
int azb_destroy(azb_handle *h);
int azb_get_config(const azb_handle *h, azb_config *out);   /* with derived fields filled in */

/* synthetic roots of SURVEY.md §8d, the example's distribution (04-c21-tree.rs:85,108-112;
 * rooted_tree/mod.rs:14-20; modify_parent_once.rs:14-25).  Host-only helper. */
int azb_generate_roots(uint64_t seed, uint64_t first_root, uint32_t count, uint32_t n_vertices, uint32_t k_min,
                       uint32_t k_max, uint8_t *parents /*[count*N]*/, uint32_t *permitted /*[count*W]*/);
/* roots: Vec<Space::State> (optimizer/mod.rs:9,61) */
int azb_set_roots(azb_handle *h, const uint8_t *parents /*[B*N]*/, const uint32_t *permitted /*[B*W]*/);
int azb_get_roots(azb_handle *h, uint8_t *parents, uint32_t *permitted);

/* ---- the model: NablaModel (nabla/model/mod.rs:4-8), ActionModel (nabla/model/dfdx.rs:18-53) ---- */
/* parameter block, dfdx order: for each of the 4 Linear layers weight[out][in] then bias[out], f32 */
size_t azb_mlp_num_params(const azb_handle *h);
int azb_mlp_init(azb_handle *h, uint64_t seed);             /* U(-1/sqrt(fan_in), 1/sqrt(fan_in)) */
int azb_mlp_set_params(azb_handle *h, const float *params);
int azb_mlp_get_params(azb_handle *h, float *params);
/* NablaModel::write_predictions (nabla/model/dfdx.rs:69-84): host [rows*S] -> host [rows*A]; rows <= B */
int azb_model_write_predictions(azb_handle *h, const float *states, float *predictions, uint32_t rows);

/* ---- tail of par_new / par_reset_trees (optimizer/mod.rs:62-101, 347-359): root costs, root vectors, one
 *      prior evaluation, root node + root predictions, argmin over roots.  In AZB_PRIOR_INJECTED mode the
 *      priors come from the last azb_set_priors. ---- */
int azb_set_priors(azb_handle *h, const float *priors /*[B*A]*/);
int azb_init_trees(azb_handle *h);

/* ---- the hot path: NablaOptimizer::par_roll_out_episodes (optimizer/mod.rs:121-191), n_steps times, fused on
 *      the device (no host round trip per step).  The reference reports an improvement per step
 *      (04-c21-tree.rs:143-148); `improvements` receives up to `cap` of them, `*n_improved` the total. ---- */
int azb_step(azb_handle *h, uint32_t n_steps, azb_improvement *improvements, uint32_t cap, uint32_t *n_improved);
/* enqueue only (no wait, no read-back; needs max_episodes = 0): lets several handles share one GPU concurrently.
 * azb_step(h, 0, ...) or any reading call completes the work and reports errors. */
int azb_step_enqueue(azb_handle *h, uint32_t n_steps);
/* par_roll_out_episodes' per-step return value (ArgminImprovement, optimizer/mod.rs:24-27,121-191) for steps enqueued
 * with azb_step_enqueue, one step per call in order, as soon as EVERY tree has finished that step — later steps keep
 * running (with async_workers > 0 the trees run ahead of the caller, so a loop of one call per step advances at the
 * speed of the fused loop).  *improved = 1 and *out = the record (may be null) when the step lowered the best cost
 * (the reference's rule: first minimum over trees, strictly below the running best, optimizer/mod.rs:194-246).
 * azb_step(h, 0, ...) after the last step completes the batch; it logs exactly the improvements reported here.
 * AZB_ERR_STATE when no enqueued step is left to report. */
int azb_step_poll(azb_handle *h, azb_improvement *out, int *improved);
/* same as azb_step, timed with CUDA events on the library's stream; *ms = device time of the n_steps */
int azb_step_timed(azb_handle *h, uint32_t n_steps, float *ms, uint32_t *n_improved);

/* same again, with an event between the kernels of every step: *tree_ms = summed device time of the search kernel
 * launches, *mlp_ms = of the model forward launches (roofline.achieved in bench.py divides by the former) */
int azb_step_profile(azb_handle *h, uint32_t n_steps, float *tree_ms, float *mlp_ms);

/* the same step split at the NablaModel boundary, for a host-side model (host buffers, copies included):
 *   rollout_host      = first parallel region (optimizer/mod.rs:159-174): walks + write_vec, D2H of state_vecs
 *   add_actions_host  = second region + par_update_argmmim_data (optimizer/mod.rs:177-190): H2D of h_theta */
int azb_rollout_host(azb_handle *h, float *state_vecs /*[B*S]*/);
int azb_add_actions_host(azb_handle *h, const float *h_theta /*[B*A]*/, int *improved);

/* ---- results ---- */
/* ArgminData{state, cost, eval} (log.rs:1-11; optimizer/mod.rs:361-363) */
int azb_get_argmin(azb_handle *h, uint8_t *parents /*[N]*/, uint32_t *permitted /*[W]*/, double *lambda1,
                   uint32_t *mu, float *eval);
/* The state of one node: its tree's root with the node's action set replayed in ascending order — the rebuild
 * par_update_argmmim_data does for the winning node (optimizer/mod.rs:224-239).  Does not wait for enqueued steps, so
 * an azb_step_poll record (tree, node) can be turned into ArgminData while later steps still run. */
int azb_get_node_state(azb_handle *h, uint32_t tree, uint32_t node, uint8_t *parents /*[N]*/, uint32_t *permitted /*[W]*/);
/* persistent walker state: states, paths, last_positions (optimizer/mod.rs:10-14) */
int azb_get_walkers(azb_handle *h, uint8_t *parents, uint32_t *permitted, uint32_t *path, uint32_t *pos,
                    uint32_t *path_len);
/* get_trees (optimizer/mod.rs:34-36), as a canonical dump of one SearchTree (tree/mod.rs:28-32):
 * nodes 6 x u32 {c bits, c* bits, n_t, exhausted_children, lo, hi} (state_weight.rs:4-10), keys W words per node,
 * preds 3 x u32 {a_id, g bits, edge_id|AZB_NONE} (arc_weight.rs:11-16), arcs 3 x u32 {src, dst, prediction_pos}
 * in creation order (petgraph EdgeIndex order). */
int azb_tree_sizes(azb_handle *h, uint32_t tree, uint32_t *n_nodes, uint32_t *n_arcs, uint32_t *n_preds);
int azb_dump_tree(azb_handle *h, uint32_t tree, uint32_t *nodes, uint32_t *keys, uint32_t *preds, uint32_t *arcs);
int azb_get_counters(azb_handle *h, azb_counters *out);
int azb_reset_counters(azb_handle *h);
/* full != 0 (default): all 16 workload counters; 0: only n_live, n_noop, n_ins (the timed bench pass) */
int azb_set_counter_mode(azb_handle *h, int full);
/* current state vectors / priors held on the device (state_vecs, h_theta_host: optimizer/mod.rs:15-16) */
int azb_get_state_vecs(azb_handle *h, float *state_vecs /*[B*S]*/);
int azb_get_priors(azb_handle *h, float *priors /*[B*A]*/);

/* ---- stand-alone cost kernel: Space::cost + evaluate (rooted_tree/ordered_edge.rs:72-124; 04-c21-tree.rs:96-102)
 *      for M independent trees.  ms (may be NULL) = device time of the kernel alone. ---- */
int azb_eval_costs(azb_handle *h, const uint8_t *parents /*[M*N]*/, uint32_t m, double *lambda1, uint32_t *mu,
                   float *c, float *ms);

/* ---- the same cost for CONNECTED GRAPHS held as neighbourhood bit sets
 *      (ConnectedBitsetGraph<N, B32>: simple_graph/connected_bitset_graph/mod.rs).  nbr[g*n + v] = neighbours of v.
 *      lambda1 = largest eigenvalue of A + 1e-4 I (adjacency_matrix :200-216, conjecture_2_1_cost :319-337),
 *      mu = matching_number (:218-317), kinds[g][ceil(n(n-1)/32)] = action_kinds (:134-154) in the index space of
 *      AddOrDeleteEdge::action_index (bitset_graph/space/action.rs:10-19): bit colex(e) set <=> Add(e) is available,
 *      bit n(n-1)/2 + colex(e) set <=> Delete(e) is (e is an edge and not a cut edge, is_cut_edge :45-71).
 *      AZB_ERR_INVALID for n outside 2..32, loops, asymmetric or disconnected input (what try_from / to_connected
 *      reject); AZB_ERR_LAMBDA where the reference asserts lambda_1 > 1.4.  Any output pointer may be NULL. ---- */
int azb_eval_graph_costs(azb_handle *h, const uint32_t *nbr /*[M*n]*/, uint32_t m, uint32_t n, double *lambda1,
                         uint32_t *mu, uint32_t *kinds, float *ms);

/* ---- epoch boundary ("next" rows): par_update_model's observation pass (tree/mod.rs:242-264,
 *      optimizer/mod.rs:262-278) ---- */
int azb_write_observations(azb_handle *h, uint32_t n_obs_tol, float *state_vecs, float *observations,
                           float *weights);

/* ---- epoch boundary: the training step.  NablaModel::update_model (nabla/model/mod.rs:7; ActionModel:
 *      nabla/model/dfdx.rs:86-131): w_n = w / sum(w); loss = sum (forward(states) - observations)^2 * w_n; backward;
 *      Adam (04-c21-tree.rs:87-92: lr 1e-4, betas .9/.999, eps 1e-8, WeightDecay::L2(1e-6)); gradients zeroed.
 *      f32 throughout, like the reference.  Returns the loss (dfdx.rs:125,130). ---- */
int azb_adam_config(azb_handle *h, float lr, float beta1, float beta2, float eps, float l2);  /* AdamConfig */
/* update_model with host buffers [rows*S], [rows*A], [rows*A]; rows <= B */
int azb_model_update(azb_handle *h, const float *states, const float *observations, const float *action_weights,
                     uint32_t rows, float *loss);
/* test hook: loss and d loss / d params (parameter order) of the same objective, no Adam step */
int azb_model_gradients(azb_handle *h, const float *states, const float *observations, const float *action_weights,
                        uint32_t rows, float *loss, float *grads);
/* NablaOptimizer::par_update_model (optimizer/mod.rs:249-281): root vectors + observations + update_model, all on
 * the device.  With a communicator attached (below) the weight sum, the gradient and the loss are all-reduced, so
 * every rank applies the same global-batch Adam step. */
int azb_update_model(azb_handle *h, uint32_t n_obs_tol, float *loss);

/* NablaOptimizer::par_reset_trees (optimizer/mod.rs:284-360) with the example's modify_root policy
 * (04-c21-tree.rs:172-206) evaluated on the device, followed by the tail shared with par_new (= azb_init_trees):
 * a root that found nothing better moves to an equal-cost node and widens its permitted set (or is regenerated
 * once that set is at k_max); otherwise it moves to a uniformly chosen node with c <= (c + 3 c*) / 4 and re-draws
 * the permitted set.  The reference draws from an unseeded thread_rng; here every draw is a counter hash of
 * (seed, number of resets so far, GLOBAL root index), so a sharded run re-selects exactly like an unsharded one.
 * k_min / k_max = num_permitted_actions_range (0, 0 = the example's 5 ..= ACTION_DIM / 2). */
int azb_reset_trees(azb_handle *h, uint64_t seed, uint32_t k_min, uint32_t k_max);

/* ---- epoch-boundary communicator (NCCL, loaded on demand; no collective ever runs inside azb_step).  Rank 0 makes
 *      an id, the host program hands it to every rank (any out-of-band channel), each rank attaches its handle. ---- */
int azb_comm_unique_id(uint8_t *id128 /*[128]*/);
int azb_comm_init(azb_handle *h, const uint8_t *id128, int rank, int world);
int azb_comm_destroy(azb_handle *h);
/* argmin over ALL ranks' roots (the sharded form of optimizer/mod.rs:221 + argmin_data): same outputs as
 * azb_get_argmin plus the owning rank; the lowest rank wins ties */
int azb_comm_argmin(azb_handle *h, uint8_t *parents, uint32_t *permitted, double *lambda1, uint32_t *mu, float *eval,
                    int *owner_rank);

/* measurement helper: `reps` NCCL all-reduces of the gradient buffer (azb_mlp_num_params floats), timed with events on
 * the library's stream; *ms = milliseconds per all-reduce */
int azb_comm_allreduce_bench(azb_handle *h, uint32_t reps, float *ms);

/* ---- measurement helpers ---- */
int azb_kernel_launches(const azb_handle *h, uint64_t *n);   /* kernels launched by this handle so far */
int azb_device_bytes(const azb_handle *h, uint64_t *bytes);  /* HBM held by this handle */
int azb_flush_l2(azb_handle *h);                             /* overwrite a >L2 scratch buffer */
/* cascades (cascade_new_terminal / cascade_old_node, nabla/tree/empty_transitions.rs:50-127) whose ancestor wave
 * outgrew the on-chip work list and continued in the tree's HBM scratch, since azb_create.  The reference's BTreeMap
 * frontier has no limit and neither has this path; the count exists so that tests can prove they exercised it. */
int azb_debug_cascade_spills(azb_handle *h, uint32_t *n);

#ifdef __cplusplus
}
#endif
#endif
