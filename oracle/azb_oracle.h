/*
 * oracle/azb_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C interface of the CPU restatement ("oracle") of the reference's batched
 * search step for the c21 example.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product library (azdopt_b200/lib/libazb.so, include/azb.h) never links,
 * loads or calls it.
 *
 * Parity status: the reference's own tests pin only the indexing tables and two
 * analytic cost values (see tests/test_oracle_golden.py).  The search-DAG logic
 * (nabla::tree, nabla::optimizer) has no reference test, and the reference
 * cannot be built here (no Rust toolchain): for that part the header says
 * "parity unpinned" — it is defended by invariants and an independent Python
 * restatement in tests/.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef AZB_ORACLE_H
#define AZB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NONE 0xFFFFFFFFu

/* lambda_1 methods */
#define ORC_LAMBDA_DENSE 0        /* Householder tridiagonalisation + implicit QL: what the reference's faer call does (ordered_edge.rs:72-78) */
#define ORC_LAMBDA_JACOBI 1       /* cyclic Jacobi, independent cross-check */
#define ORC_LAMBDA_MULTISECTION 2 /* what the CUDA kernels do, bit-identical: matching-polynomial Newton for N <= 22, 32-ary section of the subtree recursion above */
#define ORC_LAMBDA_SECTION_ONLY 3 /* the 32-ary section at every N */
#define ORC_LAMBDA_POLY 4         /* the matching-polynomial method (N <= 22 only) */

typedef struct orc_counters {
    uint64_t n_sel;    /* next_action calls on an active node            (next_action.rs:11) */
    uint64_t d_sel;    /* children examined by revisit_choice            (next_action.rs:32-52) */
    uint64_t n_cur;    /* max_curiosity calls                            (next_action.rs:55) */
    uint64_t n_cand;   /* predictions scanned by max_curiosity           (next_action.rs:64-71) */
    uint64_t n_probe;  /* transposition-map lookups                      (tree/mod.rs:170) */
    uint64_t n_ins;    /* nodes inserted (= cost evaluations)            (tree/mod.rs:181-188) */
    uint64_t n_term;   /* of which terminal                              (tree/mod.rs:199) */
    uint64_t n_hit;    /* transposition hits                             (tree/mod.rs:172) */
    uint64_t n_arc;    /* arcs added                                     (graph_operations.rs:26) */
    uint64_t n_pred;   /* predictions appended by add_actions            (graph_operations.rs:50) */
    uint64_t n_cn;     /* cascade node updates                           (empty_transitions.rs:61,100) */
    uint64_t d_cn;     /* cascade parent links followed                  (empty_transitions.rs:78,118) */
    uint64_t n_reset;  /* walker resets to the root                      (tree/mod.rs:176,206) */
    uint64_t n_live;   /* root-steps that ended on a new node (= simulations) */
    uint64_t n_noop;   /* root-steps on an exhausted root                (tree/mod.rs:221-224) */
    uint64_t n_visit;  /* Visited(e) moves                               (tree/mod.rs:139) */
} orc_counters;

/* ---- stand-alone domain functions (tests) ---- */
uint32_t orc_colex_position(uint32_t u, uint32_t v);                    /* edge.rs:48-53 */
void orc_from_colex_position(uint32_t pos, uint32_t *max, uint32_t *min); /* edge.rs:55-65 */
uint32_t orc_action_dim(uint32_t n);                                    /* space.rs:48 */
uint32_t orc_c_upper(uint32_t n);                                       /* 04-c21-tree.rs:59-68 */
/* cost of one tree: lambda_1 (f64), mu, squished eval c (f32). returns 0, or 1 if lambda_1 < 1.4 (ordered_edge.rs:79 panics) */
int orc_cost(uint32_t n, const uint8_t *parents, int method, float c_lower, float c_upper,
             double *lambda1, uint32_t *mu, float *c);
/* SURVEY 8(f) row 3 -- connected graphs as neighbourhood bit sets (B32, N <= 32; connected_bitset_graph/mod.rs) */
int orc_graph_is_cut_edge(uint32_t n, const uint32_t *nbr, uint32_t v, uint32_t u);          /* :45-71 */
void orc_graph_action_kinds(uint32_t n, const uint32_t *nbr, uint32_t *kinds /*[ceil(N(N-1)/32)]*/); /* :134-154, action.rs:10-19 */
uint32_t orc_graph_matching_number(uint32_t n, const uint32_t *nbr);                         /* :218-317 */
int orc_graph_cost(uint32_t n, const uint32_t *nbr, double *lambda1, uint32_t *mu);          /* :200-216, :319-337 */
uint32_t orc_matching_greedy(uint32_t n, const uint8_t *parents);       /* independent O(N) matching for cross-checks */
uint32_t orc_matching_poly(uint32_t n, const uint8_t *parents);         /* degree of the matching polynomial (N <= 22) */
/* legal actions of (parents, permitted mask) ascending; returns count (space.rs:75-89) */
uint32_t orc_action_data(uint32_t n, const uint8_t *parents, const uint32_t *permitted_mask, uint32_t *out);
void orc_act(uint32_t n, uint8_t *parents, uint32_t *permitted_mask, uint32_t action); /* space.rs:56-73 */
void orc_write_vec(uint32_t n, const uint8_t *parents, const uint32_t *permitted_mask, float *vec); /* space.rs:91-101 */

/* ---- synthetic inputs ---- */
void orc_generate_roots(uint64_t seed, uint64_t first_root, uint32_t count, uint32_t n,
                        uint32_t k_min, uint32_t k_max, uint8_t *parents, uint32_t *permitted_mask);
void orc_hash_priors(uint64_t seed, uint64_t first_root, uint32_t count, uint32_t a_dim,
                     uint64_t step, float *out);

/* ---- fp32 MLP forward, row-major weights [in][out] per layer then bias[out] (model/dfdx.rs:69-84; 04-c21-tree.rs:46-52) ---- */
void orc_mlp_forward(const float *params, const uint32_t *dims /*5*/, uint32_t rows,
                     const float *x, float *y, int n_threads);

/* ---- the optimizer (optimizer/mod.rs:7-22) ---- */
typedef struct orc_optimizer orc_optimizer;

orc_optimizer *orc_create(uint32_t n, uint32_t n_roots, float c_lower, float c_upper,
                          const uint32_t *n_as_tol, uint32_t n_as_tol_len, uint32_t n_as_tol_default,
                          int lambda_method, int n_threads);
void orc_destroy(orc_optimizer *o);
void orc_set_roots(orc_optimizer *o, const uint8_t *parents, const uint32_t *permitted_mask);
/* tail of par_new (optimizer/mod.rs:62-101): costs, root node, add_actions(root, priors), argmin over roots */
int orc_init_trees(orc_optimizer *o, const float *priors);
/* par_reset_trees' modify_root half with the example's policy (optimizer/mod.rs:284-339; 04-c21-tree.rs:172-206);
 * counter-generator draws keyed by (seed, epoch, global root). Follow with orc_init_trees. Returns 6 on the reference's panics */
int orc_modify_roots(orc_optimizer *o, uint64_t seed, uint64_t epoch, uint64_t first_root, uint32_t k_min, uint32_t k_max);
void orc_get_roots(orc_optimizer *o, uint8_t *parents, uint32_t *permitted_mask);
/* tail of par_reset_trees (optimizer/mod.rs:340-359): like orc_init_trees but argmin_data is kept */
int orc_reinit_trees(orc_optimizer *o, const float *priors);
/* write_vec of every root state (optimizer/mod.rs:65-70) */
void orc_root_vecs(orc_optimizer *o, float *state_vecs);
/* first parallel region of par_roll_out_episodes (optimizer/mod.rs:159-174); state_vecs may be NULL */
int orc_rollout(orc_optimizer *o, float *state_vecs);
/* second parallel region + par_update_argmmim_data (optimizer/mod.rs:177-190); *improved = 1 if the argmin improved */
int orc_add_actions(orc_optimizer *o, const float *priors, int *improved);
/* n_steps full steps with counter-hash priors of (seed, first_root, step0 + i); improved_steps gets the improving step indices */
int orc_steps_hash(orc_optimizer *o, uint64_t seed, uint64_t first_root, uint64_t step0, uint32_t n_steps,
                   uint32_t *improved_steps, uint32_t cap, uint32_t *n_improved);
void orc_get_counters(orc_optimizer *o, orc_counters *out);
void orc_reset_counters(orc_optimizer *o);
/* argmin (log.rs:1-11) */
void orc_get_argmin(orc_optimizer *o, uint8_t *parents, uint32_t *permitted_mask, double *lambda1,
                    uint32_t *mu, float *eval);
/* walkers: persistent per-root state (optimizer/mod.rs:10-14) */
void orc_get_walkers(orc_optimizer *o, uint8_t *parents, uint32_t *permitted_mask, uint32_t *path_mask,
                     uint32_t *pos, uint32_t *path_len);
/* canonical dump of one search DAG */
void orc_tree_sizes(orc_optimizer *o, uint32_t tree, uint32_t *n_nodes, uint32_t *n_arcs, uint32_t *n_preds);
/* nodes: 6 x u32 per node {c bits, c* bits, n_t, exhausted, lo, hi}; keys: W words per node;
 * preds: 3 x u32 {a_id, g bits, edge_id|ORC_NONE}; arcs: 3 x u32 {src, dst, prediction_pos} */
void orc_dump_tree(orc_optimizer *o, uint32_t tree, uint32_t *nodes, uint32_t *keys, uint32_t *preds, uint32_t *arcs);
/* epoch boundary (next rows): observations (tree/mod.rs:242-264, optimizer/mod.rs:262-278) */
void orc_write_observations(orc_optimizer *o, uint32_t n_obs_tol, float *state_vecs, float *observations, float *weights);

#ifdef __cplusplus
}
#endif
#endif
