/* lgcn_b200.h -- C ABI of the B200-native (sm_100a) LightGCN hot path.
 *
 * Drop-in boundary for the operators below the reference's Python module API
 * (paths relative to the reference checkout /root/reference):
 *
 *   lgcn_graph_build      replaces gcn_norm's per-layer degree scatter + the implicit COO order
 *                         that LGConv consumes           models/light_gcn.py:33  (PyG gcn_conv.py::gcn_norm)
 *   lgcn_propagate_fwd    replaces `for conv in self.convs` + stack/mean/split
 *                                                        models/light_gcn.py:29-38
 *   lgcn_propagate_bwd    replaces autograd of the above (loss.backward())
 *                                                        utils/train_test.py:94
 *   lgcn_bpr_fwd_bwd      replaces the six row gathers + bpr_loss + their autograd
 *                                                        utils/train_test.py:128-132, :18-64
 *   lgcn_clip_adam        replaces clip_grad_norm_(1) + Adam.step (dense, both tables)
 *                                                        utils/train_test.py:95-96, :236
 *   lgcn_train_step       one call = the loop body       utils/train_test.py:88-96
 *   lgcn_cluster_extract  replaces ClusterData partition/permute/__getitem__ + n_id remap
 *                                                        data/dataset_handler.py:273-282
 *   lgcn_partition_metis  the METIS call ClusterData makes (host, third-party library)
 *                                                        data/dataset_handler.py:273
 *   lgcn_score_topk       replaces normalise + matmul + sort + exclusion loop, batched over users
 *                                                        utils/recommend.py:39-61, utils/train_test.py:191-197
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a CUDA stream passed as void* (cudaStream_t).
 *   - the CALLER allocates every output and workspace (sizes from the *_bytes queries);
 *     the library keeps no global state and never allocates device memory.
 *   - node ids are int64 at the boundary exactly as the reference passes them
 *     (edge_index LongTensor [2,E], users 0..U-1, movies U..U+I-1), int32 internally.
 *   - every function returns 0 on success or a negative LGCN_E_* code; lgcn_last_error()
 *     gives the message for the calling thread.  No function synchronises the stream unless
 *     its comment says so.
 *   - embedding dimension is fixed at LGCN_DIM = 64 fp32 (256-byte rows), the reference's dim_h
 *     (utils/train_test.py:274, models/light_gcn.py:14).
 */
#ifndef LGCN_B200_H
#define LGCN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCN_DIM 64
#define LGCN_ROW_SPLIT 512          /* max edges handled by one warp task (large graphs)  */
#ifndef LGCN_ROW_SPLIT_SMALL          /* (compile-time tuning knob of the library build) */
#define LGCN_ROW_SPLIT_SMALL 64     /* same, for edge lists below LGCN_SMALL_GRAPH edges: the longest
                                       task is the critical path of a launch-sized Cluster-GCN batch */
#endif
#define LGCN_SMALL_GRAPH (1 << 20)

#define LGCN_OK 0
#define LGCN_E_INVALID   (-1)       /* bad argument (null pointer, negative size, K<1 ...) */
#define LGCN_E_CUDA      (-2)       /* a CUDA runtime call or launch failed                */
#define LGCN_E_WORKSPACE (-3)       /* workspace smaller than the *_bytes query            */
#define LGCN_E_RANGE     (-4)       /* ids do not fit the int32 internal range             */
#define LGCN_E_METIS     (-5)       /* METIS returned an error                             */

const char *lgcn_last_error(void);
int lgcn_version(void);

/* One warp task: edges [begin,end) of `row` in one of the two CSRs.  slot < 0: the task owns
 * the whole row.  slot >= 0: the row is split; this task writes partial `slot`, the row's
 * first partial is `slot - part`, and it has `nparts` of them.  deg_in / deg_out: the row's
 * in- and out-degree in this edge list, so a kernel that walks tasks needs no per-node array. */
typedef struct lgcn_task {
    int32_t row, begin, end, slot, part, nparts, deg_in, deg_out;
} lgcn_task;

/* Device-resident, immutable description of one edge list (built once per edge list and
 * cached by the host side; the reference rebuilds the normalisation in every layer). */
typedef struct lgcn_graph {
    int32_t num_nodes, num_users;
    int64_t num_edges;
    int64_t num_triplets;            /* P = #edges with source < num_users (utils/helpers.py:98) */
    /* CSR by TARGET (forward):   in_nbr = source ids, stable in original edge order */
    const int32_t *in_ptr, *in_nbr, *in_trip;
    /* CSR by SOURCE (backward):  out_nbr = target ids                               */
    const int32_t *out_ptr, *out_nbr, *out_trip;
    const float *dis;                /* [N] in-degree^-1/2, 0 where the in-degree is 0       */
    const uint8_t *active;           /* [N] 1 if the node has any incident edge              */
    const lgcn_task *in_tasks, *out_tasks;
    int32_t n_in_tasks, n_out_tasks;       /* tasks are sorted by row                        */
    int32_t n_in_user_tasks, n_out_user_tasks; /* how many of them have row < num_users      */
    int32_t n_in_slots, n_out_slots;       /* partial slots (rows longer than LGCN_ROW_SPLIT) */
    float *partials;                 /* [max(n_in_slots,n_out_slots) * 80] scratch            */
    int32_t *slot_counters;          /* [max slots] zero between launches                     */
    int32_t num_active;              /* nodes with active[n] == 1                             */
    int32_t row_split;               /* max edges per task used for this graph                */
    const int32_t *active_list;      /* [num_active] ascending ids of the active nodes        */
    int32_t *sched;                  /* [2] zero between launches: dynamic task scheduler state */
    int32_t in_src_sorted;           /* 1: inside every CSR-by-target row the source ids ascend (true for edge lists in
                                        the reference's (row, col) order and their subsets) -- lets a rank find its own
                                        users' slice of an item row by bisection (lgcn_bpr_fwd_bwd_range) */
    int32_t reserved0;
} lgcn_graph;

/* ---- K0: graph build ------------------------------------------------------------------ */

/* Sizes (bytes) of the arrays a graph of N nodes / E edges needs; the caller allocates them. */
typedef struct lgcn_graph_sizes {
    size_t ptr_bytes;      /* in_ptr / out_ptr: (N+1) int32 each                   */
    size_t nbr_bytes;      /* in_nbr / in_trip / out_nbr / out_trip: E int32 each  */
    size_t dis_bytes;      /* N float                                              */
    size_t active_bytes;   /* N uint8                                              */
    size_t task_bytes;     /* upper bound for each task list                       */
    size_t partial_bytes;  /* upper bound for partials                             */
    size_t counter_bytes;  /* upper bound for slot_counters                        */
    size_t workspace_bytes;/* temporary storage for lgcn_graph_build               */
    size_t active_list_bytes; /* N int32                                            */
} lgcn_graph_sizes;

int lgcn_graph_sizes_query(int64_t num_nodes, int64_t num_edges, lgcn_graph_sizes *out);

/* Builds both CSRs, degrees, dis, active flags and task lists from edge_index [2,E] int64
 * (row-major: E sources then E targets, device memory).  `g` is a HOST struct whose pointer
 * fields the caller has pointed at device arrays of the queried sizes; the scalar fields are
 * filled on return.  SYNCHRONISES the stream once (to read the counts back).
 * Returns LGCN_E_RANGE if N or E exceed int32, LGCN_E_INVALID if an id is outside [0,N) or an
 * edge is not user<->movie (source<U xor target<U): the reference's two triplet masks
 * (utils/helpers.py:98-99) only agree on bipartite lists. */
int lgcn_graph_build(const int64_t *edge_index, int64_t num_edges, int64_t num_nodes,
                     int64_t num_users, lgcn_graph *g, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---- K0b: many edge lists -> many graphs in one pass --------------------------------------
 * A Cluster-GCN epoch hands the path ~100 small edge lists (data/dataset_handler.py:277-285 ->
 * utils/train_test.py:86-88 `for batch in train_loader: batch.to(device)`).  lgcn_graph_build_batched
 * builds the lgcn_graph of every list with a fixed number of launches and ONE stream sync; each
 * graph is bit-identical to what lgcn_graph_build gives for that list alone.
 *
 *   edges      device int64: the lists back to back, list b = its contiguous [2,E_b] tensor
 *              (E_b sources then E_b targets) at int64 offset 2*edge_off[b]
 *   edge_off   HOST int64 [B+1], edge_off[0] = 0, ascending (edge counts prefix sum)
 *   graphs     HOST lgcn_graph [B], completely filled on return; all pointers point into `arena`,
 *              which must stay alive as long as the graphs are used
 *   workspace  scratch, free again when the call returns
 * The B graphs SHARE partials / slot_counters / sched: use them one after another on one stream.
 * LGCN_E_RANGE when B*(N+1) or the total edge count exceed int32: build in chunks of lists. */
typedef struct lgcn_batched_sizes {
    size_t arena_bytes;
    size_t workspace_bytes;
} lgcn_batched_sizes;

int lgcn_graph_batched_sizes(int64_t num_nodes, int64_t num_lists, const int64_t *edge_off,
                             lgcn_batched_sizes *out);
int lgcn_graph_build_batched(const int64_t *edges, const int64_t *edge_off, int64_t num_lists,
                             int64_t num_nodes, int64_t num_users, lgcn_graph *graphs, void *arena,
                             size_t arena_bytes, void *workspace, size_t workspace_bytes, void *stream);

/* `batch.to(device)` (utils/train_test.py:87) for all lists of an epoch: HOST list b (contiguous [2,E_b] int64,
 * pinned or pageable) -> dst + 2*edge_off[b] (device), one asynchronous copy per list on `stream`. */
int lgcn_upload_lists(const int64_t *const *host_lists, const int64_t *edge_off, int64_t num_lists,
                      int64_t *dst, void *stream);

/* ---- K1/K2: propagation ----------------------------------------------------------------- */

/* final[N,64] = (sum_{k=0..K} A^k e0) / (K+1)^2 with A[c,r] = dis[r] dis[c] per edge r->c
 * (models/light_gcn.py:28-40).  e0 is given as the two weight tables.  work: (K-1)*N*64 floats
 * (none for K == 1).  rnorm (optional, [N]) receives 1/||final[n]||_2. */
int lgcn_propagate_fwd(const lgcn_graph *g, const float *user_w, const float *item_w,
                       int num_layers, float *final_out, float *rnorm, float *work,
                       size_t work_bytes, void *stream);

/* grad_e0[N,64] = (sum_{k=0..K} (A^T)^k G) / (K+1)^2, evaluated Horner-style.  work:
 * 2*N*64 floats (none for K == 0).  If reg_coef != 0 the BPR regulariser's gradient
 * reg_coef * cnt[n] * e0[n] is added and sum(cnt * ||e0||^2) goes to accum[1]; the squared
 * norm of grad_e0 is ADDED to accum[2] (double, device).  cnt may be null when reg_coef == 0. */
int lgcn_propagate_bwd(const lgcn_graph *g, const float *grad_final, int num_layers,
                       const float *user_w, const float *item_w, const int32_t *neg_count,
                       float reg_coef, float *grad_e0, double *accum, float *work,
                       size_t work_bytes, void *stream);

/* One LGConv layer on its own: out[N,64] = A x (transpose = 0) or A^T x (transpose != 0) -- the
 * reference's inner operator `conv(x=emb, edge_index=edge_index)` (models/light_gcn.py:33). */
int lgcn_spmm(const lgcn_graph *g, const float *x, float *out, int transpose, void *stream);

/* ---- K3: BPR loss forward + gradient w.r.t. the final embeddings ------------------------ */

/* Triplets are (source u, target p, neg[t]) for the t-th edge with source < U in edge order
 * (utils/helpers.py:84-102); neg [P] int64 item ids in [0,I).  Writes grad_final [N,64]
 * (every row), neg_count [I] (histogram of neg), trip_scratch [2*P] floats, and ADDS
 * sum_t softplus(10 (cos+ - cos-)) to accum[0] (double).  final/rnorm come from
 * lgcn_propagate_fwd. */
int lgcn_bpr_fwd_bwd(const lgcn_graph *g, const float *final_emb, const float *rnorm,
                     const int64_t *neg, float *grad_final, int32_t *neg_count,
                     float *trip_scratch, double *accum, void *stream);

/* bpr_loss on six already-gathered [P,64] tensors (utils/train_test.py:18-51), for callers that
 * compose compute_embeddings + bpr_loss themselves.  accum: 2 doubles of scratch; loss_out: device
 * float (optional); gradients (optional, all six or none) are multiplied by *grad_scale (device
 * float, null = 1). */
int lgcn_bpr_rows(const float *uf, const float *u0, const float *pf, const float *p0, const float *nf,
                  const float *n0, int64_t P, float coeff, double *accum, float *loss_out,
                  const float *grad_scale, float *g_uf, float *g_u0, float *g_pf, float *g_p0,
                  float *g_nf, float *g_n0, void *stream);

/* ---- K6: clip_grad_norm_(max_norm) + Adam ------------------------------------------------ */

typedef struct lgcn_adam {
    double lr, beta1, beta2, eps, max_norm;   /* doubles: torch derives 1-beta, lr/(1-beta^t) in double */
    int64_t *step;          /* device int64: incremented by lgcn_step_begin             */
    float *m, *v;           /* [N,64] exp_avg / exp_avg_sq                              */
    /* optional: bias-correction table computed by the host exactly as torch does (Python doubles):
     * bc_table[2t] = lr/(1-beta1^t), bc_table[2t+1] = sqrt(1-beta2^t) for t < bc_len (device floats).
     * Steps beyond the table fall back to in-kernel pow(). */
    const float *bc_table;
    int64_t bc_len;
    /* optional (sparse steps only): row_step[r] = last optimiser step applied to row r (device) */
    int32_t *row_step;
} lgcn_adam;

/* Zeroes accum[0..3] and increments *opt->step (one tiny launch). */
int lgcn_step_begin(const lgcn_adam *opt, double *accum, void *stream);

/* grad scaled by min(1, max_norm / (sqrt(accum[2]) + 1e-6)), then torch.optim.Adam's update
 * on both tables.  Also writes loss = -accum[0]/(10 P) + bpr_coeff/(64 P) * accum[1] to
 * loss_out[0] if loss_out != null.
 * Arithmetic: torch's formula and scalar handling (bias corrections in double, one cast to fp32); the square
 * root and the two divisions per element use the hardware approximations (<= 1 / 2 ulp), so an update differs
 * from torch's correctly rounded one by <= ~6e-10 at lr = 1e-3.  All Adam entry points (dense, row lists, the
 * replay of deferred steps) share this arithmetic and are bit-identical to each other. */
int lgcn_clip_adam(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users,
                   int64_t num_items, const float *grad, const double *accum,
                   int64_t num_triplets, float bpr_coeff, float *loss_out, void *stream);

/* ---- sharded (owner-computes by node range) building blocks -------------------------------
 * Every rank holds the FULL lgcn_graph and executes only the warp tasks whose rows it owns
 * ([task_begin,task_end) of the row-sorted task list covering rows [row_begin,row_end)); the
 * caller exchanges the produced row slabs between layers (NCCL all-gather / all-reduce, see
 * lgcn_b200/sharded.py).  Layer 1 here reads the PRE-SCALED table y0 = dis (.) e0. */
/* Fused compute + all-gather.  When `peers` is given (world > 1) the tables the kernels PRODUCE
 * (y0 / yout / final_out / rnorm / zout) must live at the same offset of a symmetric-memory region on
 * every rank: base[p] is rank p's mapping of that region in this process (peer access over NVLink),
 * mc_base the NVLS multicast alias (or NULL).  Each produced row is then stored into every rank's copy
 * from the SpMM epilogue itself -- one multimem store through the NVSwitch, or `world` peer stores --
 * so the transfer overlaps the gather-sum row by row and no separate all-gather runs; the caller only
 * places a cross-rank barrier between layers.  peers == NULL: plain local stores. */
typedef struct lgcn_peers {
    int32_t world, rank;
    void *mc_base;
    void *base[8];
} lgcn_peers;

/* Barrier between the ranks of `peers` on `stream`: flags_local = this rank's int32[8] flag array
 * inside the symmetric region (zero-initialised), epoch = private device int32 (zero-initialised).
 * Orders the peer / multicast stores of the preceding kernels before the following ones. */
int lgcn_peer_barrier(const lgcn_peers *peers, int32_t *flags_local, int32_t *epoch, void *stream);

/* Sum of accum[0..3] (doubles, device) over the ranks, in rank order, in place; also a barrier like lgcn_peer_barrier
 * (same flag array and epoch).  slots_local = this rank's double[8*4] slot table inside the symmetric region.
 * Replaces the all-reduce of the loss / clip sums (utils/train_test.py:95: clip_grad_norm_ needs the GLOBAL norm). */
int lgcn_peer_allreduce4(const lgcn_peers *peers, int32_t *flags_local, int32_t *epoch, double *slots_local,
                         double *accum, void *stream);

int lgcn_prescale(const lgcn_graph *g, const float *user_w, const float *item_w, int64_t row_begin,
                  int64_t row_end, float *y0, const lgcn_peers *peers, void *stream);
int lgcn_fwd_layer(const lgcn_graph *g, const float *user_w, const float *item_w, int k, int num_layers,
                   const float *yin, float *yout, const float *y1, const float *y2, const float *y3,
                   float *final_out, float *rnorm, int task_begin, int task_end, int64_t row_begin,
                   int64_t row_end, const lgcn_peers *peers, void *stream);
/* lgcn_fwd_layer with flags.  LGCN_FWD_NORMALIZED (last layer only): final_out receives the L2-NORMALISED final rows
 * final / ||final|| (every rank's copy) and rnorm[row] = 1 / ||final[row]|| is stored LOCALLY only -- the form
 * lgcn_bpr_owner consumes (cosines become plain dot products, utils/train_test.py:53-64 applied once per row). */
#define LGCN_FWD_NORMALIZED 1
int lgcn_fwd_layer_ex(const lgcn_graph *g, const float *user_w, const float *item_w, int k, int num_layers,
                      const float *yin, float *yout, const float *y1, const float *y2, const float *y3,
                      float *final_out, float *rnorm, int task_begin, int task_end, int64_t row_begin,
                      int64_t row_end, int flags, const lgcn_peers *peers, void *stream);
/* j == 1: zin may be the pre-scaled table z_0 = dis (.) grad_final (what lgcn_bpr_owner stores as zG); NULL: the layer
 * gathers grad_final itself and applies dis per edge. */
int lgcn_bwd_layer(const lgcn_graph *g, const float *grad_final, int j, int num_layers, const float *zin,
                   float *zout, const float *user_w, const float *item_w, const int32_t *neg_count,
                   float reg_coef, float *grad_e0, double *accum, int task_begin, int task_end,
                   int64_t row_begin, int64_t row_end, const lgcn_peers *peers, void *stream);
/* lgcn_bpr_fwd_bwd restricted to the triplets of users [user_row_begin,user_row_end) (out-task range
 * [user_task_begin,user_task_end)); grad_final / neg_count are zeroed first so that the per-rank
 * results sum to the unsharded ones. */
int lgcn_bpr_fwd_bwd_range(const lgcn_graph *g, const float *final_emb, const float *rnorm,
                           const int64_t *neg, float *grad_final, int32_t *neg_count,
                           float *trip_scratch, double *accum, int user_task_begin, int user_task_end,
                           int64_t user_row_begin, int64_t user_row_end, void *stream);
/* ---- owner-computes BPR (full-graph step, any world size) ------------------------------------------
 * lgcn_bpr_fwd_bwd with every row of dL/dfinal produced by ONE warp task (no float atomics, bit-stable):
 * the triplets' negatives are bucketed by item each step, user rows / item rows are walked by their owner.
 * Replaces the same reference lines (utils/train_test.py:18-64,128-132 + autograd), for the rows
 *   users: out-tasks [user_task_begin,user_task_end) of g      items: [item_begin,item_end) (item ids),
 *   whose in-tasks are [item_task_begin,item_task_end).
 * Writes G[row] (local, rows owned) and zG[row] = dis[row] * G[row] into every rank's copy (`peers`, see above),
 * neg_count[item_begin..item_end) (histogram of neg over the owned items), and ADDS sum_t softplus(..) of the owned
 * users' triplets to accum[0].  out_trip / in_trip of g must hold GLOBAL triplet ids (lgcn_graph_remap_triplets when
 * g was built from a shard of the edge list); num_triplets = the GLOBAL count P (loss normalisation);
 * final_hat = the L2-NORMALISED final rows (lgcn_fwd_layer_ex with LGCN_FWD_NORMALIZED), complete on this rank for every
 * row referenced; rnorm = 1/||final|| of the OWNED rows; trip_user / trip_pos complete for every triplet.
 * G rows of owned nodes without tasks are never written: zero-initialise G once.
 * ws.scalars != NULL (4*P floats; valid only when this call covers ALL users and items): per-triplet scalars are
 * passed from the user pass to the item passes; NULL: they are recomputed from the rows (bit-identical). */
typedef struct lgcn_bpr_owner_ws {
    const int32_t *trip_user;   /* [P] user row of triplet t           (lgcn_triplet_index) */
    const int32_t *trip_pos;    /* [P] positive item NODE id of triplet t                    */
    int32_t *bucket_ptr;        /* [item_end - item_begin + 1]                                */
    int32_t *bucket_cursor;     /* [item_end - item_begin]                                    */
    int32_t *bucket;            /* [bucket_cap] triplet ids grouped by negative item          */
    int64_t bucket_cap;         /* >= P                                                       */
    float *scalars;             /* [4*P] or NULL                                              */
    int32_t *sched;             /* [2] zero between launches                                  */
} lgcn_bpr_owner_ws;

int lgcn_bpr_owner(const lgcn_graph *g, const float *final_hat, const float *rnorm, const int64_t *neg,
                   int64_t num_triplets, float *G, float *zG, int32_t *neg_count, double *accum,
                   const lgcn_bpr_owner_ws *ws, int user_task_begin, int user_task_end, int item_task_begin,
                   int item_task_end, int64_t item_begin, int64_t item_end, const lgcn_peers *peers, void *stream);

/* The two halves of lgcn_bpr_owner, for callers that overlap them with an exchange: lgcn_bpr_buckets groups this
 * step's negatives by item for items [item_begin,item_end) (needs only `neg`: it can run while the forward tables are
 * still in flight); lgcn_bpr_owner_passes runs the user / negative / positive passes on buckets built that way.
 * lgcn_bpr_owner == lgcn_bpr_buckets followed by lgcn_bpr_owner_passes. */
int lgcn_bpr_buckets(const int64_t *neg, int64_t num_triplets, int64_t item_begin, int64_t item_end,
                     int32_t *neg_count, const lgcn_bpr_owner_ws *ws, void *stream);
int lgcn_bpr_owner_passes(const lgcn_graph *g, const float *final_hat, const float *rnorm, const int64_t *neg,
                          int64_t num_triplets, float *G, float *zG, int32_t *neg_count, double *accum,
                          const lgcn_bpr_owner_ws *ws, int user_task_begin, int user_task_end, int item_task_begin,
                          int item_task_end, int64_t item_begin, int64_t item_end, const lgcn_peers *peers, void *stream);

/* trip_user[t] / trip_pos[t] for the triplets of the user rows covered by out-tasks [user_task_begin,user_task_end),
 * stored into every rank's copy of the two tables (symmetric memory when peers != NULL).  One-off per edge list. */
int lgcn_triplet_index(const lgcn_graph *g, int user_task_begin, int user_task_end, int32_t *trip_user,
                       int32_t *trip_pos, const lgcn_peers *peers, void *stream);

/* g was built from a SUBSET of an edge list (a rank's shard): rewrites its local triplet numbers
 * (rank among the subset's user->movie edges) to the global ones, trip_global [g->num_triplets] int32 (device). */
int lgcn_graph_remap_triplets(lgcn_graph *g, const int32_t *trip_global, void *stream);

/* lgcn_clip_adam over rows [row_begin,row_end) only (accum must already hold the GLOBAL sums). */
int lgcn_clip_adam_rows(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users,
                        int64_t num_items, const float *grad, const double *accum,
                        int64_t num_triplets, float bpr_coeff, float *loss_out, int64_t row_begin,
                        int64_t row_end, void *stream);

/* ---- fused training step ------------------------------------------------------------------ */

typedef struct lgcn_step_buffers {
    float *final_emb;      /* [N,64] */
    float *rnorm;          /* [N]    */
    float *grad_final;     /* [N,64] */
    float *grad_e0;        /* [N,64] */
    float *work;           /* max(K-1, 2) * N * 64 floats */
    size_t work_bytes;
    int32_t *neg_count;    /* [I]    */
    float *trip_scratch;   /* [2*Pmax] */
    double *accum;         /* [4]    */
    /* sparse steps only */
    int32_t *neg_flag;     /* [I] zero-initialised: last step in which the item was a negative */
    int32_t *neg_list;     /* [I] distinct INACTIVE negative items of the current step          */
    int32_t *neg_list_count; /* [1]                                                              */
    /* lgcn_train_steps_sparse only */
    int32_t *act_stamp;    /* [2*N + 4*I] zero-initialised: per step parity, the step in which a node is
                            * active [2][N], an item is an inactive negative [2][I], and the list of those [2][I] */
} lgcn_step_buffers;

/* utils/train_test.py:88-96 for one batch: forward, BPR loss, backward, clip, Adam.
 * loss_out: device float receiving this batch's loss (the reference's train_loss.item(),
 * read back by the caller whenever it chooses -- no per-step sync). */
int lgcn_train_step(const lgcn_graph *g, float *user_w, float *item_w, int num_layers,
                    const int64_t *neg, float bpr_coeff, const lgcn_adam *opt,
                    const lgcn_step_buffers *buf, float *loss_out, void *stream);

/* Same step, work proportional to the rows the batch TOUCHES (nodes with an incident edge + sampled
 * negatives) instead of N: identical arithmetic, but rows the batch does not touch are not visited --
 * their (zero-gradient) Adam updates are replayed exactly, per row, the next time the row is touched
 * (opt->row_step) or by lgcn_adam_flush.  For Cluster-GCN batches, where the reference pays full-table
 * passes per batch.  Requires opt->row_step, opt->bc_table, buf->neg_*;
 * buf->grad_final and buf->neg_count must be all-zero on entry and are all-zero again on exit. */
int lgcn_train_step_sparse(const lgcn_graph *g, float *user_w, float *item_w, int num_layers,
                           const int64_t *neg, float bpr_coeff, const lgcn_adam *opt,
                           const lgcn_step_buffers *buf, float *loss_out, void *stream);

/* A RUN of sparse steps in one persistent cooperative launch (epoch_kernel.cu): the loop
 * `for batch in train_loader:` of utils/train_test.py:86-101 for `num_steps` consecutive batches.
 * graphs: HOST array of built graphs (lgcn_graph_build / lgcn_graph_build_batched), every one with
 * num_triplets > 0; neg: device int64, the steps' negatives back to back (step b has
 * graphs[b].num_triplets of them); loss_out: device float [num_steps]; workspace: device,
 * lgcn_train_steps_workspace_bytes(num_steps).  Same arithmetic and invariants as
 * lgcn_train_step_sparse; additionally needs buf->act_stamp.  buf->trip_scratch must hold
 * 2 * max_b num_triplets floats; opt->bc_table must cover the steps of the run (current step +
 * num_steps < bc_len).  Does not synchronise. */
size_t lgcn_train_steps_workspace_bytes(int64_t num_steps);
int lgcn_train_steps_sparse(const lgcn_graph *graphs, int64_t num_steps, float *user_w, float *item_w,
                            int num_layers, const int64_t *neg, float bpr_coeff, const lgcn_adam *opt,
                            const lgcn_step_buffers *buf, float *loss_out, void *workspace,
                            size_t workspace_bytes, void *stream);

/* Brings every row up to the current step (replays the pending zero-gradient updates). */
int lgcn_adam_flush(const lgcn_adam *opt, float *user_w, float *item_w, int64_t num_users,
                    int64_t num_items, void *stream);

/* Loss only (evaluate(), utils/train_test.py:153-156): forward + BPR value, no gradients. */
int lgcn_eval_loss(const lgcn_graph *g, const float *user_w, const float *item_w, int num_layers,
                   const int64_t *neg, float bpr_coeff, const lgcn_step_buffers *buf,
                   float *loss_out, void *stream);

/* ---- K4: Cluster-GCN sub-graph extraction ------------------------------------------------- */

/* Host call into METIS with torch_sparse's arguments (no weights, default options, kway).
 * indptr [N+1], index [E] and part_out [N] are HOST int64 arrays. */
int lgcn_partition_metis(int64_t num_nodes, const int64_t *indptr, const int64_t *index,
                         int64_t num_parts, int64_t *part_out);

size_t lgcn_cluster_extract_workspace_bytes(int64_t num_nodes, int64_t num_edges, int64_t num_parts);

/* Given train edges [2,E] int64 and the partition vector cluster [N] int64 (device), writes the
 * concatenation over parts p = 0..P-1 of {(r,c): cluster[r]==cluster[c]==p} ordered by
 * (inv[r], inv[c]) -- inv = inverse of the stable sort of `cluster` -- as GLOBAL ids into
 * out_edges [2,E] (row-major with stride E), and part_ptr [P+1] offsets into it (int64).
 * part_ptr[P] = number of intra-cluster edges kept. */
int lgcn_cluster_extract(const int64_t *edge_index, int64_t num_edges, int64_t num_nodes,
                         const int64_t *cluster, int64_t num_parts, int64_t *out_edges,
                         int64_t *part_ptr, void *workspace, size_t workspace_bytes, void *stream);

/* GPU partitioner (alternative to the host METIS call above; partitions differ from METIS, so this is a set-up-time /
 * quality trade, not a parity item -- SURVEY sec. 8f rank 4).  One voting pass of balanced label propagation over rows
 * [row_begin,row_end) of a CSR by source (ptr [N+1], nbr [E], device int32): want[v] = the part label carried by most
 * out-neighbours of v (ties: the smallest label; v's own label if it has no out-neighbour), best[v] = that count,
 * own[v] = out-neighbours carrying v's current label.  num_parts <= 4096.  Deterministic integer work; the
 * capacity-constrained acceptance of the moves is host orchestration (lgcn_b200/data/partition_gpu.py). */
int lgcn_label_vote(const int32_t *ptr, const int32_t *nbr, const int32_t *labels, int64_t row_begin, int64_t row_end,
                    int num_parts, int32_t *want, int32_t *best, int32_t *own, void *stream);

/* to_undirected (PyG 2.4.0 utils/undirected.py:to_undirected + coalesce, called at
 * data/dataset_handler.py:141): both directions of every edge of edge_index [2,E] int64 (device), sorted by
 * (row, col), duplicates dropped.  out_edges: device int64, 4*E cells; on return cells [0,count) hold the rows
 * and [count, 2*count) the columns, i.e. a contiguous [2,count] tensor.  *count_out (HOST) = number of distinct
 * directed edges.  Ids outside [0,num_nodes) are rejected (LGCN_E_INVALID).  Synchronises the stream. */
size_t lgcn_to_undirected_workspace_bytes(int64_t num_edges);
int lgcn_to_undirected(const int64_t *edge_index, int64_t num_edges, int64_t num_nodes, int64_t *out_edges,
                       int64_t *count_out, void *workspace, size_t workspace_bytes, void *stream);

/* ---- K5: scoring GEMM + train-edge mask + top-k ------------------------------------------- */

/* For users [u_begin,u_end): score(u,i) = <U[u],I[i]> (rows optionally L2-normalised first, as
 * utils/recommend.py:39-40), items listed in the user's exclusion CSR row (excl_ptr [U+1],
 * excl_idx sorted item ids; may be null) are removed, top-k by (score desc, id asc) written to
 * topk_idx [n_users,k] int32 and topk_val [n_users,k].  The score matrix is never stored. */
int lgcn_score_topk(const float *user_emb, const float *item_emb, int64_t num_items,
                    int64_t u_begin, int64_t u_end, int normalize, const int64_t *excl_ptr,
                    const int32_t *excl_idx, int k, int32_t *topk_idx, float *topk_val,
                    void *stream);

/* Same with an explicit kernel choice: LGCN_SCORE_TENSOR = tcgen05 TF32 MMA with a 3-term hi/lo
 * operand split (fp32-level accuracy), accumulator in TMEM, k <= 32; LGCN_SCORE_FFMA = fp32 FFMA tiles,
 * k <= 128; LGCN_SCORE_AUTO picks the tensor-core kernel whenever k allows. */
#define LGCN_SCORE_AUTO 0
#define LGCN_SCORE_FFMA 1
#define LGCN_SCORE_TENSOR 2
#define LGCN_SCORE_AUTO_USES_TENSOR 1      /* validated on B200 (tests/test_gpu_cluster_score.py) */
int lgcn_score_topk_ex(const float *user_emb, const float *item_emb, int64_t num_items,
                       int64_t u_begin, int64_t u_end, int normalize, const int64_t *excl_ptr,
                       const int32_t *excl_idx, int k, int32_t *topk_idx, float *topk_val, int algo,
                       void *workspace, size_t workspace_bytes, void *stream);
/* Optional workspace of the tensor-core kernel (device, 128-byte aligned): with it the item table is
 * normalised and split once into ready-made shared-memory tile images that the kernel fetches with TMA
 * bulk copies; without it (NULL) every CTA prepares the item tiles itself (slower). */
size_t lgcn_score_topk_workspace_bytes(int64_t num_items);

/* ---- diagnostics ------------------------------------------------------------------------------ */

/* Gathers ctas * 16 * rows_per_halfwarp pseudo-random 256-byte rows of table [nrows,64] with the
 * propagation kernels' access shape and nothing else (no index array, no dependent address): the
 * measured ceiling their gather rate is quoted against (tools/gather_peak.py, bench.py).  sink:
 * ctas * 16 floats, never written in practice. */
int lgcn_probe_gather(const float *table, int64_t nrows, int ctas, int rows_per_halfwarp, float *sink,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LGCN_B200_H */
