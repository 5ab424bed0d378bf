/*
 * gsage_b200.h -- C ABI of the B200-native GraphSAGE sample -> gather -> aggregate -> project engine.
 *
 * This header IS the drop-in boundary.  Every entry point is `extern "C"`, takes plain pointers and
 * sizes (no torch / scipy / numpy types) and names the reference interface it stands in for
 * (file:line under bkj/pytorch-graphsage).  The reference has no FFI of its own (it is 100 % Python);
 * INTEGRATION.md shows the ctypes stub a maintainer would add to nn_modules.py / models.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative gsage_status otherwise; `gsage_last_error()`
 *     returns a thread-local message (the Python host turns it into RuntimeError / IndexError, like
 *     the reference's asserts, nn_modules.py:73,81);
 *   - `*_dev` pointers are device pointers on the current device; `*_host` pointers are host pointers;
 *   - `stream` is a `cudaStream_t` passed as void* (0 = legacy default stream).  Nothing synchronises
 *     the stream unless the name ends in `_host` or the doc says so;
 *   - ids are int64 in the reference's id space (sparse convention: row 0 = dummy node, node i = row i+1,
 *     utils/convert.py:100-126);
 *   - float tables/activations are row-major with an explicit leading dimension `ld` (in elements).
 *     Rows must start 16-byte aligned (ld * sizeof(elem) % 16 == 0) and padding columns must be zero.
 */
#ifndef GSAGE_B200_H
#define GSAGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSAGE_ABI_VERSION 3   /* 2: gsage_layer_weights gained the LSTM weights, gsage_linear_seg.a_rows, new entry points
                               * 3: gsage_engine_profile_read reports bytes AND flops over 8 categories; gsage_engine_inputs_ready,
                               *    gsage_engine_poll_errors, narrow-API backward kernels, metric kernels */

typedef enum gsage_status {
    GSAGE_OK = 0,
    GSAGE_ERR_INVALID = -1,     /* bad argument (shape / dtype / alignment / S <= 0)          */
    GSAGE_ERR_CUDA = -2,        /* a CUDA runtime call failed                                   */
    GSAGE_ERR_INDEX = -3,       /* an id outside the adjacency / table (scipy raises IndexError) */
    GSAGE_ERR_RNG = -4,         /* the device stream ran out of look-ahead (never silently wrong) */
    GSAGE_ERR_NOMEM = -5
} gsage_status;

typedef enum gsage_dtype { GSAGE_F32 = 0, GSAGE_BF16 = 1 } gsage_dtype;
typedef enum gsage_act { GSAGE_ACT_NONE = 0, GSAGE_ACT_RELU = 1, GSAGE_ACT_TANH = 2 } gsage_act;
typedef enum gsage_reduce { GSAGE_RED_MEAN = 0, GSAGE_RED_MAX = 1, GSAGE_RED_SUM = 2 } gsage_reduce;
typedef enum gsage_aggregator {
    GSAGE_AGG_MEAN = 0, GSAGE_AGG_MAX_POOL = 1, GSAGE_AGG_MEAN_POOL = 2, GSAGE_AGG_ATTENTION = 3, GSAGE_AGG_LSTM = 4
} gsage_aggregator;
typedef enum gsage_prep { GSAGE_PREP_IDENTITY = 0, GSAGE_PREP_NODE_EMBEDDING = 1, GSAGE_PREP_LINEAR = 2 } gsage_prep;

typedef struct gsage_graph gsage_graph;     /* device-resident adjacency (both samplers' `adj`)      */
typedef struct gsage_rng gsage_rng;         /* device-resident numpy-legacy MT19937 stream           */
typedef struct gsage_engine gsage_engine;   /* GSSupervised.forward re-expressed over ids            */

int gsage_abi_version(void);
const char* gsage_last_error(void);
/* name, SM count, HBM bytes of the current device */
int gsage_device_info(char* name_out, int name_cap, int* sm_count, int64_t* hbm_bytes, int* cc_major, int* cc_minor);
/* one process per GPU: bind this library's CUDA runtime to `device` (call before creating any object) */
int gsage_set_device(int device);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
int64_t gsage_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Adjacency.  Replaces: `parse_csr_matrix` (problem.py:70-72) and
 * `SparseUniformNeighborSampler.__init__` (nn_modules.py:72-78: scipy CSR + degree table).
 * ------------------------------------------------------------------------------------------------ */

/* From scipy-canonical CSR arrays on the host (sorted indices, duplicates summed).  `indices_host` may be
 * NULL when the matrix follows the reference's file convention (columns of a row are 0..deg-1). */
int gsage_graph_from_csr(const int64_t* indptr_host, const int64_t* indices_host, const int64_t* data_host,
                         int64_t n_rows, int64_t n_cols, gsage_graph** out);
/* From the 3 x nnz [v; r; c] triplets of a sparse problem file (problem.py:70-72 semantics: shape inferred
 * as (max r + 1, max c + 1), duplicate (r, c) summed). */
int gsage_graph_from_triplets(const int64_t* v_host, const int64_t* r_host, const int64_t* c_host, int64_t nnz,
                              gsage_graph** out);
void gsage_graph_destroy(gsage_graph* g);
/* shape = (n_rows, n_cols); canonical = 1 when the column array is redundant and was dropped on device */
int gsage_graph_info(const gsage_graph* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, int* canonical,
                     int64_t* device_bytes);
/* the reference's `sampler.degrees` (nn_modules.py:76-78), copied to the host (n_rows int64) */
int gsage_graph_degrees_host(const gsage_graph* g, int64_t* degrees_host);

/* ------------------------------------------------------------------------------------------------
 * Random stream.  Replaces the global numpy legacy RandomState the reference seeds in
 * helpers.set_seeds (helpers.py:14-18) and draws from in nn_modules.py:88 / problem.py:146.
 * ------------------------------------------------------------------------------------------------ */
int gsage_rng_create(gsage_rng** out);
void gsage_rng_destroy(gsage_rng* r);
int gsage_rng_seed(gsage_rng* r, uint32_t seed, void* stream);                          /* np.random.seed    */
int gsage_rng_set_state(gsage_rng* r, const uint32_t key_host[624], int pos, void* stream);   /* set_state  */
int gsage_rng_get_state(gsage_rng* r, uint32_t key_host[624], int* pos, void* stream);  /* get_state; syncs  */
/* next `count` raw tempered 32-bit words (== np.frombuffer(RandomState.bytes(4*count), '<u4')) */
int gsage_rng_raw(gsage_rng* r, int64_t count, uint32_t* out_dev, void* stream);
/* np.random.choice(hi, count) / legacy randint(0, hi, count): masked rejection, one word per attempt */
int gsage_rng_randint(gsage_rng* r, uint32_t hi, int64_t count, uint32_t* out_dev, void* stream);
/* np.random.permutation(np.arange(n)) (problem.py:146), int64 out */
int gsage_rng_permutation(gsage_rng* r, int64_t n, int64_t* out_dev, void* stream);
/* raises the sticky device error flag, if any (look-ahead exhausted); syncs the stream */
int gsage_rng_check(gsage_rng* r, void* stream);
/* total raw words consumed since the last seed / set_state; syncs the stream */
int gsage_rng_consumed(gsage_rng* r, int64_t* words, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Samplers.  Replace SparseUniformNeighborSampler.__call__ (nn_modules.py:80-101) and
 * UniformNeighborSampler.__call__ (nn_modules.py:42-49).
 * ------------------------------------------------------------------------------------------------ */

/* out[i*S + j] = A[ids[i], sel[i*S + j] % degree(ids[i])]  (0 when the row is empty).  `sel_dev` holds the
 * n*S bounded draws in [0, n_cols) -- drawn by the host from np.random exactly like nn_modules.py:88
 * ("mode A"), or by gsage_rng_randint.  Out-of-range ids raise the graph's sticky index-error flag. */
int gsage_sample_sparse(gsage_graph* g, const int64_t* ids_dev, int64_t n, int S, const uint32_t* sel_dev,
                        int64_t* out_dev, void* stream);
/* same, drawing from the device stream ("mode B": no host round trip, same indices bit for bit) */
int gsage_sample_sparse_rng(gsage_graph* g, gsage_rng* r, const int64_t* ids_dev, int64_t n, int S,
                            int64_t* out_dev, void* stream);
/* GSAGE_ERR_INDEX if any sampler call since the last check saw an id outside [0, n_rows); syncs */
int gsage_graph_check(gsage_graph* g, void* stream);
/* dense 2-D edgelist: out[i, j] = adj[ids[i], perm[j]] for j < S; one shared permutation (torch.randperm(K)) */
int gsage_sample_dense(const int64_t* adj_dev, int64_t n_rows, int K, const int64_t* ids_dev, int64_t n,
                       const int64_t* perm_dev, int S, int64_t* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gather / aggregate.  Replace `feats[ids]` (models.py:76,80), nn.Embedding lookups (nn_modules.py:146,149)
 * and the reductions inside the aggregators (nn_modules.py:197-198, 225-226/240/252, 314-315).
 * ------------------------------------------------------------------------------------------------ */

/* out[i, :d] = table[ids[i], :d] (ids NULL -> identity).  out may be a column slice: out_dev + col, ld_out. */
int gsage_gather_rows(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d,
                      const int64_t* ids_dev, int64_t n, void* out_dev, int out_dtype, int64_t ld_out, void* stream);
/* THE fused gather+aggregate kernel: out[p, :] = reduce_j w[p*S+j] * table[ids[p*S + j], :]
 *   ids NULL      -> rows p*S+j of `table` itself (the contiguous `neibs.view(N, S, d)` case)
 *   weights NULL  -> 1 (MEAN divides by S including dummy rows; MAX ignores weights)
 * fp32 accumulation whatever the table dtype. */
int gsage_gather_reduce(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d,
                        const int64_t* ids_dev, int64_t n_parents, int S, int reduce, const float* weights_dev,
                        void* out_dev, int out_dtype, int64_t ld_out, void* stream);
/* attention weights (nn_modules.py:307-311): w[p, j] = softmax_j <na[p*S+j, :H], xa[p, :H]>, fp32 */
int gsage_attention_weights(const void* na_dev, const void* xa_dev, int dtype, int64_t ld, int H,
                            int64_t n_parents, int S, float* w_dev, void* stream);
/* The whole reduction of the attention aggregator (nn_modules.py:307-315) in one launch:
 *   out[p, :d] = sum_j softmax_j(<a(n_pj), xa[p]>) n_pj,   a(v) = W2 tanh(W1 v + b1),   n_pj = table[ids[p*S + j]]
 * (ids NULL: row p*S + j of `table`).  Scores on the tensor cores, softmax and weighted sum from the same shared-memory
 * tile: every neighbour row is read from HBM once.  `xa_dev` = a(x_p), (n_parents, 32) fp32, computed by the caller with
 * gsage_linear.  bf16 table and W1 (32 x d), W2 (32 x 32) fp32, 2 <= S <= 128; dummy rows are not masked (their score is
 * a(0)-dependent, exactly like the reference).  GSAGE_ERR_INVALID when the operands do not qualify. */
int gsage_attention_aggregate(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids_dev,
                              int64_t n_parents, int S,
                              const void* w1_dev, int w1_dtype, int64_t ldw, int H, const float* b1_dev, const float* w2_dev,
                              const float* xa_dev, void* out_dev, int out_dtype, int64_t ld_out, void* stream);
/* EXPERIMENTAL (refuses unless the environment has GSAGE_FUSED_LAYER=1): the neighbour half of the mean aggregator in one
 * kernel -- out[p, col0 : col0+O] = act( mean_j table[ids[p*S + j]] . W^T + bias ) (nn_modules.py:197-200) -- the reduced rows
 * never travel through HBM.  bf16 table and W (O x d, O <= 128), S <= 32. */
int gsage_gather_mean_project(const void* table_dev, int dtype, int64_t ld, int64_t n_table_rows, int d, const int64_t* ids_dev,
                              int64_t n_parents, int S, const void* w_dev, int w_dtype, int64_t ldw, int O, const float* bias_dev,
                              int act, void* out_dev, int out_dtype, int64_t ld_out, int64_t col0, void* stream);
/* One time step of the LSTM aggregator's cell (nn_modules.py:266,276-278: nn.LSTM, one layer, unidirectional, batch_first):
 *   gates = gx + gh + b_ih + b_hh   (n x 4H fp32, torch's gate order i, f, g, o; gx = x_t . W_ih^T and gh = h_{t-1} . W_hh^T
 *                                    come from gsage_linear; separate row strides, so gx may be step t of a (n, S, 4H) block
 *                                    projected in one go: gx_dev = block + t*4H, ldgx = S*4H)
 *   c = sigmoid(f) c + sigmoid(i) tanh(g);   h = sigmoid(o) tanh(c)          (c fp32 in place, h fp32 or bf16)
 * `first` != 0: zero initial state -- c is not read and gh is ignored (may be NULL). */
int gsage_lstm_cell(const float* gx_dev, int64_t ldgx, const float* gh_dev, int64_t ldgh, const float* b_ih_dev, const float* b_hh_dev,
                    float* c_dev, void* h_dev, int h_dtype, int64_t ldh, int64_t n, int H, int first, void* stream);
/* gsage_lstm_cell backwards: one step of back-propagation through time (nn.LSTM under loss.backward(), models.py:101).  The gate
 * activations are recomputed from the pre-activations the forward saw (same gx / gh / biases) and c_{t-1}.
 *   in : dh_dev = d loss / d h_t (n x H),  dc_dev = d loss / d c_t (n x H, in place -> d loss / d c_{t-1})
 *   out: dgates_dev (n x 4H, row stride ldg) = d loss / d (gx + gh + b), gate order i, f, g, o.
 * The caller turns dgates into dW_ih / dW_hh (gsage_wgrad), db (gsage_colsum), dx_t and dh_{t-1} (gsage_linear, transposed). */
int gsage_lstm_cell_backward(const float* gx_dev, int64_t ldgx, const float* gh_dev, int64_t ldgh, const float* b_ih_dev,
                             const float* b_hh_dev, const float* c_prev_dev, const float* dh_dev, float* dc_dev, float* dgates_dev,
                             int64_t ldg, int64_t n, int H, int first, void* stream);
/* F.normalize(dim=1, eps=1e-12) (models.py:90), fp32 out */
int gsage_l2_normalize(const void* x_dev, int dtype, int64_t ld, int64_t n, int d, float* out_dev, int64_t ld_out,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Projection.  Replaces the nn.Linear calls of the aggregators and preps
 * (nn_modules.py:150,166,200,224,228,307-308,317; models.py:91).
 *   out[:, col0 : col0+O] = act( A[ids] (n x d) . W^T (O x d, row-major, ldw) + bias )
 * `ids` NULL -> A read in place.  A / W / out dtypes independent (fp32 or bf16); fp32 accumulate.
 * `exact` == 1 forces the fp32 FFMA kernel; 0 lets bf16 operands (or fp32 operands, as TF32) run on the tcgen05
 * tensor-core kernel when they qualify (16-byte aligned rows, O % 16 == 0); 2 = fp32 accuracy on the tensor cores: fp32
 * operands whose weights fit in shared memory twice run as 3 x TF32 (a = a_hi + a_lo, w = w_hi + w_lo, the three leading
 * products accumulated in fp32; ~1e-6 relative), anything else falls back to the FFMA kernel.
 * ------------------------------------------------------------------------------------------------ */
typedef struct gsage_linear_seg {
    const void* a_dev; int a_dtype; int64_t lda; const int64_t* ids_dev;    /* A source (+ optional gather) */
    const void* w_dev; int w_dtype; int64_t ldw; int d; int O;               /* W (O x d)                    */
    const float* bias_dev;                                                   /* O floats or NULL             */
    int64_t col0;                                                            /* output column offset         */
    int reduce_S;   /* <= 1: A row r = A[ids[r]].  S > 1: A row r = mean_j A[ids[r*S + j]] -- the neighbour gather+mean
                       (nn_modules.py:197-198) fused into the projection's operand load; the mean never touches HBM */
    int w_transposed; /* != 0: W is stored (d x O) row-major, i.e. out = A . W (the data-gradient form of a Linear) */
    int64_t a_rows;   /* rows of the table behind a_dev when ids_dev is given (0 = unknown): ids outside [0, a_rows) then read
                         as ZERO rows, like gsage_gather_reduce; with 0 every id must be a valid row */
} gsage_linear_seg;

/* up to two segments writing disjoint column ranges of one output: the "concat-with-self" of
 * nn_modules.py:200 ([fc_x(x) | fc_neib(agg)]) is ONE launch */
int gsage_linear(const gsage_linear_seg* segs, int n_segs, int64_t n, int act, void* out_dev, int out_dtype,
                 int64_t ld_out, int exact, void* stream);

/* Pool aggregators (nn_modules.py:223-226,240,252): out[p, col0+o] = reduce_{j<S} act(A[ids[p*S+j]] . W[o] + bias[o]),
 * reduce = GSAGE_RED_MAX | GSAGE_RED_MEAN.  The per-neighbour MLP runs on the tensor cores and the pool over the S
 * rows of a parent happens in the kernel's epilogue: the (n*S, O) hidden rows never reach HBM.  Requires operands that
 * qualify for the tensor-core kernel (bf16, or fp32 run as TF32; 16-byte aligned rows; O % 16 == 0), S <= 128. */
int gsage_linear_pooled(const gsage_linear_seg* seg, int64_t n_parents, int S, int reduce, int act, void* out_dev,
                        int out_dtype, int64_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Engine.  Replaces GSSupervised.forward (models.py:71-91) for 2-layer stacks: owns the hop buffers, the
 * workspace and the kernel order (hop-0 draws before hop-1 draws).
 * ------------------------------------------------------------------------------------------------ */
typedef struct gsage_engine_config {
    int aggregator;                 /* gsage_aggregator                                                  */
    int prep;                       /* gsage_prep                                                        */
    int n_layers;                   /* must be 2 (train.py:105-118 hard-codes two)                        */
    int fanout[2];                  /* n_samples per hop                                                  */
    int out_dim[2];                 /* O per layer (aggregator output is 2*O)                             */
    int act[2];                     /* gsage_act per layer (train.py:110,116: relu, identity)             */
    int n_classes;
    int compute_dtype;              /* GSAGE_F32: every kernel fp32-exact (projections: 3 x TF32 on the tensor cores where the
                                       weights fit in shared memory, else FFMA; env GSAGE_FP32_FFMA=1 forces FFMA).
                                       GSAGE_BF16: bf16 tables/activations, fp32 accumulate, tensor-core projections */
    /* node features (`feats`, problem.py:118-121); NULL for feats=None (Pokec) */
    const void* feats_dev; int feats_dtype; int64_t feats_ld; int feats_dim; int64_t feats_rows;
    /* NodeEmbeddingPrep (nn_modules.py:126-155): table (n_nodes+1, 64), fc 64x64 + bias; n_nodes = adj.shape[0] */
    const void* emb_dev; int emb_dtype; int64_t emb_ld; int emb_dim; int64_t n_nodes;
    int hidden_dim;                 /* pool MLP width (512) / attention width (32) / LSTM state width (512) */
    int64_t max_batch;              /* workspace is sized for this many seeds                             */
    int allow_tf32;                 /* fp32 mode only: run the projections on the tensor cores as TF32 (10-bit mantissa
                                       products, fp32 accumulate; ~1e-3 relative) instead of the exact FFMA kernel */
} gsage_engine_config;

/* fp32 weights, named after the reference's state_dict keys; unused ones NULL. All row-major (out, in). */
typedef struct gsage_layer_weights {
    const float* fc_x;              /* agg_layers.k.fc_x.weight    (O, d_in)                              */
    const float* fc_neib;           /* agg_layers.k.fc_neib.weight (O, d_in | hidden)                     */
    const float* mlp_w;             /* agg_layers.k.mlp.0.weight   (hidden, d_in)   pool                  */
    const float* mlp_b;             /* agg_layers.k.mlp.0.bias     (hidden)         pool                  */
    const float* att_w1;            /* agg_layers.k.att.0.weight   (hidden, d_in)   attention             */
    const float* att_w2;            /* agg_layers.k.att.2.weight   (hidden, hidden) attention             */
    const float* lstm_w_ih;         /* agg_layers.k.lstm.weight_ih_l0 (4*hidden, d_in)   lstm             */
    const float* lstm_w_hh;         /* agg_layers.k.lstm.weight_hh_l0 (4*hidden, hidden) lstm             */
    const float* lstm_b_ih;         /* agg_layers.k.lstm.bias_ih_l0   (4*hidden)         lstm             */
    const float* lstm_b_hh;         /* agg_layers.k.lstm.bias_hh_l0   (4*hidden)         lstm             */
} gsage_layer_weights;

typedef struct gsage_weights {
    gsage_layer_weights layer[2];
    const float* fc_w;              /* fc.weight (n_classes, 2*O_last)                                    */
    const float* fc_b;              /* fc.bias                                                            */
    const float* prep_fc_w;         /* prep.fc.weight: (64,64) node_embedding | (32, d) linear            */
    const float* prep_fc_b;         /* prep.fc.bias   (node_embedding only)                               */
    int prep_out_dim;               /* LinearPrep output_dim (32)                                         */
} gsage_weights;

int gsage_engine_create(const gsage_engine_config* cfg, gsage_engine** out);
void gsage_engine_destroy(gsage_engine* e);
/* (re)loads weights: converts to the compute dtype / padded layouts the kernels want */
int gsage_engine_set_weights(gsage_engine* e, const gsage_weights* w, void* stream);
/* logits[B, n_classes] (fp32) = fc(normalize(agg2(agg1(...))))  for seeds ids_dev[0:B].
 * Sampling draws from `rng` (device stream), hop 0 then hop 1. */
int gsage_engine_forward(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                         float* logits_dev, void* stream);
/* Seed-sharded forward that stays bit-exact with the single-process run: `ids_dev` holds seeds [first, first+B) of a
 * global batch of `global_B` seeds; the rank consumes the whole global batch's draws from its (identically seeded)
 * stream and samples with its slice of them.  global_B == B, first == 0 is gsage_engine_forward. */
int gsage_engine_forward_sharded(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                                 int64_t global_B, int64_t first, float* logits_dev, void* stream);
/* the same call for host buffers: H2D of the seed ids, forward, D2H of the logits, stream-synchronised.
 * This is the end-to-end entry a reference-side caller binds (ids and logits are what models.py:71 takes/returns). */
int gsage_engine_forward_host(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                              float* logits_host, void* stream);
/* GSSupervised.forward with the DENSE 2-D edgelist sampler (nn_modules.py:19-49; `uniform_neighbor_sampler` is the default
 * of train.py:55): hop k takes adj[ids][:, perm_k][:, :S_k], one permutation shared by all rows of the hop.  perm0 / perm1
 * are the reference's two torch.randperm(K) draws (CPU generator, hop 0 first), made by the caller and passed as K int64
 * each; `adj_dev` is the (n_rows, K) int64 table.  Everything after the sampling is gsage_engine_forward. */
int gsage_engine_forward_dense(gsage_engine* e, const int64_t* adj_dev, int64_t n_rows, int K, const int64_t* perm0_dev,
                               const int64_t* perm1_dev, const int64_t* ids_dev, int64_t B, float* logits_dev, void* stream);
/* Sample-ahead.  Draws and samples both hops of a batch NOW, on the engine's own high-priority stream, into a spare id
 * buffer; the next gsage_engine_forward[_sharded|_host] call -- which must name the same batch: same ids pointer, B,
 * (global_B, first), graph and rng -- skips its sampling section and only waits for that stream.  The (latency-bound)
 * sampling of batch i+1 then runs underneath the HBM-bound aggregation of batch i.  The draw order on the RNG stream is
 * the call order, exactly as if the sampling had happened inside the forwards (models.py:78-79), so the sampled ids are
 * bit-identical with and without it.  Call it right AFTER queueing the forward (and backward) it should overlap with:
 * the sampler stream only waits for the batch BEFORE that one (the last reader of the spare id buffer).  The ids are
 * copied when the call executes on the device: keep `ids_dev` / `ids_host` (pinned) alive until the matching forward.
 * Other state the forward reads (weights, tables) is not touched. */
int gsage_engine_sample_ahead(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                              int64_t global_B, int64_t first, void* stream);
int gsage_engine_sample_ahead_host(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                                   void* stream);
/* gsage_engine_forward_host for a caller that knows its next batch (problem.py:141-153 does): before the call blocks on its
 * own logits it queues the NEXT batch's H2D copy of the ids, its sampling (sampler stream) and its whole forward (`stream`),
 * and the logits of THIS batch travel on a copy stream -- the GPU never idles across the host round trip.  The next call
 * must then name that batch (same pointer and size); it finds its forward already queued and only collects the result.
 * Draw order and results are those of the unpipelined calls.  `next_ids_host` NULL = plain gsage_engine_forward_host.
 * While a forward is queued the engine's other entry points refuse to run, and gsage_engine_peek shows the queued batch. */
int gsage_engine_forward_host_next(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                                   const int64_t* next_ids_host, int64_t next_B, float* logits_host, void* stream);
/* 1 while a sampled-ahead batch waits for its forward */
int gsage_engine_sample_ahead_pending(const gsage_engine* e);
/* Ordering of gsage_engine_sample_ahead against the producer of `ids_dev` (e.g. the index kernel of problem.iterate,
 * problem.py:149-152, or an H2D copy on `stream`): the sampler stream waits for an event on the caller's stream before it
 * copies the ids.  By default that event is recorded when gsage_engine_sample_ahead is called -- always correct, but it also
 * covers the forward queued just before, so nothing overlaps.  A caller that already holds the next batch's ids when it
 * queues the CURRENT forward calls gsage_engine_inputs_ready first (ids complete on `stream` -> forward -> sample_ahead):
 * the event is then the one recorded here and the sampling overlaps that forward.  One-shot: consumed by the next
 * gsage_engine_sample_ahead. */
int gsage_engine_inputs_ready(gsage_engine* e, void* stream);
/* Non-blocking look at the sticky error flags (they live in mapped host memory) of the graph / rng of the last forward:
 * GSAGE_ERR_INDEX if an earlier forward saw an id outside the adjacency (where `feats[ids]`, models.py:76, raises
 * IndexError; cleared on report), GSAGE_ERR_RNG if the rng's look-ahead window ran short.  No synchronisation: it reports
 * what has REACHED the host, i.e. typically the forward before the one in flight. */
int gsage_engine_poll_errors(gsage_engine* e);
/* device views of the last forward's intermediates (valid until the next forward): hop ids and layer outputs.
 * what: 0 ids0, 1 ids1, 2 ids2 (int64) ; 10 layer-1 output (26B x 2*O1) ; 11 layer-2 output (B x 2*O2) */
int gsage_engine_peek(gsage_engine* e, int what, const void** ptr_dev, int64_t* rows, int64_t* cols, int64_t* ld,
                      int* dtype);
int64_t gsage_engine_workspace_bytes(const gsage_engine* e);

/* ------------------------------------------------------------------------------------------------
 * Backward (mean aggregator + identity prep).  Replaces `loss.backward()` (models.py:101) for the parameters of
 * the hot path; must follow a gsage_engine_forward of the same batch (reads its ids and activations).
 * `dlogits_dev`: d loss / d logits (B x n_classes, fp32).  `grads`: fp32 device buffers laid out like the
 * weights (same shapes, contiguous), overwritten.  The two halves let the caller start the gradient
 * all-reduce of the head (fc + layer 2, tiny) while the big layer-1 weight gradients are still being computed:
 *   _head    fc.weight, fc.bias, layer-2 fc_x / fc_neib, and the layer-1 output gradient (kept internally)
 *   _layer1  layer-1 fc_x / fc_neib (reduction over all 26*B parent rows)
 * ------------------------------------------------------------------------------------------------ */
typedef struct gsage_grads {
    float* fc_x[2];                 /* agg_layers.k.fc_x.weight.grad    */
    float* fc_neib[2];              /* agg_layers.k.fc_neib.weight.grad */
    float* fc_w;                    /* fc.weight.grad                   */
    float* fc_b;                    /* fc.bias.grad                     */
} gsage_grads;
int gsage_engine_backward_head(gsage_engine* e, const float* dlogits_dev, const gsage_grads* grads, void* stream);
int gsage_engine_backward_layer1(gsage_engine* e, const gsage_grads* grads, void* stream);

/* The weight gradient of one projection on its own: dW (O x d, fp32, overwritten) = G^T . A[ids] (ids NULL: A in place),
 * G = the (n, O) output gradient.  exact != 0: fp32 FFMA kernel, fp32 G.  exact == 0: split-K tcgen05 kernel reading both
 * row-major operands as MN-major tiles (bf16 G and A, O == 128, 16-byte aligned rows; GSAGE_ERR_INVALID otherwise). */
int gsage_wgrad(const void* g_dev, int g_dtype, int64_t ldg, int O, const void* a_dev, int a_dtype, int64_t lda, int64_t n_table_rows,
                const int64_t* ids_dev, int d, int64_t n, float* dw_dev, int64_t lddw, int exact, void* stream);

/* Layer-1 gradients of the Pokec recipe (utils/pokec.sh: mean aggregator + NodeEmbeddingPrep without features, fp32).
 * Replaces gsage_engine_backward_layer1 for such models; same preconditions.  The library does every reduction over rows;
 * the caller finishes with four (O x 64)(64 x 64) products (formulas in engine.cu, done by model.GSSupervised.backward):
 *   gx_raw (O1, emb_dim) = Gx^T . E[self ids]        gn_raw (O1, emb_dim) = Gn^T . mean_j E[neighbour ids]
 *   csum   (2 * O1)      = column sums of G           d_table (n_nodes + 1, emb_dim) = dense gradient of the embedding table
 * where G = d loss / d (layer-1 pre-activation) of all 26*B parent rows.  All buffers fp32, overwritten. */
typedef struct gsage_embedding_grads { float* gx_raw; float* gn_raw; float* csum; float* d_table; } gsage_embedding_grads;
int gsage_engine_backward_layer1_embedding(gsage_engine* e, const gsage_embedding_grads* g, void* stream);

/* Layer-1 gradients of a mean-aggregator model behind LinearPrep (nn_modules.py:158-166: X = F Wp^T, no bias).  The prep is
 * linear, so layer 1 = act([F_self (Wx Wp)^T | mean_j F_nb (Wn Wp)^T]) and every gradient follows from two reductions against
 * the RAW feature rows:
 *   gx_raw (O1, feats_dim) = Gx^T . F[self ids]          gn_raw (O1, feats_dim) = Gn^T . mean_j F[neighbour ids]
 * (G as above; the neighbour means of the raw rows are gathered here -- the forward only ever reduced the 32-wide X rows).
 * The caller finishes with three small products: dWx = gx_raw.Wp^T, dWn = gn_raw.Wp^T, dWp = Wx^T.gx_raw + Wn^T.gn_raw.
 * Call after gsage_engine_backward_head.  Buffers fp32, overwritten. */
typedef struct gsage_linear_prep_grads { float* gx_raw; float* gn_raw; } gsage_linear_prep_grads;
int gsage_engine_backward_layer1_linear(gsage_engine* e, const gsage_linear_prep_grads* g, void* stream);

/* Every parameter gradient of a max / mean pool model (bf16 compute, identity prep, output_dim 128) in one call; replaces
 * the _head / _layer1 pair for such models (nn_modules.py:207-256 through loss.backward(), models.py:101).  `pg`: gradients
 * of agg_layers.k.mlp.0.weight (hidden, d_in) / .bias (hidden).  All buffers fp32, overwritten. */
typedef struct gsage_pool_grads { float* mlp_w[2]; float* mlp_b[2]; } gsage_pool_grads;
int gsage_engine_backward_pool(gsage_engine* e, const float* dlogits_dev, const gsage_grads* grads, const gsage_pool_grads* pool_grads,
                               void* stream);
/* The same for BASELINE config C3 (Pokec: pool aggregator + NodeEmbeddingPrep without features, bf16).  The prep's affine is
 * folded into layer 1 (W' = W.Wp), so grads->fc_x[0] (O1, emb_dim) and pool_grads->mlp_w[0] (hidden, emb_dim) receive the RAW
 * reductions against the embedding rows; with csum_x (O1: column sums of Gx) and mlp_b[0] the caller unfolds them
 * (dW = raw.Wp^T + csum (x) bp, dWp = sum W^T.raw, dbp = sum W^T.csum).  d_table: dense (n_nodes + 1, emb_dim) fp32 gradient. */
typedef struct gsage_pool_embedding_grads { float* csum_x; float* d_table; } gsage_pool_embedding_grads;
int gsage_engine_backward_pool_embedding(gsage_engine* e, const float* dlogits_dev, const gsage_grads* grads,
                                         const gsage_pool_grads* pool_grads, const gsage_pool_embedding_grads* emb_grads, void* stream);

/* torch.nn.utils.clip_grad_norm(params, max_norm) + torch.optim.Adam.step() (models.py:102-103) on ONE flat fp32 parameter
 * buffer and its flat gradient / moment buffers (two launches instead of ~25).  `step` is the 1-based step count of the bias
 * corrections; max_norm <= 0 skips the clip; `scratch_dev` is one float of device memory. */
int gsage_adam_step(float* param_dev, float* grad_dev, float* m_dev, float* v_dev, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int64_t step, float max_norm, float* scratch_dev, void* stream);

/* Every parameter gradient of an attention model (bf16 compute, identity prep, output_dim 128, feature width % 16 == 0) in
 * one call (nn_modules.py:289-321 through loss.backward(), models.py:101).  `ag`: gradients of agg_layers.k.att.0.weight
 * (32, d_in) and agg_layers.k.att.2.weight (32, 32).  All buffers fp32, overwritten. */
typedef struct gsage_attention_grads { float* att_w1[2]; float* att_w2[2]; } gsage_attention_grads;
int gsage_engine_backward_attention(gsage_engine* e, const float* dlogits_dev, const gsage_grads* grads,
                                    const gsage_attention_grads* att_grads, void* stream);

/* ---- gradients behind the NARROW plug-in API ------------------------------------------------------------------------
 * The reference back-propagates through its plug-ins with torch autograd (models.py:100-101 -> nn_modules.py:196-204,
 * 223-232, 305-321, 144-155).  operators.py wraps the narrow calls (`agg(x, neibs)`, `prep(ids, feats, layer_idx)`) in
 * torch.autograd.Functions; their backward passes are gsage_wgrad (dW = G^T A), gsage_linear with a transposed weight
 * (dA = G W) and the elementwise / segment kernels below.  Everything fp32, rows in place, buffers overwritten.
 *   act_backward           dpre = dout * act'(out)  (out = the post-activation value)
 *   segment_broadcast      dst[p*S + j, :] = scale * src[p, :]            (mean over S rows, backwards: scale = 1/S)
 *   segment_max_backward   the first row attaining the maximum over the S rows of a parent receives dpooled (torch.max(dim))
 *   attention_sum_backward m_p = sum_j w_j n_j, w = softmax_j <na_j, xa_p>: dn = w_j dM_p (direct path only), dna_j = ds_j xa_p,
 *                          dxa_p = sum_j ds_j na_j with ds = softmax'(dw), dw_j = <dM_p, n_j>;  scratch: n*S floats
 *   colsum                 out[c] = sum_r x[r, c]   (x contiguous)        (bias gradients)
 *   embedding_backward     table_grad[ids[i], :] += drows[i, :]           (nn.Embedding's dense gradient; caller zeroes it)
 *   l2_normalize_backward  F.normalize(z, dim=1) backwards (models.py:90; z, dzn, dz contiguous (n, d)) */
int gsage_act_backward(const float* dout_dev, int64_t ld_dout, const float* out_dev, int64_t ld_out, int64_t n, int width, int act,
                       float* dpre_dev, int64_t ld_dpre, void* stream);
int gsage_segment_broadcast(const float* src_dev, int64_t ld_src, int64_t n, int d, int S, float scale, float* dst_dev, int64_t ld_dst,
                            void* stream);
int gsage_segment_max_backward(const float* h_dev, int64_t ld_h, const float* dpooled_dev, int64_t ld_dp, int64_t n, int S, int H,
                               float* dh_dev, int64_t ld_dh, void* stream);
int gsage_attention_sum_backward(const float* neibs_dev, int64_t ld_nb, int d, int64_t n, int S, const float* dm_dev, int64_t ld_dm,
                                 const float* w_dev, const float* na_dev, const float* xa_dev, int H, float* dn_dev, int64_t ld_dn,
                                 float* dna_dev, float* dxa_dev, float* scratch_dev, void* stream);
int gsage_colsum(const float* x_dev, int64_t n, int d, float* out_dev, void* stream);
int gsage_embedding_backward(const float* drows_dev, int64_t ld, int d, const int64_t* ids_dev, int64_t n_ids, float* table_grad_dev,
                             int64_t ld_table, int64_t table_rows, void* stream);
int gsage_l2_normalize_backward(const float* z_dev, const float* dzn_dev, int64_t n, int d, float* dz_dev, void* stream);

/* ---- the gradient all-reduce over NVLink peer memory (the path's one collective) --------------------------------
 * One-shot sum all-reduce for small buffers (the 0.6-0.9 MB gradient bucket): every rank's bucket lives in symmetric memory
 * (one allocation mapped into every process of the box; `peer_ptrs_host[p]` = rank p's bucket as seen from THIS process), and
 * one kernel per rank reads all `world` buckets directly over NVLink, writes  out[i] = sum_p scale_p * bucket_p[i]  into the
 * rank's own private buffer and accumulates sum(out^2) (the norm of clip_grad_norm, models.py:102) into `sumsq_dev` (may be NULL).
 * Every rank calls it once per step with the same `epoch` (1, 2, ...); flags inside the symmetric buffers order the kernels
 * of the ranks against each other (no host synchronisation, no NCCL).  The symmetric buffer of a rank must hold
 * gsage_peer_allreduce_words(n) fp32 words, zero-initialised; n a multiple of 4. */
int64_t gsage_peer_allreduce_words(int64_t n);
int gsage_peer_allreduce(const uint64_t* peer_ptrs_host, int world, int rank, int64_t n, uint32_t epoch, float scale, float* out_dev,
                         float* sumsq_dev, void* stream);

/* ---- per-batch training metric on the device (train.py:150 `problem.metric_fn(to_numpy(targets), to_numpy(preds))`) -------
 * gsage_metric_f1: sklearn micro / macro F1 as problem.py:44-58 computes them.  multilabel == 0 (`classification`):
 *   `targets_dev` int64 (n), prediction = argmax of the n_classes logits (first maximum), macro averages over the labels
 *   that occur in targets or predictions.  multilabel != 0: `targets_dev` float32 (n, ld_targets) of 0 / 1, prediction =
 *   logit > 0, macro averages over all n_classes columns (empty label: F1 0).  `scratch_dev`: 3 * n_classes * 8 bytes.
 *   out_dev[0] = micro, out_dev[1] = macro (double).
 * gsage_metric_mae: mean |preds - targets| over n float32 elements (problem.py:60-64) -> out_dev[0] (double). */
int gsage_metric_f1(const float* preds_dev, int64_t ld, const void* targets_dev, int64_t ld_targets, int64_t n, int n_classes,
                    int multilabel, void* scratch_dev, double* out_dev, void* stream);
int gsage_metric_mae(const float* preds_dev, const float* targets_dev, int64_t n, double* out_dev, void* stream);

/* keep != 0: the next forwards keep every activation the backward pass needs (training).  0 (default): forward-only
 * streaming -- intermediates may be processed in L2-sized chunks that reuse their buffers. */
int gsage_engine_keep_activations(gsage_engine* e, int keep);

/* Live stopwatch (CUDA events on the launching stream) around kernel groups of the forward, for bench.py:
 *   FORWARD  the whole gsage_engine_forward        SAMPLE   rng draws + sample kernels, both hops (on the stream they run on)
 *   REDUCE   the DOMINANT aggregate launch: layer 1 on the (x1, x2) pair -- the fused gather+mean kernel, the pooled MLP
 *            kernel (pool aggregators) or the fused attention kernel
 *   PROJECT  the concat-with-self projection launch that follows it
 *   APP0     layer 1 on the (x0, x1) pair, whole        LAYER2   layer 2, whole        HEAD   normalise + classifier
 *   WAIT     the main stream waiting for the sampled-ahead batch (0 when the sampler stream finished first)
 * `bytes_out` / `flops_out`: algorithmic bytes and flops of the recorded launches.  Arrays of
 * GSAGE_PROF_CATS entries.  Reading synchronises. */
enum { GSAGE_PROF_FORWARD = 0, GSAGE_PROF_SAMPLE = 1, GSAGE_PROF_REDUCE = 2, GSAGE_PROF_PROJECT = 3, GSAGE_PROF_APP0 = 4,
       GSAGE_PROF_LAYER2 = 5, GSAGE_PROF_HEAD = 6, GSAGE_PROF_WAIT = 7, GSAGE_PROF_CATS = 8 };
int gsage_engine_profile(gsage_engine* e, int enable);
int gsage_engine_profile_read(gsage_engine* e, double* ms_out, int64_t* launches_out, double* bytes_out, double* flops_out,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSAGE_B200_H */
