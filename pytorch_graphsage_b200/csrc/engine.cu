// engine.cu -- GSSupervised.forward re-expressed over ids (the "wide" drop-in entry).
//
// Replaces  GSSupervised.forward  (/root/reference/models.py:71-91): sample both hops, prep every hop,
// apply aggregator layer 1 to the pairs (x0,x1) and (x1,x2), layer 2 to the result pair, L2-normalise, fc.
//
// What changes versus the reference's dataflow (results identical up to fp32 rounding):
//   * ids never leave the device: hop-0 draws, then hop-1 draws, from the device MT19937 stream (same order
//     as the reference consumes np.random); the three hop id arrays live back to back in one
//     buffer so "all parents of layer 1" is a prefix of it;
//   * with the identity prep the neighbour rows are never materialised: every consumer (reduce kernel,
//     projection kernels) takes (table, ids) and gathers on the fly;
//   * `torch.cat([fc_x(x), fc_neib(agg)])` is one launch writing two column ranges.
#include "graph.cuh"
#include "mt19937.cuh"
#include "linear.cuh"
#include "backward.cuh"

#include <vector>
#include <algorithm>
#include <stdlib.h>

namespace gsage {
int sample_sparse_launch(gsage_graph* g, const int64_t* ids, int64_t n, int S, const uint32_t* sel, int64_t* out,
                         cudaStream_t s);
int gather_reduce_launch(const void* table, int dtype, int64_t ld, int64_t rows, int d, const int64_t* ids,
                         int64_t n_parents, int S, int reduce, const float* weights, void* out, int out_dtype,
                         int64_t ld_out, cudaStream_t s, int l2_hint = 0);
bool head_fused_eligible(int d, int n_classes);
int head_fused_launch(const float* z, int64_t ld, int64_t n, int d, const float* w, const float* b, int n_classes, float* zn, int64_t ld_zn,
                      float* logits, int64_t ld_logits, cudaStream_t s);

// a set of rows some kernel will read: either (table, ids) gathered on the fly or a materialised buffer
struct RowSrc {
    const void* base; int dtype; int64_t ld; int64_t table_rows; const int64_t* ids; int d;
    RowSrc shifted(int64_t rows) const {       // the same source starting `rows` further down
        RowSrc r = *this;
        if (ids) r.ids = ids + rows;
        else { r.base = (const char*)base + rows * ld * (int64_t)dtype_size(dtype); r.table_rows = table_rows - rows; }
        return r;
    }
};

// a weight matrix as a kernel will read it: the caller's fp32 tensor, or the engine's padded bf16 copy
struct WRef { const void* p; int dtype; int64_t ld; const float* hi = nullptr; const float* lo = nullptr; };   // hi/lo: tf32-exact split (fp32-exact mode)

// fp32 (rows, cols) -> bf16 (rows, ld) with zero padding (the tcgen05 kernel wants 16-byte aligned bf16 rows)
__global__ void __launch_bounds__(256) to_bf16_padded_kernel(const float* __restrict__ src, int rows, int cols,
                                                             __nv_bfloat16* __restrict__ dst, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * ld) return;
    const int r = (int)(i / ld), c = (int)(i % ld);
    dst[i] = __float2bfloat16_rn(c < cols ? src[(int64_t)r * cols + c] : 0.0f);
}

// fp32 (rows, cols) -> bf16 TRANSPOSED (cols, ld) with zero padding: the K-major operand of a data-gradient projection
// (dX = dY . W is `out = A . W'^T` with W' = W^T)
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const float* __restrict__ src, int rows, int cols,
                                                                __nv_bfloat16* __restrict__ dst, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)cols * ld) return;
    const int c = (int)(i / ld), r = (int)(i % ld);
    dst[i] = __float2bfloat16_rn(r < rows ? src[(int64_t)r * cols + c] : 0.0f);
}

__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * cols) dst[(i % cols) * rows + i / cols] = src[i];
}

__global__ void __launch_bounds__(256) axpy_kernel(float* y, const float* x, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

__global__ void __launch_bounds__(256) fill_i64_kernel(int64_t* p, int64_t n, int64_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace gsage

using namespace gsage;

// CUDA-event stopwatch around kernel groups of one forward (bench.py's live roofline / breakdown numbers)
struct EngineProf {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    struct Rec { int cat; int a, b; };
    std::vector<Rec> recs;
    int used = 0;
    double bytes[GSAGE_PROF_CATS] = {0}; double flops[GSAGE_PROF_CATS] = {0};
    void work(int rec, int cat, double b, double f) { if (rec >= 0) { bytes[cat] += b; flops[cat] += f; } }
    int begin(int cat, cudaStream_t s) {
        if (!on) return -1;
        if (used + 2 > (int)pool.size()) {
            if (pool.size() >= 16384) return -1;
            for (int i = 0; i < 64; ++i) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        }
        recs.push_back(Rec{cat, used, used + 1});
        cudaEventRecord(pool[used], s);
        used += 2;
        return (int)recs.size() - 1;
    }
    void end(int rec, cudaStream_t s) { if (rec >= 0) cudaEventRecord(pool[recs[rec].b], s); }
};

struct gsage_engine {
    EngineProf prof;
    gsage_engine_config cfg;
    gsage_weights w;
    bool have_weights = false;
    int T = GSAGE_F32;                 // compute dtype of tables / activations
    int d_prep = 0, ld_prep = 0;       // row width after prep (+ padded ld)
    int hid = 0;
    int64_t maxB = 0, n0 = 0, n1 = 0, n2 = 0;   // capacities in rows per hop
    int64_t B = 0;                      // batch of the last forward
    char* ws = nullptr; int64_t ws_bytes = 0;
    // carved views
    int64_t* ids = nullptr; uint32_t* sel = nullptr; int64_t* look0 = nullptr;
    // sample-ahead (gsage_engine_sample_ahead): the hop ids of the NEXT batch are drawn on the engine's own high-priority
    // stream into the spare id buffer while the current batch aggregates; `ids` always points at the slot in use
    int64_t* ids_slot[2] = {nullptr, nullptr}; int cur = 0; uint32_t* sel_ahead = nullptr;
    cudaStream_t ss = nullptr; cudaEvent_t ev_ahead = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}; bool done_valid[2] = {false, false};   // last reader of each id slot (forward / backward) finished
    // mean aggregator: the sample-ahead stream starts AFTER the dominant (HBM-bound) gather+mean launch of the forward in
    // flight, so its kernels share the SMs with the projection / layer-2 tail instead of slowing the bandwidth-bound kernel
    cudaEvent_t ev_mid = nullptr; bool mid_valid = false; int ahead_after_gather = 1; int ahead_split = 0;   // split: the RNG draws of the next batch run ahead of the gate, i.e. under the dominant launch (measured twice: off)
    struct Ahead { bool valid = false; const void* src = nullptr; int64_t B = 0, global_B = 0, first = 0;
                   gsage_graph* g = nullptr; gsage_rng* rng = nullptr; } ahead;
    // the producer of the NEXT batch's ids on the caller's stream: recorded by gsage_engine_inputs_ready (before the forward in
    // flight was queued) or, failing that, at the entry of gsage_engine_sample_ahead; the sampler stream waits for it before it
    // reads the ids
    cudaEvent_t ev_src = nullptr; bool src_marked = false;
    // whose sticky error flags gsage_engine_poll_errors looks at (the graph / rng of the last forward)
    gsage_graph* last_g = nullptr; gsage_rng* last_rng = nullptr;
    void* X = nullptr;                  // materialised prepped rows (non-identity preps)
    void* M = nullptr;                  // reduced neighbour rows: [layer-1 app 1 (n0) | layer-1 app 2 (n1) | layer 2 (n0)] x ld_m
    int64_t ld_m = 0;                   // (kept after the forward: the backward pass reads them)
    void* HN = nullptr; void* Pp = nullptr;          // pool: hidden rows / pooled rows
    void* T1 = nullptr; void* NA = nullptr; void* T1x = nullptr; void* XA = nullptr; float* AW = nullptr;   // attention
    void* H1 = nullptr;                 // layer-1 output (n0 + n1, 2*O1)
    float* Z = nullptr; float* ZN = nullptr; float* LG = nullptr;
    // host-buffer entry, pipelined (gsage_engine_forward_host_next with a next batch): two logits buffers, a copy stream for
    // the D2H of batch i while the forward of batch i+1 (queued before the call blocks) already runs
    float* LG2[2] = {nullptr, nullptr}; int lg_cur = 0;
    cudaStream_t cs = nullptr; cudaEvent_t ev_fwd = nullptr, ev_copy = nullptr;
    struct Pre { bool valid = false; const void* src = nullptr; int64_t B = 0; int buf = 0; gsage_graph* g = nullptr; gsage_rng* rng = nullptr; } pre;
    int64_t ld_h1 = 0;
    // weights as the kernels read them (fp32 originals, or bf16 copies when compute dtype is bf16)
    char* wb = nullptr; int64_t wb_bytes = 0;
    WRef w_x[2], w_n[2], w_mlp[2], w_att1[2], w_ih[2], w_hh[2];
    void* MF = nullptr;                 // mean + LinearPrep backward: neighbour means of the RAW feature rows, (n0 + n1) x feats_ld
    // LSTM aggregator: input half of the gates for all S steps of a block of parents (lgx_rows rows x 4H), recurrent half of the
    // current step, cell / hidden state
    float* LGX = nullptr; float* LGH = nullptr; float* LC = nullptr; void* LH = nullptr; int64_t lgx_rows = 0;
    float* wsplit = nullptr; int64_t wsplit_floats = 0;      // fp32-exact mode: (hi, lo) tf32 halves of fc_x / fc_neib / mlp.0 / att.0 for the 3 x TF32 projections
    WRef w_nT[2], w_mlpT[2];            // pool backward (bf16): fc_neib^T (H x O) and mlp.0.weight^T (d_in x H), K-major
    WRef w_xT0;                         // pool + folded node_embedding backward (bf16): (Wx.Wp)^T (emb_dim x O1), K-major
    WRef w_x2T, w_n2T;                  // mean backward (bf16): layer-2 fc_x^T / fc_neib^T (2*O1 x O2), K-major, for the head's data gradients
    void* DZB = nullptr;                // bf16 copy of d loss / d z (operand of the tensor-core head gradients)
    // attention backward (bf16, identity prep): d loss / d aggregated rows, d softmax weights, d a(n), d tanh input, d a(x), and
    // the four partial input gradients of layer 2 (self / neighbour rows x direct / through-the-attention-MLP)
    float* ADM = nullptr; float* ADW = nullptr; float* ADA = nullptr; float* ADT1 = nullptr; float* ADXA = nullptr; float* ADT1X = nullptr;
    float* ASA = nullptr; float* ASB = nullptr; float* ANA = nullptr; float* ANB = nullptr;
    float* FCP = nullptr; int fc_pad = 0;          // classifier weights + bias zero-padded to a multiple of 16 rows (fp32): the bf16 engine
                                                   // runs fc as a TF32 projection on the tensor cores and stores only n_classes columns
    float* AW2T[2] = {nullptr, nullptr};           // att.2.weight^T per layer (fp32, 32 x 32): d t1 = dA . W2 as a plain projection
    void* ADPB = nullptr; float* AW1S = nullptr;   // d tanh input as bf16 rows of 128 (zero padded) + a (128, d) scratch for its weight gradient
    float* DP = nullptr;                // pool backward: d loss / d pooled rows, (n0 + n1) x H fp32
    void* DHID = nullptr;               // pool backward: d loss / d hidden rows, (n1 + n2) x H bf16
    float* DN2 = nullptr;               // pool backward: d loss / d (layer-2 neighbour rows), n1 x 2*O1 fp32
    // backward scratch (fp32): d zn / d z (B x 2*O2), d h0 / d m2 (B x 2*O1), d H (26B x 2*O1)
    float* DZN = nullptr; float* DZ = nullptr; float* DH0 = nullptr; float* DM2 = nullptr; float* DH = nullptr;
    // NodeEmbeddingPrep without feats is an affine map of the gathered embedding row; every consumer of layer 1 is
    // linear in its input, so the affine is FOLDED into the layer-1 weights (W' = W.Wp, b' = W.bp [+ b]) at
    // set_weights time: the aggregators then read the raw (n_nodes+1, 64) table by id and the prepped rows are
    // never materialised (algebraically identical; fp32 rounding differs at the 1e-6 level)
    bool fold_prep = false;
    float* fold = nullptr; int64_t fold_floats = 0;
    const float* fold_wx = nullptr; const float* fold_wn = nullptr;     // W'x = Wx.Wp, W'n = Wn.Wp (O1 x emb_dim, fp32): backward needs them
    float* DXE = nullptr;               // backward scratch (node_embedding): d loss / d (raw embedding row), (n0 + n1) x emb_dim, fp32
    const float* b_x[2] = {nullptr, nullptr}; const float* b_n[2] = {nullptr, nullptr};
    const float* b_mlp[2] = {nullptr, nullptr}; const float* b_att[2] = {nullptr, nullptr};
    // forward-only streaming (keep_activations == false): the big layer-1 application runs in chunks of `chunk_parents`
    // parents that reuse ONE small reduced-row buffer, so the rows the reduce kernel writes are still in L2 (126 MB) when
    // the projection kernel reads them, and are overwritten there before they are ever written back to HBM
    bool keep_activations = false;
    int64_t chunk_parents = 0; int l2_hint = 0;
    bool fuse_mean = false;   // GSAGE_FUSE_MEAN=1: route the mean aggregator through the one-kernel fused layer (experimental)
};

static int64_t pad_to(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

static WRef f32w(const float* p, int64_t ld) { return WRef{p, GSAGE_F32, ld, nullptr, nullptr}; }

static int linear_call(const RowSrc& a, const WRef& W, int O, const float* bias, int64_t n, int act,
                       void* out, int out_dtype, int64_t ld_out, int64_t col0, int exact, cudaStream_t s) {
    LinearParams P;
    P.n_segs = 1; P.n = n; P.act = act; P.out = out; P.out_dtype = out_dtype; P.ld_out = ld_out;
    P.seg[0] = LinearSeg{a.base, a.dtype, a.ld, a.ids, W.p, W.dtype, W.ld, a.d, O, bias, col0};
    P.seg[0].a_rows = a.ids ? a.table_rows : 0;
    P.seg[0].w_hi = W.hi; P.seg[0].w_lo = W.lo;
    if (n == 0) return GSAGE_OK;
    return linear_dispatch(P, exact, s);
}

static int combine_call(const RowSrc& x, const WRef& Wx, const RowSrc& m, const WRef& Wn, int O, int64_t n, int act,
                        void* out, int out_dtype, int64_t ld_out, int exact, cudaStream_t s, const float* bx = nullptr,
                        const float* bn = nullptr) {
    LinearParams P;
    P.n_segs = 2; P.n = n; P.act = act; P.out = out; P.out_dtype = out_dtype; P.ld_out = ld_out;
    P.seg[0] = LinearSeg{x.base, x.dtype, x.ld, x.ids, Wx.p, Wx.dtype, Wx.ld, x.d, O, bx, 0};
    P.seg[1] = LinearSeg{m.base, m.dtype, m.ld, m.ids, Wn.p, Wn.dtype, Wn.ld, m.d, O, bn, (int64_t)O};
    P.seg[0].a_rows = x.ids ? x.table_rows : 0;
    P.seg[1].a_rows = m.ids ? m.table_rows : 0;
    P.seg[0].w_hi = Wx.hi; P.seg[0].w_lo = Wx.lo;
    P.seg[1].w_hi = Wn.hi; P.seg[1].w_lo = Wn.lo;
    if (n == 0) return GSAGE_OK;
    return linear_dispatch(P, exact, s);
}

static int linear_trans_call(const float* a, int64_t lda, int d, const float* W, int64_t ldw, int O, int64_t n, float* out,
                             int64_t ld_out, cudaStream_t s) {
    LinearParams P;                                  // out (n x O) = a (n x d) . W (d x O): the data gradient of a Linear
    P.n_segs = 1; P.n = n; P.act = GSAGE_ACT_NONE; P.out = out; P.out_dtype = GSAGE_F32; P.ld_out = ld_out;
    P.seg[0] = LinearSeg{a, GSAGE_F32, lda, nullptr, W, GSAGE_F32, ldw, d, O, nullptr, 0};
    P.seg[0].w_trans = 1;
    if (n == 0) return GSAGE_OK;
    return linear_simt_launch(P, s);
}

// one aggregator application: out[n, 2*O] = act([fc_x(x) | fc_neib(reduce(nb))])   (nn_modules.py:196-204 etc.)
static int apply_aggregator(gsage_engine* e, int layer, const RowSrc& x, const RowSrc& nb, int64_t n, int S, void* out,
                            int out_dtype, int64_t ld_out, int64_t m_row0, cudaStream_t s) {
    const gsage_layer_weights& L = e->w.layer[layer];
    const int O = e->cfg.out_dim[layer], act = e->cfg.act[layer], T = e->T;
    const int exact = (T == GSAGE_F32 && !e->cfg.allow_tf32);
    const int d = x.d;
    const int64_t ldm = e->ld_m;
    void* const Mb = (char*)e->M + m_row0 * ldm * (int64_t)dtype_size(T);
    switch (e->cfg.aggregator) {
    case GSAGE_AGG_MEAN: {
        if (T == GSAGE_BF16 && e->w_x[layer].dtype == GSAGE_BF16 && e->fuse_mean) {
            // one kernel: gather self rows + gather-and-mean the S neighbour rows in the loader warps, both
            // projections on tcgen05, activation in the epilogue -- the aggregated rows never exist in HBM
            LinearParams P;
            P.n_segs = 2; P.n = n; P.act = act; P.out = out; P.out_dtype = out_dtype; P.ld_out = ld_out;
            P.seg[0] = LinearSeg{x.base, x.dtype, x.ld, x.ids, e->w_x[layer].p, GSAGE_BF16, e->w_x[layer].ld, d, O, nullptr, 0};
            P.seg[1] = LinearSeg{nb.base, nb.dtype, nb.ld, nb.ids, e->w_n[layer].p, GSAGE_BF16, e->w_n[layer].ld, d, O, nullptr, (int64_t)O};
            P.seg[1].S = S;
            if (linear_umma_eligible(P)) {
                const int p_f = e->prof.begin(GSAGE_PROF_REDUCE, s);
                const int st = linear_umma_launch(P, s);
                e->prof.end(p_f, s);
                {
                    const double es_in = (double)dtype_size(nb.dtype), es_out = (double)dtype_size(out_dtype);
                    e->prof.work(p_f, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * es_in + (nb.ids ? 8.0 * S : 0.0) + d * es_in +
                                                                      (x.ids ? 8.0 : 0.0) + 2.0 * O * es_out), 4.0 * (double)n * d * O);
                }
                return st;
            }
        }
        // the stopwatch covers the dominant launch only (layer 1 on the (x1, x2) pair: n = B*S1 parents)
        const bool dominant = layer == 0 && n > e->B;
        if (T == GSAGE_BF16 && !e->keep_activations && e->w_n[layer].dtype == GSAGE_BF16 && out_dtype == GSAGE_BF16 &&
            gather_mean_project_eligible(nb.base, nb.dtype, nb.ld, d, S, e->w_n[layer].p, GSAGE_BF16, e->w_n[layer].ld, O)) {
            // forward only (the reduced rows M are never written, so no backward; GSAGE_FUSED_LAYER=0 turns it off): the
            // neighbour half gather + mean + projection in one kernel, the self half as a one-segment projection
            const int p_f = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
            GS_TRY(gather_mean_project_launch(nb.base, nb.ld, nb.table_rows, d, nb.ids, n, S, e->w_n[layer].p, e->w_n[layer].ld, O,
                                              e->b_n[layer], act, out, out_dtype, ld_out, O, s));
            e->prof.end(p_f, s);
            e->prof.work(p_f, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + (double)O * dtype_size(out_dtype)),
                         2.0 * (double)n * d * O);
            if (dominant && e->ahead_after_gather) {
                if (!e->ev_mid) GS_CUDA(cudaEventCreateWithFlags(&e->ev_mid, cudaEventDisableTiming));
                GS_CUDA(cudaEventRecord(e->ev_mid, s));
                e->mid_valid = true;
            }
            const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
            const int st = linear_call(x, e->w_x[layer], O, e->b_x[layer], n, act, out, out_dtype, ld_out, 0, exact, s);
            e->prof.end(p_prj, s);
            e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)O * dtype_size(out_dtype)),
                         2.0 * (double)n * d * O);
            return st;
        }
        const int p_red = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
        GS_TRY(gather_reduce_launch(nb.base, nb.dtype, nb.ld, nb.table_rows, d, nb.ids, n, S, GSAGE_RED_MEAN, nullptr, Mb, T, ldm, s,
                                    (e->l2_hint && !e->keep_activations) ? 1 : 0));
        e->prof.end(p_red, s);
        if (dominant && e->ahead_after_gather) {
            if (!e->ev_mid) GS_CUDA(cudaEventCreateWithFlags(&e->ev_mid, cudaEventDisableTiming));
            GS_CUDA(cudaEventRecord(e->ev_mid, s));
            e->mid_valid = true;
        }
        // algorithmic bytes of the fused gather+mean launch: S rows + S ids in, one row out
        e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + (double)d * dtype_size(T)),
                     (double)n * S * d);
        RowSrc m{Mb, T, ldm, n, nullptr, d};
        const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
        const int st = combine_call(x, e->w_x[layer], m, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], e->b_n[layer]);
        e->prof.end(p_prj, s);
        e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)d * dtype_size(T) +
                                                             2.0 * O * dtype_size(out_dtype)), 4.0 * (double)n * d * O);
        return st;
    }
    case GSAGE_AGG_MAX_POOL:
    case GSAGE_AGG_MEAN_POOL: {
        const int H = e->hid;
        if (!exact) {
            // MLP on tcgen05 with the pool (max / mean over the S neighbour rows) done in the epilogue: the (n*S, 512)
            // hidden rows never exist in HBM
            LinearParams P;
            void* const Pb = (char*)e->Pp + m_row0 * (int64_t)H * (int64_t)dtype_size(T);       // this application's pooled rows
            P.n_segs = 1; P.n = n * S; P.act = GSAGE_ACT_RELU; P.out = Pb; P.out_dtype = T; P.ld_out = H;
            P.seg[0] = LinearSeg{nb.base, nb.dtype, nb.ld, nb.ids, e->w_mlp[layer].p, e->w_mlp[layer].dtype, e->w_mlp[layer].ld, d, H,
                                 e->b_mlp[layer], 0};
            P.pool_S = S; P.pool_max = e->cfg.aggregator == GSAGE_AGG_MAX_POOL ? 1 : 0;
            P.seg[0].a_rows = nb.ids ? nb.table_rows : 0;
            if (linear_pool_umma_eligible(P)) {
                // stopwatch: the dominant launch (layer 1 on the (x1, x2) pair) -- tensor-bound: 2*d*H flop per neighbour row
                const bool dominant = layer == 0 && n > e->B;
                const int p_red = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
                GS_TRY(linear_dispatch(P, 0, s));
                e->prof.end(p_red, s);
                e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + (double)H * dtype_size(T)),
                             2.0 * (double)n * S * d * H);
                RowSrc p{Pb, T, H, n, nullptr, H};
                const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
                const int st = combine_call(x, e->w_x[layer], p, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], nullptr);
                e->prof.end(p_prj, s);
                e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)H * dtype_size(T) +
                                                                     2.0 * O * dtype_size(out_dtype)), 2.0 * (double)n * (d + H) * O);
                return st;
            }
        }
        // fp32-exact mode (or operands the fused kernel does not take): the MLP as a projection (3 x TF32 in 128-column blocks when
        // exact), the hidden rows through HBM, the pool as a segment reduce.  Stopwatch: MLP + pool of the dominant application.
        const bool dominant = layer == 0 && n > e->B;
        const int p_red = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
        GS_TRY(linear_call(nb, e->w_mlp[layer], H, e->b_mlp[layer], n * S, GSAGE_ACT_RELU, e->HN, T, H, 0, exact, s));
        GS_TRY(gather_reduce_launch(e->HN, T, H, n * S, H, nullptr, n, S,
                                    e->cfg.aggregator == GSAGE_AGG_MAX_POOL ? GSAGE_RED_MAX : GSAGE_RED_MEAN, nullptr,
                                    e->Pp, T, H, s));
        e->prof.end(p_red, s);
        e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + (double)H * dtype_size(T)),
                     2.0 * (double)n * S * d * H);
        RowSrc p{e->Pp, T, H, n, nullptr, H};
        const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
        const int st = combine_call(x, e->w_x[layer], p, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], nullptr);
        e->prof.end(p_prj, s);
        e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)H * dtype_size(T) +
                                                             2.0 * O * dtype_size(out_dtype)), 2.0 * (double)n * (d + H) * O);
        return st;
    }
    case GSAGE_AGG_ATTENTION: {
        const int H = e->hid;
        GS_CHECK_ARG(S > 1, "attention aggregator: S must be > 1");
        // a(x_i) for the n parents: two small projections (n rows, not n*S)
        GS_TRY(linear_call(x, e->w_att1[layer], H, e->b_att[layer], n, GSAGE_ACT_TANH, e->T1x, GSAGE_F32, H, 0, exact, s));
        RowSrc t1x{e->T1x, GSAGE_F32, H, n, nullptr, H};
        GS_TRY(linear_call(t1x, f32w(L.att_w2, H), H, nullptr, n, GSAGE_ACT_NONE, e->XA, GSAGE_F32, H, 0, 1, s));
        if (!exact && attention_fused_eligible(nb.base, nb.dtype, nb.ld, d, e->w_att1[layer].p, e->w_att1[layer].dtype, e->w_att1[layer].ld, H, S,
                                               n, Mb, ldm, T)) {
            // bf16 mode: scores (tcgen05), softmax and the weighted sum in ONE kernel -- every neighbour row is read once
            // stopwatch: the dominant launch (layer 1 on the (x1, x2) pair) -- HBM-bound by bytes: S rows + S ids + a(x) in, one row out
            const bool dominant = layer == 0 && n > e->B;
            const int p_red = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
            GS_TRY(attention_fused_launch(nb.base, nb.ld, nb.ids, d, e->w_att1[layer].p, e->w_att1[layer].ld, e->b_att[layer], L.att_w2,
                                          (const float*)e->XA, n, S, Mb, T, ldm, s, nb.ids ? nb.table_rows : 0));
            e->prof.end(p_red, s);
            e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + 4.0 * H + (double)d * dtype_size(T)),
                         (double)n * S * (2.0 * d * H + 2.0 * H * H + 2.0 * H + 2.0 * d));
            RowSrc mf{Mb, T, ldm, n, nullptr, d};
            const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
            const int st = combine_call(x, e->w_x[layer], mf, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], e->b_n[layer]);
            e->prof.end(p_prj, s);
            e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)d * dtype_size(T) +
                                                                 2.0 * O * dtype_size(out_dtype)), 4.0 * (double)n * d * O);
            return st;
        }
        // fp32-exact mode: the unfused chain -- a(n) for every neighbour row (W1 as 3 x TF32 when exact), scores + softmax, then the
        // weighted gather+sum (the rows are read twice).  Stopwatch: the whole reduction of the dominant application.
        const bool dominant = layer == 0 && n > e->B;
        const int p_red = dominant ? e->prof.begin(GSAGE_PROF_REDUCE, s) : -1;
        GS_TRY(linear_call(nb, e->w_att1[layer], H, e->b_att[layer], n * S, GSAGE_ACT_TANH, e->T1, GSAGE_F32, H, 0, exact, s));
        RowSrc t1{e->T1, GSAGE_F32, H, n * S, nullptr, H};
        GS_TRY(linear_call(t1, f32w(L.att_w2, H), H, nullptr, n * S, GSAGE_ACT_NONE, e->NA, GSAGE_F32, H, 0, 1, s));
        GS_TRY(gsage_attention_weights(e->NA, e->XA, GSAGE_F32, H, H, n, S, e->AW, s));
        GS_TRY(gather_reduce_launch(nb.base, nb.dtype, nb.ld, nb.table_rows, d, nb.ids, n, S, GSAGE_RED_SUM, e->AW, Mb, T, ldm, s));
        e->prof.end(p_red, s);
        e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n * ((double)S * d * dtype_size(nb.dtype) + (nb.ids ? 8.0 * S : 0.0) + 4.0 * H + (double)d * dtype_size(T)),
                     (double)n * S * (2.0 * d * H + 2.0 * H * H + 2.0 * H + 2.0 * d));
        RowSrc m{Mb, T, ldm, n, nullptr, d};
        const int p_prj = dominant ? e->prof.begin(GSAGE_PROF_PROJECT, s) : -1;
        const int st = combine_call(x, e->w_x[layer], m, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], e->b_n[layer]);
        e->prof.end(p_prj, s);
        e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)n * ((double)d * dtype_size(x.dtype) + (x.ids ? 8.0 : 0.0) + (double)d * dtype_size(T) +
                                                             2.0 * O * dtype_size(out_dtype)), 4.0 * (double)n * d * O);
        return st;
    }
    case GSAGE_AGG_LSTM: {
        // nn_modules.py:276-279: the S neighbour rows of a parent, in sampled order, through a one-layer LSTM; the last hidden
        // state is the aggregate.  For a block of parents: ONE projection x . W_ih^T over all their S rows (4H gate columns,
        // fp32), then per step h . W_hh^T and the cell update (lstm.cu), which reads step t of the block with a row stride of
        // S * 4H.  Parents are independent, so blocks (<= 2 GiB of gate rows) run one after the other.
        const int H = e->hid;
        const int64_t G4 = 4 * (int64_t)H, esT = (int64_t)dtype_size(T);
        GS_CHECK_ARG(e->LGX && n <= e->n1 && S <= e->lgx_rows, "lstm aggregator: workspace too small for %lld parents x %d", (long long)n, S);
        const int64_t np_max = e->lgx_rows / S;
        for (int64_t p0 = 0; p0 < n; p0 += np_max) {
            const int64_t np = std::min(np_max, n - p0);
            GS_TRY(linear_call(nb.shifted(p0 * S), e->w_ih[layer], (int)G4, nullptr, np * S, GSAGE_ACT_NONE, e->LGX, GSAGE_F32, G4, 0, exact, s));
            RowSrc hc{(char*)e->LH + p0 * H * esT, T, H, np, nullptr, H};
            for (int t = 0; t < S; ++t) {
                if (t > 0) GS_TRY(linear_call(hc, e->w_hh[layer], (int)G4, nullptr, np, GSAGE_ACT_NONE, e->LGH, GSAGE_F32, G4, 0, exact, s));
                GS_TRY(gsage_lstm_cell(e->LGX + (int64_t)t * G4, (int64_t)S * G4, e->LGH, G4, L.lstm_b_ih, L.lstm_b_hh, e->LC + p0 * H,
                                       (char*)e->LH + p0 * H * esT, T, H, np, H, t == 0, s));
            }
        }
        RowSrc hrow{e->LH, T, H, n, nullptr, H};
        return combine_call(x, e->w_x[layer], hrow, e->w_n[layer], O, n, act, out, out_dtype, ld_out, exact, s, e->b_x[layer], nullptr);
    }
    }
    set_error("engine: unknown aggregator %d", e->cfg.aggregator);
    return GSAGE_ERR_INVALID;
}

extern "C" {

int gsage_engine_create(const gsage_engine_config* cfg, gsage_engine** out) {
    GS_CHECK_ARG(cfg && out, "engine_create: NULL argument");
    GS_CHECK_ARG(cfg->n_layers == 2, "engine_create: exactly two layers (train.py:105-118)");
    GS_CHECK_ARG(cfg->fanout[0] > 0 && cfg->fanout[1] > 0, "engine_create: fanouts must be > 0");
    GS_CHECK_ARG(cfg->out_dim[0] > 0 && cfg->out_dim[1] > 0 && cfg->n_classes > 0, "engine_create: bad dims");
    GS_CHECK_ARG(cfg->max_batch > 0, "engine_create: max_batch must be > 0");
    GS_CHECK_ARG(cfg->compute_dtype == GSAGE_F32 || cfg->compute_dtype == GSAGE_BF16, "engine_create: bad compute dtype");
    GS_CHECK_ARG(cfg->aggregator >= GSAGE_AGG_MEAN && cfg->aggregator <= GSAGE_AGG_LSTM, "engine_create: bad aggregator");
    GS_CHECK_ARG(cfg->prep >= GSAGE_PREP_IDENTITY && cfg->prep <= GSAGE_PREP_LINEAR, "engine_create: bad prep");
    const bool has_feats = cfg->feats_dev != nullptr;
    if (cfg->prep == GSAGE_PREP_IDENTITY || cfg->prep == GSAGE_PREP_LINEAR)
        GS_CHECK_ARG(has_feats, "engine_create: this prep needs node features");
    if (cfg->prep == GSAGE_PREP_NODE_EMBEDDING)
        GS_CHECK_ARG(cfg->emb_dev && cfg->emb_dim > 0 && cfg->n_nodes > 0, "engine_create: node_embedding needs the table");
    if (has_feats) GS_CHECK_ARG(cfg->feats_dtype == cfg->compute_dtype, "engine_create: feats dtype must equal compute dtype");

    gsage_engine* e = new gsage_engine();
    e->cfg = *cfg;
    e->T = cfg->compute_dtype;
    if (const char* f = getenv("GSAGE_FUSE_MEAN")) e->fuse_mean = atoi(f) != 0;
    if (const char* f = getenv("GSAGE_CHUNK")) e->chunk_parents = atoll(f);
    if (const char* f = getenv("GSAGE_L2HINT")) e->l2_hint = atoi(f);
    if (const char* f = getenv("GSAGE_AHEAD_AFTER_GATHER")) e->ahead_after_gather = atoi(f);
    // GSAGE_AHEAD_SPLIT=1: the next batch's RNG draws (count / scan / scatter, refill included) under the dominant launch instead of
    // under the projection tail.  Two calls on the final build (profiles/r02_ahead_split.txt): step -1.8 % then +-0 on reddit, the
    // gather kernel 2 % slower both times (0.986 -> 0.966 of the copy peak), the attention kernel 11 % slower: off
    if (const char* f = getenv("GSAGE_AHEAD_SPLIT")) e->ahead_split = atoi(f);
    const int64_t es = (int64_t)dtype_size(e->T);
    const int64_t vec = 16 / es;
    switch (cfg->prep) {
    case GSAGE_PREP_IDENTITY: e->d_prep = cfg->feats_dim; break;
    case GSAGE_PREP_LINEAR: e->d_prep = 32; break;            // overwritten by weights.prep_out_dim at set_weights
    case GSAGE_PREP_NODE_EMBEDDING: e->d_prep = (has_feats ? cfg->feats_dim : 0) + cfg->emb_dim; break;
    }
    e->fold_prep = cfg->prep == GSAGE_PREP_NODE_EMBEDDING && !has_feats && cfg->aggregator != GSAGE_AGG_LSTM;   // the LSTM reads materialised rows
    if (const char* f = getenv("GSAGE_FOLD_PREP")) e->fold_prep = e->fold_prep && atoi(f) != 0;
    e->ld_prep = (int)pad_to(e->d_prep, vec * 2);
    e->hid = cfg->hidden_dim > 0 ? cfg->hidden_dim : (cfg->aggregator == GSAGE_AGG_ATTENTION ? 32 : 512);
    e->maxB = cfg->max_batch;
    e->n0 = e->maxB; e->n1 = e->n0 * cfg->fanout[0]; e->n2 = e->n1 * cfg->fanout[1];
    const int64_t O1 = cfg->out_dim[0], O2 = cfg->out_dim[1];
    e->ld_h1 = pad_to(2 * O1, vec * 2);

    // carve the workspace
    int64_t off = 0;
    auto carve = [&](int64_t bytes) { int64_t at = off; off += pad_to(std::max<int64_t>(bytes, 16), 256); return at; };
    const int64_t o_ids = carve(8 * (e->n0 + e->n1 + e->n2));
    const int64_t o_ids1 = carve(8 * (e->n0 + e->n1 + e->n2));
    const int64_t o_sel = carve(4 * (e->n1 + e->n2));
    const int64_t o_sel1 = carve(4 * (e->n1 + e->n2));
    const int64_t o_look = carve(8 * e->n0);
    const int64_t o_X = (cfg->prep == GSAGE_PREP_IDENTITY || e->fold_prep) ? -1 : carve(es * e->ld_prep * (e->n0 + e->n1 + e->n2));
    const int64_t dmax = std::max<int64_t>(e->ld_prep, e->ld_h1);
    e->ld_m = pad_to(dmax, vec * 2);
    const int64_t o_M = carve(es * e->ld_m * (e->n0 + e->n1 + e->n0));
    int64_t o_HN = -1, o_P = -1, o_T1 = -1, o_NA = -1, o_T1x = -1, o_XA = -1, o_AW = -1;
    if (cfg->aggregator == GSAGE_AGG_MAX_POOL || cfg->aggregator == GSAGE_AGG_MEAN_POOL) {
        o_HN = carve(es * e->hid * e->n2);
        o_P = carve(es * e->hid * (e->n0 + e->n1 + e->n0));       // pooled rows of all three applications (kept for the backward)
    }
    if (cfg->aggregator == GSAGE_AGG_ATTENTION) {
        o_T1 = carve(4 * e->hid * e->n2); o_NA = carve(4 * e->hid * e->n2);
        o_T1x = carve(4 * e->hid * e->n1); o_XA = carve(4 * e->hid * e->n1);
        o_AW = carve(4 * e->n2);
    }
    int64_t o_LGX = -1, o_LGH = -1, o_LC = -1, o_LH = -1;
    if (cfg->aggregator == GSAGE_AGG_LSTM) {
        if (e->hid % 8 != 0) { set_error("engine_create: the LSTM state width must be a multiple of 8 (got %d)", e->hid); delete e; return GSAGE_ERR_INVALID; }
        // x_t . W_ih^T for every step of a block of parents at once: at most 2 GiB of gate rows, at least one parent
        const int64_t row_bytes = 4 * 4 * (int64_t)e->hid;
        e->lgx_rows = std::min<int64_t>(e->n2, std::max<int64_t>((int64_t(2) << 30) / row_bytes, std::max(cfg->fanout[0], cfg->fanout[1])));
        if (const char* f = getenv("GSAGE_LSTM_BLOCK_ROWS"))          // tests: force several blocks on a tiny batch
            e->lgx_rows = std::min<int64_t>(e->n2, std::max<int64_t>(atoll(f), std::max(cfg->fanout[0], cfg->fanout[1])));
        o_LGX = carve(row_bytes * e->lgx_rows); o_LGH = carve(row_bytes * e->n1);
        o_LC = carve(4 * (int64_t)e->hid * e->n1); o_LH = carve(es * (int64_t)e->hid * e->n1);
    }
    const int64_t o_MF = (cfg->aggregator == GSAGE_AGG_MEAN && cfg->prep == GSAGE_PREP_LINEAR) ? carve(es * cfg->feats_ld * (e->n0 + e->n1)) : -1;
    const int64_t o_H1 = carve(es * e->ld_h1 * (e->n0 + e->n1));
    const int64_t o_Z = carve(4 * 2 * O2 * e->n0);
    const int64_t o_ZN = carve(4 * 2 * O2 * e->n0);
    const int64_t o_LG = carve(4 * (int64_t)cfg->n_classes * e->n0);
    const int64_t o_LGb = carve(4 * (int64_t)cfg->n_classes * e->n0);
    const int64_t o_DZN = carve(4 * 2 * O2 * e->n0), o_DZ = carve(4 * 2 * O2 * e->n0);
    const int64_t o_DH0 = carve(4 * 2 * O1 * e->n0), o_DM2 = carve(4 * 2 * O1 * e->n0);
    const int64_t o_DZB = carve(2 * 2 * O2 * e->n0);
    const bool att_bwd = cfg->aggregator == GSAGE_AGG_ATTENTION && e->T == GSAGE_BF16 && cfg->prep == GSAGE_PREP_IDENTITY &&
                         getenv("GSAGE_NO_ATTENTION_BACKWARD") == nullptr;
    const int64_t o_ADM = att_bwd ? carve(4 * e->ld_m * (e->n0 + e->n1)) : -1, o_ADW = att_bwd ? carve(4 * e->n2) : -1;
    const int64_t o_ADA = att_bwd ? carve(4 * (int64_t)e->hid * e->n2) : -1, o_ADT1 = att_bwd ? carve(4 * (int64_t)e->hid * e->n2) : -1;
    const int64_t o_ADXA = att_bwd ? carve(4 * (int64_t)e->hid * e->n1) : -1, o_ADT1X = att_bwd ? carve(4 * (int64_t)e->hid * e->n1) : -1;
    const int64_t o_ASA = att_bwd ? carve(4 * 2 * O1 * e->n0) : -1, o_ASB = att_bwd ? carve(4 * 2 * O1 * e->n0) : -1;
    const int64_t o_ANA = att_bwd ? carve(4 * 2 * O1 * e->n1) : -1, o_ANB = att_bwd ? carve(4 * 2 * O1 * e->n1) : -1;
    const int64_t o_ADPB = att_bwd ? carve(2 * 128 * e->n2) : -1, o_AW1S = att_bwd ? carve(4 * 128 * e->ld_m) : -1;
    const int64_t o_AW2T = att_bwd ? carve(4 * 2 * (int64_t)e->hid * e->hid) : -1;
    const int64_t o_DH = carve(4 * 2 * O1 * (e->n0 + e->n1));
    // (n0 + n1) self rows for the mean recipe; the pool recipe also needs one row per sampled neighbour (n1 + n2)
    const int64_t o_DXE = e->fold_prep ? carve(4 * (int64_t)cfg->emb_dim * ((cfg->aggregator == GSAGE_AGG_MEAN ? 0 : e->n2) + e->n0 + e->n1)) : -1;
    const bool pool_cfg = cfg->aggregator == GSAGE_AGG_MAX_POOL || cfg->aggregator == GSAGE_AGG_MEAN_POOL;
    const bool pool_bwd = pool_cfg && e->T == GSAGE_BF16 && (cfg->prep == GSAGE_PREP_IDENTITY || e->fold_prep) &&
                          getenv("GSAGE_NO_POOL_BACKWARD") == nullptr;
    const int64_t o_DP = pool_bwd ? carve(4 * (int64_t)e->hid * (e->n0 + e->n1)) : -1;
    const int64_t o_DHID = pool_bwd ? carve(2 * (int64_t)e->hid * (e->n1 + e->n2)) : -1;
    const int64_t o_DN2 = pool_bwd ? carve(4 * 2 * O1 * e->n1) : -1;
    e->ws_bytes = off;
    if (cudaMalloc((void**)&e->ws, (size_t)off) != cudaSuccess) {
        set_error("engine_create: cudaMalloc of %lld workspace bytes failed", (long long)off);
        delete e;
        return GSAGE_ERR_NOMEM;
    }
    cudaMemset(e->ws, 0, (size_t)off);      // padding columns of every intermediate stay zero forever
    if (e->fold_prep) {
        e->fold_floats = 4 * ((int64_t)(2 * O1 + e->hid) * (cfg->emb_dim + 1) + 64);
        if (cudaMalloc((void**)&e->fold, sizeof(float) * (size_t)e->fold_floats) != cudaSuccess) {
            set_error("engine_create: cudaMalloc of the folded-weight arena failed");
            cudaFree(e->ws); delete e; return GSAGE_ERR_NOMEM;
        }
    }
    if (e->T == GSAGE_BF16) {                // arena for the padded bf16 weight copies the tensor-core kernel reads
        const int64_t dmax0 = std::max<int64_t>(e->ld_prep, 2 * O1) + 8;
        e->wb_bytes = 2 * (2 * 2 * ((O1 + O2) * 2 * (dmax0 + e->hid) + 2 * (int64_t)e->hid * dmax0) + 16 * 256);   // + the transposed copies
        if (cfg->aggregator == GSAGE_AGG_LSTM) e->wb_bytes += 2 * 2 * 4 * (int64_t)e->hid * (dmax0 + e->hid + 16) + 8 * 256;   // W_ih, W_hh of both layers
        if (cudaMalloc((void**)&e->wb, (size_t)e->wb_bytes) != cudaSuccess) {
            set_error("engine_create: cudaMalloc of the weight arena failed");
            cudaFree(e->ws);
            delete e;
            return GSAGE_ERR_NOMEM;
        }
    }
    auto at = [&](int64_t o) -> char* { return o < 0 ? nullptr : e->ws + o; };
    e->ids = (int64_t*)at(o_ids); e->sel = (uint32_t*)at(o_sel); e->look0 = (int64_t*)at(o_look);
    e->ids_slot[0] = e->ids; e->ids_slot[1] = (int64_t*)at(o_ids1); e->cur = 0; e->sel_ahead = (uint32_t*)at(o_sel1);
    e->X = at(o_X); e->M = at(o_M); e->HN = at(o_HN); e->Pp = at(o_P);
    e->T1 = at(o_T1); e->NA = at(o_NA); e->T1x = at(o_T1x); e->XA = at(o_XA); e->AW = (float*)at(o_AW);
    e->MF = at(o_MF);
    e->LGX = (float*)at(o_LGX); e->LGH = (float*)at(o_LGH); e->LC = (float*)at(o_LC); e->LH = at(o_LH);
    e->H1 = at(o_H1); e->Z = (float*)at(o_Z); e->ZN = (float*)at(o_ZN); e->LG = (float*)at(o_LG);
    e->LG2[0] = e->LG; e->LG2[1] = (float*)at(o_LGb);
    e->DXE = (float*)at(o_DXE); e->DZB = at(o_DZB);
    e->ADM = (float*)at(o_ADM); e->ADW = (float*)at(o_ADW); e->ADA = (float*)at(o_ADA); e->ADT1 = (float*)at(o_ADT1);
    e->ADXA = (float*)at(o_ADXA); e->ADT1X = (float*)at(o_ADT1X);
    e->ADPB = at(o_ADPB); e->AW1S = (float*)at(o_AW1S);
    if (o_AW2T >= 0) { e->AW2T[0] = (float*)at(o_AW2T); e->AW2T[1] = e->AW2T[0] + (int64_t)e->hid * e->hid; }
    e->ASA = (float*)at(o_ASA); e->ASB = (float*)at(o_ASB); e->ANA = (float*)at(o_ANA); e->ANB = (float*)at(o_ANB);
    e->DP = (float*)at(o_DP); e->DHID = at(o_DHID); e->DN2 = (float*)at(o_DN2);
    e->DZN = (float*)at(o_DZN); e->DZ = (float*)at(o_DZ); e->DH0 = (float*)at(o_DH0); e->DM2 = (float*)at(o_DM2); e->DH = (float*)at(o_DH);
    *out = e;
    return GSAGE_OK;
}

void gsage_engine_destroy(gsage_engine* e) {
    if (!e) return;
    for (cudaEvent_t ev : e->prof.pool) cudaEventDestroy(ev);
    if (e->ss) { cudaStreamSynchronize(e->ss); cudaStreamDestroy(e->ss); }
    if (e->cs) { cudaStreamSynchronize(e->cs); cudaStreamDestroy(e->cs); }
    if (e->ev_fwd) cudaEventDestroy(e->ev_fwd);
    if (e->ev_copy) cudaEventDestroy(e->ev_copy);
    if (e->ev_src) cudaEventDestroy(e->ev_src);
    for (int i = 0; i < 2; ++i) if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    if (e->ev_mid) cudaEventDestroy(e->ev_mid);
    if (e->ev_ahead) cudaEventDestroy(e->ev_ahead);
    cudaFree(e->ws);
    cudaFree(e->wb);
    cudaFree(e->fold);
    cudaFree(e->wsplit);
    cudaFree(e->FCP);
    delete e;
}

int64_t gsage_engine_workspace_bytes(const gsage_engine* e) { return e ? e->ws_bytes : 0; }

int gsage_engine_keep_activations(gsage_engine* e, int keep) {
    GS_CHECK_ARG(e, "engine_keep_activations: NULL engine");
    e->keep_activations = keep != 0;
    return GSAGE_OK;
}

int gsage_engine_profile(gsage_engine* e, int enable) {
    GS_CHECK_ARG(e, "engine_profile: NULL engine");
    e->prof.on = enable != 0;
    e->prof.recs.clear();
    e->prof.used = 0;
    for (int i = 0; i < GSAGE_PROF_CATS; ++i) e->prof.bytes[i] = e->prof.flops[i] = 0;
    return GSAGE_OK;
}

int gsage_engine_profile_read(gsage_engine* e, double* ms_out, int64_t* launches_out, double* bytes_out, double* flops_out, void* stream) {
    GS_CHECK_ARG(e && ms_out && launches_out && bytes_out && flops_out, "engine_profile_read: NULL argument");
    GS_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (e->ss) GS_CUDA(cudaStreamSynchronize(e->ss));          // SAMPLE records live on the sampler stream when sampling runs ahead
    for (int i = 0; i < GSAGE_PROF_CATS; ++i) { ms_out[i] = 0; launches_out[i] = 0; bytes_out[i] = e->prof.bytes[i]; flops_out[i] = e->prof.flops[i]; }
    for (const EngineProf::Rec& r : e->prof.recs) {
        float ms = 0;
        GS_CUDA(cudaEventElapsedTime(&ms, e->prof.pool[r.a], e->prof.pool[r.b]));
        ms_out[r.cat] += ms;
        launches_out[r.cat] += 1;
    }
    e->prof.recs.clear();
    e->prof.used = 0;
    for (int i = 0; i < GSAGE_PROF_CATS; ++i) e->prof.bytes[i] = e->prof.flops[i] = 0;
    return GSAGE_OK;
}

int gsage_engine_set_weights(gsage_engine* e, const gsage_weights* w, void* stream) {
    GS_CHECK_ARG(e && w, "engine_set_weights: NULL argument");
    for (int l = 0; l < 2; ++l) {
        GS_CHECK_ARG(w->layer[l].fc_x && w->layer[l].fc_neib, "engine_set_weights: fc_x / fc_neib missing (layer %d)", l);
        if (e->cfg.aggregator == GSAGE_AGG_MAX_POOL || e->cfg.aggregator == GSAGE_AGG_MEAN_POOL)
            GS_CHECK_ARG(w->layer[l].mlp_w && w->layer[l].mlp_b, "engine_set_weights: pool MLP missing (layer %d)", l);
        if (e->cfg.aggregator == GSAGE_AGG_ATTENTION)
            GS_CHECK_ARG(w->layer[l].att_w1 && w->layer[l].att_w2, "engine_set_weights: attention MLP missing (layer %d)", l);
        if (e->cfg.aggregator == GSAGE_AGG_LSTM)
            GS_CHECK_ARG(w->layer[l].lstm_w_ih && w->layer[l].lstm_w_hh && w->layer[l].lstm_b_ih && w->layer[l].lstm_b_hh,
                         "engine_set_weights: LSTM weights missing (layer %d)", l);
    }
    GS_CHECK_ARG(w->fc_w && w->fc_b, "engine_set_weights: classifier missing");
    if (e->cfg.prep == GSAGE_PREP_NODE_EMBEDDING) GS_CHECK_ARG(w->prep_fc_w && w->prep_fc_b, "engine_set_weights: prep.fc missing");
    if (e->cfg.prep == GSAGE_PREP_LINEAR) {
        GS_CHECK_ARG(w->prep_fc_w && w->prep_out_dim > 0, "engine_set_weights: prep.fc missing");
        GS_CHECK_ARG(w->prep_out_dim <= e->ld_prep, "engine_set_weights: LinearPrep wider than the workspace rows");
        e->d_prep = w->prep_out_dim;
    }
    e->w = *w;
    cudaStream_t s = as_stream(stream);
    gsage_weights eff = *w;                                   // the weights the kernels will actually use
    for (int l = 0; l < 2; ++l) { e->b_x[l] = e->b_n[l] = e->b_att[l] = nullptr; e->b_mlp[l] = w->layer[l].mlp_b; }
    if (e->fold_prep) {
        // W' = W . Wp (rows x de), b' = W . bp (+ b): three or four tiny FFMA launches per set_weights
        const int de = e->cfg.emb_dim, O = e->cfg.out_dim[0], H = e->hid;
        float* at = e->fold;
        auto fold_w = [&](const float* W, int rows, const float* extra_bias, const float** w_out, const float** b_out) -> int {
            float* Wf = at; at += (int64_t)rows * de;
            float* bf = at; at += rows;
            GS_TRY(linear_trans_call(W, de, de, w->prep_fc_w, de, de, rows, Wf, de, s));               // W . Wp
            LinearParams P;                                                                             // W . bp
            P.n_segs = 1; P.n = rows; P.act = GSAGE_ACT_NONE; P.out = bf; P.out_dtype = GSAGE_F32; P.ld_out = 1;
            P.seg[0] = LinearSeg{W, GSAGE_F32, (int64_t)de, nullptr, w->prep_fc_b, GSAGE_F32, (int64_t)de, de, 1, nullptr, 0};
            GS_TRY(linear_simt_launch(P, s));
            if (extra_bias) { axpy_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, s>>>(bf, extra_bias, rows); GS_LAUNCHED(); }
            *w_out = Wf; *b_out = bf;
            return GSAGE_OK;
        };
        const bool pool0 = e->cfg.aggregator == GSAGE_AGG_MAX_POOL || e->cfg.aggregator == GSAGE_AGG_MEAN_POOL;
        GS_TRY(fold_w(w->layer[0].fc_x, O, nullptr, &eff.layer[0].fc_x, &e->b_x[0]));
        if (!pool0) GS_TRY(fold_w(w->layer[0].fc_neib, O, nullptr, &eff.layer[0].fc_neib, &e->b_n[0]));
        e->fold_wx = eff.layer[0].fc_x;
        e->fold_wn = pool0 ? nullptr : eff.layer[0].fc_neib;
        if (pool0) GS_TRY(fold_w(w->layer[0].mlp_w, H, w->layer[0].mlp_b, &eff.layer[0].mlp_w, &e->b_mlp[0]));
        if (e->cfg.aggregator == GSAGE_AGG_ATTENTION) GS_TRY(fold_w(w->layer[0].att_w1, H, nullptr, &eff.layer[0].att_w1, &e->b_att[0]));
        GS_CHECK_ARG(at - e->fold <= e->fold_floats, "engine_set_weights: folded-weight arena too small");
    }
    w = &eff;
    const bool pool = e->cfg.aggregator == GSAGE_AGG_MAX_POOL || e->cfg.aggregator == GSAGE_AGG_MEAN_POOL;
    const bool att = e->cfg.aggregator == GSAGE_AGG_ATTENTION;
    int64_t off = 0, split_off = 0;
    if (e->T == GSAGE_F32 && !e->cfg.allow_tf32 && !e->wsplit) {
        int64_t need = 0;
        for (int l = 0; l < 2; ++l) {
            const int64_t d_in = l == 0 ? e->ld_prep : 2 * e->cfg.out_dim[0];
            need += 2 * 2 * pad_to((int64_t)e->cfg.out_dim[l] * (d_in > e->hid ? d_in : e->hid), 64);      // fc_x, fc_neib: hi and lo
            if (pool || att) need += 2 * pad_to((int64_t)e->hid * d_in, 64);                                // mlp.0 / att.0: hi and lo
        }
        GS_CUDA(cudaMalloc((void**)&e->wsplit, need * sizeof(float)));
        e->wsplit_floats = need;
    }
    for (int l = 0; l < 2; ++l) {
        const int d_in = l == 0 ? e->d_prep : 2 * e->cfg.out_dim[0];
        const bool lstm = e->cfg.aggregator == GSAGE_AGG_LSTM;
        const int d_nb = (pool || lstm) ? e->hid : d_in;
        const int O = e->cfg.out_dim[l];
        struct Item { const float* src; int rows, cols; WRef* dst; } items[6] = {
            {w->layer[l].fc_x, O, d_in, &e->w_x[l]}, {w->layer[l].fc_neib, O, d_nb, &e->w_n[l]},
            {pool ? w->layer[l].mlp_w : nullptr, e->hid, d_in, &e->w_mlp[l]}, {att ? w->layer[l].att_w1 : nullptr, e->hid, d_in, &e->w_att1[l]},
            {lstm ? w->layer[l].lstm_w_ih : nullptr, 4 * e->hid, d_in, &e->w_ih[l]}, {lstm ? w->layer[l].lstm_w_hh : nullptr, 4 * e->hid, e->hid, &e->w_hh[l]}};
        for (const Item& it : items) {
            if (!it.src) { *it.dst = WRef{nullptr, GSAGE_F32, 0}; continue; }
            if (e->T != GSAGE_BF16) {
                *it.dst = f32w(it.src, it.cols);
                const int64_t cnt = pad_to((int64_t)it.rows * it.cols, 64);
                if (e->wsplit && (it.dst == &e->w_x[l] || it.dst == &e->w_n[l] || it.dst == &e->w_mlp[l] || it.dst == &e->w_att1[l]) &&
                    split_off + 2 * cnt <= e->wsplit_floats) {
                    float* hi = e->wsplit + split_off; float* lo = hi + cnt;
                    GS_TRY(split_tf32_launch(it.src, (int64_t)it.rows * it.cols, hi, lo, s));
                    it.dst->hi = hi; it.dst->lo = lo;
                    split_off += 2 * cnt;
                }
                continue;
            }
            const int64_t ld = pad_to(it.cols, 8);
            const int64_t bytes = pad_to(2 * ld * it.rows, 256);
            GS_CHECK_ARG(off + bytes <= e->wb_bytes, "engine_set_weights: bf16 weight arena too small");
            __nv_bfloat16* dst = (__nv_bfloat16*)(e->wb + off);
            to_bf16_padded_kernel<<<(unsigned)ceil_div((int64_t)it.rows * ld, 256), 256, 0, s>>>(it.src, it.rows, it.cols, dst, ld);
            GS_LAUNCHED();
            *it.dst = WRef{dst, GSAGE_BF16, ld};
            off += bytes;
        }
        if (att && e->AW2T[l]) {
            transpose_f32_kernel<<<(unsigned)ceil_div((int64_t)e->hid * e->hid, 256), 256, 0, s>>>(w->layer[l].att_w2, e->hid, e->hid, e->AW2T[l]);
            GS_LAUNCHED();
        }
        if (att && e->T == GSAGE_BF16 && e->ADM) {
            // attention backward: fc_neib^T (d_in x O) for d M = Gn . Wn, and the layer-2 fc_x^T for d h0 = Gx . Wx2
            struct TItem { const float* src; int rows, cols; WRef* dst; } titems[2] = {
                {w->layer[l].fc_neib, O, d_in, &e->w_nT[l]}, {l == 1 ? w->layer[1].fc_x : nullptr, O, d_in, &e->w_x2T}};
            for (const TItem& it : titems) {
                if (!it.src) continue;
                const int64_t ld = pad_to(it.rows, 8);
                const int64_t bytes = pad_to(2 * ld * it.cols, 256);
                GS_CHECK_ARG(off + bytes <= e->wb_bytes, "engine_set_weights: bf16 weight arena too small (transposed copies)");
                __nv_bfloat16* dst = (__nv_bfloat16*)(e->wb + off);
                transpose_to_bf16_kernel<<<(unsigned)ceil_div((int64_t)it.cols * ld, 256), 256, 0, s>>>(it.src, it.rows, it.cols, dst, ld);
                GS_LAUNCHED();
                *it.dst = WRef{dst, GSAGE_BF16, ld};
                off += bytes;
            }
        }
        if (l == 1 && e->cfg.aggregator == GSAGE_AGG_MEAN && e->T == GSAGE_BF16) {
            // K-major transposed copies of the layer-2 weights: the head's data gradients d h0 = dz_x . Wx2, d m2 = dz_n . Wn2
            // run on the projection kernel as out = A . W'^T with W' = W^T
            struct TItem { const float* src; int rows, cols; WRef* dst; } titems[2] = {
                {w->layer[1].fc_x, O, d_in, &e->w_x2T}, {w->layer[1].fc_neib, O, d_in, &e->w_n2T}};
            for (const TItem& it : titems) {
                const int64_t ld = pad_to(it.rows, 8);
                const int64_t bytes = pad_to(2 * ld * it.cols, 256);
                GS_CHECK_ARG(off + bytes <= e->wb_bytes, "engine_set_weights: bf16 weight arena too small (transposed copies)");
                __nv_bfloat16* dst = (__nv_bfloat16*)(e->wb + off);
                transpose_to_bf16_kernel<<<(unsigned)ceil_div((int64_t)it.cols * ld, 256), 256, 0, s>>>(it.src, it.rows, it.cols, dst, ld);
                GS_LAUNCHED();
                *it.dst = WRef{dst, GSAGE_BF16, ld};
                off += bytes;
            }
        }
        if (pool && e->T == GSAGE_BF16 && e->DP) {
            // K-major transposed copies for the pool backward's data-gradient projections
            struct TItem { const float* src; int rows, cols; WRef* dst; } titems[3] = {
                {w->layer[l].fc_neib, O, e->hid, &e->w_nT[l]}, {w->layer[l].mlp_w, e->hid, d_in, &e->w_mlpT[l]},
                {(l == 0 && e->fold_prep) ? w->layer[0].fc_x : nullptr, O, d_in, &e->w_xT0}};
            for (const TItem& it : titems) {
                if (!it.src) continue;
                const int64_t ld = pad_to(it.rows, 8);
                const int64_t bytes = pad_to(2 * ld * it.cols, 256);
                GS_CHECK_ARG(off + bytes <= e->wb_bytes, "engine_set_weights: bf16 weight arena too small (transposed copies)");
                __nv_bfloat16* dst = (__nv_bfloat16*)(e->wb + off);
                transpose_to_bf16_kernel<<<(unsigned)ceil_div((int64_t)it.cols * ld, 256), 256, 0, s>>>(it.src, it.rows, it.cols, dst, ld);
                GS_LAUNCHED();
                *it.dst = WRef{dst, GSAGE_BF16, ld};
                off += bytes;
            }
        }
    }
    if (e->T == GSAGE_BF16) {
        // classifier: (n_classes, 2*O2) + bias padded with zero rows to a multiple of 16 (one small copy per set_weights)
        const int C = e->cfg.n_classes, Cp = (C + 15) / 16 * 16, K = 2 * e->cfg.out_dim[1];
        if (!e->FCP) GS_CUDA(cudaMalloc((void**)&e->FCP, sizeof(float) * (size_t)Cp * (K + 1)));
        GS_CUDA(cudaMemsetAsync(e->FCP, 0, sizeof(float) * (size_t)Cp * (K + 1), s));
        GS_CUDA(cudaMemcpyAsync(e->FCP, e->w.fc_w, sizeof(float) * (size_t)C * K, cudaMemcpyDeviceToDevice, s));
        GS_CUDA(cudaMemcpyAsync(e->FCP + (size_t)Cp * K, e->w.fc_b, sizeof(float) * C, cudaMemcpyDeviceToDevice, s));
        e->fc_pad = Cp;
    }
    e->have_weights = true;
    return GSAGE_OK;
}

// ---- sample: hop 0 draws first, then hop 1 (models.py:78-79), into `ids` = [ids0 | ids1 | ids2] -----------------
// `gate` / `n_gate`: events the stream must wait for AFTER the draws and BEFORE anything touches `ids` (sample-ahead: the
// draws do not depend on the ids, so they may start before the id slot is free)
static int sample_hops(gsage_engine* e, gsage_graph* g, gsage_rng* rng, int64_t* ids, uint32_t* sel, const int64_t* ids_src,
                       bool src_host, int64_t B, int64_t global_B, int64_t first, cudaStream_t s, cudaEvent_t* gate = nullptr, int n_gate = 0) {
    const int S1 = e->cfg.fanout[0], S2 = e->cfg.fanout[1];
    const int64_t n0 = B, n1 = B * S1, n2 = n1 * S2;
    const int p_smp = e->prof.begin(GSAGE_PROF_SAMPLE, s);
    int64_t* ids0 = ids; int64_t* ids1 = ids0 + n0; int64_t* ids2 = ids1 + n1;
    if (global_B == B) {
        // both hops draw from the same range [0, maxdeg) and the hop sizes do not depend on what was sampled, so the
        // n1 hop-0 draws followed by the n2 hop-1 draws are ONE bounded draw of n1 + n2 values from the stream (same words,
        // same order, same final position as two np.random.choice calls): one count / scan / scatter pass instead of two
        GS_TRY(rng_randint_internal(rng, (uint32_t)g->n_cols, n1 + n2, sel, s));
        for (int i = 0; i < n_gate; ++i) GS_CUDA(cudaStreamWaitEvent(s, gate[i], 0));
        if (ids_src != ids0) GS_CUDA(cudaMemcpyAsync(ids0, ids_src, 8 * n0, src_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
        GS_TRY(sample_sparse_launch(g, ids0, n0, S1, sel, ids1, s));
        GS_TRY(sample_sparse_launch(g, ids1, n1, S2, sel + n1, ids2, s));
    } else {
        // seed-sharded, still bit-exact with the single-process run: every rank consumes the draws of the WHOLE
        // global batch (hop-0 block, then hop-1 block -- the cheap part) and uses the slice that belongs to its seeds;
        // the gather / aggregate / project work is what is sharded
        for (int i = 0; i < n_gate; ++i) GS_CUDA(cudaStreamWaitEvent(s, gate[i], 0));
        if (ids_src != ids0) GS_CUDA(cudaMemcpyAsync(ids0, ids_src, 8 * n0, src_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
        const int64_t g1 = global_B * S1, g2 = g1 * S2;
        uint32_t* gsel = nullptr;
        GS_CUDA(cudaMallocAsync((void**)&gsel, sizeof(uint32_t) * g2, s));
        int st = rng_randint_internal(rng, (uint32_t)g->n_cols, g1, gsel, s);
        if (st == GSAGE_OK) st = sample_sparse_launch(g, ids0, n0, S1, gsel + first * S1, ids1, s);
        if (st == GSAGE_OK) st = rng_randint_internal(rng, (uint32_t)g->n_cols, g2, gsel, s);
        if (st == GSAGE_OK) st = sample_sparse_launch(g, ids1, n1, S2, gsel + first * S1 * S2, ids2, s);
        cudaFreeAsync(gsel, s);
        GS_TRY(st);
    }
    e->prof.end(p_smp, s);
    return GSAGE_OK;
}

// the id slot in use has been read by everything queued on `s` so far (sample-ahead may recycle it after this point)
static int mark_slot_done(gsage_engine* e, cudaStream_t s) {
    if (!e->ev_done[e->cur]) GS_CUDA(cudaEventCreateWithFlags(&e->ev_done[e->cur], cudaEventDisableTiming));
    GS_CUDA(cudaEventRecord(e->ev_done[e->cur], s));
    e->done_valid[e->cur] = true;
    return GSAGE_OK;
}

static int check_batch_args(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const void* ids, int64_t B, int64_t global_B, int64_t first) {
    GS_CHECK_ARG(e && g && rng && ids, "engine_forward: NULL argument");
    GS_CHECK_ARG(global_B >= B && first >= 0 && first + B <= global_B, "engine_forward_sharded: slice [%lld, %lld) outside the global batch of %lld",
                 (long long)first, (long long)(first + B), (long long)global_B);
    GS_CHECK_ARG(B > 0 && B <= e->maxB, "engine_forward: batch %lld outside (0, max_batch=%lld]", (long long)B, (long long)e->maxB);
    GS_CHECK_ARG(g->n_cols >= 1 && g->n_cols <= 0xFFFFFFFFLL, "engine_forward: adjacency width out of range");
    return GSAGE_OK;
}

static int sample_ahead_impl(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_src, bool src_host, int64_t B,
                             int64_t global_B, int64_t first, cudaStream_t main) {
    GS_TRY(check_batch_args(e, g, rng, ids_src, B, global_B, first));
    GS_CHECK_ARG(!e->ahead.valid, "engine_sample_ahead: a sampled-ahead batch is already pending (run its forward first)");
    GS_CHECK_ARG(!e->pre.valid, "engine_sample_ahead: a host forward is already queued (gsage_engine_forward_host_next): collect it first");
    if (!e->ss) {
        int lo = 0, hi = 0;
        GS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // hi = numerically lowest = greatest priority
        GS_CUDA(cudaStreamCreateWithPriority(&e->ss, cudaStreamNonBlocking, hi));
        GS_CUDA(cudaEventCreateWithFlags(&e->ev_ahead, cudaEventDisableTiming));
    }
    // the spare slot was last read by the forward (and backward) BEFORE the one in flight: wait for that one only, so
    // the draws run underneath the forward that was queued just before this call
    // the draws (count / scan / scatter of the RNG stream: no dependence on the ids) start right away; everything that reads or
    // writes the id slot waits for the slot's last reader and -- mean aggregator -- for the dominant gather launch of the forward
    // in flight, so that the random-access sample kernels share the SMs with the projection tail, not with that kernel
    cudaEvent_t gate[3]; int n_gate = 0;
    if (!src_host) {
        // device ids: whatever produced them on the caller's stream must have finished before the sampler stream copies them.
        // gsage_engine_inputs_ready marks that point BEFORE the forward in flight was queued (the overlap survives); without
        // it the point is "now", i.e. behind that forward -- always correct, no overlap
        if (!e->ev_src) GS_CUDA(cudaEventCreateWithFlags(&e->ev_src, cudaEventDisableTiming));
        if (!e->src_marked) GS_CUDA(cudaEventRecord(e->ev_src, main));
        e->src_marked = false;
        gate[n_gate++] = e->ev_src;
    }
    if (e->done_valid[e->cur ^ 1]) gate[n_gate++] = e->ev_done[e->cur ^ 1];
    if (e->mid_valid) { gate[n_gate++] = e->ev_mid; e->mid_valid = false; }
    if (e->ahead_split) {
        GS_TRY(sample_hops(e, g, rng, e->ids_slot[e->cur ^ 1], e->sel_ahead, ids_src, src_host, B, global_B, first, e->ss, gate, n_gate));
    } else {
        for (int i = 0; i < n_gate; ++i) GS_CUDA(cudaStreamWaitEvent(e->ss, gate[i], 0));
        GS_TRY(sample_hops(e, g, rng, e->ids_slot[e->cur ^ 1], e->sel_ahead, ids_src, src_host, B, global_B, first, e->ss));
    }
    GS_CUDA(cudaEventRecord(e->ev_ahead, e->ss));
    e->ahead.valid = true; e->ahead.src = ids_src; e->ahead.B = B; e->ahead.global_B = global_B; e->ahead.first = first;
    e->ahead.g = g; e->ahead.rng = rng;
    return GSAGE_OK;
}

int gsage_engine_sample_ahead(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                              int64_t global_B, int64_t first, void* stream) {
    return sample_ahead_impl(e, g, rng, ids_dev, false, B, global_B, first, as_stream(stream));
}

int gsage_engine_sample_ahead_host(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                                   void* stream) {
    return sample_ahead_impl(e, g, rng, ids_host, true, B, B, 0, as_stream(stream));
}

int gsage_engine_sample_ahead_pending(const gsage_engine* e) { return (e && e->ahead.valid) ? 1 : 0; }

int gsage_engine_inputs_ready(gsage_engine* e, void* stream) {
    GS_CHECK_ARG(e, "engine_inputs_ready: NULL engine");
    if (!e->ev_src) GS_CUDA(cudaEventCreateWithFlags(&e->ev_src, cudaEventDisableTiming));
    GS_CUDA(cudaEventRecord(e->ev_src, as_stream(stream)));
    e->src_marked = true;
    return GSAGE_OK;
}

// The sticky error flags (rng look-ahead window, out-of-range ids) live in mapped pinned host memory (graph.cu, mt19937.cu):
// polling them costs neither a copy nor a synchronisation, so every forward can look at what the PREVIOUS forwards reported --
// an error surfaces one call late instead of never.
int gsage_engine_poll_errors(gsage_engine* e) {
    GS_CHECK_ARG(e, "engine_poll_errors: NULL engine");
    if (e->last_rng && *(volatile int*)e->last_rng->err_flag) {
        set_error("rng: the look-ahead window held fewer accepted draws than requested (12-sigma event) -- re-seed; results since the last "
                  "check are invalid");
        return GSAGE_ERR_RNG;
    }
    if (e->last_g && *(volatile int*)e->last_g->err_flag) {
        *(volatile int*)e->last_g->err_flag = 0;
        set_error("index out of range: a seed or sampled id of an earlier forward lies outside the adjacency (%lld rows) -- "
                  "`feats[ids]` (models.py:76) / scipy raise IndexError there; its samples were read as the dummy node",
                  (long long)e->last_g->n_rows);
        return GSAGE_ERR_INDEX;
    }
    return GSAGE_OK;
}

static int forward_impl(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_src, bool src_host, int64_t B,
                        int64_t global_B, int64_t first, float* logits_dev, cudaStream_t s);
static int forward_layers(gsage_engine* e, int64_t B, float* logits_dev, cudaStream_t s, int p_all);

int gsage_engine_forward(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                         float* logits_dev, void* stream) {
    GS_CHECK_ARG(e && !e->pre.valid, "engine_forward: a host forward is already queued (gsage_engine_forward_host_next): collect it first");
    return forward_impl(e, g, rng, ids_dev, false, B, B, 0, logits_dev, as_stream(stream));
}

int gsage_engine_forward_sharded(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_dev, int64_t B,
                                 int64_t global_B, int64_t first, float* logits_dev, void* stream) {
    GS_CHECK_ARG(e && !e->pre.valid, "engine_forward_sharded: a host forward is already queued (gsage_engine_forward_host_next): collect it first");
    return forward_impl(e, g, rng, ids_dev, false, B, global_B, first, logits_dev, as_stream(stream));
}

static int forward_impl(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_src, bool src_host, int64_t B,
                        int64_t global_B, int64_t first, float* logits_dev, cudaStream_t s) {
    GS_TRY(check_batch_args(e, g, rng, ids_src, B, global_B, first));
    GS_CHECK_ARG(logits_dev, "engine_forward: NULL argument");
    e->last_g = g; e->last_rng = rng;
    GS_CHECK_ARG(e->have_weights, "engine_forward: call gsage_engine_set_weights first");

    const int p_all = e->prof.begin(GSAGE_PROF_FORWARD, s);
    if (e->ahead.valid) {
        // this batch was sampled ahead: same draws, same order on the RNG stream -- only wait for the sampler stream
        GS_CHECK_ARG(e->ahead.src == ids_src && e->ahead.B == B && e->ahead.global_B == global_B && e->ahead.first == first &&
                     e->ahead.g == g && e->ahead.rng == rng,
                     "engine_forward: a different batch was sampled ahead (its draws are already consumed): run the forward of "
                     "that batch (same ids pointer, batch size, graph and rng) first");
        const int p_wait = e->prof.begin(GSAGE_PROF_WAIT, s);
        GS_CUDA(cudaStreamWaitEvent(s, e->ev_ahead, 0));
        e->prof.end(p_wait, s);
        e->cur ^= 1;
        e->ids = e->ids_slot[e->cur];
        e->ahead.valid = false;
    } else {
        GS_TRY(sample_hops(e, g, rng, e->ids, e->sel, ids_src, src_host, B, global_B, first, s));
    }
    return forward_layers(e, B, logits_dev, s, p_all);
}

// everything after the sampling: prep, both aggregator layers, normalise, classifier -- reads the hop ids of e->ids
static int forward_layers(gsage_engine* e, int64_t B, float* logits_dev, cudaStream_t s, int p_all) {
    const gsage_engine_config& c = e->cfg;
    const int S1 = c.fanout[0], S2 = c.fanout[1], T = e->T;
    const int64_t n0 = B, n1 = B * S1, n2 = n1 * S2;
    const int64_t es = (int64_t)dtype_size(T);
    e->B = B;
    int64_t* ids0 = e->ids; int64_t* ids1 = ids0 + n0;

    // ---- prep (models.py:76-81) ------------------------------------------------------------------------
    RowSrc lvl;       // all three hops, hop k starts `offset_k` rows in
    RowSrc x_seed;    // the self rows of the seeds (layer_idx == 0); differs from `lvl` only under the folded prep
    if (c.prep == GSAGE_PREP_IDENTITY) {
        lvl = RowSrc{c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, ids0, c.feats_dim};
    } else if (e->fold_prep) {
        // aggregators read the raw embedding table by id; seeds look up the masked row n_nodes (nn_modules.py:149)
        fill_i64_kernel<<<(unsigned)ceil_div(n0, 256), 256, 0, s>>>(e->look0, n0, c.n_nodes);
        GS_LAUNCHED();
        lvl = RowSrc{c.emb_dev, c.emb_dtype, c.emb_ld, c.n_nodes + 1, ids0, c.emb_dim};
    } else {
        const int64_t ntot = n0 + n1 + n2;
        const int64_t ldx = e->ld_prep;
        if (c.prep == GSAGE_PREP_LINEAR) {
            RowSrc f{c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, ids0, c.feats_dim};
            GS_TRY(linear_call(f, f32w(e->w.prep_fc_w, c.feats_dim), e->d_prep, nullptr, ntot, GSAGE_ACT_NONE, e->X, T, ldx, 0,
                               T == GSAGE_F32, s));
        } else {
            const int dfe = c.feats_dev ? c.feats_dim : 0;
            if (c.feats_dev)
                GS_TRY(gather_reduce_launch(c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, c.feats_dim, ids0, ntot, 1,
                                            GSAGE_RED_SUM, nullptr, e->X, T, ldx, s));
            // seeds look up the masked row n_nodes (nn_modules.py:149); deeper hops their own id
            fill_i64_kernel<<<(unsigned)ceil_div(n0, 256), 256, 0, s>>>(e->look0, n0, c.n_nodes);
            GS_LAUNCHED();
            RowSrc e0{c.emb_dev, c.emb_dtype, c.emb_ld, c.n_nodes + 1, e->look0, c.emb_dim};
            GS_TRY(linear_call(e0, f32w(e->w.prep_fc_w, c.emb_dim), c.emb_dim, e->w.prep_fc_b, n0, GSAGE_ACT_NONE, e->X, T, ldx,
                               dfe, T == GSAGE_F32, s));
            RowSrc e12{c.emb_dev, c.emb_dtype, c.emb_ld, c.n_nodes + 1, ids1, c.emb_dim};
            GS_TRY(linear_call(e12, f32w(e->w.prep_fc_w, c.emb_dim), c.emb_dim, e->w.prep_fc_b, n1 + n2, GSAGE_ACT_NONE,
                               (char*)e->X + n0 * ldx * es, T, ldx, dfe, T == GSAGE_F32, s));
        }
        lvl = RowSrc{e->X, T, ldx, ntot, nullptr, e->d_prep};
    }

    // ---- layer 1 on (x0, x1) and (x1, x2) with shared weights (models.py:85-86) ---------------------------
    const int64_t ldh = e->ld_h1;
    const int O1 = c.out_dim[0], O2 = c.out_dim[1];
    x_seed = lvl;
    if (e->fold_prep) x_seed.ids = e->look0;
    // mean aggregator: layer 1's two applications share their weights, the self rows of all n0 + n1 parents are the prefix
    // [ids0 | ids1] of the id buffer and their reduced rows are adjacent in M -- so both reductions run first and ONE projection
    // launch covers the 26*B parents (one launch, one pipeline fill and one tail instead of two; GSAGE_MERGE_L1=0 keeps them apart)
    static const bool merge_l1 = [] { const char* f = getenv("GSAGE_MERGE_L1"); return !f || atoi(f) != 0; }();
    const bool fused_mean_taken = T == GSAGE_BF16 && !e->keep_activations && getenv("GSAGE_FUSED_LAYER") != nullptr && atoi(getenv("GSAGE_FUSED_LAYER")) != 0;
    if (merge_l1 && c.aggregator == GSAGE_AGG_MEAN && !e->fold_prep && !e->fuse_mean && !fused_mean_taken && e->chunk_parents == 0) {
        const int exact = (T == GSAGE_F32 && !c.allow_tf32);
        const int d = lvl.d;
        const RowSrc nb0 = lvl.shifted(n0), nb1 = lvl.shifted(n0 + n1);
        const int p_app0 = e->prof.begin(GSAGE_PROF_APP0, s);
        GS_TRY(gather_reduce_launch(nb0.base, nb0.dtype, nb0.ld, nb0.table_rows, d, nb0.ids, n0, S1, GSAGE_RED_MEAN, nullptr, e->M, T, e->ld_m, s));
        e->prof.end(p_app0, s);
        const int p_red = e->prof.begin(GSAGE_PROF_REDUCE, s);
        GS_TRY(gather_reduce_launch(nb1.base, nb1.dtype, nb1.ld, nb1.table_rows, d, nb1.ids, n1, S2, GSAGE_RED_MEAN, nullptr,
                                    (char*)e->M + n0 * e->ld_m * es, T, e->ld_m, s));
        e->prof.end(p_red, s);
        e->prof.work(p_red, GSAGE_PROF_REDUCE, (double)n1 * ((double)S2 * d * dtype_size(nb1.dtype) + (nb1.ids ? 8.0 * S2 : 0.0) + (double)d * dtype_size(T)),
                     (double)n1 * S2 * d);
        if (e->ahead_after_gather) {
            if (!e->ev_mid) GS_CUDA(cudaEventCreateWithFlags(&e->ev_mid, cudaEventDisableTiming));
            GS_CUDA(cudaEventRecord(e->ev_mid, s));
            e->mid_valid = true;
        }
        RowSrc m{e->M, T, e->ld_m, n0 + n1, nullptr, d};
        const int p_prj = e->prof.begin(GSAGE_PROF_PROJECT, s);
        GS_TRY(combine_call(lvl, e->w_x[0], m, e->w_n[0], O1, n0 + n1, c.act[0], e->H1, T, ldh, exact, s, e->b_x[0], e->b_n[0]));
        e->prof.end(p_prj, s);
        e->prof.work(p_prj, GSAGE_PROF_PROJECT, (double)(n0 + n1) * ((double)d * dtype_size(lvl.dtype) + (lvl.ids ? 8.0 : 0.0) + (double)d * dtype_size(T) +
                                                                     2.0 * O1 * dtype_size(T)), 4.0 * (double)(n0 + n1) * d * O1);
    } else {
    const int p_app0 = e->prof.begin(GSAGE_PROF_APP0, s);
    GS_TRY(apply_aggregator(e, 0, x_seed, lvl.shifted(n0), n0, S1, e->H1, T, ldh, 0, s));
    e->prof.end(p_app0, s);
    if (!e->keep_activations && e->chunk_parents > 0 && c.aggregator == GSAGE_AGG_MEAN && n1 > e->chunk_parents) {
        for (int64_t p0 = 0; p0 < n1; p0 += e->chunk_parents) {
            const int64_t np = std::min(e->chunk_parents, n1 - p0);
            GS_TRY(apply_aggregator(e, 0, lvl.shifted(n0 + p0), lvl.shifted(n0 + n1 + p0 * S2), np, S2,
                                    (char*)e->H1 + (n0 + p0) * ldh * es, T, ldh, n0, s));      // same M rows every chunk
        }
    } else {
        GS_TRY(apply_aggregator(e, 0, lvl.shifted(n0), lvl.shifted(n0 + n1), n1, S2, (char*)e->H1 + n0 * ldh * es, T, ldh, n0, s));
    }
    }

    // ---- layer 2 on (h0, h1) ----------------------------------------------------------------------------------
    RowSrc h{e->H1, T, ldh, n0 + n1, nullptr, 2 * O1};
    const int p_l2 = e->prof.begin(GSAGE_PROF_LAYER2, s);
    GS_TRY(apply_aggregator(e, 1, h, h.shifted(n0), n0, S1, e->Z, GSAGE_F32, 2 * O2, n0 + n1, s));
    e->prof.end(p_l2, s);

    // ---- normalise + classifier (models.py:90-91) -----------------------------------------------------------------
    const int p_head = e->prof.begin(GSAGE_PROF_HEAD, s);
    if (head_fused_eligible(2 * O2, c.n_classes)) {           // few classes: normalise + classifier as one launch, fp32 FFMA (gather_reduce.cu)
        GS_TRY(head_fused_launch(e->Z, 2 * O2, n0, 2 * O2, e->w.fc_w, e->w.fc_b, c.n_classes, e->ZN, 2 * O2, logits_dev, c.n_classes, s));
        e->prof.end(p_head, s);
        e->prof.end(p_all, s);
        GS_TRY(mark_slot_done(e, s));
        return GSAGE_OK;
    }
    GS_TRY(gsage_l2_normalize(e->Z, GSAGE_F32, 2 * O2, n0, 2 * O2, e->ZN, 2 * O2, s));
    RowSrc zn{e->ZN, GSAGE_F32, 2 * O2, n0, nullptr, 2 * O2};
    if (T == GSAGE_BF16 && e->fc_pad > 0 && (2 * O2) % 4 == 0 && getenv("GSAGE_FC_EXACT") == nullptr) {
        // bf16 mode: the classifier as a TF32 projection on the tensor cores (fp32 operands, 10-bit mantissa products, fp32
        // accumulate -- well inside the mode's bf16 tolerance); the FFMA kernel took 40 us of a 1.5 ms step
        LinearParams P;
        P.n_segs = 1; P.n = n0; P.act = GSAGE_ACT_NONE; P.out = logits_dev; P.out_dtype = GSAGE_F32; P.ld_out = c.n_classes;
        P.seg[0] = LinearSeg{e->ZN, GSAGE_F32, 2 * (int64_t)O2, nullptr, e->FCP, GSAGE_F32, 2 * (int64_t)O2, 2 * O2, e->fc_pad,
                             e->FCP + (size_t)e->fc_pad * 2 * O2, 0};
        P.seg[0].O_store = c.n_classes;
        if (linear_ws_umma_eligible(P)) {
            GS_TRY(linear_ws_umma_launch(P, s));
            e->prof.end(p_head, s);
            e->prof.end(p_all, s);
            GS_TRY(mark_slot_done(e, s));
            return GSAGE_OK;
        }
    }
    GS_TRY(linear_call(zn, f32w(e->w.fc_w, 2 * O2), c.n_classes, e->w.fc_b, n0, GSAGE_ACT_NONE, logits_dev, GSAGE_F32, c.n_classes,
                       0, 1, s));
    e->prof.end(p_head, s);
    e->prof.end(p_all, s);
    GS_TRY(mark_slot_done(e, s));
    return GSAGE_OK;
}

int gsage_engine_forward_host_next(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                                   const int64_t* next_ids_host, int64_t next_B, float* logits_host, void* stream) {
    GS_CHECK_ARG(e && g && rng && ids_host && logits_host, "engine_forward_host: NULL argument");
    GS_CHECK_ARG(B > 0 && B <= e->maxB, "engine_forward_host: batch outside (0, max_batch]");
    GS_CHECK_ARG(!next_ids_host || (next_B > 0 && next_B <= e->maxB), "engine_forward_host: next batch outside (0, max_batch]");
    cudaStream_t s = as_stream(stream);
    if (!e->cs) {
        GS_CUDA(cudaStreamCreateWithFlags(&e->cs, cudaStreamNonBlocking));
        GS_CUDA(cudaEventCreateWithFlags(&e->ev_fwd, cudaEventDisableTiming));
        GS_CUDA(cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming));
    }
    // 1. this batch's forward: already queued by the previous call when that call was given this batch as its `next`
    int buf;
    if (e->pre.valid) {
        GS_CHECK_ARG(e->pre.src == ids_host && e->pre.B == B && e->pre.g == g && e->pre.rng == rng,
                     "engine_forward_host: the forward of a different batch is already queued (the previous call named it as its next batch)");
        buf = e->pre.buf;
        e->pre.valid = false;
    } else {
        buf = (e->lg_cur ^= 1);
        // the seed ids go straight into the hop-0 slot of the id buffer (H2D inside sample_hops, or already done by a sample-ahead)
        GS_TRY(forward_impl(e, g, rng, ids_host, true, B, B, 0, e->LG2[buf], s));
        GS_CUDA(cudaEventRecord(e->ev_fwd, s));
    }
    // 2. its logits (and the two sticky error flags) travel on the copy stream, behind that forward only
    GS_CUDA(cudaStreamWaitEvent(e->cs, e->ev_fwd, 0));
    GS_CUDA(cudaMemcpyAsync(logits_host, e->LG2[buf], 4 * B * e->cfg.n_classes, cudaMemcpyDeviceToHost, e->cs));
    GS_CUDA(cudaEventRecord(e->ev_copy, e->cs));
    // 3. the next batch: H2D of its ids + sampling (sampler stream) AND its whole forward (caller's stream) are queued BEFORE
    //    this call blocks on its own result -- the GPU never waits for the host round trip of the logits
    if (next_ids_host) {
        GS_TRY(sample_ahead_impl(e, g, rng, next_ids_host, true, next_B, next_B, 0, s));
        const int nbuf = buf ^ 1;
        GS_TRY(forward_impl(e, g, rng, next_ids_host, true, next_B, next_B, 0, e->LG2[nbuf], s));
        GS_CUDA(cudaEventRecord(e->ev_fwd, s));
        e->lg_cur = nbuf;
        e->pre.valid = true; e->pre.src = next_ids_host; e->pre.B = next_B; e->pre.buf = nbuf; e->pre.g = g; e->pre.rng = rng;
    }
    // 4. block on THIS batch's copy only
    if (cudaEventSynchronize(e->ev_copy) != cudaSuccess) {
        set_error("engine_forward_host: D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return GSAGE_ERR_CUDA;
    }
    // the sticky flags live in mapped host memory; this batch's kernels have all retired (its logits are here)
    if (*(volatile int*)rng->err_flag) {
        set_error("rng: the look-ahead window held fewer accepted draws than requested (12-sigma event) -- re-seed; results since the last "
                  "check are invalid");
        return GSAGE_ERR_RNG;
    }
    if (*(volatile int*)g->err_flag) return gsage_graph_check(g, stream);     // reports (and clears) the out-of-range id like the device entry
    return GSAGE_OK;
}

int gsage_engine_forward_host(gsage_engine* e, gsage_graph* g, gsage_rng* rng, const int64_t* ids_host, int64_t B,
                              float* logits_host, void* stream) {
    return gsage_engine_forward_host_next(e, g, rng, ids_host, B, nullptr, 0, logits_host, stream);
}

// layer-1 weight gradients run on tcgen05 when everything is bf16 and O1 == 128 (wgrad_umma.cu); fp32 mode stays exact
static bool layer1_wgrad_on_tensor_cores(const gsage_engine* e) {
    if (e->T != GSAGE_BF16 || e->cfg.out_dim[0] != 128 || e->cfg.feats_dtype != GSAGE_BF16) return false;
    WgradJob probe{e->DH, GSAGE_BF16, 2 * (int64_t)e->cfg.out_dim[0], 128, e->cfg.feats_dev, GSAGE_BF16, e->cfg.feats_ld, e->ids,
                   e->cfg.feats_dim, 1, nullptr, e->cfg.feats_dim};
    WgradJob probe2 = probe;
    probe2.A = e->M; probe2.lda = e->ld_m; probe2.ids = nullptr;
    return wgrad_umma_eligible(probe) && wgrad_umma_eligible(probe2);
}

// GSSupervised.forward with the DENSE sampler (nn_modules.py:19-49, train.py:55's default): hop k is
// adj[ids][:, perm_k][:, :S_k] with the caller's torch.randperm(K) draws; everything after the sampling is shared.
int gsage_engine_forward_dense(gsage_engine* e, const int64_t* adj_dev, int64_t n_rows, int K, const int64_t* perm0_dev,
                               const int64_t* perm1_dev, const int64_t* ids_dev, int64_t B, float* logits_dev, void* stream) {
    GS_CHECK_ARG(e && adj_dev && perm0_dev && perm1_dev && ids_dev && logits_dev, "engine_forward_dense: NULL argument");
    GS_CHECK_ARG(e->have_weights, "engine_forward_dense: call gsage_engine_set_weights first");
    GS_CHECK_ARG(B > 0 && B <= e->maxB, "engine_forward_dense: batch %lld outside (0, max_batch=%lld]", (long long)B, (long long)e->maxB);
    GS_CHECK_ARG(!e->ahead.valid, "engine_forward_dense: a sampled-ahead batch is pending (run its forward first)");
    const int S1 = e->cfg.fanout[0], S2 = e->cfg.fanout[1];
    GS_CHECK_ARG(K >= S1 && K >= S2, "engine_forward_dense: fanout larger than the table width K = %d", K);
    cudaStream_t s = as_stream(stream);
    const int p_all = e->prof.begin(GSAGE_PROF_FORWARD, s);
    const int p_smp = e->prof.begin(GSAGE_PROF_SAMPLE, s);
    int64_t* ids0 = e->ids; int64_t* ids1 = ids0 + B; int64_t* ids2 = ids1 + B * S1;
    if (ids_dev != ids0) GS_CUDA(cudaMemcpyAsync(ids0, ids_dev, 8 * B, cudaMemcpyDeviceToDevice, s));
    GS_TRY(gsage_sample_dense(adj_dev, n_rows, K, ids0, B, perm0_dev, S1, ids1, stream));
    GS_TRY(gsage_sample_dense(adj_dev, n_rows, K, ids1, B * S1, perm1_dev, S2, ids2, stream));
    e->prof.end(p_smp, s);
    return forward_layers(e, B, logits_dev, s, p_all);
}

static int backward_supported(gsage_engine* e) {
    GS_CHECK_ARG(e && e->have_weights && e->B > 0, "engine_backward: run gsage_engine_forward first");
    GS_CHECK_ARG(e->cfg.aggregator == GSAGE_AGG_MEAN &&
                 (e->cfg.prep == GSAGE_PREP_IDENTITY || e->cfg.prep == GSAGE_PREP_LINEAR ||
                  (e->cfg.prep == GSAGE_PREP_NODE_EMBEDDING && e->fold_prep && e->T == GSAGE_F32)),
                 "engine_backward: implemented for the mean aggregator with the identity or linear prep, or with the node_embedding "
                 "prep without features in fp32 (the Pokec recipe); the other plug-ins are forward-only");
    GS_CHECK_ARG(!e->fuse_mean, "engine_backward: needs the reduced rows the fused (GSAGE_FUSE_MEAN) layer never writes");
    GS_CHECK_ARG(e->keep_activations || e->chunk_parents == 0, "engine_backward: call gsage_engine_keep_activations(e, 1) before the forward");
    return GSAGE_OK;
}

int gsage_engine_backward_head(gsage_engine* e, const float* dlogits, const gsage_grads* g, void* stream) {
    GS_TRY(backward_supported(e));
    GS_CHECK_ARG(dlogits && g && g->fc_w && g->fc_b && g->fc_x[1] && g->fc_neib[1], "engine_backward_head: NULL argument");
    cudaStream_t s = as_stream(stream);
    const gsage_engine_config& c = e->cfg;
    const int64_t n0 = e->B, n1 = n0 * c.fanout[0];
    const int O1 = c.out_dim[0], O2 = c.out_dim[1], C = c.n_classes, S1 = c.fanout[0];
    const int64_t es = (int64_t)dtype_size(e->T);
    // classifier: logits = zn . Wfc^T + b
    GS_TRY(wgrad_launch(dlogits, C, C, e->ZN, GSAGE_F32, 2 * O2, nullptr, 2 * O2, n0, g->fc_w, 2 * O2, s));
    GS_TRY(colsum_launch(dlogits, n0, C, g->fc_b, s));
    GS_TRY(linear_trans_call(dlogits, C, C, e->w.fc_w, 2 * O2, 2 * O2, n0, e->DZN, 2 * O2, s));
    // F.normalize + layer-2 activation
    const void* m2 = (const char*)e->M + (n0 + n1) * e->ld_m * es;
    WgradJob head_jobs[2] = {
        {e->DZB, GSAGE_BF16, 2 * (int64_t)O2, O2, e->H1, e->T, e->ld_h1, nullptr, 2 * O1, n0, g->fc_x[1], 2 * (int64_t)O1},
        {(const __nv_bfloat16*)e->DZB + O2, GSAGE_BF16, 2 * (int64_t)O2, O2, m2, e->T, e->ld_m, nullptr, 2 * O1, n0, g->fc_neib[1], 2 * (int64_t)O1}};
    const bool head_tc = e->T == GSAGE_BF16 && e->cfg.aggregator == GSAGE_AGG_MEAN && e->w_x2T.p && O2 % 16 == 0 &&
                         wgrad_umma_eligible(head_jobs[0]) && wgrad_umma_eligible(head_jobs[1]);
    GS_TRY(l2_normalize_bwd_launch(e->Z, e->DZN, n0, 2 * O2, c.act[1], e->DZ, s, head_tc ? e->DZB : nullptr));
    // layer 2: z = [h0 Wx2^T | m2 Wn2^T]
    if (head_tc) {
        // bf16 mode: both weight gradients as two jobs of one split-K tcgen05 launch, both data gradients on the projection
        // kernel with the transposed weights (the FFMA generation of these four took 0.26 ms of a 2.4 ms train step)
        GS_TRY(wgrad_umma_launch(head_jobs, 2, s));
        const WRef* wt[2] = {&e->w_x2T, &e->w_n2T};
        float* outs[2] = {e->DH0, e->DM2};
        for (int k = 0; k < 2; ++k) {
            LinearParams P;
            P.n_segs = 1; P.n = n0; P.act = GSAGE_ACT_NONE; P.out = outs[k]; P.out_dtype = GSAGE_F32; P.ld_out = 2 * O1;
            P.seg[0] = LinearSeg{(const __nv_bfloat16*)e->DZB + k * O2, GSAGE_BF16, 2 * (int64_t)O2, nullptr, wt[k]->p, GSAGE_BF16, wt[k]->ld,
                                 O2, 2 * O1, nullptr, 0};
            GS_TRY(linear_dispatch(P, 0, s));
        }
    } else {
        GS_TRY(wgrad_launch(e->DZ, 2 * O2, O2, e->H1, e->T, e->ld_h1, nullptr, 2 * O1, n0, g->fc_x[1], 2 * O1, s));
        GS_TRY(wgrad_launch(e->DZ + O2, 2 * O2, O2, m2, e->T, e->ld_m, nullptr, 2 * O1, n0, g->fc_neib[1], 2 * O1, s));
        GS_TRY(linear_trans_call(e->DZ, 2 * O2, O2, e->w.layer[1].fc_x, 2 * O1, 2 * O1, n0, e->DH0, 2 * O1, s));
        GS_TRY(linear_trans_call(e->DZ + O2, 2 * O2, O2, e->w.layer[1].fc_neib, 2 * O1, 2 * O1, n0, e->DM2, 2 * O1, s));
    }
    // mean over S1 children + concat + layer-1 activation, backwards
    GS_TRY(layer1_grad_launch(e->DH0, e->DM2, e->H1, e->T, e->ld_h1, n0, n1, S1, 2 * O1, c.act[0], e->DH,
                              layer1_wgrad_on_tensor_cores(e) ? GSAGE_BF16 : GSAGE_F32, s));
    return mark_slot_done(e, s);
}

int gsage_engine_backward_layer1(gsage_engine* e, const gsage_grads* g, void* stream) {
    GS_TRY(backward_supported(e));
    GS_CHECK_ARG(e->cfg.prep == GSAGE_PREP_IDENTITY, "engine_backward_layer1: node_embedding models use gsage_engine_backward_layer1_embedding");
    GS_CHECK_ARG(g && g->fc_x[0] && g->fc_neib[0], "engine_backward_layer1: NULL argument");
    cudaStream_t s = as_stream(stream);
    const gsage_engine_config& c = e->cfg;
    const int64_t n0 = e->B, n1 = n0 * c.fanout[0];
    const int O1 = c.out_dim[0], d = c.feats_dim;
    // h = act([table[ids01] Wx1^T | M01 Wn1^T]): both weight gradients reduce over all n0 + n1 parent rows
    if (layer1_wgrad_on_tensor_cores(e)) {
        // bf16 mode: split-K tcgen05 GEMMs straight from the row-major operands (MN-major), self rows gathered by id;
        // the output gradient was written as bf16 by layer1_grad_kernel
        const __nv_bfloat16* dh = (const __nv_bfloat16*)e->DH;
        WgradJob jobs[2] = {
            {dh, GSAGE_BF16, 2 * (int64_t)O1, O1, c.feats_dev, c.feats_dtype, c.feats_ld, e->ids, d, n0 + n1, g->fc_x[0], d, c.feats_rows},
            {dh + O1, GSAGE_BF16, 2 * (int64_t)O1, O1, e->M, e->T, e->ld_m, nullptr, d, n0 + n1, g->fc_neib[0], d}};
        GS_TRY(wgrad_umma_launch(jobs, 2, s));
        return mark_slot_done(e, s);
    }
    GS_TRY(wgrad_launch(e->DH, 2 * O1, O1, c.feats_dev, c.feats_dtype, c.feats_ld, e->ids, d, n0 + n1, g->fc_x[0], d, s));
    GS_TRY(wgrad_launch(e->DH + O1, 2 * O1, O1, e->M, e->T, e->ld_m, nullptr, d, n0 + n1, g->fc_neib[0], d, s));
    return mark_slot_done(e, s);          // the weight gradients gather self rows by id: the slot is busy until here
}

// Layer-1 gradients behind LinearPrep (nn_modules.py:158-166): see gsage_b200.h.  Two reductions against the raw feature rows;
// the neighbour means of the raw rows are gathered here (one more pass of the fused gather+mean kernel over the table).
int gsage_engine_backward_layer1_linear(gsage_engine* e, const gsage_linear_prep_grads* g, void* stream) {
    GS_TRY(backward_supported(e));
    GS_CHECK_ARG(e->cfg.prep == GSAGE_PREP_LINEAR && e->MF, "engine_backward_layer1_linear: not a mean + LinearPrep model");
    GS_CHECK_ARG(g && g->gx_raw && g->gn_raw, "engine_backward_layer1_linear: NULL argument");
    cudaStream_t s = as_stream(stream);
    const gsage_engine_config& c = e->cfg;
    const int64_t n0 = e->B, n1 = n0 * c.fanout[0];
    const int O1 = c.out_dim[0], d = c.feats_dim, S1 = c.fanout[0], S2 = c.fanout[1];
    const int64_t es = (int64_t)dtype_size(e->T);
    const int64_t* ids1 = e->ids + n0; const int64_t* ids2 = ids1 + n1;
    GS_TRY(gather_reduce_launch(c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, d, ids1, n0, S1, GSAGE_RED_MEAN, nullptr, e->MF, e->T,
                                c.feats_ld, s));
    GS_TRY(gather_reduce_launch(c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, d, ids2, n1, S2, GSAGE_RED_MEAN, nullptr,
                                (char*)e->MF + n0 * c.feats_ld * es, e->T, c.feats_ld, s));
    if (layer1_wgrad_on_tensor_cores(e)) {               // the head wrote G as bf16 (layer1_grad_kernel)
        const __nv_bfloat16* dh = (const __nv_bfloat16*)e->DH;
        WgradJob jobs[2] = {
            {dh, GSAGE_BF16, 2 * (int64_t)O1, O1, c.feats_dev, c.feats_dtype, c.feats_ld, e->ids, d, n0 + n1, g->gx_raw, d, c.feats_rows},
            {dh + O1, GSAGE_BF16, 2 * (int64_t)O1, O1, e->MF, e->T, c.feats_ld, nullptr, d, n0 + n1, g->gn_raw, d}};
        GS_CHECK_ARG(wgrad_umma_eligible(jobs[0]) && wgrad_umma_eligible(jobs[1]), "engine_backward_layer1_linear: operands do not qualify for the tensor-core weight gradient");
        GS_TRY(wgrad_umma_launch(jobs, 2, s));
        return mark_slot_done(e, s);
    }
    GS_TRY(wgrad_launch(e->DH, 2 * O1, O1, c.feats_dev, c.feats_dtype, c.feats_ld, e->ids, d, n0 + n1, g->gx_raw, d, s));
    GS_TRY(wgrad_launch(e->DH + O1, 2 * O1, O1, e->MF, e->T, c.feats_ld, nullptr, d, n0 + n1, g->gn_raw, d, s));
    return mark_slot_done(e, s);
}

// Layer-1 gradients of the Pokec recipe (mean aggregator, NodeEmbeddingPrep without features, nn_modules.py:126-155):
//   H_r = act([W'x e_self(r) + b'x | W'n mean_j e_nb(r,j) + b'n]),  W' = W.Wp, b' = W.bp  (the prep's affine folded into layer 1)
// With G = d loss / d pre-activation (layer1_grad_kernel):
//   gx_raw = Gx^T . E[self ids]     gn_raw = Gn^T . M     cx / cn = column sums of Gx / Gn       (all the row-reductions)
//   d_table[id] += Gx[r] . W'x  for the self row of r,   += (1/S) Gn[r] . W'n  for each of its S sampled neighbours
// The parameter gradients follow from these by (O x 64)(64 x 64) products, left to the caller:
//   dWx = gx_raw.Wp^T + cx (x) bp,  dWn = gn_raw.Wp^T + cn (x) bp,  dWp = Wx^T.gx_raw + Wn^T.gn_raw,  dbp = Wx^T.cx + Wn^T.cn
int gsage_engine_backward_layer1_embedding(gsage_engine* e, const gsage_embedding_grads* g, void* stream) {
    GS_TRY(backward_supported(e));
    GS_CHECK_ARG(e->cfg.prep == GSAGE_PREP_NODE_EMBEDDING && e->fold_prep, "engine_backward_layer1_embedding: not a node_embedding model");
    GS_CHECK_ARG(g && g->gx_raw && g->gn_raw && g->csum && g->d_table, "engine_backward_layer1_embedding: NULL argument");
    cudaStream_t s = as_stream(stream);
    const gsage_engine_config& c = e->cfg;
    const int64_t n0 = e->B, n1 = n0 * c.fanout[0], n2 = n1 * c.fanout[1];
    const int O1 = c.out_dim[0], de = c.emb_dim, S1 = c.fanout[0], S2 = c.fanout[1];
    const int64_t rows = n0 + n1, trows = c.n_nodes + 1;
    const int64_t* ids0 = e->ids; const int64_t* ids1 = ids0 + n0; const int64_t* ids2 = ids1 + n1;
    (void)ids0;
    // row reductions: seeds read the masked row n_nodes (look0), hop-1 parents their own id
    GS_TRY(wgrad_launch(e->DH, 2 * O1, O1, c.emb_dev, c.emb_dtype, c.emb_ld, e->look0, de, n0, g->gx_raw, de, s));
    GS_TRY(wgrad_launch(e->DH + (int64_t)n0 * 2 * O1, 2 * O1, O1, c.emb_dev, c.emb_dtype, c.emb_ld, ids1, de, n1, g->gx_raw, de, s, true));
    GS_TRY(wgrad_launch(e->DH + O1, 2 * O1, O1, e->M, e->T, e->ld_m, nullptr, de, rows, g->gn_raw, de, s));
    GS_TRY(colsum_launch(e->DH, rows, 2 * O1, g->csum, s));
    // table gradient: dense (n_nodes + 1, de), like nn.Embedding's
    GS_CUDA(cudaMemsetAsync(g->d_table, 0, sizeof(float) * (size_t)trows * de, s));
    GS_TRY(linear_trans_call(e->DH, 2 * O1, O1, e->fold_wx, de, de, rows, e->DXE, de, s));                 // Gx . W'x
    GS_TRY(embedding_scatter_launch(e->DXE, de, de, e->look0, n0, 1, 1.0f, g->d_table, de, trows, s));
    GS_TRY(embedding_scatter_launch(e->DXE + n0 * de, de, de, ids1, n1, 1, 1.0f, g->d_table, de, trows, s));
    GS_TRY(linear_trans_call(e->DH + O1, 2 * O1, O1, e->fold_wn, de, de, rows, e->DXE, de, s));            // Gn . W'n
    GS_TRY(embedding_scatter_launch(e->DXE, de, de, ids1, n1, S1, 1.0f / (float)S1, g->d_table, de, trows, s));
    GS_TRY(embedding_scatter_launch(e->DXE + n0 * de, de, de, ids2, n2, S2, 1.0f / (float)S2, g->d_table, de, trows, s));
    return mark_slot_done(e, s);
}

// Full parameter-gradient pass for the pool aggregators (bf16 compute, identity prep; nn_modules.py:207-256).
//   out = act([Wx x | Wn p]),  p = pool_j relu(W1 n_j + b1)
// Per application, from G = d loss / d pre-activation:  dWx = Gx^T x,  dWn = Gn^T p,  dP = Gn Wn,
//   dHid = pool'(dP) -- the forward's MLP is RECOMPUTED on the tensor cores and the gradient routed to the arg-max row (or
//   spread over the open relus for the mean pool) by linear_pool_ws_umma_kernel<.., BWD>; it is the one large intermediate
//   (rows x H bf16) --  dW1 = dHid^T n (wgrad_umma, 128-unit blocks as parallel jobs),  db1 = column sums (in the same kernel),
//   and for layer 2 the input gradient dN = dHid W1 that flows into layer 1.
static int backward_pool_impl(gsage_engine* e, const float* dlogits, const gsage_grads* g, const gsage_pool_grads* pg,
                              const gsage_pool_embedding_grads* eg, void* stream) {
    GS_CHECK_ARG(e && e->have_weights && e->B > 0, "engine_backward_pool: run gsage_engine_forward first");
    const gsage_engine_config& c = e->cfg;
    const bool pool = c.aggregator == GSAGE_AGG_MAX_POOL || c.aggregator == GSAGE_AGG_MEAN_POOL;
    const bool fold = e->fold_prep;
    GS_CHECK_ARG(pool && (c.prep == GSAGE_PREP_IDENTITY || fold) && e->T == GSAGE_BF16 && e->DP,
                 "engine_backward_pool: implemented for the max / mean pool aggregators in bf16 compute mode, with the identity prep or "
                 "the node_embedding prep without features");
    GS_CHECK_ARG(fold == (eg != nullptr), fold ? "engine_backward_pool: node_embedding models use gsage_engine_backward_pool_embedding"
                                               : "engine_backward_pool_embedding: not a node_embedding model");
    if (eg) GS_CHECK_ARG(eg->csum_x && eg->d_table && c.emb_dtype == GSAGE_BF16, "engine_backward_pool_embedding: NULL argument / table not bf16");
    GS_CHECK_ARG(e->keep_activations, "engine_backward_pool: call gsage_engine_keep_activations(e, 1) before the forward");
    GS_CHECK_ARG(dlogits && g && pg && g->fc_w && g->fc_b && g->fc_x[0] && g->fc_x[1] && g->fc_neib[0] && g->fc_neib[1] &&
                 pg->mlp_w[0] && pg->mlp_w[1] && pg->mlp_b[0] && pg->mlp_b[1], "engine_backward_pool: NULL argument");
    const int O1 = c.out_dim[0], O2 = c.out_dim[1], C = c.n_classes, S1 = c.fanout[0], S2 = c.fanout[1], H = e->hid;
    // layer 1 reads rows of the feature table, or (folded node_embedding prep) of the raw embedding table
    const void* tab = fold ? c.emb_dev : c.feats_dev;
    const int64_t tab_ld = fold ? c.emb_ld : c.feats_ld, tab_rows = fold ? c.n_nodes + 1 : c.feats_rows;
    const int d = fold ? c.emb_dim : c.feats_dim;
    GS_CHECK_ARG(O1 == 128 && O2 == 128 && H % 128 == 0, "engine_backward_pool: needs output_dim 128 and a hidden width that is a multiple of 128 "
                 "(the tensor-core weight-gradient kernel works on 128-row blocks)");
    cudaStream_t s = as_stream(stream);
    const int64_t n0 = e->B, n1 = n0 * S1, n2 = n1 * S2, rows = n0 + n1;
    const int pool_max = c.aggregator == GSAGE_AGG_MAX_POOL ? 1 : 0;
    const __nv_bfloat16* H1 = (const __nv_bfloat16*)e->H1;
    const __nv_bfloat16* Pp = (const __nv_bfloat16*)e->Pp;
    __nv_bfloat16* DHID = (__nv_bfloat16*)e->DHID;
    const int64_t* ids0 = e->ids; const int64_t* ids1 = ids0 + n0; const int64_t* ids2 = ids1 + n1;

    auto pool_bwd = [&](int layer, const void* a, int64_t lda, const int64_t* ids, int d_in, int64_t n_rows, int S, const float* dP,
                        __nv_bfloat16* dhid, float* db) -> int {
        LinearParams P;
        P.n_segs = 1; P.n = n_rows; P.act = GSAGE_ACT_RELU; P.out = nullptr; P.out_dtype = GSAGE_BF16; P.ld_out = H;
        P.seg[0] = LinearSeg{a, GSAGE_BF16, lda, ids, e->w_mlp[layer].p, GSAGE_BF16, e->w_mlp[layer].ld, d_in, H, e->b_mlp[layer], 0};
        P.pool_S = S; P.pool_max = pool_max;
        P.seg[0].a_rows = ids ? tab_rows : 0;
        return linear_pool_ws_umma_backward_launch(P, dP, H, dhid, H, db, s);
    };
    auto mlp_wgrad = [&](const __nv_bfloat16* dhid, const void* a, int64_t lda, const int64_t* ids, int d_in, int64_t n_rows, float* dW) -> int {
        for (int b0 = 0; b0 < H / 128; b0 += 4) {                      // 128-unit blocks of the hidden layer as parallel jobs
            WgradJob jobs[4];
            const int nj = std::min(4, H / 128 - b0);
            for (int j = 0; j < nj; ++j)
                jobs[j] = WgradJob{dhid + (b0 + j) * 128, GSAGE_BF16, (int64_t)H, 128, a, GSAGE_BF16, lda, ids, d_in, n_rows,
                                   dW + (int64_t)(b0 + j) * 128 * d_in, (int64_t)d_in, ids ? tab_rows : 0};
            GS_TRY(wgrad_umma_launch(jobs, nj, s));
        }
        return GSAGE_OK;
    };

    // ---- classifier + F.normalize (as gsage_engine_backward_head) --------------------------------------------------
    GS_TRY(wgrad_launch(dlogits, C, C, e->ZN, GSAGE_F32, 2 * O2, nullptr, 2 * O2, n0, g->fc_w, 2 * O2, s));
    GS_TRY(colsum_launch(dlogits, n0, C, g->fc_b, s));
    GS_TRY(linear_trans_call(dlogits, C, C, e->w.fc_w, 2 * O2, 2 * O2, n0, e->DZN, 2 * O2, s));
    GS_TRY(l2_normalize_bwd_launch(e->Z, e->DZN, n0, 2 * O2, c.act[1], e->DZ, s));
    // ---- layer 2 on (h0, h1): x = H1[:n0], neighbours = H1[n0:] in place, pooled rows P2 ----------------------------
    const __nv_bfloat16* P2 = Pp + (n0 + n1) * (int64_t)H;
    GS_TRY(wgrad_launch(e->DZ, 2 * O2, O2, H1, GSAGE_BF16, e->ld_h1, nullptr, 2 * O1, n0, g->fc_x[1], 2 * O1, s));
    GS_TRY(wgrad_launch(e->DZ + O2, 2 * O2, O2, P2, GSAGE_BF16, H, nullptr, H, n0, g->fc_neib[1], H, s));
    GS_TRY(linear_trans_call(e->DZ, 2 * O2, O2, e->w.layer[1].fc_x, 2 * O1, 2 * O1, n0, e->DH0, 2 * O1, s));      // d h0
    GS_TRY(linear_trans_call(e->DZ + O2, 2 * O2, O2, e->w.layer[1].fc_neib, H, H, n0, e->DP, H, s));              // d P2
    GS_CUDA(cudaMemsetAsync(pg->mlp_b[1], 0, sizeof(float) * H, s));
    GS_TRY(pool_bwd(1, H1 + n0 * e->ld_h1, e->ld_h1, nullptr, 2 * O1, n1, S1, e->DP, DHID, pg->mlp_b[1]));
    GS_TRY(mlp_wgrad(DHID, H1 + n0 * e->ld_h1, e->ld_h1, nullptr, 2 * O1, n1, pg->mlp_w[1]));
    {   // d H1[n0:] = dHid2 . W1_2  (n1 x H -> 2*O1): the projection kernel on the transposed weights
        LinearParams P;
        P.n_segs = 1; P.n = n1; P.act = GSAGE_ACT_NONE; P.out = e->DN2; P.out_dtype = GSAGE_F32; P.ld_out = 2 * O1;
        P.seg[0] = LinearSeg{DHID, GSAGE_BF16, (int64_t)H, nullptr, e->w_mlpT[1].p, GSAGE_BF16, e->w_mlpT[1].ld, H, 2 * O1, nullptr, 0};
        GS_TRY(linear_dispatch(P, 0, s));
    }
    // d (layer-1 pre-activation): [d h0 ; d H1[n0:]] * relu'(H1), bf16 (operand of the tensor-core weight gradients)
    GS_TRY(layer1_grad_launch(e->DH0, e->DN2, e->H1, e->T, e->ld_h1, n0, n1, 1, 2 * O1, c.act[0], e->DH, GSAGE_BF16, s));
    // ---- layer 1 on (x0, x1) and (x1, x2), shared weights ------------------------------------------------------------------
    const __nv_bfloat16* dh = (const __nv_bfloat16*)e->DH;
    {
        if (!fold) {
            WgradJob jx{dh, GSAGE_BF16, 2 * (int64_t)O1, O1, tab, GSAGE_BF16, tab_ld, ids0, d, rows, g->fc_x[0], (int64_t)d, tab_rows};
            GS_CHECK_ARG(wgrad_umma_eligible(jx), "engine_backward_pool: the feature table does not qualify for the tensor-core weight gradient");
            GS_TRY(wgrad_umma_launch(&jx, 1, s));
        } else {
            // seeds read the masked row n_nodes (look0), hop-1 parents their own id: two launches into the same buffer
            WgradJob j0{dh, GSAGE_BF16, 2 * (int64_t)O1, O1, tab, GSAGE_BF16, tab_ld, e->look0, d, n0, g->fc_x[0], (int64_t)d, tab_rows};
            WgradJob j1{dh + n0 * 2 * (int64_t)O1, GSAGE_BF16, 2 * (int64_t)O1, O1, tab, GSAGE_BF16, tab_ld, ids1, d, n1, g->fc_x[0], (int64_t)d, tab_rows};
            GS_CHECK_ARG(wgrad_umma_eligible(j0), "engine_backward_pool: the embedding table does not qualify for the tensor-core weight gradient");
            GS_TRY(wgrad_umma_launch(&j0, 1, s));
            GS_TRY(wgrad_umma_launch(&j1, 1, s, true));
        }
        WgradJob jn{dh + O1, GSAGE_BF16, 2 * (int64_t)O1, O1, Pp, GSAGE_BF16, (int64_t)H, nullptr, H, rows, g->fc_neib[0], (int64_t)H};
        GS_TRY(wgrad_umma_launch(&jn, 1, s));
    }
    {   // d P1 = Gn . Wn1  ((n0 + n1) x O1 -> H)
        LinearParams P;
        P.n_segs = 1; P.n = rows; P.act = GSAGE_ACT_NONE; P.out = e->DP; P.out_dtype = GSAGE_F32; P.ld_out = H;
        P.seg[0] = LinearSeg{dh + O1, GSAGE_BF16, 2 * (int64_t)O1, nullptr, e->w_nT[0].p, GSAGE_BF16, e->w_nT[0].ld, O1, H, nullptr, 0};
        GS_TRY(linear_dispatch(P, 0, s));
    }
    GS_CUDA(cudaMemsetAsync(pg->mlp_b[0], 0, sizeof(float) * H, s));
    GS_TRY(pool_bwd(0, tab, tab_ld, ids1, d, n1, S1, e->DP, DHID, pg->mlp_b[0]));                              // (x0, x1)
    GS_TRY(pool_bwd(0, tab, tab_ld, ids2, d, n2, S2, e->DP + n0 * (int64_t)H, DHID + n1 * (int64_t)H, pg->mlp_b[0]));   // (x1, x2)
    GS_TRY(mlp_wgrad(DHID, tab, tab_ld, ids1, d, n1 + n2, pg->mlp_w[0]));
    if (fold) {
        // With the prep folded in (W1' = W1.Wp, Wx' = Wx.Wp), fc_x[0] / mlp_w[0] above are the RAW reductions against the
        // embedding rows (the caller unfolds them, see model.GSSupervised.backward); here the rest: column sums of Gx and
        // the dense gradient of the embedding table = every row's d loss / d (raw embedding), scatter-added by id
        GS_TRY(colsum_bf16_launch(dh, 2 * O1, rows, O1, eg->csum_x, s));
        GS_CUDA(cudaMemsetAsync(eg->d_table, 0, sizeof(float) * (size_t)tab_rows * d, s));
        LinearParams P;
        P.n_segs = 1; P.act = GSAGE_ACT_NONE; P.out = e->DXE; P.out_dtype = GSAGE_F32; P.ld_out = d;
        P.n = n1 + n2;                                                       // neighbours: dHid . W1'
        P.seg[0] = LinearSeg{DHID, GSAGE_BF16, (int64_t)H, nullptr, e->w_mlpT[0].p, GSAGE_BF16, e->w_mlpT[0].ld, H, d, nullptr, 0};
        GS_TRY(linear_dispatch(P, 0, s));
        GS_TRY(embedding_scatter_launch(e->DXE, d, d, ids1, n1 + n2, 1, 1.0f, eg->d_table, d, tab_rows, s));
        P.n = rows;                                                          // self rows: Gx . Wx'
        P.seg[0] = LinearSeg{dh, GSAGE_BF16, 2 * (int64_t)O1, nullptr, e->w_xT0.p, GSAGE_BF16, e->w_xT0.ld, O1, d, nullptr, 0};
        GS_TRY(linear_dispatch(P, 0, s));
        GS_TRY(embedding_scatter_launch(e->DXE, d, d, e->look0, n0, 1, 1.0f, eg->d_table, d, tab_rows, s));
        GS_TRY(embedding_scatter_launch(e->DXE + n0 * (int64_t)d, d, d, ids1, n1, 1, 1.0f, eg->d_table, d, tab_rows, s));
    }
    return mark_slot_done(e, s);
}

int gsage_engine_backward_pool(gsage_engine* e, const float* dlogits, const gsage_grads* g, const gsage_pool_grads* pg, void* stream) {
    return backward_pool_impl(e, dlogits, g, pg, nullptr, stream);
}

int gsage_engine_backward_pool_embedding(gsage_engine* e, const float* dlogits, const gsage_grads* g, const gsage_pool_grads* pg,
                                         const gsage_pool_embedding_grads* eg, void* stream) {
    GS_CHECK_ARG(eg, "engine_backward_pool_embedding: NULL argument");
    return backward_pool_impl(e, dlogits, g, pg, eg, stream);
}

// Full parameter-gradient pass for the attention aggregator (bf16 compute, identity prep; nn_modules.py:289-321).
//   out = act([Wx x | Wn m]),  m_p = sum_j w_pj n_pj,  w_p = softmax_j <a(n_pj), a(x_p)>,  a(v) = W2 tanh(W1 v)
// Per application the attention MLP's intermediates (tanh outputs, a(n), a(x), softmax weights) are RECOMPUTED with the
// unfused forward kernels (the fused forward kernel keeps them on chip), then:
//   dWx = Gx^T x, dWn = Gn^T m (tcgen05 split-K);  dM = Gn Wn (projection kernel, transposed weights);
//   dw_pj = <dM_p, n_pj> (one more pass over the neighbour rows);  softmax', da(n) = ds xa, da(x) = sum_j ds a(n_pj);
//   through W2 and tanh:  dW2 += da^T t1,  dpre = (da W2)(1 - t1^2),  dW1 += dpre^T rows  -- for the neighbour and the self side.
// Layer 2 also returns the gradient of its input rows (w dM + dpre W1 for neighbours, Gx Wx + dpre_x W1 for self rows).
int gsage_engine_backward_attention(gsage_engine* e, const float* dlogits, const gsage_grads* g, const gsage_attention_grads* ag, void* stream) {
    GS_CHECK_ARG(e && e->have_weights && e->B > 0, "engine_backward_attention: run gsage_engine_forward first");
    const gsage_engine_config& c = e->cfg;
    GS_CHECK_ARG(c.aggregator == GSAGE_AGG_ATTENTION && c.prep == GSAGE_PREP_IDENTITY && e->T == GSAGE_BF16 && e->ADM,
                 "engine_backward_attention: implemented for the attention aggregator with the identity prep in bf16 compute mode");
    GS_CHECK_ARG(e->keep_activations, "engine_backward_attention: call gsage_engine_keep_activations(e, 1) before the forward");
    GS_CHECK_ARG(dlogits && g && ag && g->fc_w && g->fc_b && g->fc_x[0] && g->fc_x[1] && g->fc_neib[0] && g->fc_neib[1] &&
                 ag->att_w1[0] && ag->att_w1[1] && ag->att_w2[0] && ag->att_w2[1], "engine_backward_attention: NULL argument");
    const int O1 = c.out_dim[0], O2 = c.out_dim[1], C = c.n_classes, S1 = c.fanout[0], S2 = c.fanout[1], H = e->hid, d1 = c.feats_dim;
    GS_CHECK_ARG(O1 == 128 && O2 == 128 && H <= 32 && d1 % 16 == 0,
                 "engine_backward_attention: needs output_dim 128, attention width <= 32 and a feature width that is a multiple of 16");
    cudaStream_t s = as_stream(stream);
    const int64_t n0 = e->B, n1 = n0 * S1, n2 = n1 * S2;
    const int64_t es = (int64_t)dtype_size(e->T);
    const int64_t* ids0 = e->ids; const int64_t* ids1 = ids0 + n0; const int64_t* ids2 = ids1 + n1;
    (void)n2;

    // one aggregator application, backwards.  G: (n, 2*O) bf16 pre-activation gradient.  `accumulate`: the layer's weights
    // are shared with an earlier application.  in_self / in_nb (layer 2): where the two halves of the input-row gradient go.
    auto app_bwd = [&](int layer, const RowSrc& x, const RowSrc& nb, int64_t n, int S, const void* Mrows, const __nv_bfloat16* G, bool accumulate,
                       float* in_self_a, float* in_self_b, float* in_nb_a, float* in_nb_b) -> int {
        const gsage_layer_weights& L = e->w.layer[layer];
        const int O = c.out_dim[layer], d = x.d;
        const int64_t ldg = 2 * (int64_t)O;
        // recompute a(n), a(x), softmax weights (the unfused forward chain of apply_aggregator)
        GS_TRY(linear_call(nb, e->w_att1[layer], H, e->b_att[layer], n * S, GSAGE_ACT_TANH, e->T1, GSAGE_F32, H, 0, 0, s));
        RowSrc t1{e->T1, GSAGE_F32, H, n * S, nullptr, H};
        // (the 32 x 32 products run as TF32 on the tensor cores here: n*S rows each, the FFMA kernel spent 1.6 ms on them)
        GS_TRY(linear_call(t1, f32w(L.att_w2, H), H, nullptr, n * S, GSAGE_ACT_NONE, e->NA, GSAGE_F32, H, 0, 0, s));
        GS_TRY(linear_call(x, e->w_att1[layer], H, e->b_att[layer], n, GSAGE_ACT_TANH, e->T1x, GSAGE_F32, H, 0, 0, s));
        RowSrc t1x{e->T1x, GSAGE_F32, H, n, nullptr, H};
        GS_TRY(linear_call(t1x, f32w(L.att_w2, H), H, nullptr, n, GSAGE_ACT_NONE, e->XA, GSAGE_F32, H, 0, 0, s));
        GS_TRY(gsage_attention_weights(e->NA, e->XA, GSAGE_F32, H, H, n, S, e->AW, s));
        // projection weights
        WgradJob jobs[2] = {
            {G, GSAGE_BF16, ldg, O, x.base, GSAGE_BF16, x.ld, x.ids, d, n, g->fc_x[layer], (int64_t)d, x.ids ? x.table_rows : 0},
            {G + O, GSAGE_BF16, ldg, O, Mrows, GSAGE_BF16, e->ld_m, nullptr, d, n, g->fc_neib[layer], (int64_t)d, 0}};
        GS_CHECK_ARG(wgrad_umma_eligible(jobs[0]) && wgrad_umma_eligible(jobs[1]), "engine_backward_attention: operands do not qualify for the "
                     "tensor-core weight gradient");
        GS_TRY(wgrad_umma_launch(jobs, 2, s, accumulate));
        {   // dM = Gn . Wn  (n x O -> d)
            LinearParams P;
            P.n_segs = 1; P.n = n; P.act = GSAGE_ACT_NONE; P.out = e->ADM; P.out_dtype = GSAGE_F32; P.ld_out = e->ld_m;
            P.seg[0] = LinearSeg{G + O, GSAGE_BF16, ldg, nullptr, e->w_nT[layer].p, GSAGE_BF16, e->w_nT[layer].ld, O, d, nullptr, 0};
            GS_TRY(linear_dispatch(P, 0, s));
        }
        GS_TRY(attention_dw_launch(nb.base, nb.ld, nb.ids ? nb.table_rows : n * S, d, nb.ids, n, S, e->ADM, e->ld_m, e->AW, e->ADW,
                                   in_nb_a, d, s));
        GS_TRY(attention_softmax_bwd_launch(e->AW, e->ADW, (const float*)e->NA, (const float*)e->XA, H, n, S, e->ADA, e->ADXA, s));
        // neighbour side of the attention MLP
        GS_TRY(wgrad_launch(e->ADA, H, H, e->T1, GSAGE_F32, H, nullptr, H, n * S, ag->att_w2[layer], H, s, accumulate));       // dW2 += dA^T t1
        {   // dA . W2 = dA . (W2^T)^T: a plain projection with the transposed copy (TF32 on the tensor cores)
            RowSrc da{e->ADA, GSAGE_F32, H, n * S, nullptr, H};
            GS_TRY(linear_call(da, f32w(e->AW2T[layer], H), H, nullptr, n * S, GSAGE_ACT_NONE, e->ADT1, GSAGE_F32, H, 0, 0, s));
        }
        GS_TRY(tanh_bwd_launch(e->ADT1, (const float*)e->T1, n * S * H, s, H, e->ADPB));
        {   // dW1 += dpre^T n over all n*S neighbour rows: the 32 gradient columns ride in a zero-padded 128-column bf16 operand so
            // that the split-K tcgen05 kernel takes it (the FFMA kernel needed 3 ms for this product at B = 8192)
            WgradJob jw{e->ADPB, GSAGE_BF16, 128, 128, nb.base, GSAGE_BF16, nb.ld, nb.ids, d, n * S, e->AW1S, (int64_t)d, nb.ids ? nb.table_rows : 0};
            if (wgrad_umma_eligible(jw)) {
                GS_TRY(wgrad_umma_launch(&jw, 1, s));
                if (accumulate) { axpy_kernel<<<(unsigned)ceil_div((int64_t)H * d, 256), 256, 0, s>>>(ag->att_w1[layer], e->AW1S, H * d); GS_LAUNCHED(); }
                else GS_CUDA(cudaMemcpyAsync(ag->att_w1[layer], e->AW1S, sizeof(float) * (size_t)H * d, cudaMemcpyDeviceToDevice, s));
            } else {
                GS_TRY(wgrad_launch(e->ADT1, H, H, nb.base, nb.dtype, nb.ld, nb.ids, d, n * S, ag->att_w1[layer], d, s, accumulate));
            }
        }
        // self side
        GS_TRY(wgrad_launch(e->ADXA, H, H, e->T1x, GSAGE_F32, H, nullptr, H, n, ag->att_w2[layer], H, s, true));
        GS_TRY(linear_trans_call(e->ADXA, H, H, L.att_w2, H, H, n, e->ADT1X, H, s));
        GS_TRY(tanh_bwd_launch(e->ADT1X, (const float*)e->T1x, n * H, s));
        GS_TRY(wgrad_launch(e->ADT1X, H, H, x.base, x.dtype, x.ld, x.ids, d, n, ag->att_w1[layer], d, s, true));
        if (in_self_a) {
            // gradient of the input rows (layer 2): self = Gx . Wx + dpre_x . W1;  neighbours = w dM (written by attention_dw) + dpre . W1
            LinearParams P;
            P.n_segs = 1; P.n = n; P.act = GSAGE_ACT_NONE; P.out = in_self_a; P.out_dtype = GSAGE_F32; P.ld_out = d;
            P.seg[0] = LinearSeg{G, GSAGE_BF16, ldg, nullptr, e->w_x2T.p, GSAGE_BF16, e->w_x2T.ld, O, d, nullptr, 0};
            GS_TRY(linear_dispatch(P, 0, s));
            GS_TRY(linear_trans_call(e->ADT1X, H, H, L.att_w1, d, d, n, in_self_b, d, s));
            GS_TRY(linear_trans_call(e->ADT1, H, H, L.att_w1, d, d, n * S, in_nb_b, d, s));
        }
        return GSAGE_OK;
    };

    // ---- classifier + F.normalize --------------------------------------------------------------------------------------
    GS_TRY(wgrad_launch(dlogits, C, C, e->ZN, GSAGE_F32, 2 * O2, nullptr, 2 * O2, n0, g->fc_w, 2 * O2, s));
    GS_TRY(colsum_launch(dlogits, n0, C, g->fc_b, s));
    GS_TRY(linear_trans_call(dlogits, C, C, e->w.fc_w, 2 * O2, 2 * O2, n0, e->DZN, 2 * O2, s));
    GS_TRY(l2_normalize_bwd_launch(e->Z, e->DZN, n0, 2 * O2, c.act[1], e->DZ, s, e->DZB));
    // ---- layer 2 on (h0, h1) -----------------------------------------------------------------------------------------------
    RowSrc h{e->H1, e->T, e->ld_h1, n0 + n1, nullptr, 2 * O1};
    const void* M2 = (const char*)e->M + (n0 + n1) * e->ld_m * es;
    GS_TRY(app_bwd(1, h, h.shifted(n0), n0, S1, M2, (const __nv_bfloat16*)e->DZB, false, e->ASA, e->ASB, e->ANA, e->ANB));
    GS_TRY(sum_act_grad_launch(e->ASA, e->ASB, e->ANA, e->ANB, e->H1, e->T, e->ld_h1, n0, n1, 2 * O1, c.act[0], e->DH, GSAGE_BF16, s));
    // ---- layer 1 on (x0, x1) and (x1, x2), shared weights -----------------------------------------------------------------------
    const __nv_bfloat16* dh = (const __nv_bfloat16*)e->DH;
    RowSrc lvl{c.feats_dev, c.feats_dtype, c.feats_ld, c.feats_rows, ids0, d1};
    GS_TRY(app_bwd(0, lvl, lvl.shifted(n0), n0, S1, e->M, dh, false, nullptr, nullptr, nullptr, nullptr));
    GS_TRY(app_bwd(0, lvl.shifted(n0), lvl.shifted(n0 + n1), n1, S2, (const char*)e->M + n0 * e->ld_m * es, dh + n0 * 2 * (int64_t)O1, true,
                   nullptr, nullptr, nullptr, nullptr));
    (void)ids1; (void)ids2;
    return mark_slot_done(e, s);
}

int gsage_engine_peek(gsage_engine* e, int what, const void** ptr, int64_t* rows, int64_t* cols, int64_t* ld, int* dtype) {
    GS_CHECK_ARG(e && ptr && rows && cols && ld && dtype, "engine_peek: NULL argument");
    const int64_t B = e->B, n1 = B * e->cfg.fanout[0], n2 = n1 * e->cfg.fanout[1];
    switch (what) {
    case 0: *ptr = e->ids; *rows = B; *cols = 1; *ld = 1; *dtype = -1; return GSAGE_OK;
    case 1: *ptr = e->ids + B; *rows = n1; *cols = 1; *ld = 1; *dtype = -1; return GSAGE_OK;
    case 2: *ptr = e->ids + B + n1; *rows = n2; *cols = 1; *ld = 1; *dtype = -1; return GSAGE_OK;
    case 10: *ptr = e->H1; *rows = B + n1; *cols = 2 * e->cfg.out_dim[0]; *ld = e->ld_h1; *dtype = e->T; return GSAGE_OK;
    case 11: *ptr = e->Z; *rows = B; *cols = 2 * e->cfg.out_dim[1]; *ld = 2 * e->cfg.out_dim[1]; *dtype = GSAGE_F32; return GSAGE_OK;
    }
    set_error("engine_peek: unknown view %d", what);
    return GSAGE_ERR_INVALID;
}

}  // extern "C"
