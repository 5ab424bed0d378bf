// graph.cu -- device-resident adjacency: ingest of the reference's sparse problem format.
//
// Replaces  parse_csr_matrix                        (/root/reference/problem.py:70-72)
//           SparseUniformNeighborSampler.__init__   (/root/reference/nn_modules.py:72-78)
//
// HBM layout
//   indptr  int64 [n_rows + 1]
//   val     int32 [nnz]  (int64 when a stored value does not fit) -- the k-th stored value of the row;
//                        in the reference's file convention this IS the k-th neighbour (+1 id space)
//   "fast" graphs (every row's columns are exactly 0..len-1 and no stored zero; utils/convert.py:100-126)
//           keep nothing else: degree = indptr[r+1]-indptr[r] comes from the same 16-byte pair.
//   general graphs additionally keep   col int32 [nnz] (sorted per row)  and  deg int32 [n_rows]
//           (deg = number of non-zero stored values, what `adj.nonzero()` counts).
#include "graph.cuh"

#include <algorithm>
#include <vector>

namespace gsage {

static thread_local std::string tl_error;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    tl_error = buf;
}

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
    }
    return cached;
}

static int upload(const void* host, size_t bytes, void** dev) {
    *dev = nullptr;
    if (bytes == 0) bytes = 16;
    GS_CUDA(cudaMalloc(dev, bytes));
    if (host) GS_CUDA(cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return GSAGE_OK;
}

// shared tail of both constructors: host CSR (sorted, duplicate-free) -> device
static int build(const int64_t* indptr, const int64_t* indices, const int64_t* data, int64_t n_rows, int64_t n_cols,
                 gsage_graph** out) {
    const int64_t nnz = indptr[n_rows];
    bool fast = true, fits32 = true;
    std::vector<int32_t> deg(n_rows);
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        if (hi < lo) { set_error("graph: indptr not monotone at row %lld", (long long)r); return GSAGE_ERR_INVALID; }
        if (hi - lo > INT32_MAX) { set_error("graph: row %lld too long", (long long)r); return GSAGE_ERR_INVALID; }
        int32_t nz = 0;
        for (int64_t k = lo; k < hi; ++k) {
            const int64_t c = indices ? indices[k] : (k - lo);
            if (c != k - lo) fast = false;
            if (c < 0 || c >= n_cols) { set_error("graph: column %lld out of range", (long long)c); return GSAGE_ERR_INVALID; }
            if (indices && k > lo && indices[k] <= indices[k - 1]) {
                set_error("graph: column indices must be sorted and unique per row (scipy canonical form)");
                return GSAGE_ERR_INVALID;
            }
            if (data[k] != 0) ++nz; else fast = false;
            if (data[k] < INT32_MIN || data[k] > INT32_MAX) fits32 = false;
        }
        deg[r] = nz;
    }
    gsage_graph* g = new gsage_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->nnz = nnz; g->fast = fast; g->val64 = !fits32;
    g->host_deg.assign(deg.begin(), deg.end());
    int st = upload(indptr, sizeof(int64_t) * (n_rows + 1), (void**)&g->indptr);
    if (st == GSAGE_OK) {
        if (fits32) {
            std::vector<int32_t> v32(nnz);
            for (int64_t k = 0; k < nnz; ++k) v32[k] = (int32_t)data[k];
            st = upload(v32.data(), sizeof(int32_t) * nnz, (void**)&g->val);
        } else {
            st = upload(data, sizeof(int64_t) * nnz, (void**)&g->val);
        }
    }
    if (st == GSAGE_OK && !fast) {
        std::vector<int32_t> c32(nnz);
        for (int64_t r = 0; r < n_rows; ++r)
            for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) c32[k] = (int32_t)(indices ? indices[k] : k - indptr[r]);
        st = upload(c32.data(), sizeof(int32_t) * nnz, (void**)&g->col);
        if (st == GSAGE_OK) st = upload(deg.data(), sizeof(int32_t) * n_rows, (void**)&g->deg);
    }
    // the sticky "id out of range" flag lives in MAPPED pinned host memory: kernels store to it through the unified address,
    // the host polls it without a copy or a synchronisation (gsage_engine_poll_errors)
    if (st == GSAGE_OK && cudaHostAlloc((void**)&g->err_flag, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { g->err_flag = nullptr; st = GSAGE_ERR_NOMEM; }
    if (st == GSAGE_OK) *g->err_flag = 0;
    if (st != GSAGE_OK) { gsage_graph_destroy(g); return st; }
    g->device_bytes = sizeof(int64_t) * (n_rows + 1) + (fits32 ? 4 : 8) * nnz + (fast ? 0 : 4 * nnz + 4 * n_rows);
    *out = g;
    return GSAGE_OK;
}

}  // namespace gsage

using namespace gsage;

extern "C" {

int gsage_abi_version(void) { return GSAGE_ABI_VERSION; }
const char* gsage_last_error(void) { return tl_error.c_str(); }
int64_t gsage_launch_count(void) { return (int64_t)g_launches.load(); }

int gsage_set_device(int device) {
    GS_CUDA(cudaSetDevice(device));
    return GSAGE_OK;
}

int gsage_device_info(char* name_out, int name_cap, int* sms, int64_t* hbm_bytes, int* cc_major, int* cc_minor) {
    int dev = 0;
    GS_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GS_CUDA(cudaGetDeviceProperties(&p, dev));
    if (name_out && name_cap > 0) snprintf(name_out, name_cap, "%s", p.name);
    if (sms) *sms = p.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return GSAGE_OK;
}

int gsage_graph_from_csr(const int64_t* indptr, const int64_t* indices, const int64_t* data, int64_t n_rows,
                         int64_t n_cols, gsage_graph** out) {
    GS_CHECK_ARG(indptr && out && n_rows >= 0 && n_cols >= 0, "graph_from_csr: bad arguments");
    GS_CHECK_ARG(indptr[0] == 0, "graph_from_csr: indptr[0] != 0");
    GS_CHECK_ARG(data || indptr[n_rows] == 0, "graph_from_csr: data is NULL");
    return build(indptr, indices, data, n_rows, n_cols, out);
}

int gsage_graph_from_triplets(const int64_t* v, const int64_t* r, const int64_t* c, int64_t nnz, gsage_graph** out) {
    GS_CHECK_ARG(out && nnz >= 0 && (nnz == 0 || (v && r && c)), "graph_from_triplets: bad arguments");
    int64_t n_rows = 0, n_cols = 0;
    for (int64_t k = 0; k < nnz; ++k) {
        GS_CHECK_ARG(r[k] >= 0 && c[k] >= 0, "graph_from_triplets: negative index at entry %lld", (long long)k);
        n_rows = std::max(n_rows, r[k] + 1);
        n_cols = std::max(n_cols, c[k] + 1);
    }
    // counting sort by row, then per-row sort by column and duplicate merge (scipy: coo -> csr, sum_duplicates)
    std::vector<int64_t> indptr(n_rows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) indptr[r[k] + 1]++;
    for (int64_t i = 0; i < n_rows; ++i) indptr[i + 1] += indptr[i];
    std::vector<int64_t> fill(indptr.begin(), indptr.end() - 1);
    std::vector<std::pair<int64_t, int64_t>> ent(nnz);      // (col, value)
    for (int64_t k = 0; k < nnz; ++k) ent[fill[r[k]]++] = std::make_pair(c[k], v[k]);
    std::vector<int64_t> o_indptr(n_rows + 1, 0), o_col, o_val;
    o_col.reserve(nnz); o_val.reserve(nnz);
    for (int64_t i = 0; i < n_rows; ++i) {
        auto b = ent.begin() + indptr[i], e = ent.begin() + indptr[i + 1];
        if (!std::is_sorted(b, e, [](const std::pair<int64_t, int64_t>& x, const std::pair<int64_t, int64_t>& y) { return x.first < y.first; }))
            std::stable_sort(b, e, [](const std::pair<int64_t, int64_t>& x, const std::pair<int64_t, int64_t>& y) { return x.first < y.first; });
        for (auto it = b; it != e; ++it) {
            if (it != b && it->first == o_col.back() && (int64_t)o_col.size() > o_indptr[i]) o_val.back() += it->second;
            else { o_col.push_back(it->first); o_val.push_back(it->second); }
        }
        o_indptr[i + 1] = (int64_t)o_col.size();
    }
    return build(o_indptr.data(), o_col.data(), o_val.data(), n_rows, n_cols, out);
}

void gsage_graph_destroy(gsage_graph* g) {
    if (!g) return;
    cudaFree(g->indptr); cudaFree(g->val); cudaFree(g->col); cudaFree(g->deg); if (g->err_flag) cudaFreeHost(g->err_flag);
    delete g;
}

int gsage_graph_info(const gsage_graph* g, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, int* canonical,
                     int64_t* device_bytes) {
    GS_CHECK_ARG(g, "graph_info: NULL graph");
    if (n_rows) *n_rows = g->n_rows;
    if (n_cols) *n_cols = g->n_cols;
    if (nnz) *nnz = g->nnz;
    if (canonical) *canonical = g->fast ? 1 : 0;
    if (device_bytes) *device_bytes = g->device_bytes;
    return GSAGE_OK;
}

int gsage_graph_degrees_host(const gsage_graph* g, int64_t* degrees_host) {
    GS_CHECK_ARG(g && degrees_host, "graph_degrees_host: NULL argument");
    for (int64_t i = 0; i < g->n_rows; ++i) degrees_host[i] = g->host_deg[i];
    return GSAGE_OK;
}

int gsage_graph_check(gsage_graph* g, void* stream) {
    GS_CHECK_ARG(g, "graph_check: NULL graph");
    GS_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (*(volatile int*)g->err_flag) {
        *(volatile int*)g->err_flag = 0;
        set_error("sampler: id out of range of the adjacency (%lld rows) -- scipy would raise IndexError",
                  (long long)g->n_rows);
        return GSAGE_ERR_INDEX;
    }
    return GSAGE_OK;
}

}  // extern "C"
