// Flow -> clusters for batches of frame pairs held in HOST memory: the body of the reference's driver
// loop between the two BEVs and the EKF (Optical_flow/main.py:577-615 — compute_velocity_vectors,
// continuity_mask, the moving-cell filter, dbscan_clustering, extract_cluster_data) as one pipelined
// C-ABI object.  sm_100a.
//
// A chain owns n_slots submission slots.  submit(slot) enqueues, without blocking the host: the
// host-to-device copy of the slot's frames on a copy stream, the whole kernel chain on the handle's
// stream, a gather that packs the ragged per-pair results into ONE contiguous buffer per array
// (labels as int16 whenever every pair has fewer than 32 768 clusters, (row << 16) | col indices, the
// first kmax summary rows of every pair), and the read-back of the 2 x batch counters.  collect(slot)
// waits for the counters, sizes the read-back from them and issues one device-to-host copy per array
// on a second copy stream.  With two slots the copies of batch i overlap the kernels of batch i + 1.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

struct datmo_chain {
    datmo_ctx* h = nullptr;
    datmo_chain_config cfg;
    size_t n = 0, in_bytes = 0;   // cells per frame, bytes of one frame array of the batch
    cudaStream_t h2d = nullptr, d2h = nullptr;
    // shared by all slots (the kernels of successive submissions are ordered on the handle's stream)
    float *flow = nullptr, *vx_f = nullptr, *vy_f = nullptr;
    uint8_t* valid = nullptr;
    int32_t *labels = nullptr, *indices = nullptr, *n_valid = nullptr, *n_clusters = nullptr;
    double* summary = nullptr;
    char* dev = nullptr;      // one allocation behind all device buffers
    char* pinned = nullptr;   // one allocation behind all pinned host buffers
    struct Slot {
        char *prev = nullptr, *next = nullptr;       // device frames
        int32_t* d_counts = nullptr;                 // device: n_valid[batch], n_clusters[batch], wide flag, kmax
        int64_t* d_offsets = nullptr;                // device [batch + 1]
        char* d_labels = nullptr;                    // device, compact int16 / int32
        uint32_t* d_cells = nullptr;                 // device, compact (row << 16) | col
        double* d_summary = nullptr;                 // device, compact [batch][kmax][8]
        int32_t* h_counts = nullptr;                 // pinned mirrors
        int64_t* h_offsets = nullptr;
        char* h_labels = nullptr;
        uint32_t* h_cells = nullptr;
        double* h_summary = nullptr;
        cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
        bool busy = false;
        // the slot's launch train as a CUDA graph: captured on the second submission (the first one grows
        // the workspace, opts kernels into large shared memory and uploads the per-layer tables) and replayed
        // while the buffers it was captured against are still the handle's
        cudaGraphExec_t graph = nullptr;
        const char* graph_ws = nullptr;
        const float* graph_tab = nullptr;
        int64_t graph_launches = 0;
        int runs = 0;
    };
    bool use_graph = false;
    std::vector<Slot> slots;
    std::string err;
};

namespace {

// offsets of the pairs inside the compact arrays, label width, summary rows: one small CTA
__global__ void __launch_bounds__(256) k_chain_plan(const int32_t* __restrict__ n_valid,
                                                    const int32_t* __restrict__ n_clusters, int batch, int cap,
                                                    int max_clusters, int32_t* __restrict__ counts,
                                                    int64_t* __restrict__ offsets) {
    __shared__ long long s_part[256];
    __shared__ int s_max[256];
    const int per = (batch + 255) / 256;
    const int lo = min(static_cast<int>(threadIdx.x) * per, batch), hi = min(lo + per, batch);
    long long t = 0;
    int m = 0;
    for (int i = lo; i < hi; ++i) {
        t += min(n_valid[i], cap);
        m = max(m, n_clusters[i]);
        counts[i] = n_valid[i];
        counts[batch + i] = n_clusters[i];
    }
    s_part[threadIdx.x] = t;
    s_max[threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        int mm = 0;
        for (int i = 0; i < 256; ++i) {
            const long long v = s_part[i];
            s_part[i] = run;
            run += v;
            mm = max(mm, s_max[i]);
        }
        offsets[batch] = run;
        counts[2 * batch] = mm >= 32768;                 // labels need int32
        counts[2 * batch + 1] = min(mm, max_clusters);   // summary rows read back per pair
    }
    __syncthreads();
    long long run = s_part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        offsets[i] = run;
        run += min(n_valid[i], cap);
    }
}

// ragged [batch][cap] labels / (row, col) pairs -> contiguous arrays in pair order
__global__ void __launch_bounds__(256) k_chain_gather(const int32_t* __restrict__ labels,
                                                      const int2* __restrict__ indices,
                                                      const int32_t* __restrict__ n_valid, int cap, int batch,
                                                      const int32_t* __restrict__ counts,
                                                      const int64_t* __restrict__ offsets, void* __restrict__ out_labels,
                                                      uint32_t* __restrict__ out_cells) {
    const int b = blockIdx.y;
    const int n = min(n_valid[b], cap);
    const bool wide = counts[2 * batch] != 0;
    const long long o = offsets[b];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int lab = labels[static_cast<size_t>(b) * cap + i];
        const int2 rc = indices[static_cast<size_t>(b) * cap + i];
        if (wide)
            static_cast<int32_t*>(out_labels)[o + i] = lab;
        else
            static_cast<int16_t*>(out_labels)[o + i] = static_cast<int16_t>(lab);
        out_cells[o + i] = (static_cast<uint32_t>(rc.x) << 16) | static_cast<uint32_t>(rc.y);
    }
}

// [batch][max_clusters][8] -> [batch][kmax][8]
__global__ void __launch_bounds__(256) k_chain_gather_summary(const double* __restrict__ summary, int max_clusters,
                                                              int batch, const int32_t* __restrict__ counts,
                                                              double* __restrict__ out) {
    const int kmax = counts[2 * batch + 1];
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kmax * 8; i += gridDim.x * blockDim.x)
        out[static_cast<size_t>(b) * kmax * 8 + i] = summary[static_cast<size_t>(b) * max_clusters * 8 + i];
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

int chain_fail(datmo_chain* c, int status, const char* what) {
    c->err = what;
    c->h->err = what;
    return status;
}

#define CHAIN_CUDA(c, expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            char _buf[512];                                                                         \
            snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return chain_fail(c, DATMO_E_CUDA, _buf);                                               \
        }                                                                                           \
    } while (0)

}  // namespace

// everything one submission runs on the handle's stream, from the frames in the slot's device buffers to
// the counters in its pinned mirror
static int chain_enqueue(datmo_chain* c, datmo_chain::Slot& s) {
    datmo_ctx* h = c->h;
    const datmo_chain_config& g = c->cfg;
    const int B = g.batch;
    int st = datmo_farneback_dev(h, s.prev, s.next, g.dtype, g.H, g.W, B, &g.fb, c->flow);
    if (st == DATMO_OK)
        st = datmo_velocity_mask_dev(h, c->flow, g.H, g.W, B, g.px_x, g.px_y, g.alpha_cont, g.thresh, nullptr, nullptr,
                                     nullptr, nullptr, c->vx_f, c->vy_f, nullptr, c->valid, nullptr);
    if (st == DATMO_OK)
        st = datmo_dbscan_grid_dev(h, c->vx_f, c->vy_f, c->valid, g.H, g.W, B, g.eps, g.min_samples, g.cap, c->n_valid,
                                   c->labels, c->indices, c->n_clusters);
    if (st == DATMO_OK && g.max_clusters > 0)
        st = datmo_cluster_summary_dev(h, c->vx_f, c->vy_f, g.H, g.W, B, g.cap, c->n_valid, c->labels, c->indices,
                                       g.max_clusters, c->summary);
    if (st != DATMO_OK) {
        c->err = h->err;
        return st;
    }
    {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        k_chain_plan<<<1, 256, 0, h->stream>>>(c->n_valid, c->n_clusters, B, g.cap, g.max_clusters, s.d_counts, s.d_offsets);
    }
    CHAIN_CUDA(c, cudaGetLastError());
    if (g.want_cells) {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        const dim3 grid(std::max(1, std::min(64, ceil_div(4 * h->sm_count, B))), B);
        k_chain_gather<<<grid, 256, 0, h->stream>>>(c->labels, reinterpret_cast<const int2*>(c->indices), c->n_valid, g.cap,
                                                   B, s.d_counts, s.d_offsets, s.d_labels, s.d_cells);
    }
    CHAIN_CUDA(c, cudaGetLastError());
    if (g.max_clusters > 0) {
        LaunchScope ls(h, DATMO_TAG_CLUSTER);
        k_chain_gather_summary<<<dim3(4, B), 256, 0, h->stream>>>(c->summary, g.max_clusters, B, s.d_counts, s.d_summary);
    }
    CHAIN_CUDA(c, cudaGetLastError());
    CHAIN_CUDA(c, cudaMemcpyAsync(s.h_counts, s.d_counts, (2 * B + 2) * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CHAIN_CUDA(c, cudaMemcpyAsync(s.h_offsets, s.d_offsets, (B + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    return DATMO_OK;
}

extern "C" {

void datmo_chain_default_config(datmo_chain_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->dtype = DATMO_U8;
    cfg->batch = 1;
    cfg->alpha_cont = 0.2;    // config.yaml masks.alpha_cont[0]
    cfg->thresh = 0.1;        // main.py:609
    cfg->eps = 5.0;           // config.yaml dbscan_params
    cfg->min_samples = 3;
    cfg->max_clusters = 1024;
    cfg->want_cells = 1;
    cfg->n_slots = 2;
    datmo_farneback_default_params(&cfg->fb);
}

int datmo_chain_destroy(datmo_chain_t c) {
    if (!c) return DATMO_E_INVALID;
    cudaSetDevice(c->h->device);
    cudaStreamSynchronize(c->h->stream);
    if (c->h2d) cudaStreamSynchronize(c->h2d), cudaStreamDestroy(c->h2d);
    if (c->d2h) cudaStreamSynchronize(c->d2h), cudaStreamDestroy(c->d2h);
    for (auto& s : c->slots) {
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
        if (s.graph) cudaGraphExecDestroy(s.graph);
    }
    if (c->dev) cudaFree(c->dev);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->h->chain_cache == c) c->h->chain_cache = nullptr;
    delete c;
    return DATMO_OK;
}

int datmo_chain_create(datmo_handle_t h, const datmo_chain_config* cfg, datmo_chain_t* out) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, cfg && out, "null config / output pointer");
    *out = nullptr;
    DATMO_REQUIRE(h, cfg->H >= 2 && cfg->W >= 2 && cfg->batch >= 1 && cfg->batch <= 32767, "bad frame geometry / batch");
    DATMO_REQUIRE(h, cfg->H <= 65535 && cfg->W <= 65535, "H and W must fit the packed (row << 16) | col cell index");
    DATMO_REQUIRE(h, cfg->dtype == DATMO_U8 || cfg->dtype == DATMO_F32, "dtype must be DATMO_U8 or DATMO_F32");
    DATMO_REQUIRE(h, cfg->cap >= 1 && cfg->max_clusters >= 0 && cfg->n_slots >= 1 && cfg->n_slots <= 16, "bad cap / max_clusters / n_slots");
    DATMO_REQUIRE(h, cfg->px_x > 0 && cfg->px_y > 0, "pixel sizes must be positive");
    datmo_chain* c = new (std::nothrow) datmo_chain();
    DATMO_REQUIRE(h, c != nullptr, "out of host memory");
    c->h = h;
    c->cfg = *cfg;
    const size_t B = cfg->batch, cap = cfg->cap, K = cfg->max_clusters;
    c->n = static_cast<size_t>(cfg->H) * cfg->W;
    c->in_bytes = B * c->n * (cfg->dtype == DATMO_U8 ? 1 : 4);
    c->slots.resize(cfg->n_slots);
    // one device allocation, one pinned allocation
    size_t dev_bytes = 0, pin_bytes = 0;
    auto carve = [&](char* dbase, char* pbase) {
        size_t d = 0, p = 0;
        auto dtake = [&](size_t bytes) {
            char* r = dbase ? dbase + d : nullptr;
            d += align_up(bytes);
            return r;
        };
        auto ptake = [&](size_t bytes) {
            char* r = pbase ? pbase + p : nullptr;
            p += align_up(bytes);
            return r;
        };
        c->flow = reinterpret_cast<float*>(dtake(B * c->n * 8));
        c->vx_f = reinterpret_cast<float*>(dtake(B * c->n * 4));
        c->vy_f = reinterpret_cast<float*>(dtake(B * c->n * 4));
        c->valid = reinterpret_cast<uint8_t*>(dtake(B * c->n));
        c->labels = reinterpret_cast<int32_t*>(dtake(B * cap * 4));
        c->indices = reinterpret_cast<int32_t*>(dtake(B * cap * 8));
        c->n_valid = reinterpret_cast<int32_t*>(dtake(B * 4));
        c->n_clusters = reinterpret_cast<int32_t*>(dtake(B * 4));
        c->summary = reinterpret_cast<double*>(dtake(std::max<size_t>(B * K * 64, 8)));
        for (auto& s : c->slots) {
            s.prev = dtake(c->in_bytes);
            s.next = dtake(c->in_bytes);
            s.d_counts = reinterpret_cast<int32_t*>(dtake((2 * B + 2) * 4));
            s.d_offsets = reinterpret_cast<int64_t*>(dtake((B + 1) * 8));
            s.d_labels = dtake(cfg->want_cells ? B * cap * 4 : 8);
            s.d_cells = reinterpret_cast<uint32_t*>(dtake(cfg->want_cells ? B * cap * 4 : 8));
            s.d_summary = reinterpret_cast<double*>(dtake(std::max<size_t>(B * K * 64, 8)));
            s.h_counts = reinterpret_cast<int32_t*>(ptake((2 * B + 2) * 4));
            s.h_offsets = reinterpret_cast<int64_t*>(ptake((B + 1) * 8));
            s.h_labels = ptake(cfg->want_cells ? B * cap * 4 : 8);
            s.h_cells = reinterpret_cast<uint32_t*>(ptake(cfg->want_cells ? B * cap * 4 : 8));
            s.h_summary = reinterpret_cast<double*>(ptake(std::max<size_t>(B * K * 64, 8)));
        }
        dev_bytes = d, pin_bytes = p;
    };
    carve(nullptr, nullptr);
    auto bail = [&](const char* what) {
        h->err = what;
        datmo_chain_destroy(c);
        return DATMO_E_CUDA;
    };
    if (cudaMalloc(reinterpret_cast<void**>(&c->dev), dev_bytes) != cudaSuccess) return bail("cudaMalloc of the chain buffers failed");
    if (cudaMallocHost(reinterpret_cast<void**>(&c->pinned), pin_bytes) != cudaSuccess) return bail("cudaMallocHost of the chain buffers failed");
    carve(c->dev, c->pinned);
    if (cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking) != cudaSuccess)
        return bail("cudaStreamCreate failed");
    for (auto& s : c->slots)
        if (cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming) != cudaSuccess)
            return bail("cudaEventCreate failed");
    // opt-in (DATMO_CHAIN_GRAPH=1): measured, the replayed graph buys nothing — the train is bound by the
    // kernels' own dependent latencies, not by launch overhead (profiles/r02_latency.txt)
    c->use_graph = getenv("DATMO_CHAIN_GRAPH") ? atoi(getenv("DATMO_CHAIN_GRAPH")) != 0 : false;
    *out = c;
    return DATMO_OK;
}

const char* datmo_chain_last_error(datmo_chain_t c) { return c ? c->err.c_str() : "null chain"; }

int datmo_chain_submit(datmo_chain_t c, int slot, const void* prev_host, const void* next_host) {
    if (!c) return DATMO_E_INVALID;
    datmo_ctx* h = c->h;
    CHAIN_CUDA(c, cudaSetDevice(h->device));
    if (slot < 0 || slot >= static_cast<int>(c->slots.size()) || !prev_host || !next_host)
        return chain_fail(c, DATMO_E_INVALID, "invalid: bad slot or null frame pointer");
    datmo_chain::Slot& s = c->slots[slot];
    if (s.busy) return chain_fail(c, DATMO_E_INVALID, "invalid: slot still in flight, collect it first");
    const datmo_chain_config& g = c->cfg;
    const int B = g.batch;
    CHAIN_CUDA(c, cudaMemcpyAsync(s.prev, prev_host, c->in_bytes, cudaMemcpyHostToDevice, c->h2d));
    CHAIN_CUDA(c, cudaMemcpyAsync(s.next, next_host, c->in_bytes, cudaMemcpyHostToDevice, c->h2d));
    CHAIN_CUDA(c, cudaEventRecord(s.ev_h2d, c->h2d));
    CHAIN_CUDA(c, cudaStreamWaitEvent(h->stream, s.ev_h2d, 0));
    // the kernels of one submission: replayed as a graph when one is at hand, else launched (and, from the
    // second submission on, captured)
    const bool graphs = c->use_graph && !h->prof;
    if (graphs && s.graph && s.graph_ws == h->ws && s.graph_tab == h->fb_tab) {
        CHAIN_CUDA(c, cudaGraphLaunch(s.graph, h->stream));
        h->launches += s.graph_launches;
    } else {
        if (s.graph) {
            cudaGraphExecDestroy(s.graph);
            s.graph = nullptr;
        }
        const bool capture = graphs && s.runs >= 1;
        const int64_t launches0 = h->launches;
        if (capture) CHAIN_CUDA(c, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int st = chain_enqueue(c, s);
        if (capture) {
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
            if (st == DATMO_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&s.graph, graph, 0) == cudaSuccess) {
                s.graph_ws = h->ws;
                s.graph_tab = h->fb_tab;
                s.graph_launches = h->launches - launches0;
                st = cudaGraphLaunch(s.graph, h->stream) == cudaSuccess ? DATMO_OK : DATMO_E_CUDA;
            } else {
                // capture refused (something in the train synchronised or allocated): plain launches from now on
                cudaGetLastError();
                s.graph = nullptr;
                c->use_graph = false;
                h->launches = launches0;
                st = chain_enqueue(c, s);
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (st != DATMO_OK) {
            if (c->err.empty()) c->err = h->err;
            return st;
        }
        ++s.runs;
    }
    CHAIN_CUDA(c, cudaEventRecord(s.ev_done, h->stream));
    s.busy = true;
    return DATMO_OK;
}

int datmo_chain_collect(datmo_chain_t c, int slot, datmo_chain_result* out) {
    if (!c) return DATMO_E_INVALID;
    datmo_ctx* h = c->h;
    CHAIN_CUDA(c, cudaSetDevice(h->device));
    if (slot < 0 || slot >= static_cast<int>(c->slots.size()) || !out)
        return chain_fail(c, DATMO_E_INVALID, "invalid: bad slot or null result pointer");
    datmo_chain::Slot& s = c->slots[slot];
    if (!s.busy) return chain_fail(c, DATMO_E_INVALID, "invalid: nothing submitted on this slot");
    const datmo_chain_config& g = c->cfg;
    const int B = g.batch;
    CHAIN_CUDA(c, cudaEventSynchronize(s.ev_done));
    const int64_t total = s.h_offsets[B];
    const int label_bytes = s.h_counts[2 * B] ? 4 : 2;
    const int kmax = g.max_clusters > 0 ? s.h_counts[2 * B + 1] : 0;
    int64_t d2h = (2 * B + 2) * sizeof(int32_t) + (B + 1) * sizeof(int64_t);
    CHAIN_CUDA(c, cudaStreamWaitEvent(c->d2h, s.ev_done, 0));
    if (g.want_cells && total > 0) {
        CHAIN_CUDA(c, cudaMemcpyAsync(s.h_labels, s.d_labels, total * label_bytes, cudaMemcpyDeviceToHost, c->d2h));
        CHAIN_CUDA(c, cudaMemcpyAsync(s.h_cells, s.d_cells, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->d2h));
        d2h += total * (label_bytes + 4);
    }
    if (kmax > 0) {
        CHAIN_CUDA(c, cudaMemcpyAsync(s.h_summary, s.d_summary, static_cast<size_t>(B) * kmax * 64, cudaMemcpyDeviceToHost, c->d2h));
        d2h += static_cast<int64_t>(B) * kmax * 64;
    }
    CHAIN_CUDA(c, cudaEventRecord(s.ev_d2h, c->d2h));
    CHAIN_CUDA(c, cudaEventSynchronize(s.ev_d2h));
    s.busy = false;
    out->n_valid = s.h_counts;
    out->n_clusters = s.h_counts + B;
    out->offsets = s.h_offsets;
    out->labels = g.want_cells ? s.h_labels : nullptr;
    out->label_bytes = label_bytes;
    out->cells = g.want_cells ? s.h_cells : nullptr;
    out->summary = kmax > 0 ? s.h_summary : nullptr;
    out->summary_rows = kmax;
    out->h2d_bytes = 2 * static_cast<int64_t>(c->in_bytes);
    out->d2h_bytes = d2h;
    out->truncated = 0;
    for (int b = 0; b < B; ++b) out->truncated |= s.h_counts[b] > g.cap;
    return DATMO_OK;
}

int datmo_flow_to_clusters_host(datmo_handle_t h, const void* prev, const void* next, const datmo_chain_config* cfg,
                                int32_t* n_valid, int32_t* n_clusters, int64_t* offsets, int32_t* labels,
                                int32_t* indices, int64_t capacity_cells, double* summary) {
    DATMO_ENTER(h);
    DATMO_REQUIRE(h, prev && next && cfg && n_valid && offsets, "null pointer");
    // the chain (device + pinned buffers, streams) is kept on the handle between calls with the same configuration
    datmo_chain* c = h->chain_cache;
    datmo_chain_config want = *cfg;
    want.n_slots = 1;
    want.want_cells = labels != nullptr || indices != nullptr;
    if (!summary) want.max_clusters = 0;
    auto same = [](const datmo_chain_config& a, const datmo_chain_config& b) {
        return a.H == b.H && a.W == b.W && a.batch == b.batch && a.dtype == b.dtype && a.px_x == b.px_x &&
               a.px_y == b.px_y && a.alpha_cont == b.alpha_cont && a.thresh == b.thresh && a.eps == b.eps &&
               a.min_samples == b.min_samples && a.cap == b.cap && a.max_clusters == b.max_clusters &&
               a.want_cells == b.want_cells && a.n_slots == b.n_slots && a.fb.pyr_scale == b.fb.pyr_scale &&
               a.fb.levels == b.fb.levels && a.fb.winsize == b.fb.winsize && a.fb.iterations == b.fb.iterations &&
               a.fb.poly_n == b.fb.poly_n && a.fb.poly_sigma == b.fb.poly_sigma && a.fb.flags == b.fb.flags &&
               a.fb.variant == b.fb.variant;
    };
    if (!c || !same(c->cfg, want)) {
        if (c) datmo_chain_destroy(c);
        h->chain_cache = nullptr;
        DATMO_TRY(datmo_chain_create(h, &want, &c));
        h->chain_cache = c;
    }
    DATMO_TRY(datmo_chain_submit(c, 0, prev, next));
    datmo_chain_result r;
    DATMO_TRY(datmo_chain_collect(c, 0, &r));
    const int B = want.batch;
    memcpy(n_valid, r.n_valid, sizeof(int32_t) * B);
    if (n_clusters) memcpy(n_clusters, r.n_clusters, sizeof(int32_t) * B);
    memcpy(offsets, r.offsets, sizeof(int64_t) * (B + 1));
    const int64_t total = r.offsets[B];
    if (want.want_cells) {
        if (total > capacity_cells) {
            h->err = "capacity: more moving cells than the caller's labels / indices arrays hold";
            return DATMO_E_CAPACITY;
        }
        for (int64_t i = 0; i < total; ++i) {
            if (labels)
                labels[i] = r.label_bytes == 2 ? static_cast<const int16_t*>(r.labels)[i] : static_cast<const int32_t*>(r.labels)[i];
            if (indices) {
                indices[2 * i] = static_cast<int32_t>(r.cells[i] >> 16);
                indices[2 * i + 1] = static_cast<int32_t>(r.cells[i] & 0xffffu);
            }
        }
    }
    if (summary) {
        // caller's layout: [batch][max_clusters][8], rows past the pair's cluster count zero
        memset(summary, 0, sizeof(double) * 8 * static_cast<size_t>(B) * cfg->max_clusters);
        for (int b = 0; b < B; ++b)
            if (r.summary_rows > 0)
                memcpy(summary + static_cast<size_t>(b) * cfg->max_clusters * 8, r.summary + static_cast<size_t>(b) * r.summary_rows * 8,
                       sizeof(double) * 8 * r.summary_rows);
    }
    return r.truncated ? DATMO_E_CAPACITY : DATMO_OK;
}

}  // extern "C"
