// Handle lifecycle, workspace arena and per-tag kernel timing of libdatmo_b200.
#include "common.cuh"

int datmo_ws_reserve(datmo_ctx* h, size_t bytes) {
    if (bytes <= h->ws_cap) return DATMO_OK;
    // in-flight work may still use the old arena
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->ws) DATMO_CHECK_CUDA(h, cudaFree(h->ws));
    h->ws = nullptr;
    h->ws_cap = 0;
    size_t want = bytes + (bytes >> 3);
    DATMO_CHECK_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->ws), want));
    h->ws_cap = want;
    return DATMO_OK;
}

int datmo_pinned_reserve(datmo_ctx* h, size_t bytes) {
    if (bytes <= h->pinned_cap) return DATMO_OK;
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->pinned) DATMO_CHECK_CUDA(h, cudaFreeHost(h->pinned));
    h->pinned = nullptr;
    h->pinned_cap = 0;
    DATMO_CHECK_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->pinned), bytes));
    h->pinned_cap = bytes;
    return DATMO_OK;
}

int datmo_io_reserve(datmo_ctx* h, size_t bytes) {
    if (bytes <= h->io_cap) return DATMO_OK;
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->io) DATMO_CHECK_CUDA(h, cudaFree(h->io));
    h->io = nullptr;
    h->io_cap = 0;
    DATMO_CHECK_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->io), bytes));
    h->io_cap = bytes;
    return DATMO_OK;
}

LaunchScope::LaunchScope(datmo_ctx* h_, int tag, int n_launches) : h(h_) {
    h->launches += n_launches;
    if (!h->prof || !((h->prof_mask >> tag) & 1u)) return;
    h->prof_launches[tag] += n_launches;
    size_t used = h->ev_used.size();
    if (used >= h->ev_pool.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        h->ev_pool.emplace_back(a, b);
    }
    idx = static_cast<int>(used);
    h->ev_used.emplace_back(tag, idx);
    cudaEventRecord(h->ev_pool[idx].first, h->stream);
}

LaunchScope::~LaunchScope() {
    if (idx >= 0) cudaEventRecord(h->ev_pool[idx].second, h->stream);
}

static int profile_drain(datmo_ctx* h) {
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (auto& u : h->ev_used) {
        float ms = 0.f;
        auto& ev = h->ev_pool[u.second];
        DATMO_CHECK_CUDA(h, cudaEventElapsedTime(&ms, ev.first, ev.second));
        h->prof_ms[u.first] += ms;
    }
    h->ev_used.clear();
    return DATMO_OK;
}

extern "C" {

int datmo_abi_version(void) { return DATMO_ABI_VERSION; }

int datmo_create(int device, void* stream, datmo_handle_t* out) {
    if (!out) return DATMO_E_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return DATMO_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DATMO_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DATMO_E_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "datmo_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                prop.major, prop.minor);
        return DATMO_E_CUDA;
    }
    datmo_ctx* h = new datmo_ctx();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    if (stream) {
        h->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete h;
            return DATMO_E_CUDA;
        }
        h->own_stream = true;
    }
    *out = h;
    return DATMO_OK;
}

int datmo_destroy(datmo_handle_t h) {
    if (!h) return DATMO_E_INVALID;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->chain_cache) datmo_chain_destroy(h->chain_cache);
    for (auto& ev : h->ev_pool) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    if (h->ws) cudaFree(h->ws);
    if (h->fb_tab) cudaFree(h->fb_tab);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->io) cudaFree(h->io);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return DATMO_OK;
}

const char* datmo_last_error(datmo_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int datmo_synchronize(datmo_handle_t h) {
    if (!h) return DATMO_E_INVALID;
    DATMO_CHECK_CUDA(h, cudaStreamSynchronize(h->stream));
    return DATMO_OK;
}

size_t datmo_workspace_bytes(datmo_handle_t h) { return h ? h->ws_cap : 0; }

int datmo_profile_enable(datmo_handle_t h, int on) { return datmo_profile_tags(h, on ? 0xffffffffu : 0u); }

int datmo_profile_tags(datmo_handle_t h, unsigned mask) {
    if (!h) return DATMO_E_INVALID;
    DATMO_TRY(profile_drain(h));
    h->prof = mask != 0;
    h->prof_mask = mask;
    return DATMO_OK;
}

int datmo_profile_reset(datmo_handle_t h) {
    if (!h) return DATMO_E_INVALID;
    DATMO_TRY(profile_drain(h));
    for (int i = 0; i < DATMO_TAG_COUNT; ++i) {
        h->prof_launches[i] = 0;
        h->prof_ms[i] = 0;
    }
    return DATMO_OK;
}

int datmo_profile_read(datmo_handle_t h, int64_t launches[DATMO_TAG_COUNT], double ms[DATMO_TAG_COUNT]) {
    if (!h) return DATMO_E_INVALID;
    DATMO_TRY(profile_drain(h));
    for (int i = 0; i < DATMO_TAG_COUNT; ++i) {
        if (launches) launches[i] = h->prof_launches[i];
        if (ms) ms[i] = h->prof_ms[i];
    }
    return DATMO_OK;
}

int64_t datmo_launch_count(datmo_handle_t h) { return h ? h->launches : -1; }

}  // extern "C"
