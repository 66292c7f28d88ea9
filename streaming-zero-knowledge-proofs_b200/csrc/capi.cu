// extern "C" surface of libsezkp_cuda.so (include/sezkp_cuda.h).  Every entry point converts C++
// exceptions into status codes + ctx->last_error; nothing aborts.
#include <sys/mman.h>
#include <sys/stat.h>

#include <chrono>
#include <memory>
#include <cstdio>
#include <thread>

#include "jsonl.hpp"
#include <cstdlib>
#include <cstring>
#include <string>

#include "gl.cuh"
#include "group.cuh"
#include "hash.cuh"
#include "ntt.cuh"
#include "stark.cuh"
#include "wide.cuh"

static std::string g_create_error;

struct sezkp_tree {
    Commit cm;
};
struct sezkp_fri {
    FriLayers fl;
};
struct sezkp_trace_dev {
    DeviceTraceOwner owner;                // the ctx's own GPU (rank 0 of a group)
    std::vector<DeviceTraceOwner> peers;   // context group: the same trace on ranks 1..world-1
};
struct sezkp_columns {
    std::vector<WideColumns> part;  // one per rank (a single entry without a group)
    int c = 0, log_n = 0;
};
static int group_world(const sezkp_ctx* ctx) { return ctx->group ? ctx->group->world : 1; }
// run fn(rank, ctx_of_rank) on every GPU of the ctx's group (or inline on a plain single-GPU ctx)
static void on_all_gpus(sezkp_ctx* ctx, const std::function<void(int, sezkp_ctx*)>& fn) {
    if (ctx->group) group_run(ctx->group, fn);
    else fn(0, ctx);
}
#define API_BEGIN(ctx)                                         \
    if (!(ctx)) return SEZKP_CUDA_EINVAL;                      \
    try {                                                      \
        CUDA_CHECK(cudaSetDevice((ctx)->device));
#define API_END(ctx)                                           \
        return SEZKP_CUDA_OK;                                  \
    } catch (const SezkpError& e) {                            \
        (ctx)->last_error = e.what();                          \
        cudaGetLastError();                                    \
        return e.code;                                         \
    } catch (const std::bad_alloc&) {                          \
        (ctx)->last_error = "host allocation failed";          \
        return SEZKP_CUDA_ENOMEM;                              \
    } catch (const std::exception& e) {                        \
        (ctx)->last_error = e.what();                          \
        return SEZKP_CUDA_ECUDA;                               \
    }

static void check_canonical(const u64* v, size_t n, const char* what) {
    for (size_t i = 0; i < n; i++)
        if (v[i] >= gl::P) sezkp_fail(SEZKP_CUDA_EINVAL, "%s[%zu] is not a canonical field element", what, i);
}
static void h2d(sezkp_ctx* ctx, void* d, const void* h, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
}
static void d2h(sezkp_ctx* ctx, void* h, const void* d, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

extern "C" {

uint32_t sezkp_cuda_abi_version(void) { return SEZKP_CUDA_ABI_VERSION; }

int32_t sezkp_cuda_create(int device_id, sezkp_ctx** out) {
    if (!out) return SEZKP_CUDA_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libsezkp_cuda has no CPU fallback)";
        cudaGetLastError();
        return SEZKP_CUDA_ENODEV;
    }
    if (device_id < 0) {
        if (cudaGetDevice(&device_id) != cudaSuccess) device_id = 0;
    }
    if (device_id >= count) {
        g_create_error = "device id out of range";
        return SEZKP_CUDA_EINVAL;
    }
    sezkp_ctx* ctx = new (std::nothrow) sezkp_ctx();
    if (!ctx) return SEZKP_CUDA_ENOMEM;
    ctx->device = device_id;
    try {
        CUDA_CHECK(cudaSetDevice(device_id));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device_id));
        ctx->sm_count = prop.multiProcessorCount;
        if (prop.major < 10) sezkp_fail(SEZKP_CUDA_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
        CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
        const char* nd = getenv("SEZKP_NO_DEDUP");
        if (nd && nd[0] == '1') ctx->dedup_enabled = false;
    } catch (const SezkpError& err) {
        g_create_error = err.what();
        int32_t code = err.code;
        delete ctx;
        return code;
    }
    *out = ctx;
    return SEZKP_CUDA_OK;
}

int32_t sezkp_cuda_create_multi(const int* device_ids, int n_dev, sezkp_ctx** out) {
    if (!out) return SEZKP_CUDA_EINVAL;
    *out = nullptr;
    try {
        sezkp_group* g = group_create(device_ids, n_dev);
        *out = g->ctx[0];
    } catch (const SezkpError& e) {
        g_create_error = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return SEZKP_CUDA_ENOMEM;
    }
    return SEZKP_CUDA_OK;
}
int32_t sezkp_cuda_group_size(const sezkp_ctx* ctx) { return ctx ? group_world(ctx) : 0; }
int32_t sezkp_cuda_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

void sezkp_cuda_destroy(sezkp_ctx* ctx) {
    if (!ctx) return;
    if (ctx->group) {  // the caller holds ctx[0] of a group: tear the whole group down (it calls back here per member)
        if (ctx->group->ctx[0] == ctx) group_destroy(ctx->group);
        return;
    }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ntt_free_tables(ctx);
    for (auto& pb : ctx->pinned) pb.release();
    for (auto& pb : ctx->stream_stage) pb.release();
    for (auto& kv : ctx->deep_tables) cudaFree(kv.second);
    ctx->deep_tables.clear();
    for (auto& kv : ctx->power_tables) cudaFree(kv.second);
    ctx->power_tables.clear();
    for (auto& b : ctx->scratch) b.release();
    for (auto& kv : ctx->pool.live) cudaFree(kv.first);
    ctx->pool.live.clear();
    ctx->pool.trim();
    for (auto e : ctx->slab_events) cudaEventDestroy(e);
    for (auto e : ctx->phase_events) cudaEventDestroy(e);
    delete ctx->open_reqs;
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sezkp_cuda_last_error(const sezkp_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int32_t sezkp_cuda_set_stream(sezkp_ctx* ctx, void* cuda_stream, int use_own) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (use_own) {
        if (!ctx->own_stream) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->own_stream = true;
        }
    } else {  // NULL is the legacy default stream
        if (ctx->own_stream) CUDA_CHECK(cudaStreamDestroy(ctx->stream));
        ctx->stream = (cudaStream_t)cuda_stream;
        ctx->own_stream = false;
    }
    API_END(ctx)
}

int32_t sezkp_cuda_synchronize(sezkp_ctx* ctx) {
    API_BEGIN(ctx)
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    API_END(ctx)
}

static void set_option_one(sezkp_ctx* ctx, const char* name, int64_t value) {
    if (std::strcmp(name, "dedup") == 0) {
        ctx->dedup_enabled = value != 0;
        if (value == 1 || value == 2) ctx->dedup_variant = (int)value;
    } else if (std::strcmp(name, "tabled") == 0) {
        ctx->tabled_enabled = value != 0;
    } else if (std::strcmp(name, "ntt_gen") == 0) {
        REQUIRE(value >= 1 && value <= 4, "ntt_gen must be 1..4");
        ctx->ntt_gen = (int)value;
    } else if (std::strcmp(name, "lde_fuse") == 0) {
        ctx->lde_fuse = value != 0;
    } else if (std::strcmp(name, "phase_sync") == 0) {
        ctx->phase_sync = value != 0;
    } else if (std::strcmp(name, "deep_fused") == 0) {
        ctx->deep_fused = value != 0;
    } else if (std::strcmp(name, "fri_coset") == 0) {
        ctx->fri_coset = value != 0;
    } else if (std::strcmp(name, "tab_cache") == 0) {
        ctx->tab_cache_enabled = value != 0;
        ctx->tab_cache_key.clear();
    } else sezkp_fail(SEZKP_CUDA_EINVAL, "unknown option '%s'", name);
}
int32_t sezkp_cuda_set_option(sezkp_ctx* ctx, const char* name, int64_t value) {
    API_BEGIN(ctx)
    REQUIRE(name != nullptr, "bad argument");
    for (int r = 0; r < group_world(ctx); r++) set_option_one(ctx->group ? ctx->group->ctx[r] : ctx, name, value);  // every GPU of a group
    API_END(ctx)
}

uint64_t sezkp_cuda_launch_count(sezkp_ctx* ctx, int reset) {
    if (!ctx) return 0;
    uint64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

static std::string timings_json(const sezkp_ctx* c) {
    std::string s = "{";
    for (size_t i = 0; i < c->timings.size(); i++) {
        char t[128];
        snprintf(t, sizeof t, "%s\"%s\": %.4f", i ? ", " : "", c->timings[i].first.c_str(), c->timings[i].second);
        s += t;
    }
    return s + "}";
}
int32_t sezkp_cuda_get_timings_gpu(sezkp_ctx* ctx, int rank, char* json_buf, size_t cap) {
    API_BEGIN(ctx)
    REQUIRE(rank >= 0 && rank < group_world(ctx), "rank out of range");
    const std::string s = timings_json(ctx->group ? ctx->group->ctx[rank] : ctx);
    if (!json_buf || cap < s.size() + 1) sezkp_fail(SEZKP_CUDA_ERANGE, "timings buffer too small (need %zu)", s.size() + 1);
    std::memcpy(json_buf, s.c_str(), s.size() + 1);
    API_END(ctx)
}
int32_t sezkp_cuda_get_timings(sezkp_ctx* ctx, char* json_buf, size_t cap) {
    API_BEGIN(ctx)
    std::string s = timings_json(ctx);
    if (!json_buf || cap < s.size() + 1) sezkp_fail(SEZKP_CUDA_ERANGE, "timings buffer too small (need %zu)", s.size() + 1);
    std::memcpy(json_buf, s.c_str(), s.size() + 1);
    API_END(ctx)
}

/* ------------------------------------------------------------ NTT / LDE ------ */
int32_t sezkp_ntt_batch_dev(sezkp_ctx* ctx, uint64_t* data_dev, int log_n, int cols, int inverse) {
    API_BEGIN(ctx)
    REQUIRE(data_dev != nullptr && cols >= 0 && log_n >= 0 && log_n <= 30, "bad argument");
    u64* tmp = log_n > 10 ? (u64*)ctx->scratch[0].ensure(((size_t)cols << log_n) * 8) : nullptr;
    ntt_batch_device(ctx, data_dev, tmp, log_n, (u64)cols, inverse != 0);
    API_END(ctx)
}
int32_t sezkp_ntt_batch(sezkp_ctx* ctx, uint64_t* data, int log_n, int cols, int inverse) {
    API_BEGIN(ctx)
    REQUIRE(data != nullptr && cols >= 0 && log_n >= 0 && log_n <= 30, "bad argument");
    const size_t count = (size_t)cols << log_n;
    if (count == 0) return SEZKP_CUDA_OK;
    check_canonical(data, count, "data");
    u64* d = (u64*)ctx->scratch[1].ensure(count * 8);
    u64* tmp = log_n > 10 ? (u64*)ctx->scratch[0].ensure(count * 8) : nullptr;
    h2d(ctx, d, data, count * 8);
    ntt_batch_device(ctx, d, tmp, log_n, (u64)cols, inverse != 0);
    d2h(ctx, data, d, count * 8);
    API_END(ctx)
}

int32_t sezkp_coset_lde_batch_dev(sezkp_ctx* ctx, const uint64_t* coeffs_dev, int log_n, int log_blow, uint64_t shift, int cols,
                                  uint64_t* out_dev) {
    API_BEGIN(ctx)
    REQUIRE(coeffs_dev && out_dev && cols >= 0, "bad argument");
    u64* inter = log_n > 10 ? (u64*)ctx->scratch[1].ensure(((size_t)cols << (log_n + log_blow)) * 8) : nullptr;
    coset_lde_device(ctx, coeffs_dev, out_dev, inter, log_n, log_blow, shift, (u64)cols);
    API_END(ctx)
}
int32_t sezkp_coset_lde_batch(sezkp_ctx* ctx, const uint64_t* coeffs, int log_n, int log_blow, uint64_t shift, int cols,
                              uint64_t* out) {
    API_BEGIN(ctx)
    REQUIRE(coeffs && out && cols >= 0 && log_n >= 1 && log_n <= 30 && log_blow >= 0 && log_blow <= 4, "bad argument");
    const size_t n_in = (size_t)cols << log_n, n_out = n_in << log_blow;
    if (n_in == 0) return SEZKP_CUDA_OK;
    check_canonical(coeffs, n_in, "coeffs");
    u64* d_in = (u64*)ctx->scratch[2].ensure(n_in * 8);
    u64* d_out = (u64*)ctx->scratch[3].ensure(n_out * 8);
    u64* inter = log_n > 10 ? (u64*)ctx->scratch[1].ensure(n_out * 8) : nullptr;
    h2d(ctx, d_in, coeffs, n_in * 8);
    coset_lde_device(ctx, d_in, d_out, inter, log_n, log_blow, shift, (u64)cols);
    d2h(ctx, out, d_out, n_out * 8);
    API_END(ctx)
}

static void lde_from_evals_device(sezkp_ctx* ctx, const u64* evals_dev, int log_n, int log_blow, u64 shift, int cols, u64* out_dev) {
    const size_t n_in = (size_t)cols << log_n, n_out = n_in << log_blow;
    u64* coeffs = (u64*)ctx->scratch[4].ensure(n_in * 8);
    CUDA_CHECK(cudaMemcpyAsync(coeffs, evals_dev, n_in * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    u64* tmp = log_n > 10 ? (u64*)ctx->scratch[0].ensure(n_in * 8) : nullptr;
    ntt_batch_device(ctx, coeffs, tmp, log_n, (u64)cols, true);
    u64* inter = log_n > 10 ? (u64*)ctx->scratch[1].ensure(n_out * 8) : nullptr;
    coset_lde_device(ctx, coeffs, out_dev, inter, log_n, log_blow, shift, (u64)cols);
}
int32_t sezkp_lde_from_evals_batch_dev(sezkp_ctx* ctx, const uint64_t* evals_dev, int log_n, int log_blow, uint64_t shift,
                                       int cols, uint64_t* out_dev) {
    API_BEGIN(ctx)
    REQUIRE(evals_dev && out_dev && cols >= 0 && log_n >= 1 && log_n <= 30, "bad argument");
    if (cols) lde_from_evals_device(ctx, evals_dev, log_n, log_blow, shift, cols, out_dev);
    API_END(ctx)
}
int32_t sezkp_lde_from_evals_batch(sezkp_ctx* ctx, const uint64_t* evals, int log_n, int log_blow, uint64_t shift, int cols,
                                   uint64_t* out) {
    API_BEGIN(ctx)
    REQUIRE(evals && out && cols >= 0 && log_n >= 1 && log_n <= 30 && log_blow >= 0 && log_blow <= 4, "bad argument");
    const size_t n_in = (size_t)cols << log_n, n_out = n_in << log_blow;
    if (n_in == 0) return SEZKP_CUDA_OK;
    check_canonical(evals, n_in, "evals");
    u64* d_in = (u64*)ctx->scratch[2].ensure(n_in * 8);
    u64* d_out = (u64*)ctx->scratch[3].ensure(n_out * 8);
    h2d(ctx, d_in, evals, n_in * 8);
    lde_from_evals_device(ctx, d_in, log_n, log_blow, shift, cols, d_out);
    d2h(ctx, out, d_out, n_out * 8);
    API_END(ctx)
}

// LDE + commit (BASELINE config 4, first half): see wide.cu.  On a context group the columns are sharded c % world.
static void lde_commit_device(sezkp_ctx* ctx, const u64* evals_dev, const char* const* labels, int c, int log_n, int log_blow, u64 shift,
                              int chunk_log2, u8* roots_host) {
    u8* d_roots = (u8*)ctx->scratch[10].ensure((size_t)c * 32 + 64);
    lde_commit_columns(ctx, evals_dev, labels, c, log_n, log_blow, shift, chunk_log2, d_roots);
    CUDA_CHECK(cudaMemcpyAsync(roots_host, d_roots, (size_t)c * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
int32_t sezkp_lde_commit_batch_dev(sezkp_ctx* ctx, const uint64_t* evals_dev, const char* const* labels, int c, int log_n, int log_blow,
                                   uint64_t shift, int chunk_log2, uint8_t* roots) {
    API_BEGIN(ctx)
    REQUIRE(evals_dev && labels && roots && c >= 1 && log_n >= 1 && log_n <= 29 && log_blow >= 0 && log_blow <= 4, "bad argument");
    lde_commit_device(ctx, evals_dev, labels, c, log_n, log_blow, shift, chunk_log2, roots);
    API_END(ctx)
}
int32_t sezkp_lde_commit_batch(sezkp_ctx* ctx, const uint64_t* evals, const char* const* labels, int c, int log_n, int log_blow,
                               uint64_t shift, int chunk_log2, uint8_t* roots) {
    API_BEGIN(ctx)
    REQUIRE(evals && labels && roots && c >= 1 && log_n >= 1 && log_n <= 29 && log_blow >= 0 && log_blow <= 4, "bad argument");
    const size_t count = (size_t)c << log_n;
    check_canonical(evals, count, "evals");
    if (ctx->group) {  // columns sharded c % world: every GPU uploads and commits its own columns, roots are gathered on the host
        const int world = group_world(ctx), max_local = (c + world - 1) / world;
        std::vector<u8> all((size_t)world * max_local * 32, 0);
        group_run(ctx->group, [&](int r, sezkp_ctx* cx) {
            WideColumns wc;
            try {
                wide_columns_upload(cx, wc, evals, c, log_n, r, world);
                if (wc.n_local) {
                    std::vector<const char*> ll;
                    for (int j = 0; j < wc.n_local; j++) ll.push_back(labels[r + j * world]);
                    lde_commit_device(cx, wc.evals.as<u64>(), ll.data(), wc.n_local, log_n, log_blow, shift, chunk_log2,
                                      &all[(size_t)r * max_local * 32]);
                }
            } catch (...) {
                wc.release();
                throw;
            }
            wc.release();
        });
        for (int k = 0; k < c; k++) std::memcpy(roots + 32 * (size_t)k, &all[((size_t)(k % world) * max_local + k / world) * 32], 32);
        return SEZKP_CUDA_OK;
    }
    u64* d = (u64*)ctx->scratch[2].ensure(count * 8);
    h2d(ctx, d, evals, count * 8);
    lde_commit_device(ctx, d, labels, c, log_n, log_blow, shift, chunk_log2, roots);
    API_END(ctx)
}

int32_t sezkp_deep_lde_dev(sezkp_ctx* ctx, const uint64_t* base_evals_dev, int log_n, int log_blow, uint64_t shift, uint64_t z,
                           uint64_t* out_dev) {
    API_BEGIN(ctx)
    REQUIRE(base_evals_dev && out_dev && log_n >= 1 && log_n <= 29 && log_blow >= 0 && log_blow <= 4, "bad argument");
    const size_t n = (size_t)1 << log_n;
    u64* base = (u64*)ctx->scratch[4].ensure(n * 8);
    CUDA_CHECK(cudaMemcpyAsync(base, base_evals_dev, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    deep_lde_device(ctx, base, out_dev, log_n, log_blow, shift, z);
    API_END(ctx)
}
int32_t sezkp_deep_lde(sezkp_ctx* ctx, const uint64_t* base_evals, int log_n, int log_blow, uint64_t shift, uint64_t z,
                       uint64_t* out) {
    API_BEGIN(ctx)
    REQUIRE(base_evals && out && log_n >= 1 && log_n <= 29 && log_blow >= 0 && log_blow <= 4, "bad argument");
    const size_t n = (size_t)1 << log_n, N = n << log_blow;
    check_canonical(base_evals, n, "base_evals");
    u64* base = (u64*)ctx->scratch[4].ensure(n * 8);
    u64* d_out = (u64*)ctx->scratch[3].ensure(N * 8);
    h2d(ctx, base, base_evals, n * 8);
    deep_lde_device(ctx, base, d_out, log_n, log_blow, shift, z);
    d2h(ctx, out, d_out, N * 8);
    API_END(ctx)
}

/* ---- resident column sets + the config-4 pipeline (wide.cu) ---- */
int32_t sezkp_columns_upload(sezkp_ctx* ctx, const uint64_t* evals, int c, int log_n, sezkp_columns** out) {
    API_BEGIN(ctx)
    REQUIRE(evals && out && c >= 1 && log_n >= 1 && log_n <= 29, "bad argument");
    *out = nullptr;
    check_canonical(evals, (size_t)c << log_n, "evals");
    const int world = group_world(ctx);
    std::unique_ptr<sezkp_columns> h(new sezkp_columns());
    h->c = c;
    h->log_n = log_n;
    h->part.resize(world);
    try {
        on_all_gpus(ctx, [&](int r, sezkp_ctx* cx) { wide_columns_upload(cx, h->part[r], evals, c, log_n, r, world); });
    } catch (...) {
        sezkp_columns_free(ctx, h.release());
        throw;
    }
    *out = h.release();
    API_END(ctx)
}
int32_t sezkp_columns_synth(sezkp_ctx* ctx, uint64_t seed, int c, int log_n, sezkp_columns** out) {
    API_BEGIN(ctx)
    REQUIRE(out && c >= 1 && log_n >= 1 && log_n <= 29, "bad argument");
    *out = nullptr;
    const int world = group_world(ctx);
    std::unique_ptr<sezkp_columns> h(new sezkp_columns());
    h->c = c;
    h->log_n = log_n;
    h->part.resize(world);
    try {
        on_all_gpus(ctx, [&](int r, sezkp_ctx* cx) {
            wide_columns_synth(cx, h->part[r], seed, c, log_n, r, world);
            CUDA_CHECK(cudaStreamSynchronize(cx->stream));
        });
    } catch (...) {
        sezkp_columns_free(ctx, h.release());
        throw;
    }
    *out = h.release();
    API_END(ctx)
}
void sezkp_columns_free(sezkp_ctx* ctx, sezkp_columns* cols) {
    if (!cols) return;
    if (ctx) {
        for (size_t r = 0; r < cols->part.size(); r++) {
            sezkp_ctx* cx = ctx->group ? ctx->group->ctx[r] : ctx;
            cudaSetDevice(cx->device);
            cudaStreamSynchronize(cx->stream);
            cols->part[r].release();
        }
        cudaSetDevice(ctx->device);
    }
    delete cols;
}
int32_t sezkp_lde_commit_fri(sezkp_ctx* ctx, const sezkp_columns* cols, const char* const* labels, int log_blow, uint64_t shift,
                             int chunk_log2, uint8_t* col_roots, uint8_t* fri_roots, uint64_t* final_value) {
    API_BEGIN(ctx)
    REQUIRE(cols && labels && col_roots && fri_roots && final_value, "bad argument");
    REQUIRE((int)cols->part.size() == group_world(ctx), "column set was created on a different context");
    for (int k = 0; k < cols->c; k++) REQUIRE(labels[k] != nullptr, "label %d is NULL", k);
    on_all_gpus(ctx, [&](int r, sezkp_ctx* cx) {
        wide_commit_fri_rank(cx, cols->part[r], labels, log_blow, shift, chunk_log2, r == 0 ? col_roots : nullptr,
                             r == 0 ? fri_roots : nullptr, r == 0 ? final_value : nullptr, nullptr);
    });
    double dev_max = 0;  // every GPU timed its own part with CUDA events on its stream: report the maximum
    for (int r = 0; r < group_world(ctx); r++)
        for (auto& kv : (ctx->group ? ctx->group->ctx[r] : ctx)->timings)
            if (kv.first == "device_ms" && kv.second > dev_max) dev_max = kv.second;
    ctx->timings.push_back({"device_ms_max_over_gpus", dev_max});
    API_END(ctx)
}

/* ------------------------------------------------------ hashing / Merkle ----- */
int32_t sezkp_leaf_hash(sezkp_ctx* ctx, const uint64_t* vals, size_t n, const char* label, uint8_t* out) {
    API_BEGIN(ctx)
    REQUIRE((vals && out) || n == 0, "bad argument");
    if (n == 0) return SEZKP_CUDA_OK;
    check_canonical(vals, n, "vals");
    u64* d_v = (u64*)ctx->scratch[2].ensure(n * 8);
    u32* d_o = (u32*)ctx->scratch[3].ensure(n * 32);
    h2d(ctx, d_v, vals, n * 8);
    leaf_hash_device(ctx, d_v, n, label, d_o);
    d2h(ctx, out, d_o, n * 32);
    API_END(ctx)
}

int32_t sezkp_merkle_root(sezkp_ctx* ctx, const uint8_t* leaves, size_t n, uint8_t out_root[32]) {
    API_BEGIN(ctx)
    REQUIRE(out_root && (leaves || n == 0), "bad argument");
    if (n == 0) {  // MerkleTree::from_leaves(&[]) -> single zero leaf (reference v1/merkle.rs:48-50)
        std::memset(out_root, 0, 32);
        return SEZKP_CUDA_OK;
    }
    u32* a = (u32*)ctx->scratch[2].ensure(n * 32);
    u32* b = (u32*)ctx->scratch[3].ensure(((n + 1) / 2) * 32);
    h2d(ctx, a, leaves, n * 32);
    merkle_root_device(ctx, a, b, n, out_root);
    API_END(ctx)
}

static int32_t column_commit_impl(sezkp_ctx* ctx, const u64* cols_host, const u64* cols_dev, const char* const* labels, int c,
                                  size_t n, int chunk_log2, uint8_t* roots, sezkp_tree** keep) {
    API_BEGIN(ctx)
    REQUIRE((cols_host || cols_dev) && roots && c >= 1, "bad argument");
    REQUIRE(n >= 1 && (n & (n - 1)) == 0, "column length must be a power of two");
    if (keep) *keep = nullptr;
    const size_t bytes = (size_t)c * n * 8;
    const u64* src = cols_dev;
    u64* owned = nullptr;
    if (cols_host) check_canonical(cols_host, (size_t)c * n, "cols");
    if (keep || cols_host) {
        if (keep) owned = (u64*)ctx->pool.alloc(bytes);
        else owned = (u64*)ctx->scratch[3].ensure(bytes);
        if (cols_host) h2d(ctx, owned, cols_host, bytes);
        else CUDA_CHECK(cudaMemcpyAsync(owned, cols_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        src = owned;
    }
    sezkp_tree* t = new sezkp_tree();
    try {
        CommitOpts o;
        o.dedup = true;
        o.roots_host = roots;
        // root-only commit through the plain kernel (value-aware kernels off): the configuration the prover uses for
        // its large FRI layers — a CTA hashes 2^10 leaves and leaves 32 sub-roots to upper_reduce (stark.cu)
        if (!keep && !ctx->dedup_enabled && chunk_log2 == 10 && n >= ((size_t)1 << 20)) {
            o.cta_log2 = 10;
            chunk_log2 = 5;
        }
        commit_build(ctx, t->cm, src, n, c, chunk_log2, labels, o);
    } catch (...) {
        t->cm.release(ctx);
        if (keep && owned) ctx->pool.free(owned);
        delete t;
        throw;
    }
    if (keep) {
        t->cm.owns_values = true;
        *keep = t;
    } else {
        t->cm.release(ctx);
        delete t;
    }
    API_END(ctx)
}
int32_t sezkp_column_commit_batch(sezkp_ctx* ctx, const uint64_t* cols, const char* const* labels, int c, size_t n,
                                  int chunk_log2, uint8_t* roots, sezkp_tree** keep) {
    return column_commit_impl(ctx, cols, nullptr, labels, c, n, chunk_log2, roots, keep);
}
int32_t sezkp_column_commit_batch_dev(sezkp_ctx* ctx, const uint64_t* cols_dev, const char* const* labels, int c, size_t n,
                                      int chunk_log2, uint8_t* roots, sezkp_tree** keep) {
    return column_commit_impl(ctx, nullptr, cols_dev, labels, c, n, chunk_log2, roots, keep);
}

int32_t sezkp_column_open(sezkp_ctx* ctx, const sezkp_tree* tree, const uint32_t* col_idx, const uint64_t* row_idx, size_t k,
                          uint64_t* values, uint8_t* chunk_roots, uint8_t* path_in_chunk, uint8_t* path_to_chunk, int* depth_in,
                          int* depth_out) {
    API_BEGIN(ctx)
    REQUIRE(tree && depth_in && depth_out, "bad argument");
    *depth_in = tree->cm.cl;
    *depth_out = ilog2(tree->cm.n_ch);
    if (k) {
        REQUIRE(col_idx && row_idx && values && chunk_roots && (path_in_chunk || *depth_in == 0) && (path_to_chunk || *depth_out == 0),
                "bad argument");
        u8 dummy[32];
        commit_open(ctx, tree->cm, col_idx, row_idx, k, values, chunk_roots, path_in_chunk ? path_in_chunk : dummy,
                    path_to_chunk ? path_to_chunk : dummy);
    }
    API_END(ctx)
}
int32_t sezkp_verify_openings(sezkp_ctx* ctx, const uint8_t* col_roots, const char* const* labels_or_null, int c, const uint32_t* col_idx,
                              const uint64_t* values, const uint64_t* index_in_chunk, const uint64_t* chunk_index,
                              const uint8_t* chunk_roots, const uint8_t* path_in_chunk, int depth_in, const uint8_t* path_to_chunk,
                              int depth_out, size_t k, uint8_t* ok) {
    API_BEGIN(ctx)
    REQUIRE(col_roots && c >= 1 && depth_in >= 0 && depth_in <= 64 && depth_out >= 0 && depth_out <= 64, "bad argument");
    if (k) {
        REQUIRE(values && index_in_chunk && ok && (path_in_chunk || depth_in == 0) && (depth_out == 0 || (path_to_chunk && chunk_index)),
                "bad argument");
        REQUIRE(col_idx || c == 1, "col_idx is required with more than one column");
        verify_paths_device(ctx, col_roots, labels_or_null, c, col_idx, values, index_in_chunk, chunk_index, chunk_roots, path_in_chunk,
                            depth_in, path_to_chunk, depth_out, k, ok);
    }
    API_END(ctx)
}
void sezkp_tree_free(sezkp_ctx* ctx, sezkp_tree* tree) {
    if (!tree) return;
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tree->cm.release(ctx);
    delete tree;
}

/* ------------------------------------------------------------------ FRI ------ */
static int32_t fri_commit_impl(sezkp_ctx* ctx, const u64* l0_host, const u64* l0_dev, int log_N, const uint64_t* betas,
                               uint8_t* roots, uint64_t* final_value, sezkp_fri** keep) {
    API_BEGIN(ctx)
    REQUIRE((l0_host || l0_dev) && betas && roots && final_value && log_N >= 1 && log_N <= 32, "bad argument");
    if (keep) *keep = nullptr;
    const size_t N = (size_t)1 << log_N;
    check_canonical(betas, (size_t)log_N, "betas");
    const u64* src = l0_dev;
    if (l0_host) {
        check_canonical(l0_host, N, "layer0");
        u64* d = (u64*)ctx->scratch[5].ensure(N * 8);
        h2d(ctx, d, l0_host, N * 8);
        src = d;
    }
    sezkp_fri* f = new sezkp_fri();
    try {
        fri_commit_device(ctx, f->fl, src, log_N, betas, roots, final_value, nullptr);
    } catch (...) {
        f->fl.release(ctx);
        delete f;
        throw;
    }
    if (keep) *keep = f;
    else {
        f->fl.release(ctx);
        delete f;
    }
    API_END(ctx)
}
int32_t sezkp_fri_commit(sezkp_ctx* ctx, const uint64_t* layer0, int log_N, const uint64_t* betas, uint8_t* roots,
                         uint64_t* final_value, sezkp_fri** keep) {
    return fri_commit_impl(ctx, layer0, nullptr, log_N, betas, roots, final_value, keep);
}
int32_t sezkp_fri_commit_dev(sezkp_ctx* ctx, const uint64_t* layer0_dev, int log_N, const uint64_t* betas, uint8_t* roots,
                             uint64_t* final_value, sezkp_fri** keep) {
    return fri_commit_impl(ctx, nullptr, layer0_dev, log_N, betas, roots, final_value, keep);
}
int32_t sezkp_fri_open(sezkp_ctx* ctx, const sezkp_fri* fri, const uint64_t* idx0, size_t k, uint64_t* positions, uint64_t* values,
                       uint8_t* paths) {
    API_BEGIN(ctx)
    REQUIRE(fri && (k == 0 || (idx0 && positions && values && paths)), "bad argument");
    if (k) fri_open_device(ctx, fri->fl, idx0, k, positions, values, paths);
    API_END(ctx)
}
void sezkp_fri_free(sezkp_ctx* ctx, sezkp_fri* fri) {
    if (!fri) return;
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    fri->fl.release(ctx);
    delete fri;
}

/* --------------------------------------------------- feeder (columns + AIR) -- */
int32_t sezkp_trace_columns(sezkp_ctx* ctx, const sezkp_trace_desc* trace, uint64_t* out) {
    API_BEGIN(ctx)
    REQUIRE(out != nullptr, "bad argument");
    validate_trace(trace);
    DeviceTraceOwner dt;
    dt.buf = ctx->scratch[2];
    ctx->scratch[2] = DevBuf();
    try {
        dt.upload(ctx, trace);
    } catch (...) {
        ctx->scratch[2] = dt.buf;
        throw;
    }
    ctx->scratch[2] = dt.buf;
    const size_t count = (size_t)(3 + 7 * trace->tau) * trace->n_rows;
    u64* cols = (u64*)ctx->scratch[3].ensure(count * 8);
    expand_columns_device(ctx, dt.t, cols);
    d2h(ctx, out, cols, count * 8);
    API_END(ctx)
}

int32_t sezkp_compose_base(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint64_t alphas8[8], const uint64_t* mask_coeffs,
                           size_t mask_deg, uint64_t* out) {
    API_BEGIN(ctx)
    REQUIRE(out && alphas8 && (mask_coeffs || mask_deg == 0), "bad argument");
    validate_trace(trace);
    DeviceTraceOwner dt;
    dt.buf = ctx->scratch[2];
    ctx->scratch[2] = DevBuf();
    try {
        dt.upload(ctx, trace);
    } catch (...) {
        ctx->scratch[2] = dt.buf;
        throw;
    }
    ctx->scratch[2] = dt.buf;
    const size_t n = trace->n_rows, count = (size_t)(3 + 7 * trace->tau) * n;
    u64* cols = (u64*)ctx->scratch[3].ensure(count * 8);
    expand_columns_device(ctx, dt.t, cols);
    u64* base = (u64*)ctx->scratch[4].ensure(n * 8);
    compose_device(ctx, cols, n, trace->tau, alphas8, mask_coeffs, mask_deg, base);
    d2h(ctx, out, base, n * 8);
    API_END(ctx)
}

/* --------------------------------------------------------------- prover ------ */
static void deliver(const ProofSink& proof, uint8_t* buf, size_t cap, size_t* len) {
    *len = proof.len;
    if (!buf) return;  // size query
    if (cap < proof.len) sezkp_fail(SEZKP_CUDA_ERANGE, "proof buffer too small: need %zu bytes, have %zu", proof.len, cap);
}

int32_t sezkp_stark_v1_prove(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint8_t manifest_root[32], uint8_t* proof_buf,
                             size_t cap, size_t* len) {
    API_BEGIN(ctx)
    REQUIRE(manifest_root && len, "bad argument");
    if (ctx->group && trace && 3 + 7 * (int)trace->tau >= ctx->group->world) {
        // one proof over all GPUs of the group: columns c % world per GPU, FRI hashing split by chunk range, every exchange
        // inside the library (group.cu); rank 0 serialises into the caller's buffer, the other ranks only count bytes
        validate_trace(trace);
        std::vector<size_t> lens(ctx->group->world, 0);
        group_run(ctx->group, [&](int r, sezkp_ctx* cx) {
            ShardInfo sh{r, cx->group->world, group_allgather_host, &cx->group->ranks[r], group_gather_root_host};
            ProofSink sink(r == 0 ? proof_buf : nullptr, r == 0 ? cap : 0);
            prove_v1_device(cx, trace, manifest_root, sink, &sh);
            lens[r] = sink.len;
        });
        ProofSink proof(proof_buf, cap);
        proof.len = lens[0];
        deliver(proof, proof_buf, cap, len);
        return SEZKP_CUDA_OK;
    }
    ProofSink proof(proof_buf, cap);
    prove_v1_device(ctx, trace, manifest_root, proof);
    deliver(proof, proof_buf, cap, len);
    API_END(ctx)
}

int32_t sezkp_stark_v1_prove_sharded(sezkp_ctx* ctx, const sezkp_trace_desc* trace, const uint8_t manifest_root[32], int rank, int world,
                                     sezkp_allgather_fn allgather, void* user, uint8_t* proof_buf, size_t cap, size_t* len) {
    API_BEGIN(ctx)
    REQUIRE(manifest_root && len && world >= 1 && rank >= 0 && rank < world, "bad argument");
    REQUIRE(world == 1 || allgather != nullptr, "allgather callback is NULL");
    ShardInfo sh{rank, world, allgather, user};
    ProofSink proof(proof_buf, cap);
    prove_v1_device(ctx, trace, manifest_root, proof, world > 1 ? &sh : nullptr);
    deliver(proof, proof_buf, cap, len);
    API_END(ctx)
}

int32_t sezkp_cuda_set_allgather_dev(sezkp_ctx* ctx, sezkp_allgather_dev_fn fn, void* user) {
    if (!ctx) return SEZKP_CUDA_EINVAL;
    ctx->allgather_dev = fn;
    ctx->allgather_dev_user = user;
    return SEZKP_CUDA_OK;
}

int32_t sezkp_trace_upload(sezkp_ctx* ctx, const sezkp_trace_desc* trace, sezkp_trace_dev** out) {
    API_BEGIN(ctx)
    REQUIRE(out != nullptr, "bad argument");
    *out = nullptr;
    validate_trace(trace);
    sezkp_trace_dev* t = new sezkp_trace_dev();
    try {
        t->peers.resize(group_world(ctx) - 1);
        on_all_gpus(ctx, [&](int r, sezkp_ctx* cx) { (r == 0 ? t->owner : t->peers[r - 1]).upload(cx, trace); });
    } catch (...) {
        sezkp_trace_free(ctx, t);
        throw;
    }
    *out = t;
    API_END(ctx)
}
void sezkp_trace_free(sezkp_ctx* ctx, sezkp_trace_dev* t) {
    if (!t) return;
    if (ctx) cudaSetDevice(ctx->device);
    t->owner.buf.release();
    for (size_t r = 0; r < t->peers.size(); r++) {
        if (ctx && ctx->group) cudaSetDevice(ctx->group->ctx[r + 1]->device);
        t->peers[r].buf.release();
    }
    if (ctx) cudaSetDevice(ctx->device);
    delete t;
}
int32_t sezkp_stark_v1_prove_resident(sezkp_ctx* ctx, const sezkp_trace_dev* trace, const uint8_t manifest_root[32],
                                      uint8_t* proof_buf, size_t cap, size_t* len) {
    API_BEGIN(ctx)
    REQUIRE(trace && manifest_root && len, "bad argument");
    if (ctx->group && (int)trace->peers.size() == ctx->group->world - 1 && 3 + 7 * (int)trace->owner.t.tau >= ctx->group->world) {
        std::vector<size_t> lens(ctx->group->world, 0);
        group_run(ctx->group, [&](int r, sezkp_ctx* cx) {
            ShardInfo sh{r, cx->group->world, group_allgather_host, &cx->group->ranks[r], group_gather_root_host};
            ProofSink sink(r == 0 ? proof_buf : nullptr, r == 0 ? cap : 0);
            prove_v1_resident(cx, r == 0 ? trace->owner.t : trace->peers[r - 1].t, manifest_root, sink, &sh);
            lens[r] = sink.len;
        });
        ProofSink proof(proof_buf, cap);
        proof.len = lens[0];
        deliver(proof, proof_buf, cap, len);
        return SEZKP_CUDA_OK;
    }
    ProofSink proof(proof_buf, cap);
    prove_v1_resident(ctx, trace->owner.t, manifest_root, proof);
    deliver(proof, proof_buf, cap, len);
    API_END(ctx)
}

int32_t sezkp_stark_v1_prove_resident_sharded(sezkp_ctx* ctx, const sezkp_trace_dev* trace, const uint8_t manifest_root[32], int rank,
                                              int world, sezkp_allgather_fn allgather, void* user, uint8_t* proof_buf, size_t cap,
                                              size_t* len) {
    API_BEGIN(ctx)
    REQUIRE(trace && manifest_root && len && world >= 1 && rank >= 0 && rank < world, "bad argument");
    REQUIRE(world == 1 || allgather != nullptr, "allgather callback is NULL");
    ShardInfo sh{rank, world, allgather, user};
    ProofSink proof(proof_buf, cap);
    prove_v1_resident(ctx, trace->owner.t, manifest_root, proof, world > 1 ? &sh : nullptr);
    deliver(proof, proof_buf, cap, len);
    API_END(ctx)
}

int32_t sezkp_stark_v1_begin(sezkp_ctx* ctx, uint32_t tau, const uint8_t manifest_root[32], uint64_t expected_rows, sezkp_stream** out) {
    API_BEGIN(ctx)
    REQUIRE(out && manifest_root && tau >= 1 && tau <= 4096 && expected_rows <= (1ULL << 29), "bad argument");
    *out = stream_begin(ctx, tau, manifest_root, expected_rows);
    API_END(ctx)
}
int32_t sezkp_stark_v1_ingest(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* blocks) {
    API_BEGIN(ctx)
    if (!st) sezkp_fail(SEZKP_CUDA_ESTATE, "stream handle is NULL");
    stream_ingest(ctx, st, blocks);
    API_END(ctx)
}
int32_t sezkp_stark_v1_finish(sezkp_ctx* ctx, sezkp_stream* st, uint8_t* proof_buf, size_t cap, size_t* len) {
    API_BEGIN(ctx)
    if (!st) sezkp_fail(SEZKP_CUDA_ESTATE, "stream handle is NULL");
    REQUIRE(proof_buf && len, "finish needs a proof buffer (use sezkp_stark_v1_proof_bound for its size)");
    ProofSink proof(proof_buf, cap);
    stream_finish(ctx, st, proof);
    deliver(proof, proof_buf, cap, len);
    stream_free(ctx, st);  // the handle is consumed once the proof has been delivered
    API_END(ctx)
}
void sezkp_stark_v1_abort(sezkp_ctx* ctx, sezkp_stream* st) {
    if (!ctx || !st) return;
    cudaSetDevice(ctx->device);
    stream_free(ctx, st);
}

/* ---- native JSONL front-end (jsonl.cpp) ---- */
struct sezkp_jsonl_trace {
    jsonl::Trace t;
};
static thread_local std::string g_jsonl_error;
static int jsonl_threads(int n) {
    if (n > 0) return n;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}
const char* sezkp_jsonl_last_error(void) { return g_jsonl_error.c_str(); }
int32_t sezkp_jsonl_parse(const char* text, size_t len, int n_threads, sezkp_jsonl_trace** out, sezkp_trace_desc* desc,
                          const sezkp_block_scalars** scalars) {
    if (!out || !desc || (!text && len)) {
        g_jsonl_error = "bad argument";
        return SEZKP_CUDA_EINVAL;
    }
    *out = nullptr;
    sezkp_jsonl_trace* h = new (std::nothrow) sezkp_jsonl_trace();
    if (!h) return SEZKP_CUDA_ENOMEM;
    try {
        jsonl::parse(text, len, jsonl_threads(n_threads), 0, 1, h->t);
    } catch (const std::bad_alloc&) {
        delete h;
        g_jsonl_error = "host allocation failed";
        return SEZKP_CUDA_ENOMEM;
    } catch (const std::exception& e) {
        delete h;
        g_jsonl_error = e.what();
        return SEZKP_CUDA_EINVAL;
    }
    h->t.fill_desc(*desc);
    if (scalars) *scalars = h->t.manifest.data();
    *out = h;
    return SEZKP_CUDA_OK;
}
void sezkp_jsonl_free(sezkp_jsonl_trace* t) { delete t; }
int32_t sezkp_jsonl_write_file(const char* path, const sezkp_trace_desc* trace, const sezkp_block_scalars* scalars, int n_threads,
                               uint64_t* bytes_written) {
    if (!path || !trace) {
        g_jsonl_error = "bad argument";
        return SEZKP_CUDA_EINVAL;
    }
    try {
        if (!trace->block_len || !trace->win_left || !trace->win_right || !trace->head_in_off || !trace->head_out_off || !trace->input_mv ||
            !trace->mv || !trace->write_flag || !trace->write_sym || trace->tau < 1)
            throw std::runtime_error("trace descriptor has NULL arrays");
        uint64_t sum = 0;
        for (uint64_t k = 0; k < trace->n_blocks; k++) sum += trace->block_len[k];
        if (sum != trace->n_rows) throw std::runtime_error("n_rows != sum(block_len)");
        const size_t n = jsonl::write_file(path, *trace, scalars, jsonl_threads(n_threads));
        if (bytes_written) *bytes_written = n;
    } catch (const std::bad_alloc&) {
        g_jsonl_error = "host allocation failed";
        return SEZKP_CUDA_ENOMEM;
    } catch (const std::exception& e) {
        g_jsonl_error = e.what();
        return SEZKP_CUDA_EINVAL;
    }
    return SEZKP_CUDA_OK;
}

// One piece of JSONL text parsed on the host threads: the workers' outputs in file order (not concatenated).
struct ParsedPiece {
    std::vector<jsonl::Trace> parts;
    size_t lines = 0;
    u32 tau = 0;
    double parse_ms = 0;
    std::string error;      // non-empty: parsing failed (reported by the consumer, in file order)
    bool oom = false;
};
static void jsonl_parse_piece(const char* text, size_t len, int n_threads, u32 tau_hint, size_t first_line, ParsedPiece& out,
                              jsonl::WorkerPool* pool = nullptr) {
    const auto t0 = std::chrono::steady_clock::now();
    try {
        out.tau = tau_hint;
        out.lines = jsonl::parse_parts(text, len, jsonl_threads(n_threads), tau_hint, first_line, out.parts, out.tau, pool);
    } catch (const std::bad_alloc&) {
        out.oom = true;
    } catch (const std::exception& e) {
        out.error = e.what();
    }
    out.parse_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
static void jsonl_ingest_piece(sezkp_ctx* ctx, sezkp_stream*& st, const uint8_t* manifest_root, uint64_t expected_rows, ParsedPiece& pc,
                               uint64_t* blocks, uint64_t* rows, const HostParallelFor* par = nullptr) {
    if (pc.oom) throw std::bad_alloc();
    if (!pc.error.empty()) sezkp_fail(SEZKP_CUDA_EINVAL, "%s", pc.error.c_str());
    // all workers' outputs of the piece go in as ONE ingest (file order): validated together, packed into the staging ring
    // on several threads when the caller provides them
    std::vector<sezkp_trace_desc> descs;
    descs.reserve(pc.parts.size());
    for (auto& t : pc.parts) {
        if (t.block_len.empty()) continue;
        sezkp_trace_desc d;
        t.tau = pc.tau;
        t.fill_desc(d);
        descs.push_back(d);
        if (blocks) *blocks += t.block_len.size();
        if (rows) *rows += t.input_mv.size();
    }
    if (!descs.empty()) {
        if (!st) {
            REQUIRE(manifest_root != nullptr, "internal: stream not started");
            st = stream_begin(ctx, pc.tau, manifest_root, expected_rows);
        }
        stream_ingest_parts(ctx, st, descs.data(), descs.size(), par);
    }
    for (auto& t : pc.parts) t.clear_keep_capacity();  // the next piece parsed into these slots reuses the arrays
}
// parse, then ingest (the one-shot text entry point)
static void jsonl_parse_and_ingest(sezkp_ctx* ctx, sezkp_stream*& st, const uint8_t* manifest_root, uint64_t expected_rows,
                                   const char* text, size_t len, int n_threads, size_t first_line, size_t* lines, uint64_t* blocks,
                                   uint64_t* rows, double* parse_ms) {
    ParsedPiece pc;
    jsonl_parse_piece(text, len, n_threads, st ? stream_tau(st) : 0, first_line, pc);
    if (lines) *lines = pc.lines;
    if (parse_ms) *parse_ms += pc.parse_ms;
    jsonl_ingest_piece(ctx, st, manifest_root, expected_rows, pc, blocks, rows);
}
int32_t sezkp_stark_v1_ingest_jsonl(sezkp_ctx* ctx, sezkp_stream* st, const char* text, size_t len, int n_threads,
                                    uint64_t* n_blocks, uint64_t* n_rows) {
    API_BEGIN(ctx)
    if (!st) sezkp_fail(SEZKP_CUDA_ESTATE, "stream handle is NULL");
    REQUIRE(text || !len, "text is NULL");
    uint64_t nb = 0, nr = 0;
    jsonl_parse_and_ingest(ctx, st, nullptr, 0, text, len, n_threads, 1, nullptr, &nb, &nr, nullptr);
    if (n_blocks) *n_blocks = nb;
    if (n_rows) *n_rows = nr;
    API_END(ctx)
}
int32_t sezkp_stark_v1_prove_jsonl_file(sezkp_ctx* ctx, const char* path, const uint8_t manifest_root[32], int n_threads,
                                        size_t chunk_bytes, uint64_t expected_rows, uint8_t* proof_buf, size_t cap,
                                        size_t* len) {
    API_BEGIN(ctx)
    REQUIRE(path && manifest_root && proof_buf && len, "bad argument");
    REQUIRE(expected_rows <= (1ULL << 29), "expected_rows too large");
    if (chunk_bytes == 0) chunk_bytes = (size_t)64 << 20;
    FILE* f = std::fopen(path, "rb");
    if (!f) sezkp_fail(SEZKP_CUDA_EINVAL, "cannot open %s", path);
    sezkp_stream* st = nullptr;
    auto ms_now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double read_ms = 0, parse_ms = 0, pack_ms = 0;
    size_t total_bytes = 0, line_no = 1;
    // Two text buffers: while the parser threads work on one, a reader thread fills the other (fread of the next piece).
    // A piece is cut at its last newline; the tail is carried to the front of the next piece.
    struct Piece {
        std::unique_ptr<char[]> mem;
        size_t cap = 0, len = 0;
        bool eof = false, failed = false;
        void need(size_t n, size_t keep) {  // grow without zero-filling; the first `keep` bytes are preserved
            if (n <= cap) return;
            std::unique_ptr<char[]> m(new char[n]);
            if (keep) std::memcpy(m.get(), mem.get(), keep);
            mem = std::move(m);
            cap = n;
        }
    } piece[2];
    auto fill = [&](Piece& p, size_t carry) {  // p.mem[0, carry) already holds the carried tail
        p.need(carry + chunk_bytes, carry);
        const double t0 = ms_now();
        const size_t got = std::fread(p.mem.get() + carry, 1, chunk_bytes, f);
        read_ms += ms_now() - t0;
        p.failed = got < chunk_bytes && std::ferror(f);
        p.eof = got < chunk_bytes;
        p.len = carry + got;
        total_bytes += got;
    };
    std::thread reader, unmapper;
    // state the parser thread writes: declared outside the try block so that it outlives the thread on every unwind path
    // (the handler joins `reader` before these go out of scope)
    ParsedPiece pc[2];
    std::unique_ptr<jsonl::WorkerPool> pool;  // used by `reader`: must outlive it on every unwind path as well
    std::vector<std::pair<size_t, size_t>> pieces;  // (offset, length) of the mapped file's pieces, cut at newlines
    void* map = MAP_FAILED;
    size_t map_len = 0;
    try {
        // Regular files are mapped: the parser threads read the page cache directly (no copy, page faults spread over
        // the workers).  Pipes / special files fall back to the double-buffered fread loop below.
        struct stat sb;
        if (fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
            const double t0 = ms_now();
            map_len = (size_t)sb.st_size;
            map = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fileno(f), 0);
            if (map != MAP_FAILED) madvise(map, map_len, MADV_WILLNEED);  // the parser threads read disjoint ranges concurrently: no SEQUENTIAL hint
            read_ms += ms_now() - t0;
        }
        if (map != MAP_FAILED) {
            // Two-stage pipeline over pieces of the mapped file: while this thread packs piece k into the pinned staging ring
            // (stream_ingest: memcpy + H2D submission), the parser threads already work on piece k + 1.
            const char* text = (const char*)map;
            for (size_t pos = 0; pos < map_len;) {
                size_t use = std::min(chunk_bytes, map_len - pos);
                if (pos + use < map_len) {  // cut at the last newline of the piece (or extend to the next one)
                    size_t e = use;
                    while (e > 0 && text[pos + e - 1] != '\n') e--;
                    if (e == 0) {
                        const char* nl = (const char*)std::memchr(text + pos + use, '\n', map_len - pos - use);
                        e = nl ? (size_t)(nl - (text + pos)) + 1 : map_len - pos;
                    }
                    use = e;
                }
                pieces.push_back({pos, use});
                pos += use;
            }
            u32 tau_known = 0;
            // One pool of host threads for the whole file (not 32 thread creations per piece), used in two steps per piece:
            // the workers parse their runs of lines, then the same threads pack what they have just produced — still hot in
            // their caches — into the pinned staging ring.  (An earlier version packed piece k on this thread while the pool
            // parsed piece k+1; with every core busy parsing, the packing threads only ran when the scheduler pre-empted a
            // parser, and packing — 4 GB/s from one thread — had become the bound of the T = 2^26 path: 0.51 s of 0.60 s.)
            if (jsonl_threads(n_threads) > 1) pool.reset(new jsonl::WorkerPool(jsonl_threads(n_threads)));
            jsonl::WorkerPool* const pl = pool.get();
            HostParallelFor pack_par = [&](int tasks, const std::function<void(int)>& fn) { pl->run(tasks, fn); };
            const HostParallelFor* const cp = pl ? &pack_par : nullptr;
            for (size_t k = 0; k < pieces.size(); k++) {
                ParsedPiece& p = pc[0];
                {  // fresh state, but the workers' output arrays of the previous piece are kept for reuse
                    std::vector<jsonl::Trace> keep = std::move(p.parts);
                    p = ParsedPiece();
                    p.parts = std::move(keep);
                }
                jsonl_parse_piece(text + pieces[k].first, pieces[k].second, n_threads, tau_known, line_no, p, pl);
                parse_ms += p.parse_ms;
                line_no += p.lines;
                if (p.error.empty() && !p.oom && p.tau) tau_known = p.tau;
                const double t_pack = ms_now();
                jsonl_ingest_piece(ctx, st, manifest_root, expected_rows, p, nullptr, nullptr, cp);
                pack_ms += ms_now() - t_pack;
            }
            total_bytes = map_len;
        } else {
        int cur = 0;
        fill(piece[cur], 0);
        for (;;) {
            Piece& p = piece[cur];
            if (p.failed) sezkp_fail(SEZKP_CUDA_EINVAL, "read error on %s", path);
            size_t use = p.len;
            if (!p.eof) {
                while (use > 0 && p.mem[use - 1] != '\n') use--;
                if (use == 0) {  // a single line longer than the piece: read on into the same buffer
                    chunk_bytes *= 2;
                    const size_t have = p.len;
                    fill(p, have);
                    continue;
                }
            }
            Piece& nx = piece[cur ^ 1];
            const size_t carry = p.len - use;
            const bool more = !p.eof;
            if (more) {  // start reading the next piece while this one is parsed
                nx.need(carry + chunk_bytes, 0);
                if (carry) std::memcpy(nx.mem.get(), p.mem.get() + use, carry);
                reader = std::thread([&fill, &nx, carry] { fill(nx, carry); });
            }
            size_t nl = 0;
            jsonl_parse_and_ingest(ctx, st, manifest_root, expected_rows, p.mem.get(), use, n_threads, line_no, &nl, nullptr, nullptr, &parse_ms);
            line_no += nl;
            if (!more) break;
            reader.join();
            cur ^= 1;
        }
        }
        // The mapping is torn down on a helper thread behind the proof: unmapping a large file is serial page-table work
        // (tmpfs: ~40 ns per 4 KiB page, 0.16 s for the 13 GB file of BASELINE configs[4]) that nothing below depends on.
        if (map != MAP_FAILED) {
            void* const m = map;
            const size_t ml = map_len;
            FILE* const ff = f;
            map = MAP_FAILED;
            f = nullptr;
            unmapper = std::thread([m, ml, ff] {
                munmap(m, ml);
                std::fclose(ff);
            });
        } else {
            std::fclose(f);
            f = nullptr;
        }
        if (!st) sezkp_fail(SEZKP_CUDA_EINVAL, "%s holds no blocks", path);
        ProofSink proof(proof_buf, cap);
        stream_finish(ctx, st, proof);
        deliver(proof, proof_buf, cap, len);
        stream_free(ctx, st);
        st = nullptr;
        if (unmapper.joinable()) unmapper.join();
        ctx->timings.insert(ctx->timings.begin(), {"jsonl_bytes", (double)total_bytes});
        ctx->timings.insert(ctx->timings.begin(), {"jsonl_pack_ms", pack_ms});
        ctx->timings.insert(ctx->timings.begin(), {"jsonl_parse_ms", parse_ms});
        ctx->timings.insert(ctx->timings.begin(), {"jsonl_read_ms", read_ms});
    } catch (...) {
        if (reader.joinable()) reader.join();
        if (unmapper.joinable()) unmapper.join();
        if (map != MAP_FAILED) munmap(map, map_len);
        if (f) std::fclose(f);
        if (st) stream_free(ctx, st);
        throw;
    }
    API_END(ctx)
}
/* upper bound of the ProofV1 size for a trace of n_rows rows and tau tapes */
size_t sezkp_stark_v1_proof_bound(uint64_t n_rows, uint32_t tau) {
    size_t ln = 1;
    while ((1ULL << ln) < n_rows) ln++;
    const size_t lN = ln + 3;
    const size_t openings = 30 * (9 * (size_t)tau + 3) * (8 + 24 + 32 + 16 + 32 * ln);
    const size_t fri = 30 * (16 + 8 * (lN + 1) + lN * 2 * (8 + 8 + 32 * lN));
    return 4096 + (3 + 7 * (size_t)tau) * 64 + openings + fri + 32 * (lN + 1);
}

}  // extern "C"
