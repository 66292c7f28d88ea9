// Shared host-side plumbing for libsezkp_cuda: context, error reporting, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sezkp_cuda.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

struct SezkpError : std::runtime_error {
    int32_t code;
    SezkpError(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void sezkp_fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw SezkpError(code, buf);
}
#define CUDA_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            sezkp_fail(SEZKP_CUDA_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define REQUIRE(cond, ...)                                 \
    do {                                                   \
        if (!(cond)) sezkp_fail(SEZKP_CUDA_EINVAL, __VA_ARGS__); \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            cap = 0;
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) {
                p = nullptr;
                sezkp_fail(SEZKP_CUDA_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            }
            cap = bytes;
        }
        return p;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

// Grow-only pinned host buffer (staging for small H2D/D2H transfers: pageable copies go through the driver's bounce
// buffer at a fraction of the PCIe rate and block the host).
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    void* ensure(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) {
                p = nullptr;
                sezkp_fail(SEZKP_CUDA_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            }
            cap = bytes;
        }
        return p;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct NttTables;  // ntt.cu
struct OpenReq;    // hash.cuh
struct sezkp_group;  // group.cuh

// Size-keyed caching device allocator: repeated proofs reuse their buffers instead of paying
// cudaMalloc / cudaFree (which synchronises the device) on every call.
struct DevPool {
    std::map<size_t, std::vector<void*>> free_lists;
    std::map<void*, size_t> live;
    size_t cached_bytes = 0;
    void* alloc(size_t bytes) {
        if (bytes == 0) bytes = 256;
        bytes = (bytes + 255) & ~(size_t)255;
        auto it = free_lists.find(bytes);
        void* p = nullptr;
        if (it != free_lists.end() && !it->second.empty()) {
            p = it->second.back();
            it->second.pop_back();
            cached_bytes -= bytes;
        } else {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e != cudaSuccess) {
                trim();  // give cached blocks back and retry once
                cudaGetLastError();
                e = cudaMalloc(&p, bytes);
            }
            if (e != cudaSuccess) {
                cudaGetLastError();
                sezkp_fail(SEZKP_CUDA_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            }
        }
        live[p] = bytes;
        return p;
    }
    void free(void* p) {
        if (!p) return;
        auto it = live.find(p);
        if (it == live.end()) {
            cudaFree(p);
            return;
        }
        free_lists[it->second].push_back(p);
        cached_bytes += it->second;
        live.erase(it);
    }
    void trim() {
        for (auto& kv : free_lists)
            for (void* p : kv.second) cudaFree(p);
        free_lists.clear();
        cached_bytes = 0;
    }
};

struct sezkp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;          // side stream for slab uploads (created on first use)
    std::vector<cudaEvent_t> slab_events;
    std::vector<cudaEvent_t> phase_events;      // phase clock of the prover (timing events, created on first use)
    std::string last_error;
    std::map<u64, NttTables*> ntt_tables;  // key: (log_n, inverse, coset params)
    std::map<std::pair<int, u64>, u64*> deep_tables;  // (log N, shift) -> device table of the first coset point of every DEEP CTA (capped)
    std::map<int, u64*> power_tables;      // log_n -> device table of the low powers of w_n (composition mask)
    DevPool pool;
    DevBuf scratch[12];                      // reusable work buffers (per purpose, see users)
    std::vector<OpenReq>* open_reqs = nullptr;  // request list of the prover's single opening launch, reused across proofs
    std::vector<u8> host_stage[2];            // grow-only pageable staging (opening-record exchange of the sharded prover)
    PinnedBuf pinned[2];                     // host staging: [0] opening requests / results
    PinnedBuf stream_stage[3];               // staging ring of the streaming ingest, kept across streams (cudaHostAlloc of ~100 MB costs ~50 ms)
    bool stream_stage_busy = false;          // one stream at a time borrows the ring; a second concurrent stream allocates its own
    std::vector<std::pair<std::string, double>> timings;  // phase -> ms (last prove)
    int dedup_variant = 2;                  // 1: 256-thread kernel, 2: 128-thread kernel (more chunks in flight per SM)
    bool dedup_enabled = true;              // value-aware column commit (SEZKP_NO_DEDUP=1 or sezkp_cuda_set_option disables)
    bool tabled_enabled = true;             // subtree tables for structured columns (option "tabled"; needs dedup)
    u64 tab_redone_chunks = 0;              // chunks the tabled pass handed back to the generic kernel (last commit)
    int tab_columns = 0;                    // columns served from subtree tables (last commit)
    bool tab_cache_enabled = true;           // option "tab_cache"
    std::vector<u8> tab_cache_key;           // subtree tables in scratch[11] were built for exactly these (ColTab[], label templates)
    int ntt_gen = 3;                        // option "ntt_gen": pass-kernel generation (1 = round-1 kernel; 2-4 = fused global I/O, see ntt.cu kernel_for_bits)
    bool lde_fuse = true;                   // option "lde_fuse" (K7): fuse the LDE's last pass with the labeled leaf hash in lde_commit / lde_commit_fri
    bool phase_sync = true;                 // option "phase_sync" (default 1): phase clock = host time with a stream synchronisation per phase; 0 = CUDA events read back at the end (no added syncs, but measured 0.3-0.4 ms SLOWER end to end on the slab-pipelined path, tools/e2e_ab.py)
    bool fri_coset = true;                  // option "fri_coset": coset-resident FRI layers for one proof over a context group (0: all-gather layer 0, replicated folds)
    bool deep_fused = false;                // option "deep_fused": one-launch DEEP kernel (per-CTA inversion) also for large domains
    std::map<const void*, size_t> func_smem;  // kernel -> dynamic shared memory already granted on this ctx's device
    u64 launches = 0;                       // kernels launched since last reset
    sezkp_group* group = nullptr;           // non-null: member of a multi-GPU group (sezkp_cuda_create_multi); ctx[0] is the caller's handle
    int group_rank = 0;
    sezkp_allgather_dev_fn allgather_dev = nullptr;  // optional device-side collective of the sharded prover
    void* allgather_dev_user = nullptr;
};

// Table upload from pageable host memory, ordered before everything later launched on ctx->stream.
// A plain cudaMemcpy is NOT enough: for pageable sources it returns once the data sits in the driver's staging buffer
// and the DMA runs on the legacy stream, which the library's non-blocking stream does not wait for — the first
// kernel after a table build could read a half-written table.
inline void upload_table(sezkp_ctx* ctx, void* dst, const void* src, size_t bytes) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

inline int ilog2(u64 n) {
    int k = 0;
    while ((1ULL << k) < n) k++;
    return k;
}
