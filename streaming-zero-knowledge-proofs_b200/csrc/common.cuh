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

struct NttTables;  // ntt.cu

struct sezkp_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string last_error;
    std::map<u64, NttTables*> ntt_tables;  // key: (log_n, inverse, coset params)
    DevBuf scratch[8];                      // reusable work buffers (per purpose, see users)
    std::vector<std::pair<std::string, double>> timings;  // phase -> ms (last prove)
    u64 launches = 0;                       // kernels launched since last reset
};

inline int ilog2(u64 n) {
    int k = 0;
    while ((1ULL << k) < n) k++;
    return k;
}
