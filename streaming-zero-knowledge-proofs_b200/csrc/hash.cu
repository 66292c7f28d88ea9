// BLAKE3 leaf hashing and chunked Merkle commitments (sm_100a).
// Replaces hash_field_leaves{,_labeled} (reference v1/merkle.rs:132-159), MerkleTree::from_leaves / open
// (v1/merkle.rs:46-108), OnDemandOpenings::{build_roots, open} (v1/openings.rs:306-497) and
// StreamingLayerBuilder (v1/fri_stream.rs:55-122).  For a power-of-two number of leaves the chunked
// tree (2^10-leaf chunk trees under an outer tree of chunk roots) is the plain binary tree, so one
// kernel reduces a whole chunk in shared memory and only chunk roots and the levels above are stored.
#include <cstring>

#include "hash.cuh"

namespace {

constexpr int HASH_THREADS = 256;
constexpr int MAX_CL = 10;

// Shared-memory digest array in word-major layout: word w of digest i at s[w*pitch + i]; a thread reads
// (left.w, right.w) of pair i as one 64-bit load at index 2i.
__device__ __forceinline__ void pair_parent(const u32* s, int pitch, int i, u32 (&out)[8]) {
    u32 l[8], r[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const uint2 p = *reinterpret_cast<const uint2*>(&s[w * pitch + 2 * i]);
        l[w] = p.x;
        r[w] = p.y;
    }
    b3::parent(l, r, out);
}
__device__ __forceinline__ void put_digest(u32* s, int pitch, int i, const u32 (&d)[8], u32* gdst) {
#pragma unroll
    for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
    if (gdst) {
        uint4* d4 = reinterpret_cast<uint4*>(gdst);
        d4[0] = make_uint4(d[0], d[1], d[2], d[3]);
        d4[1] = make_uint4(d[4], d[5], d[6], d[7]);
    }
}
// Reduces `count` (power of two, <= 1024) digests held in s to one (left in slot 0).  Optionally stores every
// produced level into the retained upper-tree array (`levels_base`, level l0+1.. of a column with n_ch_total
// chunk roots, this group being number `grp` at level l0) and/or records the sibling path of leaf `path_idx`.
__device__ __forceinline__ void reduce_levels_smem(u32* s, int pitch, int count, u32* levels_base, u64 grp, int l0,
                                                   u64 n_ch_total, u32* path_out, int path_idx) {
    int lvl = 0;
    while (count > 1) {
        if (path_out != nullptr && threadIdx.x < 8) {
            const int sib = (path_idx >> lvl) ^ 1;
            path_out[lvl * 8 + threadIdx.x] = s[threadIdx.x * pitch + sib];
        }
        const int half = count >> 1;
        const int i0 = threadIdx.x, i1 = threadIdx.x + HASH_THREADS;  // half <= 512 = 2*HASH_THREADS
        u32 d0[8], d1[8];
        if (i0 < half) pair_parent(s, pitch, i0, d0);
        if (i1 < half) pair_parent(s, pitch, i1, d1);
        __syncthreads();
        u32* gl = nullptr;
        if (levels_base) {
            const int l = l0 + lvl + 1;
            gl = levels_base + ((2 * n_ch_total - ((2 * n_ch_total) >> l)) + grp * (u64)half) * 8;
        }
        if (i0 < half) put_digest(s, pitch, i0, d0, gl ? gl + (u64)i0 * 8 : nullptr);
        if (i1 < half) put_digest(s, pitch, i1, d1, gl ? gl + (u64)i1 * 8 : nullptr);
        __syncthreads();
        count = half;
        lvl++;
    }
}

// values -> leaf hashes -> chunk root.  grid (n_ch, cols).  Writes chunk roots into level 0 of `upper`.
__global__ void __launch_bounds__(HASH_THREADS) chunk_commit_kernel(const u64* __restrict__ values, u64 n, int cl,
                                                                    const b3::LabelTemplate* __restrict__ templates,
                                                                    u32* __restrict__ upper, u64 n_ch) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 chunk = blockIdx.x;
    const int col = blockIdx.y;
    const int leaves = 1 << cl;
    const u64* v = values + (u64)col * n + (chunk << cl);
    b3::LabelTemplate t;
    if (templates) t = templates[col];
    if (templates) {
        B3_DISPATCH_LABELED(t, {
            for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
                u32 d[8];
                b3::leaf_labeled_w<B3W>(t, v[i], d);
#pragma unroll
                for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
            }
        })
    } else {
        for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
            u32 d[8];
            b3::leaf(v[i], d);
#pragma unroll
            for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
        }
    }
    __syncthreads();
    reduce_levels_smem(s, pitch, leaves, nullptr, 0, 0, 0, nullptr, 0);
    if (threadIdx.x < 8) upper[((u64)col * (2 * n_ch - 1) + chunk) * 8 + threadIdx.x] = s[threadIdx.x * pitch];
}

// digests at level l0 of `upper` -> reduce groups of 2^k -> levels l0+1..l0+k stored.  grid (count>>k, cols)
__global__ void __launch_bounds__(HASH_THREADS) upper_reduce_kernel(u32* __restrict__ upper, u64 n_ch, int l0, int k) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 grp = blockIdx.x;
    const int col = blockIdx.y;
    u32* base = upper + (u64)col * (2 * n_ch - 1) * 8;
    const u32* src = base + ((2 * n_ch - ((2 * n_ch) >> l0)) + (grp << k)) * 8;
    const int cnt = 1 << k;
    for (int i = threadIdx.x; i < cnt * 8; i += HASH_THREADS) s[(i & 7) * pitch + (i >> 3)] = src[i];
    __syncthreads();
    reduce_levels_smem(s, pitch, cnt, base, grp, l0, n_ch, nullptr, 0);
}

struct OpenReq {
    u32 col;
    u32 pad;
    u64 row;
};
// One CTA per opening: rebuild the chunk, record the in-chunk sibling path and chunk root, gather the upper path.
__global__ void __launch_bounds__(HASH_THREADS) open_kernel(const u64* __restrict__ values, u64 n, int cl,
                                                            const b3::LabelTemplate* __restrict__ templates,
                                                            const u32* __restrict__ upper, u64 n_ch, int depth_out,
                                                            const OpenReq* __restrict__ reqs, u64* __restrict__ out_values,
                                                            u32* __restrict__ out_chunk_roots, u32* __restrict__ out_path_in,
                                                            u32* __restrict__ out_path_to) {
    __shared__ __align__(16) u32 s[8 * ((1 << MAX_CL) + 2)];
    const int pitch = (1 << MAX_CL) + 2;
    const u64 q = blockIdx.x;
    const OpenReq rq = reqs[q];
    const u64 chunk = rq.row >> cl;
    const int idx_in = (int)(rq.row & ((1ULL << cl) - 1));
    const int leaves = 1 << cl;
    const u64* v = values + (u64)rq.col * n + (chunk << cl);
    b3::LabelTemplate t;
    if (templates) t = templates[rq.col];
    if (templates) {
        B3_DISPATCH_LABELED(t, {
            for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
                u32 d[8];
                b3::leaf_labeled_w<B3W>(t, v[i], d);
#pragma unroll
                for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
            }
        })
    } else {
        for (int i = threadIdx.x; i < leaves; i += HASH_THREADS) {
            u32 d[8];
            b3::leaf(v[i], d);
#pragma unroll
            for (int w = 0; w < 8; w++) s[w * pitch + i] = d[w];
        }
    }
    if (threadIdx.x == 0) out_values[q] = v[idx_in];
    __syncthreads();
    reduce_levels_smem(s, pitch, leaves, nullptr, 0, 0, 0, out_path_in + q * (u64)cl * 8, idx_in);
    if (threadIdx.x < 8) out_chunk_roots[q * 8 + threadIdx.x] = s[threadIdx.x * pitch];
    const u32* base = upper + (u64)rq.col * (2 * n_ch - 1) * 8;
    for (int i = threadIdx.x; i < depth_out * 8; i += HASH_THREADS) {
        const int l = i >> 3, w = i & 7;
        const u64 sib = (chunk >> l) ^ 1;
        out_path_to[q * (u64)depth_out * 8 + i] = base[((2 * n_ch - ((2 * n_ch) >> l)) + sib) * 8 + w];
    }
}

__global__ void leaf_hash_kernel(const u64* __restrict__ vals, size_t n, const b3::LabelTemplate* __restrict__ tpl,
                                 u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 d[8];
    if (tpl) {
        b3::LabelTemplate t = *tpl;
        B3_DISPATCH_LABELED(t, { b3::leaf_labeled_w<B3W>(t, vals[i], d); })
    } else b3::leaf(vals[i], d);
    uint4* o = reinterpret_cast<uint4*>(out + i * 8);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

// One Merkle level with odd promotion: out[i] = H(in[2i], in[2i+1]) or in[2i] when 2i+1 == n.
__global__ void merkle_level_kernel(const u32* __restrict__ in, size_t n, u32* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_out = (n + 1) / 2;
    if (i >= n_out) return;
    const uint4* p = reinterpret_cast<const uint4*>(in + 16 * i);
    uint4 a0 = p[0], a1 = p[1];
    u32 d[8];
    if (2 * i + 1 < n) {
        uint4 b0 = p[2], b1 = p[3];
        u32 l[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        u32 r[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        b3::parent(l, r, d);
    } else {
        d[0] = a0.x; d[1] = a0.y; d[2] = a0.z; d[3] = a0.w;
        d[4] = a1.x; d[5] = a1.y; d[6] = a1.z; d[7] = a1.w;
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 8);
    o[0] = make_uint4(d[0], d[1], d[2], d[3]);
    o[1] = make_uint4(d[4], d[5], d[6], d[7]);
}

}  // namespace

b3::LabelTemplate make_label_template(const char* label) {
    const size_t L = std::strlen(label);
    REQUIRE(L <= 44, "column label longer than 44 bytes (labeled leaf must fit one BLAKE3 block)");
    u8 bytes[64] = {0};
    std::memcpy(bytes, "col_leaf", 8);  // reference v1/params.rs:58 DS_COL_LEAF
    const u32 l32 = (u32)L;
    std::memcpy(bytes + 8, &l32, 4);
    std::memcpy(bytes + 12, label, L);
    b3::LabelTemplate t;
    std::memcpy(t.words, bytes, 64);
    t.off = (u32)(12 + L);
    t.block_len = (u32)(20 + L);
    return t;
}

void Commit::release(sezkp_ctx* ctx) {
    if (owns_values && values) ctx->pool.free((void*)values);
    ctx->pool.free(upper);
    ctx->pool.free(templates);
    values = nullptr;
    upper = nullptr;
    templates = nullptr;
}

void commit_build(sezkp_ctx* ctx, Commit& cm, const u64* values_dev, u64 n, int cols, int chunk_log2,
                  const char* const* labels, u8* roots_host) {
    REQUIRE(n >= 1 && (n & (n - 1)) == 0, "column length %llu is not a power of two", (unsigned long long)n);
    REQUIRE(cols >= 1 && cols <= 65535, "column count %d out of range", cols);
    REQUIRE(chunk_log2 >= 0 && chunk_log2 <= MAX_CL, "chunk_log2 %d out of range (0..10)", chunk_log2);
    const int ln = ilog2(n);
    cm.values = values_dev;
    cm.n = n;
    cm.cols = cols;
    cm.cl = ln < chunk_log2 ? ln : chunk_log2;
    cm.n_ch = n >> cm.cl;
    REQUIRE(cm.n_ch < (1ULL << 31), "too many chunks per column");
    if (labels) {
        std::vector<b3::LabelTemplate> h(cols);
        for (int c = 0; c < cols; c++) {
            REQUIRE(labels[c] != nullptr, "label %d is NULL", c);
            h[c] = make_label_template(labels[c]);
        }
        cm.templates = (b3::LabelTemplate*)ctx->pool.alloc(sizeof(b3::LabelTemplate) * cols);
        CUDA_CHECK(cudaMemcpyAsync(cm.templates, h.data(), sizeof(b3::LabelTemplate) * cols, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // h goes out of scope
    }
    const size_t upper_bytes = (size_t)cols * (2 * cm.n_ch - 1) * 32;
    cm.upper = (u32*)ctx->pool.alloc(upper_bytes);
    dim3 grid((unsigned)cm.n_ch, (unsigned)cols);
    chunk_commit_kernel<<<grid, HASH_THREADS, 0, ctx->stream>>>(values_dev, n, cm.cl, cm.templates, cm.upper, cm.n_ch);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    const int depth = ilog2(cm.n_ch);
    int l0 = 0;
    while (l0 < depth) {
        const int k = (depth - l0) < MAX_CL ? (depth - l0) : MAX_CL;
        dim3 g((unsigned)(cm.n_ch >> (l0 + k)), (unsigned)cols);
        upper_reduce_kernel<<<g, HASH_THREADS, 0, ctx->stream>>>(cm.upper, cm.n_ch, l0, k);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        l0 += k;
    }
    if (roots_host) {
        CUDA_CHECK(cudaMemcpy2DAsync(roots_host, 32, (const u8*)cm.upper + (2 * cm.n_ch - 2) * 32, (2 * cm.n_ch - 1) * 32, 32, cols,
                                     cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
}

void commit_open(sezkp_ctx* ctx, const Commit& cm, const u32* col_idx, const u64* row_idx, size_t k, u64* values,
                 u8* chunk_roots, u8* path_in, u8* path_to) {
    if (k == 0) return;
    const int depth_out = ilog2(cm.n_ch);
    std::vector<OpenReq> reqs(k);
    for (size_t i = 0; i < k; i++) {
        REQUIRE(col_idx[i] < (u32)cm.cols, "opening %zu: column %u out of range", i, col_idx[i]);
        REQUIRE(row_idx[i] < cm.n, "opening %zu: row %llu out of range", i, (unsigned long long)row_idx[i]);
        reqs[i] = OpenReq{col_idx[i], 0, row_idx[i]};
    }
    const size_t b_req = k * sizeof(OpenReq), b_val = k * 8, b_cr = k * 32, b_in = k * (size_t)cm.cl * 32,
                 b_to = k * (size_t)depth_out * 32;
    u8* d = (u8*)ctx->scratch[7].ensure(b_req + b_val + b_cr + b_in + b_to + 64);
    OpenReq* d_req = (OpenReq*)d;
    u64* d_val = (u64*)(d + b_req);
    u32* d_cr = (u32*)(d + b_req + b_val);
    u32* d_in = (u32*)(d + b_req + b_val + b_cr);
    u32* d_to = (u32*)(d + b_req + b_val + b_cr + b_in);
    CUDA_CHECK(cudaMemcpyAsync(d_req, reqs.data(), b_req, cudaMemcpyHostToDevice, ctx->stream));
    open_kernel<<<(unsigned)k, HASH_THREADS, 0, ctx->stream>>>(cm.values, cm.n, cm.cl, cm.templates, cm.upper, cm.n_ch, depth_out,
                                                               d_req, d_val, d_cr, d_in, d_to);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
    CUDA_CHECK(cudaMemcpyAsync(values, d_val, b_val, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaMemcpyAsync(chunk_roots, d_cr, b_cr, cudaMemcpyDeviceToHost, ctx->stream));
    if (b_in) CUDA_CHECK(cudaMemcpyAsync(path_in, d_in, b_in, cudaMemcpyDeviceToHost, ctx->stream));
    if (b_to) CUDA_CHECK(cudaMemcpyAsync(path_to, d_to, b_to, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}

void leaf_hash_device(sezkp_ctx* ctx, const u64* vals_dev, size_t n, const char* label, u32* out_dev) {
    if (n == 0) return;
    b3::LabelTemplate* d_t = nullptr;
    if (label) {
        b3::LabelTemplate t = make_label_template(label);
        d_t = (b3::LabelTemplate*)ctx->scratch[6].ensure(sizeof t);
        CUDA_CHECK(cudaMemcpyAsync(d_t, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    leaf_hash_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(vals_dev, n, d_t, out_dev);
    CUDA_CHECK(cudaGetLastError());
    ctx->launches++;
}

void merkle_root_device(sezkp_ctx* ctx, u32* level, u32* tmp, size_t n, u8* root_host) {
    u32 *a = level, *b = tmp;
    while (n > 1) {
        const size_t n_out = (n + 1) / 2;
        merkle_level_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, ctx->stream>>>(a, n, b);
        CUDA_CHECK(cudaGetLastError());
        ctx->launches++;
        u32* t = a;
        a = b;
        b = t;
        n = n_out;
    }
    CUDA_CHECK(cudaMemcpyAsync(root_host, a, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
}
