// Streaming ingest: the push API of the reference's ProvingBackendStream
// (crates/sezkp-core/src/prover.rs:21-33: begin_stream / ingest_block / finish_stream; driver loop
// StreamingProver::prove_stream_iter :104-150 over stream_block_summaries_auto, core/io.rs:111-139).
//
// Blocks arrive one at a time from a host-side parser.  Their rows are packed into pinned staging buffers and, each
// time a buffer fills (2^20 rows), copied to the device trace on a dedicated copy stream while the host keeps
// parsing / ingesting the next blocks — so by finish_stream only the tail of the trace is still in flight.
// finish_stream uploads the per-block metadata, joins the copy stream and runs the resident prover.
#include <chrono>
#include <cstring>

#include "group.cuh"
#include "stark.cuh"

namespace {
constexpr size_t STAGE_ROWS = 1 << 20;
constexpr int N_STAGE = 3;
double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

struct sezkp_stream {
    u32 tau = 0;
    u8 manifest_root[32];
    // per-block metadata (small, kept on the host until finish)
    std::vector<u64> block_len;
    std::vector<int64_t> win_left, win_right;
    std::vector<u32> in_off, out_off;
    // device trace arrays (row-indexed), capacity in rows
    size_t cap_rows = 0, rows = 0;
    int8_t* d_input_mv = nullptr;
    int8_t* d_mv = nullptr;
    u8* d_wflag = nullptr;
    uint16_t* d_wsym = nullptr;
    // pinned staging ring
    struct Stage {
        u8* host = nullptr;  // [input_mv | mv | wflag | wsym] for STAGE_ROWS rows
        cudaEvent_t done = nullptr, start = nullptr;
        size_t rows = 0, base_row = 0;
        bool in_flight = false;
    } stage[N_STAGE];
    int cur = 0;
    bool borrowed_ring = false;  // stage[].host belong to ctx->stream_stage
    cudaStream_t copy_stream = nullptr;
    double copy_ms = 0, ingest_t0 = 0, stall_ms = 0;
    size_t h2d_bytes = 0;

    size_t row_bytes() const { return 1 + 4 * (size_t)tau; }
    void free_all(sezkp_ctx* ctx) {
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        for (auto& s : stage) {
            if (s.host && !borrowed_ring) cudaFreeHost(s.host);
            if (s.done) cudaEventDestroy(s.done);
            if (s.start) cudaEventDestroy(s.start);
        }
        ctx->pool.free(d_input_mv);
        ctx->pool.free(d_mv);
        ctx->pool.free(d_wflag);
        ctx->pool.free(d_wsym);
        if (borrowed_ring) ctx->stream_stage_busy = false;
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

static void reserve_rows(sezkp_ctx* ctx, sezkp_stream* st, size_t need) {
    if (need <= st->cap_rows) return;
    size_t cap = st->cap_rows ? st->cap_rows : STAGE_ROWS;
    while (cap < need) cap *= 2;
    const size_t tau = st->tau;
    int8_t* n_imv = (int8_t*)ctx->pool.alloc(cap);
    int8_t* n_mv = (int8_t*)ctx->pool.alloc(cap * tau);
    u8* n_wf = (u8*)ctx->pool.alloc(cap * tau);
    uint16_t* n_ws = (uint16_t*)ctx->pool.alloc(cap * tau * 2);
    if (st->rows) {  // carry over what is already on the device (ordered after the copies in flight)
        CUDA_CHECK(cudaMemcpyAsync(n_imv, st->d_input_mv, st->rows, cudaMemcpyDeviceToDevice, st->copy_stream));
        CUDA_CHECK(cudaMemcpyAsync(n_mv, st->d_mv, st->rows * tau, cudaMemcpyDeviceToDevice, st->copy_stream));
        CUDA_CHECK(cudaMemcpyAsync(n_wf, st->d_wflag, st->rows * tau, cudaMemcpyDeviceToDevice, st->copy_stream));
        CUDA_CHECK(cudaMemcpyAsync(n_ws, st->d_wsym, st->rows * tau * 2, cudaMemcpyDeviceToDevice, st->copy_stream));
        CUDA_CHECK(cudaStreamSynchronize(st->copy_stream));
    }
    ctx->pool.free(st->d_input_mv);
    ctx->pool.free(st->d_mv);
    ctx->pool.free(st->d_wflag);
    ctx->pool.free(st->d_wsym);
    st->d_input_mv = n_imv;
    st->d_mv = n_mv;
    st->d_wflag = n_wf;
    st->d_wsym = n_ws;
    st->cap_rows = cap;
}

// issue the H2D copies of the current staging buffer and move on to the next one
static void flush_stage(sezkp_ctx* ctx, sezkp_stream* st) {
    sezkp_stream::Stage& s = st->stage[st->cur];
    if (s.rows == 0) return;
    const size_t tau = st->tau, r = s.rows, base = s.base_row;
    reserve_rows(ctx, st, base + r);
    const u8* h = s.host;
    CUDA_CHECK(cudaEventRecord(s.start, st->copy_stream));
    CUDA_CHECK(cudaMemcpyAsync(st->d_input_mv + base, h, r, cudaMemcpyHostToDevice, st->copy_stream));
    CUDA_CHECK(cudaMemcpyAsync(st->d_mv + base * tau, h + STAGE_ROWS, r * tau, cudaMemcpyHostToDevice, st->copy_stream));
    CUDA_CHECK(cudaMemcpyAsync(st->d_wflag + base * tau, h + STAGE_ROWS * (1 + tau), r * tau, cudaMemcpyHostToDevice, st->copy_stream));
    CUDA_CHECK(cudaMemcpyAsync(st->d_wsym + base * tau, h + STAGE_ROWS * (1 + 2 * tau), r * tau * 2, cudaMemcpyHostToDevice, st->copy_stream));
    CUDA_CHECK(cudaEventRecord(s.done, st->copy_stream));
    s.in_flight = true;
    st->h2d_bytes += r * st->row_bytes();
    st->rows = base + r;
    st->cur = (st->cur + 1) % N_STAGE;
    sezkp_stream::Stage& nx = st->stage[st->cur];
    if (nx.in_flight) {  // ring wrapped: wait until that buffer's copy has landed before overwriting it
        const double t0 = now_ms();
        CUDA_CHECK(cudaEventSynchronize(nx.done));
        st->stall_ms += now_ms() - t0;
        float ms = 0;
        cudaEventElapsedTime(&ms, nx.start, nx.done);
        st->copy_ms += ms;
        nx.in_flight = false;
    }
    nx.rows = 0;
    nx.base_row = st->rows;
}

sezkp_stream* stream_begin(sezkp_ctx* ctx, u32 tau, const u8 manifest_root[32], u64 expected_rows) {
    sezkp_stream* st = new sezkp_stream();
    st->tau = tau;
    std::memcpy(st->manifest_root, manifest_root, 32);
    try {
        CUDA_CHECK(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
        st->borrowed_ring = !ctx->stream_stage_busy;
        if (st->borrowed_ring) ctx->stream_stage_busy = true;
        for (int i = 0; i < N_STAGE; i++) {
            auto& s = st->stage[i];
            if (st->borrowed_ring) s.host = (u8*)ctx->stream_stage[i].ensure(STAGE_ROWS * st->row_bytes());
            else CUDA_CHECK(cudaHostAlloc((void**)&s.host, STAGE_ROWS * st->row_bytes(), cudaHostAllocDefault));
            CUDA_CHECK(cudaEventCreate(&s.done));
            CUDA_CHECK(cudaEventCreate(&s.start));
        }
        if (expected_rows) reserve_rows(ctx, st, (size_t)expected_rows);
    } catch (...) {
        st->free_all(ctx);
        delete st;
        throw;
    }
    st->ingest_t0 = now_ms();
    return st;
}

void stream_ingest(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* b) { stream_ingest_parts(ctx, st, b, 1, nullptr); }

// `count` descriptors in block order, ingested as one: all are validated before the stream is touched, the per-block metadata
// is appended, and the row arrays are packed into the pinned staging ring — with `par` on several host threads (one task per
// descriptor and ring buffer segment): a single thread packs ~4 GB/s out of arrays other cores have just written, which made
// this copy, not the JSON parser, the bound of the T = 2^26 JSONL path once the parser got faster.
void stream_ingest_parts(sezkp_ctx* ctx, sezkp_stream* st, const sezkp_trace_desc* descs, size_t count, const HostParallelFor* par) {
    REQUIRE(descs && count >= 1, "ingest: no descriptors");
    const size_t tau = st->tau;
    // validate everything before touching the stream: a rejected ingest must leave the handle as it was
    u64 total = 0;
    for (size_t p = 0; p < count; p++) {
        const sezkp_trace_desc* b = &descs[p];
        REQUIRE(b->n_blocks >= 1 && b->tau == st->tau, "ingest: bad block descriptor (tau mismatch or empty)");
        REQUIRE(b->flags == 0, "ingest: packed descriptors are not accepted by the streaming ingest");
        REQUIRE(b->block_len && b->win_left && b->win_right && b->head_in_off && b->head_out_off && b->input_mv && b->mv && b->write_flag &&
                    b->write_sym,
                "ingest: descriptor has NULL arrays");
        u64 rows = 0;
        for (u64 k = 0; k < b->n_blocks; k++) {
            REQUIRE(b->block_len[k] >= 1, "ingest: empty block");
            rows += b->block_len[k];
        }
        REQUIRE(rows == b->n_rows, "ingest: n_rows != sum(block_len)");
        total += rows;
    }
    std::vector<u64> first(count + 1, 0);  // first row of every descriptor within this call
    for (size_t p = 0; p < count; p++) {
        const sezkp_trace_desc* b = &descs[p];
        first[p + 1] = first[p] + b->n_rows;
        st->block_len.insert(st->block_len.end(), b->block_len, b->block_len + b->n_blocks);
        st->win_left.insert(st->win_left.end(), b->win_left, b->win_left + b->n_blocks * tau);
        st->win_right.insert(st->win_right.end(), b->win_right, b->win_right + b->n_blocks * tau);
        st->in_off.insert(st->in_off.end(), b->head_in_off, b->head_in_off + b->n_blocks * tau);
        st->out_off.insert(st->out_off.end(), b->head_out_off, b->head_out_off + b->n_blocks * tau);
    }
    size_t done = 0, p0 = 0;
    while (done < total) {
        sezkp_stream::Stage& s = st->stage[st->cur];
        const size_t take = std::min((size_t)total - done, STAGE_ROWS - s.rows);
        u8* h = s.host;
        const size_t base = s.rows;
        while (first[p0 + 1] <= done) p0++;
        size_t p1 = p0;
        while (p1 < count && first[p1] < done + take) p1++;
        // descriptor p contributes its rows [max(first[p], done), min(first[p+1], done + take)) to this ring buffer
        auto copy_part = [&](int i) {
            const size_t p = p0 + (size_t)i;
            const sezkp_trace_desc* b = &descs[p];
            const size_t lo = std::max<size_t>(first[p], done), hi = std::min<size_t>(first[p + 1], done + take);
            if (hi <= lo) return;
            const size_t src = lo - first[p], dst = base + (lo - done), n = hi - lo;
            std::memcpy(h + dst, b->input_mv + src, n);
            std::memcpy(h + STAGE_ROWS + dst * tau, b->mv + src * tau, n * tau);
            std::memcpy(h + STAGE_ROWS * (1 + tau) + dst * tau, b->write_flag + src * tau, n * tau);
            std::memcpy(h + STAGE_ROWS * (1 + 2 * tau) + dst * tau * 2, b->write_sym + src * tau, n * tau * 2);
        };
        const int tasks = (int)(p1 - p0);
        if (par && tasks > 1 && take * st->row_bytes() >= ((size_t)1 << 20)) (*par)(tasks, copy_part);
        else
            for (int i = 0; i < tasks; i++) copy_part(i);
        s.rows += take;
        done += take;
        if (s.rows == STAGE_ROWS) flush_stage(ctx, st);
    }
}

void stream_finish(sezkp_ctx* ctx, sezkp_stream* st, ProofSink& proof) {
    flush_stage(ctx, st);
    const double ingest_ms = now_ms() - st->ingest_t0;
    const u64 n = st->rows, nb = st->block_len.size();
    REQUIRE(nb >= 1, "finish: no blocks were ingested");
    REQUIRE(n >= 2 && (n & (n - 1)) == 0, "n_rows = %llu: the STARK v1 path needs a power-of-two trace length >= 2",
            (unsigned long long)n);
    REQUIRE(n <= (1ULL << 29), "n_rows too large");
    const size_t tau = st->tau;
    // per-block metadata: one small packed upload on the copy stream
    std::vector<u64> starts(nb);
    u64 acc = 0;
    for (u64 k = 0; k < nb; k++) {
        starts[k] = acc;
        acc += st->block_len[k];
    }
    size_t off = 0;
    auto sect = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 15) & ~(size_t)15;
        return o;
    };
    const size_t o_start = sect(nb * 8), o_len = sect(nb * 8), o_wl = sect(nb * tau * 8), o_wr = sect(nb * tau * 8),
                 o_io = sect(nb * tau * 4), o_oo = sect(nb * tau * 4);
    u8* meta = (u8*)ctx->scratch[2].ensure(off);
    auto put = [&](size_t o, const void* src, size_t bytes) {
        CUDA_CHECK(cudaMemcpyAsync(meta + o, src, bytes, cudaMemcpyHostToDevice, st->copy_stream));
    };
    put(o_start, starts.data(), nb * 8);
    put(o_len, st->block_len.data(), nb * 8);
    put(o_wl, st->win_left.data(), nb * tau * 8);
    put(o_wr, st->win_right.data(), nb * tau * 8);
    put(o_io, st->in_off.data(), nb * tau * 4);
    put(o_oo, st->out_off.data(), nb * tau * 4);
    const double t_wait0 = now_ms();
    CUDA_CHECK(cudaStreamSynchronize(st->copy_stream));  // join: only the tail of the trace can still be in flight
    const double exposed_ms = now_ms() - t_wait0;
    for (auto& s : st->stage)
        if (s.in_flight) {
            float ms = 0;
            cudaEventElapsedTime(&ms, s.start, s.done);
            st->copy_ms += ms;
            s.in_flight = false;
        }
    DeviceTrace t{};
    t.ops = nullptr;
    t.tau = st->tau;
    t.n_blocks = nb;
    t.n_rows = n;
    t.block_start = (const u64*)(meta + o_start);
    t.block_len = (const u64*)(meta + o_len);
    t.win_left = (const int64_t*)(meta + o_wl);
    t.win_right = (const int64_t*)(meta + o_wr);
    t.head_in_off = (const u32*)(meta + o_io);
    t.head_out_off = (const u32*)(meta + o_oo);
    t.input_mv = st->d_input_mv;
    t.mv = st->d_mv;
    t.write_flag = st->d_wflag;
    t.write_sym = st->d_wsym;
    if (ctx->group && 3 + 7 * (int)st->tau >= ctx->group->world) {
        // Context group: the trace was ingested on this GPU only (the host parser is the bound of this path); replicate it to
        // the peers over NVLink and let every GPU prove its share (columns c % world, FRI hashing by chunk range).
        sezkp_group* g = ctx->group;
        const int world = g->world;
        // The peers' copies live in their contexts' grow-only scratch slot 2 (the slot of the one-shot trace upload, idle
        // during a resident prove): cudaMalloc + cudaFree of a 2.2 GB buffer on seven GPUs per call cost ~170 ms at T = 2^26.
        struct Peer {
            DeviceTrace t;
        };
        std::vector<Peer> peers(world - 1);
        const size_t row_off = (off + 255) & ~(size_t)255;
        const size_t o_imv = row_off, o_mv = o_imv + ((n + 15) & ~(size_t)15), o_wf = o_mv + n * tau, o_ws = o_wf + n * tau;
        const size_t total = o_ws + n * tau * 2;
        const double t_rep0 = now_ms();
        try {
            for (int r = 1; r < world; r++) {
                sezkp_ctx* px = g->ctx[r];
                Peer& p = peers[r - 1];
                CUDA_CHECK(cudaSetDevice(px->device));
                u8* base = (u8*)px->scratch[2].ensure(total);
                CUDA_CHECK(cudaSetDevice(ctx->device));
                auto cp = [&](size_t o, const void* src, size_t bytes) {
                    CUDA_CHECK(cudaMemcpyPeerAsync(base + o, px->device, src, ctx->device, bytes, ctx->stream));
                };
                cp(0, meta, off);
                cp(o_imv, st->d_input_mv, n);
                cp(o_mv, st->d_mv, n * tau);
                cp(o_wf, st->d_wflag, n * tau);
                cp(o_ws, st->d_wsym, n * tau * 2);
                p.t = t;
                p.t.block_start = (const u64*)(base + o_start);
                p.t.block_len = (const u64*)(base + o_len);
                p.t.win_left = (const int64_t*)(base + o_wl);
                p.t.win_right = (const int64_t*)(base + o_wr);
                p.t.head_in_off = (const u32*)(base + o_io);
                p.t.head_out_off = (const u32*)(base + o_oo);
                p.t.input_mv = (const int8_t*)(base + o_imv);
                p.t.mv = (const int8_t*)(base + o_mv);
                p.t.write_flag = (const u8*)(base + o_wf);
                p.t.write_sym = (const uint16_t*)(base + o_ws);
            }
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            const double rep_ms = now_ms() - t_rep0;
            group_run(g, [&](int r, sezkp_ctx* cx) {
                ShardInfo sh{r, world, group_allgather_host, &g->ranks[r], group_gather_root_host};
                ProofSink sink(r == 0 ? proof.buf : nullptr, r == 0 ? proof.cap : 0);
                prove_v1_resident(cx, r == 0 ? t : peers[r - 1].t, st->manifest_root, sink, &sh);
                if (r == 0) proof.len = sink.len;
            });
            ctx->timings.insert(ctx->timings.begin(), {"stream_replicate_ms", rep_ms});
        } catch (...) {
            cudaSetDevice(ctx->device);
            throw;
        }
        CUDA_CHECK(cudaSetDevice(ctx->device));
    } else {
        prove_v1_resident(ctx, t, st->manifest_root, proof, nullptr);
    }
    const double hidden = st->copy_ms > 0 ? 1.0 - (exposed_ms + st->stall_ms) / st->copy_ms : 0.0;
    ctx->timings.insert(ctx->timings.begin(), {"stream_copy_hidden_frac", hidden < 0 ? 0.0 : hidden});
    ctx->timings.insert(ctx->timings.begin(), {"stream_copy_exposed_ms", exposed_ms + st->stall_ms});
    ctx->timings.insert(ctx->timings.begin(), {"stream_h2d_copy_ms", st->copy_ms});
    ctx->timings.insert(ctx->timings.begin(), {"stream_ingest_ms", ingest_ms});
}

u32 stream_tau(const sezkp_stream* st) { return st->tau; }

void stream_free(sezkp_ctx* ctx, sezkp_stream* st) {
    if (!st) return;
    st->free_all(ctx);
    delete st;
}
