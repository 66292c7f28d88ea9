// Multi-GPU context group (see group.cuh).
#include "group.cuh"

#include <chrono>
#include <cstring>

void GroupBarrier::reset(int w) {
    std::lock_guard<std::mutex> lk(mu);
    world = w;
    count.store(0);
    failed.store(false);
}
void GroupBarrier::wait() {
    if (failed.load(std::memory_order_acquire)) sezkp_fail(SEZKP_CUDA_ECOMM, "another GPU of the group failed");
    const u64 my = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == world) {  // last arrival releases the others
        count.store(0, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu);  // pairs with the sleepers' predicate check
            gen.store(my + 1, std::memory_order_release);
        }
        cv.notify_all();
        return;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int spins = 0;; spins++) {
        if (gen.load(std::memory_order_acquire) != my) return;
        if (failed.load(std::memory_order_acquire)) break;
        if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(50)) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return gen.load(std::memory_order_acquire) != my || failed.load(std::memory_order_acquire); });
            if (gen.load(std::memory_order_acquire) != my) return;
            break;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    sezkp_fail(SEZKP_CUDA_ECOMM, "another GPU of the group failed");
}
void GroupBarrier::fail() {
    {
        std::lock_guard<std::mutex> lk(mu);
        failed.store(true, std::memory_order_release);
    }
    cv.notify_all();
}

// one rank's share of a group job; errors are recorded, never thrown (a failing rank fails the barrier so that the others unwind)
static void run_rank(sezkp_group* g, int rank, const std::function<void(int)>& job) {
    int32_t rc = SEZKP_CUDA_OK;
    std::string msg;
    try {
        CUDA_CHECK(cudaSetDevice(g->ctx[rank]->device));
        job(rank);
    } catch (const SezkpError& e) {
        rc = e.code;
        msg = e.what();
    } catch (const std::bad_alloc&) {
        rc = SEZKP_CUDA_ENOMEM;
        msg = "host allocation failed";
    } catch (const std::exception& e) {
        rc = SEZKP_CUDA_ECUDA;
        msg = e.what();
    }
    if (rc != SEZKP_CUDA_OK) {
        cudaGetLastError();
        g->bar.fail();
    }
    g->rc[rank] = rc;  // own slot; read by group_run after the completion hand-shake
    g->err[rank] = msg;
}

static void worker_main(sezkp_group* g, int rank) {
    u64 seen = 0;
    for (;;) {
        std::function<void(int)> job;
        // back-to-back calls (a prover loop): the next job usually arrives within microseconds of the last one — poll for it
        // briefly before sleeping on the condition variable (a wake-up costs ~50 us per rank, a tenth of a sharded proof phase)
        const auto t0 = std::chrono::steady_clock::now();
        for (int spins = 0; g->job_seq.load(std::memory_order_acquire) == seen && !g->quit_flag.load(std::memory_order_acquire); spins++) {
            if ((spins & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->quit || g->job_seq.load(std::memory_order_acquire) != seen; });
            if (g->quit) return;
            seen = g->job_seq.load(std::memory_order_acquire);
            job = g->job;
        }
        run_rank(g, rank, job);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// Rank 0's share runs on the calling thread (it is the longest: rank 0 serialises the proof), ranks 1.. on their workers: the
// caller is not woken up at the end and one dispatch wake-up disappears.
void group_run(sezkp_group* g, const std::function<void(int, sezkp_ctx*)>& fn) {
    g->bar.reset(g->world);
    const std::function<void(int)> job = [g, &fn](int r) { fn(r, g->ctx[r]); };
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = job;
        g->pending = g->world - 1;
        g->job_seq.fetch_add(1, std::memory_order_release);
    }
    if (g->world > 1) g->cv_job.notify_all();
    run_rank(g, 0, job);
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    cudaSetDevice(g->ctx[0]->device);
    int bad = -1;
    for (int r = 0; r < g->world; r++)
        if (g->rc[r] != SEZKP_CUDA_OK && (bad < 0 || (g->rc[bad] == SEZKP_CUDA_ECOMM && g->rc[r] != SEZKP_CUDA_ECOMM))) bad = r;
    if (bad >= 0) sezkp_fail(g->rc[bad], "GPU %d (rank %d of %d): %s", g->ctx[bad]->device, bad, g->world, g->err[bad].c_str());
}

sezkp_group* group_create(const int* device_ids, int n_dev) {
    REQUIRE(device_ids != nullptr && n_dev >= 1 && n_dev <= 64, "bad device list");
    // a device may be listed more than once: its ranks then share that GPU (how the group logic is tested on one GPU)
    sezkp_group* g = new sezkp_group();
    g->world = n_dev;
    try {
        for (int i = 0; i < n_dev; i++) {
            sezkp_ctx* c = nullptr;
            const int32_t rc = sezkp_cuda_create(device_ids[i], &c);
            if (rc != SEZKP_CUDA_OK) sezkp_fail(rc, "device %d: %s", device_ids[i], sezkp_cuda_last_error(nullptr));
            g->ctx.push_back(c);
        }
        g->ranks.resize(n_dev);
        g->send.assign(n_dev, nullptr);
        g->root_slots.assign(n_dev, nullptr);
        g->rc.assign(n_dev, 0);
        g->err.assign(n_dev, "");
        g->ev_ready.assign(n_dev, nullptr);
        g->ev_done.assign(n_dev, nullptr);
        g->p2p = true;
        for (int i = 0; i < n_dev; i++) {
            CUDA_CHECK(cudaSetDevice(device_ids[i]));
            CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_ready[i], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&g->ev_done[i], cudaEventDisableTiming));
            for (int j = 0; j < n_dev; j++) {
                if (device_ids[i] == device_ids[j]) continue;
                int can = 0;
                CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]));
                if (!can) {
                    g->p2p = false;
                    continue;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) {
                    cudaGetLastError();
                    g->p2p = false;
                }
            }
        }
        for (int i = 0; i < n_dev; i++) {
            g->ranks[i] = GroupRank{g, i};
            g->ctx[i]->group = g;
            g->ctx[i]->group_rank = i;
            g->ctx[i]->allgather_dev = group_allgather_dev;
            g->ctx[i]->allgather_dev_user = &g->ranks[i];
        }
        CUDA_CHECK(cudaSetDevice(device_ids[0]));
        for (int i = 1; i < n_dev; i++) g->workers.emplace_back(worker_main, g, i);  // rank 0 runs on the caller's thread
    } catch (...) {
        group_destroy(g);
        throw;
    }
    return g;
}

void group_destroy(sezkp_group* g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->quit = true;
        g->quit_flag.store(true, std::memory_order_release);
    }
    g->cv_job.notify_all();
    for (auto& t : g->workers)
        if (t.joinable()) t.join();
    for (size_t i = 0; i < g->ctx.size(); i++) {
        cudaSetDevice(g->ctx[i]->device);
        if (i < g->ev_ready.size() && g->ev_ready[i]) cudaEventDestroy(g->ev_ready[i]);
        if (i < g->ev_done.size() && g->ev_done[i]) cudaEventDestroy(g->ev_done[i]);
    }
    for (sezkp_ctx* c : g->ctx) {
        c->group = nullptr;  // plain single-GPU teardown from here on
        sezkp_cuda_destroy(c);
    }
    delete g;
}

int32_t group_allgather_host(void* user, const void* send, size_t bytes, void* recv_all) {
    GroupRank* gr = (GroupRank*)user;
    sezkp_group* g = gr->g;
    try {
        g->send[gr->rank] = send;
        g->bar.wait();
        for (int s = 0; s < g->world; s++) std::memcpy((u8*)recv_all + (size_t)s * bytes, g->send[s], bytes);
        g->bar.wait();  // nobody reuses its send buffer before every rank has copied it
    } catch (const SezkpError&) {
        return -1;
    }
    return 0;
}

int32_t group_gather_root_host(void* user, const void* send, const void** all_ptrs) {
    GroupRank* gr = (GroupRank*)user;
    sezkp_group* g = gr->g;
    try {
        g->root_slots[gr->rank] = send;  // a slot array of its own: the buffers stay published after the barrier
        g->bar.wait();
        for (int s = 0; s < g->world; s++) all_ptrs[s] = g->root_slots[s];
    } catch (const SezkpError&) {
        return -1;
    }
    return 0;
}

// Pull model: every rank copies each peer's block into its own receive buffer on its own stream, after that peer's
// "ready" event; a second event per rank ("done") keeps a peer from overwriting its send buffer while others still pull.
// Two host barriers (thread rendezvous, microseconds), no host-device synchronisation.
int32_t group_allgather_dev(void* user, const void* send_dev, size_t bytes, void* recv_all_dev, void* cuda_stream) {
    GroupRank* gr = (GroupRank*)user;
    sezkp_group* g = gr->g;
    const int rank = gr->rank, world = g->world;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    try {
        g->send[rank] = send_dev;
        CUDA_CHECK(cudaEventRecord(g->ev_ready[rank], st));
        g->bar.wait();
        for (int k = 0; k < world; k++) {
            const int s = (rank + k) % world;  // start with the own block, then a different peer per rank: spreads the NVLink load
            u8* dst = (u8*)recv_all_dev + (size_t)s * bytes;
            if (s == rank) {
                if ((const void*)dst != send_dev) CUDA_CHECK(cudaMemcpyAsync(dst, send_dev, bytes, cudaMemcpyDeviceToDevice, st));
            } else {
                CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_ready[s], 0));
                CUDA_CHECK(cudaMemcpyPeerAsync(dst, g->ctx[rank]->device, g->send[s], g->ctx[s]->device, bytes, st));
            }
        }
        CUDA_CHECK(cudaEventRecord(g->ev_done[rank], st));
        g->bar.wait();
        for (int s = 0; s < world; s++)
            if (s != rank) CUDA_CHECK(cudaStreamWaitEvent(st, g->ev_done[s], 0));
    } catch (const SezkpError&) {
        cudaGetLastError();
        return -1;
    }
    return 0;
}

void group_publish_peers(GroupRank* gr, const void* mine, cudaStream_t stream, const void** all) {
    sezkp_group* g = gr->g;
    g->send[gr->rank] = mine;
    CUDA_CHECK(cudaEventRecord(g->ev_ready[gr->rank], stream));
    g->bar.wait();
    for (int s = 0; s < g->world; s++) {
        all[s] = g->send[s];
        if (s != gr->rank) CUDA_CHECK(cudaStreamWaitEvent(stream, g->ev_ready[s], 0));
    }
    // the slots (and the ready events) are shared with the other collectives: nobody may enter the next one — and overwrite
    // its slot or re-record its event — before every rank has read all of them
    g->bar.wait();
}
void group_release_peers(GroupRank* gr, cudaStream_t stream) {
    sezkp_group* g = gr->g;
    CUDA_CHECK(cudaEventRecord(g->ev_done[gr->rank], stream));
    g->bar.wait();
    for (int s = 0; s < g->world; s++)
        if (s != gr->rank) CUDA_CHECK(cudaStreamWaitEvent(stream, g->ev_done[s], 0));
}
