// Multi-GPU context group: ONE process drives the GPUs of a box (sezkp_cuda_create_multi), so the reference's
// stateless single-process backend call (sezkp-core/src/backend.rs:41-61) can use all of them.
//
//   * one sezkp_ctx per GPU (rank r = position in the device list); rank 0's share of a call runs on the caller's thread,
//     the other ranks on persistent worker threads;
//   * the collectives of the sharded prover (stark.cu: column roots, FRI subtree roots, opening records, compact trace)
//     are implemented inside the library: host all-gather = shared-memory exchange between the rank threads, device
//     all-gather = cudaMemcpyPeerAsync pulls over NVLink ordered by CUDA events (no host synchronisation), and
//     peer-pointer exchange for kernels that read the other GPUs' HBM directly (wide.cu);
//   * an error on one rank fails the group barrier, so the other ranks unwind instead of waiting for ever.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"

// Sense-reversing barrier of the rank threads.  The waits inside a collective are microseconds long (the ranks run in
// lockstep), and a mutex + condition-variable wake-up costs tens of microseconds per rank — with 8 ranks and a dozen
// collectives per proof that was ~0.6 ms of a 5 ms proof.  So waiters spin on an atomic generation counter first and only
// fall back to the condition variable after ~50 us (a rank that is still busy hashing its share).
struct GroupBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int world = 1;
    std::atomic<int> count{0};
    std::atomic<u64> gen{0};
    std::atomic<bool> failed{false};
    void reset(int w);
    void wait();  // throws SezkpError(ECOMM) once the barrier has been failed
    void fail();
};

struct sezkp_group;
struct GroupRank {
    sezkp_group* g;
    int rank;
};

struct sezkp_group {
    int world = 0;
    bool p2p = false;                    // every pair of devices has peer access enabled
    std::vector<sezkp_ctx*> ctx;         // ctx[0] is the handle the caller holds
    std::vector<GroupRank> ranks;
    std::vector<std::thread> workers;
    // job dispatch
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<void(int)> job;
    std::atomic<u64> job_seq{0};
    int pending = 0;
    bool quit = false;
    std::atomic<bool> quit_flag{false};  // the same, for the workers' lock-free polling
    std::vector<int32_t> rc;
    std::vector<std::string> err;
    GroupBarrier bar;
    // exchange slots (valid between two barrier phases of one collective)
    std::vector<const void*> send;
    std::vector<const void*> root_slots;  // group_gather_root_host
    std::vector<cudaEvent_t> ev_ready, ev_done;
};

sezkp_group* group_create(const int* device_ids, int n_dev);  // throws SezkpError
void group_destroy(sezkp_group* g);                           // destroys every member ctx, including ctx[0]
// Run fn(rank, ctx[rank]) on every rank thread and wait; rethrows the first primary error (a rank that merely saw the
// failed barrier reports ECOMM, which is only reported when nothing else went wrong).
void group_run(sezkp_group* g, const std::function<void(int, sezkp_ctx*)>& fn);

// Collectives, callable from inside a group_run job.  Signatures match the callbacks of the sharded prover
// (sezkp_allgather_fn / sezkp_allgather_dev_fn, include/sezkp_cuda.h); `user` is a GroupRank*.
int32_t group_allgather_host(void* user, const void* send, size_t bytes, void* recv_all);
int32_t group_allgather_dev(void* user, const void* send_dev, size_t bytes, void* recv_all_dev, void* cuda_stream);
// Gather to rank 0 without copies (ShardInfo::gather_root): every rank publishes its buffer, one barrier, all_ptrs[s] = rank
// s's buffer.  The buffers must outlive the group call.
int32_t group_gather_root_host(void* user, const void* send, const void** all_ptrs);
// Peer-pointer exchange: publishes `mine` (device memory of this rank, complete on `stream`), returns every rank's pointer
// in all[world] and makes `stream` wait until each of them is complete.  Pair with group_release_peers(), which keeps every
// rank from reusing its buffer before all peers' kernels that read it have finished.
void group_publish_peers(GroupRank* gr, const void* mine, cudaStream_t stream, const void** all);
void group_release_peers(GroupRank* gr, cudaStream_t stream);
