// Single-process multi-GPU driver over the C ABI + NCCL (no Python, no torch): one handle per
// GPU, one host thread.  Two modes, the two multi-GPU shapes of the north star:
//   default        restart chains sharded across the GPUs (chain k of GPU g has global id
//                  g * chains + k = its Philox stream); every --exchange steps the packed best
//                  keys (score << 32 | global chain id) are min-all-reduced in place on the device
//                  pointers and the elite board is broadcast from the GPU that owns it
//   --partitioned  ONE large instance, every GPU holds a replica and scans a triangular-balanced
//                  slice of the swap neighbourhood; per step one 8-byte min-all-reduce of the packed
//                  (delta, i, j) key, every replica applies the same move -- no state crosses NVLink
// Build: make -C examples/cpp nqueens_multi_gpu   (needs nccl.h / libnccl and the CUDA runtime)
#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cs_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)
#define NK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { std::fprintf(stderr, "NCCL %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); std::exit(1); } } while (0)
#define CS(h, x) do { int32_t s_ = (x); if (s_ != CS_OK) { std::fprintf(stderr, "%s: %s (%s)\n", #x, cs_status_string(s_), (h) ? cs_nq_last_error(h) : ""); std::exit(1); } } while (0)

int main(int argc, char** argv) {
    uint32_t n = 10000, chains = 1024, steps = 8, exchange = 4;
    int gpus = cs_device_count();
    bool partitioned = false;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        auto val = [&]() { if (k + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", a.c_str()); std::exit(2); } return std::atoll(argv[++k]); };
        if (a == "-b" || a == "--board-size") n = (uint32_t)val();
        else if (a == "--chains") chains = (uint32_t)val();
        else if (a == "--steps") steps = (uint32_t)val();
        else if (a == "--exchange") exchange = (uint32_t)val();
        else if (a == "--gpus") gpus = (int)val();
        else if (a == "--partitioned") partitioned = true;
        else { std::fprintf(stderr, "usage: %s [--board-size N] [--chains C] [--steps S] [--exchange K] [--gpus G] [--partitioned]\n", argv[0]); return 2; }
    }
    if (gpus < 1) { std::fprintf(stderr, "fatal: no CUDA device (this library has no CPU fallback)\n"); return 101; }
    const int G = gpus;
    std::vector<int> devs(G);
    for (int g = 0; g < G; ++g) devs[g] = g;
    std::vector<ncclComm_t> comm(G);
    NK(ncclCommInitAll(comm.data(), G, devs.data()));
    std::vector<cudaStream_t> stream(G);
    std::vector<cs_nq_handle*> h(G, nullptr);
    for (int g = 0; g < G; ++g) {
        CK(cudaSetDevice(g));
        CK(cudaStreamCreateWithFlags(&stream[g], cudaStreamNonBlocking));
        cs_nq_config cfg{};
        cfg.n = n;
        cfg.n_chains = partitioned ? 1u : chains;
        cfg.chain_offset = partitioned ? 0u : (uint32_t)g * chains;  // replicas share stream 0; shards get their own ids
        cfg.seed = 42;
        cfg.device = g;
        cfg.neighbourhood = CS_NQ_SWAP;
        cfg.flags = partitioned ? CS_NQ_FLAG_GLOBAL : 0u;
        CS(h[g], cs_nq_create(&cfg, &h[g]));
        CS(h[g], cs_nq_set_stream(h[g], stream[g]));
        CS(h[g], cs_nq_init_random(h[g]));
        if (partitioned) CS(h[g], cs_nq_set_partition(h[g], (uint32_t)g, (uint32_t)G));
    }
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long moves = 0;
    long long best = -1;
    if (partitioned) {
        std::vector<void*> key(G);
        for (int g = 0; g < G; ++g) CS(h[g], cs_nq_part_key_device_ptr(h[g], &key[g]));
        for (uint32_t s = 0; s < steps; ++s) {
            for (int g = 0; g < G; ++g) CS(h[g], cs_nq_part_scan(h[g]));  // enqueued; the GPUs scan concurrently
            NK(ncclGroupStart());
            for (int g = 0; g < G; ++g) NK(ncclAllReduce(key[g], key[g], 1, ncclInt64, ncclMin, comm[g], stream[g]));
            NK(ncclGroupEnd());
            for (int g = 0; g < G; ++g) {
                cs_step_stats st{};
                CS(h[g], cs_nq_part_apply(h[g], &st));
                moves += st.moves_scored;
                best = st.best_score;
            }
        }
        // every replica holds the same board
        std::vector<int64_t> r0(n), rg(n);
        CS(h[0], cs_nq_get_chains(h[0], 0, 1, r0.data()));
        for (int g = 1; g < G; ++g) {
            CS(h[g], cs_nq_get_chains(h[g], 0, 1, rg.data()));
            if (std::memcmp(r0.data(), rg.data(), sizeof(int64_t) * n) != 0) { std::fprintf(stderr, "replica %d diverged\n", g); return 1; }
        }
    } else {
        std::vector<void*> key(G);
        std::vector<void*> elite(G);
        uint32_t stride = 0;
        for (int g = 0; g < G; ++g) {
            CS(h[g], cs_nq_best_key_device_ptr(h[g], &key[g]));
            void* p = nullptr;
            CS(h[g], cs_nq_chain_device_ptr(h[g], 0, &p, &stride));
            CK(cudaSetDevice(g));
            CK(cudaMalloc(&elite[g], (size_t)stride * sizeof(uint16_t)));
        }
        for (uint32_t s = 0; s < steps; ++s) {
            // cs_nq_step returns when its GPU is done, so the GPUs are driven by one thread each
            // (a handle is Send-not-Sync: one thread at a time, any thread)
            std::vector<cs_step_stats> st(G);
            std::vector<std::thread> th;
            for (int g = 0; g < G; ++g)
                th.emplace_back([&, g] { CK(cudaSetDevice(g)); CS(h[g], cs_nq_step(h[g], 1, &st[g])); });
            for (auto& t : th) t.join();
            for (int g = 0; g < G; ++g) moves += st[g].moves_scored;
            if ((s + 1) % exchange == 0 || s + 1 == steps) {
                NK(ncclGroupStart());
                for (int g = 0; g < G; ++g) NK(ncclAllReduce(key[g], key[g], 1, ncclInt64, ncclMin, comm[g], stream[g]));
                NK(ncclGroupEnd());
                long long k = 0;
                CK(cudaSetDevice(0));
                CK(cudaMemcpyAsync(&k, key[0], sizeof k, cudaMemcpyDeviceToHost, stream[0]));
                CK(cudaStreamSynchronize(stream[0]));
                best = k >> 32;
                const uint32_t gid = (uint32_t)(k & 0xffffffffll), owner = gid / chains, local = gid % chains;
                void* src = nullptr;
                CS(h[owner], cs_nq_chain_device_ptr(h[owner], local, &src, &stride));
                NK(ncclGroupStart());  // elite broadcast from the owning GPU (n x 2 bytes)
                for (int g = 0; g < G; ++g)
                    NK(ncclBroadcast(g == (int)owner ? src : elite[g], elite[g], (size_t)stride * 2, ncclUint8, (int)owner, comm[g], stream[g]));
                NK(ncclGroupEnd());
                for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(stream[g])); }
            }
        }
        for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaFree(elite[g])); }
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%s n=%u gpus=%d steps=%u: %llu moves scored in %.3f s (%.3e moves/s, wall clock incl. NCCL/module start-up), best score %lld\n",
                partitioned ? "partitioned" : "sharded", n, G, steps, moves, dt, (double)moves / dt, best);
    for (int g = 0; g < G; ++g) {
        CK(cudaSetDevice(g));
        cs_nq_destroy(h[g]);
        ncclCommDestroy(comm[g]);
        CK(cudaStreamDestroy(stream[g]));
    }
    return 0;
}
