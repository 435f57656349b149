// Single-process multi-GPU driver over the C ABI + NCCL (no Python, no torch): one handle per
// GPU, one host thread.  Two modes, the two multi-GPU shapes of the north star:
//   default        restart chains sharded across the GPUs (chain k of GPU g has global id
//                  g * chains + k = its Philox stream); every --exchange steps the packed best
//                  keys (score << 32 | global chain id) are min-all-reduced on the device and the
//                  elite board is delivered by cs_nq_exchange_select + a sum-all-reduce; the steps
//                  are enqueued (cs_nq_step_enqueue), so the host never waits between them
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
        // One host thread drives every GPU and never waits inside a step: cs_nq_step_enqueue puts the
        // chain-steps on the handle's stream, the exchange (min-all-reduce of the packed key, the
        // library's owner-masked gather, sum-all-reduce of the elite) follows on the same streams,
        // and cs_nq_step_wait collects the statistics afterwards.
        std::vector<void*> key(G), kbuf(G), elite(G);
        uint32_t stride = 0;
        for (int g = 0; g < G; ++g) {
            CS(h[g], cs_nq_best_key_device_ptr(h[g], &key[g]));
            void* p = nullptr;
            CS(h[g], cs_nq_chain_device_ptr(h[g], 0, &p, &stride));
            CK(cudaSetDevice(g));
            CK(cudaMalloc(&kbuf[g], sizeof(long long)));
            CK(cudaMalloc(&elite[g], (size_t)stride * sizeof(uint16_t)));  // the stride is even: reduced as int32
        }
        for (uint32_t s = 0; s < steps; ++s) {
            for (int g = 0; g < G; ++g) CS(h[g], cs_nq_step_enqueue(h[g], 1));
            const bool xchg = (s + 1) % exchange == 0 || s + 1 == steps;
            if (xchg) {
                for (int g = 0; g < G; ++g) {
                    CK(cudaSetDevice(g));
                    CK(cudaMemcpyAsync(kbuf[g], key[g], sizeof(long long), cudaMemcpyDeviceToDevice, stream[g]));
                }
                NK(ncclGroupStart());
                for (int g = 0; g < G; ++g) NK(ncclAllReduce(kbuf[g], kbuf[g], 1, ncclInt64, ncclMin, comm[g], stream[g]));
                NK(ncclGroupEnd());
                for (int g = 0; g < G; ++g) CS(h[g], cs_nq_exchange_select(h[g], kbuf[g], elite[g], stride));
                NK(ncclGroupStart());  // the owner contributes the chain, everyone else zeros
                for (int g = 0; g < G; ++g) NK(ncclAllReduce(elite[g], elite[g], stride / 2, ncclInt32, ncclSum, comm[g], stream[g]));
                NK(ncclGroupEnd());
            }
            for (int g = 0; g < G; ++g) {
                cs_step_stats st{};
                CS(h[g], cs_nq_step_wait(h[g], &st));
                moves += st.moves_scored;
            }
            if (xchg) {
                long long k = 0;
                CK(cudaSetDevice(0));
                CK(cudaMemcpy(&k, kbuf[0], sizeof k, cudaMemcpyDeviceToHost));
                best = k >> 32;
                // every GPU now holds the same elite: check it against the owner's chain
                const uint32_t gid = (uint32_t)(k & 0xffffffffll), owner = gid / chains, local = gid % chains;
                std::vector<int64_t> row(n);
                CS(h[owner], cs_nq_get_chains(h[owner], local, 1, row.data()));
                std::vector<uint16_t> e(stride);
                for (int g = 0; g < G; ++g) {
                    CK(cudaSetDevice(g));
                    CK(cudaMemcpy(e.data(), elite[g], (size_t)stride * 2, cudaMemcpyDeviceToHost));
                    for (uint32_t i = 0; i < n; ++i)
                        if ((int64_t)e[i] != row[i]) { std::fprintf(stderr, "elite on GPU %d differs from its owner's chain\n", g); return 1; }
                }
            }
        }
        for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaFree(kbuf[g])); }
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
