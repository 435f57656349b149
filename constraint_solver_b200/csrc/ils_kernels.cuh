// Iterated-local-search shell around the LS kernels, generic over both plug-ins: a solution
// is a vector of `len` uint16 with an int64 score key (n-queens: score; scheduling:
// hard << 32 | soft); the derived Ord of ScoredSolution is lexicographic (key, vector)
// (local-search/src/local_search.rs:29-37).
//
// One ILS round (local-search/src/iterated_local_search.rs:173-202) is three launches over all
// chains, no host sync in between:
//   ils_perturb_kernel : early-out when the chain's best is_best (:175-184); random restart
//                        every 50th round (:185-191); perturbation (nqueens lib.rs:291-319,
//                        employee-scheduling lib.rs:588-612) -> LS working buffer
//   <problem LS kernel>: LocalSearch::execute (:195-197)
//   ils_accept_kernel  : History::local_search_chose_solution (local_search.rs:205-218, bounded
//                        best-set ordered by (key, vector), BTreeSet dedup) and
//                        AcceptanceCriterion::choose {existing 1, new 5, random best 1} (:51-71)
// All random choices come from ONE Philox stream per chain (purpose 1) with the draw order
// documented in oracle/cs_oracle.c (ils_core), so a chain replays bit-identically on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace csb {

constexpr int ILS_MAX_CAP = 64;
constexpr int ILS_THREADS = 128;

struct IlsChainState {
    unsigned long long t;     // next Philox draw index
    unsigned long long used;  // bitmask of occupied best-set slots
    unsigned int round;
    unsigned int size;        // best-set size
    unsigned int done;        // best is_best: nothing more to do
    unsigned int log_len;
};

struct IlsLogEntry {
    long long new_key;
    unsigned int choice;  // 0 existing, 1 new, 2 random best
    unsigned int pad;
};

struct IlsParams {
    int n_chains, len, stride;
    int value_range;       // perturbed positions get mulhi(u, value_range)
    int restart_is_perm;   // n-queens: Fisher-Yates permutation; scheduling: uniform per slot
    int do_nothing_first;  // strategy table order (scheduling lists DoNothing first)
    int k_before_shuffle;  // scheduling draws the subset size before shuffling
    int best_cap;
    uint16_t* cur;         // [chains][stride] ILS current
    long long* cur_key;    // [chains] (-1 = not scored yet, after a restart)
    uint16_t* work;        // [chains][stride] LS start / working buffer
    const uint16_t* neu;   // [chains][stride] LS result
    const long long* neu_key;
    uint16_t* bset;        // [chains][cap][stride]
    long long* bset_key;   // [chains][cap]
    unsigned char* order;  // [chains][cap] slot ids in ascending (key, vector) order
    IlsChainState* st;
    unsigned int* skip;    // [chains] 1 = LS kernel must leave the chain alone
    IlsLogEntry* log;
    int log_cap;
    unsigned long long seed;
    unsigned int chain_offset;
};

// block-wide three-way lexicographic compare of two vectors (global memory); all threads get it
__device__ __forceinline__ int ils_block_cmp(const uint16_t* a, const uint16_t* b, int len, int* sh) {
    int first = 0x7fffffff;
    for (int i = threadIdx.x; i < len; i += blockDim.x)
        if (a[i] != b[i]) {
            first = i;
            break;
        }
    __syncthreads();
    if (threadIdx.x == 0) sh[0] = 0x7fffffff;
    __syncthreads();
    first = __reduce_min_sync(0xffffffffu, first);
    if ((threadIdx.x & 31) == 0 && first != 0x7fffffff) atomicMin(&sh[0], first);
    __syncthreads();
    const int f = sh[0];
    __syncthreads();
    if (f == 0x7fffffff) return 0;
    return a[f] < b[f] ? -1 : 1;
}

__global__ void __launch_bounds__(ILS_THREADS) ils_perturb_kernel(IlsParams p) {
    extern __shared__ __align__(16) unsigned char ils_smem[];
    uint16_t* w = (uint16_t*)ils_smem;       // [len] working copy
    uint16_t* idx = w + ((p.len + 7) & ~7);  // [len] shuffled positions
    __shared__ int sh[4];
    const int tid = threadIdx.x, nt = blockDim.x, len = p.len;
    for (int chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        __syncthreads();
        IlsChainState* st = p.st + chain;
        uint16_t* cur = p.cur + (size_t)chain * p.stride;
        const unsigned char* order = p.order + (size_t)chain * p.best_cap;
        const long long* bkey = p.bset_key + (size_t)chain * p.best_cap;
        const int size = (int)st->size;
        const bool done = st->done || (size > 0 && bkey[order[0]] == 0);
        if (done) {
            if (tid == 0) {
                st->done = 1;
                p.skip[chain] = 1;
            }
            continue;
        }
        PhiloxDraws rng(p.seed, p.chain_offset + (unsigned)chain, 1u, st->t);
        const unsigned int round = st->round + 1;
        if (round % 50 == 0) {  // reset from random
            if (tid == 0) {
                if (p.restart_is_perm) {
                    for (int i = 0; i < len; ++i) w[i] = (uint16_t)i;
                    for (int q = len - 1; q >= 1; --q) {
                        const unsigned j = rng.below((unsigned)q + 1u);
                        const uint16_t t = w[q];
                        w[q] = w[j];
                        w[j] = t;
                    }
                } else {
                    for (int i = 0; i < len; ++i) w[i] = (uint16_t)rng.below((unsigned)p.value_range);
                }
                p.cur_key[chain] = -1;
            }
            __syncthreads();
            for (int i = tid; i < len; i += nt) cur[i] = w[i];
        } else {
            for (int i = tid; i < len; i += nt) w[i] = cur[i];
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned pick = rng.below(110u);
            sh[1] = p.do_nothing_first ? (pick >= 10u) : (pick < 100u);
        }
        __syncthreads();
        const int change = sh[1];
        if (change) {
            int in_best = 0;  // history.is_best_solution(current)
            for (int e = 0; e < size && !in_best; ++e) {
                const uint16_t* ent = p.bset + ((size_t)chain * p.best_cap + order[e]) * p.stride;
                in_best = (ils_block_cmp(ent, cur, len, sh) == 0);
            }
            if (tid == 0) {
                int kmax = in_best ? len / 20 : len / 2;
                kmax = kmax < 1 ? 1 : (kmax > len ? len : kmax);
                int k = 0;
                if (p.k_before_shuffle) k = 1 + (int)rng.below((unsigned)kmax);
                for (int i = 0; i < len; ++i) idx[i] = (uint16_t)i;
                for (int q = len - 1; q >= 1; --q) {
                    const unsigned j = rng.below((unsigned)q + 1u);
                    const uint16_t t = idx[q];
                    idx[q] = idx[j];
                    idx[j] = t;
                }
                if (!p.k_before_shuffle) k = 1 + (int)rng.below((unsigned)kmax);
                for (int q = 0; q < k; ++q) w[idx[q]] = (uint16_t)rng.below((unsigned)p.value_range);
            }
        }
        __syncthreads();
        uint16_t* work = p.work + (size_t)chain * p.stride;
        for (int i = tid; i < len; i += nt) work[i] = w[i];
        if (tid == 0) {
            st->t = rng.t;
            st->round = round;
            p.skip[chain] = 0;
        }
    }
}

__global__ void __launch_bounds__(ILS_THREADS) ils_accept_kernel(IlsParams p) {
    __shared__ int sh[4];
    const int tid = threadIdx.x, nt = blockDim.x, len = p.len, cap = p.best_cap;
    for (int chain = blockIdx.x; chain < p.n_chains; chain += gridDim.x) {
        __syncthreads();
        IlsChainState* st = p.st + chain;
        if (st->done) continue;
        uint16_t* cur = p.cur + (size_t)chain * p.stride;
        const uint16_t* neu = p.neu + (size_t)chain * p.stride;
        uint16_t* bset = p.bset + (size_t)chain * cap * p.stride;
        long long* bkey = p.bset_key + (size_t)chain * cap;
        unsigned char* order = p.order + (size_t)chain * cap;
        const long long nkey = p.neu_key[chain];
        int size = (int)st->size;
        unsigned long long used = st->used;
        // history.local_search_chose_solution(new)
        bool do_insert = false;
        if (size < cap) {
            do_insert = true;
        } else if (nkey <= bkey[order[size - 1]]) {
            used &= ~(1ull << order[size - 1]);  // remove the worst
            --size;
            do_insert = true;
        }
        if (do_insert) {
            int pos = 0, dup = 0;
            for (; pos < size; ++pos) {
                const int slot = order[pos];
                int c;
                if (bkey[slot] != nkey) c = bkey[slot] < nkey ? -1 : 1;
                else c = ils_block_cmp(bset + (size_t)slot * p.stride, neu, len, sh);
                if (c == 0) dup = 1;
                if (c >= 0) break;
            }
            if (!dup) {
                const int slot = __ffsll((long long)~used) - 1;
                for (int i = tid; i < len; i += nt) bset[(size_t)slot * p.stride + i] = neu[i];
                __syncthreads();
                if (tid == 0) {
                    for (int q = size; q > pos; --q) order[q] = order[q - 1];
                    order[pos] = (unsigned char)slot;
                    bkey[slot] = nkey;
                }
                used |= 1ull << slot;
                ++size;
            }
        }
        __syncthreads();
        // acceptance_criterion.choose(current, new, history)
        if (tid == 0) {
            PhiloxDraws rng(p.seed, p.chain_offset + (unsigned)chain, 1u, st->t);
            const unsigned rb = rng.below((unsigned)size);
            const unsigned wgt = rng.below(7u);
            sh[2] = wgt == 0 ? 0 : (wgt <= 5 ? 1 : 2);
            sh[3] = (int)order[rb];
            st->t = rng.t;
            st->size = (unsigned)size;
            st->used = used;
            if (p.log && st->log_len < (unsigned)p.log_cap) {
                IlsLogEntry e;
                e.new_key = nkey;
                e.choice = (unsigned)sh[2];
                e.pad = 0;
                p.log[(size_t)chain * p.log_cap + st->log_len] = e;
            }
            st->log_len += 1;
        }
        __syncthreads();
        const int choice = sh[2];
        if (choice == 1) {
            for (int i = tid; i < len; i += nt) cur[i] = neu[i];
            if (tid == 0) p.cur_key[chain] = nkey;
        } else if (choice == 2) {
            const int slot = sh[3];
            for (int i = tid; i < len; i += nt) cur[i] = bset[(size_t)slot * p.stride + i];
            if (tid == 0) p.cur_key[chain] = bkey[slot];
        }
    }
}

__global__ void ils_reset_kernel(IlsChainState* st, long long* cur_key, unsigned int* skip, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    IlsChainState z;
    z.t = 0;
    z.used = 0;
    z.round = 0;
    z.size = 0;
    z.done = 0;
    z.log_len = 0;
    st[k] = z;
    cur_key[k] = -1;
    skip[k] = 0;
}

struct IlsSummary {
    long long best_key;       // min over chains of the chain's best-set minimum (INT64_MAX if none)
    unsigned int best_chain;  // local index
    unsigned int chains_done; // chains whose best is_best
    unsigned int min_round, max_round;
};

__global__ void ils_summary_kernel(IlsParams p, IlsSummary* out) {
    __shared__ long long skey[32];
    __shared__ unsigned int sdone[32], smin[32], smax[32], schain[32];
    __shared__ long long sbest;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    long long key = 0x7fffffffffffffffll;
    unsigned int done = 0, rmin = 0xffffffffu, rmax = 0;
    for (int c = threadIdx.x; c < p.n_chains; c += blockDim.x) {
        const IlsChainState st = p.st[c];
        if (st.size) {
            const long long k = p.bset_key[(size_t)c * p.best_cap + p.order[(size_t)c * p.best_cap]];
            key = k < key ? k : key;
            done += (k == 0);
        }
        rmin = min(rmin, st.round);
        rmax = max(rmax, st.round);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        key = ok < key ? ok : key;
        done += __shfl_xor_sync(0xffffffffu, done, o);
        rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
        rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if (l == 0) {
        skey[w] = key;
        sdone[w] = done;
        smin[w] = rmin;
        smax[w] = rmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) {
            key = skey[k] < key ? skey[k] : key;
            done += sdone[k];
            rmin = min(rmin, smin[k]);
            rmax = max(rmax, smax[k]);
        }
        sbest = key;
        out->best_key = key;
        out->chains_done = done;
        out->min_round = rmin;
        out->max_round = rmax;
    }
    __syncthreads();
    const long long best = sbest;
    unsigned int chain = 0xffffffffu;  // lowest chain holding the best key
    for (int c = threadIdx.x; c < p.n_chains; c += blockDim.x) {
        if (!p.st[c].size) continue;
        const long long k = p.bset_key[(size_t)c * p.best_cap + p.order[(size_t)c * p.best_cap]];
        if (k == best) {
            chain = (unsigned)c;
            break;
        }
    }
    chain = __reduce_min_sync(0xffffffffu, chain);
    if (l == 0) schain[w] = chain;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; ++k) chain = min(chain, schain[k]);
        out->best_chain = chain == 0xffffffffu ? 0u : chain;
    }
}

}  // namespace csb
