// N-queens "big board" kernels (K3): boards that do not fit one CTA's shared memory
// (CS_NQ_MAX_N_SMEM < n <= 2^20).  State (rows, c, D1, D2 as u32) lives in global memory and
// is L2-resident (n = 10^6: 4 + 4 + 8 + 8 MB << 126 MB L2).  The swap neighbourhood of ONE
// instance is split by column ranges [i_begin, i_end) so several GPUs (or several partitions on
// one GPU) scan disjoint slices; each slice produces a packed 64-bit key
// ((delta/2 + BIAS) << 40 | i << 20 | j), the slices are min-reduced (NCCL all-reduce across
// GPUs) and every replica applies the same winning move -- no state exchange.
//
// Same delta formulae and the same (delta, i, j) tie-break as nq_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nq_kernels.cuh"

namespace csb {

constexpr int NQB_TI = 8;             // column slots per warp tile (8 u32 = one 32 B sector)
constexpr long long NQB_BIAS = 1ll << 22;  // |delta/2| <= 4n+6 < 2^22 for n <= 10^6; key stays positive
constexpr long long NQB_KEY_NONE = 0x7fffffffffffffffll;

struct NqBig {
    int n, n_pad, ld;
    unsigned int* rows;  // [n_pad]
    unsigned int* c;     // [n_pad]
    unsigned int* R;     // [n_pad]
    unsigned int* D1;    // [ld]
    unsigned int* D2;    // [ld]
    long long* score;    // [1] current score (device)
    unsigned long long* ident_pairs;  // [1]
    unsigned int* tile_counter;       // [1]
    unsigned long long* key1;         // [1] (v + BIAS) << 32 | i   (scan result)
    unsigned int* jmin;               // [1] row re-scan result
    unsigned long long* scored;       // [1] non-identity candidates scanned by this partition
    long long* key;                   // [1] packed (v, i, j) of this partition / after reduce
    int i_begin, i_end;               // this partition's column range
    long long* dump;
};

__host__ __device__ inline int nqb_ld(int n_pad) { return 2 * n_pad + 128; }

__global__ void nqb_zero_kernel(NqBig b) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nt = (long long)gridDim.x * blockDim.x;
    for (long long k = tid; k < b.ld; k += nt) {
        b.D1[k] = 0;
        b.D2[k] = 0;
    }
    for (long long k = tid; k < b.n_pad; k += nt) {
        b.R[k] = 0;
        b.c[k] = 0;
    }
    if (tid == 0) {
        *b.score = 0;
        *b.ident_pairs = 0;
    }
}

__global__ void nqb_count_kernel(NqBig b) {
    const int n = b.n;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
         j += (long long)gridDim.x * blockDim.x) {
        const int r = (int)b.rows[j];
        atomicAdd(&b.R[r], 1u);
        atomicAdd(&b.D1[j - r + n - 1], 1u);
        atomicAdd(&b.D2[j + r], 1u);
    }
}

__global__ void nqb_score_kernel(NqBig b) {
    const int n = b.n;
    long long acc = 0, idp = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < 2ll * n - 1;
         k += (long long)gridDim.x * blockDim.x) {
        const long long x = b.D1[k], y = b.D2[k];
        acc += x * (x - 1) + y * (y - 1);
        if (k < n) {
            const long long z = b.R[k];
            idp += z * (z - 1);
        }
    }
    acc += idp;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        idp += __shfl_xor_sync(0xffffffffu, idp, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (acc) atomicAdd((unsigned long long*)b.score, (unsigned long long)acc);
        if (idp) atomicAdd(b.ident_pairs, (unsigned long long)(idp / 2));
    }
}

__global__ void nqb_compute_c_kernel(NqBig b) {
    const int n = b.n;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
         j += (long long)gridDim.x * blockDim.x) {
        const int r = (int)b.rows[j];
        b.c[j] = b.D1[j - r + n - 1] + b.D2[j + r];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *b.tile_counter = 0;
        *b.key1 = ~0ull;
        *b.jmin = 0xffffffffu;
        *b.scored = 0;
    }
}

__device__ __forceinline__ int nqb_swap_half(const NqBig& b, int i, int j) {
    const int n = b.n;
    const int ri = (int)b.rows[i], rj = (int)b.rows[j];
    const int ci = (int)(b.D1[i - ri + n - 1] + b.D2[i + ri]);
    const int cj = (int)(b.D1[j - rj + n - 1] + b.D2[j + rj]);
    const int g = (int)(b.D1[i - rj + n - 1] + b.D2[i + rj] + b.D1[j - ri + n - 1] + b.D2[j + ri]);
    const int d = j - i, t = rj - ri;
    return g - ci - cj + 4 + 2 * ((t == d) | (t == -d));
}

// Scan of the partition's columns.  Each warp steals tiles of NQB_TI consecutive columns i;
// lanes sweep j > i.  The TI gathers D1[i_a - r_j] of one lane are consecutive words (one
// or two 32 B sectors), so L2 traffic per move is ~0.4 sectors; the other two reads are
// lane-consecutive (coalesced).
template <bool PERM, bool DUMP>
__global__ void __launch_bounds__(256) nqb_scan_kernel(NqBig b) {
    constexpr int TI = NQB_TI;
    const int n = b.n, lane = threadIdx.x & 31;
    const unsigned int* __restrict__ rows = b.rows;
    const unsigned int* __restrict__ cc = b.c;
    const unsigned int* __restrict__ D1 = b.D1;
    const unsigned int* __restrict__ D2 = b.D2;
    const int num_tiles = (b.i_end - b.i_begin + TI - 1) / TI;
    int best_v = NQ_INF;
    unsigned int best_i = 0xffffffffu;
    unsigned long long pairs = 0;  // lane 0: candidate pairs of the tiles this warp took
    unsigned int ident = 0;        // identity pairs met (non-permutation boards only)
    for (;;) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(b.tile_counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= num_tiles) break;
        const int i0 = b.i_begin + t * TI;
        int ri[TI], m[TI], ui[TI], wi[TI];
        const unsigned int* p1[TI];
        const unsigned int* p2[TI];
        const unsigned int* q1[TI];
        const unsigned int* q2[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) {
            const int i = i0 + a;
            const int ic = i < n ? i : n - 1;  // clamp loads; such slots are masked below
            ri[a] = (int)__ldg(rows + ic);
            ui[a] = ri[a] - i;
            wi[a] = ri[a] + i;
            p1[a] = D1 + (n - 1 - ri[a]);  // + j
            p2[a] = D2 + ri[a];            // + j
            q1[a] = D1 + (ic + n - 1);     // - r_j
            q2[a] = D2 + ic;               // + r_j
            m[a] = NQ_INF;
            if (lane == 0 && i < b.i_end && i < n - 1) pairs += (unsigned long long)(n - 1 - i);
        }
        for (int jc = (i0 + 1) & ~31; jc < n; jc += 32) {
            const int j = jc + lane;
            const bool jin = j < n;
            const int jl = jin ? j : n - 1;
            const int rj = (int)__ldg(rows + jl);
            const int ncj = -(int)__ldg(cc + jl);
            const int uj = rj - j, wj = rj + j;
#pragma unroll
            for (int a = 0; a < TI; ++a) {
                const int i = i0 + a;
                int x = (int)__ldg(q1[a] - rj) + (int)__ldg(q2[a] + rj) + (int)__ldg(p1[a] + jl) +
                        (int)__ldg(p2[a] + jl) + ncj;
                if (uj == ui[a] || wj == wi[a]) x += 2;
                const bool valid = jin && j > i && i < b.i_end;
                if (!PERM) {
                    if (rj == ri[a]) {
                        x = NQ_INF;
                        ident += valid;
                    }
                }
                if (!valid) x = NQ_INF;
                if (DUMP) {
                    if (jin && j > i && i < b.i_end)
                        b.dump[nq_swap_index(n, i, j)] =
                            (x >= NQ_INF) ? INT64_MAX
                                          : 2ll * (long long)(x + 4 - (int)__ldg(cc + i));
                }
                m[a] = min(m[a], x);
            }
        }
#pragma unroll
        for (int a = 0; a < TI; ++a) {
            const int i = i0 + a;
            if (m[a] < NQ_INF) {
                const int v = m[a] + 4 - (int)__ldg(cc + i);
                if (v < best_v || (v == best_v && (unsigned)i < best_i)) {
                    best_v = v;
                    best_i = (unsigned)i;
                }
            }
        }
    }
    const int wv = __reduce_min_sync(0xffffffffu, best_v);
    const unsigned wi2 = __reduce_min_sync(0xffffffffu, best_v == wv ? best_i : 0xffffffffu);
    if (lane == 0 && wv < NQ_INF)
        atomicMin(b.key1, ((unsigned long long)(wv + NQB_BIAS) << 32) | wi2);
    if (!PERM) ident = __reduce_add_sync(0xffffffffu, ident);
    if (lane == 0 && pairs) atomicAdd(b.scored, pairs - ident);
}

// lowest partner j of the winning column that attains the minimum
__global__ void nqb_rowscan_kernel(NqBig b) {
    const unsigned long long k1 = *b.key1;
    if (k1 == ~0ull) return;
    const int v = (int)((long long)(k1 >> 32) - NQB_BIAS);
    const int i = (int)(k1 & 0xffffffffu);
    const int ri = (int)b.rows[i];
    unsigned int jm = 0xffffffffu;
    for (long long j = i + 1 + blockIdx.x * (long long)blockDim.x + threadIdx.x; j < b.n;
         j += (long long)gridDim.x * blockDim.x) {
        if ((int)b.rows[j] != ri && nqb_swap_half(b, i, (int)j) == v) {
            jm = (unsigned)j;
            break;
        }
    }
    jm = __reduce_min_sync(0xffffffffu, jm);
    if ((threadIdx.x & 31) == 0 && jm != 0xffffffffu) atomicMin(b.jmin, jm);
}

__global__ void nqb_pack_key_kernel(NqBig b) {
    const unsigned long long k1 = *b.key1;
    if (k1 == ~0ull) {
        *b.key = NQB_KEY_NONE;
        return;
    }
    const long long vb = (long long)(k1 >> 32);
    const long long i = (long long)(k1 & 0xffffffffu);
    *b.key = (vb << 40) | (i << 20) | (long long)*b.jmin;
}

struct NqBigStep {
    long long score_after;
    unsigned int a, b;
    unsigned int applied;  // 0: nothing applied (is_best / empty / stalled)
    unsigned int status;
};

// apply the move in *b.key (after any cross-partition reduce) to this replica
__global__ void nqb_apply_kernel(NqBig b, NqChainState* st, NqTraceEntry* trace, int trace_cap,
                                 NqBigStep* out) {
    const long long key = *b.key;
    out->applied = 0;
    out->status = 0;
    out->score_after = *b.score;
    if (key == NQB_KEY_NONE) {
        out->status = 3;
        st->status = 3;
        return;
    }
    const int n = b.n;
    const int v = (int)((key >> 40) - NQB_BIAS);
    const int i = (int)((key >> 20) & 0xfffff), j = (int)(key & 0xfffff);
    const int ri = (int)b.rows[i], rj = (int)b.rows[j];
    b.D1[i - ri + n - 1] -= 1;
    b.D2[i + ri] -= 1;
    b.D1[j - rj + n - 1] -= 1;
    b.D2[j + rj] -= 1;
    b.D1[i - rj + n - 1] += 1;
    b.D2[i + rj] += 1;
    b.D1[j - ri + n - 1] += 1;
    b.D2[j + ri] += 1;
    b.rows[i] = (unsigned)rj;
    b.rows[j] = (unsigned)ri;
    const long long ns = *b.score + 2ll * v;
    *b.score = ns;
    if (trace && st->steps < (unsigned)trace_cap) {
        NqTraceEntry e;
        e.a = (unsigned)i;
        e.b = (unsigned)j;
        e.score_after = ns;
        trace[st->steps] = e;
    }
    st->steps += 1;
    st->score = ns;
    st->status = ns == 0 ? 1u : 0u;
    out->applied = 1;
    out->a = (unsigned)i;
    out->b = (unsigned)j;
    out->score_after = ns;
    out->status = st->status;
}

__global__ void nqb_eval_kernel(NqBig b, const uint2* __restrict__ moves, unsigned long long n_moves,
                                long long* __restrict__ delta) {
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
         k < n_moves; k += (unsigned long long)gridDim.x * blockDim.x) {
        const int i = (int)min(moves[k].x, moves[k].y), j = (int)max(moves[k].x, moves[k].y);
        delta[k] = (i == j || b.rows[i] == b.rows[j]) ? INT64_MAX : 2ll * nqb_swap_half(b, i, j);
    }
}

__global__ void nqb_widen_rows_kernel(const uint16_t* __restrict__ src, unsigned int* dst, int n_pad) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += gridDim.x * blockDim.x)
        dst[k] = src[k];
}

__global__ void nqb_pack_rows32_kernel(const long long* __restrict__ src, unsigned int* dst, int n,
                                       int n_pad, int* bad) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += gridDim.x * blockDim.x) {
        unsigned v = 0;
        if (k < n) {
            const long long x = src[k];
            if (x < 0 || x >= n) *bad = 1;
            v = (unsigned)x;
        }
        dst[k] = v;
    }
}

__global__ void nqb_unpack_rows32_kernel(const unsigned int* __restrict__ src, long long* dst, int n) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        dst[k] = (long long)src[k];
}

__global__ void nqb_init_kernel(unsigned int* rows, int n, int n_pad, unsigned long long seed,
                                unsigned int chain) {
    // sequential Fisher-Yates (one thread): identical draw order to nq_init_kernel / the host mirror
    if (blockIdx.x || threadIdx.x) return;
    for (int i = 0; i < n_pad; ++i) rows[i] = i < n ? (unsigned)i : 0u;
    PhiloxDraws d(seed, chain, 0u);
    for (int k = n - 1; k >= 1; --k) {
        const unsigned idx = d.below((unsigned)k + 1u);
        const unsigned t = rows[k];
        rows[k] = rows[idx];
        rows[idx] = t;
    }
}

}  // namespace csb
