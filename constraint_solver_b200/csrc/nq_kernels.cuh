// N-queens chain kernels for sm_100a.
//
// One CTA holds one restart chain at a time, entirely in shared memory: rows[], the row and
// both diagonal occupancy counters, and the per-column "own lines" sum c[].  Per step the CTA
// delta-scores the whole neighbourhood (swap: n(n-1)/2 pairs, change: n*n), reduces to the
// lexicographic (delta, a, b) minimum with warp REDUX + a block stage, accepts on device and
// patches the counters.  Chains are pulled from a global work counter (persistent CTAs).
//
// Score definition follows examples/nqueens/src/lib.rs:74-87,126-140 (sum over columns of
// attacking pairs, no blocking) == sum over row/diagonal/anti-diagonal lines of k(k-1).
// Acceptance follows local-search/src/local_search.rs:309-338.
//
// Delta formulae (half-deltas; score delta = 2x):
//   swap (i<j, ri != rj):  G(i,rj) + G(j,ri) - c_i - c_j + 4 + 2*[|ri-rj| == j-i]
//        with G(c,r) = D1[c-r+n-1] + D2[c+r],  c_k = G(k, r_k)   (row lines cancel)
//   change (c: r -> v != r): (R[v] + G(c,v)) - (R[r] + G(c,r)) + 3
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace csb {

constexpr int NQ_INF = 0x3fffffff;
constexpr int NQ_THREADS = 1024;
constexpr int NQ_PAD = 64;  // rows stride multiple and array slack

struct NqChainState {
    long long score;
    long long best_score;
    unsigned long long moves_scored;  // cumulative non-identity candidates delta-scored
    unsigned int steps;               // accepted moves since set/init (== trace length)
    unsigned int status;
    unsigned int is_perm;
    unsigned int pad;
};

struct NqTraceEntry {
    unsigned int a, b;
    long long score_after;
};

struct NqParams {
    int n, n_pad;
    int first_chain, n_chains;  // chain range of this launch
    uint16_t* rows;             // [*, n_pad]
    uint16_t* best_rows;        // [*, n_pad]
    NqChainState* st;
    NqTraceEntry* trace;
    int trace_cap;
    unsigned int* work_counter;
    unsigned long long* totals;    // [2] moves scored / steps accepted by this launch
    unsigned long long max_steps;  // per chain, this launch
    unsigned long long allow_no_improve;  // 0 = never stall-break
    int ls_mode;                          // 1: LocalSearch::execute bookkeeping
    int kind;                             // 0 swap, 1 change
    long long* dump;                      // debug: every candidate delta (one chain)
    const unsigned int* skip;             // optional [chains]: 1 = leave the chain alone (ILS)
    int force_scalar;                     // v2 kernel: never take the packed path
    // reference mode (change kind): the reference's own proposer + window + tie-break
    int ref_mode;
    unsigned long long window;            // .take(window_size), local_search.rs:321
    unsigned long long* ls_rng_t;         // [chains] draw counter of the LocalSearch-owned rng
    unsigned long long seed;
    unsigned int chain_offset;
};

// ------------------------------------------------------------------ shared memory view
struct NqSmem {
    uint16_t* rows;  // [n_pad + PAD]
    uint16_t* c;     // [n_pad + PAD]   own-lines sum per column
    uint16_t* R;     // [n_pad + PAD]   row occupancy
    uint16_t* D1;    // [ld]            index c - r + n - 1
    uint16_t* D2;    // [ld]            index c + r
    int* red;        // [128] reduction scratch / broadcast
    int ld;
};

__host__ __device__ inline int nq_ld(int n_pad) { return 2 * n_pad + 2 * NQ_PAD; }
__host__ __device__ inline size_t nq_smem_bytes(int n_pad) {
    return (size_t)3 * (n_pad + NQ_PAD) * 2 + (size_t)2 * nq_ld(n_pad) * 2 + 128 * 4;
}

__device__ __forceinline__ NqSmem nq_carve(unsigned char* base, int n_pad) {
    NqSmem s;
    s.ld = nq_ld(n_pad);
    s.rows = (uint16_t*)base;
    s.c = s.rows + (n_pad + NQ_PAD);
    s.R = s.c + (n_pad + NQ_PAD);
    s.D1 = s.R + (n_pad + NQ_PAD);
    s.D2 = s.D1 + s.ld;
    s.red = (int*)(s.D2 + s.ld);
    return s;
}

__device__ __forceinline__ void smem_inc16(uint16_t* arr, int idx) {
    atomicAdd((unsigned int*)arr + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
}

// Block-wide sum of a long long; result broadcast to all threads.
__device__ __forceinline__ long long block_sum_ll(long long v, int* red) {
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    long long* r = (long long*)red;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) r[w] = v;
    __syncthreads();
    if (w == 0) {
        long long x = (l < nw) ? r[l] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (l == 0) r[32] = x;
    }
    __syncthreads();
    const long long out = r[32];
    __syncthreads();
    return out;
}

// Lexicographic (v, id) block argmin; result broadcast.  ids are < 2^31.
__device__ __forceinline__ void block_argmin(int& v, unsigned int& id, int* red) {
    const unsigned full = 0xffffffffu;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    int wv = __reduce_min_sync(full, v);
    unsigned wi = __reduce_min_sync(full, v == wv ? id : 0xffffffffu);
    __syncthreads();
    if (l == 0) {
        red[w] = wv;
        red[32 + w] = (int)wi;
    }
    __syncthreads();
    if (w == 0) {
        int x = (l < nw) ? red[l] : NQ_INF;
        unsigned xi = (l < nw) ? (unsigned)red[32 + l] : 0xffffffffu;
        int bv = __reduce_min_sync(full, x);
        unsigned bi = __reduce_min_sync(full, x == bv ? xi : 0xffffffffu);
        if (l == 0) {
            red[64] = bv;
            red[65] = (int)bi;
        }
    }
    __syncthreads();
    v = red[64];
    id = (unsigned)red[65];
    __syncthreads();
}

// ------------------------------------------------------------------ chain load / counters
// Loads rows (global u16) into smem, builds R/D1/D2, returns the full score
// sum_lines k(k-1) and the number of identity swap pairs sum_r R(R-1)/2.
__device__ void nq_load_chain(const NqSmem& s, const uint16_t* __restrict__ grow, int n, int n_pad,
                              long long& score, long long& ident_pairs) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // zero everything (rows/c/R pads included) with 16-byte stores
    {
        const int total16 = (int)((nq_smem_bytes(n_pad) - 128 * 4) / 16);
        uint4* z = (uint4*)s.rows;
        for (int k = tid; k < total16; k += nt) z[k] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    {
        const uint4* g = (const uint4*)grow;  // n_pad multiple of 64 -> 16B aligned rows
        uint4* d = (uint4*)s.rows;
        for (int k = tid; k < n_pad / 8; k += nt) d[k] = g[k];
    }
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        const int r = s.rows[j];
        smem_inc16(s.R, r);
        smem_inc16(s.D1, j - r + n - 1);
        smem_inc16(s.D2, j + r);
    }
    __syncthreads();
    long long acc = 0, idp = 0;
    for (int k = tid; k < 2 * n - 1; k += nt) {
        const long long a = s.D1[k], b = s.D2[k];
        acc += a * (a - 1) + b * (b - 1);
    }
    for (int k = tid; k < n; k += nt) {
        const long long a = s.R[k];
        idp += a * (a - 1);
    }
    score = block_sum_ll(acc + idp, s.red);
    ident_pairs = block_sum_ll(idp, s.red) / 2;
}

__device__ __forceinline__ void nq_compute_c(const NqSmem& s, int n) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int r = s.rows[j];
        s.c[j] = (uint16_t)(s.D1[j - r + n - 1] + s.D2[j + r]);
    }
}

// Exact half-delta of one swap (requires ri != rj), scalar reference form.
__device__ __forceinline__ int nq_swap_half(const NqSmem& s, int n, int i, int j) {
    const int ri = s.rows[i], rj = s.rows[j];
    const int ci = s.D1[i - ri + n - 1] + s.D2[i + ri];
    const int cj = s.D1[j - rj + n - 1] + s.D2[j + rj];
    const int g = s.D1[i - rj + n - 1] + s.D2[i + rj] + s.D1[j - ri + n - 1] + s.D2[j + ri];
    const int d = j - i, t = rj - ri;
    const int att = (t == d) | (t == -d);
    return g - ci - cj + 4 + 2 * att;
}

// Exact half-delta of one change (requires v != r).
__device__ __forceinline__ int nq_change_half(const NqSmem& s, int n, int c, int v) {
    const int r = s.rows[c];
    const int neu = s.R[v] + s.D1[c - v + n - 1] + s.D2[c + v];
    const int old = s.R[r] + s.D1[c - r + n - 1] + s.D2[c + r];
    return neu - old + 3;
}

__device__ __forceinline__ long long nq_swap_index(long long n, long long i, long long j) {
    return i * n - i * (i + 1) / 2 + (j - i - 1);
}

// ------------------------------------------------------------------ neighbourhood scans
// Swap scan.  Warp owns a tile of TI consecutive columns i (warp-uniform), lanes sweep j.
// Per (i,j): two data-dependent gathers (D1[i-rj], D2[i+rj]) and two lane-consecutive,
// conflict-free reads (D1[j-ri], D2[j+ri]); r_j and c_j are one coalesced read per j shared
// by the TI slots.  Only the minimum VALUE per column i is tracked here; the winning j is
// recovered afterwards by re-scanning the single winning row.
template <int TI, bool PERM, bool DUMP>
__device__ __forceinline__ void nq_scan_swap(const NqSmem& s, int n, int& best_v,
                                             unsigned int& best_i, long long* dump) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int num_tiles = (n - 1 + TI - 1) / TI;  // columns 0..n-2 have partners
    const unsigned char* D1b = (const unsigned char*)s.D1;
    const unsigned char* D2b = (const unsigned char*)s.D2;
    best_v = NQ_INF;
    best_i = 0xffffffffu;

    for (int k = 0; k * W < num_tiles; ++k) {
        const int t = k * W + ((k & 1) ? (W - 1 - w) : w);  // boustrophedon: balances row lengths
        if (t >= num_tiles) continue;
        const int i0 = t * TI;
        int ri[TI], m[TI], a1[TI], a2[TI], ui[TI], wi[TI];
        const unsigned char* p1[TI];
        const unsigned char* p2[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) {
            const int i = i0 + a;
            ri[a] = s.rows[i];  // padded: in bounds, rows beyond n-1 read as 0 and get masked
            a1[a] = 2 * (i + n - 1);
            a2[a] = 2 * i;
            ui[a] = ri[a] - i;
            wi[a] = ri[a] + i;
            p1[a] = D1b + 2 * (n - 1 - ri[a]);
            p2[a] = D2b + 2 * ri[a];
            m[a] = NQ_INF;
        }

        auto body = [&](int jc, bool masked) {
            const int j = jc + lane;
            const int rj = s.rows[j];
            const int ncj = -(int)s.c[j];
            const int rj2 = 2 * rj, j2 = 2 * j;
            const int uj = rj - j, wj = rj + j;
#pragma unroll
            for (int a = 0; a < TI; ++a) {
                int x = (int)*(const uint16_t*)(D1b + (a1[a] - rj2)) +
                        (int)*(const uint16_t*)(D2b + (a2[a] + rj2)) +
                        (int)*(const uint16_t*)(p1[a] + j2) + (int)*(const uint16_t*)(p2[a] + j2) +
                        ncj;
                if (uj == ui[a] || wj == wi[a]) x += 2;
                if (!PERM) {
                    if (rj == ri[a]) x = NQ_INF;
                }
                if (masked) {
                    if (!(j > i0 + a && j < n)) x = NQ_INF;
                }
                if (DUMP) {
                    const int i = i0 + a;
                    if (j > i && j < n) {
                        const int ci = (int)s.c[i];
                        dump[nq_swap_index(n, i, j)] =
                            (x >= NQ_INF) ? INT64_MAX : 2ll * (long long)(x + 4 - ci);
                    }
                }
                m[a] = min(m[a], x);
            }
        };

        int jc = (i0 + 1) & ~31;
        const int jm0 = (i0 + TI + 31) & ~31;
        const int jm1 = n & ~31;
        const int head_end = jm0 < n ? jm0 : n;
        for (; jc < head_end; jc += 32) body(jc, true);
#pragma unroll 2
        for (; jc < jm1; jc += 32) body(jc, false);
        if (jc < n) body(jc, true);

#pragma unroll
        for (int a = 0; a < TI; ++a) {
            const int i = i0 + a;
            if (m[a] < NQ_INF) {
                const int v = m[a] + 4 - (int)s.c[i];
                if (v < best_v) {
                    best_v = v;
                    best_i = (unsigned)i;
                }
            }
        }
    }
}

// Change scan: column c is warp-uniform, lanes sweep the new row v; all three reads
// (R[v], D1[c-v], D2[c+v]) are lane-consecutive.
template <int TI, bool DUMP>
__device__ __forceinline__ void nq_scan_change(const NqSmem& s, int n, int& best_v,
                                               unsigned int& best_c, long long* dump) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int num_tiles = (n + TI - 1) / TI;
    best_v = NQ_INF;
    best_c = 0xffffffffu;
    for (int t = w; t < num_tiles; t += W) {
        const int c0 = t * TI;
        int r[TI], m[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) {
            r[a] = s.rows[c0 + a];
            m[a] = NQ_INF;
        }
        for (int vc = 0; vc < n; vc += 32) {
            const int v = vc + lane;
            const int Rv = s.R[v];
#pragma unroll
            for (int a = 0; a < TI; ++a) {
                const int c = c0 + a;
                int x = Rv + (int)s.D1[c - v + n - 1] + (int)s.D2[c + v];
                if (v == r[a] || v >= n || c >= n) x = NQ_INF;
                if (DUMP) {
                    if (c < n && v < n) {
                        const int old = s.R[r[a]] + s.D1[c - r[a] + n - 1] + s.D2[c + r[a]];
                        dump[(long long)c * n + v] =
                            (x >= NQ_INF) ? INT64_MAX : 2ll * (long long)(x - old + 3);
                    }
                }
                m[a] = min(m[a], x);
            }
        }
#pragma unroll
        for (int a = 0; a < TI; ++a) {
            const int c = c0 + a;
            if (c < n && m[a] < NQ_INF) {
                const int old = s.R[r[a]] + s.D1[c - r[a] + n - 1] + s.D2[c + r[a]];
                const int v = m[a] - old + 3;
                if (v < best_v) {
                    best_v = v;
                    best_c = (unsigned)c;
                }
            }
        }
    }
}

// NQueensMoveProposer::iter_local_moves, examples/nqueens/src/lib.rs:177-205, run by one thread
// on the chain's counters: per-column conflicts = R[r] + D1 + D2 - 3 (== get_col_scores),
// conflicted columns in ascending order, `amount` weighted draws without replacement, then a
// random prefix of a partial shuffle.  Same draw order as oracle/cs_oracle.c nq_ref_propose.
// cols / sc: scratch of n entries each (u16); returns the number of chosen columns (in cols[]).
__device__ int nq_ref_propose(const NqSmem& s, int n, PhiloxDraws& rng, uint16_t* cols,
                              uint16_t* sc) {
    int len = 0;
    for (int c = 0; c < n; ++c) {
        const int r = s.rows[c];
        const int v = (int)s.R[r] + (int)s.D1[c - r + n - 1] + (int)s.D2[c + r] - 3;
        if (v != 0) {
            cols[len] = (uint16_t)c;
            sc[len] = (uint16_t)v;
            ++len;
        }
    }
    if (len == 0) return 0;
    int amount = n / 20;
    amount = amount < 1 ? 1 : (amount > len ? len : amount);
    // picked columns are compacted to the front of a second region: reuse the tail of sc[]
    uint16_t* picked = sc + n;  // caller provides 2n entries behind sc
    int npicked = 0;
    for (int k = 0; k < amount; ++k) {
        unsigned total = 0;
        for (int q = 0; q < len; ++q) total += sc[q];
        const unsigned x = rng.below(total);
        unsigned acc = 0;
        int idx = 0;
        for (; idx < len; ++idx) {
            acc += sc[idx];
            if (acc > x) break;
        }
        picked[npicked++] = cols[idx];
        for (int q = idx; q + 1 < len; ++q) {
            cols[q] = cols[q + 1];
            sc[q] = sc[q + 1];
        }
        --len;
    }
    const int num_cols = 1 + (int)rng.below((unsigned)npicked);
    for (int k = 0; k < num_cols; ++k) {
        const int j = k + (int)rng.below((unsigned)(npicked - k));
        const uint16_t t = picked[k];
        picked[k] = picked[j];
        picked[j] = t;
    }
    for (int k = 0; k < num_cols; ++k) cols[k] = picked[k];
    return num_cols;
}

// ------------------------------------------------------------------ the step kernel
template <int TI>
__global__ void __launch_bounds__(NQ_THREADS, 1) nq_step_kernel(NqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NqSmem s = nq_carve(smem_raw, p.n_pad);
    const int n = p.n, tid = threadIdx.x;
    // reference mode keeps its proposer scratch (2 x n u16) behind the regular layout
    uint16_t* ref_scratch = (uint16_t*)(smem_raw + nq_smem_bytes(p.n_pad));

    for (;;) {
        __syncthreads();
        if (tid == 0) s.red[96] = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const int local = s.red[96];
        if (local >= p.n_chains) break;
        const int chain = p.first_chain + local;
        if (p.skip && p.skip[chain]) continue;
        uint16_t* grow = p.rows + (size_t)chain * p.n_pad;
        NqChainState* st = p.st + chain;

        long long score, ident_pairs;
        nq_load_chain(s, grow, n, p.n_pad, score, ident_pairs);
        const bool perm = (ident_pairs == 0);
        const long long nbh = (p.kind == 0) ? ((long long)n * (n - 1) / 2 - ident_pairs)
                                            : ((long long)n * n - n);
        long long best_score = p.ls_mode ? score : st->best_score;
        unsigned long long no_improve = 0;
        const unsigned int steps0 = st->steps;
        unsigned int steps = steps0;
        unsigned long long scored = 0;
        unsigned int status = 0;  // RUNNING
        bool best_dirty = false;
        if (p.ls_mode) {  // best_solution = current_solution.clone(), local_search.rs:307
            for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
                ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] = ((const uint4*)s.rows)[k];
        }

        for (unsigned long long it = 0; it < p.max_steps; ++it) {
            if (score == 0 && !p.dump) {  // is_best, local_search.rs:311-314 (returns current)
                status = 1;
                if (p.ls_mode && best_score != 0) best_dirty = true;
                best_score = 0;
                break;
            }
            int v;
            unsigned int a;
            unsigned long long ref_cand = 0;
            if (p.kind == 0) {
                nq_compute_c(s, n);
                __syncthreads();
                if (perm) {
                    if (p.dump) nq_scan_swap<TI, true, true>(s, n, v, a, p.dump);
                    else nq_scan_swap<TI, true, false>(s, n, v, a, nullptr);
                } else {
                    if (p.dump) nq_scan_swap<TI, false, true>(s, n, v, a, p.dump);
                    else nq_scan_swap<TI, false, false>(s, n, v, a, nullptr);
                }
            } else if (!p.ref_mode) {
                if (p.dump) nq_scan_change<TI, true>(s, n, v, a, p.dump);
                else nq_scan_change<TI, false>(s, n, v, a, nullptr);
            } else {
                // reference mode: sampled conflicted columns, window, (score, solution) order
                __syncthreads();
                if (tid == 0) {
                    PhiloxDraws rng(p.seed, p.chain_offset + (unsigned)chain, 2u, p.ls_rng_t[chain]);
                    // scratch: s.c holds the chosen columns; D-array slack is not touched:
                    // column scores and the picked list live in the reduction-free tail of s.c
                    s.red[97] = nq_ref_propose(s, n, rng, s.c, ref_scratch);
                    p.ls_rng_t[chain] = rng.t;
                }
                __syncthreads();
                const int ncols = s.red[97];
                long long cand = (long long)ncols * (n - 1);
                if ((unsigned long long)cand > p.window) cand = (long long)p.window;
                v = NQ_INF;
                a = 0xffffffffu;
                for (long long q = tid; q < cand; q += blockDim.x) {
                    const int col = s.c[q / (n - 1)];
                    const int vv = (int)(q % (n - 1));
                    const int r = s.rows[col];
                    const int val = vv + (vv >= r ? 1 : 0);  // the q-th non-identity candidate
                    const int h = nq_change_half(s, n, col, val);
                    // derived Ord on the resulting vectors: lowering an entry beats every raise;
                    // among lowerings the lowest column wins, among raises the highest column
                    const unsigned key = val < r ? (unsigned)(col * n + val)
                                                 : (unsigned)(n * n + (n - 1 - col) * n + val);
                    if (h < v || (h == v && key < a)) {
                        v = h;
                        a = key;
                    }
                }
                ref_cand = (unsigned long long)cand;
            }
            block_argmin(v, a, s.red);
            scored += p.ref_mode ? ref_cand : (unsigned long long)nbh;
            if (p.dump) break;  // debug dump: evaluate once, accept nothing
            if (v >= NQ_INF) {  // empty neighbourhood, local_search.rs:336-338
                status = 3;
                break;
            }
            // recover b: lowest partner of row/column `a` that attains v
            unsigned int b = 0xffffffffu;
            if (p.ref_mode) {  // the key encodes (column, value)
                const unsigned nn = (unsigned)n * (unsigned)n;
                if (a < nn) {
                    b = a % (unsigned)n;
                    a = a / (unsigned)n;
                } else {
                    const unsigned k2 = a - nn;
                    b = k2 % (unsigned)n;
                    a = (unsigned)(n - 1) - k2 / (unsigned)n;
                }
            } else if (p.kind == 0) {
                const int ra = s.rows[a];
                for (int j = (int)a + 1 + tid; j < n; j += blockDim.x)
                    if (s.rows[j] != ra && nq_swap_half(s, n, (int)a, j) == v) {
                        b = (unsigned)j;
                        break;
                    }
            } else {
                const int ra = s.rows[a];
                for (int x = tid; x < n; x += blockDim.x)
                    if (x != ra && nq_change_half(s, n, (int)a, x) == v) {
                        b = (unsigned)x;
                        break;
                    }
            }
            if (!p.ref_mode) {
                int dummy = 0;
                block_argmin(dummy, b, s.red);
            }

            const long long new_score = score + 2ll * v;
            bool improved = new_score < score;
            if (!improved) {  // local_search.rs:329-334
                ++no_improve;
                if (p.allow_no_improve && no_improve >= p.allow_no_improve) {
                    status = 2;
                    break;
                }
            } else {
                no_improve = 0;
            }
            if (tid == 0) {  // apply, local_search.rs:335 (current = best of neighbourhood)
                if (p.kind == 0) {
                    const int i = (int)a, j = (int)b;
                    const int ri = s.rows[i], rj = s.rows[j];
                    s.D1[i - ri + n - 1] -= 1;
                    s.D2[i + ri] -= 1;
                    s.D1[j - rj + n - 1] -= 1;
                    s.D2[j + rj] -= 1;
                    s.D1[i - rj + n - 1] += 1;
                    s.D2[i + rj] += 1;
                    s.D1[j - ri + n - 1] += 1;
                    s.D2[j + ri] += 1;
                    s.rows[i] = (uint16_t)rj;
                    s.rows[j] = (uint16_t)ri;
                } else {
                    const int c = (int)a, nv = (int)b, r = s.rows[c];
                    s.R[r] -= 1;
                    s.D1[c - r + n - 1] -= 1;
                    s.D2[c + r] -= 1;
                    s.R[nv] += 1;
                    s.D1[c - nv + n - 1] += 1;
                    s.D2[c + nv] += 1;
                    s.rows[c] = (uint16_t)nv;
                }
                if (p.trace && steps < (unsigned)p.trace_cap) {
                    NqTraceEntry e;
                    e.a = a;
                    e.b = b;
                    e.score_after = new_score;
                    p.trace[(size_t)chain * p.trace_cap + steps] = e;
                }
            }
            score = new_score;
            ++steps;
            __syncthreads();
            if (improved) {  // best_solution = neighborhood_best.clone(), local_search.rs:326-328
                best_score = new_score;
                for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
                    ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] =
                        ((const uint4*)s.rows)[k];
            }
        }
        __syncthreads();
        if (p.dump) continue;
        if (best_dirty) {
            for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
                ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] = ((const uint4*)s.rows)[k];
        }
        for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
            ((uint4*)grow)[k] = ((const uint4*)s.rows)[k];
        if (tid == 0) {
            st->score = score;
            st->best_score = best_score;
            st->moves_scored += scored;
            st->steps = steps;
            st->status = status;
            st->is_perm = perm ? 1u : 0u;
            atomicAdd(p.totals, scored);
            atomicAdd(p.totals + 1, (unsigned long long)(steps - steps0));
        }
    }
}

// ------------------------------------------------------------------ small kernels
// Fisher-Yates permutation per chain (examples/nqueens/src/lib.rs:156-160), one thread per
// chain; draw t of stream (seed, chain, INIT) picks idx = mulhi(u, k+1) for k = n-1-t.
__global__ void nq_init_kernel(uint16_t* rows, NqChainState* st, int n, int n_pad, int n_chains,
                               unsigned long long seed, unsigned int chain_offset) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    uint16_t* r = rows + (size_t)chain * n_pad;
    for (int i = 0; i < n_pad; ++i) r[i] = (uint16_t)(i < n ? i : 0);
    PhiloxDraws d(seed, chain_offset + (unsigned)chain, 0u);
    for (int k = n - 1; k >= 1; --k) {
        const unsigned idx = d.below((unsigned)k + 1u);
        const uint16_t t = r[k];
        r[k] = r[idx];
        r[idx] = t;
    }
    NqChainState z;
    z.score = -1;
    z.best_score = -1;
    z.moves_scored = 0;
    z.steps = 0;
    z.status = 0;
    z.is_perm = 1;
    z.pad = 0;
    st[chain] = z;
}

// score of every chain from counters-free closed form is done by the step kernel's loader;
// this kernel only (re)computes scores after set/init: one CTA per chain, same loader.
__global__ void __launch_bounds__(NQ_THREADS, 1)
    nq_rescore_kernel(NqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NqSmem s = nq_carve(smem_raw, p.n_pad);
    for (int local = blockIdx.x; local < p.n_chains; local += gridDim.x) {
        const int chain = p.first_chain + local;
        long long score, idp;
        __syncthreads();
        nq_load_chain(s, p.rows + (size_t)chain * p.n_pad, p.n, p.n_pad, score, idp);
        if (threadIdx.x == 0) {
            p.st[chain].score = score;
            p.st[chain].best_score = score;
            p.st[chain].is_perm = (idp == 0);
            p.st[chain].status = (score == 0) ? 1u : 0u;
        }
        for (int k = threadIdx.x; k < p.n_pad / 8; k += blockDim.x)
            ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] = ((const uint4*)s.rows)[k];
    }
}

// Full re-score by the reference's pair test (examples/nqueens/src/lib.rs:74-87), grid over
// col1; accumulates the number of conflicting pairs into *pairs.
template <typename T>
__global__ void nq_pair_score_kernel(const T* __restrict__ rows, int n,
                                     unsigned long long* pairs) {
    unsigned long long local = 0;
    for (int col1 = blockIdx.x; col1 < n; col1 += gridDim.x) {
        const int row1 = rows[col1];
        for (int col2 = col1 + 1 + threadIdx.x; col2 < n; col2 += blockDim.x) {
            const int row_diff = (int)rows[col2] - row1;
            const int column_diff = col2 - col1;
            local += (row_diff == 0) | (abs(row_diff) == column_diff);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(pairs, local);
}

// Explicit-move deltas against one chain (parity hook).
__global__ void __launch_bounds__(NQ_THREADS, 1)
    nq_eval_kernel(NqParams p, int chain, int kind, const uint2* __restrict__ moves,
                   unsigned long long n_moves, long long* __restrict__ delta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NqSmem s = nq_carve(smem_raw, p.n_pad);
    long long score, idp;
    nq_load_chain(s, p.rows + (size_t)chain * p.n_pad, p.n, p.n_pad, score, idp);
    __syncthreads();
    const int n = p.n;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
         k < n_moves; k += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 mv = moves[k];
        long long d;
        if (kind == 0) {
            const int i = (int)min(mv.x, mv.y), j = (int)max(mv.x, mv.y);
            d = (i == j || s.rows[i] == s.rows[j]) ? INT64_MAX : 2ll * nq_swap_half(s, n, i, j);
        } else {
            d = ((int)mv.y == (int)s.rows[mv.x]) ? INT64_MAX
                                                 : 2ll * nq_change_half(s, n, (int)mv.x, (int)mv.y);
        }
        delta[k] = d;
    }
}

struct NqStats {
    unsigned long long moves_scored;
    unsigned long long steps;
    long long best_score;
    long long best_key;  // (score << 32) | global chain id
    unsigned int best_chain;
    unsigned int chains_at_best;
};

// One CTA: reduce per-chain state into NqStats (deltas vs. the snapshot taken before).
__global__ void nq_stats_kernel(const NqChainState* __restrict__ st, int n_chains,
                                unsigned int chain_offset, NqStats* out) {
    __shared__ long long skey[32];
    __shared__ unsigned long long smv[32], sst[32];
    __shared__ unsigned int sab[32];
    long long key = INT64_MAX;
    unsigned long long mv = 0, steps = 0;
    unsigned int ab = 0;
    for (int c = threadIdx.x; c < n_chains; c += blockDim.x) {
        const NqChainState x = st[c];
        const long long k = (x.score << 32) | (long long)(unsigned)c;
        key = k < key ? k : key;
        mv += x.moves_scored;
        steps += x.steps;
        ab += (x.score == 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        key = ok < key ? ok : key;
        mv += __shfl_xor_sync(0xffffffffu, mv, o);
        steps += __shfl_xor_sync(0xffffffffu, steps, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        skey[w] = key;
        smv[w] = mv;
        sst[w] = steps;
        sab[w] = ab;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; ++k) {
            key = skey[k] < key ? skey[k] : key;
            mv += smv[k];
            steps += sst[k];
            ab += sab[k];
        }
        out->moves_scored = mv;
        out->steps = steps;
        out->best_score = key >> 32;
        out->best_chain = (unsigned)(key & 0xffffffffll);
        out->best_key = ((key >> 32) << 32) | (long long)((unsigned)(key & 0xffffffffll) + chain_offset);
        out->chains_at_best = ab;
    }
}

// int64 (reference element type) <-> uint16 device layout
__global__ void nq_pack_rows_kernel(const long long* __restrict__ src, uint16_t* __restrict__ dst,
                                    int n, int n_pad, int count, int* bad) {
    const long long total = (long long)count * n_pad;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(k / n_pad), j = (int)(k % n_pad);
        uint16_t v = 0;
        if (j < n) {
            const long long x = src[(long long)c * n + j];
            // an out-of-range row is flagged AND stored as 0: whatever runs on these rows before the
            // host sees the flag (the counter build indexes shared memory by row) stays in bounds
            if (x < 0 || x >= n) *bad = 1;
            else v = (uint16_t)x;
        }
        dst[k] = v;
    }
}

// u16 rows already on the device (elite broadcast target): same range check and clamp
__global__ void nq_copy_rows_checked_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
                                            int n, int n_pad, int* bad) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_pad; j += gridDim.x * blockDim.x) {
        uint16_t v = 0;
        if (j < n) {
            const uint16_t x = src[j];
            if ((int)x >= n) *bad = 1;
            else v = x;
        }
        dst[j] = v;
    }
}

__global__ void nq_unpack_rows_kernel(const uint16_t* __restrict__ src, long long* __restrict__ dst,
                                      int n, int n_pad, int count) {
    const long long total = (long long)count * n;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total;
         k += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(k / n), j = (int)(k % n);
        dst[k] = (long long)src[(long long)c * n_pad + j];
    }
}

__global__ void nq_reset_state_kernel(NqChainState* st, int first, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    NqChainState z;
    z.score = -1;
    z.best_score = -1;
    z.moves_scored = 0;
    z.steps = 0;
    z.status = 0;
    z.is_perm = 0;
    z.pad = 0;
    st[first + k] = z;
}

}  // namespace csb
