// N-queens "big board" kernels (K3): boards that do not fit one CTA's shared memory
// (CS_NQ_MAX_N_SMEM < n <= 2^20).  State (rows, c, D1, D2 as u32) lives in global memory and
// is L2-resident (n = 10^6: 4 + 4 + 8 + 8 MB << 126 MB L2).  The swap neighbourhood of ONE
// instance is split by column ranges [i_begin, i_end) so several GPUs (or several partitions on
// one GPU) scan disjoint slices; each slice produces a packed 64-bit key
// ((delta/2 + BIAS) << 40 | i << 20 | j), the slices are min-reduced (NCCL all-reduce across
// GPUs) and every replica applies the same winning move -- no state exchange.
//
// Same delta formulae and the same (delta, i, j) tie-break as nq_kernels.cuh.
//
// Two scans: nqb_scan_kernel (u32 counters, any board) and nqb_scan_packed_kernel (byte counters
// in 8 byte-shifted copies rebuilt per step, 16x2 SIMD, segment-major work units; permutation
// boards with n >= 256 and no line above 62 queens -- decided on the device, the scalar scan
// answers otherwise).
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
    // packed-window fast scan (nqb_scan_packed_kernel): byte counters, 8 byte-shifted copies of
    // each diagonal array (copy c, byte y = D[y + c]; Q2's copies follow Q1's), rebuilt per step
    unsigned char* Q;                 // [16][ldb]
    unsigned char* cb;                // [n_pad + 128] own-lines sum per column as a byte
    unsigned int* maxcount;           // [1] largest diagonal line count (set by nqb_pack_kernel)
    int ldb;
    int use_packed;                   // host: permutation board, n >= NQBP_MIN_N, not forced scalar
    int seg;                          // chunks (of 128 columns j) per work unit of the packed scan
};

constexpr int NQBP_MIN_N = 256;
constexpr int NQBP_COPIES = 16;     // byte shifts 0..15: any 16-counter window is one aligned 128-bit word
constexpr int NQBP_MAX_COUNT = 62;   // four byte counters + slack stay below 256 (as nq_packed.cuh)
__host__ __device__ inline int nqb_ldb(int n_pad) { return (2 * n_pad + 256 + 15) & ~15; }
__host__ __device__ inline int nqb_ld(int n_pad) { return nqb_ldb(n_pad) + 64; }

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
        const unsigned int cj = b.D1[j - r + n - 1] + b.D2[j + r];
        b.c[j] = cj;
        b.cb[j] = (unsigned char)(cj < 255u ? cj : 255u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *b.maxcount = 0;
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
    if (PERM && b.use_packed && *b.maxcount <= (unsigned)NQBP_MAX_COUNT) return;  // the packed scan runs instead
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


// ---------------------------------------------------------------------------------------------
// Packed-window scan for big boards: the shared-memory fast scan of nq_packed.cuh with the byte
// counters in global memory (L2-resident, gathers served by L1).  A CTA takes 128 consecutive
// columns (NQBP_WARPS warps x NQBP_TI column slots) so its warps sweep the same j chunks together and their
// adjacent 16-byte gather windows share L1 sectors; a lane owns 4 consecutive columns j.
// Diagonal ids need 21 bits at n = 10^6: the 16x2 attack test compares the low 15 bits and a rare
// exact pass repairs the aliases (low halves equal, ids different).
// Identical integer value per move as nqb_scan_kernel (parity: cs_nq_neighbourhood_deltas runs
// THIS scan with the dump flag whenever the board qualifies).

// u32 counters -> the NQBP_COPIES byte-shifted copies of both arrays, and the largest line count
__global__ void nqb_pack_kernel(NqBig b) {
    const int ldb = b.ldb;
    unsigned int mx = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ldb / 16;
         k += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(16 * k);
#pragma unroll
        for (int arr = 0; arr < 2; ++arr) {
            const unsigned int* __restrict__ D = arr ? b.D2 : b.D1;
            unsigned int d[16 + NQBP_COPIES - 1];
#pragma unroll
            for (int t = 0; t < 16 + NQBP_COPIES - 1; ++t) {
                d[t] = (y + t < b.ld) ? D[y + t] : 0u;
                mx = max(mx, d[t]);
                d[t] &= 0xffu;
            }
#pragma unroll
            for (int c = 0; c < NQBP_COPIES; ++c) {
                uint4 v;
                v.x = d[c] | (d[c + 1] << 8) | (d[c + 2] << 16) | (d[c + 3] << 24);
                v.y = d[c + 4] | (d[c + 5] << 8) | (d[c + 6] << 16) | (d[c + 7] << 24);
                v.z = d[c + 8] | (d[c + 9] << 8) | (d[c + 10] << 16) | (d[c + 11] << 24);
                v.w = d[c + 12] | (d[c + 13] << 8) | (d[c + 14] << 16) | (d[c + 15] << 24);
                *(uint4*)(b.Q + (size_t)(NQBP_COPIES * arr + c) * ldb + y) = v;
            }
        }
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(b.maxcount, mx);
}

#ifndef NQBP_TI_VALUE
#define NQBP_TI_VALUE 16
#endif
#ifndef NQBP_WARPS_VALUE
#define NQBP_WARPS_VALUE 8
#endif
constexpr int NQBP_WARPS = NQBP_WARPS_VALUE;  // warps per CTA: they sweep the same j chunks and share gather sectors in L1
constexpr int NQBP_TI = NQBP_TI_VALUE, NQBP_TJ = 4, NQBP_CHUNK = 128, NQBP_GROUP = NQBP_WARPS * NQBP_TI;
constexpr int NQBP_INF16 = 0x3fff, NQBP_BIAS = 128;
static_assert(NQBP_GROUP == NQBP_CHUNK, "the work-unit map assumes one column group per j chunk");
constexpr int NQBP_SEG = 64;  // default chunks (of 128 columns j) per work unit (NqBig::seg); measured at n = 10^6:
                              // 512 -> 1.58e12, 256 -> 1.75e12, 128 -> 1.90e12, 64 -> 1.91e12, 32 -> 1.79e12 moves/s

// the lane-consecutive operand is read once per tile: keep it out of L1 so the gather windows
// (shared by the CTA's warps) stay resident
__device__ __forceinline__ unsigned int nqbp_ld_stream(const unsigned char* p) {
    unsigned int v;
    asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool DUMP>
__global__ void __launch_bounds__(32 * NQBP_WARPS, 16 / NQBP_WARPS) nqb_scan_packed_kernel(NqBig b) {
    if (*b.maxcount > (unsigned)NQBP_MAX_COUNT) return;  // a line too long for byte sums: nqb_scan_kernel runs
    constexpr int TI = NQBP_TI;
    __shared__ int s_group;
    const int n = b.n, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned char* __restrict__ Q = b.Q;
    const unsigned int* __restrict__ rows = b.rows;
    const unsigned char* __restrict__ cb = b.cb;
    const int ldbm1 = b.ldb - 1, q2off = NQBP_COPIES * b.ldb;
    const int g_first = b.i_begin / NQBP_GROUP;
    const int g_count = b.i_end > b.i_begin ? (b.i_end + NQBP_GROUP - 1) / NQBP_GROUP - g_first : 0;
    int best_v = NQ_INF;
    unsigned int best_i = 0xffffffffu;
    unsigned long long pairs = 0;

    // Work unit = (column group, segment of NQBP_SEG chunks of its j sweep), claimed from one
    // counter in segment-major order.  A unit reports its own best (value, column); the min over
    // units of that pair is the min over the whole slice, so nothing else has to be combined.
    // Segments keep the early (long-sweep) partitions of a multi-GPU split from running in a few
    // coarse waves.
    const int n_chunks_all = (n + NQBP_CHUNK - 1) / NQBP_CHUNK;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_group = (int)atomicAdd(b.tile_counter, 1u);
        __syncthreads();
        int g = s_group, seg = 0;
        bool none = false;
        for (;;) {  // segment seg exists for the groups whose sweep is longer than seg * b.seg chunks
            int gs = n_chunks_all - seg * b.seg - g_first;
            gs = gs < 0 ? 0 : (gs > g_count ? g_count : gs);
            if (gs == 0) {
                none = true;
                break;
            }
            if (g < gs) break;
            g -= gs;
            ++seg;
        }
        if (none) break;
        const int i0 = (g_first + g) * NQBP_GROUP + w * TI;
        if (i0 >= b.i_end || i0 + TI <= b.i_begin || i0 >= n - 1) continue;
        const int jbase = i0 & ~(NQBP_CHUNK - 1);

        int pv1[TI], pv2[TI];
        unsigned NUl[TI / 2], NWl[TI / 2], m[TI / 2];
#pragma unroll
        for (int p = 0; p < TI / 2; ++p) {
            unsigned ul = 0, wl = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int a = 2 * p + h, i = i0 + a;
                const int ri = (int)__ldg(rows + i);  // rows are zero-padded past n (slots masked later)
                const int x1 = n - 1 - ri, x2 = ri;
                // copy (x & 3), word (x & ~3):  (x & 3) * ldb + (x & ~3) == x + (x & 3) * (ldb - 1)
                pv1[a] = x1 + (x1 & 3) * ldbm1 + 4 * lane + jbase;
                pv2[a] = q2off + x2 + (x2 & 3) * ldbm1 + 4 * lane + jbase;
                const int idu = ri - i + n, idw = ri + i;  // diagonal ids, 0 <= id < 2^21 for real columns
                ul |= (unsigned)((2 * (idu & 0x7fff)) & 0xffff) << (16 * h);
                wl |= (unsigned)((2 * (idw & 0x7fff)) & 0xffff) << (16 * h);
                if (lane == 0 && seg == 0 && i >= b.i_begin && i < b.i_end && i < n - 1) pairs += (unsigned long long)(n - 1 - i);
            }
            NUl[p] = ~ul;  // x ^ ~y == ~(x ^ y): equal halves give 0xFFFF
            NWl[p] = ~wl;
            m[p] = (unsigned)NQBP_INF16 * 0x10001u;
        }
        const int A1 = i0 + n - 1, A2 = i0;

        auto chunk = [&](int dj, bool masked) {
            const int j0 = jbase + dj + NQBP_TJ * lane;
            const uint4 r4 = __ldg((const uint4*)(rows + j0));
            const unsigned c4 = __ldg((const unsigned*)(cb + j0));
            // lane-consecutive windows: T[a] = D1[j0..j0+3 - r_ia] + D2[j0..j0+3 + r_ia] (streamed once)
            unsigned T[TI];
#pragma unroll
            for (int a = 0; a < TI; ++a)
                T[a] = nqbp_ld_stream(Q + pv1[a] + dj) + nqbp_ld_stream(Q + pv2[a] + dj);
            unsigned TP[NQBP_TJ][TI / 4];  // byte transpose: TP[b][g] = slots 4g..4g+3 at j_b
#pragma unroll
            for (int g4 = 0; g4 < TI / 4; ++g4) {
                const unsigned x0 = __byte_perm(T[4 * g4 + 0], T[4 * g4 + 1], 0x5140);
                const unsigned x1 = __byte_perm(T[4 * g4 + 2], T[4 * g4 + 3], 0x5140);
                const unsigned y0 = __byte_perm(T[4 * g4 + 0], T[4 * g4 + 1], 0x7362);
                const unsigned y1 = __byte_perm(T[4 * g4 + 2], T[4 * g4 + 3], 0x7362);
                TP[0][g4] = __byte_perm(x0, x1, 0x5410);
                TP[1][g4] = __byte_perm(x0, x1, 0x7632);
                TP[2][g4] = __byte_perm(y0, y1, 0x5410);
                TP[3][g4] = __byte_perm(y0, y1, 0x7632);
            }
#pragma unroll
            for (int bb = 0; bb < NQBP_TJ; ++bb) {
                const int j = j0 + bb;
                const int rj = (int)(bb == 0 ? r4.x : bb == 1 ? r4.y : bb == 2 ? r4.z : r4.w);
                const int cj = (int)(c4 >> (8 * bb)) & 0xff;
                // data-dependent windows over the 16 column slots (shared with the neighbouring warps)
                const int t1 = A1 - rj, t2 = A2 + rj;
                // copy (t & 15), 128-bit word (t & ~15): one sector per load, 16 column slots
                const uint4 v1 = __ldg((const uint4*)(Q + t1 + (t1 & 15) * ldbm1));
                const uint4 v2 = __ldg((const uint4*)(Q + q2off + t2 + (t2 & 15) * ldbm1));
                unsigned X[TI / 4];
                X[0] = v1.x + v2.x + TP[bb][0];
                X[1] = v1.y + v2.y + TP[bb][1];
                X[2] = v1.z + v2.z + TP[bb][2];
                X[3] = v1.w + v2.w + TP[bb][3];
                const unsigned kj = (unsigned)(7 + NQBP_BIAS - cj) * 0x10001u;
                const int idu = rj - j + n, idw = rj + j;
                const unsigned ul = (unsigned)(2 * (idu & 0x7fff)) * 0x10001u;
                const unsigned wl = (unsigned)(2 * (idw & 0x7fff)) * 0x10001u;
                unsigned hit = 0;  // some slot's LOW id half matched: a real attack or a 2^-15 alias
#pragma unroll
                for (int p = 0; p < TI / 2; ++p) {
                    const unsigned x16 = __byte_perm(X[p / 2], 0u, (p & 1) ? 0x4342 : 0x4140);
                    unsigned y = x16 + kj;
                    // low 15 id bits equal: XNOR all ones (-1), else <= 0xFFFD (-3).  An alias (low halves
                    // equal, ids different) makes this value 2 too HIGH; the exact pass below repairs it.
                    const unsigned a2 = __vimax3_u16x2(ul ^ NUl[p], wl ^ NWl[p], 0xFFFDFFFDu);
                    hit |= a2 ^ 0xFFFDFFFDu;
                    if (masked) {
                        const int ia = i0 + 2 * p;
                        unsigned pen = 0;
                        if (!(j > ia && j < n)) pen |= 0x00004000u;
                        if (!(j > ia + 1 && j < n)) pen |= 0x40000000u;
                        y = __vadd2(y, pen);
                    }
                    if (DUMP) {
                        const unsigned z = __vadd2(y, a2);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = i0 + 2 * p + h;
                            if (j > i && j < n && i >= b.i_begin && i < b.i_end) {
                                const int zz = (int)(short)((z >> (16 * h)) & 0xffff);
                                b.dump[nq_swap_index(n, i, j)] = 2ll * (long long)(zz - NQBP_BIAS - (int)cb[i]);
                            }
                        }
                    }
                    m[p] = __viaddmin_s16x2(y, a2, m[p]);
                }
                if (hit) {  // rare (2 * TI / 2^15 per column j): redo this j with the full ids
#pragma unroll
                    for (int p = 0; p < TI / 2; ++p) {  // unrolled: m[] and X[] stay in registers
                        const unsigned x16 = __byte_perm(X[p / 2], 0u, (p & 1) ? 0x4342 : 0x4140);
                        unsigned y = x16 + kj, a2 = 0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = i0 + 2 * p + h;
                            const int ri = (int)__ldg(rows + i);
                            const bool att = (ri - i == rj - j) || (ri + i == rj + j);
                            a2 |= (att ? 0xFFFFu : 0xFFFDu) << (16 * h);
                            if (masked && !(j > i && j < n)) y = __vadd2(y, 0x4000u << (16 * h));
                        }
                        if (DUMP) {
                            const unsigned z = __vadd2(y, a2);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int i = i0 + 2 * p + h;
                                if (j > i && j < n && i >= b.i_begin && i < b.i_end) {
                                    const int zz = (int)(short)((z >> (16 * h)) & 0xffff);
                                    b.dump[nq_swap_index(n, i, j)] = 2ll * (long long)(zz - NQBP_BIAS - (int)cb[i]);
                                }
                            }
                        }
                        m[p] = __viaddmin_s16x2(y, a2, m[p]);  // exact <= the aliased value: the min repairs it
                    }
                }
            }
        };

        const int dj_full = (n & ~(NQBP_CHUNK - 1)) - jbase;  // end of the full chunks
        const int span = n_chunks_all * NQBP_CHUNK - jbase;    // the whole sweep, partial last chunk included
        int dj = seg * b.seg * NQBP_CHUNK;
        int dj_hi = dj + b.seg * NQBP_CHUNK;
        dj_hi = dj_hi < span ? dj_hi : span;
        if (dj == 0) {
            chunk(0, true);  // the chunk holding the tile: needs the j > i mask
            dj = NQBP_CHUNK;
        }
        const int full_end = dj_hi < dj_full ? dj_hi : dj_full;
        for (; dj < full_end; dj += NQBP_CHUNK) chunk(dj, false);
        if (dj < dj_hi) chunk(dj, true);  // the partial last chunk: needs the j < n mask

#pragma unroll
        for (int p = 0; p < TI / 2; ++p)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + 2 * p + h;
                const int mv = (int)(short)((m[p] >> (16 * h)) & 0xffff);
                if (i < n - 1 && i >= b.i_begin && i < b.i_end && mv < NQBP_INF16 / 2) {
                    const int v = mv - NQBP_BIAS - (int)cb[i];
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
    if (lane == 0 && pairs) atomicAdd(b.scored, pairs);
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
            if (x < 0 || x >= n) *bad = 1;  // flagged and stored as 0 (in bounds for the counter build)
            else v = (unsigned)x;
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
