// Packed-window swap scan for n-queens (the fast path of nq_step_kernel_v2).
//
// The scalar scan (nq_kernels.cuh) sits on the shared-memory roofline at 9.6 wavefronts per 32
// moves, 5 of them bank-conflict replays of the two data-dependent gathers.  This path cuts
// wavefronts per move ~4x by making every 32-bit shared-memory word carry FOUR counters:
//
//   * diagonal counters are bytes; each array is stored as 4 copies shifted by 0..3 bytes
//     (copy c, byte y = D[y + c]) so the 4-counter window starting at ANY index x is the
//     aligned word (x & ~3) of copy (x & 3);
//   * a warp owns 8 consecutive columns i (two windows), a lane owns 4 consecutive columns j:
//     the data-dependent gather D[i0 - r_j .. +7] is two words per array per j (serves 8
//     moves), the lane-consecutive read D[j0 - r_i .. +3] is one word per array per i
//     (serves 4 moves, conflict-free);
//   * sums of four byte counters are done as plain 32-bit adds (no carries: the path is only
//     taken while every counter <= 62), transposed with PRMT, widened to 16x2 and finished
//     with VIADD.16x2 / VIMNMX3.U16x2 / VIADDMNMX.S16x2.
//
// Exactness: identical integer value per move as the scalar path (parity-tested through
// cs_nq_neighbourhood_deltas, which runs THIS scan with a dump flag).  Chains that are not
// permutations, have a line with more than 62 queens, or boards outside [NQC_MIN_N,
// NQC_MAX_N] stay on the scalar path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nq_kernels.cuh"

namespace csb {

#ifndef NQC_TI_VALUE
#define NQC_TI_VALUE 16
#endif
#ifndef NQC_UNROLL
#define NQC_UNROLL 2  // chunks per loop trip of the unmasked sweep (the second copy addresses with immediates)
#endif
constexpr int NQC_UNROLL_C = NQC_UNROLL;
constexpr int NQC_TI = NQC_TI_VALUE;  // column slots per warp tile (TI/4 four-byte windows)
constexpr int NQC_TJ = 4;          // columns per lane
constexpr int NQC_CHUNK = 32 * NQC_TJ;
constexpr int NQC_MAX_COUNT = 62;  // 4 counters + slack stay below 256
constexpr int NQC_MIN_N = 256;
constexpr int NQC_MAX_N = 12096;
constexpr int NQC_THREADS = 512;
constexpr int NQC_INF16 = 0x3fff;
constexpr int NQC_UBIAS = 32768;    // 2*(r - c) + UBIAS is a positive 16-bit id for |r - c| < 16384
constexpr int NQC_BIAS = 128;       // keeps X - c_j + 7 positive in each 16-bit half (c_j <= 124)

struct NqSmemC {
    int* red;        // [128]
    uint16_t* rows;  // [n_pad + PAD]
    uint8_t* cb;     // [n_pad + PAD]  own-lines sum per column (<= 63)
    uint8_t* Q1;     // [4][ldb]  copy c, byte y = D1[y + c]
    uint8_t* Q2;     // [4][ldb]
    int ldb;
};

__host__ __device__ inline int nqc_ldb(int n_pad) { return (2 * n_pad + 2 * NQ_PAD + 15) & ~15; }
__host__ __device__ inline size_t nqc_smem_bytes(int n_pad) {
    return 512 + (size_t)(n_pad + NQ_PAD) * 2 + (size_t)(n_pad + NQ_PAD) + (size_t)8 * nqc_ldb(n_pad);
}

// v2 layouts keep the reduction scratch first and rows second so both views share them
__device__ __forceinline__ NqSmemC nqc_carve(unsigned char* base, int n_pad) {
    NqSmemC s;
    s.ldb = nqc_ldb(n_pad);
    s.red = (int*)base;
    s.rows = (uint16_t*)(base + 512);
    s.cb = (uint8_t*)(s.rows + (n_pad + NQ_PAD));
    s.Q1 = s.cb + (n_pad + NQ_PAD);
    s.Q2 = s.Q1 + 4 * s.ldb;
    return s;
}

__device__ __forceinline__ NqSmem nq_carve_v2(unsigned char* base, int n_pad) {
    NqSmem s;
    s.ld = nq_ld(n_pad);
    s.red = (int*)base;
    s.rows = (uint16_t*)(base + 512);
    s.c = s.rows + (n_pad + NQ_PAD);
    s.R = s.c + (n_pad + NQ_PAD);
    s.D1 = s.R + (n_pad + NQ_PAD);
    s.D2 = s.D1 + s.ld;
    return s;
}
__host__ __device__ inline size_t nq_smem_bytes_v2(int n_pad) {
    return 512 + (size_t)3 * (n_pad + NQ_PAD) * 2 + (size_t)2 * nq_ld(n_pad) * 2;
}

__device__ __forceinline__ void nqc_bump(uint8_t* Q, int ldb, int x, int delta) {
    // counter x of the plain array lives at byte (x - c) of copy c
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (x - c >= 0) Q[c * ldb + (x - c)] = (uint8_t)(Q[c * ldb + (x - c)] + delta);
}

// rows (already in smem) -> byte counters in all four copies.  Caller guarantees every line
// count <= NQC_MAX_COUNT (checked on the u16 build).
__device__ void nqc_build(const NqSmemC& s, int n, int n_pad) {
    const int tid = threadIdx.x, nt = blockDim.x;
    {
        const int total16 = ((n_pad + NQ_PAD) + 8 * s.ldb) / 16;  // cb + Q1 + Q2
        uint4* z = (uint4*)s.cb;
        for (int k = tid; k < total16; k += nt) z[k] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        const int r = s.rows[j];
        const int x = j - r + n - 1, y = j + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (x - c >= 0) {
                const int b = c * s.ldb + (x - c);
                atomicAdd((unsigned int*)(s.Q1 + (b & ~3)), 1u << (8 * (b & 3)));
            }
            if (y - c >= 0) {
                const int b = c * s.ldb + (y - c);
                atomicAdd((unsigned int*)(s.Q2 + (b & ~3)), 1u << (8 * (b & 3)));
            }
        }
    }
    __syncthreads();
}

// c_j bytes from copy 0 (the plain arrays); returns the block-wide max LINE COUNT (all threads)
__device__ __forceinline__ int nqc_compute_cb(const NqSmemC& s, int n) {
    int mx = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int r = s.rows[j];
        const int a = (int)s.Q1[j - r + n - 1], b = (int)s.Q2[j + r];
        s.cb[j] = (uint8_t)(a + b);
        mx = max(mx, max(a, b));  // every occupied line has a queen, so this is the max line count
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    __syncthreads();
    if (threadIdx.x == 0) s.red[100] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&s.red[100], mx);
    __syncthreads();
    const int out = s.red[100];
    __syncthreads();
    return out;
}

__device__ __forceinline__ int nqc_swap_half(const NqSmemC& s, int n, int i, int j) {
    const int ri = s.rows[i], rj = s.rows[j];
    const int ci = s.Q1[i - ri + n - 1] + s.Q2[i + ri];
    const int cj = s.Q1[j - rj + n - 1] + s.Q2[j + rj];
    const int g = s.Q1[i - rj + n - 1] + s.Q2[i + rj] + s.Q1[j - ri + n - 1] + s.Q2[j + ri];
    const int d = j - i, t = rj - ri;
    return g - ci - cj + 4 + 2 * ((t == d) | (t == -d));
}

__device__ __forceinline__ void nqc_apply_swap(const NqSmemC& s, int n, int i, int j) {
    const int ri = s.rows[i], rj = s.rows[j];
    nqc_bump(s.Q1, s.ldb, i - ri + n - 1, -1);
    nqc_bump(s.Q2, s.ldb, i + ri, -1);
    nqc_bump(s.Q1, s.ldb, j - rj + n - 1, -1);
    nqc_bump(s.Q2, s.ldb, j + rj, -1);
    nqc_bump(s.Q1, s.ldb, i - rj + n - 1, +1);
    nqc_bump(s.Q2, s.ldb, i + rj, +1);
    nqc_bump(s.Q1, s.ldb, j - ri + n - 1, +1);
    nqc_bump(s.Q2, s.ldb, j + ri, +1);
    s.rows[i] = (uint16_t)rj;
    s.rows[j] = (uint16_t)ri;
}

__device__ __forceinline__ unsigned bcast16(unsigned x) { return __byte_perm(x, x, 0x1010); }

// The packed scan.  Tracks, per column i of the tile, min over j>i of
//   Z = X - c_j + 4 + 2*att   (X = sum of the four counters),   value = Z - c_i.
template <bool DUMP>
__device__ __forceinline__ void nqc_scan_swap(const NqSmemC& s, int n, int& best_v,
                                              unsigned int& best_i, long long* dump) {
    constexpr int TI = NQC_TI;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int num_tiles = (n - 1 + TI - 1) / TI;
    const unsigned char* smem0 = (const unsigned char*)s.Q1;  // Q2 = Q1 + 4*ldb
    const int ldbm1 = s.ldb - 1, q2off = 4 * s.ldb;
    best_v = NQ_INF;
    best_i = 0xffffffffu;

    for (int k = 0; k * W < num_tiles; ++k) {
        const int t = k * W + ((k & 1) ? (W - 1 - w) : w);
        if (t >= num_tiles) continue;
        const int i0 = t * TI;  // multiple of TI (>= 8, so the copy select is tile-invariant)
        const int jbase = i0 & ~(NQC_CHUNK - 1);  // first chunk (holds the tile); the loop
                                                  // offset dj below is warp-uniform by construction
        // per-slot lane-consecutive read offsets (copy fixed by the slot's row, word = 4*lane)
        int pv1[TI], pv2[TI];
        unsigned NU[TI / 2], NW[TI / 2], m[TI / 2];  // complemented diagonal ids: x ^ ~y == ~(x ^ y)
#pragma unroll
        for (int p = 0; p < TI / 2; ++p) {
            unsigned uu = 0, ww = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int a = 2 * p + h, i = i0 + a;
                const int ri = s.rows[i];  // padded read for i >= n (masked later)
                const int x1 = n - 1 - ri, x2 = ri;
                // copy (x & 3), word (x & ~3):  (x & 3) * ldb + (x & ~3) == x + (x & 3) * (ldb - 1)
                pv1[a] = x1 + (x1 & 3) * ldbm1 + 4 * lane + jbase;
                pv2[a] = q2off + x2 + (x2 & 3) * ldbm1 + 4 * lane + jbase;
                uu |= (unsigned)((2 * (ri - i) + NQC_UBIAS) & 0xffff) << (16 * h);
                ww |= (unsigned)((2 * (ri + i)) & 0xffff) << (16 * h);
            }
            NU[p] = ~uu;
            NW[p] = ~ww;
            m[p] = (unsigned)NQC_INF16 * 0x10001u;
        }
        const int A1 = i0 + n - 1, A2 = i0;

        const unsigned char* rowp = (const unsigned char*)(s.rows + jbase + NQC_TJ * lane);
        const unsigned char* cbp = (const unsigned char*)(s.cb + jbase + NQC_TJ * lane);
        auto chunk = [&](int dj, bool masked) {
            const int jc = jbase + dj;
            const int j0 = jc + NQC_TJ * lane;
            const uint2 r4 = *(const uint2*)(rowp + 2 * dj);  // 4 rows (u16)
            const unsigned c4 = *(const unsigned*)(cbp + dj);  // 4 c_j bytes
            // lane-consecutive windows: T[a] = D1[j0..j0+3 - r_ia] + D2[j0..j0+3 + r_ia]
            unsigned T[TI];
#pragma unroll
            for (int a = 0; a < TI; ++a)
                T[a] = *(const unsigned*)(smem0 + pv1[a] + dj) + *(const unsigned*)(smem0 + pv2[a] + dj);
            // transpose bytes: TP[b][g] = slots 4g..4g+3 at j_b
            unsigned TP[NQC_TJ][TI / 4];
#pragma unroll
            for (int g = 0; g < TI / 4; ++g) {
                const unsigned x0 = __byte_perm(T[4 * g + 0], T[4 * g + 1], 0x5140);
                const unsigned x1 = __byte_perm(T[4 * g + 2], T[4 * g + 3], 0x5140);
                const unsigned y0 = __byte_perm(T[4 * g + 0], T[4 * g + 1], 0x7362);
                const unsigned y1 = __byte_perm(T[4 * g + 2], T[4 * g + 3], 0x7362);
                TP[0][g] = __byte_perm(x0, x1, 0x5410);
                TP[1][g] = __byte_perm(x0, x1, 0x7632);
                TP[2][g] = __byte_perm(y0, y1, 0x5410);
                TP[3][g] = __byte_perm(y0, y1, 0x7632);
            }
#pragma unroll
            for (int b = 0; b < NQC_TJ; ++b) {
                const int j = j0 + b;
                const int rj = (int)((b < 2 ? r4.x : r4.y) >> (16 * (b & 1))) & 0xffff;
                const int cj = (int)(c4 >> (8 * b)) & 0xff;
                // data-dependent windows over the 8 column slots
                const int t1 = A1 - rj, t2 = A2 + rj;
                const unsigned char* g1 = smem0 + t1 + (t1 & 3) * ldbm1;
                const unsigned char* g2 = smem0 + q2off + t2 + (t2 & 3) * ldbm1;
                unsigned X[TI / 4];
#pragma unroll
                for (int g = 0; g < TI / 4; ++g)
                    X[g] = *(const unsigned*)(g1 + 4 * g) + *(const unsigned*)(g2 + 4 * g) + TP[b][g];
                // bcast16 takes the low 16 bits; kj is biased by NQC_BIAS so both halves stay
                // positive and the packed add below is a plain 32-bit add (no inter-half carry)
                // positive 16-bit values broadcast to both halves by a multiply (FMA pipe, no PRMT)
                const unsigned kj = (unsigned)(7 + NQC_BIAS - cj) * 0x10001u;
                const unsigned ub = (unsigned)(2 * (rj - j) + NQC_UBIAS) * 0x10001u;
                const unsigned wb = (unsigned)(2 * (rj + j)) * 0x10001u;
#pragma unroll
                for (int p = 0; p < TI / 2; ++p) {
                    const unsigned x16 = __byte_perm(X[p / 2], 0u, (p & 1) ? 0x4342 : 0x4140);
                    unsigned y = x16 + kj;
                    // att: equal diagonal ids -> XNOR = 0xFFFF (= -1), else <= 0xFFFD (= -3)
                    const unsigned a2 = __vimax3_u16x2(ub ^ NU[p], wb ^ NW[p], 0xFFFDFFFDu);
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
                            if (j > i && j < n) {
                                const int zz = (int)(short)((z >> (16 * h)) & 0xffff);
                                dump[nq_swap_index(n, i, j)] = 2ll * (long long)(zz - NQC_BIAS - (int)s.cb[i]);
                            }
                        }
                    }
                    m[p] = __viaddmin_s16x2(y, a2, m[p]);
                }
            }
        };

        chunk(0, true);  // the chunk holding the tile: needs the j > i mask
        const int dj_full = (n & ~(NQC_CHUNK - 1)) - jbase;  // end of the full chunks
        int dj = NQC_CHUNK;
#pragma unroll NQC_UNROLL_C
        for (; dj < dj_full; dj += NQC_CHUNK) chunk(dj, false);
        if (jbase + dj < n) chunk(dj, true);

#pragma unroll
        for (int p = 0; p < TI / 2; ++p)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + 2 * p + h;
                const int mv = (int)(short)((m[p] >> (16 * h)) & 0xffff);
                if (i < n - 1 && mv < NQC_INF16 / 2) {
                    const int v = mv - NQC_BIAS - (int)s.cb[i];
                    if (v < best_v) {
                        best_v = v;
                        best_i = (unsigned)i;
                    }
                }
            }
    }
}


// u16 counters from rows already in smem (scalar layout), e.g. when a chain leaves the packed path
__device__ void nq_count_from_rows(const NqSmem& s, int n, int n_pad) {
    const int tid = threadIdx.x, nt = blockDim.x;
    {
        const int total16 = (int)(((size_t)2 * (n_pad + NQ_PAD) * 2 + (size_t)2 * s.ld * 2) / 16);
        uint4* z = (uint4*)s.c;  // c, R, D1, D2 are contiguous after rows
        for (int k = tid; k < total16; k += nt) z[k] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        const int r = s.rows[j];
        smem_inc16(s.R, r);
        smem_inc16(s.D1, j - r + n - 1);
        smem_inc16(s.D2, j + r);
    }
    __syncthreads();
}

__device__ __forceinline__ int nq_max_line_count(const NqSmem& s, int n) {
    int mx = 0;
    for (int k = threadIdx.x; k < 2 * n - 1; k += blockDim.x) mx = max(mx, max((int)s.D1[k], (int)s.D2[k]));
    mx = __reduce_max_sync(0xffffffffu, mx);
    __syncthreads();
    if (threadIdx.x == 0) s.red[100] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(&s.red[100], mx);
    __syncthreads();
    const int out = s.red[100];
    __syncthreads();
    return out;
}

// Step kernel with the packed fast path.  Same chain loop and LocalSearch bookkeeping as
// nq_step_kernel; per chain-step the scan runs on the packed layout while the chain is a
// permutation with every line count <= NQC_MAX_COUNT, otherwise on the scalar layout.
template <int TI_A>
__global__ void __launch_bounds__(NQC_THREADS, 1) nq_step_kernel_v2(NqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NqSmem s = nq_carve_v2(smem_raw, p.n_pad);
    const NqSmemC sc = nqc_carve(smem_raw, p.n_pad);
    const int n = p.n, tid = threadIdx.x;

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
        bool packed = perm && !p.force_scalar && nq_max_line_count(s, n) <= NQC_MAX_COUNT;
        if (packed) nqc_build(sc, n, p.n_pad);
        const long long nbh = (long long)n * (n - 1) / 2 - ident_pairs;
        long long best_score = p.ls_mode ? score : st->best_score;
        unsigned long long no_improve = 0;
        const unsigned int steps0 = st->steps;
        unsigned int steps = steps0;
        unsigned long long scored = 0;
        unsigned int status = 0;
        if (p.ls_mode) {
            for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
                ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] = ((const uint4*)s.rows)[k];
        }

        for (unsigned long long it = 0; it < p.max_steps; ++it) {
            if (score == 0 && !p.dump) {
                status = 1;
                best_score = 0;
                break;
            }
            int v;
            unsigned int a;
            if (packed) {
                if (nqc_compute_cb(sc, n) > NQC_MAX_COUNT) {  // a line grew too long: leave the fast path
                    packed = false;
                    nq_count_from_rows(s, n, p.n_pad);
                }
            }
            if (packed) {
                if (p.dump) nqc_scan_swap<true>(sc, n, v, a, p.dump);
                else nqc_scan_swap<false>(sc, n, v, a, nullptr);
            } else {
                nq_compute_c(s, n);
                __syncthreads();
                if (perm) {
                    if (p.dump) nq_scan_swap<TI_A, true, true>(s, n, v, a, p.dump);
                    else nq_scan_swap<TI_A, true, false>(s, n, v, a, nullptr);
                } else {
                    if (p.dump) nq_scan_swap<TI_A, false, true>(s, n, v, a, p.dump);
                    else nq_scan_swap<TI_A, false, false>(s, n, v, a, nullptr);
                }
            }
            block_argmin(v, a, s.red);
            scored += (unsigned long long)nbh;
            if (p.dump) break;
            if (v >= NQ_INF) {
                status = 3;
                break;
            }
            unsigned int b = 0xffffffffu;
            {
                const int ra = s.rows[a];
                for (int j = (int)a + 1 + tid; j < n; j += blockDim.x) {
                    if (s.rows[j] == ra) continue;
                    const int h = packed ? nqc_swap_half(sc, n, (int)a, j) : nq_swap_half(s, n, (int)a, j);
                    if (h == v) {
                        b = (unsigned)j;
                        break;
                    }
                }
            }
            int dummy = 0;
            block_argmin(dummy, b, s.red);

            const long long new_score = score + 2ll * v;
            const bool improved = new_score < score;
            if (!improved) {
                ++no_improve;
                if (p.allow_no_improve && no_improve >= p.allow_no_improve) {
                    status = 2;
                    break;
                }
            } else {
                no_improve = 0;
            }
            if (tid == 0) {
                const int i = (int)a, j = (int)b;
                if (packed) {
                    nqc_apply_swap(sc, n, i, j);
                } else {
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
            if (improved) {
                best_score = new_score;
                for (int k = tid; k < p.n_pad / 8; k += blockDim.x)
                    ((uint4*)(p.best_rows + (size_t)chain * p.n_pad))[k] = ((const uint4*)s.rows)[k];
            }
        }
        __syncthreads();
        if (p.dump) continue;
        for (int k = tid; k < p.n_pad / 8; k += blockDim.x) ((uint4*)grow)[k] = ((const uint4*)s.rows)[k];
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

}  // namespace csb
