// Employee-scheduling (on-call rota) chain kernels for sm_100a.
//
// Score definition: examples/employee-scheduling/src/lib.rs:261-375 (+ :194-218).  One
// employee per calendar day; D scored days (<= 64), E employees (dense index 0..E-1).
//
// Device formulation.  For every employee e keep the 64-bit day mask m_e (bit d <=> a[d]==e).
// All four hard terms and S1 are sums over employees of a function of m_e alone:
//   H1_e = popc(m & holiday_e)                                   (:273-280)
//   H2_e = popc(m & m>>1)                                        (:286-292)
//   H3_e = popc(m & m>>7 & SATF) + popc(m & m>>8 & SATF)
//        + popc(m>>1 & m>>7 & SATF) + popc(m>>1 & m>>8 & SATF)   (:295-315)
//   H4_e = #{w : popc(m & W14<<w) > 3}                           (:318-327)
//   S1_e = #{w : popc(m & W7<<w)  > 2}                           (:330-339)
// so a move's delta on these terms is F(m') - F(m) over the (at most two) employees whose
// mask changes, with the window loops restricted to windows that overlap a changed day.
// Hot-loop formulation (es_prepare + es_change_delta / es_swap_delta): once per chain-step,
// for every PRESENT employee (<= D of them) all sliding-window counts are computed at once by a
// bit-sliced adder over the day mask, kept as "count == k" window-start masks
// (EQ3_14, EQ4_14, EQ2_7, EQ3_7); per day d the terms that only depend on the day's current
// employee are tabulated.  A candidate then needs three 64-bit popcounts for H2+H3, H4 and S1:
//   gain of day d for employee e = popc(m_e & PART[d]) + popc(EQ3_14[e] & CONT14[d]) ...
// where PART[d] = days paired with d by H2/H3 and CONT[d] = window starts whose window holds d.
// S2 (weekday affinity, min over employees present on that weekday), S3 (max-min of total
// days over PRESENT employees) and S4 (max-min of weekend days over present employees) are
// kept as count histograms + occupancy bitsets, so min/max after a move are bit scans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace csb {

typedef unsigned long long u64;

constexpr int ES_MAX_DAYS = 64;
constexpr int ES_CBINS = 12;   // per-weekday count bins 0..11 (a weekday occurs <= 10 times in 64 days)
constexpr int ES_TBINS = 66;   // total-days bins 0..65
constexpr int ES_WBINS = 24;   // weekend-day bins 0..23
constexpr long long ES_KEY_INF = 0x7fffffffffffffffll;
constexpr int ES_KEYS = 66 * 32;  // (total days 0..65) x (weekend days 0..31)
constexpr int ES_MAXCLS = 68;

struct EsConst {
    int D, E, start_wd, n14, n7;
    u64 valid, wkend, satf;
    u64 wd[7];
};

struct EsChainState {
    long long hard, soft, best_hard, best_soft;
    unsigned long long moves_scored;
    unsigned int steps, status;
};

struct EsTraceEntry {
    unsigned int kind, x, y, pad;
    int hard_after, soft_after;
};

struct EsParams {
    EsConst K;
    int first_chain, n_chains, stride;  // stride = D + 1 slots (phantom last, lib.rs:405-412)
    uint16_t* a;                        // [*, stride] employee index per slot
    uint16_t* best_a;
    const u64* hol;  // [E] holiday day-mask per employee
    const u64* dayconst;  // [3][64]: PART (H2/H3 partner days), CONT14, CONT7 (window starts holding d)
    EsChainState* st;
    EsTraceEntry* trace;
    int trace_cap;
    unsigned int* work_counter;
    unsigned long long* totals;
    unsigned long long max_steps, allow_no_improve;
    int ls_mode;
    long long* dump_h;  // debug: every candidate's (dhard, dsoft)
    long long* dump_s;
    const unsigned int* skip;  // optional [chains]: 1 = leave the chain alone (ILS)
};

// per-chain shared state
struct EsSmem {
    u64* mask;          // [E]
    uint16_t* a;        // [stride]
    uint16_t* hist2;    // [5][ES_CBINS]
    uint16_t* histT;    // [ES_TBINS]
    uint16_t* histW;    // [ES_WBINS]
    unsigned int* occ2; // [5]  bit c: some employee has exactly c days on that weekday (c>=1)
    u64* occT;          // [1]  bit c: some employee has exactly c days in total (c>=1)
    unsigned int* occW; // [1]  bit c: some PRESENT employee has exactly c weekend days (c>=0)
    int* misc;          // [16] present, distinct[5], hard, soft, ...
    u64* red;           // [40] reduction scratch
    u64* part;          // [64] H2/H3 partner-day mask per day
    u64* cont14;        // [64] 14-day window starts whose window contains the day
    u64* cont7;         // [64]
    u64* eq;            // [64][4] per present-employee slot: EQ3_14, EQ4_14, EQ2_7, EQ3_7
    unsigned char* dayb;   // [5][64] per day, for its current employee: lossH, lossS1, total, weekend, weekday count
    unsigned char* pslot;  // [E] present-employee slot (0xFF = absent)
    unsigned char* cls;    // [E] class of the employee's (total days, weekend days) pair
    unsigned char* keymap; // [ES_KEYS] (total << 5 | weekend) -> class id
    unsigned int* keybits; // [ES_KEYS / 32] keys in use
    signed char* s34;      // [64][ES_MAXCLS] S3+S4 delta of giving day d to an employee of the class
    signed char* s2t;      // [64][ES_CBINS] S2 delta of giving day d to an employee with cn days on that weekday
};

struct EsLayout {
    size_t mask, a, hist, occ, occT, misc, red, day, eq, dayb, pslot, cls, keymap, keybits, s34, s2t, total;
};
__host__ __device__ inline size_t es_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline EsLayout es_layout(int D, int E) {
    EsLayout L;
    size_t o = 0;
    L.mask = o;    o += (size_t)E * 8;
    L.a = o;       o = es_align(o + (size_t)(D + 1) * 2, 8);
    L.hist = o;    o = es_align(o + (5 * ES_CBINS + ES_TBINS + ES_WBINS) * 2, 8);
    L.occ = o;     o = es_align(o + 6 * 4, 8);
    L.occT = o;    o += 8;
    L.misc = o;    o += 16 * 4;
    L.red = o;     o += 40 * 8;
    L.day = o;     o += 3 * 64 * 8;
    L.eq = o;      o += 64 * 4 * 8;
    L.dayb = o;    o += 5 * 64;
    L.pslot = o;   o = es_align(o + (size_t)E, 8);
    L.cls = o;     o = es_align(o + (size_t)E, 8);
    L.keymap = o;  o = es_align(o + ES_KEYS, 8);
    L.keybits = o; o = es_align(o + (ES_KEYS / 32) * 4, 8);
    L.s34 = o;     o = es_align(o + 64 * ES_MAXCLS, 8);
    L.s2t = o;     o = es_align(o + 64 * ES_CBINS, 8);
    L.total = o;
    return L;
}
__host__ __device__ inline size_t es_smem_bytes(int D, int E) { return es_layout(D, E).total; }

// offsets only (no pointer <-> integer casts) so the compiler keeps the shared address space
__device__ __forceinline__ EsSmem es_carve(unsigned char* p, int D, int E) {
    const EsLayout L = es_layout(D, E);
    EsSmem s;
    s.mask = (u64*)(p + L.mask);
    s.a = (uint16_t*)(p + L.a);
    s.hist2 = (uint16_t*)(p + L.hist);
    s.histT = s.hist2 + 5 * ES_CBINS;
    s.histW = s.histT + ES_TBINS;
    s.occ2 = (unsigned int*)(p + L.occ);
    s.occW = s.occ2 + 5;
    s.occT = (u64*)(p + L.occT);
    s.misc = (int*)(p + L.misc);
    s.red = (u64*)(p + L.red);
    s.part = (u64*)(p + L.day);
    s.cont14 = s.part + 64;
    s.cont7 = s.cont14 + 64;
    s.eq = (u64*)(p + L.eq);
    s.dayb = p + L.dayb;
    s.pslot = p + L.pslot;
    s.cls = p + L.cls;
    s.keymap = p + L.keymap;
    s.keybits = (unsigned int*)(p + L.keybits);
    s.s34 = (signed char*)(p + L.s34);
    s.s2t = (signed char*)(p + L.s2t);
    return s;
}

enum { ES_PRESENT = 0, ES_DISTINCT0 = 1, ES_HARD = 6, ES_SOFT = 7, ES_BCAST = 8, ES_NSLOT = 10, ES_NCLS = 11 };
enum { ES_DB_LOSSH = 0, ES_DB_LOSSS = 64, ES_DB_TOT = 128, ES_DB_WK = 192, ES_DB_WD = 256 };

// ------------------------------------------------------------------ per-employee terms
__device__ __forceinline__ int es_pair_terms(u64 m, u64 hol, const EsConst& K) {
    const u64 m1 = m >> 1, m7 = m >> 7, m8 = m >> 8;
    return __popcll(m & hol) + __popcll(m & m1) + __popcll(m & m7 & K.satf) +
           __popcll(m & m8 & K.satf) + __popcll(m1 & m7 & K.satf) + __popcll(m1 & m8 & K.satf);
}

// windows w in [lo, hi]: #{popc(m & W<<w) > thr}
__device__ __forceinline__ int es_win_viol(u64 m, u64 W, int lo, int hi, int thr) {
    int v = 0;
    for (int w = lo; w <= hi; ++w) v += (__popcll(m & (W << w)) > thr);
    return v;
}

// (hard, S1) of employee mask m over all windows
__device__ __forceinline__ void es_emp_full(u64 m, u64 hol, const EsConst& K, int& hard, int& s1) {
    hard = es_pair_terms(m, hol, K) + es_win_viol(m, 0x3fffull, 0, K.n14 - 1, 3);
    s1 = es_win_viol(m, 0x7full, 0, K.n7 - 1, 2);
}

// delta of (hard, S1) when an employee's mask changes m -> m2
__device__ __forceinline__ void es_emp_delta(u64 m, u64 m2, u64 hol, const EsConst& K, int& dh,
                                             int& ds) {
    dh += es_pair_terms(m2, hol, K) - es_pair_terms(m, hol, K);
    const u64 x = m ^ m2;
    const int lo_d = __ffsll((long long)x) - 1, hi_d = 63 - __clzll((long long)x);
    {   // 14-day windows touching a changed day
        const int lo = max(0, lo_d - 13), hi = min(K.n14 - 1, hi_d);
        for (int w = lo; w <= hi; ++w) {
            const u64 W = 0x3fffull << w;
            if (!(W & x)) continue;
            dh += (__popcll(m2 & W) > 3) - (__popcll(m & W) > 3);
        }
    }
    {   // 7-day windows
        const int lo = max(0, lo_d - 6), hi = min(K.n7 - 1, hi_d);
        for (int w = lo; w <= hi; ++w) {
            const u64 W = 0x7full << w;
            if (!(W & x)) continue;
            ds += (__popcll(m2 & W) > 2) - (__popcll(m & W) > 2);
        }
    }
}

// ------------------------------------------------------------------ histogram helpers
// up to four (bin, +-1) histogram adjustments, always four slots so everything stays in
// registers (unused slots carry delta 0 on a valid bin)
struct EsAdj {
    int b0, b1, b2, b3, d0, d1, d2, d3, n;
    __device__ __forceinline__ EsAdj() : b0(0), b1(0), b2(0), b3(0), d0(0), d1(0), d2(0), d3(0), n(0) {}
    __device__ __forceinline__ void add(int b, int delta) {
        if (n == 0) { b0 = b; d0 = delta; }
        else if (n == 1) { b1 = b; d1 = delta; }
        else if (n == 2) { b2 = b; d2 = delta; }
        else { b3 = b; d3 = delta; }
        ++n;
    }
};

__device__ __forceinline__ u64 es_occ_one(const uint16_t* hist, u64 occ, int b, const EsAdj& A) {
    const int tot = (A.b0 == b ? A.d0 : 0) + (A.b1 == b ? A.d1 : 0) + (A.b2 == b ? A.d2 : 0) +
                    (A.b3 == b ? A.d3 : 0);
    const int c = (int)hist[b] + tot;
    const u64 bit = 1ull << b;
    return c > 0 ? (occ | bit) : (occ & ~bit);
}

// occupancy bitset after applying the adjustments to the histogram
__device__ __forceinline__ u64 es_occ_after(const uint16_t* hist, u64 occ, const EsAdj& A) {
    occ = es_occ_one(hist, occ, A.b0, A);
    occ = es_occ_one(hist, occ, A.b1, A);
    occ = es_occ_one(hist, occ, A.b2, A);
    occ = es_occ_one(hist, occ, A.b3, A);
    return occ;
}

__device__ __forceinline__ int es_spread(u64 occ, int members) {  // max-min, lib.rs:349,363
    if (members < 2 || occ == 0) return 0;
    return (63 - __clzll((long long)occ)) - (__ffsll((long long)occ) - 1);
}

__device__ __forceinline__ int es_s2_term(unsigned int occ, int distinct) {  // lib.rs:206-214
    return (distinct >= 2 && occ) ? (__ffs((int)occ) - 1) : 0;
}

// weekday-affinity delta on weekday wd when employee counts change: cm (count c -> c-1) and
// cp (count c -> c+1); pass -1 to skip one side.
__device__ __forceinline__ int es_s2_delta(const EsSmem& s, int wd, int cm, int cp) {
    EsAdj A;
    int distinct = s.misc[ES_DISTINCT0 + wd];
    const int old = es_s2_term(s.occ2[wd], distinct);
    if (cm >= 1) {
        A.add(cm, -1);
        if (cm - 1 >= 1) A.add(cm - 1, +1);
        else --distinct;
    }
    if (cp >= 0) {
        if (cp >= 1) A.add(cp, -1);
        else ++distinct;
        A.add(cp + 1, +1);
    }
    const unsigned int occ = (unsigned int)es_occ_after(s.hist2 + wd * ES_CBINS, s.occ2[wd], A);
    return es_s2_term(occ, distinct) - old;
}

__device__ __forceinline__ int es_weekday(const EsConst& K, int d) { return (K.start_wd + d) % 7; }

// S3 + S4 delta when a day (weekend flag isw) moves from an employee with (to, wo) total /
// weekend days to one with (tn, wn); lib.rs:345-365, min/max over PRESENT employees only.
__device__ __forceinline__ int es_s34_change(const EsSmem& s, int to, int wo, int isw, int tn, int wn) {
    int present = s.misc[ES_PRESENT];
    const int oldT = es_spread(*s.occT, present), oldW = es_spread((u64)*s.occW, present);
    EsAdj T, W;
    T.add(to, -1);
    W.add(wo, -1);
    if (to - 1 >= 1) {
        T.add(to - 1, +1);
        W.add(wo - isw, +1);
    } else {
        --present;
    }
    if (tn >= 1) {
        T.add(tn, -1);
        W.add(wn, -1);
    } else {
        ++present;
    }
    T.add(tn + 1, +1);
    W.add(wn + isw, +1);
    const u64 occT = es_occ_after(s.histT, *s.occT, T);
    const u64 occW = es_occ_after(s.histW, (u64)*s.occW, W);
    return es_spread(occT, present) - oldT + es_spread(occW, present) - oldW;
}

// ------------------------------------------------------------------ per-step tables
// All sliding-window counts of one day mask at once: bit-sliced adder over the L shifted
// copies of m; plane i bit w = bit i of popc(m & (ONES(L) << w)).
template <int L, int PLANES>
__device__ __forceinline__ void es_window_planes(u64 m, u64 (&pl)[PLANES]) {
#pragma unroll
    for (int i = 0; i < PLANES; ++i) pl[i] = 0;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        u64 carry = m >> k;
#pragma unroll
        for (int i = 0; i < PLANES; ++i) {
            const u64 t = pl[i] & carry;
            pl[i] ^= carry;
            carry = t;
        }
    }
}

// Once per chain-step (masks must be current): present-employee slots, their "count == k"
// window masks, and the per-day tables of the day's current employee.  Block-cooperative.
__device__ void es_prepare(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) s.misc[ES_NSLOT] = 0;
    __syncthreads();
    const u64 v14 = K.n14 >= 64 ? ~0ull : ((1ull << K.n14) - 1);  // real window starts only
    const u64 v7 = K.n7 >= 64 ? ~0ull : ((1ull << K.n7) - 1);
    for (int e = tid; e < K.E; e += nt) {
        const u64 m = s.mask[e];
        if (!m) {
            s.pslot[e] = 0xFF;
            continue;
        }
        const int slot = atomicAdd(&s.misc[ES_NSLOT], 1);
        s.pslot[e] = (unsigned char)slot;
        u64 p14[4], p7[3];
        es_window_planes<14, 4>(m, p14);
        es_window_planes<7, 3>(m, p7);
        u64* q = s.eq + slot * 4;
        q[0] = p14[0] & p14[1] & ~p14[2] & ~p14[3] & v14;   // count == 3 (one more => H4 violation)
        q[1] = ~p14[0] & ~p14[1] & p14[2] & ~p14[3] & v14;  // count == 4 (one less => violation gone)
        q[2] = ~p7[0] & p7[1] & ~p7[2] & v7;                // count == 2
        q[3] = p7[0] & p7[1] & ~p7[2] & v7;                 // count == 3
    }
    __syncthreads();
    for (int d = tid; d < K.D; d += nt) {
        const int eo = s.a[d];
        const u64 m = s.mask[eo];
        const u64* q = s.eq + (int)s.pslot[eo] * 4;
        const int wd = (K.start_wd + d) % 7;
        s.dayb[ES_DB_LOSSH + d] = (unsigned char)(((hol[eo] >> d) & 1ull) + __popcll(m & s.part[d]) +
                                                  __popcll(q[1] & s.cont14[d]));
        s.dayb[ES_DB_LOSSS + d] = (unsigned char)__popcll(q[3] & s.cont7[d]);
        s.dayb[ES_DB_TOT + d] = (unsigned char)__popcll(m);
        s.dayb[ES_DB_WK + d] = (unsigned char)__popcll(m & K.wkend);
        s.dayb[ES_DB_WD + d] = (unsigned char)(wd < 5 ? __popcll(m & K.wd[wd]) : 0);
    }
    // classes of (total days, weekend days): the soft S3+S4 delta of a change move depends on
    // the receiving employee only through this pair, so it is tabulated per (day, class)
    for (int k = tid; k < ES_KEYS / 32; k += nt) s.keybits[k] = 0;
    __syncthreads();
    for (int e = tid; e < K.E; e += nt) {
        const u64 m = s.mask[e];
        const int key = (__popcll(m) << 5) | __popcll(m & K.wkend);
        atomicOr(&s.keybits[key >> 5], 1u << (key & 31));
    }
    __syncthreads();
    if (tid == 0) {  // deterministic class ids in ascending key order
        int nc = 0;
        for (int w = 0; w < ES_KEYS / 32; ++w) {
            unsigned bits = s.keybits[w];
            while (bits) {
                const int b = __ffs((int)bits) - 1;
                bits &= bits - 1;
                s.keymap[w * 32 + b] = (unsigned char)nc;
                ((unsigned short*)s.red)[nc] = (unsigned short)(w * 32 + b);  // class -> key
                ++nc;
            }
        }
        s.misc[ES_NCLS] = nc;
    }
    __syncthreads();
    const int ncls = s.misc[ES_NCLS];
    for (int e = tid; e < K.E; e += nt) {
        const u64 m = s.mask[e];
        s.cls[e] = s.keymap[(__popcll(m) << 5) | __popcll(m & K.wkend)];
    }
    for (int k = tid; k < K.D * ncls; k += nt) {
        const int d = k / ncls, c = k - d * ncls;
        const int key = ((const unsigned short*)s.red)[c];
        const int isw = (K.wkend >> d) & 1ull ? 1 : 0;
        s.s34[d * ES_MAXCLS + c] = (signed char)es_s34_change(
            s, s.dayb[ES_DB_TOT + d], s.dayb[ES_DB_WK + d], isw, key >> 5, key & 31);
    }
    for (int k = tid; k < K.D * ES_CBINS; k += nt) {
        const int d = k / ES_CBINS, cn = k - d * ES_CBINS;
        const int wd = (K.start_wd + d) % 7;
        s.s2t[k] = (signed char)((wd < 5 && cn < ES_CBINS - 1)
                                     ? es_s2_delta(s, wd, (int)s.dayb[ES_DB_WD + d], cn) : 0);
    }
    __syncthreads();
}

// ------------------------------------------------------------------ move deltas
// change: day d gets employee en (!= current).  Returns (dhard, dsoft).
__device__ __forceinline__ void es_change_delta(const EsSmem& s, const EsConst& K,
                                                const u64* __restrict__ hol, int d, int en,
                                                int& dh, int& ds) {
    const u64 mn = s.mask[en];
    dh = (int)((hol[en] >> d) & 1ull) - (int)s.dayb[ES_DB_LOSSH + d];
    ds = -(int)s.dayb[ES_DB_LOSSS + d];
    int cn = 0;
    const int wd = es_weekday(K, d);
    if (mn) {  // an absent employee has no pairs and no window counts
        const u64* q = s.eq + (int)s.pslot[en] * 4;
        dh += __popcll(mn & s.part[d]) + __popcll(q[0] & s.cont14[d]);
        ds += __popcll(q[2] & s.cont7[d]);
        if (wd < 5) cn = __popcll(mn & K.wd[wd]);
    }
    // S2 / S3+S4: memoised per (day, weekday count) and per (day, (total, weekend) class)
    ds += (int)s.s2t[d * ES_CBINS + cn] + (int)s.s34[d * ES_MAXCLS + (int)s.cls[en]];
}

// swap: days d1 < d2 exchange employees (different).
__device__ __forceinline__ void es_swap_delta(const EsSmem& s, const EsConst& K,
                                              const u64* __restrict__ hol, int d1, int d2, int& dh,
                                              int& ds) {
    const int e1 = s.a[d1], e2 = s.a[d2];
    const u64 b1 = 1ull << d1, b2 = 1ull << d2;
    const u64 m1 = s.mask[e1], m2 = s.mask[e2];
    const u64* q1 = s.eq + (int)s.pslot[e1] * 4;
    const u64* q2 = s.eq + (int)s.pslot[e2] * 4;
    const u64 h1 = hol[e1], h2 = hol[e2];
    // windows holding exactly one of the two days change count by one for each employee
    const u64 c14a = s.cont14[d1], c14b = s.cont14[d2], c7a = s.cont7[d1], c7b = s.cont7[d2];
    const u64 only14a = c14a & ~c14b, only14b = c14b & ~c14a, only7a = c7a & ~c7b, only7b = c7b & ~c7a;
    dh = (int)((h1 >> d2) & 1ull) - (int)((h1 >> d1) & 1ull) + (int)((h2 >> d1) & 1ull) -
         (int)((h2 >> d2) & 1ull);
    // H2/H3 pairs: e1 leaves d1 and lands on d2 (its other days: m1 without d1), e2 the reverse
    dh += __popcll((m1 & ~b1) & s.part[d2]) - __popcll(m1 & s.part[d1]);
    dh += __popcll((m2 & ~b2) & s.part[d1]) - __popcll(m2 & s.part[d2]);
    // H4: e1 loses a day in windows with only d1 (count 4 -> 3), gains in windows with only d2
    dh += __popcll(q1[0] & only14b) - __popcll(q1[1] & only14a);
    dh += __popcll(q2[0] & only14a) - __popcll(q2[1] & only14b);
    ds = __popcll(q1[2] & only7b) - __popcll(q1[3] & only7a);
    ds += __popcll(q2[2] & only7a) - __popcll(q2[3] & only7b);
    const int wd1 = es_weekday(K, d1), wd2 = es_weekday(K, d2);
    if (wd1 != wd2) {
        // the two weekdays are distinct histograms, so their deltas are independent
        if (wd1 < 5)
            ds += es_s2_delta(s, wd1, __popcll(m1 & K.wd[wd1]), __popcll(m2 & K.wd[wd1]));
        if (wd2 < 5)
            ds += es_s2_delta(s, wd2, __popcll(m2 & K.wd[wd2]), __popcll(m1 & K.wd[wd2]));
    }
    const int k1 = (K.wkend & b1) ? 1 : 0, k2 = (K.wkend & b2) ? 1 : 0;
    if (k1 != k2) {  // totals (S3) unchanged; weekend counts move between the two employees
        const int present = s.misc[ES_PRESENT];
        const int w1 = __popcll(m1 & K.wkend), w2 = __popcll(m2 & K.wkend);
        EsAdj W;
        W.add(w1, -1);
        W.add(w1 - k1 + k2, +1);
        W.add(w2, -1);
        W.add(w2 + k1 - k2, +1);
        const u64 occW = es_occ_after(s.histW, (u64)*s.occW, W);
        ds += es_spread(occW, present) - es_spread((u64)*s.occW, present);
    }
}

// ------------------------------------------------------------------ tallies (K4)
// Build masks + histograms from a[] and the full score from them.  Block-cooperative.
__device__ void es_build(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol,
                         int& hard, int& soft) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < K.E; e += nt) s.mask[e] = 0;
    for (int k = tid; k < 5 * ES_CBINS + ES_TBINS + ES_WBINS; k += nt) s.hist2[k] = 0;
    if (tid < 5) s.occ2[tid] = 0;
    if (tid < 16) s.misc[tid] = 0;
    if (tid == 0) {
        *s.occT = 0;
        *s.occW = 0;
    }
    __syncthreads();
    for (int d = tid; d < K.D; d += nt) atomicOr(&s.mask[s.a[d]], 1ull << d);
    __syncthreads();
    int h = 0, s1 = 0;
    for (int e = tid; e < K.E; e += nt) {
        const u64 m = s.mask[e];
        if (!m) continue;
        int eh, es;
        es_emp_full(m, hol[e], K, eh, es);
        h += eh;
        s1 += es;
        atomicAdd(&s.misc[ES_PRESENT], 1);
        const int t = __popcll(m), w = __popcll(m & K.wkend);
        // 16-bit histogram bins: add through the containing 32-bit word
        {
            const int idx = (int)(s.histT - s.hist2) + t;
            atomicAdd((unsigned int*)s.hist2 + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
            const int idw = (int)(s.histW - s.hist2) + w;
            atomicAdd((unsigned int*)s.hist2 + (idw >> 1), (idw & 1) ? 0x10000u : 1u);
        }
        atomicOr(s.occT, 1ull << t);
        atomicOr(s.occW, 1u << w);
        for (int wd = 0; wd < 5; ++wd) {
            const int c = __popcll(m & K.wd[wd]);
            if (!c) continue;
            const int idx = wd * ES_CBINS + c;
            atomicAdd((unsigned int*)s.hist2 + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
            atomicOr(&s.occ2[wd], 1u << c);
            atomicAdd(&s.misc[ES_DISTINCT0 + wd], 1);
        }
    }
    atomicAdd(&s.misc[ES_HARD], h);
    atomicAdd(&s.misc[ES_SOFT], s1);
    __syncthreads();
    hard = s.misc[ES_HARD];
    soft = s.misc[ES_SOFT];
    const int present = s.misc[ES_PRESENT];
    for (int wd = 0; wd < 5; ++wd) soft += es_s2_term(s.occ2[wd], s.misc[ES_DISTINCT0 + wd]);
    soft += es_spread(*s.occT, present) + es_spread((u64)*s.occW, present);
    __syncthreads();
}

// single-thread application of an accepted move to the tallies
__device__ void es_hist_move(uint16_t* hist, unsigned int* occ32, u64* occ64, int from, int to) {
    // from/to < 0 : no such side
    if (from >= 0) {
        if (--hist[from] == 0) {
            if (occ64) *occ64 &= ~(1ull << from);
            else *occ32 &= ~(1u << from);
        }
    }
    if (to >= 0) {
        if (hist[to]++ == 0) {
            if (occ64) *occ64 |= 1ull << to;
            else *occ32 |= 1u << to;
        }
    }
}

__device__ void es_set_mask(const EsSmem& s, const EsConst& K, int e, u64 m2) {
    const u64 m = s.mask[e];
    const int t = __popcll(m), t2 = __popcll(m2);
    const int w = __popcll(m & K.wkend), w2 = __popcll(m2 & K.wkend);
    if (t != t2 || w != w2) {
        es_hist_move(s.histT, nullptr, s.occT, t >= 1 ? t : -1, t2 >= 1 ? t2 : -1);
        es_hist_move(s.histW, s.occW, nullptr, t >= 1 ? w : -1, t2 >= 1 ? w2 : -1);
        s.misc[ES_PRESENT] += (t2 >= 1) - (t >= 1);
    }
    for (int wd = 0; wd < 5; ++wd) {
        const int c = __popcll(m & K.wd[wd]), c2 = __popcll(m2 & K.wd[wd]);
        if (c == c2) continue;
        es_hist_move(s.hist2 + wd * ES_CBINS, &s.occ2[wd], nullptr, c >= 1 ? c : -1,
                     c2 >= 1 ? c2 : -1);
        s.misc[ES_DISTINCT0 + wd] += (c2 >= 1) - (c >= 1);
    }
    s.mask[e] = m2;
}

// move id: change (d, e) -> d*E + e ; swap (d1<d2) -> D*E + tri(d1,d2)
__device__ __forceinline__ int es_tri_index(int D, int d1, int d2) {
    return d1 * D - d1 * (d1 + 1) / 2 + (d2 - d1 - 1);
}

__device__ __forceinline__ long long es_key(int dh, int ds, int id) {
    return ((long long)(dh + 32768) << 44) | ((long long)(ds + 32768) << 24) | (long long)id;
}

__device__ __forceinline__ long long es_block_min(long long key, u64* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = (u64)key;
    __syncthreads();
    if (w == 0) {
        long long x = (l < nw) ? (long long)red[l] : ES_KEY_INF;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long other = __shfl_xor_sync(0xffffffffu, x, o);
            x = other < x ? other : x;
        }
        if (l == 0) red[32] = (u64)x;
    }
    __syncthreads();
    const long long out = (long long)red[32];
    __syncthreads();
    return out;
}

// ------------------------------------------------------------------ the step kernel (K5)
__global__ void es_step_kernel(EsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsConst& K = p.K;
    const EsSmem s = es_carve(smem_raw, K.D, K.E);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int D = K.D, E = K.E;
    const int n_change = D * E, n_swap = D * (D - 1) / 2, n_moves = n_change + n_swap;
    for (int k = tid; k < 192; k += nt) s.part[k] = p.dayconst[k];  // part | cont14 | cont7

    for (;;) {
        __syncthreads();
        if (tid == 0) s.misc[ES_BCAST] = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const int local = s.misc[ES_BCAST];
        if (local >= p.n_chains) break;
        const int chain = p.first_chain + local;
        if (p.skip && p.skip[chain]) continue;
        uint16_t* ga = p.a + (size_t)chain * p.stride;
        EsChainState* st = p.st + chain;
        __syncthreads();
        for (int k = tid; k < p.stride; k += nt) s.a[k] = ga[k];
        __syncthreads();
        int hard, soft;
        es_build(s, K, p.hol, hard, soft);
        int best_h = p.ls_mode ? hard : (int)st->best_hard;
        int best_s = p.ls_mode ? soft : (int)st->best_soft;
        if (p.ls_mode)
            for (int k = tid; k < p.stride; k += nt) p.best_a[(size_t)chain * p.stride + k] = s.a[k];
        unsigned long long no_improve = 0, scored = 0;
        const unsigned int steps0 = st->steps;
        unsigned int steps = steps0, status = 0;

        for (unsigned long long it = 0; it < p.max_steps; ++it) {
            if (hard == 0 && soft == 0 && !p.dump_h) {  // is_best, lib.rs:245-249
                status = 1;
                best_h = 0;
                best_s = 0;
                break;
            }
            es_prepare(s, K, p.hol);
            long long key = ES_KEY_INF;
            unsigned int nscored = 0;
            for (int id = tid; id < n_moves; id += nt) {
                int dh, ds;
                bool ok;
                if (id < n_change) {
                    const int d = id / E, e = id - d * E;
                    ok = (s.a[d] != e);
                    if (ok) es_change_delta(s, K, p.hol, d, e, dh, ds);
                } else {
                    // decode the triangular index (d1 < d2)
                    int r = id - n_change, d1 = 0;
                    while (r >= D - 1 - d1) {
                        r -= D - 1 - d1;
                        ++d1;
                    }
                    const int d2 = d1 + 1 + r;
                    ok = (s.a[d1] != s.a[d2]);
                    if (ok) es_swap_delta(s, K, p.hol, d1, d2, dh, ds);
                }
                if (p.dump_h) {
                    p.dump_h[id] = ok ? (long long)dh : INT64_MAX;
                    p.dump_s[id] = ok ? (long long)ds : INT64_MAX;
                }
                if (ok) {
                    ++nscored;
                    const long long k2 = es_key(dh, ds, id);
                    key = k2 < key ? k2 : key;
                }
            }
            key = es_block_min(key, s.red);
            {   // count the candidates scored (block sum through the same scratch)
                unsigned int c = nscored;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                if ((tid & 31) == 0) atomicAdd((unsigned int*)&s.misc[ES_BCAST + 1], c);
                __syncthreads();
                scored += (unsigned int)s.misc[ES_BCAST + 1];
                __syncthreads();
                if (tid == 0) s.misc[ES_BCAST + 1] = 0;
            }
            if (p.dump_h) break;
            if (key == ES_KEY_INF) {  // empty neighbourhood, local_search.rs:336-338
                status = 3;
                break;
            }
            const int dh = (int)((key >> 44) & 0xffff) - 32768;
            const int ds = (int)((key >> 24) & 0xfffff) - 32768;
            const int id = (int)(key & 0xffffff);
            const bool improved = dh < 0 || (dh == 0 && ds < 0);  // lexicographic (hard, soft)
            if (!improved) {
                ++no_improve;
                if (p.allow_no_improve && no_improve >= p.allow_no_improve) {
                    status = 2;
                    break;
                }
            } else {
                no_improve = 0;
            }
            hard += dh;
            soft += ds;
            if (tid == 0) {
                unsigned int kind, x, y;
                if (id < n_change) {
                    const int d = id / E, e = id - d * E, eo = s.a[d];
                    const u64 bit = 1ull << d;
                    es_set_mask(s, K, eo, s.mask[eo] & ~bit);
                    es_set_mask(s, K, e, s.mask[e] | bit);
                    s.a[d] = (uint16_t)e;
                    kind = 0;
                    x = (unsigned)d;
                    y = (unsigned)e;
                } else {
                    int r = id - n_change, d1 = 0;
                    while (r >= D - 1 - d1) {
                        r -= D - 1 - d1;
                        ++d1;
                    }
                    const int d2 = d1 + 1 + r, e1 = s.a[d1], e2 = s.a[d2];
                    const u64 x2 = (1ull << d1) | (1ull << d2);
                    es_set_mask(s, K, e1, s.mask[e1] ^ x2);
                    es_set_mask(s, K, e2, s.mask[e2] ^ x2);
                    s.a[d1] = (uint16_t)e2;
                    s.a[d2] = (uint16_t)e1;
                    kind = 1;
                    x = (unsigned)d1;
                    y = (unsigned)d2;
                }
                if (p.trace && steps < (unsigned)p.trace_cap) {
                    EsTraceEntry t;
                    t.kind = kind;
                    t.x = x;
                    t.y = y;
                    t.pad = 0;
                    t.hard_after = hard;
                    t.soft_after = soft;
                    p.trace[(size_t)chain * p.trace_cap + steps] = t;
                }
            }
            ++steps;
            __syncthreads();
            if (improved) {
                best_h = hard;
                best_s = soft;
                for (int k = tid; k < p.stride; k += nt)
                    p.best_a[(size_t)chain * p.stride + k] = s.a[k];
            }
        }
        __syncthreads();
        if (p.dump_h) continue;
        for (int k = tid; k < p.stride; k += nt) ga[k] = s.a[k];
        if (tid == 0) {
            st->hard = hard;
            st->soft = soft;
            st->best_hard = best_h;
            st->best_soft = best_s;
            st->moves_scored += scored;
            st->steps = steps;
            st->status = status;
            atomicAdd(p.totals, scored);
            atomicAdd(p.totals + 1, (unsigned long long)(steps - steps0));
        }
    }
}

// (re)score chains from the tallies after set/init: fills hard/soft/best and best_a
__global__ void es_rescore_kernel(EsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsSmem s = es_carve(smem_raw, p.K.D, p.K.E);
    for (int local = blockIdx.x; local < p.n_chains; local += gridDim.x) {
        const int chain = p.first_chain + local;
        __syncthreads();
        for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
            s.a[k] = p.a[(size_t)chain * p.stride + k];
        __syncthreads();
        int hard, soft;
        es_build(s, p.K, p.hol, hard, soft);
        for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
            p.best_a[(size_t)chain * p.stride + k] = s.a[k];
        if (threadIdx.x == 0) {
            EsChainState z = p.st[chain];
            z.hard = hard;
            z.soft = soft;
            z.best_hard = hard;
            z.best_soft = soft;
            z.status = (hard == 0 && soft == 0) ? 1u : 0u;
            p.st[chain] = z;
        }
    }
}

// explicit-move deltas against one chain (parity hook); kind 0 change (x=day,y=employee idx),
// 1 swap (x,y = days)
__global__ void es_eval_kernel(EsParams p, int chain, int kind, const uint2* __restrict__ moves,
                               unsigned long long n_moves, long long* __restrict__ dh_out,
                               long long* __restrict__ ds_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsSmem s = es_carve(smem_raw, p.K.D, p.K.E);
    for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
        s.a[k] = p.a[(size_t)chain * p.stride + k];
    for (int k = threadIdx.x; k < 192; k += blockDim.x) s.part[k] = p.dayconst[k];
    __syncthreads();
    int hard, soft;
    es_build(s, p.K, p.hol, hard, soft);
    es_prepare(s, p.K, p.hol);
    for (unsigned long long k = threadIdx.x; k < n_moves; k += blockDim.x) {
        const uint2 mv = moves[k];
        int dh = 0, ds = 0;
        bool ok;
        if (kind == 0) {
            ok = s.a[mv.x] != mv.y;
            if (ok) es_change_delta(s, p.K, p.hol, (int)mv.x, (int)mv.y, dh, ds);
        } else {
            const int d1 = (int)min(mv.x, mv.y), d2 = (int)max(mv.x, mv.y);
            ok = d1 != d2 && s.a[d1] != s.a[d2];
            if (ok) es_swap_delta(s, p.K, p.hol, d1, d2, dh, ds);
        }
        dh_out[k] = ok ? (long long)dh : INT64_MAX;
        ds_out[k] = ok ? (long long)ds : INT64_MAX;
    }
}

// K6: full re-score straight from a[] by the reference's own loops (no masks, no
// histograms) -- one thread per chain; cross-checks the tally formulation.
__global__ void es_full_score_kernel(const uint16_t* __restrict__ a, int stride, int n_chains,
                                     EsConst K, const u64* __restrict__ hol, long long* out8) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const uint16_t* x = a + (size_t)chain * stride;
    const int D = K.D;
    long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int d = 0; d < D; ++d) t[0] += (hol[x[d]] >> d) & 1ull;           // lib.rs:273-280
    for (int i = 0; i + 2 <= D; ++i) t[1] += (x[i] == x[i + 1]);           // :286-292
    for (int i = 0; i + 9 <= D; ++i) {                                     // :295-315
        if ((K.start_wd + i) % 7 != 5) continue;
        t[2] += (x[i] == x[i + 7]) + (x[i] == x[i + 8]) + (x[i + 1] == x[i + 7]) +
                (x[i + 1] == x[i + 8]);
    }
    for (int len = 14; len >= 7; len -= 7) {                               // :318-339
        const int limit = len == 14 ? 3 : 2;
        for (int w = 0; w + len <= D; ++w)
            for (int q = w; q < w + len; ++q) {
                bool first = true;
                for (int r = w; r < q; ++r)
                    if (x[r] == x[q]) first = false;
                if (!first) continue;
                int c = 0;
                for (int r = q; r < w + len; ++r) c += (x[r] == x[q]);
                if (c > limit) t[len == 14 ? 3 : 4] += 1;
            }
    }
    for (int wd = 0; wd < 5; ++wd) {                                       // :194-218
        int distinct = 0, minc = 1 << 30;
        for (int i = 0; i < D; ++i) {
            if ((K.start_wd + i) % 7 != wd) continue;
            bool first = true;
            for (int q = 0; q < i; ++q)
                if ((K.start_wd + q) % 7 == wd && x[q] == x[i]) first = false;
            if (!first) continue;
            int c = 0;
            for (int q = i; q < D; ++q) c += ((K.start_wd + q) % 7 == wd && x[q] == x[i]);
            ++distinct;
            minc = c < minc ? c : minc;
        }
        if (distinct >= 2) t[5] += minc;
    }
    int present = 0, mind = 1 << 30, maxd = -1, minw = 1 << 30, maxw = -1;  // :345-365
    for (int i = 0; i < D; ++i) {
        bool first = true;
        for (int q = 0; q < i; ++q)
            if (x[q] == x[i]) first = false;
        if (!first) continue;
        int days = 0, wk = 0;
        for (int q = i; q < D; ++q)
            if (x[q] == x[i]) {
                ++days;
                const int w = (K.start_wd + q) % 7;
                wk += (w == 5 || w == 6);
            }
        ++present;
        mind = days < mind ? days : mind;
        maxd = days > maxd ? days : maxd;
        minw = wk < minw ? wk : minw;
        maxw = wk > maxw ? wk : maxw;
    }
    if (present >= 2) {
        t[6] = maxd - mind;
        t[7] = maxw - minw;
    }
    for (int k = 0; k < 8; ++k) out8[(size_t)chain * 8 + k] = t[k];
}

// initial solution: uniform random employee per slot, phantom slot included
// (examples/employee-scheduling/src/lib.rs:404-419); draw s of stream (seed, chain, INIT).
__global__ void es_init_kernel(uint16_t* a, EsChainState* st, int stride, int E, int n_chains,
                               unsigned long long seed, unsigned int chain_offset) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    PhiloxDraws d(seed, chain_offset + (unsigned)chain, 0u);
    for (int k = 0; k < stride; ++k) a[(size_t)chain * stride + k] = (uint16_t)d.below((unsigned)E);
    EsChainState z;
    z.hard = z.soft = z.best_hard = z.best_soft = -1;
    z.moves_scored = 0;
    z.steps = 0;
    z.status = 0;
    st[chain] = z;
}

__global__ void es_reset_state_kernel(EsChainState* st, int first, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    EsChainState z;
    z.hard = z.soft = z.best_hard = z.best_soft = -1;
    z.moves_scored = 0;
    z.steps = 0;
    z.status = 0;
    st[first + k] = z;
}

struct EsStats {
    long long best_hard, best_soft, best_key;
    unsigned int best_chain, chains_at_best, chains_feasible, pad;
};

// best chain by lexicographic (hard, soft); key = (hard<<44 | soft<<24... ) packs into
// (hard << 48) | (soft << 32) | global chain id for the min-allreduce
__global__ void es_stats_kernel(const EsChainState* __restrict__ st, int n_chains,
                                unsigned int chain_offset, EsStats* out) {
    __shared__ long long skey[32];
    __shared__ unsigned int sa[32], sf[32];
    long long key = ES_KEY_INF;
    unsigned int ab = 0, fe = 0;
    for (int c = threadIdx.x; c < n_chains; c += blockDim.x) {
        const EsChainState x = st[c];
        const long long k = (x.hard << 48) | (x.soft << 32) | (long long)(unsigned)c;
        key = k < key ? k : key;
        ab += (x.hard == 0 && x.soft == 0);
        fe += (x.hard == 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
        key = ok < key ? ok : key;
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
        fe += __shfl_xor_sync(0xffffffffu, fe, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        skey[w] = key;
        sa[w] = ab;
        sf[w] = fe;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            key = skey[k] < key ? skey[k] : key;
            ab += sa[k];
            fe += sf[k];
        }
        out->best_hard = key >> 48;
        out->best_soft = (key >> 32) & 0xffff;
        out->best_chain = (unsigned)(key & 0xffffffffll);
        out->best_key = (key & ~0xffffffffll) | (long long)(out->best_chain + chain_offset);
        out->chains_at_best = ab;
        out->chains_feasible = fe;
    }
}

}  // namespace csb
