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
// mask changes.
// Hot-loop formulation (es_tally -> es_prepare -> es_scan).  At most D
// employees are PRESENT (hold a day); each gets a slot = rank of its first day, so every
// per-step table is built by <= D threads and nothing loops over the employee table.  Per slot
// all sliding-window counts are computed at once by a bit-sliced adder over the day mask and
// kept as "count == k" window-start masks (EQ3_14, EQ4_14, EQ2_7, EQ3_7); per day d the terms
// that only depend on the day's current employee are tabulated (base[d]).  A change candidate to
// a present employee then needs three 64-bit popcounts for H2+H3, H4 and S1:
//   gain of day d for slot s = popc(m_s & PART[d]) + popc(EQ3_14[s] & CONT14[d]) ...
// where PART[d] = days paired with d by H2/H3 and CONT[d] = window starts whose window holds d;
// a candidate to an ABSENT employee is baseW[d] plus its holiday bit; a swap is two such
// transfers (pass A's gain table) corrected where both days meet.
// S2 (weekday affinity, min over employees present on that weekday), S3 (max-min of total
// days over PRESENT employees) and S4 (max-min of weekend days over present employees) are
// kept as count histograms + occupancy bitsets; their deltas depend on the receiver only through
// a count in use, so they are memoised per (day, value in use) as closed-form occupancy moves.
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
    // reference mode: the reference's own random proposer + window (es_scan_ref)
    unsigned long long seed, window, max_draws;
    unsigned int chain_offset;
};

// per-chain shared state.  NS = min(E, D) bounds the number of PRESENT employees ("slots");
// DP = D rounded up to 4 sizes the per-day arrays.
constexpr int ES_TCOLS = 12;  // distinct total-day values among present employees (sum <= 64 => <= 10)
constexpr int ES_WCOLS = 8;   // distinct weekend-day values (0 included; sum <= 20 => <= 6)
// layout of EsSmem::val: total-day values by rank | weekend values by rank | per weekday: counts in use | their number
constexpr int ES_VAL_C2 = ES_TCOLS + ES_WCOLS, ES_VAL_N2 = ES_VAL_C2 + 5 * 12, ES_VAL_BYTES = ES_VAL_N2 + 8;
struct EsSmem {
    u64* mask;          // [E]
    uint16_t* a;        // [stride]
    uint16_t* hist2;    // [5][ES_CBINS]
    uint16_t* histT;    // [ES_TBINS]
    uint16_t* histW;    // [ES_WBINS]
    unsigned int* occ2; // [5]  bit c: some employee has exactly c days on that weekday (c>=1)
    u64* occT;          // [1]  bit c: some employee has exactly c days in total (c>=1)
    unsigned int* occW; // [1]  bit c: some PRESENT employee has exactly c weekend days (c>=0)
    u64* fmask;         // [1]  bit d: day d is the FIRST day of its employee (one bit per present employee)
    int* misc;          // [16] present, distinct[5], hard, soft, ...
    u64* red;           // [40] reduction scratch
    u64* part;          // [DP] H2/H3 partner-day mask per day
    u64* cont14;        // [DP] 14-day window starts whose window contains the day
    u64* cont7;         // [DP]
    u64* wdm;           // [DP] days on the same weekday as d (0 for weekend days)
    u64* eq;            // [NS][4] per slot: EQ3_14, EQ4_14, EQ2_7, EQ3_7 ("count == k" window starts)
    u64* smask;         // [NS] day mask of the slot's employee
    u64* shol;          // [NS] its holiday mask
    uint16_t* semp;     // [NS] its employee index
    unsigned char* srk;    // [NS][2] rank of its total-day / weekend-day count among the values in use
    unsigned char* val;    // [ES_VAL_BYTES] rank -> value lists (see ES_VAL_*)
    unsigned char* dwd;    // [DP] weekday of the day (0 = Monday)
    unsigned char* dslot;  // [DP] slot of the day's current employee
    unsigned char* dayb;   // [3][DP] per day, for its current employee: total, weekend, weekday count
    unsigned int* base;    // [DP] packed (0x8000 - lossH) << 16 | (0x8000 - lossS1) of the day's employee
    unsigned int* baseW;   // [DP] absent receiver without a holiday: (dh + 64) << 15 | (ds + 256) << 6 | d
    signed char* s2t;      // [DP][ES_CBINS] S2 delta of giving day d to an employee with cn days on that weekday
    signed char* s3t;      // [DP][ES_TCOLS] S3 delta ... to a present employee whose total has rank j
    signed char* s4t;      // [DP][ES_WCOLS] S4 delta ... to a present employee whose weekend count has rank j
    signed char* s4s;      // [DP][ES_WCOLS] S4 delta of SWAPPING weekend day d with a weekday of such an employee
    uint16_t* ga;          // [DP][NS] pass A's gains of giving day d to slot: gh | gs << 5 | (s2 + 16) << 8 (swaps reuse them)
    int ns, dp;
};

struct EsLayout {
    size_t mask, a, hist, occ, occT, fmask, misc, red, day, eq, smask, shol, semp, srk, val, dwd, dslot, dayb, base,
        baseW, s2t, s3t, s4t, s4s, ga, total;
    int ns, dp;
};
__host__ __device__ inline size_t es_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline EsLayout es_layout(int D, int E) {
    EsLayout L;
    L.ns = E < D ? E : D;
    if (L.ns < 1) L.ns = 1;
    L.dp = (D + 3) & ~3;
    const size_t dp = (size_t)L.dp;
    size_t o = 0;
    L.mask = o;    o += (size_t)E * 8;
    L.occT = o;    o += 8;
    L.fmask = o;   o += 8;
    L.red = o;     o += 40 * 8;
    L.day = o;     o += 4 * dp * 8;
    L.eq = o;      o += (size_t)L.ns * 4 * 8;
    L.smask = o;   o += (size_t)L.ns * 8;
    L.shol = o;    o += (size_t)L.ns * 8;
    o = es_align(o, 16);
    L.baseW = o;   o += dp * 4;  // read as uint4
    L.base = o;    o += dp * 4;
    L.misc = o;    o += 16 * 4;
    L.occ = o;     o = es_align(o + 6 * 4, 8);
    L.hist = o;    o = es_align(o + (5 * ES_CBINS + ES_TBINS + ES_WBINS) * 2, 8);
    L.a = o;       o = es_align(o + (size_t)(D + 1) * 2, 8);
    L.semp = o;    o = es_align(o + (size_t)L.ns * 2, 8);
    L.srk = o;     o = es_align(o + (size_t)L.ns * 2, 8);
    L.val = o;     o = es_align(o + ES_VAL_BYTES, 8);
    L.dwd = o;     o += dp;
    L.dslot = o;   o += dp;
    L.dayb = o;    o += 3 * dp;
    L.s2t = o;     o += dp * ES_CBINS;
    L.s3t = o;     o += dp * ES_TCOLS;
    L.s4t = o;     o += dp * ES_WCOLS;
    L.s4s = o;     o += dp * ES_WCOLS;
    o = es_align(o, 8);
    L.ga = o;      o += dp * (size_t)L.ns * 2;
    L.total = es_align(o, 16);
    return L;
}
__host__ __device__ inline size_t es_smem_bytes(int D, int E) { return es_layout(D, E).total; }

// offsets only (no pointer <-> integer casts) so the compiler keeps the shared address space
__device__ __forceinline__ EsSmem es_carve(unsigned char* p, int D, int E) {
    const EsLayout L = es_layout(D, E);
    EsSmem s;
    s.ns = L.ns;
    s.dp = L.dp;
    s.mask = (u64*)(p + L.mask);
    s.a = (uint16_t*)(p + L.a);
    s.hist2 = (uint16_t*)(p + L.hist);
    s.histT = s.hist2 + 5 * ES_CBINS;
    s.histW = s.histT + ES_TBINS;
    s.occ2 = (unsigned int*)(p + L.occ);
    s.occW = s.occ2 + 5;
    s.occT = (u64*)(p + L.occT);
    s.fmask = (u64*)(p + L.fmask);
    s.misc = (int*)(p + L.misc);
    s.red = (u64*)(p + L.red);
    s.part = (u64*)(p + L.day);
    s.cont14 = s.part + L.dp;
    s.cont7 = s.cont14 + L.dp;
    s.wdm = s.cont7 + L.dp;
    s.eq = (u64*)(p + L.eq);
    s.smask = (u64*)(p + L.smask);
    s.shol = (u64*)(p + L.shol);
    s.semp = (uint16_t*)(p + L.semp);
    s.srk = p + L.srk;
    s.val = p + L.val;
    s.dwd = p + L.dwd;
    s.dslot = p + L.dslot;
    s.dayb = p + L.dayb;
    s.base = (unsigned int*)(p + L.base);
    s.baseW = (unsigned int*)(p + L.baseW);
    s.s2t = (signed char*)(p + L.s2t);
    s.s3t = (signed char*)(p + L.s3t);
    s.s4t = (signed char*)(p + L.s4t);
    s.s4s = (signed char*)(p + L.s4s);
    s.ga = (uint16_t*)(p + L.ga);
    return s;
}

enum { ES_PRESENT = 0, ES_DISTINCT0 = 1, ES_HARD = 6, ES_SOFT = 7, ES_BCAST = 8, ES_NSLOT = 10, ES_SAME = 11 };
constexpr unsigned int ES_W_PAD = 0x7fff0000u;  // larger than any packed absent-candidate value, no overflow on +hol

// ------------------------------------------------------------------ histogram helpers
// bit b of a 64-bit set; bin 64 (one employee holding all 64 days) has no bit -- it can only occur
// with a single present employee, where the spread is 0 whatever the set says
__device__ __forceinline__ u64 es_bit64(int b) { return b < 64 ? 1ull << b : 0ull; }

// Occupancy bitset after one member leaves bin r0, one leaves r1 (r1 < 0: nobody), one enters
// a0 (a0 < 0: nobody) and one enters a1.  hist holds the current member count per bin.
__device__ __forceinline__ u64 es_occ_move(const uint16_t* hist, u64 occ, int r0, int r1, int a0, int a1) {
    int c0 = (int)hist[r0] - 1;
    if (r1 >= 0) {
        const int same = (r1 == r0) ? 1 : 0;
        c0 -= same;
        if ((int)hist[r1] - 1 - same <= 0) occ &= ~es_bit64(r1);
    }
    if (c0 <= 0) occ &= ~es_bit64(r0);
    if (a0 >= 0) occ |= es_bit64(a0);
    return occ | es_bit64(a1);
}
__device__ __forceinline__ unsigned int es_occ_move32(const uint16_t* hist, unsigned int occ, int r0, int r1, int a0,
                                                      int a1) {
    int c0 = (int)hist[r0] - 1;
    if (r1 >= 0) {
        const int same = (r1 == r0) ? 1 : 0;
        c0 -= same;
        if ((int)hist[r1] - 1 - same <= 0) occ &= ~(1u << r1);
    }
    if (c0 <= 0) occ &= ~(1u << r0);
    if (a0 >= 0) occ |= 1u << a0;
    return occ | (1u << a1);
}

__device__ __forceinline__ int es_spread(u64 occ, int members) {  // max-min, lib.rs:349,363
    if (members < 2 || occ == 0) return 0;
    return (63 - __clzll((long long)occ)) - (__ffsll((long long)occ) - 1);
}
__device__ __forceinline__ int es_spread32(unsigned int occ, int members) {
    if (members < 2 || occ == 0) return 0;
    return (31 - __clz((int)occ)) - (__ffs((int)occ) - 1);
}

__device__ __forceinline__ int es_s2_term(unsigned int occ, int distinct) {  // lib.rs:206-214
    return (distinct >= 2 && occ) ? (__ffs((int)occ) - 1) : 0;
}

// weekday-affinity delta on weekday wd when one employee's count there goes cm -> cm-1
// (cm >= 1) and another's goes cp -> cp+1 (cp >= 0)
__device__ __forceinline__ int es_s2_delta(const EsSmem& s, int wd, int cm, int cp) {
    const int distinct = s.misc[ES_DISTINCT0 + wd];
    const unsigned int occ0 = s.occ2[wd];
    const unsigned int occ = es_occ_move32(s.hist2 + wd * ES_CBINS, occ0, cm, cp >= 1 ? cp : -1,
                                           cm - 1 >= 1 ? cm - 1 : -1, cp + 1);
    return es_s2_term(occ, distinct - (cm == 1) + (cp == 0)) - es_s2_term(occ0, distinct);
}

// S3 delta (max-min of total days over PRESENT employees, lib.rs:345-351) when a day moves from
// an employee with `to` days to one with `tn` days (0 = absent so far)
__device__ __forceinline__ int es_s3_delta(const EsSmem& s, int to, int tn) {
    const int present = s.misc[ES_PRESENT];
    const u64 occ = es_occ_move(s.histT, *s.occT, to, tn >= 1 ? tn : -1, to - 1 >= 1 ? to - 1 : -1, tn + 1);
    return es_spread(occ, present - (to == 1) + (tn == 0)) - es_spread(*s.occT, present);
}
// S4 delta (weekend days, lib.rs:354-365): the day (weekend flag isw) leaves an employee with
// (to total, wo weekend) days for one with wn weekend days (absent = !rpresent, wn = 0)
__device__ __forceinline__ int es_s4_delta(const EsSmem& s, int to, int wo, int isw, int wn, bool rpresent) {
    const int present = s.misc[ES_PRESENT];
    const unsigned int occ = es_occ_move32(s.histW, *s.occW, wo, rpresent ? wn : -1, to - 1 >= 1 ? wo - isw : -1,
                                           wn + isw);
    return es_spread32(occ, present - (to == 1) + (rpresent ? 0 : 1)) - es_spread32(*s.occW, present);
}

// ------------------------------------------------------------------ tallies and per-step tables
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

__device__ __forceinline__ int es_pair_terms(u64 m, u64 hol, const EsConst& K) {
    const u64 m1 = m >> 1, m7 = m >> 7, m8 = m >> 8;
    return __popcll(m & hol) + __popcll(m & m1) + __popcll(m & m7 & K.satf) +
           __popcll(m & m8 & K.satf) + __popcll(m1 & m7 & K.satf) + __popcll(m1 & m8 & K.satf);
}

__device__ __forceinline__ void es_hist16_inc(uint16_t* hist2, int idx) {  // 16-bit bin through its 32-bit word
    atomicAdd((unsigned int*)hist2 + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
}

// slot of a present employee = rank of its first day among the first days
__device__ __forceinline__ int es_slot_of(const EsSmem& s, u64 m) {
    const int f = __ffsll((long long)m) - 1;
    return __popcll(*s.fmask & ((1ull << f) - 1ull));
}

// Tallies from the day masks (must be current): count histograms, occupancy sets, present /
// distinct counters, the first-day mask.  Work is per DAY (<= 64 threads busy), never per
// employee.  SCORE additionally accumulates the full (hard, S1) into misc[ES_HARD/ES_SOFT].
template <bool SCORE>
__device__ void es_tally(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = tid; k < (5 * ES_CBINS + ES_TBINS + ES_WBINS) / 2; k += nt) ((unsigned int*)s.hist2)[k] = 0;
    if (tid < 6) s.occ2[tid] = 0;  // occ2[0..4] and occW
    if (tid < 16 && tid != ES_BCAST && tid != ES_BCAST + 1) s.misc[tid] = 0;
    if (tid == 0) {
        *s.occT = 0;
        *s.fmask = 0;
    }
    __syncthreads();
    const u64 v14 = K.n14 >= 64 ? ~0ull : ((1ull << K.n14) - 1);  // real window starts only
    const u64 v7 = K.n7 >= 64 ? ~0ull : ((1ull << K.n7) - 1);
    for (int d = tid; d < K.D; d += nt) {
        const int e = s.a[d];
        const u64 m = s.mask[e];
        if (m & ((1ull << d) - 1ull)) continue;  // not the employee's first day
        atomicOr(s.fmask, 1ull << d);
        const int t = __popcll(m), w = __popcll(m & K.wkend);
        es_hist16_inc(s.hist2, (int)(s.histT - s.hist2) + t);
        es_hist16_inc(s.hist2, (int)(s.histW - s.hist2) + w);
        atomicOr(s.occT, es_bit64(t));
        atomicOr(s.occW, 1u << w);
        atomicAdd(&s.misc[ES_PRESENT], 1);
        atomicAdd(&s.misc[ES_SAME], t * (t - 1) / 2);  // day pairs held by one employee (identity swaps)
#pragma unroll
        for (int wd = 0; wd < 5; ++wd) {
            const int c = __popcll(m & K.wd[wd]);
            if (!c) continue;
            es_hist16_inc(s.hist2, wd * ES_CBINS + c);
            atomicOr(&s.occ2[wd], 1u << c);
            atomicAdd(&s.misc[ES_DISTINCT0 + wd], 1);
        }
        if (SCORE) {  // H1..H3 pairs + H4 (14-day windows with count > 3) ; S1 (7-day windows, count > 2)
            u64 p14[4], p7[3];
            es_window_planes<14, 4>(m, p14);
            es_window_planes<7, 3>(m, p7);
            const int eh = es_pair_terms(m, hol[e], K) + __popcll((p14[2] | p14[3]) & v14);
            const int es = __popcll(((p7[0] & p7[1]) | p7[2]) & v7);
            atomicAdd(&s.misc[ES_HARD], eh);
            atomicAdd(&s.misc[ES_SOFT], es);
        }
    }
    __syncthreads();
}

// Build masks + tallies from a[] and the full score from them.  Block-cooperative.  The mask
// table must be all-zero on entry (es_clear_masks after the previous chain).
__device__ __forceinline__ void es_zero_masks(const EsSmem& s, int E) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) s.mask[e] = 0;
}
// clear exactly the entries the current a[] set (<= D stores instead of E)
__device__ __forceinline__ void es_clear_masks(const EsSmem& s, int D) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) s.mask[s.a[d]] = 0;
}
__device__ void es_build(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol,
                         int& hard, int& soft) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int d = tid; d < K.D; d += nt) atomicOr(&s.mask[s.a[d]], 1ull << d);
    __syncthreads();
    es_tally<true>(s, K, hol);
    hard = s.misc[ES_HARD];
    soft = s.misc[ES_SOFT];
    const int present = s.misc[ES_PRESENT];
    for (int wd = 0; wd < 5; ++wd) soft += es_s2_term(s.occ2[wd], s.misc[ES_DISTINCT0 + wd]);
    soft += es_spread(*s.occT, present) + es_spread32(*s.occW, present);
    __syncthreads();
}

// Once per chain-step (masks + tallies must be current).  Everything is per day or per
// (day, value in use): no loop over the employee table.
__device__ void es_prepare(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int D = K.D;
    const u64 v14 = K.n14 >= 64 ? ~0ull : ((1ull << K.n14) - 1);  // real window starts only
    const u64 v7 = K.n7 >= 64 ? ~0ull : ((1ull << K.n7) - 1);
    const u64 fm = *s.fmask;
    const u64 occT = *s.occT;
    const unsigned int occW = *s.occW;
    // phase 1: slots (one per first day), their window masks; per-day counts of the day's employee
    for (int d = tid; d < D; d += nt) {
        const int e = s.a[d];
        const u64 m = s.mask[e];
        const int f = __ffsll((long long)m) - 1;
        const int slot = __popcll(fm & ((1ull << f) - 1ull));
        s.dslot[d] = (unsigned char)slot;
        const int t = __popcll(m), w = __popcll(m & K.wkend);
        s.dayb[d] = (unsigned char)t;
        s.dayb[s.dp + d] = (unsigned char)w;
        s.dayb[2 * s.dp + d] = (unsigned char)__popcll(m & s.wdm[d]);
        if (f == d) {
            s.semp[slot] = (uint16_t)e;
            s.smask[slot] = m;
            s.shol[slot] = hol[e];
            s.srk[2 * slot] = (unsigned char)__popcll(occT & (es_bit64(t) - 1ull));
            s.srk[2 * slot + 1] = (unsigned char)__popc(occW & ((1u << w) - 1u));
            u64 p14[4], p7[3];
            es_window_planes<14, 4>(m, p14);
            es_window_planes<7, 3>(m, p7);
            u64* q = s.eq + slot * 4;
            q[0] = p14[0] & p14[1] & ~p14[2] & ~p14[3] & v14;   // count == 3 (one more => H4 violation)
            q[1] = ~p14[0] & ~p14[1] & p14[2] & ~p14[3] & v14;  // count == 4 (one less => violation gone)
            q[2] = ~p7[0] & p7[1] & ~p7[2] & v7;                // count == 2
            q[3] = p7[0] & p7[1] & ~p7[2] & v7;                 // count == 3
        }
    }
    // rank -> value lists of the total / weekend counts in use (j-th set bit), and per weekday the
    // counts in use (0 first): the memo tables below are built for exactly these values
    for (int q = nt - 1 - tid; q < ES_VAL_BYTES; q += nt) {  // the last threads first: the first ones hold first days
        if (q < ES_TCOLS + ES_WCOLS) {
            u64 bits = q < ES_TCOLS ? occT : (u64)occW;
            const int r = q < ES_TCOLS ? q : q - ES_TCOLS;
            for (int k = 0; k < r; ++k) bits &= bits - 1;
            s.val[q] = (unsigned char)(bits ? __ffsll((long long)bits) - 1 : 0xff);
        } else if (q < ES_VAL_C2 + 5 * ES_CBINS) {
            const int wd = (q - ES_VAL_C2) / ES_CBINS, r = (q - ES_VAL_C2) - wd * ES_CBINS;
            unsigned int bits = (s.occ2[wd] & 0x7feu) | 1u;  // counts 1..10 in use, and 0 (a newcomer to the weekday)
            for (int k = 0; k < r; ++k) bits &= bits - 1;
            s.val[q] = (unsigned char)(bits ? __ffs((int)bits) - 1 : 0xff);
        } else if (q < ES_VAL_N2 + 5) {
            s.val[q] = (unsigned char)__popc((s.occ2[q - ES_VAL_N2] & 0x7feu) | 1u);
        }
    }
    if (tid == 0) s.misc[ES_NSLOT] = __popcll(fm);
    __syncthreads();
    // phase 2: what the day's current employee loses, the value of an absent receiver, and the
    // memo tables.  The soft deltas of a change move depend on the receiving employee only
    // through its count on the weekday (S2), its total (S3) and its weekend count (S4).
    for (int d = tid; d < s.dp; d += nt) {
        if (d >= D) {
            s.baseW[d] = ES_W_PAD;
            continue;
        }
        const int slot = s.dslot[d];
        const u64 m = s.smask[slot];
        const u64* q = s.eq + slot * 4;
        const int lossH = (int)((s.shol[slot] >> d) & 1ull) + __popcll(m & s.part[d]) + __popcll(q[1] & s.cont14[d]);
        const int lossS = __popcll(q[3] & s.cont7[d]);
        s.base[d] = ((unsigned)(0x8000 - lossH) << 16) | (unsigned)(0x8000 - lossS);
        // an absent receiver: no pairs, no window counts, zero days anywhere
        const int to = s.dayb[d], wo = s.dayb[s.dp + d], wd = s.dwd[d];
        const int isw = wd >= 5 ? 1 : 0;
        int ds = -lossS + es_s3_delta(s, to, 0) + es_s4_delta(s, to, wo, isw, 0, false);
        if (wd < 5) ds += es_s2_delta(s, wd, (int)s.dayb[2 * s.dp + d], 0);
        s.baseW[d] = ((unsigned)(64 - lossH) << 15) | ((unsigned)(256 + ds) << 6) | (unsigned)d;
    }
    // dense loops: only (day, value in use) pairs, so every lane of a warp has work
    {
        int nmax = 1;
#pragma unroll
        for (int wd = 0; wd < 5; ++wd) nmax = max(nmax, (int)s.val[ES_VAL_N2 + wd]);
        for (int k = tid; k < D * nmax; k += nt) {
            const int d = k / nmax, j = k - d * nmax;
            const int wd = s.dwd[d];
            if (wd >= 5) {
                if (j == 0) s.s2t[d * ES_CBINS] = 0;  // weekend days: the only entry ever looked up
            } else if (j < (int)s.val[ES_VAL_N2 + wd]) {
                const int cn = s.val[ES_VAL_C2 + wd * ES_CBINS + j];
                s.s2t[d * ES_CBINS + cn] = (signed char)es_s2_delta(s, wd, (int)s.dayb[2 * s.dp + d], cn);
            }
        }
    }
    {
        const int nT = __popcll(occT);
        for (int k = tid; k < D * nT; k += nt) {
            const int d = k / nT, j = k - d * nT;
            s.s3t[d * ES_TCOLS + j] = (signed char)es_s3_delta(s, s.dayb[d], (int)s.val[j]);
        }
        const int nW = __popc(occW);
        for (int k = tid; k < D * nW; k += nt) {
            const int d = k / nW, j = k - d * nW;
            const int wn = s.val[ES_TCOLS + j];
            const int isw = s.dwd[d] >= 5 ? 1 : 0, wo = s.dayb[s.dp + d];
            s.s4t[d * ES_WCOLS + j] = (signed char)es_s4_delta(s, s.dayb[d], wo, isw, wn, true);
            if (isw) {  // swap with a weekday of an employee holding wn weekend days: nobody joins or leaves
                const int present = s.misc[ES_PRESENT];
                const unsigned int occ = es_occ_move32(s.histW, occW, wo, wn, wo - 1, wn + 1);
                s.s4s[d * ES_WCOLS + j] = (signed char)(es_spread32(occ, present) - es_spread32(occW, present));
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------ move deltas
// Packed candidate value v = (0x8000 + dhard) << 16 | (0x8000 + dsoft): one unsigned compare
// orders candidates lexicographically by (dhard, dsoft).
__device__ __forceinline__ int es_v_dh(unsigned int v) { return (int)(v >> 16) - 0x8000; }
__device__ __forceinline__ int es_v_ds(unsigned int v) { return (int)(v & 0xffffu) - 0x8000; }
// absent-receiver packing w = (dh + 64) << 15 | (ds + 256) << 6 | day  ->  v
// (dh in [-21, 1], ds in [-101, 94]: bounded by the window / weekday / spread ranges for D <= 64)
__device__ __forceinline__ unsigned int es_w_to_v(unsigned int w) {
    return ((unsigned)(0x8000 - 64 + (int)(w >> 15)) << 16) | (unsigned)(0x8000 - 256 + (int)((w >> 6) & 0x1ffu));
}

// change: day d goes to the PRESENT employee of `slot` (not the day's current one).  ga = the
// receiver-side parts (hard gain, S1 gain, S2 delta) packed for the swap pass.
__device__ __forceinline__ unsigned int es_change_present_v(const EsSmem& s, int d, int slot, unsigned int& ga) {
    const u64 m = s.smask[slot];
    const u64* q = s.eq + slot * 4;
    const int gh = (int)((s.shol[slot] >> d) & 1ull) + __popcll(m & s.part[d]) + __popcll(q[0] & s.cont14[d]);
    const int gs = __popcll(q[2] & s.cont7[d]);
    const int cn = __popcll(m & s.wdm[d]);
    const int s2 = (int)s.s2t[d * ES_CBINS + cn];
    ga = (unsigned)gh | ((unsigned)gs << 5) | ((unsigned)(s2 + 16) << 8);
    return s.base[d] + ((unsigned)gh << 16) +
           (unsigned)(gs + s2 + (int)s.s3t[d * ES_TCOLS + (int)s.srk[2 * slot]] +
                      (int)s.s4t[d * ES_WCOLS + (int)s.srk[2 * slot + 1]]);
}
__device__ __forceinline__ unsigned int es_change_present_v(const EsSmem& s, int d, int slot) {
    unsigned int ga;
    return es_change_present_v(s, d, slot, ga);
}

// change: day d goes to an ABSENT employee whose holiday mask is hol
__device__ __forceinline__ unsigned int es_change_absent_v(const EsSmem& s, int d, u64 hol) {
    return es_w_to_v(s.baseW[d] + ((unsigned)((hol >> d) & 1ull) << 15));
}

// swap: days d1 < d2 exchange employees (different).
__device__ __forceinline__ unsigned int es_swap_v(const EsSmem& s, const EsConst& K, int d1, int d2) {
    const int s1 = s.dslot[d1], s2 = s.dslot[d2];
    const u64 b1 = 1ull << d1, b2 = 1ull << d2;
    const u64 m1 = s.smask[s1], m2 = s.smask[s2];
    const u64* q1 = s.eq + s1 * 4;
    const u64* q2 = s.eq + s2 * 4;
    const u64 h1 = s.shol[s1], h2 = s.shol[s2];
    // windows holding exactly one of the two days change count by one for each employee
    const u64 c14a = s.cont14[d1], c14b = s.cont14[d2], c7a = s.cont7[d1], c7b = s.cont7[d2];
    const u64 only14a = c14a & ~c14b, only14b = c14b & ~c14a, only7a = c7a & ~c7b, only7b = c7b & ~c7a;
    int dh = (int)((h1 >> d2) & 1ull) - (int)((h1 >> d1) & 1ull) + (int)((h2 >> d1) & 1ull) -
             (int)((h2 >> d2) & 1ull);
    // H2/H3 pairs: e1 leaves d1 and lands on d2 (its other days: m1 without d1), e2 the reverse
    dh += __popcll((m1 & ~b1) & s.part[d2]) - __popcll(m1 & s.part[d1]);
    dh += __popcll((m2 & ~b2) & s.part[d1]) - __popcll(m2 & s.part[d2]);
    // H4: e1 loses a day in windows with only d1 (count 4 -> 3), gains in windows with only d2
    dh += __popcll(q1[0] & only14b) - __popcll(q1[1] & only14a);
    dh += __popcll(q2[0] & only14a) - __popcll(q2[1] & only14b);
    int ds = __popcll(q1[2] & only7b) - __popcll(q1[3] & only7a);
    ds += __popcll(q2[2] & only7a) - __popcll(q2[3] & only7b);
    const int wd1 = s.dwd[d1], wd2 = s.dwd[d2];
    if (wd1 != wd2) {
        // two independent transfers on distinct weekday histograms: day d1 goes e1 -> e2, day d2
        // goes e2 -> e1; each is the memoised change-move S2 delta for the receiver's count
        // (weekend rows of s2t are zero)
        ds += (int)s.s2t[d1 * ES_CBINS + __popcll(m2 & s.wdm[d1])] + (int)s.s2t[d2 * ES_CBINS + __popcll(m1 & s.wdm[d2])];
        // totals (S3) unchanged; a weekend day and a weekday trade places (S4)
        if (wd1 >= 5 && wd2 < 5) ds += (int)s.s4s[d1 * ES_WCOLS + (int)s.srk[2 * s2 + 1]];
        if (wd2 >= 5 && wd1 < 5) ds += (int)s.s4s[d2 * ES_WCOLS + (int)s.srk[2 * s1 + 1]];
    }
    return ((unsigned)(0x8000 + dh) << 16) | (unsigned)(0x8000 + ds);
}

// The same value from pass A's table (s.ga must be complete): a swap is two simultaneous
// transfers, day d1 -> e2 and day d2 -> e1; their tabulated gains/losses are exact except where
// both days meet -- the H2/H3 pair (d1, d2) itself and the windows holding BOTH days, whose
// counts do not change.
__device__ __forceinline__ unsigned int es_swap_from_table(const EsSmem& s, int d1, int d2) {
    const int s1 = s.dslot[d1], s2 = s.dslot[d2];
    const unsigned int g12 = s.ga[d1 * s.ns + s2], g21 = s.ga[d2 * s.ns + s1];  // d1 -> e2, d2 -> e1
    const unsigned int b1 = s.base[d1], b2 = s.base[d2];
    int dh = (int)(g12 & 31u) + (int)(g21 & 31u) + (int)(b1 >> 16) + (int)(b2 >> 16) - 0x10000;
    int ds = (int)((g12 >> 5) & 7u) + (int)((g21 >> 5) & 7u) + (int)(b1 & 0xffffu) + (int)(b2 & 0xffffu) - 0x10000;
    if (d2 - d1 < 14) {
        if ((s.part[d1] >> d2) & 1ull) dh -= 2;
        const u64* q1 = s.eq + s1 * 4;
        const u64* q2 = s.eq + s2 * 4;
        const u64 both14 = s.cont14[d1] & s.cont14[d2];
        if (both14)
            dh += __popcll(q1[1] & both14) - __popcll(q1[0] & both14) + __popcll(q2[1] & both14) - __popcll(q2[0] & both14);
        const u64 both7 = s.cont7[d1] & s.cont7[d2];
        if (both7)
            ds += __popcll(q1[3] & both7) - __popcll(q1[2] & both7) + __popcll(q2[3] & both7) - __popcll(q2[2] & both7);
    }
    const int wd1 = s.dwd[d1], wd2 = s.dwd[d2];
    if (wd1 != wd2) {
        ds += (int)((g12 >> 8) & 31u) + (int)((g21 >> 8) & 31u) - 32;  // the two S2 transfers (weekend rows are 0)
        if (wd1 >= 5 && wd2 < 5) ds += (int)s.s4s[d1 * ES_WCOLS + (int)s.srk[2 * s2 + 1]];
        if (wd2 >= 5 && wd1 < 5) ds += (int)s.s4s[d2 * ES_WCOLS + (int)s.srk[2 * s1 + 1]];
    }
    return ((unsigned)(0x8000 + dh) << 16) | (unsigned)(0x8000 + ds);
}

// move id: change (d, e) -> d*E + e ; swap (d1<d2) -> D*E + tri(d1,d2)
__device__ __forceinline__ int es_tri_index(int D, int d1, int d2) {
    return d1 * D - d1 * (d1 + 1) / 2 + (d2 - d1 - 1);
}

// key = v << 24 | move id  (id < 2^24: 64 days x 65535 employees + swaps)
__device__ __forceinline__ long long es_key(unsigned int v, int id) {
    return (long long)(((u64)v << 24) | (u64)(unsigned)id);
}

__device__ __forceinline__ long long es_block_min(long long key, u64* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
    }
    if (blockDim.x == 32) return key;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = (u64)key;
    __syncthreads();
    long long x = (l < nw) ? (long long)red[l] : ES_KEY_INF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, x, o);
        x = other < x ? other : x;
    }
    return x;
}

// The neighbourhood scan of one chain-step: every non-identity candidate gets its exact packed
// (dhard, dsoft); returns this thread's minimum key.  Three passes:
//   A  change moves to PRESENT employees  (day x slot; full mask arithmetic)
//   B  change moves to ABSENT employees   (thread per employee, loop over days; the value is the
//      per-day table entry plus the employee's holiday bit, tracked with one min per candidate)
//   C  swaps
template <bool DUMP>
__device__ __forceinline__ long long es_scan(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol,
                                             const uint16_t* __restrict__ tri, long long* dump_h,
                                             long long* dump_s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int D = K.D, E = K.E;
    const int nslot = s.misc[ES_NSLOT];
    long long key = ES_KEY_INF;
    {   // A
        const int nA = nslot * D;
        int d = tid / nslot, slot = tid - d * nslot;
        const int dd = nt / nslot, dsl = nt - dd * nslot;
        for (int k = tid; k < nA; k += nt) {
            const int id = d * E + (int)s.semp[slot];
            if ((int)s.dslot[d] != slot) {
                unsigned int ga;
                const unsigned int v = es_change_present_v(s, d, slot, ga);
                s.ga[d * s.ns + slot] = (uint16_t)ga;
                const long long k2 = es_key(v, id);
                key = k2 < key ? k2 : key;
                if (DUMP) {
                    dump_h[id] = es_v_dh(v);
                    dump_s[id] = es_v_ds(v);
                }
            } else if (DUMP) {
                dump_h[id] = INT64_MAX;
                dump_s[id] = INT64_MAX;
            }
            slot += dsl;
            d += dd;
            if (slot >= nslot) {
                slot -= nslot;
                ++d;
            }
        }
    }
    {   // B
        unsigned int bw = 0xffffffffu;
        int be = 0;
        const uint4* bw4 = (const uint4*)s.baseW;
        for (int e = tid; e < E; e += nt) {
            if (s.mask[e]) continue;
            const u64 h = hol[e];
            unsigned int w = 0xffffffffu;
            for (int d0 = 0; d0 < D; d0 += 4) {
                const uint4 b = bw4[d0 >> 2];
                const unsigned int x = (unsigned int)(h >> d0);  // holiday bits of days d0..d0+3
                const unsigned int w0 = b.x + ((x << 15) & 0x8000u), w1 = b.y + ((x << 14) & 0x8000u);
                const unsigned int w2 = b.z + ((x << 13) & 0x8000u), w3 = b.w + ((x << 12) & 0x8000u);
                w = __vimin3_u32(w, w0, w1);
                w = __vimin3_u32(w, w2, w3);
                if (DUMP) {
                    const unsigned int ww[4] = {w0, w1, w2, w3};
                    for (int j = 0; j < 4 && d0 + j < D; ++j) {
                        const unsigned int v = es_w_to_v(ww[j]);
                        dump_h[(d0 + j) * E + e] = es_v_dh(v);
                        dump_s[(d0 + j) * E + e] = es_v_ds(v);
                    }
                }
            }
            if (w < bw) {  // equal value and day: the lower employee index (seen first) stays
                bw = w;
                be = e;
            }
        }
        if (bw != 0xffffffffu) {
            const long long k2 = es_key(es_w_to_v(bw), (int)(bw & 63u) * E + be);
            key = k2 < key ? k2 : key;
        }
    }
    __syncthreads();  // pass A's table is complete
    {   // C
        const int n_change = D * E, n_swap = D * (D - 1) / 2;
        const uint16_t* __restrict__ scan = tri + ((n_swap + 3) / 4 + 1) * 4;  // near pairs first (host-built)
        for (int r = tid; r < n_swap; r += nt) {
            const int dd = scan[r], d1 = dd >> 8, d2 = dd & 0xff;
            const int id = n_change + es_tri_index(D, d1, d2);
            if (s.dslot[d1] != s.dslot[d2]) {
                const unsigned int v = es_swap_from_table(s, d1, d2);
                const long long k2 = es_key(v, id);
                key = k2 < key ? k2 : key;
                if (DUMP) {
                    dump_h[id] = es_v_dh(v);
                    dump_s[id] = es_v_ds(v);
                }
            } else if (DUMP) {
                dump_h[id] = INT64_MAX;
                dump_s[id] = INT64_MAX;
            }
        }
    }
    return key;
}

// ------------------------------------------------------------------ reference mode
// The reference's OWN neighbourhood for scheduling: ScheduleRandomMoveProposer::iter_local_moves
// (examples/employee-scheduling/src/lib.rs:440-491) is an endless stream of random ChangeDay
// (weight 1) / SwapDays (weight 4) candidates drawn from a CLONE of the LocalSearch rng (:488) --
// so every step replays the same draws from t = 0 -- filtered by the tabu set (== {current},
// local_search.rs:155-199,319), scored, truncated to window_size (:321) and ordered by the derived
// Ord (score, then date_to_employee by Employee.id; :29-37,323).  Candidate k uses draws
// 3k..3k+2 of the chain's Philox stream CS_PHILOX_LS (restated in oracle/cs_oracle.c:
// orc_es_local_search_ref).
struct EsCand {
    unsigned int v;  // packed (dhard, dsoft); 0xffffffff = none
    int kind, x, y;  // change: day x -> employee index y; swap: days x < y
};

__device__ __forceinline__ unsigned int es_ref_draw(unsigned long long seed, unsigned int chain, unsigned long long t) {
    const Philox4 b = philox_stream(seed, chain, 2u /* CS_PHILOX_LS */, t >> 2);
    return b.v[t & 3];
}

// candidate k of the endless stream; returns false for an identity candidate (tabu)
__device__ __forceinline__ bool es_ref_candidate(const EsSmem& s, int D, int E, unsigned long long seed,
                                                 unsigned int chain, unsigned long long k, EsCand& c) {
    const unsigned int u0 = es_ref_draw(seed, chain, 3 * k), u1 = es_ref_draw(seed, chain, 3 * k + 1),
                       u2 = es_ref_draw(seed, chain, 3 * k + 2);
    if (philox_mulhi(u0, 5u) < 1u) {  // choose_weighted over [(ChangeDay, 1), (SwapDays, 4)]
        c.kind = 0;
        c.x = (int)philox_mulhi(u1, (unsigned)D);
        c.y = (int)philox_mulhi(u2, (unsigned)E);
        return (int)s.a[c.x] != c.y;
    }
    c.kind = 1;
    if (D < 2) return false;  // the reference would panic (xs[1] of a one-element sample)
    int d1 = (int)philox_mulhi(u1, (unsigned)D), d2 = (int)philox_mulhi(u2, (unsigned)(D - 1));
    d2 += (d2 >= d1);
    c.x = min(d1, d2);
    c.y = max(d1, d2);
    return s.a[c.x] != s.a[c.y];
}

// derived Ord of ScoredSolution: score first, then the solution vector (employee indices are in
// id order).  A candidate differs from the current rota in at most two slots.
__device__ __forceinline__ bool es_cand_less(const EsSmem& s, const EsCand& A, const EsCand& B) {
    if (A.v != B.v) return A.v < B.v;
    if (A.v == 0xffffffffu) return false;
    int pa[2], va[2], pb[2], vb[2];
    const int na = A.kind == 0 ? 1 : 2, nb = B.kind == 0 ? 1 : 2;
    pa[0] = A.x; va[0] = A.kind == 0 ? A.y : (int)s.a[A.y]; pa[1] = A.y; va[1] = (int)s.a[A.x];
    pb[0] = B.x; vb[0] = B.kind == 0 ? B.y : (int)s.a[B.y]; pb[1] = B.y; vb[1] = (int)s.a[B.x];
    int ia = 0, ib = 0;
    while (ia < na || ib < nb) {
        const int pA = ia < na ? pa[ia] : 1 << 30, pB = ib < nb ? pb[ib] : 1 << 30;
        const int p = min(pA, pB);
        const int xa = pA == p ? va[ia] : (int)s.a[p], xb = pB == p ? vb[ib] : (int)s.a[p];
        if (xa != xb) return xa < xb;
        ia += (pA == p);
        ib += (pB == p);
    }
    return false;  // the same solution
}

// One step's window: returns the block-uniform key (v << 24 | move id) of the window's minimum,
// ES_KEY_INF for an empty window; n_scored = candidates scored (<= window).
__device__ long long es_scan_ref(const EsSmem& s, const EsConst& K, const u64* __restrict__ hol,
                                 unsigned long long seed, unsigned int chain, unsigned long long window,
                                 unsigned long long max_draws, unsigned int& n_scored) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5, nw = (nt + 31) >> 5;
    const int D = K.D, E = K.E;
    int* cnt = (int*)s.red;  // [nw + 1] per-warp counts of valid candidates
    EsCand best{0xffffffffu, 0, 0, 0};
    unsigned long long count = 0;
    for (unsigned long long base = 0; count < window && base < max_draws; base += (unsigned)nt) {
        EsCand c{0xffffffffu, 0, 0, 0};
        const bool valid = es_ref_candidate(s, D, E, seed, chain, base + (unsigned)tid, c);
        const unsigned int bal = __ballot_sync(0xffffffffu, valid);
        int before = __popc(bal & ((1u << lane) - 1u)), total = __popc(bal);
        if (nw > 1) {
            __syncthreads();
            if (lane == 0) cnt[w] = total;
            __syncthreads();
            total = 0;
            for (int q = 0; q < nw; ++q) {
                if (q < w) before += cnt[q];
                total += cnt[q];
            }
        }
        if (valid && count + (unsigned long long)before < window) {
            if (c.kind == 0) {
                const u64 m = s.mask[c.y];
                c.v = m ? es_change_present_v(s, c.x, es_slot_of(s, m)) : es_change_absent_v(s, c.x, hol[c.y]);
            } else {
                c.v = es_swap_v(s, K, c.x, c.y);
            }
            if (es_cand_less(s, c, best)) best = c;
        }
        count += (unsigned long long)total;
    }
    n_scored = (unsigned int)(count < window ? count : window);
    // block argmin under the derived Ord
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        EsCand t;
        t.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        t.kind = __shfl_xor_sync(0xffffffffu, best.kind, o);
        t.x = __shfl_xor_sync(0xffffffffu, best.x, o);
        t.y = __shfl_xor_sync(0xffffffffu, best.y, o);
        if (es_cand_less(s, t, best)) best = t;
    }
    if (nw > 1) {
        EsCand* slot = (EsCand*)(s.red + 8);  // 16 B per warp, after the counts
        __syncthreads();
        if (lane == 0) slot[w] = best;
        __syncthreads();
        best = slot[0];
        for (int q = 1; q < nw; ++q)
            if (es_cand_less(s, slot[q], best)) best = slot[q];
        __syncthreads();
    }
    if (best.v == 0xffffffffu) return ES_KEY_INF;
    const int id = best.kind == 0 ? best.x * E + best.y : D * E + es_tri_index(D, best.x, best.y);
    return es_key(best.v, id);
}

// per-day constants of the handle: part | cont14 | cont7 from the host table, weekday masks
__device__ __forceinline__ void es_load_consts(const EsSmem& s, const EsConst& K, const u64* __restrict__ dayconst) {
    for (int d = threadIdx.x; d < s.dp; d += blockDim.x) {
        const int wd = (K.start_wd + d) % 7;
        s.part[d] = dayconst[d];
        s.cont14[d] = dayconst[64 + d];
        s.cont7[d] = dayconst[128 + d];
        s.wdm[d] = (d < K.D && wd < 5) ? K.wd[wd] : 0ull;
        s.dwd[d] = (unsigned char)wd;
    }
}
// (d1 << 8 | d2) of swap r, appended to the day constants by the host
__device__ __forceinline__ const uint16_t* es_tri_table(const u64* dayconst) { return (const uint16_t*)(dayconst + 192); }

// ------------------------------------------------------------------ the step kernel (K5)
template <bool REF>
__global__ void __launch_bounds__(256, 4) es_step_kernel(EsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsConst& K = p.K;
    const EsSmem s = es_carve(smem_raw, K.D, K.E);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int D = K.D, E = K.E;
    const int n_change = D * E, n_swap = D * (D - 1) / 2;
    const uint16_t* tri = es_tri_table(p.dayconst);
    es_load_consts(s, K, p.dayconst);
    es_zero_masks(s, E);

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
            // non-identity candidates: every (day, employee != current) + every day pair held by
            // two different employees -- each of them is evaluated by es_scan
            long long key;
            if (REF && !p.dump_h) {  // the reference's sampled window instead of the whole neighbourhood
                unsigned int nsc = 0;
                key = es_scan_ref(s, K, p.hol, p.seed, p.chain_offset + (unsigned)chain, p.window, p.max_draws, nsc);
                scored += nsc;
            } else {
                scored += (unsigned long long)(n_change - D) + (unsigned long long)(n_swap - s.misc[ES_SAME]);
                key = p.dump_h ? es_scan<true>(s, K, p.hol, tri, p.dump_h, p.dump_s)
                               : es_scan<false>(s, K, p.hol, tri, nullptr, nullptr);
                key = es_block_min(key, s.red);
            }
            if (p.dump_h) break;
            if (key == ES_KEY_INF) {  // empty neighbourhood, local_search.rs:336-338
                status = 3;
                break;
            }
            const unsigned int v = (unsigned int)((u64)key >> 24);
            const int dh = es_v_dh(v), ds = es_v_ds(v);
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
            __syncthreads();  // every thread has read the tables of this step
            if (tid == 0) {
                unsigned int kind, x, y;
                if (id < n_change) {
                    const int d = id / E, e = id - d * E, eo = s.a[d];
                    const u64 bit = 1ull << d;
                    s.mask[eo] &= ~bit;
                    s.mask[e] |= bit;
                    s.a[d] = (uint16_t)e;
                    kind = 0;
                    x = (unsigned)d;
                    y = (unsigned)e;
                } else {
                    const int dd = tri[id - n_change], d1 = dd >> 8, d2 = dd & 0xff;
                    const int e1 = s.a[d1], e2 = s.a[d2];
                    const u64 x2 = (1ull << d1) | (1ull << d2);
                    s.mask[e1] ^= x2;
                    s.mask[e2] ^= x2;
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
            if (it + 1 < p.max_steps) es_tally<false>(s, K, p.hol);  // tallies of the new state
        }
        __syncthreads();
        es_clear_masks(s, D);  // leave the mask table all-zero for the next chain
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
    es_zero_masks(s, p.K.E);
    for (int local = blockIdx.x; local < p.n_chains; local += gridDim.x) {
        const int chain = p.first_chain + local;
        __syncthreads();
        for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
            s.a[k] = p.a[(size_t)chain * p.stride + k];
        __syncthreads();
        int hard, soft;
        es_build(s, p.K, p.hol, hard, soft);
        es_clear_masks(s, p.K.D);
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
    es_load_consts(s, p.K, p.dayconst);
    es_zero_masks(s, p.K.E);
    __syncthreads();
    int hard, soft;
    es_build(s, p.K, p.hol, hard, soft);
    es_prepare(s, p.K, p.hol);
    for (unsigned long long k = threadIdx.x; k < n_moves; k += blockDim.x) {
        const uint2 mv = moves[k];
        unsigned int v = 0;
        bool ok;
        if (kind == 0) {
            const int d = (int)mv.x, en = (int)mv.y;
            ok = s.a[d] != en;
            if (ok) {
                const u64 m = s.mask[en];
                v = m ? es_change_present_v(s, d, es_slot_of(s, m)) : es_change_absent_v(s, d, p.hol[en]);
            }
        } else {
            const int d1 = (int)min(mv.x, mv.y), d2 = (int)max(mv.x, mv.y);
            ok = d1 != d2 && s.a[d1] != s.a[d2];
            if (ok) v = es_swap_v(s, p.K, d1, d2);
        }
        dh_out[k] = ok ? (long long)es_v_dh(v) : INT64_MAX;
        ds_out[k] = ok ? (long long)es_v_ds(v) : INT64_MAX;
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
