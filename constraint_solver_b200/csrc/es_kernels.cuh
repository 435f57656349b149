// Employee-scheduling (on-call rota) chain kernels for sm_100a.
//
// Score definition: examples/employee-scheduling/src/lib.rs:261-375 (+ :194-218).  The reference
// has one employee per calendar day; this file carries T = D x S SLOTS (D days, S shifts per
// day, slot t = day t / S, shift t % S) and at S = 1 is exactly the reference's rota.  E
// employees (dense index 0..E-1).  Template parameters: W = 64-bit words per slot mask
// (T <= 64 W); MULTI = the slot-generalised extension (S > 1 and / or a skill table; NOT pinned by
// the reference -- its full-re-score definition is oracle/cs_oracle.c: esx_terms).
//
// Device formulation.  For every employee e keep the slot mask m_e (bit t <=> a[t] == e).  All
// hard terms and S1 are sums over employees of a function of m_e alone (S = 1 shown):
//   H1_e = popc(m & holiday_e)                                   (:273-280)
//   H2_e = popc(m & m>>1)                                        (:286-292)
//   H3_e = popc(m & m>>7 & SATF) + popc(m & m>>8 & SATF)
//        + popc(m>>1 & m>>7 & SATF) + popc(m>>1 & m>>8 & SATF)   (:295-315)
//   H4_e = #{w : popc(m & W14<<w) > 3}                           (:318-327)
//   S1_e = #{w : popc(m & W7<<w)  > 2}                           (:330-339)
// (MULTI: shifts by S, 7S, 8S; windows over per-DAY slot counts; X1 = same-day pairs
// popc(m & m>>1 & SD1) + popc(m & m>>2 & SD2); X2 = popc(m & unskilled_e)) so a move's delta on
// these terms is F(m') - F(m) over the (at most two) employees whose mask changes.
// Hot-loop formulation (es_tally -> es_prepare -> es_scan).  At most T
// employees are PRESENT (hold a slot); each gets a "owner" index = rank of its first slot, so every
// per-step table is built by <= T threads and nothing loops over the employee table.  Per owner
// all sliding-window counts are computed at once by a bit-sliced adder and kept as "count == k"
// window-start masks (EQ3_14, EQ4_14, EQ2_7, EQ3_7); per slot d the terms that only depend on the
// slot's current employee are tabulated (base[d]).  A change candidate to a present employee then
// needs three mask popcounts for H2+H3(+X1), H4 and S1:
//   gain of slot d for owner s = popc(m_s & PART[d]) + popc(EQ3_14[s] & CONT14[d]) ...
// where PART[d] = slots paired with d by H2/H3 and CONT[d] = window starts whose window holds d's
// day; a candidate to an ABSENT employee is baseW[d] plus its holiday (and skill) bit; a swap is
// two such transfers (pass A's gain table) corrected where both slots meet.
// S2 (weekday affinity, min over employees present on that weekday), S3 (max-min of total
// slots over PRESENT employees) and S4 (max-min of weekend slots over present employees) are
// kept as count histograms + occupancy bitsets; their deltas depend on the receiver only through
// a count in use, so they are memoised per (slot, value in use) as closed-form occupancy moves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "es_bits.cuh"
#include "philox.cuh"

namespace csb {

constexpr int ES_MAX_SLOTS = 192;
#ifndef ES_BQ_VALUE
#define ES_BQ_VALUE 2
#endif
constexpr int ES_BQ = ES_BQ_VALUE;  // slots per row of pass B's absent-receiver table (2: 8-byte rows, 4: 16-byte rows)
static_assert(ES_BQ == 2 || ES_BQ == 4, "pass B reads its table rows as uint2 or uint4");
constexpr long long ES_KEY_INF = 0x7fffffffffffffffll;

// table dimensions as a function of the mask width (value ranges for T <= 64 W slots)
template <int W>
struct EsDim {
    static constexpr int OW = W == 1 ? 1 : W + 1;       // words of the total-count occupancy set (bins 0..T)
    static constexpr int CBINS = W == 1 ? 12 : 32;      // per-weekday count bins (a weekday holds <= 10 / 31 slots)
    static constexpr int TBINS = 64 * W + 2;            // total-slot bins 0..T
    static constexpr int WBINS = W == 1 ? 24 : 64;      // weekend-slot bins (<= 20 / 60)
    static constexpr int TCOLS = W == 1 ? 12 : 20;      // distinct totals among present employees (sum <= T)
    static constexpr int WCOLS = W == 1 ? 8 : 12;       // distinct weekend counts (0 included)
    static constexpr int VAL_C2 = TCOLS + WCOLS;
    static constexpr int VAL_N2 = VAL_C2 + 5 * CBINS;
    static constexpr int VAL_BYTES = VAL_N2 + 8;
    static constexpr unsigned int CMASK = W == 1 ? 0x7feu : 0xfffffffeu;  // counts >= 1 that can be in use
    typedef typename std::conditional<W == 1, signed char, short>::type s3_t;  // S3 deltas reach +-T
};

template <int W>
struct EsConstT {
    int D, S, T, E, start_wd, n14, n7;  // D days, S shifts/day, T = D*S slots; n14 / n7 window starts (days)
    Bits<W> valid, wkend, satf, sd1, sd2;  // slot masks: scored, weekend, Saturdays opening an H3 window,
                                           // slots whose next / next-but-one slot is on the same day
    Bits<W> wd[5];                         // slots per weekday Mon..Fri
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

template <int W>
struct EsParamsT {
    EsConstT<W> K;
    int first_chain, n_chains, stride;  // stride = T + 1 slots (phantom last, lib.rs:405-412)
    uint16_t* a;                        // [*, stride] employee index per slot
    uint16_t* best_a;
    const Bits<W>* hol;    // [E] holiday slot-mask per employee
    const Bits<W>* unsk;   // [E] MULTI: slots whose shift kind the employee is NOT qualified for
    const u64* cnt2;       // [E][2 W] MULTI: holiday + unskilled bits of every slot as a 2-bit count (pass B's addend)
    const Bits<W>* slotc;  // [4][dp]: PART (H2/H3 partner slots), CONT14, CONT7 (window starts holding the slot's day), PARTX (same-day slots)
    const uint16_t* tri;   // [n_swap] (d1 << 8 | d2), enumeration order; then the same pairs in scan order
    const uint16_t* tri_scan;
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

// per-chain shared state.  NS = min(E, T) bounds the number of PRESENT employees ("owners");
// DP = T rounded up to 4 sizes the per-slot arrays.
template <int W>
struct EsSmemT {
    typedef EsDim<W> Dm;
    Bits<W>* mask;      // [E]
    uint16_t* a;        // [stride]
    uint16_t* hist2;    // [5][CBINS]
    uint16_t* histT;    // [TBINS]
    uint16_t* histW;    // [WBINS]
    unsigned int* occ2; // [5]  bit c: some employee has exactly c slots on that weekday (c>=1)
    Bits<Dm::OW>* occT; // [1]  bit c: some employee has exactly c slots in total (c>=1)
    u64* occW;          // [1]  bit c: some PRESENT employee has exactly c weekend slots (c>=0)
    Bits<W>* fmask;     // [1]  bit d: slot d is the FIRST slot of its employee (one bit per present employee)
    int* misc;          // [16] present, distinct[5], hard, soft, ...
    u64* red;           // [40] reduction scratch
    Bits<W>* part;      // [DP] H2/H3 partner-slot mask per slot
    Bits<W>* cont14;    // [DP] 14-day window starts whose window contains the slot's day
    Bits<W>* cont7;     // [DP]
    Bits<W>* wdm;       // [DP] slots on the same weekday as d (0 for weekend slots)
    Bits<W>* partx;     // [DP] MULTI: the other slots of d's day
    Bits<W>* eq;        // [4][NS] per owner: EQ3_14, EQ4_14, EQ2_7, EQ3_7 ("count == k" window starts), plane-major
    Bits<W>* smask;     // [NS] slot mask of the owner's employee
    Bits<W>* shol;      // [NS] its holiday mask
    Bits<W>* sunsk;     // [NS] MULTI: its unskilled-slot mask
    uint16_t* semp;     // [NS] its employee index
    unsigned char* srk;    // [NS][2] rank of its total / weekend count among the values in use
    unsigned char* swd;    // [NS][8] W > 1: its slots per weekday Mon..Fri (entries 5..7 stay 0: weekend slots carry no S2 term)
    unsigned char* val;    // [VAL_BYTES] rank -> value lists
    unsigned char* dwd;    // [DP] weekday of the slot's day (0 = Monday)
    unsigned char* sday;   // [DP] day of the slot
    unsigned char* dslot;  // [DP] owner index of the slot's current employee
    unsigned char* dayb;   // [3][DP] per slot, for its current employee: total, weekend, weekday count
    unsigned int* base;    // [DP] packed (0x8000 - lossH) << 16 | (0x8000 - lossS1) of the slot's employee
    unsigned int* baseW;   // [DP] absent receiver without a holiday: (dh + 64) << 18 | (ds + 512) << 8 | d
    unsigned int* bwq;     // [DP/BQ][2^BQ][BQ] !MULTI: baseW of slots BQ*q.. plus the holiday addend of bit pattern k (pass B)
    signed char* s2t;      // [DP][CBINS] S2 delta of giving slot d to an employee with cn slots on that weekday
    typename Dm::s3_t* s3t;  // [DP][TCOLS] S3 delta ... to a present employee whose total has rank j
    signed char* s4t;      // [DP][WCOLS] S4 delta ... to a present employee whose weekend count has rank j
    signed char* s4s;      // [DP][WCOLS] S4 delta of SWAPPING weekend slot d with a weekday slot of such an employee
    uint16_t* ga;          // [DP][NS] pass A's gains of giving slot d to owner: gh | gs << 5 | (s2 + 32) << 8
    int ns, dp;
};

struct EsLayout {
    size_t mask, a, hist, occ, occT, occW, fmask, misc, red, slot, eq, smask, shol, sunsk, semp, srk, swd, val, dwd, sday,
        dslot, dayb, base, baseW, bwq, s2t, s3t, s4t, s4s, ga, total;
    int ns, dp;
};
__host__ __device__ inline size_t es_align(size_t x, size_t a) { return (x + a - 1) / a * a; }
template <int W, bool MULTI>
__host__ __device__ inline EsLayout es_layout(int T, int E) {
    typedef EsDim<W> Dm;
    EsLayout L;
    L.ns = E < T ? E : T;
    if (L.ns < 1) L.ns = 1;
    L.dp = (T + 3) & ~3;
    const size_t dp = (size_t)L.dp, B = (size_t)W * 8;
    size_t o = 0;
    L.mask = o;    o += (size_t)E * B;
    L.occT = o;    o += (size_t)Dm::OW * 8;
    L.occW = o;    o += 8;
    L.fmask = o;   o += B;
    L.red = o;     o += 40 * 8;
    L.slot = o;    o += (MULTI ? 5 : 4) * dp * B;
    L.eq = o;      o += (size_t)L.ns * 4 * B;
    L.smask = o;   o += (size_t)L.ns * B;
    L.shol = o;    o += (size_t)L.ns * B;
    L.sunsk = o;   o += MULTI ? (size_t)L.ns * B : 0;
    o = es_align(o, 16);
    L.bwq = o;     o += MULTI ? 0 : (dp / ES_BQ) * (1u << ES_BQ) * ES_BQ * 4;
    L.baseW = o;   o += dp * 4;  // read as uint4
    L.base = o;    o += dp * 4;
    L.misc = o;    o += 16 * 4;
    L.occ = o;     o = es_align(o + 5 * 4, 8);
    L.hist = o;    o = es_align(o + (5 * Dm::CBINS + Dm::TBINS + Dm::WBINS) * 2, 8);
    L.a = o;       o = es_align(o + (size_t)(T + 1) * 2, 8);
    L.semp = o;    o = es_align(o + (size_t)L.ns * 2, 8);
    L.srk = o;     o = es_align(o + (size_t)L.ns * 2, 8);
    L.swd = o;     o += W > 1 ? (size_t)L.ns * 8 : 0;  // one word: the popcount itself is as cheap as the lookup
    L.val = o;     o = es_align(o + Dm::VAL_BYTES, 8);
    L.dwd = o;     o += dp;
    L.sday = o;    o += dp;
    L.dslot = o;   o += dp;
    L.dayb = o;    o += 3 * dp;
    L.s2t = o;     o += dp * Dm::CBINS;
    L.s4t = o;     o += dp * Dm::WCOLS;
    L.s4s = o;     o += dp * Dm::WCOLS;
    o = es_align(o, 8);
    L.s3t = o;     o += dp * Dm::TCOLS * sizeof(typename Dm::s3_t);
    o = es_align(o, 8);
    L.ga = o;      o += dp * (size_t)L.ns * 2;
    L.total = es_align(o, 16);
    return L;
}
template <int W, bool MULTI>
__host__ __device__ inline size_t es_smem_bytes(int T, int E) { return es_layout<W, MULTI>(T, E).total; }

// offsets only (no pointer <-> integer casts) so the compiler keeps the shared address space
template <int W, bool MULTI>
__device__ __forceinline__ EsSmemT<W> es_carve(unsigned char* p, int T, int E) {
    typedef EsDim<W> Dm;
    const EsLayout L = es_layout<W, MULTI>(T, E);
    EsSmemT<W> s;
    s.ns = L.ns;
    s.dp = L.dp;
    s.mask = (Bits<W>*)(p + L.mask);
    s.a = (uint16_t*)(p + L.a);
    s.hist2 = (uint16_t*)(p + L.hist);
    s.histT = s.hist2 + 5 * Dm::CBINS;
    s.histW = s.histT + Dm::TBINS;
    s.occ2 = (unsigned int*)(p + L.occ);
    s.occW = (u64*)(p + L.occW);
    s.occT = (Bits<Dm::OW>*)(p + L.occT);
    s.fmask = (Bits<W>*)(p + L.fmask);
    s.misc = (int*)(p + L.misc);
    s.red = (u64*)(p + L.red);
    s.part = (Bits<W>*)(p + L.slot);
    s.cont14 = s.part + L.dp;
    s.cont7 = s.cont14 + L.dp;
    s.wdm = s.cont7 + L.dp;
    s.partx = s.wdm + (MULTI ? L.dp : 0);
    s.eq = (Bits<W>*)(p + L.eq);
    s.smask = (Bits<W>*)(p + L.smask);
    s.shol = (Bits<W>*)(p + L.shol);
    s.sunsk = (Bits<W>*)(p + L.sunsk);
    s.semp = (uint16_t*)(p + L.semp);
    s.srk = p + L.srk;
    s.swd = p + L.swd;
    s.val = p + L.val;
    s.dwd = p + L.dwd;
    s.sday = p + L.sday;
    s.dslot = p + L.dslot;
    s.dayb = p + L.dayb;
    s.base = (unsigned int*)(p + L.base);
    s.baseW = (unsigned int*)(p + L.baseW);
    s.bwq = (unsigned int*)(p + L.bwq);
    s.s2t = (signed char*)(p + L.s2t);
    s.s3t = (typename Dm::s3_t*)(p + L.s3t);
    s.s4t = (signed char*)(p + L.s4t);
    s.s4s = (signed char*)(p + L.s4s);
    s.ga = (uint16_t*)(p + L.ga);
    return s;
}

// the four "count == k" window-start masks of one owner; stored plane-major ([4][NS]) so the lanes of
// pass A (consecutive owners) read consecutive words instead of striding by 32 bytes
template <int W>
struct EsEq {
    Bits<W>* base;
    int ns;
    __device__ __forceinline__ Bits<W>& operator[](int k) const { return base[k * ns]; }
};
// PK ("packed day sets"): a rota of at most 38 days has at most 25 14-day and 32 7-day window starts, so
// both families share one 64-bit word -- 14-day sets in bits 0..31, 7-day sets in bits 32..63.  The step
// kernel then keeps CONT[d] = cont14 | cont7 << 32 in cont14[d].w[0] and, per owner, the "one more slot is a
// violation" planes eq3_14 | eq2_7 << 32 in q[0] and the "one slot fewer ends a violation" planes
// eq4_14 | eq3_7 << 32 in q[1]: one AND and the two half-word popcounts give the hard and the soft count.
constexpr int ES_PK_MAX_DAYS = 38;
__device__ __forceinline__ int es_popc_lo(u64 x) { return __popc((unsigned int)x); }
__device__ __forceinline__ int es_popc_hi(u64 x) { return __popc((unsigned int)(x >> 32)); }

template <int W>
__device__ __forceinline__ EsEq<W> es_eq(const EsSmemT<W>& s, int slot) {
    return EsEq<W>{s.eq + slot, s.ns};
}

enum { ES_PRESENT = 0, ES_DISTINCT0 = 1, ES_HARD = 6, ES_SOFT = 7, ES_BCAST = 8, ES_NSLOT = 10, ES_SAME = 11 };
constexpr unsigned int ES_W_PAD = 0x7f000000u;  // larger than any packed absent-candidate value, no overflow on +hol
constexpr int ES_W_DH = 18, ES_W_DS = 8;        // field positions of the absent-receiver packing

// ------------------------------------------------------------------ histogram helpers
// Occupancy bitset after one member leaves bin r0, one leaves r1 (r1 < 0: nobody), one enters
// a0 (a0 < 0: nobody) and one enters a1.  hist holds the current member count per bin.  A bin one
// past the set (one employee holding all 64 W slots) has no bit -- it can only occur with a single
// present employee, where the spread is 0 whatever the set says.
template <int OW>
__device__ __forceinline__ Bits<OW> es_occ_move(const uint16_t* hist, Bits<OW> occ, int r0, int r1, int a0, int a1) {
    int c0 = (int)hist[r0] - 1;
    if (r1 >= 0) {
        const int same = (r1 == r0) ? 1 : 0;
        c0 -= same;
        if ((int)hist[r1] - 1 - same <= 0) occ = occ & ~Bits<OW>::bit(r1);
    }
    if (c0 <= 0) occ = occ & ~Bits<OW>::bit(r0);
    if (a0 >= 0) occ = occ | Bits<OW>::bit(a0);
    return occ | Bits<OW>::bit(a1);
}
__device__ __forceinline__ u64 es_occ_move64(const uint16_t* hist, u64 occ, int r0, int r1, int a0, int a1) {
    int c0 = (int)hist[r0] - 1;
    if (r1 >= 0) {
        const int same = (r1 == r0) ? 1 : 0;
        c0 -= same;
        if ((int)hist[r1] - 1 - same <= 0) occ &= ~(1ull << r1);
    }
    if (c0 <= 0) occ &= ~(1ull << r0);
    if (a0 >= 0) occ |= 1ull << a0;
    return occ | (1ull << a1);
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

template <int OW>
__device__ __forceinline__ int es_spread(const Bits<OW>& occ, int members) {  // max-min, lib.rs:349,363
    if (members < 2 || !occ.any()) return 0;
    return occ.fls() - occ.ffs();
}
__device__ __forceinline__ int es_spread64(u64 occ, int members) {
    if (members < 2 || occ == 0) return 0;
    return (63 - __clzll((long long)occ)) - (__ffsll((long long)occ) - 1);
}

__device__ __forceinline__ int es_s2_term(unsigned int occ, int distinct) {  // lib.rs:206-214
    return (distinct >= 2 && occ) ? (__ffs((int)occ) - 1) : 0;
}

// weekday-affinity delta on weekday wd when one employee's count there goes cm -> cm-1
// (cm >= 1) and another's goes cp -> cp+1 (cp >= 0)
template <int W>
__device__ __forceinline__ int es_s2_delta(const EsSmemT<W>& s, int wd, int cm, int cp) {
    const int distinct = s.misc[ES_DISTINCT0 + wd];
    const unsigned int occ0 = s.occ2[wd];
    const unsigned int occ = es_occ_move32(s.hist2 + wd * EsDim<W>::CBINS, occ0, cm, cp >= 1 ? cp : -1,
                                           cm - 1 >= 1 ? cm - 1 : -1, cp + 1);
    return es_s2_term(occ, distinct - (cm == 1) + (cp == 0)) - es_s2_term(occ0, distinct);
}

// S3 delta (max-min of total slots over PRESENT employees, lib.rs:345-351) when a slot moves from
// an employee with `to` slots to one with `tn` slots (0 = absent so far)
template <int W>
__device__ __forceinline__ int es_s3_delta(const EsSmemT<W>& s, int to, int tn) {
    const int present = s.misc[ES_PRESENT];
    const Bits<EsDim<W>::OW> occ0 = *s.occT;
    const Bits<EsDim<W>::OW> occ = es_occ_move(s.histT, occ0, to, tn >= 1 ? tn : -1, to - 1 >= 1 ? to - 1 : -1, tn + 1);
    return es_spread(occ, present - (to == 1) + (tn == 0)) - es_spread(occ0, present);
}
// S4 delta (weekend slots, lib.rs:354-365): the slot (weekend flag isw) leaves an employee with
// (to total, wo weekend) slots for one with wn weekend slots (absent = !rpresent, wn = 0)
template <int W>
__device__ __forceinline__ int es_s4_delta(const EsSmemT<W>& s, int to, int wo, int isw, int wn, bool rpresent) {
    const int present = s.misc[ES_PRESENT];
    const u64 occ = es_occ_move64(s.histW, *s.occW, wo, rpresent ? wn : -1, to - 1 >= 1 ? wo - isw : -1, wn + isw);
    return es_spread64(occ, present - (to == 1) + (rpresent ? 0 : 1)) - es_spread64(*s.occW, present);
}

// ------------------------------------------------------------------ tallies and per-step tables
// All sliding-window counts of one slot mask at once (S = 1): bit-sliced adder over the L shifted
// copies of m; plane i bit w = bit i of popc(m & (ONES(L) << w)).
template <int L, int PLANES, int W>
__device__ __forceinline__ void es_window_planes(const Bits<W>& m, Bits<W> (&pl)[PLANES]) {
#pragma unroll
    for (int i = 0; i < PLANES; ++i) pl[i] = Bits<W>::zero();
#pragma unroll
    for (int k = 0; k < L; ++k) {
        Bits<W> carry = m.shr(k);
#pragma unroll
        for (int i = 0; i < PLANES; ++i) {
            const Bits<W> t = pl[i] & carry;
            pl[i] = pl[i] ^ carry;
            carry = t;
        }
    }
}
// MULTI: the per-day slot counts (0..3, planes c0 / c1 in DAY space) summed over L-day windows
template <int L, int PLANES, int W>
__device__ __forceinline__ void es_window_planes2(const Bits<W>& c0, const Bits<W>& c1, Bits<W> (&pl)[PLANES]) {
#pragma unroll
    for (int i = 0; i < PLANES; ++i) pl[i] = Bits<W>::zero();
#pragma unroll
    for (int k = 0; k < L; ++k) {
        Bits<W> carry = c0.shr(k);
#pragma unroll
        for (int i = 0; i < PLANES; ++i) {
            const Bits<W> t = pl[i] & carry;
            pl[i] = pl[i] ^ carry;
            carry = t;
        }
        carry = c1.shr(k);
#pragma unroll
        for (int i = 1; i < PLANES; ++i) {
            const Bits<W> t = pl[i] & carry;
            pl[i] = pl[i] ^ carry;
            carry = t;
        }
    }
}
// window starts whose count equals V
template <int V, int PLANES, int W>
__device__ __forceinline__ Bits<W> es_planes_eq(const Bits<W> (&pl)[PLANES]) {
    Bits<W> r = ~Bits<W>::zero();
#pragma unroll
    for (int i = 0; i < PLANES; ++i) r = r & (((V >> i) & 1) ? pl[i] : ~pl[i]);
    return r;
}
// window starts whose count exceeds V (V = 3: planes >= 2 set; V = 2: (p0 & p1) | higher)
template <int V, int PLANES, int W>
__device__ __forceinline__ Bits<W> es_planes_gt(const Bits<W> (&pl)[PLANES]) {
    static_assert(V == 2 || V == 3, "limits of the reference: 3 per 14 days, 2 per 7 days");
    Bits<W> r = V == 2 ? (pl[0] & pl[1]) : Bits<W>::zero();
#pragma unroll
    for (int i = 2; i < PLANES; ++i) r = r | pl[i];
    return r;
}

template <int W>
struct EsWin {  // the four "count == k" masks and the two violation counts of one employee
    Bits<W> eq3_14, eq4_14, eq2_7, eq3_7;
    int viol14, viol7;
};

// per-day slot counts of a slot mask as two bit planes in day space (MULTI)
template <int W>
__device__ __forceinline__ void es_day_planes(const EsSmemT<W>& s, const Bits<W>& m, Bits<W>& c0, Bits<W>& c1) {
    c0 = Bits<W>::zero();
    c1 = Bits<W>::zero();
#pragma unroll
    for (int i = 0; i < W; ++i) {
        u64 x = m.w[i];
        while (x) {
            const int t = 64 * i + __ffsll((long long)x) - 1;
            x &= x - 1ull;
            const Bits<W> b = Bits<W>::bit((int)s.sday[t]);
            const Bits<W> carry = c0 & b;
            c0 = c0 ^ b;
            c1 = c1 ^ carry;  // a day holds <= 3 slots: two planes never overflow
        }
    }
}

template <int W, bool MULTI>
__device__ __forceinline__ EsWin<W> es_windows(const EsSmemT<W>& s, const EsConstT<W>& K, const Bits<W>& m) {
    const Bits<W> v14 = Bits<W>::lowmask(K.n14), v7 = Bits<W>::lowmask(K.n7);  // real window starts only
    EsWin<W> r;
    if (!MULTI) {
        Bits<W> p14[4], p7[3];
        es_window_planes<14, 4>(m, p14);
        es_window_planes<7, 3>(m, p7);
        r.eq3_14 = es_planes_eq<3>(p14) & v14;  // one more => H4 violation
        r.eq4_14 = es_planes_eq<4>(p14) & v14;  // one less => violation gone
        r.eq2_7 = es_planes_eq<2>(p7) & v7;
        r.eq3_7 = es_planes_eq<3>(p7) & v7;
        r.viol14 = (es_planes_gt<3>(p14) & v14).popc();
        r.viol7 = (es_planes_gt<2>(p7) & v7).popc();
    } else {
        Bits<W> c0, c1, p14[6], p7[5];
        es_day_planes(s, m, c0, c1);
        es_window_planes2<14, 6>(c0, c1, p14);
        es_window_planes2<7, 5>(c0, c1, p7);
        r.eq3_14 = es_planes_eq<3>(p14) & v14;
        r.eq4_14 = es_planes_eq<4>(p14) & v14;
        r.eq2_7 = es_planes_eq<2>(p7) & v7;
        r.eq3_7 = es_planes_eq<3>(p7) & v7;
        r.viol14 = (es_planes_gt<3>(p14) & v14).popc();
        r.viol7 = (es_planes_gt<2>(p7) & v7).popc();
    }
    return r;
}

// H1 + H2 + H3 (+ X1 + X2) of one employee from its mask
template <int W, bool MULTI>
__device__ __forceinline__ int es_pair_terms(const Bits<W>& m, const Bits<W>& hol, const Bits<W>& unsk,
                                             const EsConstT<W>& K) {
    if (!MULTI) {
        const Bits<W> m1 = m.shr(1), m7 = m.shr(7), m8 = m.shr(8);
        return (m & hol).popc() + (m & m1).popc() + (m & m7 & K.satf).popc() + (m & m8 & K.satf).popc() +
               (m1 & m7 & K.satf).popc() + (m1 & m8 & K.satf).popc();
    }
    const int S = K.S;
    const Bits<W> m1 = m.shr(1), mS = m.shr(S), m7 = m.shr(7 * S), m8 = m.shr(8 * S);
    int r = (m & hol).popc() + (m & unsk).popc() + (m & m1).popc();
    r += (m & m7 & K.satf).popc() + (m & m8 & K.satf).popc() + (mS & m7 & K.satf).popc() + (mS & m8 & K.satf).popc();
    r += (m & m1 & K.sd1).popc() + (m & m.shr(2) & K.sd2).popc();  // same-day pairs (S <= 3)
    return r;
}

__device__ __forceinline__ void es_hist16_inc(uint16_t* hist2, int idx) {  // 16-bit bin through its 32-bit word
    atomicAdd((unsigned int*)hist2 + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
}

// owner index of a present employee = rank of its first slot among the first slots
template <int W>
__device__ __forceinline__ int es_slot_of(const EsSmemT<W>& s, const Bits<W>& m) {
    return s.fmask->rank_below(m.ffs());
}

// Tallies from the slot masks (must be current): count histograms, occupancy sets, present /
// distinct counters, the first-slot mask.  Work is per SLOT (<= T threads busy), never per
// employee.  SCORE additionally accumulates the full (hard, S1) into misc[ES_HARD/ES_SOFT].
template <int W, bool MULTI, bool SCORE>
__device__ void es_tally(const EsSmemT<W>& s, const EsConstT<W>& K, const Bits<W>* __restrict__ hol,
                         const Bits<W>* __restrict__ unsk) {
    typedef EsDim<W> Dm;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = tid; k < (5 * Dm::CBINS + Dm::TBINS + Dm::WBINS) / 2; k += nt) ((unsigned int*)s.hist2)[k] = 0;
    if (tid < 5) s.occ2[tid] = 0;
    if (tid < 16 && tid != ES_BCAST && tid != ES_BCAST + 1) s.misc[tid] = 0;
    if (tid == 0) {
        *s.occT = Bits<Dm::OW>::zero();
        *s.occW = 0ull;
        *s.fmask = Bits<W>::zero();
    }
    __syncthreads();
    for (int d = tid; d < K.T; d += nt) {
        const int e = s.a[d];
        const Bits<W> m = s.mask[e];
        if (m.ffs() != d) continue;  // not the employee's first slot
        bits_atomic_or(s.fmask, d);
        const int t = m.popc(), w = (m & K.wkend).popc();
        es_hist16_inc(s.hist2, (int)(s.histT - s.hist2) + t);
        es_hist16_inc(s.hist2, (int)(s.histW - s.hist2) + w);
        if (t < 64 * Dm::OW) bits_atomic_or(s.occT, t);
        bits_atomic_or32((void*)s.occW, w);
        atomicAdd(&s.misc[ES_PRESENT], 1);
        atomicAdd(&s.misc[ES_SAME], t * (t - 1) / 2);  // slot pairs held by one employee (identity swaps)
#pragma unroll
        for (int wd = 0; wd < 5; ++wd) {
            const int c = (m & K.wd[wd]).popc();
            if (!c) continue;
            es_hist16_inc(s.hist2, wd * Dm::CBINS + c);
            atomicOr(&s.occ2[wd], 1u << c);
            atomicAdd(&s.misc[ES_DISTINCT0 + wd], 1);
        }
        if (SCORE) {  // H1..H3 (+X1, X2) pairs + H4 (14-day windows with count > 3); S1 (7-day windows, count > 2)
            const EsWin<W> win = es_windows<W, MULTI>(s, K, m);
            const int eh = es_pair_terms<W, MULTI>(m, hol[e], MULTI ? unsk[e] : Bits<W>::zero(), K) + win.viol14;
            atomicAdd(&s.misc[ES_HARD], eh);
            atomicAdd(&s.misc[ES_SOFT], win.viol7);
        }
    }
    __syncthreads();
}

// Build masks + tallies from a[] and the full score from them.  Block-cooperative.  The mask
// table must be all-zero on entry (es_clear_masks after the previous chain).
template <int W>
__device__ __forceinline__ void es_zero_masks(const EsSmemT<W>& s, int E) {
    u64* m = (u64*)s.mask;
    for (int k = threadIdx.x; k < E * W; k += blockDim.x) m[k] = 0ull;
}
// clear exactly the entries the current a[] set (<= T stores instead of E)
template <int W>
__device__ __forceinline__ void es_clear_masks(const EsSmemT<W>& s, int T) {
    for (int d = threadIdx.x; d < T; d += blockDim.x) s.mask[s.a[d]] = Bits<W>::zero();
}
template <int W, bool MULTI>
__device__ void es_build(const EsSmemT<W>& s, const EsConstT<W>& K, const Bits<W>* __restrict__ hol,
                         const Bits<W>* __restrict__ unsk, int& hard, int& soft) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int d = tid; d < K.T; d += nt) bits_atomic_or(&s.mask[s.a[d]], d);
    __syncthreads();
    es_tally<W, MULTI, true>(s, K, hol, unsk);
    hard = s.misc[ES_HARD];
    soft = s.misc[ES_SOFT];
    const int present = s.misc[ES_PRESENT];
    for (int wd = 0; wd < 5; ++wd) soft += es_s2_term(s.occ2[wd], s.misc[ES_DISTINCT0 + wd]);
    soft += es_spread(*s.occT, present) + es_spread64(*s.occW, present);
    __syncthreads();
}

// bit d of a set that lives in shared memory: with several words, one 32-bit load of the half-word that
// holds it instead of loading the whole set and selecting the word
template <int W>
__device__ __forceinline__ int es_smem_bit(const Bits<W>* p, int d) {
    if (W == 1) return (int)p->test(d);
    return (int)((((const unsigned int*)p)[d >> 5] >> (d & 31)) & 1u);
}

// receiver-side availability bits of slot d for an owner: holiday (+ missing skill)
template <int W, bool MULTI>
__device__ __forceinline__ int es_avoid_bits(const EsSmemT<W>& s, int slot, int d) {
    int r = es_smem_bit(&s.shol[slot], d);
    if (MULTI) r += es_smem_bit(&s.sunsk[slot], d);
    return r;
}

// Once per chain-step (masks + tallies must be current).  Everything is per slot or per
// (slot, value in use): no loop over the employee table.
template <int W, bool MULTI, int WD = W, bool PK = false>
__device__ void es_prepare(const EsSmemT<W>& s, const EsConstT<W>& K, const Bits<W>* __restrict__ hol,
                           const Bits<W>* __restrict__ unsk) {
    typedef EsDim<W> Dm;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int T = K.T;
    const Bits<W> fm = *s.fmask;
    const Bits<Dm::OW> occT = *s.occT;
    const u64 occW = *s.occW;
    // phase 1: owners (one per first slot), their window masks; per-slot counts of the slot's employee
    for (int d = tid; d < T; d += nt) {
        const int e = s.a[d];
        const Bits<W> m = s.mask[e];
        const int f = m.ffs();
        const int slot = fm.rank_below(f);
        s.dslot[d] = (unsigned char)slot;
        const int t = m.popc(), w = (m & K.wkend).popc();
        s.dayb[d] = (unsigned char)t;
        s.dayb[s.dp + d] = (unsigned char)w;
        s.dayb[2 * s.dp + d] = (unsigned char)(m & s.wdm[d]).popc();
        if (f == d) {
            s.semp[slot] = (uint16_t)e;
            s.smask[slot] = m;
            s.shol[slot] = hol[e];
            if (MULTI) s.sunsk[slot] = unsk[e];
            s.srk[2 * slot] = (unsigned char)occT.rank_below(t);
            s.srk[2 * slot + 1] = (unsigned char)__popcll(occW & ((1ull << w) - 1ull));
            if (W > 1) {
                u64 pk = 0;  // five byte counts in one 64-bit store
#pragma unroll
                for (int k = 0; k < 5; ++k) pk |= (u64)(m & K.wd[k]).popc() << (8 * k);
                *(u64*)(s.swd + 8 * slot) = pk;
            }
            const EsWin<W> win = es_windows<W, MULTI>(s, K, m);
            const EsEq<W> q = es_eq(s, slot);
            if (PK) {
                q[0].w[0] = win.eq3_14.w[0] | (win.eq2_7.w[0] << 32);
                q[1].w[0] = win.eq4_14.w[0] | (win.eq3_7.w[0] << 32);
            } else {
                q[0] = win.eq3_14;
                q[1] = win.eq4_14;
                q[2] = win.eq2_7;
                q[3] = win.eq3_7;
            }
        }
    }
    // rank -> value lists of the total / weekend counts in use (j-th set bit), and per weekday the
    // counts in use (0 first): the memo tables below are built for exactly these values.  One thread
    // per candidate VALUE: a value in use lands at its rank (the number of smaller values in use).
    for (int q = nt - 1 - tid; q < Dm::TBINS + Dm::WBINS + 5 * Dm::CBINS + 5; q += nt) {  // the last threads first
        if (q < Dm::TBINS) {
            if (q < 64 * Dm::OW && occT.test(q)) s.val[occT.rank_below(q)] = (unsigned char)q;
        } else if (q < Dm::TBINS + Dm::WBINS) {
            const int c = q - Dm::TBINS;
            if ((occW >> c) & 1ull) s.val[Dm::TCOLS + __popcll(occW & ((1ull << c) - 1ull))] = (unsigned char)c;
        } else if (q < Dm::TBINS + Dm::WBINS + 5 * Dm::CBINS) {
            const int k = q - Dm::TBINS - Dm::WBINS, wd = k / Dm::CBINS, c = k - wd * Dm::CBINS;
            const unsigned int bits = (s.occ2[wd] & Dm::CMASK) | 1u;  // counts >= 1 in use, and 0 (a newcomer to the weekday)
            if ((bits >> c) & 1u) s.val[Dm::VAL_C2 + wd * Dm::CBINS + __popc(bits & ((1u << c) - 1u))] = (unsigned char)c;
        } else {
            const int wd = q - Dm::TBINS - Dm::WBINS - 5 * Dm::CBINS;
            s.val[Dm::VAL_N2 + wd] = (unsigned char)__popc((s.occ2[wd] & Dm::CMASK) | 1u);
        }
    }
    if (tid == 0) s.misc[ES_NSLOT] = fm.popc();
    __syncthreads();
    // phase 2: what the slot's current employee loses, the value of an absent receiver, and the
    // memo tables.  The soft deltas of a change move depend on the receiving employee only
    // through its count on the weekday (S2), its total (S3) and its weekend count (S4).
    for (int d0 = 0; d0 < s.dp; d0 += nt) {  // trip count uniform per warp: the quad shuffles below need whole warps
        const int d = d0 + tid;
        unsigned int bw = ES_W_PAD;
        if (d < T) {
            const int slot = s.dslot[d];
            const Bits<W> m = s.smask[slot];
            const EsEq<W> q = es_eq(s, slot);
            int lossH = es_avoid_bits<W, MULTI>(s, slot, d) + (m & s.part[d]).popc();
            if (MULTI) lossH += (m & s.partx[d]).popc();
            int lossS;
            if (PK) {
                const u64 x = q[1].w[0] & s.cont14[d].w[0];
                lossH += es_popc_lo(x);
                lossS = es_popc_hi(x);
            } else {
                lossH += (bits_lo<WD>(q[1]) & bits_lo<WD>(s.cont14[d])).popc();
                lossS = (bits_lo<WD>(q[3]) & bits_lo<WD>(s.cont7[d])).popc();
            }
            s.base[d] = ((unsigned)(0x8000 - lossH) << 16) | (unsigned)(0x8000 - lossS);
            // an absent receiver: no pairs, no window counts, zero slots anywhere
            const int to = s.dayb[d], wo = s.dayb[s.dp + d], wd = s.dwd[d];
            const int isw = wd >= 5 ? 1 : 0;
            int ds = -lossS + es_s3_delta(s, to, 0) + es_s4_delta(s, to, wo, isw, 0, false);
            if (wd < 5) ds += es_s2_delta(s, wd, (int)s.dayb[2 * s.dp + d], 0);
            bw = ((unsigned)(64 - lossH) << ES_W_DH) | ((unsigned)(512 + ds) << ES_W_DS) | (unsigned)d;
        }
        if (d < s.dp) s.baseW[d] = bw;
        if (!MULTI) {
            // pass B's table: for the quad of slots 4q..4q+3 and every 4-bit holiday pattern k, the four
            // absent-receiver values with the holiday addend already in (one 128-bit load per 4 candidates).
            // The quad's four lanes are neighbours; each writes 4 of the 16 patterns.
            constexpr unsigned int HB = 1u << ES_W_DH;
            const int lane = tid & 31, q0 = lane & ~(ES_BQ - 1), r = lane & (ES_BQ - 1);
            unsigned int b[ES_BQ];
#pragma unroll
            for (int j = 0; j < ES_BQ; ++j) b[j] = __shfl_sync(0xffffffffu, bw, q0 + j);
            if (d < s.dp) {  // the group's lanes are neighbours; each writes 2^BQ / BQ of the patterns
                unsigned int* row = s.bwq + (d / ES_BQ) * ((1 << ES_BQ) * ES_BQ);
                constexpr int PER = (1 << ES_BQ) / ES_BQ;
#pragma unroll
                for (int kk = 0; kk < PER; ++kk) {
                    const int k = r * PER + kk;
#pragma unroll
                    for (int j = 0; j < ES_BQ; ++j) row[k * ES_BQ + j] = b[j] + (((k >> j) & 1) ? HB : 0u);
                }
            }
        }
    }
    // dense loops: only (slot, value in use) pairs, so every lane of a warp has work
    {
        int nmax = 1;
#pragma unroll
        for (int wd = 0; wd < 5; ++wd) nmax = max(nmax, (int)s.val[Dm::VAL_N2 + wd]);
        for (int k = tid; k < T * nmax; k += nt) {
            const int d = k / nmax, j = k - d * nmax;
            const int wd = s.dwd[d];
            if (wd >= 5) {
                if (j == 0) s.s2t[d * Dm::CBINS] = 0;  // weekend slots: the only entry ever looked up
            } else if (j < (int)s.val[Dm::VAL_N2 + wd]) {
                const int cn = s.val[Dm::VAL_C2 + wd * Dm::CBINS + j];
                s.s2t[d * Dm::CBINS + cn] = (signed char)es_s2_delta(s, wd, (int)s.dayb[2 * s.dp + d], cn);
            }
        }
    }
    {
        const int nT = occT.popc();
        for (int k = tid; k < T * nT; k += nt) {
            const int d = k / nT, j = k - d * nT;
            s.s3t[d * Dm::TCOLS + j] = (typename Dm::s3_t)es_s3_delta(s, s.dayb[d], (int)s.val[j]);
        }
        const int nW = __popcll(occW);
        for (int k = tid; k < T * nW; k += nt) {
            const int d = k / nW, j = k - d * nW;
            const int wn = s.val[Dm::TCOLS + j];
            const int isw = s.dwd[d] >= 5 ? 1 : 0, wo = s.dayb[s.dp + d];
            s.s4t[d * Dm::WCOLS + j] = (signed char)es_s4_delta(s, s.dayb[d], wo, isw, wn, true);
            if (isw) {  // swap with a weekday slot of an employee holding wn weekend slots: nobody joins or leaves
                const int present = s.misc[ES_PRESENT];
                const u64 occ = es_occ_move64(s.histW, occW, wo, wn, wo - 1, wn + 1);
                s.s4s[d * Dm::WCOLS + j] = (signed char)(es_spread64(occ, present) - es_spread64(occW, present));
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
// absent-receiver packing w = (dh + 64) << 18 | (ds + 512) << 8 | slot  ->  v
// (dh in [-30, 2], ds in [-300, 300]: bounded by the window / weekday / spread ranges for T <= 192)
__device__ __forceinline__ unsigned int es_w_to_v(unsigned int w) {
    return ((unsigned)(0x8000 - 64 + (int)(w >> ES_W_DH)) << 16) |
           (unsigned)(0x8000 - 512 + (int)((w >> ES_W_DS) & 0x3ffu));
}

// change: slot d goes to the PRESENT employee of owner `slot` (not the slot's current one).  ga = the
// receiver-side parts (hard gain, S1 gain, S2 delta) packed for the swap pass.
// WD: words that hold the DAY-indexed sets (window starts: cont14 / cont7 / the "count == k" planes).  With
// several shifts per day a rota of T <= 64 W slots has only T / S days, so those sets are empty above word
// WD = ceil(D / 64) and their popcounts run on WD words instead of W.
template <int W, bool MULTI, int WD = W, bool PK = false>
__device__ __forceinline__ unsigned int es_change_present_v(const EsSmemT<W>& s, int d, int slot, unsigned int& ga) {
    typedef EsDim<W> Dm;
    const Bits<W> m = s.smask[slot];
    const EsEq<W> q = es_eq(s, slot);
    int gh = es_avoid_bits<W, MULTI>(s, slot, d) + (m & s.part[d]).popc();
    if (MULTI) gh += (m & s.partx[d]).popc();
    int gs;
    if (PK) {
        const u64 x = q[0].w[0] & s.cont14[d].w[0];
        gh += es_popc_lo(x);
        gs = es_popc_hi(x);
    } else {
        gh += (bits_lo<WD>(q[0]) & bits_lo<WD>(s.cont14[d])).popc();
        gs = (bits_lo<WD>(q[2]) & bits_lo<WD>(s.cont7[d])).popc();
    }
    // the receiver's slots on d's weekday: a table lookup instead of a W-word popcount
    const int cn = W > 1 ? (int)s.swd[8 * slot + (int)s.dwd[d]] : (m & s.wdm[d]).popc();
    const int s2 = (int)s.s2t[d * Dm::CBINS + cn];
    ga = (unsigned)gh | ((unsigned)gs << 5) | ((unsigned)(s2 + 32) << 8);
    return s.base[d] + ((unsigned)gh << 16) +
           (unsigned)(gs + s2 + (int)s.s3t[d * Dm::TCOLS + (int)s.srk[2 * slot]] +
                      (int)s.s4t[d * Dm::WCOLS + (int)s.srk[2 * slot + 1]]);
}
template <int W, bool MULTI>
__device__ __forceinline__ unsigned int es_change_present_v(const EsSmemT<W>& s, int d, int slot) {
    unsigned int ga;
    return es_change_present_v<W, MULTI>(s, d, slot, ga);
}

// change: slot d goes to an ABSENT employee with holiday mask hol (and unskilled mask unsk)
template <int W, bool MULTI>
__device__ __forceinline__ unsigned int es_change_absent_v(const EsSmemT<W>& s, int d, const Bits<W>& hol,
                                                           const Bits<W>& unsk) {
    unsigned int bits = (unsigned)hol.test(d);
    if (MULTI) bits += (unsigned)unsk.test(d);
    return es_w_to_v(s.baseW[d] + (bits << ES_W_DH));
}

// swap: slots d1 < d2 exchange employees (different).
template <int W, bool MULTI>
__device__ __forceinline__ unsigned int es_swap_v(const EsSmemT<W>& s, const EsConstT<W>& K, int d1, int d2) {
    typedef EsDim<W> Dm;
    const int s1 = s.dslot[d1], s2 = s.dslot[d2];
    const Bits<W> nb1 = ~Bits<W>::bit(d1), nb2 = ~Bits<W>::bit(d2);
    const Bits<W> m1 = s.smask[s1], m2 = s.smask[s2];
    const EsEq<W> q1 = es_eq(s, s1), q2 = es_eq(s, s2);
    // windows holding exactly one of the two slots' days change count by one for each employee
    const Bits<W> c14a = s.cont14[d1], c14b = s.cont14[d2], c7a = s.cont7[d1], c7b = s.cont7[d2];
    const Bits<W> only14a = c14a & ~c14b, only14b = c14b & ~c14a, only7a = c7a & ~c7b, only7b = c7b & ~c7a;
    int dh = es_avoid_bits<W, MULTI>(s, s1, d2) - es_avoid_bits<W, MULTI>(s, s1, d1) +
             es_avoid_bits<W, MULTI>(s, s2, d1) - es_avoid_bits<W, MULTI>(s, s2, d2);
    // H2/H3 pairs: e1 leaves d1 and lands on d2 (its other slots: m1 without d1), e2 the reverse
    dh += ((m1 & nb1) & s.part[d2]).popc() - (m1 & s.part[d1]).popc();
    dh += ((m2 & nb2) & s.part[d1]).popc() - (m2 & s.part[d2]).popc();
    if (MULTI) {  // same-day pairs
        dh += ((m1 & nb1) & s.partx[d2]).popc() - (m1 & s.partx[d1]).popc();
        dh += ((m2 & nb2) & s.partx[d1]).popc() - (m2 & s.partx[d2]).popc();
    }
    // H4: e1 loses a slot in windows with only d1's day (count 4 -> 3), gains in windows with only d2's
    dh += (q1[0] & only14b).popc() - (q1[1] & only14a).popc();
    dh += (q2[0] & only14a).popc() - (q2[1] & only14b).popc();
    int ds = (q1[2] & only7b).popc() - (q1[3] & only7a).popc();
    ds += (q2[2] & only7a).popc() - (q2[3] & only7b).popc();
    const int wd1 = s.dwd[d1], wd2 = s.dwd[d2];
    if (wd1 != wd2) {
        // two independent transfers on distinct weekday histograms: slot d1 goes e1 -> e2, slot d2
        // goes e2 -> e1; each is the memoised change-move S2 delta for the receiver's count
        // (weekend rows of s2t are zero)
        ds += (int)s.s2t[d1 * Dm::CBINS + (m2 & s.wdm[d1]).popc()] + (int)s.s2t[d2 * Dm::CBINS + (m1 & s.wdm[d2]).popc()];
        // totals (S3) unchanged; a weekend slot and a weekday slot trade places (S4)
        if (wd1 >= 5 && wd2 < 5) ds += (int)s.s4s[d1 * Dm::WCOLS + (int)s.srk[2 * s2 + 1]];
        if (wd2 >= 5 && wd1 < 5) ds += (int)s.s4s[d2 * Dm::WCOLS + (int)s.srk[2 * s1 + 1]];
    }
    (void)K;
    return ((unsigned)(0x8000 + dh) << 16) | (unsigned)(0x8000 + ds);
}

// The same value from pass A's table (s.ga must be complete): a swap is two simultaneous
// transfers, slot d1 -> e2 and slot d2 -> e1; their tabulated gains/losses are exact except where
// both slots meet -- the H2/H3 (and same-day) pair (d1, d2) itself and the windows holding BOTH
// days, whose counts do not change.
template <int W, bool MULTI, int WD = W, bool PK = false>
__device__ __forceinline__ unsigned int es_swap_from_table(const EsSmemT<W>& s, int d1, int d2) {
    typedef EsDim<W> Dm;
    const int s1 = s.dslot[d1], s2 = s.dslot[d2];
    const unsigned int g12 = s.ga[d1 * s.ns + s2], g21 = s.ga[d2 * s.ns + s1];  // d1 -> e2, d2 -> e1
    const unsigned int b1 = s.base[d1], b2 = s.base[d2];
    int dh = (int)(g12 & 31u) + (int)(g21 & 31u) + (int)(b1 >> 16) + (int)(b2 >> 16) - 0x10000;
    int ds = (int)((g12 >> 5) & 7u) + (int)((g21 >> 5) & 7u) + (int)(b1 & 0xffffu) + (int)(b2 & 0xffffu) - 0x10000;
    const int gap = MULTI ? (int)s.sday[d2] - (int)s.sday[d1] : d2 - d1;
    if (gap < 14) {
        if (W == 1) {
            if (s.part[d1].test(d2)) dh -= 2;
            if (MULTI && s.partx[d1].test(d2)) dh -= 2;
        } else {
            dh -= 2 * es_smem_bit(&s.part[d1], d2);
            if (MULTI) dh -= 2 * es_smem_bit(&s.partx[d1], d2);
        }
        const EsEq<W> q1 = es_eq(s, s1), q2 = es_eq(s, s2);
        if (PK) {
            const u64 both = s.cont14[d1].w[0] & s.cont14[d2].w[0];  // the windows (both lengths) holding both days
            if (both) {
                const u64 l1 = q1[1].w[0] & both, g1 = q1[0].w[0] & both, l2 = q2[1].w[0] & both, g2 = q2[0].w[0] & both;
                dh += es_popc_lo(l1) - es_popc_lo(g1) + es_popc_lo(l2) - es_popc_lo(g2);
                ds += es_popc_hi(l1) - es_popc_hi(g1) + es_popc_hi(l2) - es_popc_hi(g2);
            }
        } else {
            const Bits<WD> both14 = bits_lo<WD>(s.cont14[d1]) & bits_lo<WD>(s.cont14[d2]);
            if (both14.any())
                dh += (bits_lo<WD>(q1[1]) & both14).popc() - (bits_lo<WD>(q1[0]) & both14).popc() +
                      (bits_lo<WD>(q2[1]) & both14).popc() - (bits_lo<WD>(q2[0]) & both14).popc();
            const Bits<WD> both7 = bits_lo<WD>(s.cont7[d1]) & bits_lo<WD>(s.cont7[d2]);
            if (both7.any())
                ds += (bits_lo<WD>(q1[3]) & both7).popc() - (bits_lo<WD>(q1[2]) & both7).popc() +
                      (bits_lo<WD>(q2[3]) & both7).popc() - (bits_lo<WD>(q2[2]) & both7).popc();
        }
    }
    const int wd1 = s.dwd[d1], wd2 = s.dwd[d2];
    if (wd1 != wd2) {
        ds += (int)((g12 >> 8) & 63u) + (int)((g21 >> 8) & 63u) - 64;  // the two S2 transfers (weekend rows are 0)
        if (wd1 >= 5 && wd2 < 5) ds += (int)s.s4s[d1 * Dm::WCOLS + (int)s.srk[2 * s2 + 1]];
        if (wd2 >= 5 && wd1 < 5) ds += (int)s.s4s[d2 * Dm::WCOLS + (int)s.srk[2 * s1 + 1]];
    }
    return ((unsigned)(0x8000 + dh) << 16) | (unsigned)(0x8000 + ds);
}

// move id: change (d, e) -> d*E + e ; swap (d1<d2) -> T*E + tri(d1,d2)
__device__ __forceinline__ int es_tri_index(int T, int d1, int d2) {
    return d1 * T - d1 * (d1 + 1) / 2 + (d2 - d1 - 1);
}

// key = v << 24 | move id  (id < 2^24: 192 slots x 65535 employees + swaps)
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
//   A  change moves to PRESENT employees  (slot x owner; full mask arithmetic)
//   B  change moves to ABSENT employees   (thread per employee, loop over slots; the value is the
//      per-slot table entry plus the employee's holiday / skill bits, tracked with one min per candidate)
//   C  swaps
template <int W, bool MULTI, bool DUMP, int WD = W, bool PK = false>
__device__ __forceinline__ long long es_scan(const EsSmemT<W>& s, const EsConstT<W>& K,
                                             const Bits<W>* __restrict__ hol, const Bits<W>* __restrict__ unsk,
                                             const u64* __restrict__ cnt2,
                                             const uint16_t* __restrict__ scan, long long* dump_h,
                                             long long* dump_s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int T = K.T, E = K.E;
    const int nslot = s.misc[ES_NSLOT];
    long long key = ES_KEY_INF;
    {   // A
        const int nA = nslot * T;
        int d = tid / nslot, slot = tid - d * nslot;
        const int dd = nt / nslot, dsl = nt - dd * nslot;
        for (int k = tid; k < nA; k += nt) {
            const int id = d * E + (int)s.semp[slot];
            if ((int)s.dslot[d] != slot) {
                unsigned int ga;
                const unsigned int v = es_change_present_v<W, MULTI, WD, PK>(s, d, slot, ga);
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
        constexpr unsigned int HB = 1u << ES_W_DH;
        for (int e = tid; e < E; e += nt) {
            if (s.mask[e].any()) continue;
            unsigned int w = 0xffffffffu;
            if (MULTI) {
                // the holiday and missing-skill bits of a slot arrive as one 2-bit count (cnt2, built with the
                // handle): value = baseW + count << ES_W_DH, a mask and a multiply-add per candidate
                const u64* c2 = cnt2 + (size_t)e * (2 * W);
#pragma unroll
                for (int g = 0; g < 2 * W; ++g) {
                    const int lim = min(T - 32 * g, 32);
                    if (lim <= 0) break;
                    const u64 cw = c2[g];
                    for (int d0 = 0; d0 < lim; d0 += 4) {
                        const uint4 b = bw4[(32 * g + d0) >> 2];
                        const unsigned int x = (unsigned int)(cw >> (2 * d0));  // counts of slots d0..d0+3, 2 bits each
                        const unsigned int w0 = b.x + (x & 0x03u) * HB;
                        const unsigned int w1 = b.y + (x & 0x0cu) * (HB >> 2);
                        const unsigned int w2 = b.z + (x & 0x30u) * (HB >> 4);
                        const unsigned int w3 = b.w + (x & 0xc0u) * (HB >> 6);
                        w = __vimin3_u32(w, w0, w1);
                        w = __vimin3_u32(w, w2, w3);
                        if (DUMP) {
                            const unsigned int ww[4] = {w0, w1, w2, w3};
                            for (int j = 0; j < 4 && 32 * g + d0 + j < T; ++j) {
                                const unsigned int v = es_w_to_v(ww[j]);
                                dump_h[(32 * g + d0 + j) * E + e] = es_v_dh(v);
                                dump_s[(32 * g + d0 + j) * E + e] = es_v_ds(v);
                            }
                        }
                    }
                }
            } else {
                const Bits<W> h = hol[e];
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    const int lim = min(T - 64 * i, 64);
                    for (int d0 = 0; d0 < lim; d0 += 4) {
                        // the group's values for this employee's holiday pattern: one table row per load
                        unsigned int w0, w1, w2, w3;
                        const unsigned int x = (unsigned int)(h.w[i] >> d0);
                        if (ES_BQ == 4) {
                            const uint4 b = ((const uint4*)s.bwq)[((64 * i + d0) >> 2) * 16 + (x & 15u)];
                            w0 = b.x;
                            w1 = b.y;
                            w2 = b.z;
                            w3 = b.w;
                        } else {
                            const uint2* t2 = (const uint2*)s.bwq + ((64 * i + d0) >> 1) * 4;
                            const uint2 b0 = t2[x & 3u], b1 = t2[4 + ((x >> 2) & 3u)];
                            w0 = b0.x;
                            w1 = b0.y;
                            w2 = b1.x;
                            w3 = b1.y;
                        }
                        w = __vimin3_u32(w, w0, w1);
                        w = __vimin3_u32(w, w2, w3);
                        if (DUMP) {
                            const unsigned int ww[4] = {w0, w1, w2, w3};
                            for (int j = 0; j < 4 && 64 * i + d0 + j < T; ++j) {
                                const unsigned int v = es_w_to_v(ww[j]);
                                dump_h[(64 * i + d0 + j) * E + e] = es_v_dh(v);
                                dump_s[(64 * i + d0 + j) * E + e] = es_v_ds(v);
                            }
                        }
                    }
                }
            }
            if (w < bw) {  // equal value and slot: the lower employee index (seen first) stays
                bw = w;
                be = e;
            }
        }
        if (bw != 0xffffffffu) {
            const long long k2 = es_key(es_w_to_v(bw), (int)(bw & 255u) * E + be);
            key = k2 < key ? k2 : key;
        }
    }
    __syncthreads();  // pass A's table is complete
    {   // C
        const int n_change = T * E, n_swap = T * (T - 1) / 2;
        for (int r = tid; r < n_swap; r += nt) {
            const int dd = scan[r], d1 = dd >> 8, d2 = dd & 0xff;
            const int id = n_change + es_tri_index(T, d1, d2);
            if (s.dslot[d1] != s.dslot[d2]) {
                const unsigned int v = es_swap_from_table<W, MULTI, WD, PK>(s, d1, d2);
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
// orc_es_local_search_ref).  Reference rotas only (S = 1, no skill table).
struct EsCand {
    unsigned int v;  // packed (dhard, dsoft); 0xffffffff = none
    int kind, x, y;  // change: day x -> employee index y; swap: days x < y
};

__device__ __forceinline__ unsigned int es_ref_draw(unsigned long long seed, unsigned int chain, unsigned long long t) {
    const Philox4 b = philox_stream(seed, chain, 2u /* CS_PHILOX_LS */, t >> 2);
    return b.v[t & 3];
}

// candidate k of the endless stream; returns false for an identity candidate (tabu)
template <int W>
__device__ __forceinline__ bool es_ref_candidate(const EsSmemT<W>& s, int D, int E, unsigned long long seed,
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
template <int W>
__device__ __forceinline__ bool es_cand_less(const EsSmemT<W>& s, const EsCand& A, const EsCand& B) {
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
template <int W>
__device__ long long es_scan_ref(const EsSmemT<W>& s, const EsConstT<W>& K, const Bits<W>* __restrict__ hol,
                                 unsigned long long seed, unsigned int chain, unsigned long long window,
                                 unsigned long long max_draws, unsigned int& n_scored) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5, nw = (nt + 31) >> 5;
    const int D = K.T, E = K.E;
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
                const Bits<W> m = s.mask[c.y];
                c.v = m.any() ? es_change_present_v<W, false>(s, c.x, es_slot_of(s, m))
                              : es_change_absent_v<W, false>(s, c.x, hol[c.y], Bits<W>::zero());
            } else {
                c.v = es_swap_v<W, false>(s, K, c.x, c.y);
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

// per-slot constants of the handle: part | cont14 | cont7 | partx from the host table, weekday masks
template <int W, bool MULTI, bool PK = false>
__device__ __forceinline__ void es_load_consts(const EsSmemT<W>& s, const EsConstT<W>& K,
                                               const Bits<W>* __restrict__ slotc) {
    for (int d = threadIdx.x; d < s.dp; d += blockDim.x) {
        const int day = d / K.S;
        const int wd = (K.start_wd + day) % 7;
        s.part[d] = slotc[d];
        s.cont14[d] = slotc[s.dp + d];
        s.cont7[d] = slotc[2 * s.dp + d];
        if (PK) s.cont14[d].w[0] |= slotc[2 * s.dp + d].w[0] << 32;
        if (MULTI) s.partx[d] = slotc[3 * s.dp + d];
        s.wdm[d] = (d < K.T && wd < 5) ? K.wd[wd] : Bits<W>::zero();
        s.dwd[d] = (unsigned char)wd;
        s.sday[d] = (unsigned char)day;
    }
}

// ------------------------------------------------------------------ the step kernel (K5)
#define ES_LB_THREADS(W, MULTI) (((W) == 1 && !(MULTI)) ? 256 : 512)
#define ES_LB_BLOCKS(W, MULTI) (((W) == 1 && !(MULTI)) ? 4 : 1)

template <int W, bool MULTI, bool REF, int WD = W, bool PK = false>
__global__ void __launch_bounds__(ES_LB_THREADS(W, MULTI), ES_LB_BLOCKS(W, MULTI)) es_step_kernel(EsParamsT<W> p) {
    static_assert(!(PK && (REF || WD != 1)), "packed day sets: one day word, full-neighbourhood scan only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsConstT<W>& K = p.K;
    const EsSmemT<W> s = es_carve<W, MULTI>(smem_raw, K.T, K.E);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int T = K.T, E = K.E;
    const int n_change = T * E, n_swap = T * (T - 1) / 2;
    es_load_consts<W, MULTI, PK>(s, K, p.slotc);
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
        es_build<W, MULTI>(s, K, p.hol, p.unsk, hard, soft);
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
            es_prepare<W, MULTI, WD, PK>(s, K, p.hol, p.unsk);
            // non-identity candidates: every (slot, employee != current) + every slot pair held by
            // two different employees -- each of them is evaluated by es_scan
            long long key;
            if (REF && !p.dump_h) {  // the reference's sampled window instead of the whole neighbourhood
                unsigned int nsc = 0;
                key = es_scan_ref<W>(s, K, p.hol, p.seed, p.chain_offset + (unsigned)chain, p.window, p.max_draws, nsc);
                scored += nsc;
            } else {
                scored += (unsigned long long)(n_change - T) + (unsigned long long)(n_swap - s.misc[ES_SAME]);
                key = p.dump_h ? es_scan<W, MULTI, true, WD, PK>(s, K, p.hol, p.unsk, p.cnt2, p.tri_scan, p.dump_h, p.dump_s)
                               : es_scan<W, MULTI, false, WD, PK>(s, K, p.hol, p.unsk, p.cnt2, p.tri_scan, nullptr, nullptr);
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
                    const Bits<W> bit = Bits<W>::bit(d);
                    s.mask[eo] = s.mask[eo] & ~bit;
                    s.mask[e] = s.mask[e] | bit;
                    s.a[d] = (uint16_t)e;
                    kind = 0;
                    x = (unsigned)d;
                    y = (unsigned)e;
                } else {
                    const int dd = p.tri[id - n_change], d1 = dd >> 8, d2 = dd & 0xff;
                    const int e1 = s.a[d1], e2 = s.a[d2];
                    const Bits<W> x2 = Bits<W>::bit(d1) | Bits<W>::bit(d2);
                    s.mask[e1] = s.mask[e1] ^ x2;
                    s.mask[e2] = s.mask[e2] ^ x2;
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
            if (it + 1 < p.max_steps) es_tally<W, MULTI, false>(s, K, p.hol, p.unsk);  // tallies of the new state
        }
        __syncthreads();
        es_clear_masks(s, T);  // leave the mask table all-zero for the next chain
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
template <int W, bool MULTI>
__global__ void __launch_bounds__(ES_LB_THREADS(W, MULTI)) es_rescore_kernel(EsParamsT<W> p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsSmemT<W> s = es_carve<W, MULTI>(smem_raw, p.K.T, p.K.E);
    es_load_consts<W, MULTI>(s, p.K, p.slotc);
    es_zero_masks(s, p.K.E);
    for (int local = blockIdx.x; local < p.n_chains; local += gridDim.x) {
        const int chain = p.first_chain + local;
        __syncthreads();
        for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
            s.a[k] = p.a[(size_t)chain * p.stride + k];
        __syncthreads();
        int hard, soft;
        es_build<W, MULTI>(s, p.K, p.hol, p.unsk, hard, soft);
        es_clear_masks(s, p.K.T);
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

// explicit-move deltas against one chain (parity hook); kind 0 change (x=slot,y=employee idx),
// 1 swap (x,y = slots)
template <int W, bool MULTI>
__global__ void __launch_bounds__(256) es_eval_kernel(EsParamsT<W> p, int chain, int kind, const uint2* __restrict__ moves,
                               unsigned long long n_moves, long long* __restrict__ dh_out,
                               long long* __restrict__ ds_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EsSmemT<W> s = es_carve<W, MULTI>(smem_raw, p.K.T, p.K.E);
    for (int k = threadIdx.x; k < p.stride; k += blockDim.x)
        s.a[k] = p.a[(size_t)chain * p.stride + k];
    es_load_consts<W, MULTI>(s, p.K, p.slotc);
    es_zero_masks(s, p.K.E);
    __syncthreads();
    int hard, soft;
    es_build<W, MULTI>(s, p.K, p.hol, p.unsk, hard, soft);
    es_prepare<W, MULTI>(s, p.K, p.hol, p.unsk);
    for (unsigned long long k = threadIdx.x; k < n_moves; k += blockDim.x) {
        const uint2 mv = moves[k];
        unsigned int v = 0;
        bool ok;
        if (kind == 0) {
            const int d = (int)mv.x, en = (int)mv.y;
            ok = s.a[d] != en;
            if (ok) {
                const Bits<W> m = s.mask[en];
                v = m.any() ? es_change_present_v<W, MULTI>(s, d, es_slot_of(s, m))
                            : es_change_absent_v<W, MULTI>(s, d, p.hol[en], MULTI ? p.unsk[en] : Bits<W>::zero());
            }
        } else {
            const int d1 = (int)min(mv.x, mv.y), d2 = (int)max(mv.x, mv.y);
            ok = d1 != d2 && s.a[d1] != s.a[d2];
            if (ok) v = es_swap_v<W, MULTI>(s, p.K, d1, d2);
        }
        dh_out[k] = ok ? (long long)es_v_dh(v) : INT64_MAX;
        ds_out[k] = ok ? (long long)es_v_ds(v) : INT64_MAX;
    }
}

// K6: full re-score straight from a[] by the reference's own loops (no masks, no histograms) in
// slot units -- one thread per chain; cross-checks the tally formulation.  out10: H1..H4, S1..S4,
// X1 (same-day overlap), X2 (skill); S = 1 without a skill table leaves X1 = X2 = 0.
// holiday: [E][D] bytes would be large; holidays come as the per-employee slot masks (words).
__global__ void es_full_score_kernel(const uint16_t* __restrict__ a, int stride, int n_chains, int D, int S,
                                     int start_wd, const u64* __restrict__ hol, const u64* __restrict__ unsk,
                                     int words, long long* out10) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const uint16_t* x = a + (size_t)chain * stride;
    const int T = D * S;
    long long t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto wday = [&](int slot) { return (start_wd + slot / S) % 7; };
    for (int d = 0; d < T; ++d) {
        t[0] += (hol[(size_t)x[d] * words + (d >> 6)] >> (d & 63)) & 1ull;               // lib.rs:273-280
        if (unsk) t[9] += (unsk[(size_t)x[d] * words + (d >> 6)] >> (d & 63)) & 1ull;     // skill (extension)
    }
    for (int i = 0; i + 2 <= T; ++i) t[1] += (x[i] == x[i + 1]);                          // :286-292
    for (int i = 0; i + 9 <= D; ++i) {                                                    // :295-315
        if ((start_wd + i) % 7 != 5) continue;
        for (int sh = 0; sh < S; ++sh) {
            const int p = i * S + sh, q = (i + 1) * S + sh, r = (i + 7) * S + sh, u = (i + 8) * S + sh;
            t[2] += (x[p] == x[r]) + (x[p] == x[u]) + (x[q] == x[r]) + (x[q] == x[u]);
        }
    }
    for (int len = 14; len >= 7; len -= 7) {                                              // :318-339
        const int limit = len == 14 ? 3 : 2;
        for (int w = 0; w + len <= D; ++w) {
            const int lo = w * S, hi = (w + len) * S;
            for (int q = lo; q < hi; ++q) {
                bool first = true;
                for (int r = lo; r < q; ++r)
                    if (x[r] == x[q]) first = false;
                if (!first) continue;
                int c = 0;
                for (int r = q; r < hi; ++r) c += (x[r] == x[q]);
                if (c > limit) t[len == 14 ? 3 : 4] += 1;
            }
        }
    }
    for (int wd = 0; wd < 5; ++wd) {                                                      // :194-218
        int distinct = 0, minc = 1 << 30;
        for (int i = 0; i < T; ++i) {
            if (wday(i) != wd) continue;
            bool first = true;
            for (int q = 0; q < i; ++q)
                if (wday(q) == wd && x[q] == x[i]) first = false;
            if (!first) continue;
            int c = 0;
            for (int q = i; q < T; ++q) c += (wday(q) == wd && x[q] == x[i]);
            ++distinct;
            minc = c < minc ? c : minc;
        }
        if (distinct >= 2) t[5] += minc;
    }
    int present = 0, mind = 1 << 30, maxd = -1, minw = 1 << 30, maxw = -1;               // :345-365
    for (int i = 0; i < T; ++i) {
        bool first = true;
        for (int q = 0; q < i; ++q)
            if (x[q] == x[i]) first = false;
        if (!first) continue;
        int days = 0, wk = 0;
        for (int q = i; q < T; ++q)
            if (x[q] == x[i]) {
                ++days;
                const int w = wday(q);
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
    for (int d = 0; d < D; ++d)                                                           // same-day overlap (extension)
        for (int s1 = 0; s1 < S; ++s1)
            for (int s2 = s1 + 1; s2 < S; ++s2) t[8] += (x[d * S + s1] == x[d * S + s2]);
    for (int k = 0; k < 10; ++k) out10[(size_t)chain * 10 + k] = t[k];
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

// best chain by lexicographic (hard, soft); the key packs (hard << 48) | (soft << 32) | global
// chain id for the min-allreduce
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
