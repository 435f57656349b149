/*
 * cs_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT THE PRODUCT).
 *
 * Plain-C restatement of the move-evaluation hot path of asimihsan/constraint-solver
 * (Rust; cannot be compiled in this image: no cargo/rustc).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).
 *
 * Who may use this: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs -- as the CHECKER or the timed CPU baseline, never as a
 * product code path.  Nothing under constraint_solver_b200/ links or loads it.
 *
 * Parity pinning:
 *   - n-queens score: PINNED by the reference's own known-answer tests
 *     (examples/nqueens/src/lib.rs:94-105 and :108-119) -- see tests/golden/nq_kat.json.
 *   - employee-scheduling score: the reference has NO test for it => "parity unpinned"
 *     by reference tests; pinned only by the hand-derived vectors of SURVEY.md section 8(c)
 *     (tests/golden/es_kat.json).
 *   - random streams (rand 0.8.5 / rand_chacha 0.3.1 are un-vendored crates.io
 *     dependencies pinned in Cargo.lock): NOT reproduced; the oracle draws from
 *     Philox4x32-10 (Salmon et al., SC'11) instead.  "parity unpinned" at the RNG boundary.
 */
#ifndef CS_ORACLE_H
#define CS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Philox4x32-10 (published algorithm; Random123 known-answer vectors in tests) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* Stream convention shared with the product (include/cs_b200.h):
 * key = {seed lo, seed hi}, ctr = {counter lo, counter hi, chain, purpose}. */
void orc_philox_stream(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t counter,
                       uint32_t out[4]);
/* t-th 32-bit draw of a stream (block t/4, word t%4). */
uint32_t orc_philox_draw(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t t);

/* ---- n-queens ---- */
/* get_col_scores: examples/nqueens/src/lib.rs:74-87 (O(n^2) pair loop). */
void orc_nq_col_scores(const int64_t* rows, int64_t n, int64_t* out);
/* get_scored_solution: examples/nqueens/src/lib.rs:126-140 (sum of col scores). */
int64_t orc_nq_score(const int64_t* rows, int64_t n);
/* generate_initial_solution: examples/nqueens/src/lib.rs:152-161 (shuffle of 0..n),
 * shuffle restated as Fisher-Yates over Philox purpose 0: for k=n-1..1: swap(k, mulhi(u32,k+1)). */
void orc_nq_init_perm(uint64_t seed, uint32_t chain, int64_t n, int64_t* rows);

enum { ORC_NQ_SWAP = 0, ORC_NQ_CHANGE = 1 };
enum { ORC_TIE_MOVE_ORDER = 0, ORC_TIE_REFERENCE = 1 };

/* Score every candidate of the full neighbourhood the reference's way: clone the
 * solution, apply the move, full re-score (local_search.rs:315-322).  Writes the
 * candidate's score minus the current score into delta[] in enumeration order:
 * SWAP: (i,j), i<j ascending, j inner; CHANGE: (c,v), c outer, v inner.
 * Identity candidates (tabu, local_search.rs:155-199,319) get delta = INT64_MAX.
 * Returns the number of entries written (n(n-1)/2 or n*n). */
int64_t orc_nq_neighbourhood_deltas(const int64_t* rows, int64_t n, int kind, int64_t* delta);

/* Clone + full re-score of explicit moves (a[k], b[k]) of the given kind. */
void orc_nq_eval_moves(const int64_t* rows, int64_t n, int kind, const int64_t* a,
                       const int64_t* b, int64_t n_moves, int64_t* delta);

/* LocalSearch::execute, local-search/src/local_search.rs:301-342, with the FULL swap or
 * change neighbourhood as the move proposer and tabu == {current} (identity filter).
 * window_size==0 means unlimited.  tie = ORC_TIE_MOVE_ORDER picks the first minimal
 * candidate in enumeration order (the device rule); ORC_TIE_REFERENCE sorts by
 * (score, solution lexicographic) like the derived Ord (local_search.rs:29-37,323).
 * rows: in = start, out = best_solution.  current_out (optional) = last current.
 * trace arrays (optional, capacity cap): chosen move and score after each accepted step.
 * Returns the number of accepted steps. */
int64_t orc_nq_local_search(int64_t* rows, int64_t n, int kind, int tie,
                            uint64_t allow_no_improvement_for, uint64_t max_iterations,
                            uint64_t window_size, int64_t* best_score, int64_t* current_out,
                            int64_t* current_score_out, int64_t* trace_a, int64_t* trace_b,
                            int64_t* trace_score, int64_t cap);

/* Throughput helper for the CPU baseline: scores `n_moves` explicit swap candidates by
 * clone + full re-score using `threads` OpenMP threads; returns candidates scored. */
int64_t orc_nq_baseline_sample(const int64_t* rows, int64_t n, const int64_t* a,
                               const int64_t* b, int64_t n_moves, int threads,
                               int64_t* checksum);

/* ---- checker-side O(1)-per-move delta scorer (NOT a reference function; see cs_oracle.c) ----
 * Occupancy counters + the SURVEY 8(a2) formulae with all four shared-line corrections.  Proven
 * against orc_nq_neighbourhood_deltas on every candidate of small boards (tests/test_oracle_cpu.py),
 * then used to check every candidate of benchmark-sized neighbourhoods.
 * Band: rows x in [x_begin, x_end) of the enumeration (swap: column i with all j > i; change: column
 * c with all values), written band-relative in enumeration order; identity -> INT64_MAX. */
int64_t orc_nq_fast_band_deltas(const int64_t* rows, int64_t n, int kind, int64_t x_begin, int64_t x_end,
                                int threads, int64_t* delta);
/* minimum of the whole neighbourhood by (delta, a, b); returns the non-identity candidates scored */
int64_t orc_nq_fast_argmin(const int64_t* rows, int64_t n, int kind, int threads, int64_t* best_delta,
                           int64_t* best_a, int64_t* best_b);
/* "CPU delta" courtesy baseline: same counters + deltas as the GPU, one chain per thread, full swap
 * neighbourhood + steepest descent per step.  Not the reference. */
int64_t orc_nq_delta_baseline(uint64_t seed, int64_t n, int chains, int steps, int threads, int64_t* checksum);
/* orc_nq_neighbourhood_deltas, candidates spread over OpenMP threads */
int64_t orc_nq_neighbourhood_deltas_mt(const int64_t* rows, int64_t n, int kind, int threads, int64_t* delta);

/* ---- employee scheduling ---- */
/* Days since 1970-01-01 -> weekday 0=Mon..6=Sun (chrono NaiveDate::weekday restated). */
int orc_weekday_from_days(int64_t days_since_epoch);
int64_t orc_days_from_civil(int64_t y, int m, int d);

/* get_scored_solution: examples/employee-scheduling/src/lib.rs:261-375 (+ :194-218, :148-192).
 * a[0..D) = employee id on scored day i (D = end-start+1; a phantom slot a[D] may exist in
 * the caller's vector, lib.rs:405-412, and is never read here).  start_weekday: 0=Mon.
 * holidays: n_hol pairs (hol_emp[k], hol_day[k]) with day index relative to start; the
 * reference unwrap()s a None for out-of-range days (lib.rs:275) -> return -1 here.
 * Returns 0 and writes hard/soft. */
int orc_es_score(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                 const int64_t* hol_day, int64_t n_hol, int64_t* hard, int64_t* soft);
/* Same, with each of the 8 terms separately: out[0..3]=H1..H4, out[4..7]=S1..S4. */
int orc_es_score_terms(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                       const int64_t* hol_day, int64_t n_hol, int64_t out[8]);

/* generate_initial_solution, examples/employee-scheduling/src/lib.rs:400-420 (n_slots = D+1) */
void orc_es_init(uint64_t seed, uint32_t chain, int64_t n_slots, const int64_t* employees,
                 int64_t E, int64_t* out);

enum { ORC_ES_CHANGE = 0, ORC_ES_SWAP = 1 };
/* Clone + full re-score of explicit moves. CHANGE: a[x[k]] = employees[y[k]] (y = index into
 * the employee id table); SWAP: exchange days x[k], y[k].  Identity => INT64_MAX in both. */
int orc_es_eval_moves(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                      const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                      int64_t E, int kind, const int64_t* x, const int64_t* y, int64_t n_moves,
                      int64_t* dhard, int64_t* dsoft);

/* LocalSearch::execute (local_search.rs:301-342) over the full change (D*E, day outer,
 * employee-index inner) then swap (d1<d2) neighbourhood, lexicographic (hard, soft) order,
 * first minimal candidate in enumeration order.  a: in=start, out=best. */
int64_t orc_es_local_search(int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                            const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                            int64_t E, uint64_t allow_no_improvement_for,
                            uint64_t max_iterations, int64_t* best_hard, int64_t* best_soft,
                            int64_t* current_out, int64_t* trace_kind, int64_t* trace_x,
                            int64_t* trace_y, int64_t* trace_hard, int64_t* trace_soft,
                            int64_t cap);

int64_t orc_es_baseline_sample(const int64_t* a, int64_t D, int start_weekday,
                               const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                               const int64_t* employees, int64_t E, int kind, const int64_t* x,
                               const int64_t* y, int64_t n_moves, int threads,
                               int64_t* checksum);

/* every candidate of the full neighbourhood (clone + full re-score each, OpenMP over candidates) in
 * the device's enumeration order including identities (INT64_MAX): D*E change, then D(D-1)/2 swap */
int64_t orc_es_neighbourhood_deltas_mt(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                                       const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                                       int threads, int64_t* dhard, int64_t* dsoft);

/* ---- slot-generalised scheduling: EXTENSION, NOT PINNED BY THE REFERENCE (see cs_oracle.c) ----
 * slots = D days x S shifts/day (slot t = day t / S, shift t % S), the reference's 8 terms in slot
 * units plus X1 same-day overlap and X2 skill (skills[k]: bit s = employees[k] works shift s; NULL =
 * everyone qualified).  Reduces to orc_es_score_terms at S = 1, skills = NULL.
 * out[0..3] = H1..H4, out[4..7] = S1..S4, out[8] = X1, out[9] = X2. */
int orc_esx_score_terms(const int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                        const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                        const int64_t* skills, int64_t out[10]);
int64_t orc_esx_neighbourhood_deltas_mt(const int64_t* a, int64_t D, int64_t S, int start_weekday,
                                        const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                                        const int64_t* employees, int64_t E, const int64_t* skills, int threads,
                                        int64_t* dhard, int64_t* dsoft);
int64_t orc_esx_local_search(int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                             const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                             const int64_t* skills, uint64_t allow_no_improvement_for, uint64_t max_iterations,
                             int threads, int64_t* best_hard, int64_t* best_soft, int64_t* current_out,
                             int64_t* trace_kind, int64_t* trace_x, int64_t* trace_y, int64_t* trace_hard,
                             int64_t* trace_soft, int64_t cap);
int64_t orc_esx_baseline_sample(const int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                                const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                                const int64_t* skills, int kind, const int64_t* x, const int64_t* y,
                                int64_t n_moves, int threads, int64_t* checksum);

/* ---- iterated local search (iterated_local_search.rs:173-202; see cs_oracle.c) ---- */
int64_t orc_nq_ils(uint64_t seed, uint32_t chain, int64_t n, int kind, uint64_t ls_max_iterations,
                   uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                   int64_t* best_rows, int64_t* best_score, int64_t* current_out,
                   int64_t* round_new_score, int64_t* round_choice, uint64_t ref_window);
/* LocalSearch::execute with the reference's OWN proposer (examples/nqueens/src/lib.rs:177-255:
 * sampled conflicted columns, change moves), window (.take(window_size)) and derived-Ord
 * tie-break (score, solution lexicographic); rng_t = LS rng draw counter (in/out). */
int64_t orc_nq_local_search_ref(int64_t* rows, int64_t n, uint64_t seed, uint32_t chain,
                                uint64_t* rng_t, uint64_t allow_no_improvement_for,
                                uint64_t max_iterations, uint64_t window_size, int64_t* best_score,
                                int64_t* current_out, int64_t* trace_a, int64_t* trace_b,
                                int64_t* trace_score, int64_t cap);
int64_t orc_es_ils(uint64_t seed, uint32_t chain, int64_t D, int start_weekday,
                   const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                   const int64_t* employees, int64_t E, uint64_t ls_max_iterations,
                   uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                   int64_t* best_idx, int64_t* best_hard, int64_t* best_soft,
                   int64_t* round_new_key, int64_t* round_choice);

/* LocalSearch::execute with the reference's OWN scheduling proposer
 * (ScheduleRandomMoveProposer, examples/employee-scheduling/src/lib.rs:440-491: endless random
 * ChangeDay / SwapDays stream from a CLONED rng), window (.take(window_size)) and derived-Ord
 * tie-break; see cs_oracle.c. */
int64_t orc_es_local_search_ref(int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                                const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                                int64_t E, uint64_t seed, uint32_t chain,
                                uint64_t allow_no_improvement_for, uint64_t max_iterations,
                                uint64_t window_size, uint64_t max_draws, int64_t* best_hard,
                                int64_t* best_soft, int64_t* current_out, int64_t* trace_kind,
                                int64_t* trace_x, int64_t* trace_y, int64_t* trace_hard,
                                int64_t* trace_soft, int64_t cap, int64_t* scored_out);
int64_t orc_es_ils_ref(uint64_t seed, uint32_t chain, int64_t D, int start_weekday,
                       const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                       const int64_t* employees, int64_t E, uint64_t ls_max_iterations,
                       uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                       int64_t* best_idx, int64_t* best_hard, int64_t* best_soft,
                       int64_t* round_new_key, int64_t* round_choice, uint64_t ref_window,
                       uint64_t ref_max_draws);

#ifdef __cplusplus
}
#endif
#endif
