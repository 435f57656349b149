/*
 * cs_b200.h -- C ABI of the B200-native local-search move evaluator.
 *
 * Drop-in boundary for ONE path of asimihsan/constraint-solver: the `local-search` crate's
 * move-evaluation loop (enumerate neighbourhood -> score every candidate -> select best ->
 * accept) for the nqueens and employee-scheduling plug-ins.  The reference has no FFI for this
 * path; the seam is its generic Rust trait surface.  Each entry point below names the
 * reference interface (file:line under the reference root) it stands in for.  The Rust-side
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns int32_t status (CS_OK == 0, errors < 0); nothing unwinds or
 *     throws across this boundary.  Where the reference would panic (unwrap on None, e.g.
 *     examples/employee-scheduling/src/lib.rs:275) the call returns CS_ERR_INVALID_ARG.
 *   - caller owns every in/out buffer; the library owns device memory behind the opaque
 *     handle; no pointer is retained past a call except the handle.
 *   - a handle is Send-not-Sync (one thread at a time), bound to ONE CUDA device; multi-GPU
 *     is one handle per process/GPU with chain_offset giving the global chain ids.
 *   - there is NO CPU fallback: without a CUDA device *_create returns CS_ERR_NO_DEVICE.
 *   - solutions cross the boundary in the reference's own element type: int64_t rows
 *     (examples/nqueens/src/lib.rs:13,19-21) and int64_t employee ids
 *     (examples/employee-scheduling/src/lib.rs:119-122).
 */
#ifndef CS_B200_H
#define CS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_OK 0
#define CS_ERR_INVALID_ARG (-1)
#define CS_ERR_CUDA (-2)
#define CS_ERR_NO_DEVICE (-3)
#define CS_ERR_OOM (-4)
#define CS_ERR_STATE (-5)
#define CS_ERR_UNSUPPORTED (-6)

#define CS_ABI_VERSION 4

/* Largest board the shared-memory-resident chain kernels take (one CTA holds rows,
 * per-column line sums and both diagonal counter arrays in <= 227 KB). */
#define CS_NQ_MAX_N_SMEM 16384u
/* Largest board at all: bigger boards (and CS_NQ_FLAG_GLOBAL) use the L2-resident kernels,
 * one instance per handle, swap neighbourhood, optionally partitioned across handles/GPUs. */
#define CS_NQ_MAX_N 1000000u
#define CS_NQ_FLAG_GLOBAL 1u /* force the global-memory (big board) path for any n */
#define CS_NQ_FLAG_SCALAR 2u /* never use the packed-window fast scan (parity / A-B testing) */
/* Reference mode (CS_NQ_CHANGE only): the reference's own move proposer -- conflicted columns
 * sampled by weight, a random subset of them, every row value of each
 * (examples/nqueens/src/lib.rs:177-255) -- truncated to window_size candidates and ordered by
 * the derived Ord (score, then solution vector), local_search.rs:315-323.  Random choices come
 * from the chain's Philox stream CS_PHILOX_LS. */
#define CS_NQ_FLAG_REFERENCE_PROPOSER 4u

/* neighbourhood kinds */
#define CS_NQ_SWAP 0u   /* exchange rows of columns i<j (new; defined against lib.rs:74-87) */
#define CS_NQ_CHANGE 1u /* rows[col] = value, examples/nqueens/src/lib.rs:227-229 */

/* per-chain status after a step / local-search call */
#define CS_CHAIN_RUNNING 0u  /* hit the step/iteration budget */
#define CS_CHAIN_BEST 1u     /* Score::is_best, local_search.rs:311-314 */
#define CS_CHAIN_STALLED 2u  /* no_improvement_for >= allow_no_improvement_for, :329-334 */
#define CS_CHAIN_EMPTY 3u    /* empty neighbourhood, :336-338 */

/* Philox stream purposes (ctr[3]) */
#define CS_PHILOX_INIT 0u
#define CS_PHILOX_PERTURB 1u
#define CS_PHILOX_LS 2u /* the LocalSearch-owned rng (reference-mode proposer) */
#define CS_PHILOX_HOLIDAYS 3u

typedef struct cs_move {
    uint32_t a; /* swap: column i   | change: column        */
    uint32_t b; /* swap: column j>i | change: new row value */
} cs_move;

typedef struct cs_step_stats {
    uint64_t moves_scored;    /* non-identity candidates delta-scored by this call (all chains) */
    uint64_t steps_accepted;  /* accepted moves summed over chains */
    int64_t best_score;       /* min current score over this handle's chains after the call */
    uint32_t best_chain;      /* local index of that chain (lowest index on ties) */
    uint32_t chains_at_best;  /* chains whose score is_best (== 0) */
    float device_ms;          /* CUDA-event time of this call's kernels on the handle's stream */
    uint32_t kernel_launches; /* kernels this call launched */
} cs_step_stats;

/* ------------------------------------------------------------------ library-wide */
int32_t cs_abi_version(void);
/* number of visible CUDA devices (0 when there is no driver/GPU) */
int32_t cs_device_count(void);
/* Host mirror of the device RNG so any chain can be replayed on the CPU.
 * key = {seed lo, seed hi}; ctr = {counter lo, counter hi, chain, purpose}.
 * Stands in for ChaCha20Rng::from_seed (examples/nqueens/src/main.rs:39,66) -- the random
 * STREAM is not reference-identical (rand/rand_chacha are un-vendored), only replayable. */
void cs_philox4x32_10(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t counter,
                      uint32_t out[4]);
const char* cs_status_string(int32_t status);

/* On-chip bandwidth micro-benchmarks -- the measured denominators of the roofline the benchmark
 * reports (the chain kernels are bound by the shared-memory data pipe or the L2 -> SM path, not by
 * HBM; SURVEY 8(d) asks for these to be measured, not computed).  No reference equivalent.
 *   CS_MICROBENCH_SMEM_LDS32 / _LDS128: conflict-free shared-memory loads streamed by every SM;
 *   CS_MICROBENCH_L2_READ: 16-byte loads over a 32 MB L2-resident buffer with L1 bypassed.
 * *gbs = best of three timed launches (CUDA events); *sm_mhz (optional) = rated SM clock. */
#define CS_MICROBENCH_SMEM_LDS32 0u
#define CS_MICROBENCH_SMEM_LDS128 1u
#define CS_MICROBENCH_L2_READ 2u
int32_t cs_microbench(int32_t device, uint32_t which, double* gbs, double* sm_mhz);

/* ------------------------------------------------------------------ n-queens */
typedef struct cs_nq_handle cs_nq_handle;

typedef struct cs_nq_config {
    uint32_t n;              /* board size, 1..CS_NQ_MAX_N */
    uint32_t n_chains;       /* independent restart chains held by this handle */
    uint32_t chain_offset;   /* global id of local chain 0 (Philox stream id; rank sharding) */
    uint32_t trace_capacity; /* chosen-move log entries kept per chain (0 = no trace) */
    uint64_t seed;           /* Philox key */
    int32_t device;          /* CUDA ordinal; -1 = current device */
    uint32_t neighbourhood;  /* CS_NQ_SWAP or CS_NQ_CHANGE */
    uint32_t flags;          /* CS_NQ_FLAG_* */
} cs_nq_config;

/* LocalSearch::new, local-search/src/local_search.rs:277-299 (the handle owns what the
 * struct owns: proposer + score calculator + history, for n_chains chains at once). */
int32_t cs_nq_create(const cs_nq_config* cfg, cs_nq_handle** out);
int32_t cs_nq_destroy(cs_nq_handle* h);
/* NUL-terminated, owned by the handle, valid until the next call on it. */
const char* cs_nq_last_error(const cs_nq_handle* h);
/* Run this handle's kernels on a caller-provided cudaStream_t (NULL = default stream). */
int32_t cs_nq_set_stream(cs_nq_handle* h, void* cuda_stream);

/* InitialSolutionGenerator::generate_initial_solution, examples/nqueens/src/lib.rs:152-161:
 * every chain := Fisher-Yates permutation of 0..n-1 from Philox (seed, chain_offset+k, INIT). */
int32_t cs_nq_init_random(cs_nq_handle* h);
/* Load `count` solutions (row-major int64 [count][n], values 0..n-1, any multiset -- change
 * moves and the perturbation legally break the permutation, lib.rs:228,311-312).  A value outside
 * [0, n) is CS_ERR_INVALID_ARG; the affected chains then hold the input with such values replaced
 * by 0, consistently scored (nothing on the device ever indexes by an out-of-range row). */
int32_t cs_nq_set_chains(cs_nq_handle* h, uint32_t first_chain, uint32_t count,
                         const int64_t* rows);
/* Double-buffered input staging for a stream of batches (no reference equivalent; the wasm worker
 * feeds one problem per message, web/employee-scheduling/src/worker.ts:8-22).  _async starts the
 * host -> device copy of `count` solutions on the library's copy stream and returns at once
 * (rows must stay valid and should be pinned until the commit); the copy overlaps whatever the
 * handle is running (e.g. cs_nq_step on the previous batch).  _commit waits for the copy, then
 * does what cs_nq_set_chains does (validate, pack, reset the chains' state, score).  One upload
 * may be pending per handle (CS_ERR_STATE otherwise); chain path only (CS_ERR_UNSUPPORTED on the
 * big-board path). */
int32_t cs_nq_set_chains_async(cs_nq_handle* h, uint32_t first_chain, uint32_t count,
                               const int64_t* rows);
int32_t cs_nq_commit_chains(cs_nq_handle* h);
int32_t cs_nq_get_chains(cs_nq_handle* h, uint32_t first_chain, uint32_t count, int64_t* rows);
/* current score of every chain (maintained by delta, written by step/local_search/set) */
int32_t cs_nq_get_scores(cs_nq_handle* h, int64_t* scores /* [n_chains] */);
int32_t cs_nq_get_status(cs_nq_handle* h, uint32_t* status /* [n_chains] */);

/* SolutionScoreCalculator::get_scored_solution, examples/nqueens/src/lib.rs:126-140.
 * Device FULL re-score by the O(n^2) pair test of get_col_scores (:74-87) -- deliberately
 * not the counter formulation, so it cross-checks the delta path. */
int32_t cs_nq_score_full(cs_nq_handle* h, uint32_t chain, int64_t* score);

/* Parity hook: exact score delta of explicit moves against chain's current state, computed
 * from the diagonal/row occupancy counters.  Identity moves (candidate == current, filtered
 * by the tabu set, local_search.rs:155-199,319) report INT64_MAX. */
int32_t cs_nq_eval_moves(cs_nq_handle* h, uint32_t chain, uint32_t kind, const cs_move* moves,
                         uint64_t n_moves, int64_t* delta);
/* MoveProposer::iter_local_moves (local_search.rs:85-89; nqueens lib.rs:173-256), full
 * neighbourhood of the handle's kind in device enumeration order, identity moves skipped.
 * Writes up to cap moves; *n_out = total count. */
int32_t cs_nq_enumerate(cs_nq_handle* h, uint32_t chain, cs_move* moves, uint64_t cap,
                        uint64_t* n_out);
/* Debug/parity: run the PRODUCTION neighbourhood scan on one chain and write every
 * candidate's delta in enumeration order (swap: i<j row-major, n(n-1)/2 entries; change:
 * (c,v) row-major, n*n entries); identity -> INT64_MAX. */
int32_t cs_nq_neighbourhood_deltas(cs_nq_handle* h, uint32_t chain, int64_t* delta,
                                   uint64_t cap, uint64_t* n_out);

/* The same dump for a BAND of columns (swap neighbourhood): columns i in [i_begin, i_end), every
 * j > i, band-relative enumeration order (entry of (i, j) = tri(i) - tri(i_begin) + j - i - 1 with
 * tri(x) = x*n - x(x+1)/2).  This is how every-candidate parity is checked at sizes whose whole
 * neighbourhood does not fit a buffer (n = 10^6: 5e11 candidates; a band of 8 columns is 8e6).  On
 * the big-board path the production scan is restricted to the band, exactly as a partition is. */
int32_t cs_nq_band_deltas(cs_nq_handle* h, uint32_t chain, uint32_t i_begin, uint32_t i_end,
                          int64_t* delta, uint64_t cap, uint64_t* n_out);

/* window_size of LocalSearch::new (local_search.rs:281); only reference mode truncates the
 * neighbourhood (default 5 * n, examples/nqueens/src/main.rs:130). */
int32_t cs_nq_set_window(cs_nq_handle* h, uint64_t window_size);

/* The hot path: for every chain, n_steps times: enumerate the full neighbourhood,
 * delta-score every candidate, argmin by (delta, a, b), accept unconditionally
 * (local_search.rs:315-335 with window = whole neighbourhood).  A chain stops early when its
 * score is_best or its neighbourhood is empty.  stats may be NULL. */
int32_t cs_nq_step(cs_nq_handle* h, uint32_t n_steps, cs_step_stats* stats);
/* The same launch split in two so a multi-GPU driver never blocks the host between a step and the
 * collective that follows it: _enqueue puts the n_steps chain-steps on the handle's stream and
 * returns; _wait waits for the stream and reports the LAST enqueued launch (stats may be NULL).
 * cs_nq_step == _enqueue + _wait.  Not available on the big-board path (its step is a host loop of
 * scan + apply: use cs_nq_part_scan / cs_nq_part_apply). */
int32_t cs_nq_step_enqueue(cs_nq_handle* h, uint32_t n_steps);
int32_t cs_nq_step_wait(cs_nq_handle* h, cs_step_stats* stats);

/* LocalSearch::execute, local-search/src/local_search.rs:301-342, on every chain from its
 * current state: bounded non-improving acceptance, best_solution bookkeeping (:326-328),
 * max_iterations.  Afterwards get_chains = last `current`, get_best_chains = returned best. */
int32_t cs_nq_local_search(cs_nq_handle* h, uint64_t allow_no_improvement_for,
                           uint64_t max_iterations, cs_step_stats* stats);
int32_t cs_nq_get_best_chains(cs_nq_handle* h, uint32_t first_chain, uint32_t count,
                              int64_t* rows, int64_t* best_scores);
/* Single-solution convenience with execute()'s exact shape: start -> best (+ score). */
int32_t cs_nq_local_search_one(cs_nq_handle* h, const int64_t* start,
                               uint64_t allow_no_improvement_for, uint64_t max_iterations,
                               int64_t* best, int64_t* best_score);

/* Chosen-move log of a chain since the last set/init (for CPU replay through the reference
 * scorer).  score_after[k] is the chain's score after move k.  *n_out = moves logged in total
 * (may exceed cap/trace_capacity; only the first trace_capacity are kept). */
int32_t cs_nq_get_trace(cs_nq_handle* h, uint32_t chain, cs_move* moves, int64_t* score_after,
                        uint64_t cap, uint64_t* n_out);
/* Best current solution over this handle's chains. Any out pointer may be NULL. */
int32_t cs_nq_best(cs_nq_handle* h, int64_t* rows, int64_t* score, uint32_t* chain);
/* Device pointer to the packed best key of this handle ((score << 32) | global chain id,
 * int64, refreshed by step/local_search) so a host collective (NCCL min-allreduce) can run
 * on it without a host round trip. */
int32_t cs_nq_best_key_device_ptr(cs_nq_handle* h, void** dptr);
/* Elite delivery without a host round trip (the exchange of SURVEY 8e): d_key is a DEVICE int64 --
 * the min-all-reduced best key, global chain id in its low 32 bits.  One kernel on the handle's
 * stream writes d_elite_u16[0, elite_len): the rows of that chain when this handle owns it
 * (chain_offset <= id < chain_offset + n_chains; zero past n), zeros otherwise -- so a sum
 * all-reduce of the buffer lands the elite on every rank. */
int32_t cs_nq_exchange_select(cs_nq_handle* h, const void* d_key, void* d_elite_u16, uint32_t elite_len);
/* Overwrite one chain with a solution (elite broadcast target). */
int32_t cs_nq_set_chain_u16_device(cs_nq_handle* h, uint32_t chain, const void* d_rows_u16);
/* Device pointer to chain's rows (uint16 [n], padded stride available via *stride_elems). */
int32_t cs_nq_chain_device_ptr(cs_nq_handle* h, uint32_t chain, void** dptr,
                               uint32_t* stride_elems);

/* --- one very large instance with its neighbourhood split across handles / GPUs (big-board
 * path only).  Every handle holds a full replica of the same instance; partition `part` of
 * `parts` scans the columns i of a triangular-balanced slice.  Per step: cs_nq_part_scan on
 * every handle, min-reduce the 8-byte keys (NCCL all-reduce over the device pointers), then
 * cs_nq_part_apply on every handle -- all replicas accept the same move, no state moves. */
int32_t cs_nq_set_partition(cs_nq_handle* h, uint32_t part, uint32_t parts);
/* enumerate + delta-score this partition's slice; leaves the packed key
 * ((delta/2 + 2^22) << 40 | i << 20 | j, int64; INT64_MAX = empty) on the device */
int32_t cs_nq_part_scan(cs_nq_handle* h);
int32_t cs_nq_part_key_device_ptr(cs_nq_handle* h, void** dptr);
/* accept the move currently in the key (after the caller's reduce); stats->moves_scored is
 * what THIS partition scanned */
int32_t cs_nq_part_apply(cs_nq_handle* h, cs_step_stats* stats);

/* ------------------------------------------------------------------ iterated local search */
/* IteratedLocalSearch (local-search/src/iterated_local_search.rs:96-203) for every chain of a
 * handle at once: perturbation (nqueens lib.rs:285-320 / employee-scheduling lib.rs:582-613),
 * LocalSearch::execute, History::local_search_chose_solution (bounded best-set,
 * local_search.rs:205-218), AcceptanceCriterion::choose weights 1:5:1 (:51-71), random restart
 * every 50th round (:185-191).  Random choices come from the chain's Philox stream
 * (purpose CS_PHILOX_PERTURB) so a chain replays bit-identically on the CPU. */
typedef struct cs_ils_stats {
    uint64_t moves_scored;   /* candidates delta-scored by the local searches of this call */
    uint64_t ls_steps;       /* accepted LS moves */
    int64_t best_key;        /* best over chains of History::get_best: n-queens score;
                                scheduling hard << 32 | soft; INT64_MAX before any round */
    uint32_t best_chain;
    uint32_t chains_done;    /* chains whose best is_best */
    uint32_t rounds_run;     /* execute_round calls issued by this call */
    float device_ms;
    uint32_t kernel_launches;
} cs_ils_stats;

/* IteratedLocalSearch::new (:130-156): current := each chain's present solution, empty
 * history of capacity best_solutions_capacity (1..64); log_capacity rounds of
 * (new local-minimum key, acceptance choice) are kept per chain for replay checks. */
int32_t cs_nq_ils_init(cs_nq_handle* h, uint32_t best_solutions_capacity, uint32_t log_capacity);
/* `rounds` x execute_round (:173-202) with LocalSearch::execute(perturbed,
 * allow_no_improvement_for) bounded by ls_max_iterations.  stop_when_any_best != 0 checks
 * after every round and returns as soon as some chain's best is_best. */
int32_t cs_nq_ils_run(cs_nq_handle* h, uint32_t rounds, uint64_t ls_max_iterations,
                      uint64_t allow_no_improvement_for, uint32_t stop_when_any_best,
                      cs_ils_stats* stats);
/* get_best_solution (:165-167) of one chain; CS_ERR_STATE before the first round (the
 * reference unwrap()s None there). */
int32_t cs_nq_ils_get_best(cs_nq_handle* h, uint32_t chain, int64_t* rows, int64_t* score);
int32_t cs_nq_ils_get_log(cs_nq_handle* h, uint32_t chain, int64_t* new_key, uint32_t* choice,
                          uint64_t cap, uint64_t* n_out);

/* ------------------------------------------------------------------ employee scheduling */
/* One employee per calendar day (examples/employee-scheduling/src/lib.rs:127-146).  A solution
 * is the reference's `date_to_employee`: n_days + 1 int64 employee ids -- the generator pushes
 * one phantom slot past end_date (lib.rs:405-412) that is stored and returned untouched but
 * never scored or moved.  Scores are the reference's (hard, soft) pair (lib.rs:239-249;
 * OrderedFloat<f64> holding integers only) as two int64. */
typedef struct cs_es_handle cs_es_handle;

#define CS_ES_MAX_SLOTS 192u /* scored slots = n_days * shifts_per_day (three 64-bit mask words) */
#define CS_ES_MAX_DAYS CS_ES_MAX_SLOTS /* at one shift per day (the reference's rota) */
#define CS_ES_MAX_SHIFTS 3u
#define CS_ES_CHANGE 0u /* ChangeDay: a = day (slot), b = index into the sorted employee table, lib.rs:466-470 */
#define CS_ES_SWAP 1u   /* SwapDays: a < b days (slots), lib.rs:471-478 */
/* Reference mode: the reference's own move proposer -- ScheduleRandomMoveProposer, the one get_ils
 * installs (examples/employee-scheduling/src/lib.rs:60, :440-491): an endless stream of random
 * ChangeDay (weight 1) / SwapDays (weight 4) candidates drawn from a CLONE of the LocalSearch rng
 * (:488; every step replays the same draws) -- tabu-filtered, truncated to window_size candidates
 * (local_search.rs:321) and ordered by the derived Ord (score, then date_to_employee), :323.
 * Random choices come from the chain's Philox stream CS_PHILOX_LS.  Without the flag the device
 * scans the FULL change + swap neighbourhood (ScheduleMoveProposer's precedent, lib.rs:493-559). */
#define CS_ES_FLAG_REFERENCE_PROPOSER 1u

typedef struct cs_es_config {
    uint32_t n_days;         /* D = end_date - start_date + 1; D * shifts_per_day <= CS_ES_MAX_SLOTS */
    uint32_t n_employees;    /* E >= 1; the per-chain day-mask table (8 B per employee) lives in shared memory,
                              * so E <= ~26 000 on B200 (CS_ERR_INVALID_ARG beyond: "employee table too large") */
    uint32_t start_weekday;  /* weekday of start_date, 0 = Monday .. 6 = Sunday */
    uint32_t n_chains;
    uint32_t chain_offset;
    uint32_t trace_capacity;
    uint64_t seed;
    int32_t device;
    uint32_t flags;          /* CS_ES_FLAG_* */
} cs_es_config;

typedef struct cs_es_move {
    uint32_t kind; /* CS_ES_CHANGE / CS_ES_SWAP */
    uint32_t a;
    uint32_t b;
} cs_es_move;

typedef struct cs_es_step_stats {
    uint64_t moves_scored;
    uint64_t steps_accepted;
    int64_t best_hard; /* lexicographically best (hard, soft) over this handle's chains */
    int64_t best_soft;
    uint32_t best_chain;
    uint32_t chains_at_best;  /* hard == 0 && soft == 0 */
    uint32_t chains_feasible; /* hard == 0 */
    float device_ms;
    uint32_t kernel_launches;
} cs_es_step_stats;

/* get_ils / ScheduleSolutionScoreCalculator::new (lib.rs:57-117, :255-259): employee ids (any
 * order, unique; kept sorted like the reference's BTreeSet) and the holiday table as n_hol
 * (employee id, day index from start_date) pairs.  A holiday outside [0, n_days) is
 * CS_ERR_INVALID_ARG (the reference unwrap()s a None there, lib.rs:275). */
int32_t cs_es_create(const cs_es_config* cfg, const int64_t* employee_ids, const int64_t* hol_emp,
                     const int64_t* hol_day, uint64_t n_hol, cs_es_handle** out);
/* EXTENSION, not pinned by the reference (which has one employee per calendar DAY and no skills;
 * BASELINE configs[2] / [3] say "3 shifts/day", the north-star names shift-overlap and skill tallies):
 * slots = n_days x shifts_per_day, slot t = day t / S, shift t % S.  A solution is n_days * S + 1
 * employee ids (phantom last); move indices are SLOTS.  The reference's 8 terms apply in slot units
 * (H2 consecutive slots; H3 per shift; H4 / S1 per 14- / 7-DAY window over slot counts; a holiday
 * covers every slot of its day) plus two hard terms: same-day overlap (pairs of slots of one day held
 * by one employee) and skill (skills[k]: bit s set = employee_ids[k] is qualified for shift kind s;
 * NULL = everybody for everything).  The full-re-score definition is oracle/cs_oracle.c: esx_terms.
 * shifts_per_day = 1 with skills = NULL IS cs_es_create (same kernels, same results).  The reference
 * proposer flag is only available there (CS_ERR_UNSUPPORTED otherwise). */
int32_t cs_es_create_ex(const cs_es_config* cfg, const int64_t* employee_ids, const int64_t* hol_emp,
                        const int64_t* hol_day, uint64_t n_hol, uint32_t shifts_per_day,
                        const uint32_t* skills, cs_es_handle** out);
/* n_days, shifts_per_day, slots per solution (n_days * shifts_per_day + 1); any pointer may be NULL */
int32_t cs_es_get_dims(cs_es_handle* h, uint32_t* n_days, uint32_t* shifts_per_day, uint32_t* n_slots);
int32_t cs_es_destroy(cs_es_handle* h);
const char* cs_es_last_error(const cs_es_handle* h);
int32_t cs_es_set_stream(cs_es_handle* h, void* cuda_stream);
/* ScheduleInitialSolutionGenerator::generate_initial_solution, lib.rs:400-420 */
int32_t cs_es_init_random(cs_es_handle* h);
/* rows: int64 [count][n_days + 1] employee ids */
int32_t cs_es_set_chains(cs_es_handle* h, uint32_t first_chain, uint32_t count, const int64_t* rows);
/* Double-buffered input staging, as cs_nq_set_chains_async / cs_nq_commit_chains: _async starts the
 * host -> device copy on the library's copy stream and returns (rows should be pinned and stay valid
 * until the commit); _commit waits for it, converts ids to indices, resets and scores the chains. */
int32_t cs_es_set_chains_async(cs_es_handle* h, uint32_t first_chain, uint32_t count, const int64_t* rows);
int32_t cs_es_commit_chains(cs_es_handle* h);
int32_t cs_es_get_chains(cs_es_handle* h, uint32_t first_chain, uint32_t count, int64_t* rows);
int32_t cs_es_get_scores(cs_es_handle* h, int64_t* hard, int64_t* soft);
int32_t cs_es_get_status(cs_es_handle* h, uint32_t* status);
/* ScheduleSolutionScoreCalculator::get_scored_solution, lib.rs:261-375: device full re-score by
 * the reference's own loops over the day vector (not the mask/tally formulation).
 * terms (optional) = H1..H4, S1..S4. */
int32_t cs_es_score_full(cs_es_handle* h, uint32_t chain, int64_t* hard, int64_t* soft,
                         int64_t terms[8]);
/* the same with the two extension terms: terms (optional) = H1..H4, S1..S4, X1 same-day overlap,
 * X2 skill; hard = H1+H2+H3+H4+X1+X2 */
int32_t cs_es_score_full_ex(cs_es_handle* h, uint32_t chain, int64_t* hard, int64_t* soft,
                            int64_t terms[10]);
/* exact (dhard, dsoft) of explicit moves; identity moves report INT64_MAX in both */
int32_t cs_es_eval_moves(cs_es_handle* h, uint32_t chain, const cs_es_move* moves, uint64_t n_moves,
                         int64_t* dhard, int64_t* dsoft);
/* ScheduleMoveProposer::iter_local_moves precedent (lib.rs:511-559): the FULL neighbourhood,
 * change moves (day outer, employee index inner) then swaps (a<b), identity moves skipped. */
int32_t cs_es_enumerate(cs_es_handle* h, uint32_t chain, cs_es_move* moves, uint64_t cap,
                        uint64_t* n_out);
/* production scan, every candidate in enumeration order INCLUDING identities (INT64_MAX):
 * n_days*n_employees change entries then n_days*(n_days-1)/2 swap entries */
int32_t cs_es_neighbourhood_deltas(cs_es_handle* h, uint32_t chain, int64_t* dhard, int64_t* dsoft,
                                   uint64_t cap, uint64_t* n_out);
/* window_size of LocalSearch::new (local_search.rs:281); only reference mode truncates the
 * neighbourhood (default 100, examples/employee-scheduling/src/main.rs:26). */
int32_t cs_es_set_window(cs_es_handle* h, uint64_t window_size);
/* hot path: enumerate change + swap moves, delta-score the 8 constraints, lexicographic
 * (hard, soft, move id) argmin, accept (local_search.rs:315-335) */
int32_t cs_es_step(cs_es_handle* h, uint32_t n_steps, cs_es_step_stats* stats);
/* enqueue-only / wait halves of cs_es_step, as cs_nq_step_enqueue / cs_nq_step_wait */
int32_t cs_es_step_enqueue(cs_es_handle* h, uint32_t n_steps);
int32_t cs_es_step_wait(cs_es_handle* h, cs_es_step_stats* stats);
/* LocalSearch::execute, local_search.rs:301-342 */
int32_t cs_es_local_search(cs_es_handle* h, uint64_t allow_no_improvement_for,
                           uint64_t max_iterations, cs_es_step_stats* stats);
int32_t cs_es_get_best_chains(cs_es_handle* h, uint32_t first_chain, uint32_t count, int64_t* rows,
                              int64_t* best_hard, int64_t* best_soft);
int32_t cs_es_local_search_one(cs_es_handle* h, const int64_t* start,
                               uint64_t allow_no_improvement_for, uint64_t max_iterations,
                               int64_t* best, int64_t* best_hard, int64_t* best_soft);
int32_t cs_es_get_trace(cs_es_handle* h, uint32_t chain, cs_es_move* moves, int64_t* hard_after,
                        int64_t* soft_after, uint64_t cap, uint64_t* n_out);
int32_t cs_es_best(cs_es_handle* h, int64_t* rows, int64_t* hard, int64_t* soft, uint32_t* chain);
/* device int64: (hard << 48) | (soft << 32) | global chain id */
int32_t cs_es_best_key_device_ptr(cs_es_handle* h, void** dptr);
int32_t cs_es_chain_device_ptr(cs_es_handle* h, uint32_t chain, void** dptr, uint32_t* n_slots);
/* as cs_nq_exchange_select: the owning handle writes the chain's dense employee indices (uint16) */
int32_t cs_es_exchange_select(cs_es_handle* h, const void* d_key, void* d_elite_u16, uint32_t elite_len);
/* ILS shell, see cs_nq_ils_*; the perturbation may rewrite the phantom slot (lib.rs:599-608) */
int32_t cs_es_ils_init(cs_es_handle* h, uint32_t best_solutions_capacity, uint32_t log_capacity);
int32_t cs_es_ils_run(cs_es_handle* h, uint32_t rounds, uint64_t ls_max_iterations,
                      uint64_t allow_no_improvement_for, uint32_t stop_when_any_best,
                      cs_ils_stats* stats);
int32_t cs_es_ils_get_best(cs_es_handle* h, uint32_t chain, int64_t* rows, int64_t* hard,
                           int64_t* soft);
int32_t cs_es_ils_get_log(cs_es_handle* h, uint32_t chain, int64_t* new_key, uint32_t* choice,
                          uint64_t cap, uint64_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
