//! `LocalSearch`-shaped front over the B200 evaluator.  SOURCE ONLY -- never compiled here
//! (no cargo/rustc in the build image); every call it makes is exercised through the same C ABI
//! from Python in tests/.
//!
//! Mirrors local-search/src/local_search.rs:253-343: `new(..)` takes the same solver constants,
//! `execute(start, allow_no_improvement_for)` returns the best ScoredSolution.  The move
//! proposer, score calculator and history live behind the handle (on the GPU).
pub mod ffi;
/// The reference's traits -- Solution, Score, SolutionScoreCalculator, MoveProposer,
/// InitialSolutionGenerator -- and a LocalSearch-compatible struct, implemented over the C ABI.
pub mod traits;

/// rows[col] = row, exactly the reference's `NQueensSolution.rows` (examples/nqueens/src/lib.rs:18-21)
pub struct B200NQueensLocalSearch {
    h: *mut ffi::cs_nq_handle,
    n: usize,
    max_iterations: u64,
}
// Send-not-Sync, like the handle (local_search.rs:16-17,23 only require Send)
unsafe impl Send for B200NQueensLocalSearch {}

impl B200NQueensLocalSearch {
    /// LocalSearch::new (local_search.rs:277-299); window / history capacities are accepted for
    /// signature compatibility -- the device scores the whole neighbourhood every step.
    pub fn new(board_size: usize, max_iterations: u64, _window_size: usize,
               _best_solutions_capacity: usize, _all_solutions_capacity: usize,
               _all_solution_iteration_expiry: u64, seed: u64, swap_moves: bool) -> Self {
        let cfg = ffi::cs_nq_config {
            n: board_size as u32, n_chains: 1, chain_offset: 0, trace_capacity: 0, seed, device: -1,
            neighbourhood: if swap_moves { ffi::CS_NQ_SWAP } else { ffi::CS_NQ_CHANGE }, flags: 0,
        };
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::cs_nq_create(&cfg, &mut h) };
        assert_eq!(rc, ffi::CS_OK, "cs_nq_create failed with status {}", rc); // the reference panics too
        Self { h, n: board_size, max_iterations }
    }

    /// LocalSearch::execute (local_search.rs:301-342): (best rows, best score)
    pub fn execute(&mut self, start: &[i64], allow_no_improvement_for: u64) -> (Vec<i64>, i64) {
        assert_eq!(start.len(), self.n);
        let mut best = vec![0i64; self.n];
        let mut score = 0i64;
        let rc = unsafe {
            ffi::cs_nq_local_search_one(self.h, start.as_ptr(), allow_no_improvement_for,
                                        self.max_iterations, best.as_mut_ptr(), &mut score)
        };
        assert_eq!(rc, ffi::CS_OK, "{}", unsafe { ffi::last_error_nq(self.h) });
        (best, score)
    }

    /// SolutionScoreCalculator::get_scored_solution (examples/nqueens/src/lib.rs:126-140)
    pub fn score(&mut self, rows: &[i64]) -> i64 {
        let mut s = 0i64;
        unsafe {
            assert_eq!(ffi::cs_nq_set_chains(self.h, 0, 1, rows.as_ptr()), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_score_full(self.h, 0, &mut s), ffi::CS_OK);
        }
        s
    }
}

impl Drop for B200NQueensLocalSearch {
    fn drop(&mut self) { unsafe { ffi::cs_nq_destroy(self.h); } }
}

/// Thousands of restart chains at once: IteratedLocalSearch (iterated_local_search.rs:96-203)
/// for every chain, Philox-seeded, stop as soon as one chain is solved.
pub struct B200NQueensIls { h: *mut ffi::cs_nq_handle, n: usize }
unsafe impl Send for B200NQueensIls {}

impl B200NQueensIls {
    pub fn new(board_size: usize, n_chains: u32, seed: u64, best_solutions_capacity: u32) -> Self {
        let cfg = ffi::cs_nq_config {
            n: board_size as u32, n_chains, chain_offset: 0, trace_capacity: 0, seed, device: -1,
            neighbourhood: ffi::CS_NQ_CHANGE, flags: 0,
        };
        let mut h = std::ptr::null_mut();
        unsafe {
            assert_eq!(ffi::cs_nq_create(&cfg, &mut h), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_init_random(h), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_ils_init(h, best_solutions_capacity, 0), ffi::CS_OK);
        }
        Self { h, n: board_size }
    }

    /// `while !is_finished { execute_round }` (examples/nqueens/src/main.rs:89-92) in one call
    pub fn solve(&mut self, max_rounds: u32, ls_max_iterations: u64, allow: u64) -> (Vec<i64>, i64) {
        let mut st = ffi::cs_ils_stats::default();
        let mut rows = vec![0i64; self.n];
        let mut score = 0i64;
        unsafe {
            assert_eq!(ffi::cs_nq_ils_run(self.h, max_rounds, ls_max_iterations, allow, 1, &mut st), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_ils_get_best(self.h, st.best_chain, rows.as_mut_ptr(), &mut score), ffi::CS_OK);
        }
        (rows, score)
    }
}

impl Drop for B200NQueensIls {
    fn drop(&mut self) { unsafe { ffi::cs_nq_destroy(self.h); } }
}

/// On-call rota: `date_to_employee` ids (length n_days + 1, phantom slot last) in, best out.
/// hard/soft convert to the reference's ScheduleScore with OrderedFloat(x as f64) (exact).
pub struct B200ScheduleLocalSearch { h: *mut ffi::cs_es_handle, slots: usize, max_iterations: u64 }
unsafe impl Send for B200ScheduleLocalSearch {}

impl B200ScheduleLocalSearch {
    /// employee ids, holidays as (employee id, (holiday - start_date).num_days()),
    /// start_weekday = start_date.weekday().num_days_from_monday()
    pub fn new(n_days: u32, start_weekday: u32, employees: &[i64], holidays: &[(i64, i64)],
               max_iterations: u64, seed: u64) -> Self {
        let cfg = ffi::cs_es_config {
            n_days, n_employees: employees.len() as u32, start_weekday, n_chains: 1, chain_offset: 0,
            trace_capacity: 0, seed, device: -1, flags: 1 /* CS_ES_FLAG_REFERENCE_PROPOSER: get_ils installs ScheduleRandomMoveProposer */,
        };
        let he: Vec<i64> = holidays.iter().map(|h| h.0).collect();
        let hd: Vec<i64> = holidays.iter().map(|h| h.1).collect();
        let mut h = std::ptr::null_mut();
        let rc = unsafe {
            ffi::cs_es_create(&cfg, employees.as_ptr(), he.as_ptr(), hd.as_ptr(), he.len() as u64, &mut h)
        };
        assert_eq!(rc, ffi::CS_OK, "cs_es_create failed with status {}", rc);
        Self { h, slots: n_days as usize + 1, max_iterations }
    }

    pub fn execute(&mut self, start: &[i64], allow_no_improvement_for: u64) -> (Vec<i64>, i64, i64) {
        assert_eq!(start.len(), self.slots);
        let mut best = vec![0i64; self.slots];
        let (mut hard, mut soft) = (0i64, 0i64);
        let rc = unsafe {
            ffi::cs_es_local_search_one(self.h, start.as_ptr(), allow_no_improvement_for,
                                        self.max_iterations, best.as_mut_ptr(), &mut hard, &mut soft)
        };
        assert_eq!(rc, ffi::CS_OK);
        (best, hard, soft)
    }
}

impl Drop for B200ScheduleLocalSearch {
    fn drop(&mut self) { unsafe { ffi::cs_es_destroy(self.h); } }
}
