//! The reference's trait surface (local-search/src/local_search.rs:16-90) implemented over the C ABI
//! of include/cs_b200.h.  SOURCE ONLY -- there is no cargo/rustc in the build image, so this file has
//! never been compiled; every C entry point it calls is exercised from Python and C++ in tests/.
//!
//! Two ways to use it:
//!
//!  1. LITERAL drop-in, no change to the reference: the reference's own `LocalSearch` and
//!     `IteratedLocalSearch` are generic over `SolutionScoreCalculator` and `MoveProposer`
//!     (local_search.rs:253-263, iterated_local_search.rs:96-117), so
//!     `LocalSearch::new(B200NQueensMoveProposer::new(n), B200NQueensScoreCalculator::new(n), ..)`
//!     type-checks as is: candidates come from `cs_nq_enumerate`, scores from `cs_nq_score_full`.
//!     Correct and bit-identical in score, but it keeps the reference's per-candidate clone +
//!     full re-score loop (one device call per candidate) -- it is the compatibility path.
//!  2. FAST path: `B200LocalSearch` has `LocalSearch`'s 8-argument `new` (local_search.rs:277-299)
//!     and `execute(start, allow_no_improvement_for) -> ScoredSolution` (:301-305) but runs
//!     enumerate + delta-score + select + accept on the device (`cs_*_local_search_one`).
//!     `IteratedLocalSearch` owns a concrete `LocalSearch` by value (iterated_local_search.rs:108),
//!     so using the fast path inside it needs the one-line change shown in INTEGRATION.md (make
//!     the field generic over the `Execute` trait below) -- or use `cs_*_ils_run`, which runs the
//!     whole ILS shell on the device.
use std::cell::RefCell;
use std::marker::PhantomData;

use local_search::local_search::{
    InitialSolutionGenerator, MoveProposer, Score, ScoredSolution, Solution, SolutionScoreCalculator,
};

use crate::ffi;

// ------------------------------------------------------------------------------------------ n-queens
/// rows[col] = row -- the reference's `NQueensSolution` (examples/nqueens/src/lib.rs:16-21; its field is
/// private to that crate, hence the twin type).  Derived Ord = lexicographic over rows, as there.
#[derive(Clone, PartialEq, Eq, PartialOrd, Ord, Hash, Debug)]
pub struct B200NQueensSolution {
    pub rows: Vec<i64>,
}
impl Solution for B200NQueensSolution {}

/// examples/nqueens/src/lib.rs:62-71
#[derive(Clone, Debug, Eq, PartialEq, Ord, PartialOrd, Hash)]
pub struct B200NQueensScore(pub i64);
impl Score for B200NQueensScore {
    fn is_best(&self) -> bool {
        self.0 == 0
    }
}

/// One device handle, shared by the calculator / proposer / generator of one problem instance.
/// Send-not-Sync like every `cs_*_handle` (the traits only require Send, local_search.rs:16-17,23).
pub struct NqDevice {
    h: *mut ffi::cs_nq_handle,
    n: usize,
}
unsafe impl Send for NqDevice {}
impl NqDevice {
    pub fn new(board_size: usize, neighbourhood: u32, flags: u32, seed: u64) -> Self {
        let cfg = ffi::cs_nq_config {
            n: board_size as u32, n_chains: 1, chain_offset: 0, trace_capacity: 0, seed, device: -1,
            neighbourhood, flags,
        };
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::cs_nq_create(&cfg, &mut h) };
        assert_eq!(rc, ffi::CS_OK, "cs_nq_create failed with status {}", rc); // the reference panics too
        Self { h, n: board_size }
    }
    fn load(&self, rows: &[i64]) {
        assert_eq!(rows.len(), self.n);
        let rc = unsafe { ffi::cs_nq_set_chains(self.h, 0, 1, rows.as_ptr()) };
        assert_eq!(rc, ffi::CS_OK, "{}", unsafe { ffi::last_error_nq(self.h) });
    }
}
impl Drop for NqDevice {
    fn drop(&mut self) {
        unsafe { ffi::cs_nq_destroy(self.h); }
    }
}

/// `impl SolutionScoreCalculator` (local_search.rs:58-66; examples/nqueens/src/lib.rs:122-140).
pub struct B200NQueensScoreCalculator {
    dev: RefCell<NqDevice>,
}
impl B200NQueensScoreCalculator {
    pub fn new(board_size: usize) -> Self {
        Self { dev: RefCell::new(NqDevice::new(board_size, ffi::CS_NQ_CHANGE, 0, 0)) }
    }
}
impl SolutionScoreCalculator for B200NQueensScoreCalculator {
    type _Solution = B200NQueensSolution;
    type _Score = B200NQueensScore;

    fn get_scored_solution(&self, solution: Self::_Solution) -> ScoredSolution<Self::_Solution, Self::_Score> {
        let dev = self.dev.borrow_mut();
        dev.load(&solution.rows);
        let mut s = 0i64;
        let rc = unsafe { ffi::cs_nq_score_full(dev.h, 0, &mut s) };
        assert_eq!(rc, ffi::CS_OK);
        ScoredSolution { score: B200NQueensScore(s), solution }
    }
}

/// `impl InitialSolutionGenerator` (local_search.rs:68-75; examples/nqueens/src/lib.rs:142-161): a
/// Philox Fisher-Yates permutation on the device.  `R` is whatever rng the solver is built with; the
/// device stream is keyed by (seed, chain), one draw of `rng` picks the chain so repeated calls differ.
pub struct B200NQueensInitialSolutionGenerator<R: rand::Rng> {
    board_size: usize,
    seed: u64,
    _r: PhantomData<R>,
}
impl<R: rand::Rng> B200NQueensInitialSolutionGenerator<R> {
    pub fn new(board_size: usize, seed: u64) -> Self {
        Self { board_size, seed, _r: PhantomData }
    }
}
impl<R: rand::Rng> InitialSolutionGenerator for B200NQueensInitialSolutionGenerator<R> {
    type R = R;
    type Solution = B200NQueensSolution;

    fn generate_initial_solution(&self, rng: &mut Self::R) -> Self::Solution {
        let chain: u32 = rng.gen();
        let cfg = ffi::cs_nq_config {
            n: self.board_size as u32, n_chains: 1, chain_offset: chain & 0x7fff_ffff, trace_capacity: 0,
            seed: self.seed, device: -1, neighbourhood: ffi::CS_NQ_CHANGE, flags: 0,
        };
        let mut h = std::ptr::null_mut();
        let mut rows = vec![0i64; self.board_size];
        unsafe {
            assert_eq!(ffi::cs_nq_create(&cfg, &mut h), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_init_random(h), ffi::CS_OK);
            assert_eq!(ffi::cs_nq_get_chains(h, 0, 1, rows.as_mut_ptr()), ffi::CS_OK);
            ffi::cs_nq_destroy(h);
        }
        B200NQueensSolution { rows }
    }
}

/// `impl MoveProposer` (local_search.rs:77-90): the FULL neighbourhood of the handle's kind in device
/// enumeration order, identity moves skipped (`cs_nq_enumerate`), each materialised as a solution the
/// way the reference's iterator does (examples/nqueens/src/lib.rs:217-234).
pub struct B200NQueensMoveProposer<R: rand::Rng> {
    dev: RefCell<NqDevice>,
    swap: bool,
    _r: PhantomData<R>,
}
impl<R: rand::Rng> B200NQueensMoveProposer<R> {
    pub fn new(board_size: usize, swap_moves: bool) -> Self {
        let kind = if swap_moves { ffi::CS_NQ_SWAP } else { ffi::CS_NQ_CHANGE };
        Self { dev: RefCell::new(NqDevice::new(board_size, kind, 0, 0)), swap: swap_moves, _r: PhantomData }
    }
}
impl<R: rand::Rng> MoveProposer for B200NQueensMoveProposer<R> {
    type R = R;
    type Solution = B200NQueensSolution;

    fn iter_local_moves(&self, start: &Self::Solution, _rng: &mut Self::R) -> Box<dyn Iterator<Item = Self::Solution>> {
        let dev = self.dev.borrow_mut();
        dev.load(&start.rows);
        let mut n_moves = 0u64;
        unsafe { assert_eq!(ffi::cs_nq_enumerate(dev.h, 0, std::ptr::null_mut(), 0, &mut n_moves), ffi::CS_OK); }
        let mut moves = vec![ffi::cs_move::default(); n_moves as usize];
        unsafe { assert_eq!(ffi::cs_nq_enumerate(dev.h, 0, moves.as_mut_ptr(), n_moves, &mut n_moves), ffi::CS_OK); }
        let (start, swap) = (start.clone(), self.swap);
        Box::new(moves.into_iter().map(move |m| {
            let mut s = start.clone();
            if swap {
                s.rows.swap(m.a as usize, m.b as usize);
            } else {
                s.rows[m.a as usize] = m.b as i64; // rows[col] = value, lib.rs:227-229
            }
            s
        }))
    }
}

// ------------------------------------------------------------------------------------------ scheduling
/// `Employee` (examples/employee-scheduling/src/lib.rs:119-122) and the rota
/// `date_to_employee` incl. the phantom slot (:127-146, :405-412); Ord = lexicographic by id, as
/// the reference's derived Ord is once its ignored fields are dropped.
#[derive(Clone, PartialEq, Eq, PartialOrd, Ord, Hash, Debug)]
pub struct B200ScheduleSolution {
    pub date_to_employee: Vec<i64>,
}
impl Solution for B200ScheduleSolution {}

/// `ScheduleScore` (:239-249): the reference stores OrderedFloat<f64> holding integers only; the
/// derived Ord (hard, then soft) is the same on i64.  `to_f64()` gives the reference's pair.
#[derive(Clone, Debug, Eq, PartialEq, Ord, PartialOrd, Hash)]
pub struct B200ScheduleScore {
    pub hard_score: i64,
    pub soft_score: i64,
}
impl B200ScheduleScore {
    pub fn to_f64(&self) -> (f64, f64) {
        (self.hard_score as f64, self.soft_score as f64)
    }
}
impl Score for B200ScheduleScore {
    fn is_best(&self) -> bool {
        self.hard_score == 0 && self.soft_score == 0
    }
}

pub struct EsDevice {
    h: *mut ffi::cs_es_handle,
    slots: usize,
    n_scored: usize,
    n_employees: usize,
    employees_sorted: Vec<i64>,
}
unsafe impl Send for EsDevice {}
impl EsDevice {
    /// holidays as (employee id, (holiday - start_date).num_days()); start_weekday =
    /// start_date.weekday().num_days_from_monday(); flags: 1 = the reference's own random proposer
    pub fn new(n_days: u32, start_weekday: u32, employees: &[i64], holidays: &[(i64, i64)], flags: u32, seed: u64) -> Self {
        let cfg = ffi::cs_es_config {
            n_days, n_employees: employees.len() as u32, start_weekday, n_chains: 1, chain_offset: 0,
            trace_capacity: 0, seed, device: -1, flags,
        };
        let he: Vec<i64> = holidays.iter().map(|h| h.0).collect();
        let hd: Vec<i64> = holidays.iter().map(|h| h.1).collect();
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::cs_es_create(&cfg, employees.as_ptr(), he.as_ptr(), hd.as_ptr(), he.len() as u64, &mut h) };
        assert_eq!(rc, ffi::CS_OK, "cs_es_create failed with status {}", rc);
        let mut sorted = employees.to_vec();
        sorted.sort();
        Self { h, slots: n_days as usize + 1, n_scored: n_days as usize, n_employees: employees.len(), employees_sorted: sorted }
    }
    fn load(&self, rows: &[i64]) {
        assert_eq!(rows.len(), self.slots);
        assert_eq!(unsafe { ffi::cs_es_set_chains(self.h, 0, 1, rows.as_ptr()) }, ffi::CS_OK);
    }
}
impl Drop for EsDevice {
    fn drop(&mut self) {
        unsafe { ffi::cs_es_destroy(self.h); }
    }
}

/// `impl SolutionScoreCalculator` for the rota (lib.rs:251-375)
pub struct B200ScheduleScoreCalculator {
    dev: RefCell<EsDevice>,
}
impl B200ScheduleScoreCalculator {
    pub fn new(n_days: u32, start_weekday: u32, employees: &[i64], holidays: &[(i64, i64)]) -> Self {
        Self { dev: RefCell::new(EsDevice::new(n_days, start_weekday, employees, holidays, 0, 0)) }
    }
}
impl SolutionScoreCalculator for B200ScheduleScoreCalculator {
    type _Solution = B200ScheduleSolution;
    type _Score = B200ScheduleScore;

    fn get_scored_solution(&self, solution: Self::_Solution) -> ScoredSolution<Self::_Solution, Self::_Score> {
        let dev = self.dev.borrow_mut();
        dev.load(&solution.date_to_employee);
        let (mut hard, mut soft) = (0i64, 0i64);
        let rc = unsafe { ffi::cs_es_score_full(dev.h, 0, &mut hard, &mut soft, std::ptr::null_mut()) };
        assert_eq!(rc, ffi::CS_OK);
        ScoredSolution { score: B200ScheduleScore { hard_score: hard, soft_score: soft }, solution }
    }
}

/// `impl MoveProposer` for the rota: the full ChangeDay + SwapDays neighbourhood
/// (ScheduleMoveProposer's precedent, lib.rs:493-559) in device enumeration order (`cs_es_enumerate`).
pub struct B200ScheduleMoveProposer<R: rand::Rng> {
    dev: RefCell<EsDevice>,
    _r: PhantomData<R>,
}
impl<R: rand::Rng> B200ScheduleMoveProposer<R> {
    pub fn new(n_days: u32, start_weekday: u32, employees: &[i64], holidays: &[(i64, i64)]) -> Self {
        Self { dev: RefCell::new(EsDevice::new(n_days, start_weekday, employees, holidays, 0, 0)), _r: PhantomData }
    }
}
impl<R: rand::Rng> MoveProposer for B200ScheduleMoveProposer<R> {
    type R = R;
    type Solution = B200ScheduleSolution;

    fn iter_local_moves(&self, start: &Self::Solution, _rng: &mut Self::R) -> Box<dyn Iterator<Item = Self::Solution>> {
        let dev = self.dev.borrow_mut();
        dev.load(&start.date_to_employee);
        let mut n_moves = 0u64;
        unsafe { assert_eq!(ffi::cs_es_enumerate(dev.h, 0, std::ptr::null_mut(), 0, &mut n_moves), ffi::CS_OK); }
        let mut moves = vec![ffi::cs_es_move::default(); n_moves as usize];
        unsafe { assert_eq!(ffi::cs_es_enumerate(dev.h, 0, moves.as_mut_ptr(), n_moves, &mut n_moves), ffi::CS_OK); }
        let (start, ids) = (start.clone(), dev.employees_sorted.clone());
        debug_assert!(dev.n_scored * dev.n_employees + dev.n_scored * (dev.n_scored - 1) / 2 >= moves.len());
        Box::new(moves.into_iter().map(move |m| {
            let mut s = start.clone();
            if m.kind == ffi::CS_ES_CHANGE {
                s.date_to_employee[m.a as usize] = ids[m.b as usize]; // lib.rs:466-470
            } else {
                s.date_to_employee.swap(m.a as usize, m.b as usize); // lib.rs:471-478
            }
            s
        }))
    }
}

// ------------------------------------------------------------------------------------------ LocalSearch
/// What a solution type must offer for the device to run `LocalSearch::execute` on it.
pub trait DeviceProblem {
    type _Solution: Solution;
    type _Score: Score;
    /// run LocalSearch::execute on the device from `start`; returns the best solution and score
    fn execute_on_device(&mut self, start: &Self::_Solution, allow_no_improvement_for: u64, max_iterations: u64,
                         window_size: usize) -> ScoredSolution<Self::_Solution, Self::_Score>;
}

impl DeviceProblem for NqDevice {
    type _Solution = B200NQueensSolution;
    type _Score = B200NQueensScore;
    fn execute_on_device(&mut self, start: &B200NQueensSolution, allow: u64, max_iterations: u64, window_size: usize)
                         -> ScoredSolution<B200NQueensSolution, B200NQueensScore> {
        let mut best = vec![0i64; self.n];
        let mut score = 0i64;
        unsafe {
            assert_eq!(ffi::cs_nq_set_window(self.h, window_size.max(1) as u64), ffi::CS_OK);
            let rc = ffi::cs_nq_local_search_one(self.h, start.rows.as_ptr(), allow, max_iterations, best.as_mut_ptr(), &mut score);
            assert_eq!(rc, ffi::CS_OK, "{}", ffi::last_error_nq(self.h));
        }
        ScoredSolution { score: B200NQueensScore(score), solution: B200NQueensSolution { rows: best } }
    }
}

impl DeviceProblem for EsDevice {
    type _Solution = B200ScheduleSolution;
    type _Score = B200ScheduleScore;
    fn execute_on_device(&mut self, start: &B200ScheduleSolution, allow: u64, max_iterations: u64, window_size: usize)
                         -> ScoredSolution<B200ScheduleSolution, B200ScheduleScore> {
        let mut best = vec![0i64; self.slots];
        let (mut hard, mut soft) = (0i64, 0i64);
        unsafe {
            assert_eq!(ffi::cs_es_set_window(self.h, window_size.max(1) as u64), ffi::CS_OK);
            let rc = ffi::cs_es_local_search_one(self.h, start.date_to_employee.as_ptr(), allow, max_iterations,
                                                 best.as_mut_ptr(), &mut hard, &mut soft);
            assert_eq!(rc, ffi::CS_OK);
        }
        ScoredSolution { score: B200ScheduleScore { hard_score: hard, soft_score: soft },
                         solution: B200ScheduleSolution { date_to_employee: best } }
    }
}

/// The one method `IteratedLocalSearch::execute_round` needs from its local search
/// (iterated_local_search.rs:195-197); implemented by `B200LocalSearch`, and trivially by the
/// reference's `LocalSearch` (`fn execute(..) { LocalSearch::execute(self, ..) }`).
pub trait Execute<_Solution: Solution, _Score: Score> {
    fn execute(&mut self, start: _Solution, allow_no_improvement_for: u64) -> ScoredSolution<_Solution, _Score>;
}

/// `LocalSearch<R, _Solution, _Score, SSC, MP>` with the whole loop on the device.  Same type
/// parameters, the same 8-argument `new` (local_search.rs:277-299) and the same `execute`
/// (:301-305).  The proposer and calculator are kept (and usable on their own) but `execute`
/// does not call them: enumeration, delta scoring, selection and acceptance run behind `device`.
pub struct B200LocalSearch<R, _Solution, _Score, SSC, MP, DP>
where
    R: rand::Rng,
    _Solution: Solution,
    _Score: Score,
    SSC: SolutionScoreCalculator<_Solution = _Solution, _Score = _Score>,
    MP: MoveProposer<R = R, Solution = _Solution>,
    DP: DeviceProblem<_Solution = _Solution, _Score = _Score>,
{
    pub move_proposer: MP,
    pub solution_score_calculator: SSC,
    max_iterations: u64,
    window_size: usize,
    // History::new(best_solutions_capacity, all_solutions_capacity, all_solution_iteration_expiry),
    // local_search.rs:292-296: kept for signature compatibility -- the reference's tabu set is always
    // {current} (its age test is inverted, :182-195), which the device implements by skipping
    // identity moves, so neither capacity changes a result.
    _best_solutions_capacity: usize,
    _all_solutions_capacity: usize,
    _all_solution_iteration_expiry: u64,
    _rng: R,
    device: DP,
}

impl<R, _Solution, _Score, SSC, MP, DP> B200LocalSearch<R, _Solution, _Score, SSC, MP, DP>
where
    R: rand::Rng,
    _Solution: Solution,
    _Score: Score,
    SSC: SolutionScoreCalculator<_Solution = _Solution, _Score = _Score>,
    MP: MoveProposer<R = R, Solution = _Solution>,
    DP: DeviceProblem<_Solution = _Solution, _Score = _Score>,
{
    /// local_search.rs:277-299 plus the device the loop runs on
    #[allow(clippy::too_many_arguments)]
    pub fn new(move_proposer: MP, solution_score_calculator: SSC, max_iterations: u64, window_size: usize,
               best_solutions_capacity: usize, all_solutions_capacity: usize, all_solution_iteration_expiry: u64,
               rng: R, device: DP) -> Self {
        Self {
            move_proposer, solution_score_calculator, max_iterations, window_size,
            _best_solutions_capacity: best_solutions_capacity, _all_solutions_capacity: all_solutions_capacity,
            _all_solution_iteration_expiry: all_solution_iteration_expiry, _rng: rng, device,
        }
    }

    /// local_search.rs:301-342
    pub fn execute(&mut self, start: _Solution, allow_no_improvement_for: u64) -> ScoredSolution<_Solution, _Score> {
        self.device.execute_on_device(&start, allow_no_improvement_for, self.max_iterations, self.window_size)
    }
}

impl<R, _Solution, _Score, SSC, MP, DP> Execute<_Solution, _Score> for B200LocalSearch<R, _Solution, _Score, SSC, MP, DP>
where
    R: rand::Rng,
    _Solution: Solution,
    _Score: Score,
    SSC: SolutionScoreCalculator<_Solution = _Solution, _Score = _Score>,
    MP: MoveProposer<R = R, Solution = _Solution>,
    DP: DeviceProblem<_Solution = _Solution, _Score = _Score>,
{
    fn execute(&mut self, start: _Solution, allow_no_improvement_for: u64) -> ScoredSolution<_Solution, _Score> {
        B200LocalSearch::execute(self, start, allow_no_improvement_for)
    }
}

/// examples/nqueens/src/main.rs:35-93 `get_solution`, local-search part, with the reference's constants
/// (:129-135): `reference_proposer` = the reference's own sampled change-move proposer, window and
/// (score, solution) tie-break run on the device (CS_NQ_FLAG_REFERENCE_PROPOSER).
pub fn nqueens_local_search<R: rand::Rng>(board_size: usize, seed: u64, reference_proposer: bool, rng: R)
    -> B200LocalSearch<R, B200NQueensSolution, B200NQueensScore, B200NQueensScoreCalculator,
                       B200NQueensMoveProposer<R>, NqDevice> {
    let flags = if reference_proposer { ffi::CS_NQ_FLAG_REFERENCE_PROPOSER } else { 0 };
    B200LocalSearch::new(
        B200NQueensMoveProposer::new(board_size, false),
        B200NQueensScoreCalculator::new(board_size),
        10_000,            // max_iterations
        5 * board_size,    // window_size
        32,                // best_solutions_capacity
        100_000,           // all_solutions_capacity
        10_000,            // all_solution_iteration_expiry
        rng,
        NqDevice::new(board_size, ffi::CS_NQ_CHANGE, flags, seed),
    )
}
