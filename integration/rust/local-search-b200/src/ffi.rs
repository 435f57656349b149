//! Raw bindings of include/cs_b200.h (ABI version 4; a subset of its entry points).  SOURCE ONLY -- never compiled here.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

pub const CS_OK: i32 = 0;
pub const CS_NQ_SWAP: u32 = 0;
pub const CS_NQ_CHANGE: u32 = 1;
pub const CS_NQ_FLAG_REFERENCE_PROPOSER: u32 = 4;
pub const CS_ES_CHANGE: u32 = 0;
pub const CS_ES_SWAP: u32 = 1;
pub const CS_ES_FLAG_REFERENCE_PROPOSER: u32 = 1;

#[repr(C)] pub struct cs_nq_handle { _p: [u8; 0] }
#[repr(C)] pub struct cs_es_handle { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct cs_move { pub a: u32, pub b: u32 }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct cs_step_stats {
    pub moves_scored: u64, pub steps_accepted: u64, pub best_score: i64, pub best_chain: u32,
    pub chains_at_best: u32, pub device_ms: f32, pub kernel_launches: u32,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct cs_nq_config {
    pub n: u32, pub n_chains: u32, pub chain_offset: u32, pub trace_capacity: u32, pub seed: u64,
    pub device: i32, pub neighbourhood: u32, pub flags: u32,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct cs_es_config {
    pub n_days: u32, pub n_employees: u32, pub start_weekday: u32, pub n_chains: u32,
    pub chain_offset: u32, pub trace_capacity: u32, pub seed: u64, pub device: i32, pub flags: u32,
}

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct cs_es_move { pub kind: u32, pub a: u32, pub b: u32 }

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct cs_es_step_stats {
    pub moves_scored: u64, pub steps_accepted: u64, pub best_hard: i64, pub best_soft: i64,
    pub best_chain: u32, pub chains_at_best: u32, pub chains_feasible: u32, pub device_ms: f32,
    pub kernel_launches: u32,
}

#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct cs_ils_stats {
    pub moves_scored: u64, pub ls_steps: u64, pub best_key: i64, pub best_chain: u32,
    pub chains_done: u32, pub rounds_run: u32, pub device_ms: f32, pub kernel_launches: u32,
}

extern "C" {
    pub fn cs_abi_version() -> i32;
    pub fn cs_device_count() -> i32;
    pub fn cs_philox4x32_10(seed: u64, chain: u32, purpose: u32, counter: u64, out: *mut u32);

    pub fn cs_nq_create(cfg: *const cs_nq_config, out: *mut *mut cs_nq_handle) -> i32;
    pub fn cs_nq_destroy(h: *mut cs_nq_handle) -> i32;
    pub fn cs_nq_last_error(h: *const cs_nq_handle) -> *const c_char;
    pub fn cs_nq_init_random(h: *mut cs_nq_handle) -> i32;
    pub fn cs_nq_set_chains(h: *mut cs_nq_handle, first: u32, count: u32, rows: *const i64) -> i32;
    pub fn cs_nq_get_chains(h: *mut cs_nq_handle, first: u32, count: u32, rows: *mut i64) -> i32;
    pub fn cs_nq_get_scores(h: *mut cs_nq_handle, scores: *mut i64) -> i32;
    pub fn cs_nq_score_full(h: *mut cs_nq_handle, chain: u32, score: *mut i64) -> i32;
    pub fn cs_nq_eval_moves(h: *mut cs_nq_handle, chain: u32, kind: u32, moves: *const cs_move,
                            n: u64, delta: *mut i64) -> i32;
    pub fn cs_nq_enumerate(h: *mut cs_nq_handle, chain: u32, moves: *mut cs_move, cap: u64,
                           n_out: *mut u64) -> i32;
    pub fn cs_nq_neighbourhood_deltas(h: *mut cs_nq_handle, chain: u32, delta: *mut i64, cap: u64,
                                      n_out: *mut u64) -> i32;
    pub fn cs_nq_band_deltas(h: *mut cs_nq_handle, chain: u32, i_begin: u32, i_end: u32, delta: *mut i64,
                             cap: u64, n_out: *mut u64) -> i32;
    pub fn cs_nq_set_window(h: *mut cs_nq_handle, window_size: u64) -> i32;
    pub fn cs_nq_step(h: *mut cs_nq_handle, n_steps: u32, stats: *mut cs_step_stats) -> i32;
    pub fn cs_nq_step_enqueue(h: *mut cs_nq_handle, n_steps: u32) -> i32;
    pub fn cs_nq_step_wait(h: *mut cs_nq_handle, stats: *mut cs_step_stats) -> i32;
    pub fn cs_nq_exchange_select(h: *mut cs_nq_handle, d_key: *const c_void, d_elite_u16: *mut c_void, elite_len: u32) -> i32;
    pub fn cs_nq_local_search(h: *mut cs_nq_handle, allow: u64, max_iterations: u64,
                              stats: *mut cs_step_stats) -> i32;
    pub fn cs_nq_local_search_one(h: *mut cs_nq_handle, start: *const i64, allow: u64,
                                  max_iterations: u64, best: *mut i64, best_score: *mut i64) -> i32;
    pub fn cs_nq_get_trace(h: *mut cs_nq_handle, chain: u32, moves: *mut cs_move,
                           score_after: *mut i64, cap: u64, n_out: *mut u64) -> i32;
    pub fn cs_nq_best(h: *mut cs_nq_handle, rows: *mut i64, score: *mut i64, chain: *mut u32) -> i32;
    pub fn cs_nq_ils_init(h: *mut cs_nq_handle, best_solutions_capacity: u32, log_capacity: u32) -> i32;
    pub fn cs_nq_ils_run(h: *mut cs_nq_handle, rounds: u32, ls_max_iterations: u64, allow: u64,
                         stop_when_any_best: u32, stats: *mut cs_ils_stats) -> i32;
    pub fn cs_nq_ils_get_best(h: *mut cs_nq_handle, chain: u32, rows: *mut i64, score: *mut i64) -> i32;

    pub fn cs_es_create(cfg: *const cs_es_config, employee_ids: *const i64, hol_emp: *const i64,
                        hol_day: *const i64, n_hol: u64, out: *mut *mut cs_es_handle) -> i32;
    pub fn cs_es_create_ex(cfg: *const cs_es_config, employee_ids: *const i64, hol_emp: *const i64,
                           hol_day: *const i64, n_hol: u64, shifts_per_day: u32, skills: *const u32,
                           out: *mut *mut cs_es_handle) -> i32;
    pub fn cs_es_get_dims(h: *mut cs_es_handle, n_days: *mut u32, shifts_per_day: *mut u32, n_slots: *mut u32) -> i32;
    pub fn cs_es_destroy(h: *mut cs_es_handle) -> i32;
    pub fn cs_es_last_error(h: *const cs_es_handle) -> *const c_char;
    pub fn cs_es_init_random(h: *mut cs_es_handle) -> i32;
    pub fn cs_es_set_chains(h: *mut cs_es_handle, first: u32, count: u32, rows: *const i64) -> i32;
    pub fn cs_es_score_full(h: *mut cs_es_handle, chain: u32, hard: *mut i64, soft: *mut i64,
                            terms: *mut i64) -> i32;
    pub fn cs_es_score_full_ex(h: *mut cs_es_handle, chain: u32, hard: *mut i64, soft: *mut i64,
                               terms: *mut i64) -> i32;
    pub fn cs_es_enumerate(h: *mut cs_es_handle, chain: u32, moves: *mut cs_es_move, cap: u64,
                           n_out: *mut u64) -> i32;
    pub fn cs_es_eval_moves(h: *mut cs_es_handle, chain: u32, moves: *const cs_es_move, n_moves: u64,
                            dhard: *mut i64, dsoft: *mut i64) -> i32;
    pub fn cs_es_step(h: *mut cs_es_handle, n_steps: u32, stats: *mut cs_es_step_stats) -> i32;
    pub fn cs_es_step_enqueue(h: *mut cs_es_handle, n_steps: u32) -> i32;
    pub fn cs_es_step_wait(h: *mut cs_es_handle, stats: *mut cs_es_step_stats) -> i32;
    pub fn cs_es_exchange_select(h: *mut cs_es_handle, d_key: *const c_void, d_elite_u16: *mut c_void, elite_len: u32) -> i32;
    pub fn cs_es_set_chains_async(h: *mut cs_es_handle, first: u32, count: u32, rows: *const i64) -> i32;
    pub fn cs_es_commit_chains(h: *mut cs_es_handle) -> i32;
    pub fn cs_es_set_window(h: *mut cs_es_handle, window_size: u64) -> i32;
    pub fn cs_es_local_search_one(h: *mut cs_es_handle, start: *const i64, allow: u64,
                                  max_iterations: u64, best: *mut i64, best_hard: *mut i64,
                                  best_soft: *mut i64) -> i32;
    pub fn cs_es_ils_init(h: *mut cs_es_handle, best_solutions_capacity: u32, log_capacity: u32) -> i32;
    pub fn cs_es_ils_run(h: *mut cs_es_handle, rounds: u32, ls_max_iterations: u64, allow: u64,
                         stop_when_any_best: u32, stats: *mut cs_ils_stats) -> i32;
    pub fn cs_es_ils_get_best(h: *mut cs_es_handle, chain: u32, rows: *mut i64, hard: *mut i64,
                              soft: *mut i64) -> i32;
}

pub unsafe fn last_error_nq(h: *const cs_nq_handle) -> String {
    std::ffi::CStr::from_ptr(cs_nq_last_error(h)).to_string_lossy().into_owned()
}
#[allow(unused)] pub type Opaque = c_void;
