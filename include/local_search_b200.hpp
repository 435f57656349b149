// local_search_b200.hpp -- C++17 host side of the B200 move evaluator, above the C ABI
// (include/cs_b200.h, libcs_b200.so).
//
// The reference is Rust (no cargo/rustc in this image), so the host mirror of its trait
// surface is C++: the same type names, constructor argument order, method names and error
// behaviour as
//   local-search/src/local_search.rs:16-90      Solution / Score / ScoredSolution /
//                                               SolutionScoreCalculator / InitialSolutionGenerator /
//                                               MoveProposer
//   local-search/src/local_search.rs:253-343    LocalSearch::{new, execute}
//   local-search/src/local_search.rs:115-248    History::new (capacities only; the bounded
//                                               best-set itself lives on the device)
//   local-search/src/iterated_local_search.rs:96-203  IteratedLocalSearch::{new, execute_round,
//                                               is_finished, get_best_solution, get_iteration_info}
//   examples/nqueens/src/lib.rs                 NQueens{Solution, Score, SolutionScoreCalculator,
//                                               InitialSolutionGenerator, MoveProposer, Perturbation}
//   examples/employee-scheduling/src/lib.rs     Employee, Holiday, Schedule{Solution, Score, ...},
//                                               MainArgs, get_ils
// Every score, neighbourhood, selection and acceptance is computed by the CUDA kernels behind
// the handle; nothing here re-implements them on the CPU and there is NO CPU fallback: without
// libcs_b200.so / a CUDA device every constructor that needs a handle throws CsError
// (the reference panics where this throws: unwrap() at local_search.rs:213,226,
// iterated_local_search.rs:70,166, employee-scheduling lib.rs:275).
//
// Differences a caller can see, all deliberate:
//   * random streams are Philox4x32-10 keyed by the first 8 bytes of hash_str(seed)
//     (the reference keys ChaCha20 with all 32), so trajectories are replayable but not
//     stream-identical to the Rust binary;
//   * a LocalSearch / IteratedLocalSearch may carry `n_chains` independent restart chains
//     (default 1 = the reference's shape); get_best_solution() is the best over the chains.
#ifndef LOCAL_SEARCH_B200_HPP
#define LOCAL_SEARCH_B200_HPP

#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "cs_b200.h"

namespace local_search_b200 {

// ---------------------------------------------------------------------------- errors
struct CsError : std::runtime_error {
    int32_t status;
    CsError(int32_t st, const std::string& where, const char* detail)
        : std::runtime_error(where + ": " + cs_status_string(st) +
                             (detail && *detail ? std::string(" (") + detail + ")" : std::string())),
          status(st) {}
};

// ---------------------------------------------------------------------------- seed hashing
// hash_str, examples/nqueens/src/main.rs:28-33 / examples/employee-scheduling/src/lib.rs:50-55:
// BLAKE2b with a 32-byte digest (RFC 7693, unkeyed) of the seed string.
inline std::array<uint8_t, 32> hash_str(const std::string& input) {
    static const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull,
                                   0xa54ff53a5f1d36f1ull, 0x510e527fade682d1ull, 0x9b05688c2b3e6c1full,
                                   0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    static const uint8_t SIGMA[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    auto rotr = [](uint64_t x, int r) { return (x >> r) | (x << (64 - r)); };
    uint64_t h[8];
    for (int i = 0; i < 8; ++i) h[i] = IV[i];
    h[0] ^= 0x01010000ull ^ 32ull;  // digest length 32, no key, fanout = depth = 1
    const size_t len = input.size();
    size_t off = 0;
    do {
        uint8_t block[128] = {0};
        const size_t take = std::min<size_t>(128, len - off);
        std::memcpy(block, input.data() + off, take);
        off += take;
        const bool last = off >= len;
        uint64_t m[16], v[16];
        for (int i = 0; i < 16; ++i) {
            m[i] = 0;
            for (int b = 7; b >= 0; --b) m[i] = (m[i] << 8) | block[i * 8 + b];
        }
        for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = IV[i]; }
        v[12] ^= (uint64_t)off;  // byte counter (inputs here are far below 2^64)
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint64_t x, uint64_t y) {
            v[a] = v[a] + v[b] + x; v[d] = rotr(v[d] ^ v[a], 32);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y; v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];     v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; ++r) {
            const uint8_t* s = SIGMA[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);   G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);  G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);  G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]); G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
    } while (off < len);
    std::array<uint8_t, 32> out{};
    for (int i = 0; i < 32; ++i) out[i] = (uint8_t)(h[i / 8] >> (8 * (i % 8)));
    return out;
}

// The solver rng: `ChaCha20Rng::from_seed(hash_str(seed))` in the reference
// (examples/nqueens/src/main.rs:39,66).  Here it names a family of Philox4x32-10 streams
// (key = first 8 digest bytes, little endian); the device draws from the same streams and
// next_u32() is their host mirror (cs_philox4x32_10), so any chain replays on the CPU.
struct PhiloxRng {
    uint64_t seed = 42;
    uint32_t chain = 0;
    uint32_t purpose = CS_PHILOX_INIT;
    uint64_t draws = 0;
    static PhiloxRng from_seed(const std::array<uint8_t, 32>& digest) {
        PhiloxRng r;
        r.seed = 0;
        for (int i = 7; i >= 0; --i) r.seed = (r.seed << 8) | digest[i];
        return r;
    }
    static PhiloxRng seed_from_u64(uint64_t s) { PhiloxRng r; r.seed = s; return r; }
    uint32_t next_u32() {
        uint32_t out[4];
        cs_philox4x32_10(seed, chain, purpose, draws / 4, out);
        return out[draws++ % 4];
    }
};

// ---------------------------------------------------------------------------- core types
// ScoredSolution, local_search.rs:29-47: derived Ord = (score, then solution).
template <class _Solution, class _Score>
struct ScoredSolution {
    _Score score;
    _Solution solution;
    ScoredSolution() = default;
    ScoredSolution(_Solution solution_, _Score score_) : score(std::move(score_)), solution(std::move(solution_)) {}
    bool operator==(const ScoredSolution& o) const { return score == o.score && solution == o.solution; }
    bool operator<(const ScoredSolution& o) const {
        if (score < o.score) return true;
        if (o.score < score) return false;
        return solution < o.solution;
    }
};

// History::new(best_solutions_capacity, all_solutions_capacity, all_solution_iteration_expiry),
// local_search.rs:133-150.  Only best_solutions_capacity has an effect: the tabu set is
// {current} whatever the other two say (inverted age test, local_search.rs:182-195).
template <class _Solution, class _Score>
struct History {
    size_t best_solutions_capacity, all_solutions_capacity;
    uint64_t all_solution_iteration_expiry;
    History(size_t best_capacity, size_t all_capacity, uint64_t expiry)
        : best_solutions_capacity(best_capacity), all_solutions_capacity(all_capacity),
          all_solution_iteration_expiry(expiry) {}
};

template <class _Solution, class _Score>
struct AcceptanceCriterion {};  // weights {existing 1, new 5, random best 1}, iterated_local_search.rs:51-71 (on device)

struct IterationInfo { uint64_t current, total; };  // iterated_local_search.rs:90-94

// ---------------------------------------------------------------------------- n-queens
namespace nqueens {

// NQueensSolution, examples/nqueens/src/lib.rs:17-60.  debug() is the fmt::Debug board.
struct NQueensSolution {
    std::vector<int64_t> rows;
    bool operator==(const NQueensSolution& o) const { return rows == o.rows; }
    bool operator<(const NQueensSolution& o) const { return rows < o.rows; }
    std::string debug() const {
        const size_t n = rows.size();
        std::string out;
        for (size_t row = 0; row < n * 2 + 1; ++row) {
            if (row % 2 == 0) {
                out.append(n * 4 + 1, '-');
            } else {
                for (size_t col = 0; col < n; ++col) {
                    out += rows[col] == (int64_t)((row - 1) / 2) ? "| Q " : "|   ";
                    if (col == n - 1) out += "|";
                }
            }
            if (row != n * 2) out += "\n";
        }
        return out;
    }
};

// NQueensScore(i64), lib.rs:63-72
struct NQueensScore {
    int64_t value = 0;
    bool is_best() const { return value == 0; }
    bool operator==(const NQueensScore& o) const { return value == o.value; }
    bool operator<(const NQueensScore& o) const { return value < o.value; }
    std::string debug() const { return "NQueensScore(" + std::to_string(value) + ")"; }
};

using Scored = ScoredSolution<NQueensSolution, NQueensScore>;

struct HandleDeleter { void operator()(cs_nq_handle* h) const { if (h) cs_nq_destroy(h); } };
using Handle = std::unique_ptr<cs_nq_handle, HandleDeleter>;

inline void check(cs_nq_handle* h, int32_t rc, const char* where) {
    if (rc != CS_OK) throw CsError(rc, where, h ? cs_nq_last_error(h) : "");
}

inline Handle make_handle(uint32_t n, uint32_t n_chains, uint64_t seed, uint32_t neighbourhood, uint32_t flags,
                          uint32_t trace_capacity = 0, uint32_t chain_offset = 0, int32_t device = -1) {
    cs_nq_config cfg{};
    cfg.n = n; cfg.n_chains = n_chains; cfg.chain_offset = chain_offset; cfg.trace_capacity = trace_capacity;
    cfg.seed = seed; cfg.device = device; cfg.neighbourhood = neighbourhood; cfg.flags = flags;
    cs_nq_handle* h = nullptr;
    const int32_t rc = cs_nq_create(&cfg, &h);
    if (rc != CS_OK) throw CsError(rc, "cs_nq_create", "");
    return Handle(h);
}

// NQueensSolutionScoreCalculator, lib.rs:121-141: device full re-score by the reference's pair test.
class NQueensSolutionScoreCalculator {
    mutable std::map<size_t, Handle> engines_;  // one 1-chain handle per board size seen
public:
    NQueensSolutionScoreCalculator() = default;
    Scored get_scored_solution(NQueensSolution solution) const {
        const size_t n = solution.rows.size();
        auto it = engines_.find(n);
        if (it == engines_.end()) it = engines_.emplace(n, make_handle((uint32_t)n, 1, 0, CS_NQ_CHANGE, 0)).first;
        cs_nq_handle* h = it->second.get();
        check(h, cs_nq_set_chains(h, 0, 1, solution.rows.data()), "cs_nq_set_chains");
        int64_t score = 0;
        check(h, cs_nq_score_full(h, 0, &score), "cs_nq_score_full");
        return Scored(std::move(solution), NQueensScore{score});
    }
};

// NQueensInitialSolutionGenerator, lib.rs:143-162: shuffle of 0..n, drawn on the device from the
// rng's Philox stream (seed, chain, CS_PHILOX_INIT).
class NQueensInitialSolutionGenerator {
    size_t board_size_;
public:
    explicit NQueensInitialSolutionGenerator(size_t board_size) : board_size_(board_size) {}
    size_t board_size() const { return board_size_; }
    NQueensSolution generate_initial_solution(PhiloxRng& rng) const {
        Handle h = make_handle((uint32_t)board_size_, 1, rng.seed, CS_NQ_CHANGE, 0, 0, rng.chain);
        check(h.get(), cs_nq_init_random(h.get()), "cs_nq_init_random");
        NQueensSolution s;
        s.rows.resize(board_size_);
        check(h.get(), cs_nq_get_chains(h.get(), 0, 1, s.rows.data()), "cs_nq_get_chains");
        return s;
    }
};

// NQueensMoveProposer, lib.rs:164-256.  Default construction = the reference's own proposer
// (sampled conflicted columns x every row value, CS_NQ_FLAG_REFERENCE_PROPOSER); full_swap /
// full_change ask for the whole neighbourhood (the north-star hot path).
class NQueensMoveProposer {
    size_t board_size_;
    uint32_t neighbourhood_, flags_;
public:
    explicit NQueensMoveProposer(size_t board_size)
        : board_size_(board_size), neighbourhood_(CS_NQ_CHANGE), flags_(CS_NQ_FLAG_REFERENCE_PROPOSER) {}
    static NQueensMoveProposer full_swap(size_t n) { NQueensMoveProposer p(n); p.neighbourhood_ = CS_NQ_SWAP; p.flags_ = 0; return p; }
    static NQueensMoveProposer full_change(size_t n) { NQueensMoveProposer p(n); p.flags_ = 0; return p; }
    size_t board_size() const { return board_size_; }
    uint32_t neighbourhood() const { return neighbourhood_; }
    uint32_t flags() const { return flags_; }
    // iter_local_moves: the FULL neighbourhood of this proposer's kind, device enumeration order,
    // identity candidates skipped (host-visible for tests; the hot path never materialises it).
    std::vector<NQueensSolution> iter_local_moves(const NQueensSolution& start, PhiloxRng&) const {
        Handle h = make_handle((uint32_t)board_size_, 1, 0, neighbourhood_, 0);
        check(h.get(), cs_nq_set_chains(h.get(), 0, 1, start.rows.data()), "cs_nq_set_chains");
        uint64_t count = 0;
        check(h.get(), cs_nq_enumerate(h.get(), 0, nullptr, 0, &count), "cs_nq_enumerate");
        std::vector<cs_move> mv(count);
        check(h.get(), cs_nq_enumerate(h.get(), 0, mv.data(), count, &count), "cs_nq_enumerate");
        std::vector<NQueensSolution> out;
        out.reserve(count);
        for (const cs_move& m : mv) {
            NQueensSolution s = start;
            if (neighbourhood_ == CS_NQ_SWAP) std::swap(s.rows[m.a], s.rows[m.b]);
            else s.rows[m.a] = (int64_t)m.b;
            out.push_back(std::move(s));
        }
        return out;
    }
};

struct NQueensPerturbation {};  // lib.rs:258-321, runs on the device inside execute_round

struct Problem {
    using Solution = NQueensSolution;
    using Score = NQueensScore;
    using SSC = NQueensSolutionScoreCalculator;
    using MP = NQueensMoveProposer;
    using ISG = NQueensInitialSolutionGenerator;
    using P = NQueensPerturbation;
    using HandleT = cs_nq_handle;
    using HandlePtr = Handle;
    static size_t solution_len(const HandlePtr&, size_t n) { return n; }
};

}  // namespace nqueens

// ---------------------------------------------------------------------------- LocalSearch (n-queens)
// LocalSearch<R,_Solution,_Score,SSC,MP>, local_search.rs:253-343, n-queens instantiation.
class NQueensLocalSearch {
    nqueens::Handle h_;
    size_t n_;
    uint64_t max_iterations_;
    uint32_t n_chains_;
    friend class NQueensIteratedLocalSearch;
public:
    // LocalSearch::new(move_proposer, solution_score_calculator, max_iterations, window_size,
    //                  best_solutions_capacity, all_solutions_capacity,
    //                  all_solution_iteration_expiry, rng), local_search.rs:277-299
    NQueensLocalSearch(const nqueens::NQueensMoveProposer& move_proposer, const nqueens::NQueensSolutionScoreCalculator&,
                       uint64_t max_iterations, size_t window_size, size_t /*best_solutions_capacity*/,
                       size_t /*all_solutions_capacity*/, uint64_t /*all_solution_iteration_expiry*/, PhiloxRng rng,
                       uint32_t n_chains = 1)
        : h_(nqueens::make_handle((uint32_t)move_proposer.board_size(), n_chains, rng.seed, move_proposer.neighbourhood(),
                                  move_proposer.flags(), 0, rng.chain)),
          n_(move_proposer.board_size()), max_iterations_(max_iterations), n_chains_(n_chains) {
        if (move_proposer.flags() & CS_NQ_FLAG_REFERENCE_PROPOSER)
            nqueens::check(h_.get(), cs_nq_set_window(h_.get(), window_size), "cs_nq_set_window");
    }
    // LocalSearch::execute(start, allow_no_improvement_for) -> best ScoredSolution, :301-342
    nqueens::Scored execute(const nqueens::NQueensSolution& start, uint64_t allow_no_improvement_for) {
        if (start.rows.size() != n_) throw CsError(CS_ERR_INVALID_ARG, "LocalSearch::execute", "board size mismatch");
        nqueens::NQueensSolution best;
        best.rows.resize(n_);
        int64_t score = 0;
        nqueens::check(h_.get(), cs_nq_local_search_one(h_.get(), start.rows.data(), allow_no_improvement_for,
                                                        max_iterations_, best.rows.data(), &score),
                       "cs_nq_local_search_one");
        return nqueens::Scored(std::move(best), nqueens::NQueensScore{score});
    }
    cs_nq_handle* handle() const { return h_.get(); }
    uint32_t n_chains() const { return n_chains_; }
};

// IteratedLocalSearch, iterated_local_search.rs:96-203, n-queens instantiation.
class NQueensIteratedLocalSearch {
    NQueensLocalSearch ls_;
    uint64_t iteration_ = 0, max_iterations_, max_allow_no_improvement_for_;
    cs_ils_stats last_{};
    uint64_t moves_scored_ = 0;
    bool any_round_ = false;
public:
    // IteratedLocalSearch::new(initial_solution_generator, solution_score_calculator, local_search,
    //   perturbation, history, acceptance_criterion, max_iterations, max_allow_no_improvement_for,
    //   rng), :130-156.  `current` := the generator's solution for every chain (device init).
    NQueensIteratedLocalSearch(const nqueens::NQueensInitialSolutionGenerator&, const nqueens::NQueensSolutionScoreCalculator&,
                               NQueensLocalSearch local_search, nqueens::NQueensPerturbation,
                               History<nqueens::NQueensSolution, nqueens::NQueensScore> history,
                               AcceptanceCriterion<nqueens::NQueensSolution, nqueens::NQueensScore>, uint64_t max_iterations,
                               uint64_t max_allow_no_improvement_for, PhiloxRng /*rng: same seed as the handle*/)
        : ls_(std::move(local_search)), max_iterations_(max_iterations),
          max_allow_no_improvement_for_(max_allow_no_improvement_for) {
        cs_nq_handle* h = ls_.h_.get();
        nqueens::check(h, cs_nq_init_random(h), "cs_nq_init_random");
        nqueens::check(h, cs_nq_ils_init(h, (uint32_t)history.best_solutions_capacity, 0), "cs_nq_ils_init");
    }
    IterationInfo get_iteration_info() const { return {iteration_, max_iterations_}; }
    bool is_finished() const { return iteration_ >= max_iterations_; }
    void execute_round() {  // :173-202 (the early-out on a best score happens on the device)
        ++iteration_;
        cs_nq_handle* h = ls_.h_.get();
        nqueens::check(h, cs_nq_ils_run(h, 1, ls_.max_iterations_, max_allow_no_improvement_for_, 0, &last_), "cs_nq_ils_run");
        moves_scored_ += last_.moves_scored;
        any_round_ = true;
    }
    // extension: run up to `rounds` rounds without a host round trip, stop when a chain is_best
    void execute_rounds(uint32_t rounds, bool stop_when_any_best = true) {
        cs_nq_handle* h = ls_.h_.get();
        nqueens::check(h, cs_nq_ils_run(h, rounds, ls_.max_iterations_, max_allow_no_improvement_for_,
                                        stop_when_any_best ? 1u : 0u, &last_), "cs_nq_ils_run");
        iteration_ += last_.rounds_run;
        if (stop_when_any_best && last_.best_key == 0) iteration_ = std::max(iteration_, max_iterations_);
        moves_scored_ += last_.moves_scored;
        any_round_ = true;
    }
    nqueens::Scored get_best_solution() const {  // history.get_best().unwrap(), :165-167
        if (!any_round_) throw CsError(CS_ERR_STATE, "get_best_solution", "no round executed (the reference unwrap()s None)");
        nqueens::NQueensSolution s;
        s.rows.resize(ls_.n_);
        int64_t score = 0;
        nqueens::check(ls_.h_.get(), cs_nq_ils_get_best(ls_.h_.get(), last_.best_chain, s.rows.data(), &score), "cs_nq_ils_get_best");
        return nqueens::Scored(std::move(s), nqueens::NQueensScore{score});
    }
    uint64_t moves_scored() const { return moves_scored_; }
    const cs_ils_stats& last_stats() const { return last_; }
};

// ---------------------------------------------------------------------------- employee scheduling
namespace employee_scheduling {

// chrono::NaiveDate restated: proleptic Gregorian civil date <-> day number.
struct NaiveDate {
    int64_t days = 0;  // days since 1970-01-01
    static NaiveDate from_ymd(int64_t y, unsigned m, unsigned d) {
        y -= m <= 2;
        const int64_t era = (y >= 0 ? y : y - 399) / 400;
        const unsigned yoe = (unsigned)(y - era * 400);
        const unsigned doy = (153 * (m > 2 ? m - 3 : m + 9) + 2) / 5 + d - 1;
        const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
        return NaiveDate{era * 146097 + (int64_t)doe - 719468};
    }
    // NaiveDate::parse_from_str(s, "%Y-%m-%d"); throws where the reference unwrap()s an Err
    static NaiveDate parse(const std::string& s) {
        int y = 0; unsigned m = 0, d = 0;
        if (std::sscanf(s.c_str(), "%d-%u-%u", &y, &m, &d) != 3 || m < 1 || m > 12 || d < 1 || d > 31)
            throw std::invalid_argument("bad date: " + s);
        return from_ymd(y, m, d);
    }
    std::tuple<int64_t, unsigned, unsigned> ymd() const {
        const int64_t z = days + 719468;
        const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
        const unsigned doe = (unsigned)(z - era * 146097);
        const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
        const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
        const unsigned mp = (5 * doy + 2) / 153;
        const unsigned d = doy - (153 * mp + 2) / 5 + 1;
        const unsigned m = mp < 10 ? mp + 3 : mp - 9;
        return {(int64_t)yoe + era * 400 + (m <= 2), m, d};
    }
    unsigned num_days_from_monday() const { return (unsigned)(((days % 7) + 7 + 3) % 7); }  // 1970-01-01 = Thursday
    const char* weekday() const {
        static const char* N[7] = {"Mon", "Tue", "Wed", "Thu", "Fri", "Sat", "Sun"};
        return N[num_days_from_monday()];
    }
    std::string format_iso() const {  // "%Y-%m-%d"
        auto [y, m, d] = ymd();
        char buf[32];
        std::snprintf(buf, sizeof buf, "%04lld-%02u-%02u", (long long)y, m, d);
        return buf;
    }
    std::string format_a_ymd() const { return std::string(weekday()) + " " + format_iso(); }  // "%a %Y-%m-%d"
    NaiveDate operator+(int64_t n) const { return NaiveDate{days + n}; }
    bool operator==(const NaiveDate& o) const { return days == o.days; }
    bool operator<(const NaiveDate& o) const { return days < o.days; }
};

struct Employee {  // lib.rs:119-122
    int64_t id = 0;
    bool operator==(const Employee& o) const { return id == o.id; }
    bool operator<(const Employee& o) const { return id < o.id; }
};
using Holiday = NaiveDate;  // lib.rs:124-125
using EmployeeToHolidays = std::map<Employee, std::set<Holiday>>;

// ScheduleSolution, lib.rs:127-192.  date_to_employee carries the phantom slot past end_date
// (generator quirk, lib.rs:405-412); Eq/Ord look at date_to_employee only.
struct ScheduleSolution {
    NaiveDate start_date, end_date;
    std::vector<Employee> date_to_employee;
    std::vector<Employee> employees;
    bool operator==(const ScheduleSolution& o) const { return date_to_employee == o.date_to_employee; }
    bool operator<(const ScheduleSolution& o) const { return date_to_employee < o.date_to_employee; }
    std::vector<std::pair<NaiveDate, Employee>> get_days_to_employees() const {  // :181-191
        std::vector<std::pair<NaiveDate, Employee>> out;
        for (size_t i = 0; i < date_to_employee.size(); ++i) {
            const NaiveDate d = start_date + (int64_t)i;
            out.emplace_back(d, date_to_employee[i]);
            if (!(d < end_date)) break;
        }
        return out;
    }
    std::map<Employee, std::vector<NaiveDate>> get_employees_to_days() const {  // :170-179
        std::map<Employee, std::vector<NaiveDate>> out;
        for (auto& de : get_days_to_employees()) out[de.second].push_back(de.first);
        return out;
    }
    std::string debug() const {  // :224-236
        std::string out;
        for (auto& de : get_days_to_employees())
            out += std::string(de.first.weekday()) + " " + de.first.format_iso() + " - Employee { id: " +
                   std::to_string(de.second.id) + " }\n";
        return out;
    }
};

// ScheduleScore{hard_score, soft_score: OrderedFloat<f64>}, lib.rs:238-249 (integers held in f64)
struct ScheduleScore {
    double hard_score = 0, soft_score = 0;
    bool is_best() const { return hard_score == 0.0 && soft_score == 0.0; }
    bool operator==(const ScheduleScore& o) const { return hard_score == o.hard_score && soft_score == o.soft_score; }
    bool operator<(const ScheduleScore& o) const {
        return hard_score < o.hard_score || (hard_score == o.hard_score && soft_score < o.soft_score);
    }
    std::string debug() const {
        char buf[96];
        std::snprintf(buf, sizeof buf, "ScheduleScore { hard_score: OrderedFloat(%.1f), soft_score: OrderedFloat(%.1f) }",
                      hard_score, soft_score);
        return buf;
    }
};
using Scored = ScoredSolution<ScheduleSolution, ScheduleScore>;

struct HandleDeleter { void operator()(cs_es_handle* h) const { if (h) cs_es_destroy(h); } };
using Handle = std::unique_ptr<cs_es_handle, HandleDeleter>;
inline void check(cs_es_handle* h, int32_t rc, const char* where) {
    if (rc != CS_OK) throw CsError(rc, where, h ? cs_es_last_error(h) : "");
}

inline Handle make_handle(NaiveDate start, NaiveDate end, const std::vector<Employee>& employees,
                          const EmployeeToHolidays& holidays, uint32_t n_chains, uint64_t seed, uint32_t chain_offset = 0,
                          uint32_t flags = 0) {
    cs_es_config cfg{};
    cfg.n_days = (uint32_t)(end.days - start.days + 1);
    cfg.n_employees = (uint32_t)employees.size();
    cfg.start_weekday = start.num_days_from_monday();
    cfg.n_chains = n_chains; cfg.chain_offset = chain_offset; cfg.seed = seed; cfg.device = -1; cfg.flags = flags;
    std::vector<int64_t> ids, he, hd;
    for (const Employee& e : employees) ids.push_back(e.id);
    for (auto& kv : holidays)
        for (const Holiday& d : kv.second) { he.push_back(kv.first.id); hd.push_back(d.days - start.days); }
    cs_es_handle* h = nullptr;
    const int32_t rc = cs_es_create(&cfg, ids.data(), he.data(), hd.data(), he.size(), &h);
    if (rc != CS_OK) throw CsError(rc, "cs_es_create", "holiday outside the rota, too many days, or no device");
    return Handle(h);
}

inline std::vector<int64_t> ids_of(const std::vector<Employee>& v) {
    std::vector<int64_t> out;
    for (const Employee& e : v) out.push_back(e.id);
    return out;
}

// ScheduleSolutionScoreCalculator::new(employee_to_holidays), lib.rs:251-375
class ScheduleSolutionScoreCalculator {
    EmployeeToHolidays holidays_;
public:
    explicit ScheduleSolutionScoreCalculator(EmployeeToHolidays employee_to_holidays) : holidays_(std::move(employee_to_holidays)) {}
    const EmployeeToHolidays& holidays() const { return holidays_; }
    Scored get_scored_solution(ScheduleSolution solution) const {
        Handle h = make_handle(solution.start_date, solution.end_date, solution.employees, holidays_, 1, 0);
        std::vector<int64_t> rows = ids_of(solution.date_to_employee);
        rows.resize((size_t)(solution.end_date.days - solution.start_date.days + 2), rows.empty() ? 0 : rows.back());
        check(h.get(), cs_es_set_chains(h.get(), 0, 1, rows.data()), "cs_es_set_chains");
        int64_t hard = 0, soft = 0;
        check(h.get(), cs_es_score_full(h.get(), 0, &hard, &soft, nullptr), "cs_es_score_full");
        return Scored(std::move(solution), ScheduleScore{(double)hard, (double)soft});
    }
};

// ScheduleInitialSolutionGenerator::new(start_date, end_date, employees, employee_to_holidays), lib.rs:377-420
class ScheduleInitialSolutionGenerator {
public:
    NaiveDate start_date, end_date;
    std::vector<Employee> employees;
    EmployeeToHolidays employee_to_holidays;
    ScheduleInitialSolutionGenerator(NaiveDate s, NaiveDate e, std::vector<Employee> emp, EmployeeToHolidays hol)
        : start_date(s), end_date(e), employees(std::move(emp)), employee_to_holidays(std::move(hol)) {
        std::sort(employees.begin(), employees.end());
    }
    ScheduleSolution generate_initial_solution(PhiloxRng& rng) const {
        Handle h = make_handle(start_date, end_date, employees, employee_to_holidays, 1, rng.seed, rng.chain);
        check(h.get(), cs_es_init_random(h.get()), "cs_es_init_random");
        std::vector<int64_t> rows((size_t)(end_date.days - start_date.days + 2));
        check(h.get(), cs_es_get_chains(h.get(), 0, 1, rows.data()), "cs_es_get_chains");
        ScheduleSolution s{start_date, end_date, {}, employees};
        for (int64_t id : rows) s.date_to_employee.push_back(Employee{id});
        return s;
    }
};

// ScheduleRandomMoveProposer (lib.rs:429-491, the one get_ils installs, :60) = the reference's
// sampled window on the device (CS_ES_FLAG_REFERENCE_PROPOSER); ScheduleMoveProposer
// (lib.rs:493-559, exhaustive) = the FULL change + swap neighbourhood, the north-star hot path.
struct ScheduleRandomMoveProposer { static constexpr uint32_t flags = CS_ES_FLAG_REFERENCE_PROPOSER; };
struct ScheduleMoveProposer { std::vector<Employee> employees; static constexpr uint32_t flags = 0; };
struct SchedulePerturbation {};

}  // namespace employee_scheduling

// LocalSearch, scheduling instantiation (local_search.rs:253-343)
class ScheduleLocalSearch {
    employee_scheduling::Handle h_;
    employee_scheduling::NaiveDate start_, end_;
    std::vector<employee_scheduling::Employee> employees_;
    uint64_t max_iterations_;
    friend class ScheduleIteratedLocalSearch;
    size_t slots() const { return (size_t)(end_.days - start_.days + 2); }
    employee_scheduling::ScheduleSolution wrap(const std::vector<int64_t>& rows) const {
        employee_scheduling::ScheduleSolution s{start_, end_, {}, employees_};
        for (int64_t id : rows) s.date_to_employee.push_back(employee_scheduling::Employee{id});
        return s;
    }
public:
    // LocalSearch::new(...) + what get_ils knows about the rota (lib.rs:57-81): the device needs
    // the calendar and the employee table up front.
    template <class MoveProposerT>
    ScheduleLocalSearch(const MoveProposerT&, const employee_scheduling::ScheduleSolutionScoreCalculator& ssc,
                        uint64_t max_iterations, size_t window_size, size_t /*best_solutions_capacity*/,
                        size_t /*all_solutions_capacity*/, uint64_t /*all_solution_iteration_expiry*/, PhiloxRng rng,
                        employee_scheduling::NaiveDate start_date, employee_scheduling::NaiveDate end_date,
                        std::vector<employee_scheduling::Employee> employees, uint32_t n_chains = 1)
        : start_(start_date), end_(end_date), employees_(std::move(employees)), max_iterations_(max_iterations) {
        std::sort(employees_.begin(), employees_.end());
        h_ = employee_scheduling::make_handle(start_, end_, employees_, ssc.holidays(), n_chains, rng.seed, rng.chain,
                                              MoveProposerT::flags);
        if (MoveProposerT::flags & CS_ES_FLAG_REFERENCE_PROPOSER)
            employee_scheduling::check(h_.get(), cs_es_set_window(h_.get(), window_size), "cs_es_set_window");
    }
    employee_scheduling::Scored execute(const employee_scheduling::ScheduleSolution& start, uint64_t allow_no_improvement_for) {
        std::vector<int64_t> in = employee_scheduling::ids_of(start.date_to_employee), best(slots());
        if (in.size() != slots()) throw CsError(CS_ERR_INVALID_ARG, "LocalSearch::execute", "date_to_employee must hold n_days + 1 slots");
        int64_t hard = 0, soft = 0;
        employee_scheduling::check(h_.get(), cs_es_local_search_one(h_.get(), in.data(), allow_no_improvement_for, max_iterations_,
                                                                    best.data(), &hard, &soft), "cs_es_local_search_one");
        return employee_scheduling::Scored(wrap(best), employee_scheduling::ScheduleScore{(double)hard, (double)soft});
    }
    cs_es_handle* handle() const { return h_.get(); }
};

// IteratedLocalSearch, scheduling instantiation == employee_scheduling::IlsType (lib.rs:40-48)
class ScheduleIteratedLocalSearch {
    ScheduleLocalSearch ls_;
    uint64_t iteration_ = 0, max_iterations_, max_allow_no_improvement_for_;
    cs_ils_stats last_{};
    uint64_t moves_scored_ = 0;
    bool any_round_ = false;
public:
    ScheduleIteratedLocalSearch(const employee_scheduling::ScheduleInitialSolutionGenerator&,
                                const employee_scheduling::ScheduleSolutionScoreCalculator&, ScheduleLocalSearch local_search,
                                employee_scheduling::SchedulePerturbation,
                                History<employee_scheduling::ScheduleSolution, employee_scheduling::ScheduleScore> history,
                                AcceptanceCriterion<employee_scheduling::ScheduleSolution, employee_scheduling::ScheduleScore>,
                                uint64_t max_iterations, uint64_t max_allow_no_improvement_for, PhiloxRng)
        : ls_(std::move(local_search)), max_iterations_(max_iterations),
          max_allow_no_improvement_for_(max_allow_no_improvement_for) {
        cs_es_handle* h = ls_.h_.get();
        employee_scheduling::check(h, cs_es_init_random(h), "cs_es_init_random");
        employee_scheduling::check(h, cs_es_ils_init(h, (uint32_t)history.best_solutions_capacity, 0), "cs_es_ils_init");
    }
    IterationInfo get_iteration_info() const { return {iteration_, max_iterations_}; }
    bool is_finished() const { return iteration_ >= max_iterations_; }
    void execute_round() {
        ++iteration_;
        cs_es_handle* h = ls_.h_.get();
        employee_scheduling::check(h, cs_es_ils_run(h, 1, ls_.max_iterations_, max_allow_no_improvement_for_, 0, &last_), "cs_es_ils_run");
        moves_scored_ += last_.moves_scored;
        any_round_ = true;
    }
    employee_scheduling::Scored get_best_solution() const {
        if (!any_round_) throw CsError(CS_ERR_STATE, "get_best_solution", "no round executed (the reference unwrap()s None)");
        std::vector<int64_t> rows(ls_.slots());
        int64_t hard = 0, soft = 0;
        employee_scheduling::check(ls_.h_.get(), cs_es_ils_get_best(ls_.h_.get(), last_.best_chain, rows.data(), &hard, &soft),
                                   "cs_es_ils_get_best");
        return employee_scheduling::Scored(ls_.wrap(rows), employee_scheduling::ScheduleScore{(double)hard, (double)soft});
    }
    uint64_t moves_scored() const { return moves_scored_; }
    const cs_ils_stats& last_stats() const { return last_; }
};

namespace employee_scheduling {

using IlsType = ScheduleIteratedLocalSearch;

struct MainArgs {  // lib.rs:26-48
    NaiveDate start_date, end_date;
    std::set<Employee> employees;
    EmployeeToHolidays employee_to_holidays;
    std::string seed = "42";
    uint64_t local_search_max_iterations = 1000;
    uint64_t window_size = 100;
    size_t best_solutions_capacity = 64;
    size_t all_solutions_capacity = 100000;
    uint64_t all_solution_iteration_expiry = 1000;
    uint64_t iterated_local_search_max_iterations = 250;
    uint64_t max_allow_no_improvement_for = 20;
    uint32_t n_chains = 1;  // extension: independent ILS chains on the device
};

// get_ils, lib.rs:57-117 -- the same construction sequence
inline IlsType get_ils(const MainArgs& args) {
    const auto seed = hash_str(args.seed);
    std::vector<Employee> employees(args.employees.begin(), args.employees.end());
    ScheduleRandomMoveProposer move_proposer;
    ScheduleSolutionScoreCalculator solution_score_calculator(args.employee_to_holidays);
    PhiloxRng solver_rng = PhiloxRng::from_seed(seed);
    ScheduleLocalSearch local_search(move_proposer, solution_score_calculator, args.local_search_max_iterations,
                                     (size_t)args.window_size, args.best_solutions_capacity, args.all_solutions_capacity,
                                     args.all_solution_iteration_expiry, solver_rng, args.start_date, args.end_date, employees,
                                     args.n_chains);
    ScheduleInitialSolutionGenerator initial_solution_generator(args.start_date, args.end_date, employees, args.employee_to_holidays);
    History<ScheduleSolution, ScheduleScore> history(args.best_solutions_capacity, args.all_solutions_capacity,
                                                     args.all_solution_iteration_expiry);
    return IlsType(initial_solution_generator, solution_score_calculator, std::move(local_search), SchedulePerturbation{}, history,
                   AcceptanceCriterion<ScheduleSolution, ScheduleScore>{}, args.iterated_local_search_max_iterations,
                   args.max_allow_no_improvement_for, PhiloxRng::from_seed(seed));
}

}  // namespace employee_scheduling
}  // namespace local_search_b200
#endif
