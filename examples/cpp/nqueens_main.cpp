// C++ driver with the reference CLI's surface (examples/nqueens/src/main.rs:95-150):
//   --seed/-s STRING (default "42"), --board-size/-b INT (default 8)
// and its solver constants (:129-135).  get_solution() is the reference's construction sequence
// (:35-93) over the C++ host mirror (include/local_search_b200.hpp) -> C ABI -> CUDA kernels.
// Extensions: --chains N (independent ILS chains on the GPU, default 1 = the reference's shape),
// --full-change (whole change neighbourhood instead of the reference's sampled proposer).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "local_search_b200.hpp"

using namespace local_search_b200;
using namespace local_search_b200::nqueens;

struct MainArgs {
    uint64_t board_size;
    std::string seed;
    uint64_t local_search_max_iterations, window_size;
    size_t best_solutions_capacity, all_solutions_capacity;
    uint64_t all_solution_iteration_expiry, iterated_local_search_max_iterations, max_allow_no_improvement_for;
    uint32_t chains;
    bool full_change;
};

static Scored get_solution(const MainArgs& args, uint64_t* moves_scored) {
    const auto seed = hash_str(args.seed);
    NQueensMoveProposer move_proposer = args.full_change ? NQueensMoveProposer::full_change(args.board_size)
                                                         : NQueensMoveProposer(args.board_size);
    NQueensSolutionScoreCalculator solution_score_calculator;
    PhiloxRng solver_rng = PhiloxRng::from_seed(seed);
    NQueensLocalSearch local_search(move_proposer, solution_score_calculator, args.local_search_max_iterations,
                                    (size_t)args.window_size, args.best_solutions_capacity, args.all_solutions_capacity,
                                    args.all_solution_iteration_expiry, solver_rng, args.chains);
    NQueensInitialSolutionGenerator initial_solution_generator(args.board_size);
    NQueensPerturbation perturbation;
    History<NQueensSolution, NQueensScore> history(args.best_solutions_capacity, args.all_solutions_capacity,
                                                   args.all_solution_iteration_expiry);
    AcceptanceCriterion<NQueensSolution, NQueensScore> acceptance_criterion;
    PhiloxRng iterated_local_search_rng = PhiloxRng::from_seed(seed);
    NQueensIteratedLocalSearch iterated_local_search(
        initial_solution_generator, solution_score_calculator, std::move(local_search), perturbation, history,
        acceptance_criterion, args.iterated_local_search_max_iterations, args.max_allow_no_improvement_for,
        iterated_local_search_rng);
    while (!iterated_local_search.is_finished()) {
        iterated_local_search.execute_round();
        // the reference keeps calling execute_round, which early-outs once the best is_best
        // (iterated_local_search.rs:175-184); skip the remaining no-op rounds
        if (iterated_local_search.last_stats().best_key == 0) break;
    }
    *moves_scored = iterated_local_search.moves_scored();
    return iterated_local_search.get_best_solution();
}

int main(int argc, char** argv) {
    std::printf("local search n-queens example\n");
    std::string seed = "42";
    uint64_t board_size = 8;
    uint32_t chains = 1;
    bool full_change = false;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        auto value = [&](const char* what) -> const char* {
            if (k + 1 >= argc) { std::fprintf(stderr, "error: %s needs a value\n", what); std::exit(2); }
            return argv[++k];
        };
        if (a == "-s" || a == "--seed") seed = value("--seed");
        else if (a == "-b" || a == "--board-size") {
            const char* v = value("--board-size");
            char* end = nullptr;
            board_size = std::strtoull(v, &end, 10);
            if (!*v || *end) { std::fprintf(stderr, "error: invalid digit found in string\n"); return 2; }  // clap validator, :113-118
        } else if (a == "--chains") chains = (uint32_t)std::strtoul(value("--chains"), nullptr, 10);
        else if (a == "--full-change") full_change = true;
        else if (a == "-h" || a == "--help") {
            std::printf("Local Search N-Queens Example 1.0\n\nOPTIONS:\n  -s, --seed <STRING>       Random seeed, any string [default: 42]\n"
                        "  -b, --board-size <INT>    Board size [default: 8]\n      --chains <INT>        ILS chains on the GPU [default: 1]\n"
                        "      --full-change         whole change neighbourhood instead of the sampled proposer\n");
            return 0;
        } else { std::fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); return 2; }
    }
    MainArgs args{board_size, seed, 10000, board_size * 5, 32, 100000, 10000, 10000, 5, chains, full_change};
    try {
        uint64_t moves = 0;
        const Scored result = get_solution(args, &moves);
        if (board_size <= 64) std::printf("result.solution:\n%s\n", result.solution.debug().c_str());
        else {
            std::printf("result.solution rows:");
            for (int64_t r : result.solution.rows) std::printf(" %lld", (long long)r);
            std::printf("\n");
        }
        std::printf("result.score: %s\n", result.score.debug().c_str());
        std::fprintf(stderr, "moves scored on the device: %llu\n", (unsigned long long)moves);
    } catch (const CsError& e) {
        std::fprintf(stderr, "fatal: %s\n", e.what());  // the reference panics
        return 101;
    }
    return 0;
}
