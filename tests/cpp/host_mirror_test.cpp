// Tests of the C++ host mirror (include/local_search_b200.hpp) -- written to read like the
// reference's own tests: examples/nqueens/src/lib.rs:94-119 (score known answers) and
// examples/nqueens/src/main.rs:157-200 (`repeatable`).  The CPU oracle (oracle/cs_oracle.h) is
// linked here as the CHECKER only.
//   host_mirror_test --cpu   host-only logic, no device needed (and asserts the loud failure)
//   host_mirror_test --gpu   parity through the C ABI on cuda:0
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cs_oracle.h"
#include "local_search_b200.hpp"

using namespace local_search_b200;
namespace nq = local_search_b200::nqueens;
namespace es = local_search_b200::employee_scheduling;

static int g_checks = 0;
#define ASSERT_TRUE(cond)                                                                  \
    do {                                                                                   \
        ++g_checks;                                                                        \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
    } while (0)
#define ASSERT_EQ(a, b) ASSERT_TRUE((a) == (b))

static std::string hex(const std::array<uint8_t, 32>& d) {
    std::string s;
    char b[3];
    for (uint8_t x : d) { std::snprintf(b, sizeof b, "%02x", x); s += b; }
    return s;
}

static void cpu_tests() {
    // hash_str: BLAKE2b-256 known answers (RFC 7693 "abc"; the others are checked against
    // hashlib.blake2b by tests/test_cpp_host.py, which parses these lines)
    ASSERT_EQ(hex(hash_str("abc")), std::string("bddd813c634239723171ef3fee98579b94964e3bb1cb3e427262c8c068d52319"));
    for (const char* s : {"", "42", "43", "abc", "the quick brown fox jumps over the lazy dog, the quick brown fox jumps over the lazy dog, "
                                                 "the quick brown fox jumps over the lazy dog -- longer than one 128-byte block"})
        std::printf("blake2b256 %zu %s\n", std::strlen(s), hex(hash_str(s)).c_str());

    // NaiveDate: 2022-05-09 is the Monday the reference driver starts on (main.rs:11)
    const es::NaiveDate d = es::NaiveDate::parse("2022-05-09");
    ASSERT_EQ(d.num_days_from_monday(), 0u);
    ASSERT_EQ(d.format_a_ymd(), std::string("Mon 2022-05-09"));
    ASSERT_EQ((d + 30).format_iso(), std::string("2022-06-08"));
    ASSERT_EQ(es::NaiveDate::parse("1970-01-01").days, 0);
    ASSERT_EQ(es::NaiveDate::parse("2000-02-29").days, 11016);
    for (int64_t k = -800000; k < 800000; k += 997) {
        const es::NaiveDate x{k};
        auto [y, m, dd] = x.ymd();
        ASSERT_EQ(es::NaiveDate::from_ymd(y, m, dd).days, k);
        ASSERT_EQ((int)x.num_days_from_monday(), orc_weekday_from_days(k));
    }
    // get_days_to_employees stops at end_date: the phantom slot is never listed (lib.rs:181-191)
    es::ScheduleSolution s{d, d + 2, {{0}, {1}, {2}, {9}}, {{0}, {1}, {2}, {9}}};
    ASSERT_EQ(s.get_days_to_employees().size(), (size_t)3);
    ASSERT_EQ(s.get_employees_to_days().count(es::Employee{9}), (size_t)0);

    // derived Ord: score first, then the solution vector (local_search.rs:29-37)
    nq::Scored a({{0, 3, 0, 0}}, {6}), b({{0, 0, 3, 0}}, {6}), c({{9, 9, 9, 9}}, {4});
    ASSERT_TRUE(b < a);
    ASSERT_TRUE(c < b);
    ASSERT_TRUE(!(a < a));
    ASSERT_TRUE((es::ScheduleScore{0, 7} < es::ScheduleScore{1, 0}));
    ASSERT_TRUE((es::ScheduleScore{0, 0}.is_best()));

    // board pretty-printer (lib.rs:26-60): 2n+1 lines, n queens, last line has no newline
    const nq::NQueensSolution sol{{1, 3, 0, 2}};
    const std::string board = sol.debug();
    ASSERT_EQ(std::count(board.begin(), board.end(), 'Q'), 4);
    ASSERT_EQ(std::count(board.begin(), board.end(), '\n'), 8);
    ASSERT_EQ(board.substr(0, 18), std::string("-----------------\n"));
    ASSERT_EQ(board.substr(18, 18), std::string("|   |   | Q |   |\n"));  // row 0 holds column 2's queen

    // host Philox mirror == oracle Philox
    PhiloxRng rng = PhiloxRng::seed_from_u64(42);
    rng.chain = 5;
    for (uint64_t t = 0; t < 37; ++t) ASSERT_EQ(rng.next_u32(), orc_philox_draw(42, 5, CS_PHILOX_INIT, t));

    if (cs_device_count() == 0) {  // the product fails loudly without a device (no CPU fallback)
        bool threw = false;
        try {
            nq::NQueensSolutionScoreCalculator().get_scored_solution({{0, 0, 0, 0}});
        } catch (const CsError& e) {
            threw = e.status == CS_ERR_NO_DEVICE;
        }
        ASSERT_TRUE(threw);
    }
}

static nq::Scored get_solution(uint64_t board_size, const std::string& seed_str, uint64_t rounds, uint64_t all_expiry) {
    const auto seed = hash_str(seed_str);
    nq::NQueensMoveProposer move_proposer(board_size);
    nq::NQueensSolutionScoreCalculator ssc;
    NQueensLocalSearch local_search(move_proposer, ssc, 10000, board_size * 5, 32, 100000, all_expiry, PhiloxRng::from_seed(seed));
    nq::NQueensInitialSolutionGenerator isg(board_size);
    History<nq::NQueensSolution, nq::NQueensScore> history(32, 100000, all_expiry);
    NQueensIteratedLocalSearch ils(isg, ssc, std::move(local_search), nq::NQueensPerturbation{}, history, {}, rounds, 5,
                                   PhiloxRng::from_seed(seed));
    while (!ils.is_finished()) ils.execute_round();
    return ils.get_best_solution();
}

static void gpu_tests() {
    ASSERT_TRUE(cs_device_count() > 0);
    nq::NQueensSolutionScoreCalculator ssc;
    // examples/nqueens/src/lib.rs:94-105 and :108-119
    ASSERT_EQ(ssc.get_scored_solution({{0, 0, 0, 0}}).score.value, 12);
    ASSERT_EQ(ssc.get_scored_solution({{1, 3, 0, 2}}).score.value, 0);
    ASSERT_TRUE(ssc.get_scored_solution({{1, 3, 0, 2}}).score.is_best());
    // SURVEY 8(c) derived vectors + random boards against the oracle's pair loop
    ASSERT_EQ(ssc.get_scored_solution({{0, 2, 4, 6, 1, 3, 5, 7}}).score.value, 2);
    ASSERT_EQ(ssc.get_scored_solution({{3, 1, 4, 1, 5, 9, 2, 6, 5, 3}}).score.value, 8);
    for (uint32_t k = 0; k < 20; ++k) {
        const size_t n = 5 + 13 * k;
        nq::NQueensSolution s;
        for (size_t c = 0; c < n; ++c) s.rows.push_back(orc_philox_draw(7, k, 9, c) % n);
        ASSERT_EQ(ssc.get_scored_solution(s).score.value, orc_nq_score(s.rows.data(), (int64_t)n));
    }
    // generate_initial_solution == the oracle's Fisher-Yates over the same Philox stream
    {
        PhiloxRng rng = PhiloxRng::seed_from_u64(1234);
        rng.chain = 3;
        const nq::NQueensSolution s = nq::NQueensInitialSolutionGenerator(50).generate_initial_solution(rng);
        std::vector<int64_t> want(50);
        orc_nq_init_perm(1234, 3, 50, want.data());
        ASSERT_EQ(s.rows, want);
    }
    // iter_local_moves: full swap neighbourhood; every candidate re-scored == oracle clone+re-score
    {
        PhiloxRng rng;
        const nq::NQueensSolution start{{2, 0, 3, 1, 5, 4}};
        const auto cands = nq::NQueensMoveProposer::full_swap(6).iter_local_moves(start, rng);
        ASSERT_EQ(cands.size(), (size_t)15);
        for (const auto& c : cands) ASSERT_EQ(ssc.get_scored_solution(c).score.value, orc_nq_score(c.rows.data(), 6));
    }
    // LocalSearch::execute (full swap / full change) == oracle's restatement of local_search.rs:301-342
    for (int kind : {ORC_NQ_SWAP, ORC_NQ_CHANGE}) {
        const size_t n = 40;
        nq::NQueensSolution start;
        start.rows.resize(n);
        orc_nq_init_perm(99, 1, (int64_t)n, start.rows.data());
        auto mp = kind == ORC_NQ_SWAP ? nq::NQueensMoveProposer::full_swap(n) : nq::NQueensMoveProposer::full_change(n);
        NQueensLocalSearch ls(mp, ssc, 1000, 5 * n, 32, 100000, 10000, PhiloxRng::seed_from_u64(99));
        const nq::Scored got = ls.execute(start, 5);
        std::vector<int64_t> want = start.rows;
        int64_t want_score = 0;
        orc_nq_local_search(want.data(), (int64_t)n, kind, ORC_TIE_MOVE_ORDER, 5, 1000, 0, &want_score, nullptr, nullptr, nullptr,
                            nullptr, nullptr, 0);
        ASSERT_EQ(got.score.value, want_score);
        ASSERT_EQ(got.solution.rows, want);
    }
    // `repeatable`, examples/nqueens/src/main.rs:157-200: n = 8 reaches 0 for seeds "42".."49" and
    // every repeat returns the same solution; plus: the same solution as the oracle's ILS
    // restatement over the same Philox streams (reference proposer, window 5n, derived-Ord ties)
    for (int seed = 42; seed < 50; ++seed) {
        const nq::Scored first = get_solution(8, std::to_string(seed), 200, 1000);
        for (int i = 1; i < 3; ++i) ASSERT_TRUE(first == get_solution(8, std::to_string(seed), 200, 1000));
        ASSERT_EQ(first.score.value, 0);
        const PhiloxRng rng = PhiloxRng::from_seed(hash_str(std::to_string(seed)));
        std::vector<int64_t> best(8), cur(8), rn(200), rc(200);
        int64_t best_score = -1;
        orc_nq_ils(rng.seed, 0, 8, ORC_NQ_CHANGE, 10000, 5, 200, 32, best.data(), &best_score, cur.data(), rn.data(), rc.data(), 40);
        ASSERT_EQ(best_score, 0);
        ASSERT_EQ(first.solution.rows, best);
    }
    // get_best_solution before any round: the reference unwrap()s None (iterated_local_search.rs:166)
    {
        nq::NQueensMoveProposer mp(8);
        NQueensLocalSearch ls(mp, ssc, 100, 40, 32, 1000, 1000, PhiloxRng{});
        NQueensIteratedLocalSearch ils(nq::NQueensInitialSolutionGenerator(8), ssc, std::move(ls), {}, {32, 1000, 1000}, {}, 10, 5, PhiloxRng{});
        bool threw = false;
        try { ils.get_best_solution(); } catch (const CsError& e) { threw = e.status == CS_ERR_STATE; }
        ASSERT_TRUE(threw);
        ASSERT_EQ(ils.get_iteration_info().total, (uint64_t)10);
    }

    // ---- employee scheduling: SURVEY 8(c) known answers (start 2022-05-09, 31 scored days) ----
    const es::NaiveDate start = es::NaiveDate::parse("2022-05-09"), end = start + 30;
    std::vector<es::Employee> emp;
    for (int64_t id = 0; id < 7; ++id) emp.push_back({id});
    auto rota = [&](auto f) {
        es::ScheduleSolution s{start, end, {}, emp};
        for (int i = 0; i < 32; ++i) s.date_to_employee.push_back({f(i)});
        return s;
    };
    {
        es::ScheduleSolutionScoreCalculator calc({});
        auto sc = calc.get_scored_solution(rota([](int) { return (int64_t)0; })).score;
        ASSERT_EQ(sc.hard_score, 60.0);
        ASSERT_EQ(sc.soft_score, 25.0);
        sc = calc.get_scored_solution(rota([](int i) { return (int64_t)(i % 7); })).score;
        ASSERT_EQ(sc.hard_score, 6.0);
        ASSERT_EQ(sc.soft_score, 5.0);
        sc = calc.get_scored_solution(rota([](int i) { return (int64_t)(i % 2); })).score;
        ASSERT_EQ(sc.hard_score, 42.0);
        ASSERT_EQ(sc.soft_score, 61.0);
        es::EmployeeToHolidays hol;
        hol[{0}] = {es::NaiveDate::parse("2022-05-09"), es::NaiveDate::parse("2022-05-10")};
        hol[{3}] = {es::NaiveDate::parse("2022-05-12")};
        sc = es::ScheduleSolutionScoreCalculator(hol).get_scored_solution(rota([](int i) { return (int64_t)(i % 7); })).score;
        ASSERT_EQ(sc.hard_score, 8.0);
        ASSERT_EQ(sc.soft_score, 5.0);
        // a holiday outside the rota: the reference unwrap()s None (lib.rs:275)
        es::EmployeeToHolidays bad;
        bad[{1}] = {es::NaiveDate::parse("2023-01-01")};
        bool threw = false;
        try { es::ScheduleSolutionScoreCalculator(bad).get_scored_solution(rota([](int) { return (int64_t)0; })); }
        catch (const CsError& e) { threw = e.status == CS_ERR_INVALID_ARG; }
        ASSERT_TRUE(threw);
    }
    // LocalSearch::execute == oracle; get_ils reaches a feasible rota
    {
        es::EmployeeToHolidays hol;
        hol[{2}] = {start + 4, start + 11};
        es::ScheduleSolutionScoreCalculator calc(hol);
        // ScheduleMoveProposer = the exhaustive neighbourhood (lib.rs:493-559)
        ScheduleLocalSearch ls(es::ScheduleMoveProposer{emp}, calc, 1000, 100, 64, 100000, 1000, PhiloxRng::seed_from_u64(5), start,
                               end, emp);
        const es::ScheduleSolution s0 = rota([](int i) { return (int64_t)((i * 5 + i / 3) % 7); });
        const es::Scored got = ls.execute(s0, 20);
        std::vector<int64_t> a = es::ids_of(s0.date_to_employee), ids = es::ids_of(emp);
        const int64_t he[2] = {2, 2}, hd[2] = {4, 11};
        int64_t bh = 0, bs = 0;
        orc_es_local_search(a.data(), 31, 0, he, hd, 2, ids.data(), 7, 20, 1000, &bh, &bs, nullptr, nullptr, nullptr, nullptr, nullptr,
                            nullptr, 0);
        ASSERT_EQ(got.score.hard_score, (double)bh);
        ASSERT_EQ(got.score.soft_score, (double)bs);
        ASSERT_EQ(es::ids_of(got.solution.date_to_employee), a);

        // ScheduleRandomMoveProposer = the reference's own sampled window (lib.rs:440-491), against
        // the oracle's literal restatement over the same Philox stream
        {
            ScheduleLocalSearch rls(es::ScheduleRandomMoveProposer{}, calc, 200, 100, 64, 100000, 1000, PhiloxRng::seed_from_u64(5),
                                    start, end, emp);
            const es::Scored rgot = rls.execute(s0, 20);
            std::vector<int64_t> ra = es::ids_of(s0.date_to_employee);
            int64_t rh = 0, rs = 0;
            orc_es_local_search_ref(ra.data(), 31, 0, he, hd, 2, ids.data(), 7, 5, 0, 20, 200, 100, 1u << 16, &rh, &rs, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
            ASSERT_EQ(rgot.score.hard_score, (double)rh);
            ASSERT_EQ(rgot.score.soft_score, (double)rs);
            std::vector<int64_t> got_ids = es::ids_of(rgot.solution.date_to_employee);
            got_ids.resize(31);
            ra.resize(31);
            ASSERT_EQ(got_ids, ra);
        }
        es::MainArgs args;
        args.start_date = start;
        args.end_date = end;
        args.employees = std::set<es::Employee>(emp.begin(), emp.end());
        args.employee_to_holidays = hol;
        args.iterated_local_search_max_iterations = 20;
        es::IlsType ils = es::get_ils(args);
        while (!ils.is_finished()) ils.execute_round();
        const es::Scored best = ils.get_best_solution();
        ASSERT_EQ(best.score.hard_score, 0.0);
        ASSERT_EQ(best.solution.date_to_employee.size(), (size_t)32);
        const es::Scored again = calc.get_scored_solution(best.solution);  // returned score == device full re-score
        ASSERT_TRUE(again.score == best.score);
    }
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "--cpu";
    try {
        cpu_tests();
        if (mode == "--gpu") gpu_tests();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "FAILED with exception: %s\n", e.what());
        return 1;
    }
    std::printf("host_mirror_test %s: %d checks passed\n", mode.c_str(), g_checks);
    return 0;
}
