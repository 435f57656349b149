// C++ driver mirroring examples/employee-scheduling/src/main.rs:8-64 (7 employees, 2022-05-09 +
// 30 days, no holidays, solver constants :24-31) and, with --json FILE, the wasm API
// (examples/cpp/solver_context.hpp: create_solver / execute_solver_round / get_iteration_info /
// is_solver_finished / get_best_solution with the reference's JSON shapes):
//   in : {"startDate":"YYYY-MM-DD","endDate":"YYYY-MM-DD","employees":[{"id":0},..],
//         "employeeHolidays":[["YYYY-MM-DD",..],..]}          (EmployeeSchedulingInput)
//   out: {"score":{"hard_score":H,"soft_score":S},
//         "days_to_employees":[["Mon 2022-05-09",{"id":3}],..]} (ScoredSolutionWrapper)
#include <fstream>
#include <sstream>

#include "solver_context.hpp"

using namespace employee_scheduling_wasm_api;

int main(int argc, char** argv) {
    std::string json_path;
    uint32_t chains = 1;
    bool progress = false;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        if (a == "--json" && k + 1 < argc) json_path = argv[++k];
        else if (a == "--chains" && k + 1 < argc) chains = (uint32_t)std::strtoul(argv[++k], nullptr, 10);
        else if (a == "--progress") progress = true;
        else { std::fprintf(stderr, "usage: %s [--json INPUT.json] [--chains N] [--progress]\n", argv[0]); return 2; }
    }
    try {
        if (!json_path.empty()) {  // the web worker's call sequence, worker.ts:8-22
            std::ifstream f(json_path);
            if (!f) { std::fprintf(stderr, "cannot open %s\n", json_path.c_str()); return 2; }
            std::stringstream ss;
            ss << f.rdbuf();
            auto ctx = create_solver(ss.str(), chains);
            while (!is_solver_finished(*ctx)) {
                execute_solver_round(*ctx);
                if (progress) std::fprintf(stderr, "%s\n", get_iteration_info(*ctx).c_str());
                if (ctx->solver.last_stats().best_key == 0) break;  // later rounds are early-out no-ops
            }
            std::printf("%s\n", get_best_solution(*ctx).c_str());
            return 0;
        }
        std::printf("employee scheduling local search example\n");
        MainArgs args;
        args.start_date = NaiveDate::parse("2022-05-09");
        args.end_date = args.start_date + 30;
        for (int64_t id = 0; id < 7; ++id) args.employees.insert(Employee{id});
        args.n_chains = chains;
        IlsType iterated_local_search = get_ils(args);
        while (!iterated_local_search.is_finished()) {
            iterated_local_search.execute_round();
            if (iterated_local_search.last_stats().best_key == 0) break;  // later rounds are early-out no-ops
        }
        const Scored result = iterated_local_search.get_best_solution();
        std::printf("result.solution:\n%s\n", result.solution.debug().c_str());
        std::printf("result.score: %s\n", result.score.debug().c_str());
        std::printf("---\n");
        for (auto& ed : result.solution.get_employees_to_days()) {
            std::printf("employee: Employee { id: %lld }\n", (long long)ed.first.id);
            for (const NaiveDate& d : ed.second) std::printf("%s - %s\n", d.weekday(), d.format_iso().c_str());
            std::printf("---\n");
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "fatal: %s\n", e.what());
        return 101;
    }
    return 0;
}
