// The wasm-bindgen API of the reference (web/employee-scheduling-wasm-bindgen/src/lib.rs:13-84:
// an opaque SolverContext driven one round per message by the web worker,
// web/employee-scheduling/src/worker.ts:8-22) over the B200 host mirror, with the same five entry
// points and the same JSON shapes (:86-110):
//   create_solver(EmployeeSchedulingInput JSON) -> SolverContext
//   execute_solver_round(ctx); get_iteration_info(ctx) -> {"current":..,"total":..};
//   is_solver_finished(ctx); get_best_solution(ctx) -> ScoredSolutionWrapper JSON
// Solver constants are the ones create_solver hard-codes (:34-41).
#ifndef EMPLOYEE_SCHEDULING_SOLVER_CONTEXT_HPP
#define EMPLOYEE_SCHEDULING_SOLVER_CONTEXT_HPP

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

#include "local_search_b200.hpp"

namespace employee_scheduling_wasm_api {

using namespace local_search_b200;
using namespace local_search_b200::employee_scheduling;

// ---- just enough JSON for EmployeeSchedulingInput (objects, arrays, strings, integers) ----
struct Json {
    const std::string& s;
    size_t p = 0;
    explicit Json(const std::string& text) : s(text) {}
    [[noreturn]] void fail(const char* what) { throw std::invalid_argument(std::string("deserializing input failed: ") + what); }
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) ++p; }
    bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { ++p; return true; } return false; }
    void need(char c) { if (!eat(c)) fail("unexpected character"); }
    std::string str() {
        need('"');
        std::string out;
        while (p < s.size() && s[p] != '"') { if (s[p] == '\\') ++p; out += s[p++]; }
        need('"');
        return out;
    }
    int64_t integer() { ws(); char* e = nullptr; const long long v = std::strtoll(s.c_str() + p, &e, 10); if (e == s.c_str() + p) fail("integer expected"); p = e - s.c_str(); return v; }
    void skip() {  // any value
        ws();
        if (p >= s.size()) fail("eof");
        if (s[p] == '"') { str(); return; }
        if (s[p] == '{' || s[p] == '[') {
            const char close = s[p] == '{' ? '}' : ']';
            ++p;
            if (eat(close)) return;
            do { if (close == '}') { str(); need(':'); } skip(); } while (eat(','));
            need(close);
            return;
        }
        while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '-' || s[p] == '+' || s[p] == '.')) ++p;
    }
};

inline MainArgs parse_input(const std::string& text) {
    MainArgs a;
    std::vector<Employee> employees;
    std::vector<std::vector<NaiveDate>> holidays;
    bool have_start = false, have_end = false;
    Json j(text);
    j.need('{');
    do {
        const std::string key = j.str();
        j.need(':');
        if (key == "startDate") { a.start_date = NaiveDate::parse(j.str()); have_start = true; }
        else if (key == "endDate") { a.end_date = NaiveDate::parse(j.str()); have_end = true; }
        else if (key == "employees") {
            j.need('[');
            if (!j.eat(']')) {
                do {
                    j.need('{');
                    Employee e;
                    do { const std::string k = j.str(); j.need(':'); if (k == "id") e.id = j.integer(); else j.skip(); } while (j.eat(','));
                    j.need('}');
                    employees.push_back(e);
                } while (j.eat(','));
                j.need(']');
            }
        } else if (key == "employeeHolidays") {
            j.need('[');
            if (!j.eat(']')) {
                do {
                    holidays.emplace_back();
                    j.need('[');
                    if (!j.eat(']')) { do holidays.back().push_back(NaiveDate::parse(j.str())); while (j.eat(',')); j.need(']'); }
                } while (j.eat(','));
                j.need(']');
            }
        } else j.skip();
    } while (j.eat(','));
    j.need('}');
    if (!have_start || !have_end) j.fail("startDate / endDate missing");
    // itertools::zip(employees, employee_holidays), wasm lib.rs:24-33: the shorter list wins
    for (size_t k = 0; k < employees.size(); ++k) {
        a.employees.insert(employees[k]);
        if (k < holidays.size()) a.employee_to_holidays[employees[k]] = std::set<Holiday>(holidays[k].begin(), holidays[k].end());
    }
    return a;
}


struct SolverContext {
    IlsType solver;
};

// create_solver, wasm lib.rs:22-53 (n_chains: independent ILS chains on the device; 1 = the reference)
inline std::unique_ptr<SolverContext> create_solver(const std::string& input_json, uint32_t n_chains = 1) {
    MainArgs args = parse_input(input_json);  // throws where `input.into_serde().unwrap()` panics
    args.seed = "42";
    args.local_search_max_iterations = 1000;
    args.window_size = 100;
    args.best_solutions_capacity = 64;
    args.all_solutions_capacity = 100000;
    args.all_solution_iteration_expiry = 1000;
    args.iterated_local_search_max_iterations = 250;
    args.max_allow_no_improvement_for = 20;
    args.n_chains = n_chains;
    return std::unique_ptr<SolverContext>(new SolverContext{get_ils(args)});
}

inline void execute_solver_round(SolverContext& ctx) { ctx.solver.execute_round(); }  // :55-58

inline std::string get_iteration_info(const SolverContext& ctx) {  // :60-64, IterationInfo is Serialize
    const IterationInfo i = ctx.solver.get_iteration_info();
    return "{\"current\":" + std::to_string(i.current) + ",\"total\":" + std::to_string(i.total) + "}";
}

inline bool is_solver_finished(const SolverContext& ctx) { return ctx.solver.is_finished(); }  // :66-69

inline std::string get_best_solution(const SolverContext& ctx) {  // :71-84, ScoredSolutionWrapper
    const Scored result = ctx.solver.get_best_solution();
    char buf[96];
    std::snprintf(buf, sizeof buf, "{\"score\":{\"hard_score\":%.1f,\"soft_score\":%.1f},\"days_to_employees\":[",
                  result.score.hard_score, result.score.soft_score);
    std::string out = buf;
    bool first = true;
    for (auto& de : result.solution.get_days_to_employees()) {
        out += std::string(first ? "" : ",") + "[\"" + de.first.format_a_ymd() + "\",{\"id\":" + std::to_string(de.second.id) + "}]";
        first = false;
    }
    return out + "]}";
}

}  // namespace employee_scheduling_wasm_api
#endif
