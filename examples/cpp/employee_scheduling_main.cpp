// C++ driver mirroring examples/employee-scheduling/src/main.rs:8-64 (7 employees, 2022-05-09 +
// 30 days, no holidays, solver constants :24-31) and, with --json FILE, the wasm API's JSON
// shapes (web/employee-scheduling-wasm-bindgen/src/lib.rs:86-110):
//   in : {"startDate":"YYYY-MM-DD","endDate":"YYYY-MM-DD","employees":[{"id":0},..],
//         "employeeHolidays":[["YYYY-MM-DD",..],..]}          (EmployeeSchedulingInput)
//   out: {"score":{"hard_score":H,"soft_score":S},
//         "days_to_employees":[["Mon 2022-05-09",{"id":3}],..]} (ScoredSolutionWrapper)
// and the wasm call sequence create_solver / execute_solver_round / is_solver_finished /
// get_best_solution (:22-84) over get_ils.
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

#include "local_search_b200.hpp"

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

static MainArgs parse_input(const std::string& text) {
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

int main(int argc, char** argv) {
    std::string json_path;
    uint32_t chains = 1;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        if (a == "--json" && k + 1 < argc) json_path = argv[++k];
        else if (a == "--chains" && k + 1 < argc) chains = (uint32_t)std::strtoul(argv[++k], nullptr, 10);
        else { std::fprintf(stderr, "usage: %s [--json INPUT.json] [--chains N]\n", argv[0]); return 2; }
    }
    try {
        MainArgs args;
        if (json_path.empty()) {
            std::printf("employee scheduling local search example\n");
            args.start_date = NaiveDate::parse("2022-05-09");
            args.end_date = args.start_date + 30;
            for (int64_t id = 0; id < 7; ++id) args.employees.insert(Employee{id});
        } else {
            std::ifstream f(json_path);
            if (!f) { std::fprintf(stderr, "cannot open %s\n", json_path.c_str()); return 2; }
            std::stringstream ss;
            ss << f.rdbuf();
            args = parse_input(ss.str());
        }
        args.n_chains = chains;
        IlsType iterated_local_search = get_ils(args);
        while (!iterated_local_search.is_finished()) {
            iterated_local_search.execute_round();
            if (iterated_local_search.last_stats().best_key == 0) break;  // later rounds are early-out no-ops
        }
        const Scored result = iterated_local_search.get_best_solution();
        if (!json_path.empty()) {  // ScoredSolutionWrapper
            std::printf("{\"score\":{\"hard_score\":%.1f,\"soft_score\":%.1f},\"days_to_employees\":[", result.score.hard_score,
                        result.score.soft_score);
            bool first = true;
            for (auto& de : result.solution.get_days_to_employees()) {
                std::printf("%s[\"%s\",{\"id\":%lld}]", first ? "" : ",", de.first.format_a_ymd().c_str(), (long long)de.second.id);
                first = false;
            }
            std::printf("]}\n");
            return 0;
        }
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
