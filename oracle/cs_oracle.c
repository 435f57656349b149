/*
 * cs_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).  See cs_oracle.h for the rules.
 *
 * Restates, in plain C, the reference's algorithm for the move-evaluation path:
 *   examples/nqueens/src/lib.rs            (score definition, initial solution)
 *   examples/employee-scheduling/src/lib.rs (score definition)
 *   local-search/src/local_search.rs        (LocalSearch::execute, tabu quirk)
 * Deliberately the SLOW formulation the reference uses (clone + full re-score per
 * candidate): it is the independent check for the GPU's counter/delta formulation.
 */
#include "cs_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3" (SC'11). */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

void orc_philox_stream(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t counter,
                       uint32_t out[4]) {
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint32_t ctr[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), chain, purpose};
    orc_philox4x32_10(ctr, key, out);
}

uint32_t orc_philox_draw(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t t) {
    uint32_t o[4];
    orc_philox_stream(seed, chain, purpose, t >> 2, o);
    return o[t & 3];
}

/* ------------------------------------------------------------------ n-queens */

/* examples/nqueens/src/lib.rs:74-87 */
void orc_nq_col_scores(const int64_t* rows, int64_t n, int64_t* out) {
    for (int64_t c = 0; c < n; ++c) out[c] = 0;
    for (int64_t col1 = 0; col1 < n; ++col1) {
        const int64_t row1 = rows[col1];
        for (int64_t col2 = col1 + 1; col2 < n; ++col2) {
            const int64_t row_diff = rows[col2] - row1;
            const int64_t column_diff = col2 - col1;
            const int64_t ar = row_diff < 0 ? -row_diff : row_diff;
            if (row_diff == 0 || ar == column_diff) { /* :80 */
                out[col1] += 1;
                out[col2] += 1;
            }
        }
    }
}

/* examples/nqueens/src/lib.rs:130-139 : score = sum of per-column conflict counts.
 * Same pair loop as above, accumulated directly (every conflicting pair adds 2). */
int64_t orc_nq_score(const int64_t* rows, int64_t n) {
    int64_t pairs = 0;
    for (int64_t col1 = 0; col1 < n; ++col1) {
        const int64_t row1 = rows[col1];
        int64_t local = 0;
        for (int64_t col2 = col1 + 1; col2 < n; ++col2) {
            const int64_t row_diff = rows[col2] - row1;
            const int64_t column_diff = col2 - col1;
            const int64_t ar = row_diff < 0 ? -row_diff : row_diff;
            local += (row_diff == 0) | (ar == column_diff);
        }
        pairs += local;
    }
    return 2 * pairs;
}

/* examples/nqueens/src/lib.rs:156-160 */
void orc_nq_init_perm(uint64_t seed, uint32_t chain, int64_t n, int64_t* rows) {
    for (int64_t i = 0; i < n; ++i) rows[i] = i;
    uint64_t t = 0;
    for (int64_t k = n - 1; k >= 1; --k, ++t) {
        const uint32_t u = orc_philox_draw(seed, chain, 0u, t);
        const int64_t idx = (int64_t)(((uint64_t)u * (uint64_t)(k + 1)) >> 32);
        const int64_t tmp = rows[k];
        rows[k] = rows[idx];
        rows[idx] = tmp;
    }
}

static int nq_apply(int64_t* cand, int kind, int64_t a, int64_t b) {
    /* returns 0 when the candidate equals the current solution (tabu == {current},
     * local_search.rs:155-199 quirk + :319) */
    if (kind == ORC_NQ_SWAP) {
        if (cand[a] == cand[b]) return 0;
        const int64_t t = cand[a];
        cand[a] = cand[b];
        cand[b] = t;
        return 1;
    }
    if (cand[a] == b) return 0; /* examples/nqueens/src/lib.rs:227-229: rows[col] = value */
    cand[a] = b;
    return 1;
}

void orc_nq_eval_moves(const int64_t* rows, int64_t n, int kind, const int64_t* a,
                       const int64_t* b, int64_t n_moves, int64_t* delta) {
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    const int64_t cur = orc_nq_score(rows, n);
    for (int64_t k = 0; k < n_moves; ++k) {
        memcpy(cand, rows, sizeof(int64_t) * (size_t)n); /* clone, local_search.rs:315-322 */
        if (!nq_apply(cand, kind, a[k], b[k])) {
            delta[k] = INT64_MAX;
            continue;
        }
        delta[k] = orc_nq_score(cand, n) - cur;
    }
    free(cand);
}

int64_t orc_nq_neighbourhood_deltas(const int64_t* rows, int64_t n, int kind, int64_t* delta) {
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    const int64_t cur = orc_nq_score(rows, n);
    int64_t k = 0;
    for (int64_t x = 0; x < n; ++x) {
        const int64_t y0 = (kind == ORC_NQ_SWAP) ? x + 1 : 0;
        for (int64_t y = y0; y < n; ++y, ++k) {
            memcpy(cand, rows, sizeof(int64_t) * (size_t)n);
            if (!nq_apply(cand, kind, x, y)) {
                delta[k] = INT64_MAX;
                continue;
            }
            delta[k] = orc_nq_score(cand, n) - cur;
        }
    }
    free(cand);
    return k;
}

static int lex_less(const int64_t* x, const int64_t* y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] != y[i]) return x[i] < y[i];
    }
    return 0;
}

/* local-search/src/local_search.rs:301-342 */
int64_t orc_nq_local_search(int64_t* rows, int64_t n, int kind, int tie,
                            uint64_t allow_no_improvement_for, uint64_t max_iterations,
                            uint64_t window_size, int64_t* best_score, int64_t* current_out,
                            int64_t* current_score_out, int64_t* trace_a, int64_t* trace_b,
                            int64_t* trace_score, int64_t cap) {
    const size_t bytes = sizeof(int64_t) * (size_t)(n > 0 ? n : 1);
    int64_t* current = (int64_t*)malloc(bytes);
    int64_t* best = (int64_t*)malloc(bytes);
    int64_t* cand = (int64_t*)malloc(bytes);
    int64_t* nb = (int64_t*)malloc(bytes);
    memcpy(current, rows, bytes);
    int64_t current_score = orc_nq_score(current, n); /* :306 */
    memcpy(best, current, bytes);                      /* :307 */
    int64_t bscore = current_score;
    uint64_t no_improvement_for = 0;
    int64_t steps = 0;
    for (uint64_t it = 0; it < max_iterations; ++it) { /* :309 */
        /* :310 seen_solution => tabu == {current} (age test inverted, :182-195) */
        if (current_score == 0) { /* :311-314 returns current_solution */
            memcpy(best, current, bytes);
            bscore = current_score;
            break;
        }
        int have = 0;
        int64_t nb_score = 0, nb_a = -1, nb_b = -1;
        uint64_t taken = 0;
        int stop = 0;
        for (int64_t x = 0; x < n && !stop; ++x) { /* :315-322 */
            const int64_t y0 = (kind == ORC_NQ_SWAP) ? x + 1 : 0;
            for (int64_t y = y0; y < n; ++y) {
                if (window_size && taken >= window_size) { /* .take(window) :321 */
                    stop = 1;
                    break;
                }
                memcpy(cand, current, bytes);
                if (!nq_apply(cand, kind, x, y)) continue; /* tabu filter :319 */
                const int64_t s = orc_nq_score(cand, n);   /* :320 */
                ++taken;
                int better;
                if (!have) better = 1;
                else if (s != nb_score) better = s < nb_score;
                else better = (tie == ORC_TIE_REFERENCE) ? lex_less(cand, nb, n) : 0;
                if (better) { /* sort() then first(), :323-325 */
                    have = 1;
                    nb_score = s;
                    nb_a = x;
                    nb_b = y;
                    memcpy(nb, cand, bytes);
                }
            }
        }
        if (!have) break; /* :336-338 */
        if (nb_score < current_score) { /* :326-328 */
            memcpy(best, nb, bytes);
            bscore = nb_score;
            no_improvement_for = 0;
        } else { /* :329-334 */
            no_improvement_for += 1;
            if (no_improvement_for >= allow_no_improvement_for) break;
        }
        memcpy(current, nb, bytes); /* :335 */
        current_score = nb_score;
        if (steps < cap) {
            if (trace_a) trace_a[steps] = nb_a;
            if (trace_b) trace_b[steps] = nb_b;
            if (trace_score) trace_score[steps] = nb_score;
        }
        ++steps;
    }
    memcpy(rows, best, bytes); /* :341 */
    if (best_score) *best_score = bscore;
    if (current_out) memcpy(current_out, current, bytes);
    if (current_score_out) *current_score_out = current_score;
    free(current);
    free(best);
    free(cand);
    free(nb);
    return steps;
}

int64_t orc_nq_baseline_sample(const int64_t* rows, int64_t n, const int64_t* a,
                               const int64_t* b, int64_t n_moves, int threads,
                               int64_t* checksum) {
    int64_t sum = 0;
    const int64_t cur = orc_nq_score(rows, n);
#pragma omp parallel num_threads(threads) reduction(+ : sum)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t k = 0; k < n_moves; ++k) {
            memcpy(cand, rows, sizeof(int64_t) * (size_t)n);
            if (nq_apply(cand, ORC_NQ_SWAP, a[k], b[k])) sum += orc_nq_score(cand, n) - cur;
        }
        free(cand);
    }
    if (checksum) *checksum = sum;
    return n_moves;
}

/* ------------------------------------------------------------------ employee scheduling */

/* Proleptic Gregorian civil-date arithmetic (what chrono's NaiveDate does). */
int64_t orc_days_from_civil(int64_t y, int m, int d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const int64_t yoe = y - era * 400;
    const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}

int orc_weekday_from_days(int64_t z) {
    /* 1970-01-01 was a Thursday (=3 with Monday=0). */
    int64_t w = (z + 3) % 7;
    if (w < 0) w += 7;
    return (int)w;
}

static inline int es_is_weekend(int start_weekday, int64_t i) {
    const int w = (int)((start_weekday + i) % 7);
    return w == 5 || w == 6; /* lib.rs:220-222 */
}

/* count distinct ids in window with multiplicity > limit (itertools counts(), lib.rs:319-326) */
static int64_t es_window_violations(const int64_t* w, int64_t len, int64_t limit) {
    int64_t v = 0;
    for (int64_t p = 0; p < len; ++p) {
        int first = 1;
        for (int64_t q = 0; q < p; ++q)
            if (w[q] == w[p]) { first = 0; break; }
        if (!first) continue;
        int64_t c = 0;
        for (int64_t q = p; q < len; ++q) c += (w[q] == w[p]);
        if (c > limit) ++v;
    }
    return v;
}

int orc_es_score_terms(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                       const int64_t* hol_day, int64_t n_hol, int64_t out[8]) {
    for (int k = 0; k < 8; ++k) out[k] = 0;
    /* H1 holidays, lib.rs:273-280.  The reference keeps them in a HashSet<Holiday> (lib.rs:255-259),
     * so a duplicated (employee, day) entry counts ONCE: a day has one employee, hence at most one
     * distinct matching pair per day -- a per-day flag de-duplicates exactly. */
    {
        unsigned char seen_small[512];
        unsigned char* seen = D <= 512 ? seen_small : (unsigned char*)malloc((size_t)D);
        memset(seen, 0, (size_t)(D > 0 ? D : 0));
        int bad = 0;
        for (int64_t k = 0; k < n_hol; ++k) {
            if (hol_day[k] < 0 || hol_day[k] >= D) { /* unwrap() on None, :275 */
                bad = 1;
                break;
            }
            if (a[hol_day[k]] == hol_emp[k] && !seen[hol_day[k]]) {
                seen[hol_day[k]] = 1;
                out[0] += 1;
            }
        }
        if (seen != seen_small) free(seen);
        if (bad) return -1;
    }
    /* H2 consecutive days, lib.rs:286-292 */
    for (int64_t i = 0; i + 2 <= D; ++i)
        if (a[i] == a[i + 1]) out[1] += 1;
    /* H3 consecutive weekends, lib.rs:295-315 */
    for (int64_t i = 0; i + 9 <= D; ++i) {
        if (!(es_is_weekend(start_weekday, i) && es_is_weekend(start_weekday, i + 1))) continue;
        if (a[i] == a[i + 7]) out[2] += 1;
        if (a[i] == a[i + 8]) out[2] += 1;
        if (a[i + 1] == a[i + 7]) out[2] += 1;
        if (a[i + 1] == a[i + 8]) out[2] += 1;
    }
    /* H4 no more than 3 per 14 days, lib.rs:318-327 */
    for (int64_t i = 0; i + 14 <= D; ++i) out[3] += es_window_violations(a + i, 14, 3);
    /* S1 no more than 2 per 7 days, lib.rs:330-339 */
    for (int64_t i = 0; i + 7 <= D; ++i) out[4] += es_window_violations(a + i, 7, 2);
    /* S2 weekday affinity, lib.rs:194-218: per weekday Mon..Fri with >= 2 distinct
     * employees, add the minimum per-employee count. */
    for (int wd = 0; wd < 5; ++wd) {
        int64_t distinct = 0, minc = INT64_MAX;
        for (int64_t i = 0; i < D; ++i) {
            if ((start_weekday + i) % 7 != wd) continue;
            int first = 1;
            for (int64_t q = 0; q < i; ++q)
                if ((start_weekday + q) % 7 == wd && a[q] == a[i]) { first = 0; break; }
            if (!first) continue;
            int64_t c = 0;
            for (int64_t q = i; q < D; ++q)
                if ((start_weekday + q) % 7 == wd && a[q] == a[i]) ++c;
            ++distinct;
            if (c < minc) minc = c;
        }
        if (distinct >= 2) out[5] += minc; /* len()<=1 skipped :206; MinMax => += min :212-214 */
    }
    /* S3 / S4 spreads over employees present at least once, lib.rs:345-365 */
    int64_t present = 0, mind = INT64_MAX, maxd = INT64_MIN, minw = INT64_MAX, maxw = INT64_MIN;
    for (int64_t i = 0; i < D; ++i) {
        int first = 1;
        for (int64_t q = 0; q < i; ++q)
            if (a[q] == a[i]) { first = 0; break; }
        if (!first) continue;
        int64_t days = 0, wk = 0;
        for (int64_t q = i; q < D; ++q)
            if (a[q] == a[i]) {
                ++days;
                wk += es_is_weekend(start_weekday, q);
            }
        ++present;
        if (days < mind) mind = days;
        if (days > maxd) maxd = days;
        if (wk < minw) minw = wk;
        if (wk > maxw) maxw = wk;
    }
    if (present >= 2) { /* MinMaxResult::MinMax only for >= 2 elements :349,:363 */
        out[6] = maxd - mind;
        out[7] = maxw - minw;
    }
    return 0;
}

int orc_es_score(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                 const int64_t* hol_day, int64_t n_hol, int64_t* hard, int64_t* soft) {
    int64_t t[8];
    const int rc = orc_es_score_terms(a, D, start_weekday, hol_emp, hol_day, n_hol, t);
    if (rc) return rc;
    *hard = t[0] + t[1] + t[2] + t[3];
    *soft = t[4] + t[5] + t[6] + t[7];
    return 0;
}

/* examples/employee-scheduling/src/lib.rs:404-419: one uniformly random employee per slot,
 * n_slots = D + 1 (the loop pushes a phantom slot past end_date before it breaks).
 * `choose` restated as employees[mulhi(u32, E)] over Philox purpose 0, draw s for slot s. */
void orc_es_init(uint64_t seed, uint32_t chain, int64_t n_slots, const int64_t* employees,
                 int64_t E, int64_t* out) {
    for (int64_t s = 0; s < n_slots; ++s) {
        const uint32_t u = orc_philox_draw(seed, chain, 0u, (uint64_t)s);
        out[s] = employees[(int64_t)(((uint64_t)u * (uint64_t)E) >> 32)];
    }
}

static int es_apply(int64_t* cand, const int64_t* employees, int kind, int64_t x, int64_t y) {
    if (kind == ORC_ES_CHANGE) { /* lib.rs:466-470 */
        if (cand[x] == employees[y]) return 0;
        cand[x] = employees[y];
        return 1;
    }
    if (cand[x] == cand[y]) return 0; /* lib.rs:471-478 */
    const int64_t t = cand[x];
    cand[x] = cand[y];
    cand[y] = t;
    return 1;
}

int orc_es_eval_moves(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                      const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                      int64_t E, int kind, const int64_t* x, const int64_t* y, int64_t n_moves,
                      int64_t* dhard, int64_t* dsoft) {
    (void)E;
    int64_t h0, s0;
    if (orc_es_score(a, D, start_weekday, hol_emp, hol_day, n_hol, &h0, &s0)) return -1;
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(D > 0 ? D : 1));
    for (int64_t k = 0; k < n_moves; ++k) {
        memcpy(cand, a, sizeof(int64_t) * (size_t)D);
        if (!es_apply(cand, employees, kind, x[k], y[k])) {
            dhard[k] = INT64_MAX;
            dsoft[k] = INT64_MAX;
            continue;
        }
        int64_t h, s;
        orc_es_score(cand, D, start_weekday, hol_emp, hol_day, n_hol, &h, &s);
        dhard[k] = h - h0;
        dsoft[k] = s - s0;
    }
    free(cand);
    return 0;
}

int64_t orc_es_local_search(int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                            const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                            int64_t E, uint64_t allow_no_improvement_for,
                            uint64_t max_iterations, int64_t* best_hard, int64_t* best_soft,
                            int64_t* current_out, int64_t* trace_kind, int64_t* trace_x,
                            int64_t* trace_y, int64_t* trace_hard, int64_t* trace_soft,
                            int64_t cap) {
    const size_t bytes = sizeof(int64_t) * (size_t)(D > 0 ? D : 1);
    int64_t* current = (int64_t*)malloc(bytes);
    int64_t* best = (int64_t*)malloc(bytes);
    int64_t* cand = (int64_t*)malloc(bytes);
    int64_t* nb = (int64_t*)malloc(bytes);
    memcpy(current, a, bytes);
    int64_t ch, cs;
    orc_es_score(current, D, start_weekday, hol_emp, hol_day, n_hol, &ch, &cs);
    memcpy(best, current, bytes);
    int64_t bh = ch, bs = cs;
    uint64_t no_improvement_for = 0;
    int64_t steps = 0;
    for (uint64_t it = 0; it < max_iterations; ++it) {
        if (ch == 0 && cs == 0) { /* is_best lib.rs:245-249; local_search.rs:311-314 */
            memcpy(best, current, bytes);
            bh = ch;
            bs = cs;
            break;
        }
        int have = 0;
        int64_t nh = 0, ns = 0, nk = 0, nx = 0, ny = 0;
        for (int kind = 0; kind < 2; ++kind) {
            for (int64_t x = 0; x < D; ++x) {
                const int64_t y0 = (kind == ORC_ES_SWAP) ? x + 1 : 0;
                const int64_t y1 = (kind == ORC_ES_SWAP) ? D : E;
                for (int64_t y = y0; y < y1; ++y) {
                    memcpy(cand, current, bytes);
                    if (!es_apply(cand, employees, kind, x, y)) continue;
                    int64_t h, s;
                    orc_es_score(cand, D, start_weekday, hol_emp, hol_day, n_hol, &h, &s);
                    if (!have || h < nh || (h == nh && s < ns)) {
                        have = 1;
                        nh = h; ns = s; nk = kind; nx = x; ny = y;
                        memcpy(nb, cand, bytes);
                    }
                }
            }
        }
        if (!have) break;
        if (nh < ch || (nh == ch && ns < cs)) {
            memcpy(best, nb, bytes);
            bh = nh;
            bs = ns;
            no_improvement_for = 0;
        } else {
            no_improvement_for += 1;
            if (no_improvement_for >= allow_no_improvement_for) break;
        }
        memcpy(current, nb, bytes);
        ch = nh;
        cs = ns;
        if (steps < cap) {
            if (trace_kind) trace_kind[steps] = nk;
            if (trace_x) trace_x[steps] = nx;
            if (trace_y) trace_y[steps] = ny;
            if (trace_hard) trace_hard[steps] = nh;
            if (trace_soft) trace_soft[steps] = ns;
        }
        ++steps;
    }
    memcpy(a, best, bytes);
    if (best_hard) *best_hard = bh;
    if (best_soft) *best_soft = bs;
    if (current_out) memcpy(current_out, current, bytes);
    free(current);
    free(best);
    free(cand);
    free(nb);
    return steps;
}

/* LocalSearch::execute (local-search/src/local_search.rs:301-342) with the reference's OWN
 * scheduling proposer, ScheduleRandomMoveProposer::iter_local_moves
 * (examples/employee-scheduling/src/lib.rs:440-491): an endless stream of random ChangeDay
 * (weight 1) / SwapDays (weight 4) candidates (:435, :459-478) drawn from a CLONE of the
 * LocalSearch rng (:488) -- the solver rng never advances, so EVERY step replays the same draws
 * from t = 0 -- each a full clone of the rota, filtered by the tabu set (== {current},
 * local_search.rs:155-199,319), full re-scored (:320), truncated to window_size (:321) and
 * sorted by the derived Ord (score, then date_to_employee by Employee.id; :29-37,323).
 * Draw restatement (rand 0.8.5 is un-vendored: parity unpinned at the RNG boundary): candidate k
 * uses draws 3k..3k+2 of Philox stream (seed, chain, purpose 2):
 *   choose_weighted     -> mulhi(u,5) < 1 ? ChangeDay : SwapDays
 *   ChangeDay           -> day = mulhi(u,D) (scored days only, :466), employee = employees[mulhi(u,E)]
 *   SwapDays            -> choose_multiple(2): d1 = mulhi(u,D), d2 = mulhi(u,D-1), d2 += (d2 >= d1)
 * max_draws bounds the endless iterator: the reference spins forever when every candidate is tabu
 * (one employee); D = 1 SwapDays candidates are skipped (the reference indexes xs[1] and panics).
 * a: in = start, out = best_solution.  trace: kind, x, y (change: day, employee INDEX; swap:
 * min day, max day) and the score after each accepted step.  Returns the accepted steps;
 * *scored_out (optional) = candidates scored in total. */
int64_t orc_es_local_search_ref(int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                                const int64_t* hol_day, int64_t n_hol, const int64_t* employees,
                                int64_t E, uint64_t seed, uint32_t chain,
                                uint64_t allow_no_improvement_for, uint64_t max_iterations,
                                uint64_t window_size, uint64_t max_draws, int64_t* best_hard,
                                int64_t* best_soft, int64_t* current_out, int64_t* trace_kind,
                                int64_t* trace_x, int64_t* trace_y, int64_t* trace_hard,
                                int64_t* trace_soft, int64_t cap, int64_t* scored_out) {
    const size_t bytes = sizeof(int64_t) * (size_t)(D > 0 ? D : 1);
    int64_t* current = (int64_t*)malloc(bytes);
    int64_t* best = (int64_t*)malloc(bytes);
    int64_t* cand = (int64_t*)malloc(bytes);
    int64_t* nb = (int64_t*)malloc(bytes);
    memcpy(current, a, bytes);
    int64_t ch, cs;
    orc_es_score(current, D, start_weekday, hol_emp, hol_day, n_hol, &ch, &cs);
    memcpy(best, current, bytes);
    int64_t bh = ch, bs = cs, steps = 0, scored = 0;
    uint64_t no_improvement_for = 0;
    for (uint64_t it = 0; it < max_iterations; ++it) {
        if (ch == 0 && cs == 0) { /* local_search.rs:311-314 */
            memcpy(best, current, bytes);
            bh = ch;
            bs = cs;
            break;
        }
        int have = 0;
        int64_t nh = 0, ns = 0, nk = 0, nx = 0, ny = 0;
        uint64_t taken = 0;
        for (uint64_t k = 0; k < max_draws && taken < window_size; ++k) { /* rng.clone(): t restarts at 0 */
            const uint32_t u0 = orc_philox_draw(seed, chain, 2u, 3 * k);
            const uint32_t u1 = orc_philox_draw(seed, chain, 2u, 3 * k + 1);
            const uint32_t u2 = orc_philox_draw(seed, chain, 2u, 3 * k + 2);
            int64_t kind, x, y;
            memcpy(cand, current, bytes); /* self.solution.clone(), lib.rs:464 */
            if ((((uint64_t)u0 * 5u) >> 32) < 1) {
                kind = ORC_ES_CHANGE;
                x = (int64_t)(((uint64_t)u1 * (uint64_t)D) >> 32);
                y = (int64_t)(((uint64_t)u2 * (uint64_t)E) >> 32);
                cand[x] = employees[y];
            } else {
                kind = ORC_ES_SWAP;
                if (D < 2) continue;
                int64_t d1 = (int64_t)(((uint64_t)u1 * (uint64_t)D) >> 32);
                int64_t d2 = (int64_t)(((uint64_t)u2 * (uint64_t)(D - 1)) >> 32);
                d2 += (d2 >= d1);
                x = d1 < d2 ? d1 : d2;
                y = d1 < d2 ? d2 : d1;
                const int64_t t = cand[x];
                cand[x] = cand[y];
                cand[y] = t;
            }
            if (memcmp(cand, current, bytes) == 0) continue; /* tabu == {current}, local_search.rs:319 */
            ++taken;
            ++scored;
            int64_t h, s_;
            orc_es_score(cand, D, start_weekday, hol_emp, hol_day, n_hol, &h, &s_);
            int better = !have || h < nh || (h == nh && s_ < ns);
            if (have && h == nh && s_ == ns) { /* derived Ord: lexicographic date_to_employee */
                for (int64_t q = 0; q < D; ++q)
                    if (cand[q] != nb[q]) {
                        better = cand[q] < nb[q];
                        break;
                    }
            }
            if (better) {
                have = 1;
                nh = h; ns = s_; nk = kind; nx = x; ny = y;
                memcpy(nb, cand, bytes);
            }
        }
        if (!have) break; /* empty window, local_search.rs:336-338 */
        if (nh < ch || (nh == ch && ns < cs)) {
            memcpy(best, nb, bytes);
            bh = nh;
            bs = ns;
            no_improvement_for = 0;
        } else {
            no_improvement_for += 1;
            if (no_improvement_for >= allow_no_improvement_for) break;
        }
        memcpy(current, nb, bytes);
        ch = nh;
        cs = ns;
        if (steps < cap) {
            if (trace_kind) trace_kind[steps] = nk;
            if (trace_x) trace_x[steps] = nx;
            if (trace_y) trace_y[steps] = ny;
            if (trace_hard) trace_hard[steps] = nh;
            if (trace_soft) trace_soft[steps] = ns;
        }
        ++steps;
    }
    memcpy(a, best, bytes);
    if (best_hard) *best_hard = bh;
    if (best_soft) *best_soft = bs;
    if (current_out) memcpy(current_out, current, bytes);
    if (scored_out) *scored_out = scored;
    free(current);
    free(best);
    free(cand);
    free(nb);
    return steps;
}

int64_t orc_es_baseline_sample(const int64_t* a, int64_t D, int start_weekday,
                               const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                               const int64_t* employees, int64_t E, int kind, const int64_t* x,
                               const int64_t* y, int64_t n_moves, int threads,
                               int64_t* checksum) {
    (void)E;
    int64_t sum = 0;
    int64_t h0, s0;
    if (orc_es_score(a, D, start_weekday, hol_emp, hol_day, n_hol, &h0, &s0)) return -1;
#pragma omp parallel num_threads(threads) reduction(+ : sum)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(D > 0 ? D : 1));
#pragma omp for schedule(static)
        for (int64_t k = 0; k < n_moves; ++k) {
            memcpy(cand, a, sizeof(int64_t) * (size_t)D);
            if (es_apply(cand, employees, kind, x[k], y[k])) {
                int64_t h, s;
                orc_es_score(cand, D, start_weekday, hol_emp, hol_day, n_hol, &h, &s);
                sum += (h - h0) * 1000 + (s - s0);
            }
        }
        free(cand);
    }
    if (checksum) *checksum = sum;
    return n_moves;
}


/* ------------------------------------------------------------------ reference-mode proposer */
/* NQueensMoveProposer::iter_local_moves, examples/nqueens/src/lib.rs:177-255, followed by
 * LocalSearch::execute's window + sort (local_search.rs:315-323), literally: candidates are
 * full clones, each fully re-scored, ordered by the derived Ord (score, then the solution
 * vector lexicographically).
 * Random choices over the chain's Philox stream purpose 2 (the LocalSearch-owned rng):
 *   choose_multiple_weighted(amount, w = score + 1e-4)  ->  `amount` sequential draws without
 *       replacement, x = mulhi(u, sum of remaining integer scores), first column (ascending)
 *       whose cumulative score exceeds x;
 *   gen_range(1..=len) -> 1 + mulhi(u, len);   choose_multiple(num) -> partial Fisher-Yates
 *       k = 0..num-1: swap(k, k + mulhi(u, len - k)).
 * (rand 0.8.5's own sampling algorithms are un-vendored: parity unpinned at the RNG boundary.) */
static uint32_t ls_rng_next(uint64_t seed, uint32_t chain, uint64_t* t) {
    return orc_philox_draw(seed, chain, 2u, (*t)++);
}
static int64_t ls_rng_below(uint64_t seed, uint32_t chain, uint64_t* t, int64_t m) {
    return (int64_t)(((uint64_t)ls_rng_next(seed, chain, t) * (uint64_t)m) >> 32);
}

/* returns the number of chosen columns written to `chosen` (0 = no conflicts, lib.rs:193-194) */
static int64_t nq_ref_propose(const int64_t* rows, int64_t n, uint64_t seed, uint32_t chain,
                              uint64_t* t, int64_t* chosen, int64_t* scratch /* 2n */) {
    int64_t* col = scratch;
    int64_t* sc = scratch + n;
    orc_nq_col_scores(rows, n, sc); /* :182 */
    int64_t len = 0;
    for (int64_t c = 0; c < n; ++c)
        if (sc[c] != 0) { /* :183-187, already in ascending column order (:187 sort) */
            col[len] = c;
            sc[len] = sc[c];
            ++len;
        }
    if (len == 0) return 0;
    int64_t amount = n / 20; /* :196 */
    if (amount < 1) amount = 1;
    if (amount > len) amount = len;
    int64_t npicked = 0;
    for (int64_t k = 0; k < amount; ++k) { /* :197-201 */
        int64_t total = 0;
        for (int64_t q = 0; q < len; ++q) total += sc[q];
        const int64_t x = ls_rng_below(seed, chain, t, total);
        int64_t acc = 0, idx = 0;
        for (; idx < len; ++idx) {
            acc += sc[idx];
            if (acc > x) break;
        }
        chosen[npicked++] = col[idx];
        for (int64_t q = idx; q + 1 < len; ++q) {
            col[q] = col[q + 1];
            sc[q] = sc[q + 1];
        }
        --len;
    }
    const int64_t num_cols = 1 + ls_rng_below(seed, chain, t, npicked); /* :202 */
    for (int64_t k = 0; k < num_cols; ++k) {                            /* :203 */
        const int64_t j = k + ls_rng_below(seed, chain, t, npicked - k);
        const int64_t tmp = chosen[k];
        chosen[k] = chosen[j];
        chosen[j] = tmp;
    }
    return num_cols;
}

int64_t orc_nq_local_search_ref(int64_t* rows, int64_t n, uint64_t seed, uint32_t chain,
                                uint64_t* rng_t, uint64_t allow_no_improvement_for,
                                uint64_t max_iterations, uint64_t window_size, int64_t* best_score,
                                int64_t* current_out, int64_t* trace_a, int64_t* trace_b,
                                int64_t* trace_score, int64_t cap) {
    const size_t bytes = sizeof(int64_t) * (size_t)(n > 0 ? n : 1);
    int64_t* current = (int64_t*)malloc(bytes);
    int64_t* best = (int64_t*)malloc(bytes);
    int64_t* cand = (int64_t*)malloc(bytes);
    int64_t* nb = (int64_t*)malloc(bytes);
    int64_t* chosen = (int64_t*)malloc(bytes);
    int64_t* scratch = (int64_t*)malloc(bytes * 2);
    memcpy(current, rows, bytes);
    int64_t current_score = orc_nq_score(current, n);
    memcpy(best, current, bytes);
    int64_t bscore = current_score;
    uint64_t no_improvement_for = 0;
    int64_t steps = 0;
    for (uint64_t it = 0; it < max_iterations; ++it) {
        if (current_score == 0) {
            memcpy(best, current, bytes);
            bscore = 0;
            break;
        }
        const int64_t ncols = nq_ref_propose(current, n, seed, chain, rng_t, chosen, scratch);
        int have = 0;
        int64_t nb_score = 0, nb_a = -1, nb_b = -1;
        uint64_t taken = 0;
        for (int64_t k = 0; k < ncols && taken < window_size; ++k) { /* lib.rs:217-234 */
            for (int64_t v = 0; v < n && taken < window_size; ++v) {
                memcpy(cand, current, bytes);
                cand[chosen[k]] = v;
                if (v == current[chosen[k]]) continue; /* tabu == {current} */
                const int64_t sc = orc_nq_score(cand, n);
                ++taken;
                int better;
                if (!have) better = 1;
                else if (sc != nb_score) better = sc < nb_score;
                else better = lex_less(cand, nb, n); /* derived Ord: (score, solution) */
                if (better) {
                    have = 1;
                    nb_score = sc;
                    nb_a = chosen[k];
                    nb_b = v;
                    memcpy(nb, cand, bytes);
                }
            }
        }
        if (!have) break;
        if (nb_score < current_score) {
            memcpy(best, nb, bytes);
            bscore = nb_score;
            no_improvement_for = 0;
        } else {
            no_improvement_for += 1;
            if (no_improvement_for >= allow_no_improvement_for) break;
        }
        memcpy(current, nb, bytes);
        current_score = nb_score;
        if (steps < cap) {
            if (trace_a) trace_a[steps] = nb_a;
            if (trace_b) trace_b[steps] = nb_b;
            if (trace_score) trace_score[steps] = nb_score;
        }
        ++steps;
    }
    memcpy(rows, best, bytes);
    if (best_score) *best_score = bscore;
    if (current_out) memcpy(current_out, current, bytes);
    free(current);
    free(best);
    free(cand);
    free(nb);
    free(chosen);
    free(scratch);
    return steps;
}

/* ------------------------------------------------------------------ iterated local search */
/* IteratedLocalSearch::execute_round, local-search/src/iterated_local_search.rs:173-202, with
 *   - History::local_search_chose_solution (bounded best-set, BTreeSet ordered by
 *     (score, solution)), local_search.rs:205-218;
 *   - AcceptanceCriterion::choose weights {existing 1, new 5, random best 1}, :51-71;
 *   - the plug-in perturbations (nqueens lib.rs:291-319, employee-scheduling lib.rs:588-612);
 *   - random restart every 50th round, iterated_local_search.rs:185-191.
 * Random choices are restated over ONE Philox stream per chain (purpose 1), draw t after draw
 * t-1:  weighted strategy pick = mulhi(u,110); gen_range(1..=m) = 1 + mulhi(u,m);
 * shuffle = Fisher-Yates k=len-1..1 swap(k, mulhi(u,k+1)); choose = mulhi(u,len);
 * acceptance = mulhi(u,7).  (rand 0.8.5's own algorithms are un-vendored: parity unpinned.)
 * A solution is a vector of `len` small integers with an int64 score key (n-queens: the
 * score; scheduling: hard << 32 | soft) -- lexicographic (key, vector) is the derived Ord. */
typedef struct {
    uint64_t seed;
    uint32_t chain;
    uint64_t t;
} orc_rng;

static uint32_t rng_next(orc_rng* r) { return orc_philox_draw(r->seed, r->chain, 1u, r->t++); }
static int64_t rng_below(orc_rng* r, int64_t m) {
    return (int64_t)(((uint64_t)rng_next(r) * (uint64_t)m) >> 32);
}

typedef int64_t (*orc_ls_fn)(void* ctx, int64_t* sol); /* in: start, out: best; returns key */
typedef void (*orc_restart_fn)(void* ctx, orc_rng* r, int64_t* sol);

static int sol_cmp(int64_t ka, const int64_t* a, int64_t kb, const int64_t* b, int64_t len) {
    if (ka != kb) return ka < kb ? -1 : 1;
    for (int64_t i = 0; i < len; ++i)
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
}

static int64_t ils_core(void* ctx, orc_ls_fn ls, orc_restart_fn restart, uint64_t seed,
                        uint32_t chain, int64_t len, int64_t value_range, int do_nothing_first,
                        int k_before_shuffle, int best_cap, uint64_t rounds, int64_t* current,
                        int64_t* best_out, int64_t* best_key_out, int64_t* round_new_key,
                        int64_t* round_choice) {
    orc_rng rng = {seed, chain, 0};
    const size_t bytes = sizeof(int64_t) * (size_t)len;
    int64_t* set = (int64_t*)malloc(bytes * (size_t)best_cap); /* sorted ascending */
    int64_t* set_key = (int64_t*)malloc(sizeof(int64_t) * (size_t)best_cap);
    int64_t* work = (int64_t*)malloc(bytes);
    int64_t* idx = (int64_t*)malloc(bytes);
    int size = 0;
    uint64_t done_rounds = 0;
    for (uint64_t it = 1; it <= rounds; ++it) {
        if (size > 0 && set_key[0] == 0) break; /* :175-184 best is_best -> nothing more to do */
        done_rounds = it;
        if (it % 50 == 0) restart(ctx, &rng, current); /* :185-191 */
        memcpy(work, current, bytes);
        /* perturbation */
        const int64_t pick = rng_below(&rng, 110);
        const int change = do_nothing_first ? (pick >= 10) : (pick < 100);
        if (change) {
            int in_best = 0; /* history.is_best_solution(current): BTreeSet::contains */
            for (int e = 0; e < size && !in_best; ++e)
                in_best = (memcmp(set + (size_t)e * len, current, bytes) == 0);
            int64_t kmax = in_best ? len / 20 : len / 2;
            if (kmax < 1) kmax = 1;
            if (kmax > len) kmax = len;
            int64_t k = 0;
            if (k_before_shuffle) k = 1 + rng_below(&rng, kmax);
            for (int64_t i = 0; i < len; ++i) idx[i] = i;
            for (int64_t q = len - 1; q >= 1; --q) {
                const int64_t j = rng_below(&rng, q + 1);
                const int64_t tmp = idx[q];
                idx[q] = idx[j];
                idx[j] = tmp;
            }
            if (!k_before_shuffle) k = 1 + rng_below(&rng, kmax);
            for (int64_t q = 0; q < k; ++q) work[idx[q]] = rng_below(&rng, value_range);
        }
        /* local search: work := best found, returns its key (:195-197) */
        const int64_t nkey = ls(ctx, work);
        if (round_new_key) round_new_key[it - 1] = nkey;
        /* history.local_search_chose_solution(new), local_search.rs:205-218 */
        int do_insert = 0;
        if (size < best_cap) {
            do_insert = 1;
        } else if (nkey <= set_key[size - 1]) {
            --size; /* remove the worst */
            do_insert = 1;
        }
        if (do_insert) {
            int pos = 0, dup = 0;
            for (; pos < size; ++pos) {
                const int c = sol_cmp(set_key[pos], set + (size_t)pos * len, nkey, work, len);
                if (c == 0) dup = 1;
                if (c >= 0) break;
            }
            if (!dup) {
                memmove(set + (size_t)(pos + 1) * len, set + (size_t)pos * len, bytes * (size_t)(size - pos));
                memmove(set_key + pos + 1, set_key + pos, sizeof(int64_t) * (size_t)(size - pos));
                memcpy(set + (size_t)pos * len, work, bytes);
                set_key[pos] = nkey;
                ++size;
            }
        }
        /* acceptance_criterion.choose, iterated_local_search.rs:51-71 */
        const int64_t rb = rng_below(&rng, size);
        const int64_t w = rng_below(&rng, 7);
        if (round_choice) round_choice[it - 1] = w == 0 ? 0 : (w <= 5 ? 1 : 2);
        if (w >= 1 && w <= 5) memcpy(current, work, bytes);
        else if (w == 6) memcpy(current, set + (size_t)rb * len, bytes);
    }
    if (size > 0) {
        memcpy(best_out, set, bytes);
        *best_key_out = set_key[0];
    } else {
        memcpy(best_out, current, bytes);
        *best_key_out = -1;
    }
    free(set);
    free(set_key);
    free(work);
    free(idx);
    return (int64_t)done_rounds;
}

typedef struct {
    int64_t n;
    int kind;
    uint64_t allow, iters;
    uint64_t ref_window; /* != 0: the reference's proposer + window + tie-break */
    uint64_t seed;
    uint32_t chain;
    uint64_t ls_t; /* the LocalSearch-owned rng advances across execute() calls */
} nq_ils_ctx;

static int64_t nq_ils_ls(void* c, int64_t* sol) {
    nq_ils_ctx* x = (nq_ils_ctx*)c;
    int64_t best = 0;
    if (x->ref_window)
        orc_nq_local_search_ref(sol, x->n, x->seed, x->chain, &x->ls_t, x->allow, x->iters,
                                x->ref_window, &best, NULL, NULL, NULL, NULL, 0);
    else
        orc_nq_local_search(sol, x->n, x->kind, ORC_TIE_MOVE_ORDER, x->allow, x->iters, 0, &best,
                            NULL, NULL, NULL, NULL, NULL, 0);
    return best;
}

static void nq_ils_restart(void* c, orc_rng* r, int64_t* sol) {
    nq_ils_ctx* x = (nq_ils_ctx*)c;
    for (int64_t i = 0; i < x->n; ++i) sol[i] = i;
    for (int64_t q = x->n - 1; q >= 1; --q) {
        const int64_t j = rng_below(r, q + 1);
        const int64_t t = sol[q];
        sol[q] = sol[j];
        sol[j] = t;
    }
}

/* n-queens ILS chain: initial solution = orc_nq_init_perm(seed, chain) (ILS::new, :141-142).
 * Returns the rounds run; best_rows/best_score = history.get_best() (:165-167). */
int64_t orc_nq_ils(uint64_t seed, uint32_t chain, int64_t n, int kind, uint64_t ls_max_iterations,
                   uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                   int64_t* best_rows, int64_t* best_score, int64_t* current_out,
                   int64_t* round_new_score, int64_t* round_choice, uint64_t ref_window) {
    nq_ils_ctx ctx = {n, kind, allow_no_improvement_for, ls_max_iterations, ref_window, seed, chain, 0};
    int64_t* current = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    orc_nq_init_perm(seed, chain, n, current);
    const int64_t r = ils_core(&ctx, nq_ils_ls, nq_ils_restart, seed, chain, n, n, 0, 0, best_cap,
                               rounds, current, best_rows, best_score, round_new_score, round_choice);
    if (current_out) memcpy(current_out, current, sizeof(int64_t) * (size_t)n);
    free(current);
    return r;
}

typedef struct {
    int64_t D, E;
    int wd;
    const int64_t *he, *hd, *emp;
    int64_t nh;
    uint64_t allow, iters;
    uint64_t ref_window, ref_max_draws, seed; /* ref_window > 0: the reference's own proposer */
    uint32_t chain;
} es_ils_ctx;

/* the ILS vectors of the scheduling problem hold employee INDICES (0..E-1) over D+1 slots */
static int64_t es_ils_ls(void* c, int64_t* sol) {
    es_ils_ctx* x = (es_ils_ctx*)c;
    int64_t* a = (int64_t*)malloc(sizeof(int64_t) * (size_t)x->D);
    for (int64_t i = 0; i < x->D; ++i) a[i] = x->emp[sol[i]];
    int64_t bh = 0, bs = 0;
    if (x->ref_window)
        orc_es_local_search_ref(a, x->D, x->wd, x->he, x->hd, x->nh, x->emp, x->E, x->seed, x->chain, x->allow,
                                x->iters, x->ref_window, x->ref_max_draws, &bh, &bs, NULL, NULL, NULL, NULL,
                                NULL, NULL, 0, NULL);
    else
        orc_es_local_search(a, x->D, x->wd, x->he, x->hd, x->nh, x->emp, x->E, x->allow, x->iters, &bh,
                            &bs, NULL, NULL, NULL, NULL, NULL, NULL, 0);
    for (int64_t i = 0; i < x->D; ++i) {
        int64_t e = 0;
        while (x->emp[e] != a[i]) ++e;
        sol[i] = e; /* phantom slot sol[D] is untouched by the local search */
    }
    free(a);
    return (bh << 32) | bs;
}

static void es_ils_restart(void* c, orc_rng* r, int64_t* sol) {
    es_ils_ctx* x = (es_ils_ctx*)c;
    for (int64_t s = 0; s <= x->D; ++s) sol[s] = rng_below(r, x->E);
}

int64_t orc_es_ils(uint64_t seed, uint32_t chain, int64_t D, int start_weekday,
                   const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                   const int64_t* employees, int64_t E, uint64_t ls_max_iterations,
                   uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                   int64_t* best_idx /*[D+1] employee indices*/, int64_t* best_hard,
                   int64_t* best_soft, int64_t* round_new_key, int64_t* round_choice) {
    return orc_es_ils_ref(seed, chain, D, start_weekday, hol_emp, hol_day, n_hol, employees, E, ls_max_iterations,
                          allow_no_improvement_for, rounds, best_cap, best_idx, best_hard, best_soft, round_new_key,
                          round_choice, 0, 0);
}

/* orc_es_ils with LocalSearch::execute running the reference's own proposer when ref_window > 0 */
int64_t orc_es_ils_ref(uint64_t seed, uint32_t chain, int64_t D, int start_weekday,
                       const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                       const int64_t* employees, int64_t E, uint64_t ls_max_iterations,
                       uint64_t allow_no_improvement_for, uint64_t rounds, int best_cap,
                       int64_t* best_idx, int64_t* best_hard, int64_t* best_soft,
                       int64_t* round_new_key, int64_t* round_choice, uint64_t ref_window,
                       uint64_t ref_max_draws) {
    es_ils_ctx ctx = {D, E, start_weekday, hol_emp, hol_day, employees, n_hol,
                      allow_no_improvement_for, ls_max_iterations, ref_window, ref_max_draws, seed, chain};
    int64_t* current = (int64_t*)malloc(sizeof(int64_t) * (size_t)(D + 1));
    for (int64_t s = 0; s <= D; ++s) /* same draws as orc_es_init, as indices */
        current[s] = (int64_t)(((uint64_t)orc_philox_draw(seed, chain, 0u, (uint64_t)s) * (uint64_t)E) >> 32);
    int64_t key = 0;
    const int64_t r = ils_core(&ctx, es_ils_ls, es_ils_restart, seed, chain, D + 1, E, 1, 1, best_cap,
                               rounds, current, best_idx, &key, round_new_key, round_choice);
    *best_hard = key < 0 ? -1 : key >> 32;
    *best_soft = key < 0 ? -1 : key & 0xffffffffll;
    free(current);
    return r;
}

/* ------------------------------------------------------------------ O(1)-per-move delta scorer */
static inline int d_less(int64_t d, int64_t a, int64_t b, int64_t D, int64_t A, int64_t B) {
    if (A < 0) return 1;
    if (d != D) return d < D;
    if (a != A) return a < A;
    return b < B;
}

/* NOT a reference function: the reference has no delta scoring (SURVEY.md "three facts", 1).  This
 * is the checker's own counter formulation -- occupancy counters R[r], D1[c-r+n-1], D2[c+r] and the
 * SURVEY 8(a2) formulae, written with all four shared-line correction terms spelled out (the
 * device collapses them into one attack test; the two derivations are independent) -- so that
 * every candidate of a benchmark-sized neighbourhood (5e7 at n = 10 000, a column band at
 * n = 1e6) can be checked in seconds.  It is itself PROVEN against the literal clone + full
 * re-score (orc_nq_neighbourhood_deltas) on every candidate of small boards, permutations and
 * non-permutations, in tests/test_oracle_cpu.py; only then is it used as the big-size checker. */
typedef struct {
    int64_t n;
    int32_t *R, *D1, *D2;
} nq_counters;

static void nq_counters_build(nq_counters* c, const int64_t* rows, int64_t n) {
    c->n = n;
    c->R = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    c->D1 = (int32_t*)calloc((size_t)(2 * n + 1), sizeof(int32_t));
    c->D2 = (int32_t*)calloc((size_t)(2 * n + 1), sizeof(int32_t));
    for (int64_t j = 0; j < n; ++j) {
        c->R[rows[j]] += 1;
        c->D1[j - rows[j] + n - 1] += 1;
        c->D2[j + rows[j]] += 1;
    }
}

static void nq_counters_free(nq_counters* c) {
    free(c->R);
    free(c->D1);
    free(c->D2);
}

/* SURVEY 8(a2), swap i<j with r_i != r_j (rows term is 0: the row multiset is unchanged):
 *   d1 = (D1[i-rj] + D1[j-ri]) - (D1[i-ri] + D1[j-rj]) + 2 + [i-rj == j-ri] + [i-ri == j-rj]
 *   d2 = the same on D2 with c+r indices;   delta = 2 * (d1 + d2)                                  */
static inline int64_t nq_fast_swap_delta(const nq_counters* c, const int64_t* rows, int64_t i, int64_t j) {
    const int64_t n = c->n, ri = rows[i], rj = rows[j], o = n - 1;
    const int64_t d1 = (int64_t)c->D1[i - rj + o] + c->D1[j - ri + o] - c->D1[i - ri + o] - c->D1[j - rj + o] + 2 +
                       (i - rj == j - ri) + (i - ri == j - rj);
    const int64_t d2 = (int64_t)c->D2[i + rj] + c->D2[j + ri] - c->D2[i + ri] - c->D2[j + rj] + 2 +
                       (i + rj == j + ri) + (i + ri == j + rj);
    return 2 * (d1 + d2);
}

/* SURVEY 8(a2), change column c: r -> v, v != r */
static inline int64_t nq_fast_change_delta(const nq_counters* c, const int64_t* rows, int64_t col, int64_t v) {
    const int64_t n = c->n, r = rows[col], o = n - 1;
    return 2 * (((int64_t)c->R[v] + c->D1[col - v + o] + c->D2[col + v]) -
                ((int64_t)c->R[r] + c->D1[col - r + o] + c->D2[col + r]) + 3);
}

int64_t orc_nq_fast_band_deltas(const int64_t* rows, int64_t n, int kind, int64_t x_begin, int64_t x_end,
                                int threads, int64_t* delta) {
    nq_counters c;
    nq_counters_build(&c, rows, n);
    if (x_begin < 0) x_begin = 0;
    if (x_end > n) x_end = n;
    /* band-relative index: swap rows of the triangular enumeration, change rows of the n x n one */
    const int64_t base = kind == ORC_NQ_SWAP ? x_begin * n - x_begin * (x_begin + 1) / 2 : x_begin * n;
    int64_t total = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8) reduction(+ : total)
    for (int64_t x = x_begin; x < x_end; ++x) {
        if (kind == ORC_NQ_SWAP) {
            int64_t* out = delta + (x * n - x * (x + 1) / 2 - base);
            for (int64_t y = x + 1; y < n; ++y)
                out[y - x - 1] = rows[x] == rows[y] ? INT64_MAX : nq_fast_swap_delta(&c, rows, x, y);
            total += n - 1 - x;
        } else {
            int64_t* out = delta + (x * n - base);
            for (int64_t y = 0; y < n; ++y)
                out[y] = rows[x] == y ? INT64_MAX : nq_fast_change_delta(&c, rows, x, y);
            total += n;
        }
    }
    nq_counters_free(&c);
    return total;
}

int64_t orc_nq_fast_argmin(const int64_t* rows, int64_t n, int kind, int threads, int64_t* best_delta,
                           int64_t* best_a, int64_t* best_b) {
    nq_counters c;
    nq_counters_build(&c, rows, n);
    int64_t bd = INT64_MAX, ba = -1, bb = -1, scored = 0;
#pragma omp parallel num_threads(threads)
    {
        int64_t ld = INT64_MAX, la = -1, lb = -1, ls = 0;
#pragma omp for schedule(dynamic, 8) nowait
        for (int64_t x = 0; x < n; ++x) {
            const int64_t y0 = kind == ORC_NQ_SWAP ? x + 1 : 0;
            for (int64_t y = y0; y < n; ++y) {
                if (kind == ORC_NQ_SWAP ? rows[x] == rows[y] : rows[x] == y) continue;
                const int64_t d = kind == ORC_NQ_SWAP ? nq_fast_swap_delta(&c, rows, x, y)
                                                      : nq_fast_change_delta(&c, rows, x, y);
                ++ls;
                /* (delta, a, b) order: first minimum in enumeration order */
                if (d < ld || (d == ld && (x < la || (x == la && y < lb)))) {
                    ld = d;
                    la = x;
                    lb = y;
                }
            }
        }
#pragma omp critical
        {
            scored += ls;
            if (la >= 0 && (d_less(ld, la, lb, bd, ba, bb))) {
                bd = ld;
                ba = la;
                bb = lb;
            }
        }
    }
    nq_counters_free(&c);
    if (best_delta) *best_delta = bd;
    if (best_a) *best_a = ba;
    if (best_b) *best_b = bb;
    return scored;
}

/* "CPU delta" courtesy baseline (SURVEY 8(d) "CPU reference timing" (ii)):
 * the SAME counters + delta formulae as the GPU, one restart chain per host thread, full swap
 * neighbourhood per step, steepest descent with the (delta, i, j) rule.  Clearly NOT the reference
 * (which has no delta scoring); it isolates hardware from algorithm.  Returns candidates scored. */
int64_t orc_nq_delta_baseline(uint64_t seed, int64_t n, int chains, int steps, int threads, int64_t* checksum) {
    int64_t scored = 0, sum = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(+ : scored, sum)
    for (int k = 0; k < chains; ++k) {
        int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        orc_nq_init_perm(seed, (uint32_t)k, n, rows);
        nq_counters c;
        nq_counters_build(&c, rows, n);
        for (int s = 0; s < steps; ++s) {
            int64_t bd = INT64_MAX, bi = -1, bj = -1;
            for (int64_t i = 0; i + 1 < n; ++i)
                for (int64_t j = i + 1; j < n; ++j) {
                    if (rows[i] == rows[j]) continue;
                    const int64_t d = nq_fast_swap_delta(&c, rows, i, j);
                    ++scored;
                    if (d < bd) {
                        bd = d;
                        bi = i;
                        bj = j;
                    }
                }
            if (bi < 0) break;
            const int64_t ri = rows[bi], rj = rows[bj], o = n - 1;
            c.D1[bi - ri + o]--; c.D2[bi + ri]--; c.D1[bj - rj + o]--; c.D2[bj + rj]--;
            c.D1[bi - rj + o]++; c.D2[bi + rj]++; c.D1[bj - ri + o]++; c.D2[bj + ri]++;
            rows[bi] = rj;
            rows[bj] = ri;
            sum += bd;
        }
        nq_counters_free(&c);
        free(rows);
    }
    if (checksum) *checksum = sum;
    return scored;
}

/* orc_nq_neighbourhood_deltas with the candidates spread over OpenMP threads (the same clone +
 * full re-score per candidate; only the outer loop is parallel) */
int64_t orc_nq_neighbourhood_deltas_mt(const int64_t* rows, int64_t n, int kind, int threads, int64_t* delta) {
    const int64_t cur = orc_nq_score(rows, n);
#pragma omp parallel num_threads(threads)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
        for (int64_t x = 0; x < n; ++x) {
            const int64_t y0 = (kind == ORC_NQ_SWAP) ? x + 1 : 0;
            int64_t k = (kind == ORC_NQ_SWAP) ? x * n - x * (x + 1) / 2 : x * n;
            for (int64_t y = y0; y < n; ++y, ++k) {
                memcpy(cand, rows, sizeof(int64_t) * (size_t)n);
                delta[k] = nq_apply(cand, kind, x, y) ? orc_nq_score(cand, n) - cur : INT64_MAX;
            }
        }
        free(cand);
    }
    return kind == ORC_NQ_SWAP ? n * (n - 1) / 2 : n * n;
}

/* every candidate of the full scheduling neighbourhood, clone + full re-score each (OpenMP over
 * candidates), in the device's enumeration order INCLUDING identities (INT64_MAX): D*E change
 * entries (day outer, employee index inner) then D(D-1)/2 swap entries (d1 < d2 row-major) */
int64_t orc_es_neighbourhood_deltas_mt(const int64_t* a, int64_t D, int start_weekday, const int64_t* hol_emp,
                                       const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                                       int threads, int64_t* dhard, int64_t* dsoft) {
    int64_t h0, s0;
    if (orc_es_score(a, D, start_weekday, hol_emp, hol_day, n_hol, &h0, &s0)) return -1;
    const int64_t n_change = D * E, n_swap = D * (D - 1) / 2;
#pragma omp parallel num_threads(threads)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(D > 0 ? D : 1));
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < n_change + n_swap; ++k) {
            int kind;
            int64_t x, y;
            if (k < n_change) {
                kind = ORC_ES_CHANGE;
                x = k / E;
                y = k % E;
            } else {
                kind = ORC_ES_SWAP;
                int64_t r = k - n_change;
                x = 0;
                while (r >= D - 1 - x) {
                    r -= D - 1 - x;
                    ++x;
                }
                y = x + 1 + r;
            }
            memcpy(cand, a, sizeof(int64_t) * (size_t)D);
            if (!es_apply(cand, employees, kind, x, y)) {
                dhard[k] = INT64_MAX;
                dsoft[k] = INT64_MAX;
                continue;
            }
            int64_t h, s;
            orc_es_score(cand, D, start_weekday, hol_emp, hol_day, n_hol, &h, &s);
            dhard[k] = h - h0;
            dsoft[k] = s - s0;
        }
        free(cand);
    }
    return n_change + n_swap;
}

/* ================================================================== slot-generalised scheduling
 * EXTENSION -- NOT PINNED BY THE REFERENCE.  BASELINE configs[2] / [3] say "3 shifts/day" and the
 * north-star names "shift overlap" and "skill" tallies; the reference has one employee per calendar
 * DAY and neither notion (vocabulary table).  This is the in-repo CPU
 * full-re-score definition SURVEY section 7 asks for: slots = days x shifts_per_day, slot t =
 * day t / S, shift t % S, and it REDUCES to orc_es_score_terms at S = 1 with every employee
 * qualified (asserted on random rotas in tests/test_oracle_cpu.py):
 *   H1  a slot held by an employee on holiday that day                    (lib.rs:273-280 per slot)
 *   H2  the same employee on consecutive SLOTS, a[t] == a[t+1]            (:286-292 in slot units)
 *   H3  consecutive weekends, same shift: for Saturday i <= D-9 and each shift s, the four
 *       pairs {Sat, Sun} x {next Sat, next Sun} of shift s                (:295-315 per shift)
 *   H4  per 14-DAY window, employees holding more than 3 slots in it      (:318-327)
 *   S1  per 7-day window, employees holding more than 2 slots             (:330-339)
 *   S2  per weekday Mon-Fri with >= 2 employees, the minimum slot count   (:194-218)
 *   S3 / S4  max - min of total / weekend slots over employees present    (:345-365)
 *   X1  (new, hard) same-day overlap: pairs of slots of one day held by the same employee
 *   X2  (new, hard) skill: slots whose shift kind the employee is not qualified for
 * out[0..3] = H1..H4, out[4..7] = S1..S4, out[8] = X1, out[9] = X2;
 * hard = H1+H2+H3+H4+X1+X2, soft = S1+S2+S3+S4. */
typedef struct {
    int64_t D, S, T, E, nh;
    int wd;
    const int64_t *he, *hd, *emp, *skills; /* skills[k]: bit s set = employees[k] works shift s; NULL = all */
    int64_t *sid, *ssk;                    /* employee ids sorted, their skill masks */
} esx_prob;

static void esx_prob_init(esx_prob* p, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                          const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                          const int64_t* skills) {
    p->D = D; p->S = S; p->T = D * S; p->E = E; p->nh = n_hol; p->wd = start_weekday;
    p->he = hol_emp; p->hd = hol_day; p->emp = employees; p->skills = skills;
    p->sid = NULL; p->ssk = NULL;
    if (skills) { /* (id, skill) sorted by id for the lookups of X2 */
        p->sid = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E > 0 ? E : 1));
        p->ssk = (int64_t*)malloc(sizeof(int64_t) * (size_t)(E > 0 ? E : 1));
        for (int64_t k = 0; k < E; ++k) { /* insertion sort: E is small in the tests, 2000 in the bench */
            int64_t q = k;
            while (q > 0 && p->sid[q - 1] > employees[k]) {
                p->sid[q] = p->sid[q - 1];
                p->ssk[q] = p->ssk[q - 1];
                --q;
            }
            p->sid[q] = employees[k];
            p->ssk[q] = skills[k];
        }
    }
}
static void esx_prob_free(esx_prob* p) {
    free(p->sid);
    free(p->ssk);
}
static int64_t esx_skill_of(const esx_prob* p, int64_t id) {
    if (!p->skills) return -1; /* every shift */
    int64_t lo = 0, hi = p->E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (p->sid[mid] < id) lo = mid + 1;
        else hi = mid;
    }
    return (lo < p->E && p->sid[lo] == id) ? p->ssk[lo] : 0;
}

static int esx_terms(const esx_prob* p, const int64_t* a, int64_t out[10]) {
    const int64_t D = p->D, S = p->S, T = p->T;
    for (int k = 0; k < 10; ++k) out[k] = 0;
    { /* H1: a holiday covers every slot of its day; duplicate (employee, day) entries count once */
        unsigned char seen_small[1024];
        unsigned char* seen = T <= 1024 ? seen_small : (unsigned char*)malloc((size_t)T);
        memset(seen, 0, (size_t)(T > 0 ? T : 0));
        int bad = 0;
        for (int64_t k = 0; k < p->nh && !bad; ++k) {
            if (p->hd[k] < 0 || p->hd[k] >= D) { bad = 1; break; }
            for (int64_t s = 0; s < S; ++s) {
                const int64_t t = p->hd[k] * S + s;
                if (a[t] == p->he[k] && !seen[t]) {
                    seen[t] = 1;
                    out[0] += 1;
                }
            }
        }
        if (seen != seen_small) free(seen);
        if (bad) return -1;
    }
    for (int64_t t = 0; t + 2 <= T; ++t) out[1] += (a[t] == a[t + 1]); /* H2 */
    for (int64_t i = 0; i + 9 <= D; ++i) {                            /* H3 */
        if (!(es_is_weekend(p->wd, i) && es_is_weekend(p->wd, i + 1))) continue;
        for (int64_t s = 0; s < S; ++s) {
            const int64_t d1 = i * S + s, d2 = (i + 1) * S + s, d3 = (i + 7) * S + s, d4 = (i + 8) * S + s;
            out[2] += (a[d1] == a[d3]) + (a[d1] == a[d4]) + (a[d2] == a[d3]) + (a[d2] == a[d4]);
        }
    }
    for (int64_t i = 0; i + 14 <= D; ++i) out[3] += es_window_violations(a + i * S, 14 * S, 3); /* H4 */
    for (int64_t i = 0; i + 7 <= D; ++i) out[4] += es_window_violations(a + i * S, 7 * S, 2);   /* S1 */
    for (int wd = 0; wd < 5; ++wd) {                                                            /* S2 */
        int64_t distinct = 0, minc = INT64_MAX;
        for (int64_t t = 0; t < T; ++t) {
            if ((p->wd + t / S) % 7 != wd) continue;
            int first = 1;
            for (int64_t q = 0; q < t; ++q)
                if ((p->wd + q / S) % 7 == wd && a[q] == a[t]) { first = 0; break; }
            if (!first) continue;
            int64_t c = 0;
            for (int64_t q = t; q < T; ++q)
                if ((p->wd + q / S) % 7 == wd && a[q] == a[t]) ++c;
            ++distinct;
            if (c < minc) minc = c;
        }
        if (distinct >= 2) out[5] += minc;
    }
    int64_t present = 0, mind = INT64_MAX, maxd = INT64_MIN, minw = INT64_MAX, maxw = INT64_MIN; /* S3, S4 */
    for (int64_t t = 0; t < T; ++t) {
        int first = 1;
        for (int64_t q = 0; q < t; ++q)
            if (a[q] == a[t]) { first = 0; break; }
        if (!first) continue;
        int64_t slots = 0, wk = 0;
        for (int64_t q = t; q < T; ++q)
            if (a[q] == a[t]) {
                ++slots;
                wk += es_is_weekend(p->wd, q / S);
            }
        ++present;
        if (slots < mind) mind = slots;
        if (slots > maxd) maxd = slots;
        if (wk < minw) minw = wk;
        if (wk > maxw) maxw = wk;
    }
    if (present >= 2) {
        out[6] = maxd - mind;
        out[7] = maxw - minw;
    }
    for (int64_t d = 0; d < D; ++d) /* X1 same-day overlap */
        for (int64_t s = 0; s < S; ++s)
            for (int64_t s2 = s + 1; s2 < S; ++s2) out[8] += (a[d * S + s] == a[d * S + s2]);
    if (p->skills)                  /* X2 skill */
        for (int64_t t = 0; t < T; ++t)
            if (!((esx_skill_of(p, a[t]) >> (t % S)) & 1)) out[9] += 1;
    return 0;
}

static int esx_score(const esx_prob* p, const int64_t* a, int64_t* hard, int64_t* soft) {
    int64_t t[10];
    const int rc = esx_terms(p, a, t);
    if (rc) return rc;
    *hard = t[0] + t[1] + t[2] + t[3] + t[8] + t[9];
    *soft = t[4] + t[5] + t[6] + t[7];
    return 0;
}

int orc_esx_score_terms(const int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                        const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                        const int64_t* skills, int64_t out[10]) {
    esx_prob p;
    esx_prob_init(&p, D, S, start_weekday, hol_emp, hol_day, n_hol, employees, E, skills);
    const int rc = esx_terms(&p, a, out);
    esx_prob_free(&p);
    return rc;
}

/* every candidate (clone + full re-score, OpenMP over candidates), device enumeration order incl.
 * identities: T*E change entries (slot outer, employee index inner), then T(T-1)/2 swaps */
int64_t orc_esx_neighbourhood_deltas_mt(const int64_t* a, int64_t D, int64_t S, int start_weekday,
                                        const int64_t* hol_emp, const int64_t* hol_day, int64_t n_hol,
                                        const int64_t* employees, int64_t E, const int64_t* skills, int threads,
                                        int64_t* dhard, int64_t* dsoft) {
    esx_prob p;
    esx_prob_init(&p, D, S, start_weekday, hol_emp, hol_day, n_hol, employees, E, skills);
    const int64_t T = p.T;
    int64_t h0, s0;
    if (esx_score(&p, a, &h0, &s0)) {
        esx_prob_free(&p);
        return -1;
    }
    const int64_t n_change = T * E, n_swap = T * (T - 1) / 2;
#pragma omp parallel num_threads(threads)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(T > 0 ? T : 1));
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < n_change + n_swap; ++k) {
            int kind;
            int64_t x, y;
            if (k < n_change) {
                kind = ORC_ES_CHANGE;
                x = k / E;
                y = k % E;
            } else {
                kind = ORC_ES_SWAP;
                int64_t r = k - n_change;
                x = 0;
                while (r >= T - 1 - x) {
                    r -= T - 1 - x;
                    ++x;
                }
                y = x + 1 + r;
            }
            memcpy(cand, a, sizeof(int64_t) * (size_t)T);
            if (!es_apply(cand, employees, kind, x, y)) {
                dhard[k] = INT64_MAX;
                dsoft[k] = INT64_MAX;
                continue;
            }
            int64_t h, s;
            esx_score(&p, cand, &h, &s);
            dhard[k] = h - h0;
            dsoft[k] = s - s0;
        }
        free(cand);
    }
    esx_prob_free(&p);
    return n_change + n_swap;
}

/* LocalSearch::execute (local_search.rs:301-342) over the full change + swap neighbourhood of the
 * slot-generalised problem; same rule as orc_es_local_search (first minimum in enumeration order) */
int64_t orc_esx_local_search(int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                             const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                             const int64_t* skills, uint64_t allow_no_improvement_for, uint64_t max_iterations,
                             int threads, int64_t* best_hard, int64_t* best_soft, int64_t* current_out,
                             int64_t* trace_kind, int64_t* trace_x, int64_t* trace_y, int64_t* trace_hard,
                             int64_t* trace_soft, int64_t cap) {
    esx_prob p;
    esx_prob_init(&p, D, S, start_weekday, hol_emp, hol_day, n_hol, employees, E, skills);
    const int64_t T = p.T;
    const size_t bytes = sizeof(int64_t) * (size_t)(T > 0 ? T : 1);
    int64_t* current = (int64_t*)malloc(bytes);
    int64_t* best = (int64_t*)malloc(bytes);
    const int64_t n_all = T * E + T * (T - 1) / 2;
    int64_t* dh = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_all > 0 ? n_all : 1));
    int64_t* ds = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_all > 0 ? n_all : 1));
    memcpy(current, a, bytes);
    int64_t ch, cs;
    esx_score(&p, current, &ch, &cs);
    memcpy(best, current, bytes);
    int64_t bh = ch, bs = cs, steps = 0;
    uint64_t no_improvement_for = 0;
    for (uint64_t it = 0; it < max_iterations; ++it) {
        if (ch == 0 && cs == 0) {
            memcpy(best, current, bytes);
            bh = ch;
            bs = cs;
            break;
        }
        orc_esx_neighbourhood_deltas_mt(current, D, S, start_weekday, hol_emp, hol_day, n_hol, employees, E, skills,
                                        threads, dh, ds);
        int64_t kbest = -1;
        for (int64_t k = 0; k < n_all; ++k) {
            if (dh[k] == INT64_MAX) continue;
            if (kbest < 0 || dh[k] < dh[kbest] || (dh[k] == dh[kbest] && ds[k] < ds[kbest])) kbest = k;
        }
        if (kbest < 0) break;
        int64_t kind, x, y;
        if (kbest < T * E) {
            kind = ORC_ES_CHANGE;
            x = kbest / E;
            y = kbest % E;
        } else {
            kind = ORC_ES_SWAP;
            int64_t r = kbest - T * E;
            x = 0;
            while (r >= T - 1 - x) {
                r -= T - 1 - x;
                ++x;
            }
            y = x + 1 + r;
        }
        const int64_t nh = ch + dh[kbest], ns = cs + ds[kbest];
        const int improved = nh < ch || (nh == ch && ns < cs);
        if (!improved) {
            no_improvement_for += 1;
            if (no_improvement_for >= allow_no_improvement_for) break;
        }
        es_apply(current, employees, (int)kind, x, y);
        ch = nh;
        cs = ns;
        if (improved) {
            memcpy(best, current, bytes);
            bh = nh;
            bs = ns;
            no_improvement_for = 0;
        }
        if (steps < cap) {
            if (trace_kind) trace_kind[steps] = kind;
            if (trace_x) trace_x[steps] = x;
            if (trace_y) trace_y[steps] = y;
            if (trace_hard) trace_hard[steps] = nh;
            if (trace_soft) trace_soft[steps] = ns;
        }
        ++steps;
    }
    memcpy(a, best, bytes);
    if (best_hard) *best_hard = bh;
    if (best_soft) *best_soft = bs;
    if (current_out) memcpy(current_out, current, bytes);
    free(current);
    free(best);
    free(dh);
    free(ds);
    esx_prob_free(&p);
    return steps;
}

int64_t orc_esx_baseline_sample(const int64_t* a, int64_t D, int64_t S, int start_weekday, const int64_t* hol_emp,
                                const int64_t* hol_day, int64_t n_hol, const int64_t* employees, int64_t E,
                                const int64_t* skills, int kind, const int64_t* x, const int64_t* y,
                                int64_t n_moves, int threads, int64_t* checksum) {
    esx_prob p;
    esx_prob_init(&p, D, S, start_weekday, hol_emp, hol_day, n_hol, employees, E, skills);
    int64_t sum = 0, h0, s0;
    if (esx_score(&p, a, &h0, &s0)) {
        esx_prob_free(&p);
        return -1;
    }
#pragma omp parallel num_threads(threads) reduction(+ : sum)
    {
        int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p.T > 0 ? p.T : 1));
#pragma omp for schedule(static)
        for (int64_t k = 0; k < n_moves; ++k) {
            memcpy(cand, a, sizeof(int64_t) * (size_t)p.T);
            if (es_apply(cand, employees, kind, x[k], y[k])) {
                int64_t h, s;
                esx_score(&p, cand, &h, &s);
                sum += (h - h0) * 1000 + (s - s0);
            }
        }
        free(cand);
    }
    esx_prob_free(&p);
    if (checksum) *checksum = sum;
    return n_moves;
}
