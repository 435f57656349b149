// C ABI for the employee-scheduling plug-in (included by cs_api.cu; shares its helpers).
#pragma once
#include <cstdlib>
#include <algorithm>

#include "es_kernels.cuh"

struct cs_es_handle {
    cs_es_config cfg{};
    int device = 0;
    int D = 0, S = 1, T = 0;  // days, shifts per day, scored slots (T = D * S)
    int W = 1;                // 64-bit words per slot mask
    bool multi = false;       // slot-generalised extension (S > 1 or a skill table with a gap)
    int wd = 1;               // words that hold the day-indexed sets: ceil(D / 64) <= W
    bool pk = false;          // <= 38 days: the 14- and 7-day window sets share one word (ES_PK_MAX_DAYS)
    int dp = 0;               // T rounded up to 4 (stride of the per-slot constant tables)
    int stride = 0;           // T + 1 slots (phantom last)
    int threads = 128;
    int grid_cap = 1;
    size_t smem = 0;
    EsConstT<3> K{};          // masks at the widest width; truncated to W words for the kernels
    std::vector<int64_t> ids;  // sorted employee ids (BTreeSet order)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint16_t* d_a = nullptr;
    uint16_t* d_best_a = nullptr;
    u64* d_hol = nullptr;     // [E][W] holiday slot mask per employee
    u64* d_unsk = nullptr;    // [E][W] multi: slots whose shift kind the employee is not qualified for
    u64* d_cnt2 = nullptr;    // [E][2W] multi: holiday + unskilled bits per slot as 2-bit counts (pass B)
    u64* d_slotc = nullptr;   // [4][dp][W] PART, CONT14, CONT7, PARTX
    uint16_t* d_tri = nullptr;  // [2][n_swap] (d1 << 8 | d2): enumeration order, then scan order
    size_t n_swap = 0;
    EsChainState* d_st = nullptr;
    EsTraceEntry* d_trace = nullptr;
    unsigned int* d_work = nullptr;
    unsigned long long* d_totals = nullptr;
    EsStats* d_stats = nullptr;
    EsStats* h_stats = nullptr;
    unsigned long long* h_totals = nullptr;
    EsChainState* h_states = nullptr;  // pinned [n_chains]: chain states for get_scores / get_status / best
    long long* d_stage64 = nullptr;  // device staging: employee ids <-> dense indices are converted on the device
    long long* d_ids = nullptr;      // [E] sorted employee ids
    int* d_bad = nullptr;
    size_t stage_chains = 0;
    // double-buffered input staging (cs_es_set_chains_async / cs_es_commit_chains)
    long long* d_stage_async = nullptr;
    size_t stage_async_elems = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_upload = nullptr;
    uint32_t async_first = 0, async_count = 0;
    bool async_pending = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool scored = false;
    bool run_pending = false;  // a step was enqueued and not yet waited for
    bool ref_mode = false;          // CS_ES_FLAG_REFERENCE_PROPOSER
    unsigned long long window = 100;  // window_size, examples/employee-scheduling/src/main.rs:26
    IlsHost ils;
    std::string err;
};

namespace {

template <int W>
Bits<W> es_narrow(const Bits<3>& b) {
    Bits<W> r;
    for (int i = 0; i < W; ++i) r.w[i] = b.w[i];
    return r;
}

template <int W>
EsConstT<W> es_const(const cs_es_handle* h) {
    const EsConstT<3>& k = h->K;
    EsConstT<W> r;
    r.D = k.D; r.S = k.S; r.T = k.T; r.E = k.E; r.start_wd = k.start_wd; r.n14 = k.n14; r.n7 = k.n7;
    r.valid = es_narrow<W>(k.valid);
    r.wkend = es_narrow<W>(k.wkend);
    r.satf = es_narrow<W>(k.satf);
    r.sd1 = es_narrow<W>(k.sd1);
    r.sd2 = es_narrow<W>(k.sd2);
    for (int i = 0; i < 5; ++i) r.wd[i] = es_narrow<W>(k.wd[i]);
    return r;
}

// run f(integral_constant<int, W>, bool_constant<MULTI>) for the handle's kernel variant
template <typename F>
void es_dispatch(const cs_es_handle* h, F&& f) {
    using std::integral_constant;
    switch (h->W * 2 + (h->multi ? 1 : 0)) {
        case 2: f(integral_constant<int, 1>{}, std::false_type{}); break;
        case 3: f(integral_constant<int, 1>{}, std::true_type{}); break;
        case 4: f(integral_constant<int, 2>{}, std::false_type{}); break;
        case 5: f(integral_constant<int, 2>{}, std::true_type{}); break;
        case 6: f(integral_constant<int, 3>{}, std::false_type{}); break;
        default: f(integral_constant<int, 3>{}, std::true_type{}); break;
    }
}

template <int W>
EsParamsT<W> es_params(cs_es_handle* h, int first, int count) {
    EsParamsT<W> p{};
    p.K = es_const<W>(h);
    p.first_chain = first;
    p.n_chains = count;
    p.stride = h->stride;
    p.a = h->d_a;
    p.best_a = h->d_best_a;
    p.hol = (const Bits<W>*)h->d_hol;
    p.unsk = (const Bits<W>*)h->d_unsk;
    p.cnt2 = h->d_cnt2;
    p.slotc = (const Bits<W>*)h->d_slotc;
    p.tri = h->d_tri;
    p.tri_scan = h->d_tri + h->n_swap;
    p.st = h->d_st;
    p.trace = h->d_trace;
    p.trace_cap = (int)h->cfg.trace_capacity;
    p.work_counter = h->d_work;
    p.totals = h->d_totals;
    p.seed = h->cfg.seed;
    p.chain_offset = h->cfg.chain_offset;
    p.window = h->window;
    p.max_draws = 1ull << 16;  // bounds the reference's endless iterator (it spins forever if every candidate is tabu)
    return p;
}

// what differs between launches of the step kernel
struct EsRun {
    int first = 0, count = 0;
    unsigned long long max_steps = 0, allow = 0;
    int ls_mode = 0;
    long long* dump_h = nullptr;
    long long* dump_s = nullptr;
    const unsigned int* skip = nullptr;
};

void es_free(cs_es_handle* h) {
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->ils.release();
    cudaFree(h->d_a);
    cudaFree(h->d_best_a);
    cudaFree(h->d_hol);
    cudaFree(h->d_unsk);
    cudaFree(h->d_cnt2);
    cudaFree(h->d_slotc);
    cudaFree(h->d_tri);
    cudaFree(h->d_st);
    cudaFree(h->d_trace);
    cudaFree(h->d_work);
    cudaFree(h->d_totals);
    cudaFree(h->d_stats);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    if (h->h_totals) cudaFreeHost(h->h_totals);
    if (h->h_states) cudaFreeHost(h->h_states);
    cudaFree(h->d_stage64);
    cudaFree(h->d_stage_async);
    if (h->ev_upload) cudaEventDestroy(h->ev_upload);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaFree(h->d_ids);
    cudaFree(h->d_bad);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
}

void es_check_range(cs_es_handle* h, uint32_t first, uint32_t count) {
    REQUIRE(count > 0 && first < h->cfg.n_chains && count <= h->cfg.n_chains - first,
            "chain range outside [0, n_chains)");
}

// the step kernel of this handle: the reference proposer's variant, or the full scan with its day-indexed
// sets cut to ceil(days / 64) words (several shifts per day: fewer days than slots)
template <int W, bool MULTI>
void (*es_step_fn(const cs_es_handle* h))(EsParamsT<W>) {
    const bool pk = h->pk;
    if constexpr (!MULTI) {
        if (h->ref_mode) return es_step_kernel<W, false, true>;
        if constexpr (W == 1) {
            if (pk) return es_step_kernel<1, false, false, 1, true>;
        }
        return es_step_kernel<W, false, false>;
    } else {
        if (h->wd == 1 && pk) return es_step_kernel<W, true, false, 1, true>;
        if (h->wd == 1) return es_step_kernel<W, true, false, 1>;
        if constexpr (W >= 3) {
            if (h->wd == 2) return es_step_kernel<W, true, false, 2>;
        }
        return es_step_kernel<W, true, false, W>;
    }
}

void es_launch_step(cs_es_handle* h, int grid, const EsRun& r) {
    es_dispatch(h, [&](auto w, auto m) {
        constexpr int W = decltype(w)::value;
        constexpr bool MULTI = decltype(m)::value;
        EsParamsT<W> p = es_params<W>(h, r.first, r.count);
        p.max_steps = r.max_steps;
        p.allow_no_improve = r.allow;
        p.ls_mode = r.ls_mode;
        p.dump_h = r.dump_h;
        p.dump_s = r.dump_s;
        p.skip = r.skip;
        es_step_fn<W, MULTI>(h)<<<grid, h->threads, h->smem, h->stream>>>(p);
    });
    CU(cudaGetLastError());
}

void es_rescore(cs_es_handle* h, int first, int count) {
    const int grid = count < h->grid_cap ? count : h->grid_cap;
    es_dispatch(h, [&](auto w, auto m) {
        constexpr int W = decltype(w)::value;
        constexpr bool MULTI = decltype(m)::value;
        EsParamsT<W> p = es_params<W>(h, first, count);
        es_rescore_kernel<W, MULTI><<<grid, h->threads, h->smem, h->stream>>>(p);
    });
    CU(cudaGetLastError());
}

void es_refresh_stats(cs_es_handle* h) {
    es_stats_kernel<<<1, 1024, 0, h->stream>>>(h->d_st, (int)h->cfg.n_chains, h->cfg.chain_offset,
                                                h->d_stats);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(EsStats), cudaMemcpyDeviceToHost, h->stream));
}

void es_run_enqueue(cs_es_handle* h, int first, int count, unsigned long long max_steps,
                    unsigned long long allow, int ls_mode) {
    if (!h->scored) throw StateFail{"chains have no solution yet: call cs_es_init_random or cs_es_set_chains first"};
    EsRun r;
    r.first = first;
    r.count = count;
    r.max_steps = max_steps;
    r.allow = allow;
    r.ls_mode = ls_mode;
    CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
    CU(cudaMemsetAsync(h->d_totals, 0, 2 * sizeof(unsigned long long), h->stream));
    const int grid = count < h->grid_cap ? count : h->grid_cap;
    CU(cudaEventRecord(h->ev0, h->stream));
    es_launch_step(h, grid, r);
    es_refresh_stats(h);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaMemcpyAsync(h->h_totals, h->d_totals, 2 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, h->stream));
    h->run_pending = true;
}

void es_run_wait(cs_es_handle* h, cs_es_step_stats* stats) {
    if (!h->run_pending) throw StateFail{"no enqueued step to wait for"};
    CU(cudaStreamSynchronize(h->stream));
    h->run_pending = false;
    if (stats) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        stats->moves_scored = h->h_totals[0];
        stats->steps_accepted = h->h_totals[1];
        stats->best_hard = h->h_stats->best_hard;
        stats->best_soft = h->h_stats->best_soft;
        stats->best_chain = h->h_stats->best_chain;
        stats->chains_at_best = h->h_stats->chains_at_best;
        stats->chains_feasible = h->h_stats->chains_feasible;
        stats->device_ms = ms;
        stats->kernel_launches = 2;
    }
}

void es_run(cs_es_handle* h, int first, int count, unsigned long long max_steps,
            unsigned long long allow, int ls_mode, cs_es_step_stats* stats) {
    es_run_enqueue(h, first, count, max_steps, allow, ls_mode);
    es_run_wait(h, stats);
}

int es_index_of(cs_es_handle* h, int64_t id) {
    auto it = std::lower_bound(h->ids.begin(), h->ids.end(), id);
    if (it == h->ids.end() || *it != id) return -1;
    return (int)(it - h->ids.begin());
}

// employee id -> dense index by binary search over the sorted id table, on the device
__global__ void es_ids_to_index_kernel(const long long* __restrict__ rows, uint16_t* __restrict__ a, size_t total,
                                       const long long* __restrict__ ids, int E, int* bad) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const long long id = rows[k];
        int lo = 0, hi = E;
        if (id >= 0 && id < E && ids[id] == id) lo = (int)id;  // dense ids 0..E-1: no search
        else {
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ids[mid] < id) lo = mid + 1;
                else hi = mid;
            }
            if (lo >= E || ids[lo] != id) {
                *bad = 1;
                lo = 0;
            }
        }
        a[k] = (uint16_t)lo;
    }
}
__global__ void es_index_to_ids_kernel(const uint16_t* __restrict__ a, long long* __restrict__ rows, size_t total,
                                       const long long* __restrict__ ids) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x)
        rows[k] = ids[a[k]];
}

void es_upload(cs_es_handle* h, uint32_t first, uint32_t count, const int64_t* rows) {
    const size_t stride = h->stride;
    CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = (uint32_t)std::min<size_t>(count - done, h->stage_chains);
        const size_t total = (size_t)c * stride;
        CU(cudaMemcpyAsync(h->d_stage64, rows + (size_t)done * stride, total * sizeof(long long),
                           cudaMemcpyHostToDevice, h->stream));
        es_ids_to_index_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, h->stream>>>(
            h->d_stage64, h->d_a + (size_t)(first + done) * stride, total, h->d_ids, (int)h->ids.size(), h->d_bad);
        CU(cudaGetLastError());
        done += c;
    }
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    REQUIRE(!bad, "solution names an employee id that is not in the employee table");
}

void es_download(cs_es_handle* h, const uint16_t* src, uint32_t first, uint32_t count, int64_t* rows) {
    const size_t stride = h->stride;
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = (uint32_t)std::min<size_t>(count - done, h->stage_chains);
        const size_t total = (size_t)c * stride;
        es_index_to_ids_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, h->stream>>>(
            src + (size_t)(first + done) * stride, h->d_stage64, total, h->d_ids);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(rows + (size_t)done * stride, h->d_stage64, total * sizeof(long long),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        done += c;
    }
}

// chain states through the pinned staging buffer (valid until the next call on the handle)
const EsChainState* es_states(cs_es_handle* h, uint32_t first, uint32_t count) {
    CU(cudaMemcpyAsync(h->h_states, h->d_st + first, count * sizeof(EsChainState),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return h->h_states;
}

inline void es_set_bit(Bits<3>& b, int d) { b.w[d >> 6] |= 1ull << (d & 63); }

}  // namespace

extern "C" int32_t cs_es_create_ex(const cs_es_config* cfg, const int64_t* employee_ids,
                                   const int64_t* hol_emp, const int64_t* hol_day, uint64_t n_hol,
                                   uint32_t shifts_per_day, const uint32_t* skills, cs_es_handle** out) {
    if (!cfg || !out || !employee_ids) return CS_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->n_days < 1 || cfg->n_employees < 1 || cfg->n_employees > 65535 || cfg->n_chains < 1 ||
        cfg->start_weekday > 6 || shifts_per_day < 1)
        return CS_ERR_INVALID_ARG;
    if (shifts_per_day > CS_ES_MAX_SHIFTS || (uint64_t)cfg->n_days * shifts_per_day > CS_ES_MAX_SLOTS)
        return CS_ERR_UNSUPPORTED;
    if (n_hol && (!hol_emp || !hol_day)) return CS_ERR_INVALID_ARG;
    if ((uint64_t)cfg->chain_offset + cfg->n_chains > 0xffffffffull) return CS_ERR_INVALID_ARG;
    // argument validation that needs no device comes first (the reference would panic)
    std::vector<int64_t> ids(employee_ids, employee_ids + cfg->n_employees);
    std::sort(ids.begin(), ids.end());
    if (std::adjacent_find(ids.begin(), ids.end()) != ids.end()) return CS_ERR_INVALID_ARG;
    for (uint64_t k = 0; k < n_hol; ++k)
        if (hol_day[k] < 0 || hol_day[k] >= (int64_t)cfg->n_days) return CS_ERR_INVALID_ARG;
    const int D = (int)cfg->n_days, S = (int)shifts_per_day, T = D * S, E = (int)cfg->n_employees;
    // the extension kernels are only needed for several shifts per day or when somebody lacks a skill
    bool skill_gap = false;
    if (skills)
        for (int k = 0; k < E; ++k)
            if ((skills[k] & ((1u << S) - 1u)) != ((1u << S) - 1u)) skill_gap = true;
    const bool multi = S > 1 || skill_gap;
    const bool ref_mode = (cfg->flags & CS_ES_FLAG_REFERENCE_PROPOSER) != 0;
    if (ref_mode && multi) return CS_ERR_UNSUPPORTED;  // the reference's proposer belongs to the reference's rota
    int ndev = cs_device_count();
    if (ndev <= 0) return CS_ERR_NO_DEVICE;
    int dev = cfg->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) return CS_ERR_NO_DEVICE;
    }
    if (dev >= ndev) return CS_ERR_INVALID_ARG;
    cs_es_handle* h = new (std::nothrow) cs_es_handle();
    if (!h) return CS_ERR_OOM;
    h->cfg = *cfg;
    h->device = dev;
    h->ids = ids;
    h->D = D;
    h->S = S;
    h->T = T;
    h->W = (T + 63) / 64;
    h->multi = multi;
    h->wd = (D + 63) / 64;
    h->pk = D <= ES_PK_MAX_DAYS && !std::getenv("CS_ES_NO_PACKED_DAY_SETS");  // the knob is for A/B timing
    h->dp = (T + 3) & ~3;
    const int32_t rc = guarded(h, [&] {
        const int W = h->W, dp = h->dp;
        h->stride = T + 1;
        EsConstT<3>& K = h->K;
        K.D = D;
        K.S = S;
        K.T = T;
        K.E = E;
        K.start_wd = (int)cfg->start_weekday;
        K.n14 = D >= 14 ? D - 13 : 0;
        K.n7 = D >= 7 ? D - 6 : 0;
        K.valid = Bits<3>::lowmask(T);
        K.wkend = K.satf = K.sd1 = K.sd2 = Bits<3>::zero();
        for (int w = 0; w < 5; ++w) K.wd[w] = Bits<3>::zero();
        auto slot = [&](int day, int sh) { return day * S + sh; };
        for (int t = 0; t < T; ++t) {
            const int day = t / S, sh = t % S;
            const int w = (K.start_wd + day) % 7;
            if (w < 5) es_set_bit(K.wd[w], t);
            else es_set_bit(K.wkend, t);
            if (w == 5 && day + 9 <= D) es_set_bit(K.satf, t);  // windows(9) start, lib.rs:295-302
            if (sh + 1 < S) es_set_bit(K.sd1, t);
            if (sh + 2 < S) es_set_bit(K.sd2, t);
        }
        // per-slot constants of the hot loop: H2/H3 partner slots, the window starts holding the
        // slot's day, the other slots of the same day
        std::vector<Bits<3>> sc(4 * (size_t)dp, Bits<3>::zero());
        for (int t = 0; t < T; ++t) {
            const int day = t / S;
            if (t > 0) es_set_bit(sc[t], t - 1);          // H2, lib.rs:286-292 (consecutive slots)
            if (t + 1 < T) es_set_bit(sc[t], t + 1);
            for (int w = 0; w < K.n14; ++w)
                if (w <= day && day <= w + 13) es_set_bit(sc[dp + t], w);
            for (int w = 0; w < K.n7; ++w)
                if (w <= day && day <= w + 6) es_set_bit(sc[2 * dp + t], w);
            for (int sh = 0; sh < S; ++sh)
                if (slot(day, sh) != t) es_set_bit(sc[3 * dp + t], slot(day, sh));
        }
        for (int sat = 0; sat + 9 <= D; ++sat) {          // H3, lib.rs:295-315, per shift
            if ((K.start_wd + sat) % 7 != 5) continue;
            for (int sh = 0; sh < S; ++sh) {
                const int pr[4][2] = {{sat, sat + 7}, {sat, sat + 8}, {sat + 1, sat + 7}, {sat + 1, sat + 8}};
                for (auto& q : pr) {
                    es_set_bit(sc[slot(q[0], sh)], slot(q[1], sh));
                    es_set_bit(sc[slot(q[1], sh)], slot(q[0], sh));
                }
            }
        }
        std::vector<u64> slotc(4 * (size_t)dp * W);
        for (size_t k = 0; k < sc.size(); ++k)
            for (int i = 0; i < W; ++i) slotc[k * W + i] = sc[k].w[i];
        // swap r -> (d1 << 8 | d2), enumeration order d1 < d2 row-major; then the same pairs in SCAN
        // order: pairs closer than 14 days (whose windows overlap and need the both-day corrections)
        // first, so a warp's 32 swaps take the same path; move ids stay row-major
        h->n_swap = (size_t)T * (T - 1) / 2;
        std::vector<uint16_t> tri(2 * h->n_swap + 8, 0);
        {
            size_t r = 0;
            for (int d1 = 0; d1 < T; ++d1)
                for (int d2 = d1 + 1; d2 < T; ++d2) tri[r++] = (uint16_t)((d1 << 8) | d2);
            r = h->n_swap;
            for (int pass = 0; pass < 2; ++pass)
                for (int d1 = 0; d1 < T; ++d1)
                    for (int d2 = d1 + 1; d2 < T; ++d2)
                        if ((d2 / S - d1 / S < 14) == (pass == 0)) tri[r++] = (uint16_t)((d1 << 8) | d2);
        }
        std::vector<u64> hol((size_t)E * W, 0ull), unsk;
        for (uint64_t k = 0; k < n_hol; ++k) {
            const int idx = es_index_of(h, hol_emp[k]);
            if (idx < 0) continue;  // unknown employees never match a slot
            for (int sh = 0; sh < S; ++sh) {
                const int t = slot((int)hol_day[k], sh);
                hol[(size_t)idx * W + (t >> 6)] |= 1ull << (t & 63);
            }
        }
        if (multi) {
            unsk.assign((size_t)E * W, 0ull);
            if (skills)
                for (int k = 0; k < E; ++k) {
                    const int idx = es_index_of(h, employee_ids[k]);
                    for (int t = 0; t < T; ++t)
                        if (!((skills[k] >> (t % S)) & 1u)) unsk[(size_t)idx * W + (t >> 6)] |= 1ull << (t & 63);
                }
        }
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        es_dispatch(h, [&](auto w, auto m) {
            constexpr int W_ = decltype(w)::value;
            constexpr bool MULTI = decltype(m)::value;
            h->smem = es_smem_bytes<W_, MULTI>(T, E);
        });
        REQUIRE(h->smem <= (size_t)prop.sharedMemPerBlockOptin, "employee table too large for shared memory");
        const long long moves = (long long)T * E + (long long)T * (T - 1) / 2;
        if (W == 1 && !multi) {
            // per-day phases keep <= 64 threads busy between barriers, so wide CTAs mostly wait (measured: 128 beats
            // 256 by 6 % at 56 x 2000, 32 beats 64 by 10 % at 28 x 50)
            h->threads = moves <= 2048 ? 32 : moves <= 8192 ? 64 : 128;
        } else {
            // wider masks: the tables take most of an SM's shared memory, so few CTAs are resident -- make them wide
            // (measured: 84 slots x 50: 64 / 128 / 256 threads -> 8.4 / 9.7 / 8.6e10 moves/s; 168 x 2000: 256 / 384 / 512 ->
            // 5.4 / 6.5 / 7.0e11)
            h->threads = moves <= 2048 ? 64 : h->smem > (size_t)96 * 1024 ? 512 : moves <= 16384 ? 128 : 256;
        }
        const int max_threads = (W == 1 && !multi) ? 256 : 512;
        if (const char* t = std::getenv("CS_ES_THREADS")) {  // tuning knob: CTA size (multiple of 32)
            const int v = std::atoi(t);
            if (v >= 32 && v <= max_threads && v % 32 == 0) h->threads = v;
        }
        h->ref_mode = ref_mode;
        int per_sm = 1;
        es_dispatch(h, [&](auto w, auto m) {
            constexpr int W_ = decltype(w)::value;
            constexpr bool MULTI = decltype(m)::value;
            allow_max_smem(es_step_fn<W_, MULTI>(h), prop);
            allow_max_smem(es_rescore_kernel<W_, MULTI>, prop);
            allow_max_smem(es_eval_kernel<W_, MULTI>, prop);
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, es_step_fn<W_, MULTI>(h), h->threads,
                                                             h->smem));
        });
        if (per_sm < 1) per_sm = 1;
        h->grid_cap = prop.multiProcessorCount * per_sm;
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
        const size_t nc = cfg->n_chains;
        CU(cudaMalloc(&h->d_a, nc * h->stride * sizeof(uint16_t)));
        CU(cudaMalloc(&h->d_best_a, nc * h->stride * sizeof(uint16_t)));
        CU(cudaMalloc(&h->d_hol, hol.size() * sizeof(u64)));
        if (multi) CU(cudaMalloc(&h->d_unsk, unsk.size() * sizeof(u64)));
        CU(cudaMalloc(&h->d_slotc, slotc.size() * sizeof(u64)));
        CU(cudaMalloc(&h->d_tri, tri.size() * sizeof(uint16_t)));
        CU(cudaMalloc(&h->d_st, nc * sizeof(EsChainState)));
        if (cfg->trace_capacity)
            CU(cudaMalloc(&h->d_trace, nc * cfg->trace_capacity * sizeof(EsTraceEntry)));
        CU(cudaMalloc(&h->d_work, sizeof(unsigned int)));
        CU(cudaMalloc(&h->d_totals, 2 * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_stats, sizeof(EsStats)));
        CU(cudaMallocHost(&h->h_stats, sizeof(EsStats)));
        CU(cudaMallocHost(&h->h_totals, 2 * sizeof(unsigned long long)));
        CU(cudaMallocHost(&h->h_states, (size_t)cfg->n_chains * sizeof(EsChainState)));
        h->stage_chains = std::min<size_t>(nc, std::max<size_t>(1, (size_t)(64u << 20) / (h->stride * 8)));
        CU(cudaMalloc(&h->d_stage64, h->stage_chains * h->stride * sizeof(long long)));
        CU(cudaMalloc(&h->d_ids, (size_t)E * sizeof(long long)));
        CU(cudaMalloc(&h->d_bad, sizeof(int)));
        CU(cudaMemcpy(h->d_ids, h->ids.data(), (size_t)E * sizeof(long long), cudaMemcpyHostToDevice));
        CU(cudaEventCreate(&h->ev0));
        CU(cudaEventCreate(&h->ev1));
        CU(cudaMemcpy(h->d_hol, hol.data(), hol.size() * sizeof(u64), cudaMemcpyHostToDevice));
        if (multi) {
            CU(cudaMemcpy(h->d_unsk, unsk.data(), unsk.size() * sizeof(u64), cudaMemcpyHostToDevice));
            std::vector<u64> cnt2((size_t)E * 2 * W, 0ull);  // slot t -> bits 2t, 2t+1 = holiday bit + unskilled bit
            for (int e = 0; e < E; ++e)
                for (int t = 0; t < T; ++t) {
                    const u64 c = ((hol[(size_t)e * W + (t >> 6)] >> (t & 63)) & 1ull) +
                                  ((unsk[(size_t)e * W + (t >> 6)] >> (t & 63)) & 1ull);
                    cnt2[(size_t)e * 2 * W + (t >> 5)] |= c << (2 * (t & 31));
                }
            CU(cudaMalloc(&h->d_cnt2, cnt2.size() * sizeof(u64)));
            CU(cudaMemcpy(h->d_cnt2, cnt2.data(), cnt2.size() * sizeof(u64), cudaMemcpyHostToDevice));
        }
        CU(cudaMemcpy(h->d_slotc, slotc.data(), slotc.size() * sizeof(u64), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_tri, tri.data(), tri.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        CU(cudaMemsetAsync(h->d_a, 0, nc * h->stride * sizeof(uint16_t), h->stream));
        CU(cudaMemsetAsync(h->d_best_a, 0, nc * h->stride * sizeof(uint16_t), h->stream));
        es_reset_state_kernel<<<(int)((nc + 255) / 256), 256, 0, h->stream>>>(h->d_st, 0, (int)nc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    });
    if (rc != CS_OK) {
        fprintf(stderr, "cs_es_create failed: %s\n", h->err.c_str());
        es_free(h);
        delete h;
        return rc;
    }
    *out = h;
    return CS_OK;
}

extern "C" int32_t cs_es_create(const cs_es_config* cfg, const int64_t* employee_ids,
                                const int64_t* hol_emp, const int64_t* hol_day, uint64_t n_hol,
                                cs_es_handle** out) {
    return cs_es_create_ex(cfg, employee_ids, hol_emp, hol_day, n_hol, 1u, nullptr, out);
}

extern "C" int32_t cs_es_destroy(cs_es_handle* h) {
    if (!h) return CS_ERR_INVALID_ARG;
    es_free(h);
    delete h;
    return CS_OK;
}

extern "C" const char* cs_es_last_error(const cs_es_handle* h) { return h ? h->err.c_str() : ""; }

extern "C" int32_t cs_es_set_stream(cs_es_handle* h, void* s) {
    return guarded(h, [&] {
        CU(cudaStreamSynchronize(h->stream));
        if (h->own_stream) {
            CU(cudaStreamDestroy(h->stream));
            h->own_stream = false;
        }
        h->stream = (cudaStream_t)s;
    });
}

extern "C" int32_t cs_es_init_random(cs_es_handle* h) {
    return guarded(h, [&] {
        const int nc = (int)h->cfg.n_chains;
        es_init_kernel<<<(nc + 63) / 64, 64, 0, h->stream>>>(h->d_a, h->d_st, h->stride, h->K.E, nc,
                                                              h->cfg.seed, h->cfg.chain_offset);
        CU(cudaGetLastError());
        es_rescore(h, 0, nc);
        es_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
    });
}

extern "C" int32_t cs_es_set_chains(cs_es_handle* h, uint32_t first, uint32_t count, const int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        es_check_range(h, first, count);
        es_upload(h, first, count, rows);
        es_reset_state_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_st, (int)first, (int)count);
        CU(cudaGetLastError());
        if (!h->scored && !(first == 0 && count == h->cfg.n_chains))
            es_rescore(h, 0, (int)h->cfg.n_chains);
        else
            es_rescore(h, (int)first, (int)count);
        es_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
    });
}

extern "C" int32_t cs_es_set_chains_async(cs_es_handle* h, uint32_t first, uint32_t count, const int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        es_check_range(h, first, count);
        if (h->async_pending) throw StateFail{"an upload is already pending: call cs_es_commit_chains first"};
        const size_t need = (size_t)count * h->stride;
        if (!h->copy_stream) {
            CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&h->ev_upload, cudaEventDisableTiming));
        }
        if (h->stage_async_elems < need) {
            cudaFree(h->d_stage_async);
            h->d_stage_async = nullptr;
            h->stage_async_elems = 0;
            CU(cudaMalloc(&h->d_stage_async, need * sizeof(long long)));
            h->stage_async_elems = need;
        }
        // copy engine only: nothing on the handle's stream touches d_stage_async before the commit
        CU(cudaMemcpyAsync(h->d_stage_async, rows, need * sizeof(long long), cudaMemcpyHostToDevice, h->copy_stream));
        CU(cudaEventRecord(h->ev_upload, h->copy_stream));
        h->async_first = first;
        h->async_count = count;
        h->async_pending = true;
    });
}

extern "C" int32_t cs_es_commit_chains(cs_es_handle* h) {
    return guarded(h, [&] {
        if (!h->async_pending) throw StateFail{"no pending upload: call cs_es_set_chains_async first"};
        const uint32_t first = h->async_first, count = h->async_count;
        h->async_pending = false;
        const size_t total = (size_t)count * h->stride;
        CU(cudaStreamWaitEvent(h->stream, h->ev_upload, 0));
        CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
        es_ids_to_index_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, h->stream>>>(
            h->d_stage_async, h->d_a + (size_t)first * h->stride, total, h->d_ids, (int)h->ids.size(), h->d_bad);
        CU(cudaGetLastError());
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        es_reset_state_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_st, (int)first, (int)count);
        CU(cudaGetLastError());
        if (!h->scored && !(first == 0 && count == h->cfg.n_chains)) es_rescore(h, 0, (int)h->cfg.n_chains);
        else es_rescore(h, (int)first, (int)count);
        es_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
        REQUIRE(!bad, "solution names an employee id that is not in the employee table (stored as the first employee)");
    });
}

extern "C" int32_t cs_es_set_window(cs_es_handle* h, uint64_t window_size) {
    return guarded(h, [&] {
        REQUIRE(window_size >= 1, "window_size must be >= 1");
        h->window = window_size;
    });
}

extern "C" int32_t cs_es_get_chains(cs_es_handle* h, uint32_t first, uint32_t count, int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        es_check_range(h, first, count);
        es_download(h, h->d_a, first, count, rows);
    });
}

extern "C" int32_t cs_es_get_best_chains(cs_es_handle* h, uint32_t first, uint32_t count, int64_t* rows,
                                         int64_t* best_hard, int64_t* best_soft) {
    return guarded(h, [&] {
        es_check_range(h, first, count);
        if (rows) es_download(h, h->d_best_a, first, count, rows);
        if (best_hard || best_soft) {
            auto st = es_states(h, first, count);
            for (uint32_t k = 0; k < count; ++k) {
                if (best_hard) best_hard[k] = st[k].best_hard;
                if (best_soft) best_soft[k] = st[k].best_soft;
            }
        }
    });
}

extern "C" int32_t cs_es_get_scores(cs_es_handle* h, int64_t* hard, int64_t* soft) {
    return guarded(h, [&] {
        REQUIRE(hard && soft, "hard/soft is NULL");
        auto st = es_states(h, 0, h->cfg.n_chains);
        for (size_t k = 0; k < h->cfg.n_chains; ++k) {
            hard[k] = st[k].hard;
            soft[k] = st[k].soft;
        }
    });
}

extern "C" int32_t cs_es_get_status(cs_es_handle* h, uint32_t* status) {
    return guarded(h, [&] {
        REQUIRE(status, "status is NULL");
        auto st = es_states(h, 0, h->cfg.n_chains);
        for (size_t k = 0; k < h->cfg.n_chains; ++k) status[k] = st[k].status;
    });
}

namespace {
void es_score_full10(cs_es_handle* h, uint32_t chain, long long t[10]) {
    es_check_range(h, chain, 1);
    long long* d_out = nullptr;
    CU(cudaMalloc(&d_out, 10 * sizeof(long long)));
    try {
        es_full_score_kernel<<<1, 32, 0, h->stream>>>(h->d_a + (size_t)chain * h->stride, h->stride, 1, h->D, h->S,
                                                      h->K.start_wd, h->d_hol, h->d_unsk, h->W, d_out);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(t, d_out, 10 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } catch (...) {
        cudaFree(d_out);
        throw;
    }
    cudaFree(d_out);
}
}  // namespace

extern "C" int32_t cs_es_score_full(cs_es_handle* h, uint32_t chain, int64_t* hard, int64_t* soft,
                                    int64_t terms[8]) {
    return guarded(h, [&] {
        long long t[10];
        es_score_full10(h, chain, t);
        if (hard) *hard = t[0] + t[1] + t[2] + t[3] + t[8] + t[9];
        if (soft) *soft = t[4] + t[5] + t[6] + t[7];
        if (terms)
            for (int k = 0; k < 8; ++k) terms[k] = t[k];
    });
}

extern "C" int32_t cs_es_score_full_ex(cs_es_handle* h, uint32_t chain, int64_t* hard, int64_t* soft,
                                       int64_t terms[10]) {
    return guarded(h, [&] {
        long long t[10];
        es_score_full10(h, chain, t);
        if (hard) *hard = t[0] + t[1] + t[2] + t[3] + t[8] + t[9];
        if (soft) *soft = t[4] + t[5] + t[6] + t[7];
        if (terms)
            for (int k = 0; k < 10; ++k) terms[k] = t[k];
    });
}

extern "C" int32_t cs_es_get_dims(cs_es_handle* h, uint32_t* n_days, uint32_t* shifts_per_day, uint32_t* n_slots) {
    return guarded(h, [&] {
        if (n_days) *n_days = (uint32_t)h->D;
        if (shifts_per_day) *shifts_per_day = (uint32_t)h->S;
        if (n_slots) *n_slots = (uint32_t)h->stride;
    });
}

extern "C" int32_t cs_es_eval_moves(cs_es_handle* h, uint32_t chain, const cs_es_move* moves,
                                    uint64_t n_moves, int64_t* dhard, int64_t* dsoft) {
    return guarded(h, [&] {
        es_check_range(h, chain, 1);
        if (n_moves == 0) return;
        REQUIRE(moves && dhard && dsoft, "moves/dhard/dsoft is NULL");
        const uint32_t D = (uint32_t)h->T, E = h->cfg.n_employees;  // D: scored slots
        // split by kind (the kernel takes one kind per launch), keep positions
        std::vector<uint2> mv[2];
        std::vector<uint64_t> pos[2];
        for (uint64_t k = 0; k < n_moves; ++k) {
            const cs_es_move& m = moves[k];
            REQUIRE(m.kind <= CS_ES_SWAP, "unknown move kind");
            REQUIRE(m.a < D && (m.kind == CS_ES_CHANGE ? m.b < E : m.b < D), "move index out of range");
            mv[m.kind].push_back(make_uint2(m.a, m.b));
            pos[m.kind].push_back(k);
        }
        for (int kind = 0; kind < 2; ++kind) {
            const size_t cnt = mv[kind].size();
            if (!cnt) continue;
            uint2* d_m = nullptr;
            long long* d_o = nullptr;
            CU(cudaMalloc(&d_m, cnt * sizeof(uint2)));
            cudaError_t e = cudaMalloc(&d_o, 2 * cnt * sizeof(long long));
            if (e != cudaSuccess) {
                cudaFree(d_m);
                CU(e);
            }
            std::vector<long long> o(2 * cnt);
            try {
                CU(cudaMemcpyAsync(d_m, mv[kind].data(), cnt * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
                es_dispatch(h, [&](auto w, auto m) {
                    constexpr int W = decltype(w)::value;
                    constexpr bool MULTI = decltype(m)::value;
                    EsParamsT<W> p = es_params<W>(h, 0, (int)h->cfg.n_chains);
                    es_eval_kernel<W, MULTI><<<1, 256, h->smem, h->stream>>>(p, (int)chain, kind, d_m, cnt, d_o, d_o + cnt);
                });
                CU(cudaGetLastError());
                CU(cudaMemcpyAsync(o.data(), d_o, 2 * cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
            } catch (...) {
                cudaFree(d_m);
                cudaFree(d_o);
                throw;
            }
            cudaFree(d_m);
            cudaFree(d_o);
            for (size_t k = 0; k < cnt; ++k) {
                dhard[pos[kind][k]] = o[k];
                dsoft[pos[kind][k]] = o[cnt + k];
            }
        }
    });
}

extern "C" int32_t cs_es_enumerate(cs_es_handle* h, uint32_t chain, cs_es_move* moves, uint64_t cap,
                                   uint64_t* n_out) {
    return guarded(h, [&] {
        es_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        std::vector<uint16_t> a(h->stride);
        CU(cudaMemcpyAsync(a.data(), h->d_a + (size_t)chain * h->stride, h->stride * sizeof(uint16_t),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        const uint32_t D = (uint32_t)h->T, E = h->cfg.n_employees;  // D: scored slots
        uint64_t k = 0;
        for (uint32_t d = 0; d < D; ++d)
            for (uint32_t e = 0; e < E; ++e) {
                if (a[d] == e) continue;
                if (moves && k < cap) moves[k] = cs_es_move{CS_ES_CHANGE, d, e};
                ++k;
            }
        for (uint32_t d1 = 0; d1 < D; ++d1)
            for (uint32_t d2 = d1 + 1; d2 < D; ++d2) {
                if (a[d1] == a[d2]) continue;
                if (moves && k < cap) moves[k] = cs_es_move{CS_ES_SWAP, d1, d2};
                ++k;
            }
        *n_out = k;
    });
}

extern "C" int32_t cs_es_neighbourhood_deltas(cs_es_handle* h, uint32_t chain, int64_t* dhard,
                                              int64_t* dsoft, uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        es_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        if (!h->scored) throw StateFail{"no solution loaded"};
        const uint64_t D = (uint64_t)h->T, E = h->cfg.n_employees;  // D: scored slots
        const uint64_t cnt = D * E + D * (D - 1) / 2;
        *n_out = cnt;
        if (!dhard || !dsoft) return;
        REQUIRE(cap >= cnt, "delta buffers too small");
        long long* d_dump = nullptr;
        CU(cudaMalloc(&d_dump, 2 * cnt * sizeof(long long)));
        try {
            EsRun r;
            r.first = (int)chain;
            r.count = 1;
            r.max_steps = 1;
            r.dump_h = d_dump;
            r.dump_s = d_dump + cnt;
            CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
            es_launch_step(h, 1, r);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(dhard, d_dump, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(dsoft, d_dump + cnt, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } catch (...) {
            cudaFree(d_dump);
            throw;
        }
        cudaFree(d_dump);
    });
}

extern "C" int32_t cs_es_step(cs_es_handle* h, uint32_t n_steps, cs_es_step_stats* stats) {
    return guarded(h, [&] { es_run(h, 0, (int)h->cfg.n_chains, n_steps, 0, 0, stats); });
}

extern "C" int32_t cs_es_step_enqueue(cs_es_handle* h, uint32_t n_steps) {
    return guarded(h, [&] { es_run_enqueue(h, 0, (int)h->cfg.n_chains, n_steps, 0, 0); });
}

extern "C" int32_t cs_es_step_wait(cs_es_handle* h, cs_es_step_stats* stats) {
    return guarded(h, [&] { es_run_wait(h, stats); });
}

extern "C" int32_t cs_es_exchange_select(cs_es_handle* h, const void* d_key, void* d_elite_u16, uint32_t elite_len) {
    return guarded(h, [&] {
        REQUIRE(d_key && d_elite_u16, "d_key / d_elite_u16 is NULL");
        if (elite_len == 0) return;
        xchg_select_kernel<<<(elite_len + 1023) / 1024, 1024, 0, h->stream>>>(
            (const long long*)d_key, h->d_a, (size_t)h->stride, (unsigned)h->cfg.chain_offset, h->cfg.n_chains,
            (unsigned)h->stride, (uint16_t*)d_elite_u16, elite_len);
        CU(cudaGetLastError());
    });
}

extern "C" int32_t cs_es_local_search(cs_es_handle* h, uint64_t allow, uint64_t max_iterations,
                                      cs_es_step_stats* stats) {
    return guarded(h, [&] { es_run(h, 0, (int)h->cfg.n_chains, max_iterations, allow, 1, stats); });
}

extern "C" int32_t cs_es_local_search_one(cs_es_handle* h, const int64_t* start, uint64_t allow,
                                          uint64_t max_iterations, int64_t* best, int64_t* best_hard,
                                          int64_t* best_soft) {
    return guarded(h, [&] {
        REQUIRE(start, "start is NULL");
        es_upload(h, 0, 1, start);
        es_reset_state_kernel<<<1, 32, 0, h->stream>>>(h->d_st, 0, 1);
        CU(cudaGetLastError());
        es_rescore(h, 0, h->scored ? 1 : (int)h->cfg.n_chains);
        h->scored = true;
        es_run(h, 0, 1, max_iterations, allow, 1, nullptr);
        if (best) es_download(h, h->d_best_a, 0, 1, best);
        auto st = es_states(h, 0, 1);
        if (best_hard) *best_hard = st[0].best_hard;
        if (best_soft) *best_soft = st[0].best_soft;
    });
}

extern "C" int32_t cs_es_get_trace(cs_es_handle* h, uint32_t chain, cs_es_move* moves, int64_t* hard_after,
                                   int64_t* soft_after, uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        es_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        auto st = es_states(h, chain, 1);
        *n_out = st[0].steps;
        uint64_t k = st[0].steps;
        if (k > h->cfg.trace_capacity) k = h->cfg.trace_capacity;
        if (k > cap) k = cap;
        if (k == 0) return;
        std::vector<EsTraceEntry> t(k);
        CU(cudaMemcpyAsync(t.data(), h->d_trace + (size_t)chain * h->cfg.trace_capacity,
                           k * sizeof(EsTraceEntry), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (uint64_t q = 0; q < k; ++q) {
            if (moves) moves[q] = cs_es_move{t[q].kind, t[q].x, t[q].y};
            if (hard_after) hard_after[q] = t[q].hard_after;
            if (soft_after) soft_after[q] = t[q].soft_after;
        }
    });
}

extern "C" int32_t cs_es_best(cs_es_handle* h, int64_t* rows, int64_t* hard, int64_t* soft, uint32_t* chain) {
    return guarded(h, [&] {
        if (!h->scored) throw StateFail{"no solution loaded"};
        es_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        if (hard) *hard = h->h_stats->best_hard;
        if (soft) *soft = h->h_stats->best_soft;
        if (chain) *chain = h->h_stats->best_chain;
        if (rows) es_download(h, h->d_a, h->h_stats->best_chain, 1, rows);
    });
}

extern "C" int32_t cs_es_best_key_device_ptr(cs_es_handle* h, void** dptr) {
    return guarded(h, [&] {
        REQUIRE(dptr, "dptr is NULL");
        *dptr = (void*)&h->d_stats->best_key;
    });
}

extern "C" int32_t cs_es_chain_device_ptr(cs_es_handle* h, uint32_t chain, void** dptr, uint32_t* n_slots) {
    return guarded(h, [&] {
        es_check_range(h, chain, 1);
        REQUIRE(dptr, "dptr is NULL");
        *dptr = (void*)(h->d_a + (size_t)chain * h->stride);
        if (n_slots) *n_slots = (uint32_t)h->stride;
    });
}

// ------------------------------------------------------------------ scheduling ILS shell
namespace {

__global__ void es_gather_keys_kernel(const EsChainState* st, long long* best_key, long long* cur_key,
                                      int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (best_key) best_key[k] = (st[k].best_hard << 32) | st[k].best_soft;
    if (cur_key) cur_key[k] = (st[k].hard << 32) | st[k].soft;
}

IlsParams es_ils_params(cs_es_handle* h) {
    IlsParams p = h->ils.params(h->d_a, h->d_best_a, h->cfg.seed, h->cfg.chain_offset);
    p.value_range = h->K.E;
    p.restart_is_perm = 0;
    p.do_nothing_first = 1;  // employee-scheduling lib.rs:574-577: DoNothing listed first
    p.k_before_shuffle = 1;  // lib.rs:600-605: subset size, then the shuffle
    return p;
}

}  // namespace

extern "C" int32_t cs_es_ils_init(cs_es_handle* h, uint32_t cap, uint32_t log_cap) {
    return guarded(h, [&] {
        REQUIRE(cap >= 1 && cap <= ILS_MAX_CAP, "best_solutions_capacity must be 1..64");
        if (!h->scored) throw StateFail{"no solution loaded"};
        const int nc = (int)h->cfg.n_chains;
        h->ils.alloc(nc, h->stride, h->stride, (int)cap, (int)log_cap);
        ils_reset_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->ils.d_st, h->ils.d_cur_key,
                                                                    h->ils.d_skip, nc);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h->ils.d_cur, h->d_a, (size_t)nc * h->stride * sizeof(uint16_t),
                           cudaMemcpyDeviceToDevice, h->stream));
        es_gather_keys_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_st, nullptr, h->ils.d_cur_key, nc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    });
}

extern "C" int32_t cs_es_ils_run(cs_es_handle* h, uint32_t rounds, uint64_t ls_max_iterations,
                                 uint64_t allow, uint32_t stop_when_any_best, cs_ils_stats* stats) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_es_ils_init first"};
        const int nc = (int)h->cfg.n_chains;
        IlsParams ip = es_ils_params(h);
        EsRun lp;
        lp.first = 0;
        lp.count = nc;
        lp.max_steps = ls_max_iterations;
        lp.allow = allow;
        lp.ls_mode = 1;
        lp.skip = h->ils.d_skip;
        const int ls_grid = nc < h->grid_cap ? nc : h->grid_cap;
        const int ig = nc < 4096 ? nc : 4096;
        unsigned launches = 0, run = 0;
        CU(cudaMemsetAsync(h->d_totals, 0, 2 * sizeof(unsigned long long), h->stream));
        CU(cudaEventRecord(h->ev0, h->stream));
        for (uint32_t r = 0; r < rounds; ++r) {
            ils_perturb_kernel<<<ig, ILS_THREADS, h->ils.perturb_smem, h->stream>>>(ip);
            CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
            es_launch_step(h, ls_grid, lp);
            es_gather_keys_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_st, h->ils.d_neu_key, nullptr, nc);
            ils_accept_kernel<<<ig, ILS_THREADS, 0, h->stream>>>(ip);
            CU(cudaGetLastError());
            launches += 4;
            ++run;
            if (stop_when_any_best) {
                ils_summary_kernel<<<1, 1024, 0, h->stream>>>(ip, h->ils.d_sum);
                CU(cudaMemcpyAsync(h->ils.h_sum, h->ils.d_sum, sizeof(IlsSummary), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                ++launches;
                if (h->ils.h_sum->chains_done) break;
            }
        }
        ils_summary_kernel<<<1, 1024, 0, h->stream>>>(ip, h->ils.d_sum);
        CU(cudaMemcpyAsync(h->ils.h_sum, h->ils.d_sum, sizeof(IlsSummary), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->d_a, h->ils.d_cur, (size_t)nc * h->stride * sizeof(uint16_t),
                           cudaMemcpyDeviceToDevice, h->stream));
        es_rescore(h, 0, nc);
        es_refresh_stats(h);
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaMemcpyAsync(h->h_totals, h->d_totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (stats) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            stats->moves_scored = h->h_totals[0];
            stats->ls_steps = h->h_totals[1];
            stats->best_key = h->ils.h_sum->best_key;
            stats->best_chain = h->ils.h_sum->best_chain;
            stats->chains_done = h->ils.h_sum->chains_done;
            stats->rounds_run = run;
            stats->device_ms = ms;
            stats->kernel_launches = launches + 3;
        }
    });
}

extern "C" int32_t cs_es_ils_get_best(cs_es_handle* h, uint32_t chain, int64_t* rows, int64_t* hard,
                                      int64_t* soft) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_es_ils_init first"};
        es_check_range(h, chain, 1);
        IlsChainState st;
        CU(cudaMemcpyAsync(&st, h->ils.d_st + chain, sizeof st, cudaMemcpyDeviceToHost, h->stream));
        unsigned char slot = 0;
        CU(cudaMemcpyAsync(&slot, h->ils.d_order + (size_t)chain * h->ils.cap, 1, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (!st.size) throw StateFail{"no round has run yet (the reference unwrap()s None here)"};
        long long key = 0;
        CU(cudaMemcpyAsync(&key, h->ils.d_bset_key + (size_t)chain * h->ils.cap + slot, sizeof key,
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (hard) *hard = key >> 32;
        if (soft) *soft = key & 0xffffffffll;
        if (rows) es_download(h, h->ils.d_bset + ((size_t)chain * h->ils.cap + slot) * h->stride, 0, 1, rows);
    });
}

extern "C" int32_t cs_es_ils_get_log(cs_es_handle* h, uint32_t chain, int64_t* new_key, uint32_t* choice,
                                     uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_es_ils_init first"};
        es_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        IlsChainState st;
        CU(cudaMemcpyAsync(&st, h->ils.d_st + chain, sizeof st, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        *n_out = st.log_len;
        uint64_t k = st.log_len;
        if (k > (uint64_t)h->ils.log_cap) k = h->ils.log_cap;
        if (k > cap) k = cap;
        if (!k) return;
        std::vector<IlsLogEntry> log(k);
        CU(cudaMemcpyAsync(log.data(), h->ils.d_log + (size_t)chain * h->ils.log_cap, k * sizeof(IlsLogEntry),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (uint64_t q = 0; q < k; ++q) {
            if (new_key) new_key[q] = log[q].new_key;
            if (choice) choice[q] = log[q].choice;
        }
    });
}
