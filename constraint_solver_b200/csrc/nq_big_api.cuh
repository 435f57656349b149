// Host side of the big-board (global-memory, partitionable) n-queens path; included by
// cs_api.cu after cs_nq_handle is defined.
#pragma once
#include <cstdlib>

namespace {

__global__ void nqb_sync_state_kernel(NqBig b, NqChainState* st, int reset_steps) {
    const long long s = *b.score;
    st->score = s;
    st->best_score = s;
    st->is_perm = (*b.ident_pairs == 0);
    st->status = s == 0 ? 1u : 0u;
    if (reset_steps) {
        st->steps = 0;
        st->moves_scored = 0;
    }
    *b.key = NQB_KEY_NONE;
}

int nqb_grid(cs_nq_handle* h, long long work) {
    long long g = (work + 255) / 256;
    const long long cap = (long long)h->sm_count * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// triangular-balanced column range of partition `part` of `parts` over columns 0..n-2
void nqb_set_range(cs_nq_handle* h) {
    const long long n = h->cfg.n;
    auto cum = [&](long long i) { return i * (n - 1) - i * (i - 1) / 2; };  // pairs of columns < i
    const long long total = n * (n - 1) / 2;
    auto bound = [&](uint32_t k) -> long long {
        if (k == 0) return 0;
        if (k >= h->parts) return n - 1;
        const long long target = total / h->parts * k;
        long long lo = 0, hi = n - 1;
        while (lo < hi) {
            const long long mid = (lo + hi) / 2;
            if (cum(mid) >= target) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    h->big.i_begin = (int)bound(h->part);
    h->big.i_end = (int)bound(h->part + 1);
}

void nqb_alloc(cs_nq_handle* h) {
    NqBig& b = h->big;
    b.n = (int)h->cfg.n;
    b.n_pad = h->n_pad;
    b.ld = nqb_ld(h->n_pad);
    CU(cudaMalloc(&b.rows, (size_t)(b.n_pad + 128) * 4));  // zero slack: the packed scan reads whole chunks
    CU(cudaMalloc(&b.c, (size_t)b.n_pad * 4));
    b.ldb = nqb_ldb(h->n_pad);
    CU(cudaMalloc(&b.Q, (size_t)2 * NQBP_COPIES * b.ldb));
    CU(cudaMalloc(&b.cb, (size_t)b.n_pad + 128));
    CU(cudaMemset(b.cb, 0, (size_t)b.n_pad + 128));
    b.use_packed = 0;
    b.seg = NQBP_SEG;
    if (const char* e = std::getenv("CS_NQB_SEG")) {  // tuning / test knob: work-unit length in chunks
        const int v = std::atoi(e);
        if (v >= 1) b.seg = v;
    }
    CU(cudaMalloc(&b.R, (size_t)b.n_pad * 4));
    CU(cudaMalloc(&b.D1, (size_t)b.ld * 4));
    CU(cudaMalloc(&b.D2, (size_t)b.ld * 4));
    CU(cudaMalloc(&h->d_best_rows32, (size_t)b.n_pad * 4));
    // one block of scalars: score, ident_pairs, key1, scored, key | tile_counter, jmin, maxcount
    long long* sc = nullptr;
    CU(cudaMalloc(&sc, 8 * sizeof(long long)));
    CU(cudaMemset(sc, 0, 8 * sizeof(long long)));
    b.score = sc;
    b.ident_pairs = (unsigned long long*)(sc + 1);
    b.key1 = (unsigned long long*)(sc + 2);
    b.scored = (unsigned long long*)(sc + 3);
    b.key = sc + 4;
    b.tile_counter = (unsigned int*)(sc + 5);
    b.jmin = (unsigned int*)(sc + 6);
    b.maxcount = (unsigned int*)(sc + 7);
    b.dump = nullptr;
    CU(cudaMalloc(&h->d_bstep, sizeof(NqBigStep)));
    CU(cudaMallocHost(&h->h_bstep, sizeof(NqBigStep)));
    CU(cudaMallocHost(&h->h_key, sizeof(long long)));
    CU(cudaMallocHost(&h->h_scored, sizeof(unsigned long long)));
    CU(cudaMemset(b.rows, 0, (size_t)(b.n_pad + 128) * 4));
    CU(cudaMemset(h->d_best_rows32, 0, (size_t)b.n_pad * 4));
    nqb_set_range(h);
}

// rebuild counters + score from rows; refresh chain state
void nqb_rebuild(cs_nq_handle* h, bool reset_steps) {
    NqBig& b = h->big;
    nqb_zero_kernel<<<nqb_grid(h, b.ld), 256, 0, h->stream>>>(b);
    nqb_count_kernel<<<nqb_grid(h, b.n), 256, 0, h->stream>>>(b);
    nqb_score_kernel<<<nqb_grid(h, 2ll * b.n), 256, 0, h->stream>>>(b);
    nqb_sync_state_kernel<<<1, 1, 0, h->stream>>>(b, h->d_st, reset_steps ? 1 : 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->d_best_rows32, b.rows, (size_t)b.n_pad * 4, cudaMemcpyDeviceToDevice,
                       h->stream));
    nq_refresh_stats(h);
    CU(cudaStreamSynchronize(h->stream));
}

bool nqb_is_perm(cs_nq_handle* h) {
    NqChainState st;
    CU(cudaMemcpyAsync(&st, h->d_st, sizeof st, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return st.is_perm != 0;
}

// Keep the byte-copy tables (2 x 16 copies, 64 MB at n = 10^6) resident in L2: a persisting access-policy
// window over them on the handle's stream.  Without it the tables sit at the edge of what the two-die L2
// holds and a step's DRAM re-reads vary 6-19 GB (225-246 ms) from run to run.  CS_NQB_L2_PERSIST=0 disables.
void nqb_set_l2_policy(cs_nq_handle* h) {
    if (h->l2_policy_set && h->l2_policy_stream == h->stream) return;
    h->l2_policy_set = true;
    h->l2_policy_stream = h->stream;
    if (const char* e = std::getenv("CS_NQB_L2_PERSIST"))
        if (std::atoi(e) == 0) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 ||
        prop.accessPolicyMaxWindowSize <= 0) {
        cudaGetLastError();
        return;
    }
    const size_t bytes = (size_t)2 * NQBP_COPIES * h->big.ldb;
    const size_t persist = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist);
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = (void*)h->big.Q;
    attr.accessPolicyWindow.num_bytes = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
    attr.accessPolicyWindow.hitRatio = (float)((double)persist / (double)attr.accessPolicyWindow.num_bytes);
    if (attr.accessPolicyWindow.hitRatio > 1.0f) attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();  // a hint: never an error for the caller
}

// enqueue: c[], scan of this partition, row re-scan, packed key.  No host sync.
void nqb_enqueue_scan(cs_nq_handle* h, bool perm, long long* dump) {
    nqb_set_l2_policy(h);
    NqBig b = h->big;
    b.dump = dump;
    b.use_packed = (perm && b.n >= NQBP_MIN_N && !(h->cfg.flags & CS_NQ_FLAG_SCALAR)) ? 1 : 0;
    nqb_compute_c_kernel<<<nqb_grid(h, b.n), 256, 0, h->stream>>>(b);
    if (b.use_packed) {  // byte copies + largest line count, then the packed scan (no-op if a line is too long)
        nqb_pack_kernel<<<nqb_grid(h, b.ldb / 16), 256, 0, h->stream>>>(b);
        if (dump) nqb_scan_packed_kernel<true><<<h->sm_count * (16 / NQBP_WARPS), 32 * NQBP_WARPS, 0, h->stream>>>(b);
        else nqb_scan_packed_kernel<false><<<h->sm_count * (16 / NQBP_WARPS), 32 * NQBP_WARPS, 0, h->stream>>>(b);
        // the fallback below returns at once unless the packed scan declined; it needs a fresh tile counter
        CU(cudaMemsetAsync(b.tile_counter, 0, sizeof(unsigned int), h->stream));
    }
    const int grid = h->sm_count * 4;
    if (perm) {
        if (dump) nqb_scan_kernel<true, true><<<grid, 256, 0, h->stream>>>(b);
        else nqb_scan_kernel<true, false><<<grid, 256, 0, h->stream>>>(b);
    } else {
        if (dump) nqb_scan_kernel<false, true><<<grid, 256, 0, h->stream>>>(b);
        else nqb_scan_kernel<false, false><<<grid, 256, 0, h->stream>>>(b);
    }
    nqb_rowscan_kernel<<<nqb_grid(h, b.n), 256, 0, h->stream>>>(b);
    nqb_pack_key_kernel<<<1, 1, 0, h->stream>>>(b);
    CU(cudaGetLastError());
}

void nqb_enqueue_apply(cs_nq_handle* h) {
    nqb_apply_kernel<<<1, 1, 0, h->stream>>>(h->big, h->d_st, h->d_trace,
                                             (int)h->cfg.trace_capacity, h->d_bstep);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->h_bstep, h->d_bstep, sizeof(NqBigStep), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h->h_scored, h->big.scored, sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, h->stream));
}

long long nqb_host_score(cs_nq_handle* h) {
    long long s = 0;
    CU(cudaMemcpyAsync(&s, h->big.score, sizeof s, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return s;
}

// host-driven step loop (single partition): LocalSearch::execute bookkeeping when ls_mode
void nqb_run(cs_nq_handle* h, unsigned long long max_steps, unsigned long long allow, int ls_mode,
             cs_step_stats* stats) {
    if (!h->scored) throw StateFail{"no solution loaded"};
    REQUIRE(h->parts == 1, "this handle scans a partition only: drive it with cs_nq_part_scan / "
                           "reduce / cs_nq_part_apply");
    const bool perm = nqb_is_perm(h);
    long long score = nqb_host_score(h);
    long long best_score = score;
    unsigned long long moves = 0, steps = 0, no_improve = 0;
    unsigned int status = 0, launches = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    if (ls_mode)
        CU(cudaMemcpyAsync(h->d_best_rows32, h->big.rows, (size_t)h->big.n_pad * 4,
                           cudaMemcpyDeviceToDevice, h->stream));
    for (unsigned long long it = 0; it < max_steps; ++it) {
        if (score == 0) {
            status = 1;
            break;
        }
        nqb_enqueue_scan(h, perm, nullptr);
        launches += 4;
        CU(cudaMemcpyAsync(h->h_key, h->big.key, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(h->h_scored, h->big.scored, sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        moves += *h->h_scored;
        const long long key = *h->h_key;
        if (key == NQB_KEY_NONE) {
            status = 3;
            break;
        }
        const long long v = (key >> 40) - NQB_BIAS;
        const bool improved = v < 0;
        if (!improved) {
            ++no_improve;
            if (allow && no_improve >= allow) {
                status = 2;
                break;
            }
        } else {
            no_improve = 0;
        }
        nqb_enqueue_apply(h);
        launches += 1;
        CU(cudaStreamSynchronize(h->stream));
        score = h->h_bstep->score_after;
        ++steps;
        if (improved) {
            best_score = score;
            CU(cudaMemcpyAsync(h->d_best_rows32, h->big.rows, (size_t)h->big.n_pad * 4,
                               cudaMemcpyDeviceToDevice, h->stream));
        }
    }
    // publish status / best score into the chain state
    NqChainState st;
    CU(cudaMemcpyAsync(&st, h->d_st, sizeof st, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    st.status = status;
    st.moves_scored += moves;
    if (ls_mode || best_score < st.best_score || st.best_score < 0) st.best_score = best_score;
    CU(cudaMemcpyAsync(h->d_st, &st, sizeof st, cudaMemcpyHostToDevice, h->stream));
    nq_refresh_stats(h);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (stats) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        stats->moves_scored = moves;
        stats->steps_accepted = steps;
        stats->best_score = score;
        stats->best_chain = 0;
        stats->chains_at_best = score == 0;
        stats->device_ms = ms;
        stats->kernel_launches = launches + 1;
    }
}

bool nqb_upload(cs_nq_handle* h, const int64_t* rows) {  // true: some row was out of range (stored as 0)
    const int n = (int)h->cfg.n;
    CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    CU(cudaMemcpyAsync(h->d_stage, rows, (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    nqb_pack_rows32_kernel<<<nqb_grid(h, h->n_pad), 256, 0, h->stream>>>(h->d_stage, h->big.rows, n,
                                                                       h->n_pad, h->d_bad);
    CU(cudaGetLastError());
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return bad != 0;
}

void nqb_download(cs_nq_handle* h, const unsigned int* src, int64_t* rows) {
    const int n = (int)h->cfg.n;
    nqb_unpack_rows32_kernel<<<nqb_grid(h, n), 256, 0, h->stream>>>(src, h->d_stage, n);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rows, h->d_stage, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
}

}  // namespace

extern "C" int32_t cs_nq_set_partition(cs_nq_handle* h, uint32_t part, uint32_t parts) {
    return guarded(h, [&] {
        REQUIRE(h->is_big, "partitioning needs the big-board path (n > CS_NQ_MAX_N_SMEM or CS_NQ_FLAG_GLOBAL)");
        REQUIRE(parts >= 1 && part < parts, "part must be < parts");
        h->part = part;
        h->parts = parts;
        nqb_set_range(h);
    });
}

extern "C" int32_t cs_nq_part_scan(cs_nq_handle* h) {
    return guarded(h, [&] {
        REQUIRE(h->is_big, "cs_nq_part_scan needs the big-board path");
        if (!h->scored) throw StateFail{"no solution loaded"};
        CU(cudaEventRecord(h->ev0, h->stream));
        nqb_enqueue_scan(h, nqb_is_perm(h), nullptr);
    });
}

extern "C" int32_t cs_nq_part_key_device_ptr(cs_nq_handle* h, void** dptr) {
    return guarded(h, [&] {
        REQUIRE(h->is_big && dptr, "needs the big-board path and a non-NULL dptr");
        *dptr = (void*)h->big.key;
    });
}

extern "C" int32_t cs_nq_part_apply(cs_nq_handle* h, cs_step_stats* stats) {
    return guarded(h, [&] {
        REQUIRE(h->is_big, "cs_nq_part_apply needs the big-board path");
        nqb_enqueue_apply(h);
        nq_refresh_stats(h);
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (stats) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            stats->moves_scored = *h->h_scored;
            stats->steps_accepted = h->h_bstep->applied;
            stats->best_score = h->h_bstep->score_after;
            stats->best_chain = 0;
            stats->chains_at_best = h->h_bstep->score_after == 0;
            stats->device_ms = ms;
            stats->kernel_launches = 6;
        }
    });
}
