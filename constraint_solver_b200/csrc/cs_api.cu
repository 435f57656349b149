// C ABI (include/cs_b200.h) over the sm_100a chain kernels.  No CPU fallback: every entry
// point that computes needs a CUDA device and says so loudly when there is none.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/cs_b200.h"
#include "nq_kernels.cuh"
#include "nq_big.cuh"
#include "nq_packed.cuh"
#include "ils_kernels.cuh"
#include "philox.cuh"
#include "microbench.cuh"

using namespace csb;

// ------------------------------------------------------------------ helpers
namespace {

struct CudaFail {
    cudaError_t e;
    const char* what;
    int line;
};

#define CU(expr)                                            \
    do {                                                    \
        cudaError_t _e = (expr);                            \
        if (_e != cudaSuccess) throw CudaFail{_e, #expr, __LINE__}; \
    } while (0)

struct ArgFail {
    std::string msg;
};
#define REQUIRE(cond, msg)               \
    do {                                 \
        if (!(cond)) throw ArgFail{msg}; \
    } while (0)

struct StateFail {
    std::string msg;
};
struct Unsupported {
    std::string msg;
};

template <typename H, typename F>
int32_t guarded(H* h, F&& f) {
    if (!h) return CS_ERR_INVALID_ARG;
    try {
        h->err.clear();
        CU(cudaSetDevice(h->device));
        f();
        return CS_OK;
    } catch (const CudaFail& c) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at cs_api.cu:%d: %s", (int)c.e,
                 cudaGetErrorString(c.e), c.line, c.what);
        h->err = buf;
        cudaGetLastError();
        return c.e == cudaErrorMemoryAllocation ? CS_ERR_OOM : CS_ERR_CUDA;
    } catch (const ArgFail& a) {
        h->err = a.msg;
        return CS_ERR_INVALID_ARG;
    } catch (const StateFail& s) {
        h->err = s.msg;
        return CS_ERR_STATE;
    } catch (const Unsupported& u) {
        h->err = u.msg;
        return CS_ERR_UNSUPPORTED;
    } catch (const std::bad_alloc&) {
        h->err = "host out of memory";
        return CS_ERR_OOM;
    } catch (...) {
        h->err = "unknown internal error";
        return CS_ERR_CUDA;
    }
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// The dynamic shared-memory opt-in is a per-function, per-device attribute (not per handle): always
// raise it to the device's opt-in maximum, so handles of different sizes never lower each other's
// limit and concurrent creates write the same value.
template <typename K>
void allow_max_smem(K kernel, int optin_bytes) {
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, kernel));  // static + dynamic must fit the opt-in limit
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            optin_bytes - (int)fa.sharedSizeBytes));
}
template <typename K>
void allow_max_smem(K kernel, const cudaDeviceProp& prop) {
    allow_max_smem(kernel, (int)prop.sharedMemPerBlockOptin);
}

}  // namespace

#include "ils_api.cuh"

// ------------------------------------------------------------------ library-wide
extern "C" int32_t cs_abi_version(void) { return CS_ABI_VERSION; }

extern "C" int32_t cs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void cs_philox4x32_10(uint64_t seed, uint32_t chain, uint32_t purpose,
                                 uint64_t counter, uint32_t out[4]) {
    const Philox4 r = philox_stream(seed, chain, purpose, counter);
    for (int k = 0; k < 4; ++k) out[k] = r.v[k];
}

extern "C" const char* cs_status_string(int32_t s) {
    switch (s) {
        case CS_OK: return "ok";
        case CS_ERR_INVALID_ARG: return "invalid argument";
        case CS_ERR_CUDA: return "CUDA error";
        case CS_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case CS_ERR_OOM: return "out of memory";
        case CS_ERR_STATE: return "invalid handle state";
        case CS_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}

// On-chip bandwidth micro-benchmarks (microbench.cuh): measured GB/s of the conflict-free
// shared-memory stream (all SMs) and of the L2 -> SM stream; bench.py's roofline denominators.
extern "C" int32_t cs_microbench(int32_t device, uint32_t which, double* gbs, double* sm_mhz) {
    if (!gbs || which > CS_MICROBENCH_L2_READ) return CS_ERR_INVALID_ARG;
    int ndev = cs_device_count();
    if (ndev <= 0) return CS_ERR_NO_DEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return CS_ERR_NO_DEVICE;
    if (device >= ndev) return CS_ERR_INVALID_ARG;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    unsigned int* sink = nullptr;
    uint4* buf = nullptr;
    int32_t rc = CS_OK;
    int prev_device = -1;
    cudaGetDevice(&prev_device);  // restored below: the caller's current device is not ours to change
    try {
        CU(cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaMalloc(&sink, sizeof(unsigned int)));
        const int sms = prop.multiProcessorCount;
        double bytes = 0.0;
        float best_ms = 1e30f;
        if (which == CS_MICROBENCH_L2_READ) {
            const size_t n_vec = (size_t)(32u << 20) / 16;  // 32 MB: L2-resident on B200 (126 MB), far beyond L1
            CU(cudaMalloc(&buf, n_vec * 16));
            CU(cudaMemsetAsync(buf, 1, n_vec * 16, st));
            const int passes = 24;
            for (int rep = 0; rep < 4; ++rep) {  // rep 0 also warms L2
                CU(cudaEventRecord(e0, st));
                mb_l2_kernel<<<sms * 4, 512, 0, st>>>(buf, n_vec, passes, sink);
                CU(cudaEventRecord(e1, st));
                CU(cudaStreamSynchronize(st));
                CU(cudaGetLastError());
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best_ms) best_ms = ms;
            }
            bytes = (double)n_vec * 16.0 * passes;
        } else {
            const int iters = 8192;
            for (int rep = 0; rep < 4; ++rep) {
                CU(cudaEventRecord(e0, st));
                if (which == CS_MICROBENCH_SMEM_LDS128) mb_smem_kernel<true><<<sms * 2, 1024, 0, st>>>(iters, sink);
                else mb_smem_kernel<false><<<sms * 2, 1024, 0, st>>>(iters, sink);
                CU(cudaEventRecord(e1, st));
                CU(cudaStreamSynchronize(st));
                CU(cudaGetLastError());
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best_ms) best_ms = ms;
            }
            bytes = (double)sms * 2 * 1024 * (double)iters * 8 * (which == CS_MICROBENCH_SMEM_LDS128 ? 16.0 : 4.0);
        }
        *gbs = bytes / (best_ms * 1e-3) / 1e9;
        if (sm_mhz) *sm_mhz = prop.clockRate / 1e3;  // the device's rated boost clock (for the B/clk/SM reading)
    } catch (const CudaFail& c) {
        fprintf(stderr, "cs_microbench: CUDA error %d (%s) at cs_api.cu:%d\n", (int)c.e, cudaGetErrorString(c.e), c.line);
        cudaGetLastError();
        rc = CS_ERR_CUDA;
    } catch (...) {
        rc = CS_ERR_CUDA;
    }
    cudaFree(buf);
    cudaFree(sink);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (prev_device >= 0) cudaSetDevice(prev_device);
    return rc;
}

// ------------------------------------------------------------------ n-queens handle
struct cs_nq_handle {
    cs_nq_config cfg{};
    int device = 0;
    int n_pad = 0;
    int sm_count = 0;
    int threads = NQ_THREADS;
    size_t smem = 0;
    bool use_v2 = false;   // step kernel with the packed-window fast path
    size_t smem_v2 = 0;
    int grid_v2 = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint16_t* d_rows = nullptr;
    uint16_t* d_best_rows = nullptr;
    NqChainState* d_st = nullptr;
    NqTraceEntry* d_trace = nullptr;
    unsigned int* d_work = nullptr;
    unsigned long long* d_totals = nullptr;  // [2] moves, steps of the current call
    NqStats* d_stats = nullptr;
    NqStats* h_stats = nullptr;  // pinned
    unsigned long long* h_totals = nullptr;  // pinned [2]
    long long* d_stage = nullptr;
    size_t stage_elems = 0;
    // double-buffered input staging (cs_nq_set_chains_async / cs_nq_commit_chains)
    long long* d_stage_async = nullptr;
    size_t stage_async_elems = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_upload = nullptr;
    uint32_t async_first = 0, async_count = 0;
    bool async_pending = false;
    int* d_bad = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool scored = false;  // chain scores valid
    bool run_pending = false;  // a step was enqueued and not yet waited for
    // big-board (global-memory) mode: one instance, optional neighbourhood partition
    bool is_big = false;
    NqBig big{};
    cudaStream_t l2_policy_stream = nullptr;  // stream the persisting-L2 window of the byte-copy tables is set on
    bool l2_policy_set = false;
    unsigned int* d_best_rows32 = nullptr;
    NqBigStep* d_bstep = nullptr;
    NqBigStep* h_bstep = nullptr;       // pinned
    long long* h_key = nullptr;         // pinned
    unsigned long long* h_scored = nullptr;  // pinned
    uint32_t part = 0, parts = 1;
    unsigned long long ls_no_improve = 0;
    IlsHost ils;
    bool ref_mode = false;
    unsigned long long window = 0;
    unsigned long long* d_ls_rng = nullptr;  // [chains] LS-owned rng draw counters
    std::string err;
};

namespace {

constexpr int NQ_TI = 4;

NqParams nq_params(cs_nq_handle* h, int first, int count) {
    NqParams p{};
    p.n = (int)h->cfg.n;
    p.n_pad = h->n_pad;
    p.first_chain = first;
    p.n_chains = count;
    p.rows = h->d_rows;
    p.best_rows = h->d_best_rows;
    p.st = h->d_st;
    p.trace = h->d_trace;
    p.trace_cap = (int)h->cfg.trace_capacity;
    p.work_counter = h->d_work;
    p.totals = h->d_totals;
    p.max_steps = 0;
    p.allow_no_improve = 0;
    p.ls_mode = 0;
    p.kind = (int)h->cfg.neighbourhood;
    p.dump = nullptr;
    p.ref_mode = h->ref_mode ? 1 : 0;
    p.window = h->window;
    p.ls_rng_t = h->d_ls_rng;
    p.seed = h->cfg.seed;
    p.chain_offset = h->cfg.chain_offset;
    return p;
}

void nq_free(cs_nq_handle* h) {
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->ils.release();
    cudaFree(h->d_rows);
    cudaFree(h->d_best_rows);
    cudaFree(h->d_st);
    cudaFree(h->d_trace);
    cudaFree(h->d_work);
    cudaFree(h->d_totals);
    cudaFree(h->d_stats);
    cudaFree(h->d_stage);
    cudaFree(h->d_stage_async);
    if (h->ev_upload) cudaEventDestroy(h->ev_upload);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaFree(h->d_bad);
    cudaFree(h->d_ls_rng);
    if (h->is_big) {
        if (h->l2_policy_set && h->l2_policy_stream == h->stream && h->stream) {
            // drop the persisting window before the tables go away (a hint; errors are not the caller's problem)
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaCtxResetPersistingL2Cache();
            cudaGetLastError();
        }
        cudaFree(h->big.rows);
        cudaFree(h->big.Q);
        cudaFree(h->big.cb);
        cudaFree(h->big.c);
        cudaFree(h->big.R);
        cudaFree(h->big.D1);
        cudaFree(h->big.D2);
        cudaFree(h->big.score);
        cudaFree(h->d_best_rows32);
        cudaFree(h->d_bstep);
        if (h->h_bstep) cudaFreeHost(h->h_bstep);
        if (h->h_key) cudaFreeHost(h->h_key);
        if (h->h_scored) cudaFreeHost(h->h_scored);
    }
    if (h->h_stats) cudaFreeHost(h->h_stats);
    if (h->h_totals) cudaFreeHost(h->h_totals);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
}

void nq_check_range(cs_nq_handle* h, uint32_t first, uint32_t count) {
    REQUIRE(count > 0 && first < h->cfg.n_chains && count <= h->cfg.n_chains - first,
            "chain range outside [0, n_chains)");
}

// (re)compute score / best / permutation flag of a chain range after its rows changed
void nq_rescore(cs_nq_handle* h, int first, int count) {
    NqParams p = nq_params(h, first, count);
    const int grid = count < h->sm_count ? count : h->sm_count;
    nq_rescore_kernel<<<grid, h->threads, h->smem, h->stream>>>(p);
    CU(cudaGetLastError());
}

void nq_refresh_stats(cs_nq_handle* h) {
    nq_stats_kernel<<<1, 1024, 0, h->stream>>>(h->d_st, (int)h->cfg.n_chains,
                                                h->cfg.chain_offset, h->d_stats);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h->h_stats, h->d_stats, sizeof(NqStats), cudaMemcpyDeviceToHost,
                       h->stream));
}

void nq_launch_step(cs_nq_handle* h, NqParams& p, int count) {
    p.force_scalar = (h->cfg.flags & CS_NQ_FLAG_SCALAR) ? 1 : 0;
    if (h->use_v2) {
        const int grid = count < h->grid_v2 ? count : h->grid_v2;
        nq_step_kernel_v2<NQ_TI><<<grid, NQC_THREADS, h->smem_v2, h->stream>>>(p);
    } else {
        const int grid = count < h->sm_count ? count : h->sm_count;
        nq_step_kernel<NQ_TI><<<grid, h->threads, h->smem, h->stream>>>(p);
    }
    CU(cudaGetLastError());
}

void nq_run_enqueue(cs_nq_handle* h, int first, int count, unsigned long long max_steps,
                    unsigned long long allow, int ls_mode) {
    if (!h->scored) throw StateFail{"chains have no solution yet: call cs_nq_init_random or cs_nq_set_chains first"};
    NqParams p = nq_params(h, first, count);
    p.max_steps = max_steps;
    p.allow_no_improve = allow;
    p.ls_mode = ls_mode;
    CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
    CU(cudaMemsetAsync(h->d_totals, 0, 2 * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev0, h->stream));
    nq_launch_step(h, p, count);
    nq_refresh_stats(h);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaMemcpyAsync(h->h_totals, h->d_totals, 2 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, h->stream));
    h->run_pending = true;
}

void nq_run_wait(cs_nq_handle* h, cs_step_stats* stats) {
    if (!h->run_pending) throw StateFail{"no enqueued step to wait for"};
    CU(cudaStreamSynchronize(h->stream));
    h->run_pending = false;
    if (stats) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        stats->moves_scored = h->h_totals[0];
        stats->steps_accepted = h->h_totals[1];
        stats->best_score = h->h_stats->best_score;
        stats->best_chain = h->h_stats->best_chain;
        stats->chains_at_best = h->h_stats->chains_at_best;
        stats->device_ms = ms;
        stats->kernel_launches = 2;
    }
}

void nq_run(cs_nq_handle* h, int first, int count, unsigned long long max_steps,
            unsigned long long allow, int ls_mode, cs_step_stats* stats) {
    nq_run_enqueue(h, first, count, max_steps, allow, ls_mode);
    nq_run_wait(h, stats);
}

// Returns true when some row value was outside [0, n): such rows are stored as 0, so the caller can
// (and must) bring the chains' state in line with what is stored BEFORE reporting the error.
bool nq_upload(cs_nq_handle* h, uint32_t first, uint32_t count, const int64_t* rows) {
    const size_t n = h->cfg.n;
    const size_t per = h->stage_elems / n;  // chains per staging pass
    CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = (uint32_t)((count - done) < per ? (count - done) : per);
        CU(cudaMemcpyAsync(h->d_stage, rows + (size_t)done * n, (size_t)c * n * sizeof(int64_t),
                           cudaMemcpyHostToDevice, h->stream));
        const long long total = (long long)c * h->n_pad;
        const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        nq_pack_rows_kernel<<<grid, 256, 0, h->stream>>>(
            h->d_stage, h->d_rows + (size_t)(first + done) * h->n_pad, (int)n, h->n_pad, (int)c,
            h->d_bad);
        CU(cudaGetLastError());
        // the staging buffer is reused by the next pass
        CU(cudaStreamSynchronize(h->stream));
        done += c;
    }
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return bad != 0;
}

void nq_download(cs_nq_handle* h, const uint16_t* src, uint32_t first, uint32_t count,
                 int64_t* rows) {
    const size_t n = h->cfg.n;
    const size_t per = h->stage_elems / n;
    for (uint32_t done = 0; done < count;) {
        const uint32_t c = (uint32_t)((count - done) < per ? (count - done) : per);
        const long long total = (long long)c * (long long)n;
        const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        nq_unpack_rows_kernel<<<grid, 256, 0, h->stream>>>(
            src + (size_t)(first + done) * h->n_pad, h->d_stage, (int)n, h->n_pad, (int)c);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(rows + (size_t)done * n, h->d_stage, (size_t)c * n * sizeof(int64_t),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        done += c;
    }
}

}  // namespace

#include "nq_big_api.cuh"

extern "C" int32_t cs_nq_create(const cs_nq_config* cfg, cs_nq_handle** out) {
    if (!cfg || !out) return CS_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->n < 1 || cfg->n_chains < 1 || cfg->neighbourhood > CS_NQ_CHANGE)
        return CS_ERR_INVALID_ARG;
    if (cfg->n > CS_NQ_MAX_N) return CS_ERR_UNSUPPORTED;
    const bool big = cfg->n > CS_NQ_MAX_N_SMEM || (cfg->flags & CS_NQ_FLAG_GLOBAL);
    // the L2-resident path holds one instance per handle and scores the swap neighbourhood
    if (big && (cfg->n_chains != 1 || cfg->neighbourhood != CS_NQ_SWAP)) return CS_ERR_UNSUPPORTED;
    if ((uint64_t)cfg->chain_offset + cfg->n_chains > 0xffffffffull) return CS_ERR_INVALID_ARG;
    const bool ref_mode = (cfg->flags & CS_NQ_FLAG_REFERENCE_PROPOSER) != 0;
    if (ref_mode && (big || cfg->neighbourhood != CS_NQ_CHANGE)) return CS_ERR_UNSUPPORTED;
    int ndev = cs_device_count();
    if (ndev <= 0) return CS_ERR_NO_DEVICE;
    int dev = cfg->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) return CS_ERR_NO_DEVICE;
    }
    if (dev >= ndev) return CS_ERR_INVALID_ARG;
    cs_nq_handle* h = new (std::nothrow) cs_nq_handle();
    if (!h) return CS_ERR_OOM;
    h->cfg = *cfg;
    h->device = dev;
    h->is_big = big;
    const int32_t rc = guarded(h, [&] {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, dev));
        h->sm_count = prop.multiProcessorCount;
        const int n = (int)cfg->n;
        h->n_pad = round_up(n, NQ_PAD);
        if (big) {
            CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
            h->own_stream = true;
            CU(cudaMalloc(&h->d_st, sizeof(NqChainState)));
            if (cfg->trace_capacity)
                CU(cudaMalloc(&h->d_trace, cfg->trace_capacity * sizeof(NqTraceEntry)));
            CU(cudaMalloc(&h->d_totals, 2 * sizeof(unsigned long long)));
            CU(cudaMalloc(&h->d_stats, sizeof(NqStats)));
            CU(cudaMalloc(&h->d_bad, sizeof(int)));
            CU(cudaMallocHost(&h->h_stats, sizeof(NqStats)));
            CU(cudaMallocHost(&h->h_totals, 2 * sizeof(unsigned long long)));
            h->stage_elems = n;
            CU(cudaMalloc(&h->d_stage, h->stage_elems * sizeof(int64_t)));
            CU(cudaEventCreate(&h->ev0));
            CU(cudaEventCreate(&h->ev1));
            nqb_alloc(h);
            nq_reset_state_kernel<<<1, 32, 0, h->stream>>>(h->d_st, 0, 1);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(h->stream));
            return;
        }
        h->smem = nq_smem_bytes(h->n_pad);
        h->ref_mode = ref_mode;
        h->window = 5ull * cfg->n;  // examples/nqueens/src/main.rs:130
        if (ref_mode) h->smem += (size_t)4 * h->n_pad;  // proposer scratch behind the layout
        REQUIRE(h->smem <= (size_t)prop.sharedMemPerBlockOptin,
                "board does not fit the shared-memory chain kernel on this device");
        h->threads = n <= 96 ? 128 : n <= 512 ? 256 : n <= 2048 ? 512 : 1024;
        allow_max_smem(nq_step_kernel<NQ_TI>, prop);
        allow_max_smem(nq_rescore_kernel, prop);
        allow_max_smem(nq_eval_kernel, prop);
        // resident CTAs: as many as fit, so small boards run several chains per SM
        int per_sm = 1;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nq_step_kernel<NQ_TI>,
                                                         h->threads, h->smem));
        if (per_sm < 1) per_sm = 1;
        h->grid_v2 = h->sm_count;
        h->sm_count *= per_sm;  // sm_count now = resident CTA slots (grid cap)
        h->use_v2 = cfg->neighbourhood == CS_NQ_SWAP && n >= NQC_MIN_N && n <= NQC_MAX_N;
        if (h->use_v2) {
            const size_t a = nq_smem_bytes_v2(h->n_pad), c = nqc_smem_bytes(h->n_pad);
            h->smem_v2 = a > c ? a : c;
            REQUIRE(h->smem_v2 <= (size_t)prop.sharedMemPerBlockOptin, "packed layout does not fit");
            allow_max_smem(nq_step_kernel_v2<NQ_TI>, prop);
        }
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
        const size_t nc = cfg->n_chains;
        CU(cudaMalloc(&h->d_rows, nc * h->n_pad * sizeof(uint16_t)));
        CU(cudaMalloc(&h->d_best_rows, nc * h->n_pad * sizeof(uint16_t)));
        CU(cudaMalloc(&h->d_st, nc * sizeof(NqChainState)));
        if (cfg->trace_capacity)
            CU(cudaMalloc(&h->d_trace, nc * cfg->trace_capacity * sizeof(NqTraceEntry)));
        CU(cudaMalloc(&h->d_work, sizeof(unsigned int)));
        CU(cudaMalloc(&h->d_totals, 2 * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_stats, sizeof(NqStats)));
        CU(cudaMalloc(&h->d_bad, sizeof(int)));
        CU(cudaMalloc(&h->d_ls_rng, nc * sizeof(unsigned long long)));
        CU(cudaMemset(h->d_ls_rng, 0, nc * sizeof(unsigned long long)));
        CU(cudaMallocHost(&h->h_stats, sizeof(NqStats)));
        CU(cudaMallocHost(&h->h_totals, 2 * sizeof(unsigned long long)));
        // staging for int64 <-> u16 conversion: whole chains, ~64 MiB or one chain
        size_t per = (size_t)(64u << 20) / (n * sizeof(int64_t));
        if (per < 1) per = 1;
        if (per > nc) per = nc;
        h->stage_elems = per * n;
        CU(cudaMalloc(&h->d_stage, h->stage_elems * sizeof(int64_t)));
        CU(cudaEventCreate(&h->ev0));
        CU(cudaEventCreate(&h->ev1));
        CU(cudaMemsetAsync(h->d_rows, 0, nc * h->n_pad * sizeof(uint16_t), h->stream));
        CU(cudaMemsetAsync(h->d_best_rows, 0, nc * h->n_pad * sizeof(uint16_t), h->stream));
        nq_reset_state_kernel<<<(int)((nc + 255) / 256), 256, 0, h->stream>>>(h->d_st, 0, (int)nc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    });
    if (rc != CS_OK) {
        fprintf(stderr, "cs_nq_create failed: %s\n", h->err.c_str());
        nq_free(h);
        delete h;
        return rc;
    }
    *out = h;
    return CS_OK;
}

extern "C" int32_t cs_nq_destroy(cs_nq_handle* h) {
    if (!h) return CS_ERR_INVALID_ARG;
    nq_free(h);
    delete h;
    return CS_OK;
}

extern "C" const char* cs_nq_last_error(const cs_nq_handle* h) { return h ? h->err.c_str() : ""; }

extern "C" int32_t cs_nq_set_stream(cs_nq_handle* h, void* s) {
    return guarded(h, [&] {
        CU(cudaStreamSynchronize(h->stream));
        if (h->own_stream) {
            CU(cudaStreamDestroy(h->stream));
            h->own_stream = false;
        }
        h->stream = (cudaStream_t)s;
    });
}

extern "C" int32_t cs_nq_init_random(cs_nq_handle* h) {
    return guarded(h, [&] {
        if (h->is_big) {
            nqb_init_kernel<<<1, 1, 0, h->stream>>>(h->big.rows, (int)h->cfg.n, h->n_pad, h->cfg.seed,
                                                    h->cfg.chain_offset);
            CU(cudaGetLastError());
            nqb_rebuild(h, true);
            h->scored = true;
            return;
        }
        const int nc = (int)h->cfg.n_chains;
        CU(cudaMemsetAsync(h->d_ls_rng, 0, (size_t)nc * sizeof(unsigned long long), h->stream));
        nq_init_kernel<<<(nc + 63) / 64, 64, 0, h->stream>>>(h->d_rows, h->d_st, (int)h->cfg.n,
                                                              h->n_pad, nc, h->cfg.seed,
                                                              h->cfg.chain_offset);
        CU(cudaGetLastError());
        nq_rescore(h, 0, nc);
        nq_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
    });
}

extern "C" int32_t cs_nq_set_chains(cs_nq_handle* h, uint32_t first, uint32_t count,
                                    const int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        nq_check_range(h, first, count);
        if (h->is_big) {
            const bool bad = nqb_upload(h, rows);
            nqb_rebuild(h, true);
            h->scored = true;
            REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
            return;
        }
        const bool bad = nq_upload(h, first, count, rows);
        CU(cudaMemsetAsync(h->d_ls_rng + first, 0, (size_t)count * sizeof(unsigned long long), h->stream));
        nq_reset_state_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_st, (int)first,
                                                                          (int)count);
        CU(cudaGetLastError());
        nq_rescore(h, (int)first, (int)count);
        if (!h->scored && !(first == 0 && count == h->cfg.n_chains)) {
            // chains never given a solution hold the all-zero board; score them too
            nq_rescore(h, 0, (int)h->cfg.n_chains);
        }
        nq_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
        // the chains now hold the input with out-of-range rows replaced by 0, consistently scored
        REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
    });
}

extern "C" int32_t cs_nq_set_chains_async(cs_nq_handle* h, uint32_t first, uint32_t count,
                                          const int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        nq_check_range(h, first, count);
        if (h->is_big) throw Unsupported{"asynchronous staging is for the chain (shared-memory) path"};
        if (h->async_pending) throw StateFail{"an upload is already pending: call cs_nq_commit_chains first"};
        const size_t need = (size_t)count * h->cfg.n;
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
        // the copy engine works while the handle's stream runs kernels; nothing on the handle's
        // stream reads d_stage_async until cs_nq_commit_chains
        CU(cudaMemcpyAsync(h->d_stage_async, rows, need * sizeof(long long), cudaMemcpyHostToDevice, h->copy_stream));
        CU(cudaEventRecord(h->ev_upload, h->copy_stream));
        h->async_first = first;
        h->async_count = count;
        h->async_pending = true;
    });
}

extern "C" int32_t cs_nq_commit_chains(cs_nq_handle* h) {
    return guarded(h, [&] {
        if (!h->async_pending) throw StateFail{"no pending upload: call cs_nq_set_chains_async first"};
        const uint32_t first = h->async_first, count = h->async_count;
        const size_t n = h->cfg.n;
        h->async_pending = false;
        CU(cudaStreamWaitEvent(h->stream, h->ev_upload, 0));
        CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
        const long long total = (long long)count * h->n_pad;
        const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
        nq_pack_rows_kernel<<<grid, 256, 0, h->stream>>>(h->d_stage_async, h->d_rows + (size_t)first * h->n_pad,
                                                         (int)n, h->n_pad, (int)count, h->d_bad);
        CU(cudaGetLastError());
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemsetAsync(h->d_ls_rng + first, 0, (size_t)count * sizeof(unsigned long long), h->stream));
        nq_reset_state_kernel<<<(count + 255) / 256, 256, 0, h->stream>>>(h->d_st, (int)first, (int)count);
        CU(cudaGetLastError());
        nq_rescore(h, (int)first, (int)count);
        if (!h->scored && !(first == 0 && count == h->cfg.n_chains)) nq_rescore(h, 0, (int)h->cfg.n_chains);
        nq_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        h->scored = true;
        REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
    });
}

extern "C" int32_t cs_nq_get_chains(cs_nq_handle* h, uint32_t first, uint32_t count,
                                    int64_t* rows) {
    return guarded(h, [&] {
        REQUIRE(rows, "rows is NULL");
        nq_check_range(h, first, count);
        if (h->is_big) {
            nqb_download(h, h->big.rows, rows);
            return;
        }
        nq_download(h, h->d_rows, first, count, rows);
    });
}

extern "C" int32_t cs_nq_get_best_chains(cs_nq_handle* h, uint32_t first, uint32_t count,
                                         int64_t* rows, int64_t* best_scores) {
    return guarded(h, [&] {
        nq_check_range(h, first, count);
        if (rows) {
            if (h->is_big) nqb_download(h, h->d_best_rows32, rows);
            else nq_download(h, h->d_best_rows, first, count, rows);
        }
        if (best_scores) {
            std::vector<NqChainState> st(count);
            CU(cudaMemcpyAsync(st.data(), h->d_st + first, count * sizeof(NqChainState),
                               cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            for (uint32_t k = 0; k < count; ++k) best_scores[k] = st[k].best_score;
        }
    });
}

extern "C" int32_t cs_nq_get_scores(cs_nq_handle* h, int64_t* scores) {
    return guarded(h, [&] {
        REQUIRE(scores, "scores is NULL");
        std::vector<NqChainState> st(h->cfg.n_chains);
        CU(cudaMemcpyAsync(st.data(), h->d_st, st.size() * sizeof(NqChainState),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (size_t k = 0; k < st.size(); ++k) scores[k] = st[k].score;
    });
}

extern "C" int32_t cs_nq_get_status(cs_nq_handle* h, uint32_t* status) {
    return guarded(h, [&] {
        REQUIRE(status, "status is NULL");
        std::vector<NqChainState> st(h->cfg.n_chains);
        CU(cudaMemcpyAsync(st.data(), h->d_st, st.size() * sizeof(NqChainState),
                           cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (size_t k = 0; k < st.size(); ++k) status[k] = st[k].status;
    });
}

extern "C" int32_t cs_nq_score_full(cs_nq_handle* h, uint32_t chain, int64_t* score) {
    return guarded(h, [&] {
        REQUIRE(score, "score is NULL");
        nq_check_range(h, chain, 1);
        unsigned long long* d_pairs = (unsigned long long*)h->d_totals;  // reuse, restored below
        CU(cudaMemsetAsync(d_pairs, 0, sizeof(unsigned long long), h->stream));
        const int n = (int)h->cfg.n;
        const int grid = n < 2048 ? (n > 0 ? n : 1) : 2048;
        if (h->is_big)
            nq_pair_score_kernel<unsigned int><<<grid, 256, 0, h->stream>>>(h->big.rows, n, d_pairs);
        else
            nq_pair_score_kernel<uint16_t><<<grid, 256, 0, h->stream>>>(
                h->d_rows + (size_t)chain * h->n_pad, n, d_pairs);
        CU(cudaGetLastError());
        unsigned long long pairs = 0;
        CU(cudaMemcpyAsync(&pairs, d_pairs, sizeof pairs, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        *score = 2 * (int64_t)pairs;
    });
}

extern "C" int32_t cs_nq_eval_moves(cs_nq_handle* h, uint32_t chain, uint32_t kind,
                                    const cs_move* moves, uint64_t n_moves, int64_t* delta) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(kind <= CS_NQ_CHANGE, "unknown move kind");
        if (n_moves == 0) return;
        REQUIRE(moves && delta, "moves/delta is NULL");
        const uint32_t n = h->cfg.n;
        for (uint64_t k = 0; k < n_moves; ++k)
            REQUIRE(moves[k].a < n && moves[k].b < n, "move index outside the board");
        uint2* d_moves = nullptr;
        long long* d_delta = nullptr;
        CU(cudaMalloc(&d_moves, n_moves * sizeof(uint2)));
        cudaError_t e = cudaMalloc(&d_delta, n_moves * sizeof(long long));
        if (e != cudaSuccess) {
            cudaFree(d_moves);
            CU(e);
        }
        try {
            CU(cudaMemcpyAsync(d_moves, moves, n_moves * sizeof(uint2), cudaMemcpyHostToDevice,
                               h->stream));
            if (h->is_big) {
                REQUIRE(kind == CS_NQ_SWAP, "the big-board path scores swap moves only");
                nqb_eval_kernel<<<nqb_grid(h, (long long)n_moves), 256, 0, h->stream>>>(
                    h->big, d_moves, n_moves, d_delta);
            } else {
                NqParams p = nq_params(h, 0, (int)h->cfg.n_chains);
                nq_eval_kernel<<<1, h->threads, h->smem, h->stream>>>(p, (int)chain, (int)kind,
                                                                      d_moves, n_moves, d_delta);
            }
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(delta, d_delta, n_moves * sizeof(long long), cudaMemcpyDeviceToHost,
                               h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } catch (...) {
            cudaFree(d_moves);
            cudaFree(d_delta);
            throw;
        }
        cudaFree(d_moves);
        cudaFree(d_delta);
    });
}

extern "C" int32_t cs_nq_enumerate(cs_nq_handle* h, uint32_t chain, cs_move* moves, uint64_t cap,
                                   uint64_t* n_out) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        const uint32_t n = h->cfg.n;
        std::vector<int64_t> rows(n);
        if (h->is_big) nqb_download(h, h->big.rows, rows.data());
        else nq_download(h, h->d_rows, chain, 1, rows.data());
        uint64_t k = 0;
        if (h->cfg.neighbourhood == CS_NQ_SWAP) {
            for (uint32_t i = 0; i < n; ++i)
                for (uint32_t j = i + 1; j < n; ++j) {
                    if (rows[i] == rows[j]) continue;
                    if (moves && k < cap) moves[k] = cs_move{i, j};
                    ++k;
                }
        } else {
            for (uint32_t c = 0; c < n; ++c)
                for (uint32_t v = 0; v < n; ++v) {
                    if (rows[c] == (int64_t)v) continue;
                    if (moves && k < cap) moves[k] = cs_move{c, v};
                    ++k;
                }
        }
        *n_out = k;
    });
}

extern "C" int32_t cs_nq_neighbourhood_deltas(cs_nq_handle* h, uint32_t chain, int64_t* delta,
                                              uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        if (!h->scored) throw StateFail{"no solution loaded"};
        const uint64_t n = h->cfg.n;
        const uint64_t cnt = h->cfg.neighbourhood == CS_NQ_SWAP ? n * (n - 1) / 2 : n * n;
        *n_out = cnt;
        if (!delta || cnt == 0) return;
        REQUIRE(cap >= cnt, "delta buffer too small");
        long long* d_dump = nullptr;
        CU(cudaMalloc(&d_dump, cnt * sizeof(long long)));
        try {
            CU(cudaMemsetAsync(d_dump, 0x7f, cnt * sizeof(long long), h->stream));
            if (h->is_big) {
                nqb_enqueue_scan(h, nqb_is_perm(h), d_dump);
            } else {
                NqParams p = nq_params(h, (int)chain, 1);
                p.max_steps = 1;
                p.dump = d_dump;
                CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
                nq_launch_step(h, p, 1);
            }
            CU(cudaMemcpyAsync(delta, d_dump, cnt * sizeof(long long), cudaMemcpyDeviceToHost,
                               h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } catch (...) {
            cudaFree(d_dump);
            throw;
        }
        cudaFree(d_dump);
    });
}

extern "C" int32_t cs_nq_band_deltas(cs_nq_handle* h, uint32_t chain, uint32_t i_begin, uint32_t i_end,
                                     int64_t* delta, uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        if (!h->scored) throw StateFail{"no solution loaded"};
        REQUIRE(h->cfg.neighbourhood == CS_NQ_SWAP, "column bands are defined for the swap neighbourhood");
        const uint64_t n = h->cfg.n;
        REQUIRE(i_begin <= i_end && i_end <= (n ? n - 1 : 0), "band outside [0, n-1]");
        auto tri = [&](uint64_t x) { return x * n - x * (x + 1) / 2; };  // entries of columns < x
        const uint64_t base = tri(i_begin), cnt = tri(i_end) - base;
        *n_out = cnt;
        if (!delta || cnt == 0) return;
        REQUIRE(cap >= cnt, "delta buffer too small");
        long long* d_dump = nullptr;
        if (h->is_big) {
            // the scan itself is restricted to the band (the partition range), and writes through a
            // pointer shifted by the band's first index, so only the band is ever allocated.  (An explicit
            // base field in the kernel's parameter struct was measured 7 % slower on the n = 10^6 scan --
            // 246 vs 229 ms per step, more DRAM re-reads -- so the shift stays on the host side.)
            REQUIRE(h->parts == 1, "band dumps need an unpartitioned handle");
            CU(cudaMalloc(&d_dump, cnt * sizeof(long long)));
            const int ib = h->big.i_begin, ie = h->big.i_end;
            try {
                CU(cudaMemsetAsync(d_dump, 0x7f, cnt * sizeof(long long), h->stream));
                h->big.i_begin = (int)i_begin;
                h->big.i_end = (int)i_end;
                nqb_enqueue_scan(h, nqb_is_perm(h), d_dump - (long long)base);
                CU(cudaMemcpyAsync(delta, d_dump, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
            } catch (...) {
                h->big.i_begin = ib;
                h->big.i_end = ie;
                cudaFree(d_dump);
                throw;
            }
            h->big.i_begin = ib;
            h->big.i_end = ie;
            cudaFree(d_dump);
            return;
        }
        // shared-memory path: the production scan dumps the whole neighbourhood on the device (the
        // band is a contiguous slice of the row-major enumeration); only the band crosses PCIe
        const uint64_t all = n * (n - 1) / 2;
        CU(cudaMalloc(&d_dump, all * sizeof(long long)));
        try {
            CU(cudaMemsetAsync(d_dump, 0x7f, all * sizeof(long long), h->stream));
            NqParams p = nq_params(h, (int)chain, 1);
            p.max_steps = 1;
            p.dump = d_dump;
            CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
            nq_launch_step(h, p, 1);
            CU(cudaMemcpyAsync(delta, d_dump + base, cnt * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } catch (...) {
            cudaFree(d_dump);
            throw;
        }
        cudaFree(d_dump);
    });
}

extern "C" int32_t cs_nq_set_window(cs_nq_handle* h, uint64_t window_size) {
    return guarded(h, [&] {
        REQUIRE(window_size >= 1, "window_size must be >= 1");
        h->window = window_size;
    });
}

extern "C" int32_t cs_nq_step(cs_nq_handle* h, uint32_t n_steps, cs_step_stats* stats) {
    return guarded(h, [&] {
        if (h->is_big) nqb_run(h, n_steps, 0, 0, stats);
        else nq_run(h, 0, (int)h->cfg.n_chains, n_steps, 0, 0, stats);
    });
}

extern "C" int32_t cs_nq_step_enqueue(cs_nq_handle* h, uint32_t n_steps) {
    return guarded(h, [&] {
        if (h->is_big) throw Unsupported{"the big-board step is a host loop: use cs_nq_part_scan / cs_nq_part_apply"};
        nq_run_enqueue(h, 0, (int)h->cfg.n_chains, n_steps, 0, 0);
    });
}

extern "C" int32_t cs_nq_step_wait(cs_nq_handle* h, cs_step_stats* stats) {
    return guarded(h, [&] { nq_run_wait(h, stats); });
}

// One block: the reduced key's owner writes its chain, everyone else zeros (cs_b200.h).
__global__ void xchg_select_kernel(const long long* __restrict__ key, const uint16_t* __restrict__ rows,
                                   size_t stride, unsigned chain_offset, unsigned n_chains, unsigned len,
                                   uint16_t* __restrict__ elite, unsigned elite_len) {
    const unsigned gid = (unsigned)((unsigned long long)*key & 0xFFFFFFFFull);
    const bool mine = gid >= chain_offset && gid - chain_offset < n_chains;
    const uint16_t* src = rows + (size_t)(mine ? gid - chain_offset : 0u) * stride;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < elite_len; i += gridDim.x * blockDim.x)
        elite[i] = (mine && i < len) ? src[i] : (uint16_t)0;
}

extern "C" int32_t cs_nq_exchange_select(cs_nq_handle* h, const void* d_key, void* d_elite_u16, uint32_t elite_len) {
    return guarded(h, [&] {
        REQUIRE(d_key && d_elite_u16, "d_key / d_elite_u16 is NULL");
        REQUIRE(!h->is_big, "not available on the big-board path");
        if (elite_len == 0) return;
        const unsigned grid = elite_len > 65536 ? 64u : (elite_len + 1023) / 1024;
        xchg_select_kernel<<<grid, 1024, 0, h->stream>>>((const long long*)d_key, h->d_rows, (size_t)h->n_pad,
                                                         (unsigned)h->cfg.chain_offset, h->cfg.n_chains, h->cfg.n,
                                                         (uint16_t*)d_elite_u16, elite_len);
        CU(cudaGetLastError());
    });
}

extern "C" int32_t cs_nq_local_search(cs_nq_handle* h, uint64_t allow, uint64_t max_iterations,
                                      cs_step_stats* stats) {
    return guarded(h, [&] {
        if (h->is_big) nqb_run(h, max_iterations, allow, 1, stats);
        else nq_run(h, 0, (int)h->cfg.n_chains, max_iterations, allow, 1, stats);
    });
}

extern "C" int32_t cs_nq_local_search_one(cs_nq_handle* h, const int64_t* start, uint64_t allow,
                                          uint64_t max_iterations, int64_t* best,
                                          int64_t* best_score) {
    return guarded(h, [&] {
        REQUIRE(start, "start is NULL");
        if (h->is_big) {
            const bool bad = nqb_upload(h, start);
            nqb_rebuild(h, true);
            h->scored = true;
            REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
            nqb_run(h, max_iterations, allow, 1, nullptr);
            if (best) nqb_download(h, h->d_best_rows32, best);
        } else {
            const bool bad = nq_upload(h, 0, 1, start);
            nq_reset_state_kernel<<<1, 32, 0, h->stream>>>(h->d_st, 0, 1);
            CU(cudaGetLastError());
            nq_rescore(h, 0, h->scored ? 1 : (int)h->cfg.n_chains);
            h->scored = true;
            REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
            nq_run(h, 0, 1, max_iterations, allow, 1, nullptr);
            if (best) nq_download(h, h->d_best_rows, 0, 1, best);
        }
        if (best_score) {
            NqChainState st;
            CU(cudaMemcpyAsync(&st, h->d_st, sizeof st, cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            *best_score = st.best_score;
        }
    });
}

extern "C" int32_t cs_nq_get_trace(cs_nq_handle* h, uint32_t chain, cs_move* moves,
                                   int64_t* score_after, uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(n_out, "n_out is NULL");
        NqChainState st;
        CU(cudaMemcpyAsync(&st, h->d_st + chain, sizeof st, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        *n_out = st.steps;
        uint64_t k = st.steps;
        if (k > h->cfg.trace_capacity) k = h->cfg.trace_capacity;
        if (k > cap) k = cap;
        if (k == 0) return;
        std::vector<NqTraceEntry> t(k);
        CU(cudaMemcpyAsync(t.data(), h->d_trace + (size_t)chain * h->cfg.trace_capacity,
                           k * sizeof(NqTraceEntry), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (uint64_t q = 0; q < k; ++q) {
            if (moves) moves[q] = cs_move{t[q].a, t[q].b};
            if (score_after) score_after[q] = t[q].score_after;
        }
    });
}

extern "C" int32_t cs_nq_best(cs_nq_handle* h, int64_t* rows, int64_t* score, uint32_t* chain) {
    return guarded(h, [&] {
        if (!h->scored) throw StateFail{"no solution loaded"};
        nq_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        if (score) *score = h->h_stats->best_score;
        if (chain) *chain = h->h_stats->best_chain;
        if (rows) {
            if (h->is_big) nqb_download(h, h->big.rows, rows);
            else nq_download(h, h->d_rows, h->h_stats->best_chain, 1, rows);
        }
    });
}

extern "C" int32_t cs_nq_best_key_device_ptr(cs_nq_handle* h, void** dptr) {
    return guarded(h, [&] {
        REQUIRE(dptr, "dptr is NULL");
        *dptr = (void*)&h->d_stats->best_key;
    });
}

extern "C" int32_t cs_nq_chain_device_ptr(cs_nq_handle* h, uint32_t chain, void** dptr,
                                          uint32_t* stride_elems) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(dptr, "dptr is NULL");
        REQUIRE(!h->is_big, "big-board rows are uint32; use cs_nq_get_chains");
        *dptr = (void*)(h->d_rows + (size_t)chain * h->n_pad);
        if (stride_elems) *stride_elems = (uint32_t)h->n_pad;
    });
}

extern "C" int32_t cs_nq_set_chain_u16_device(cs_nq_handle* h, uint32_t chain,
                                              const void* d_rows_u16) {
    return guarded(h, [&] {
        nq_check_range(h, chain, 1);
        REQUIRE(d_rows_u16, "d_rows_u16 is NULL");
        REQUIRE(!h->is_big, "not available on the big-board path");
        CU(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->stream));
        nq_copy_rows_checked_kernel<<<(h->n_pad + 255) / 256, 256, 0, h->stream>>>(
            (const uint16_t*)d_rows_u16, h->d_rows + (size_t)chain * h->n_pad, (int)h->cfg.n, h->n_pad, h->d_bad);
        CU(cudaGetLastError());
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        nq_reset_state_kernel<<<1, 32, 0, h->stream>>>(h->d_st, (int)chain, 1);
        CU(cudaGetLastError());
        nq_rescore(h, (int)chain, 1);
        nq_refresh_stats(h);
        CU(cudaStreamSynchronize(h->stream));
        REQUIRE(!bad, "row value outside [0, n) (stored as 0)");
    });
}

// ------------------------------------------------------------------ n-queens ILS shell
namespace {

__global__ void nq_gather_keys_kernel(const NqChainState* st, long long* best_key, long long* cur_key,
                                      int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (best_key) best_key[k] = st[k].best_score;
    if (cur_key) cur_key[k] = st[k].score;
}

IlsParams nq_ils_params(cs_nq_handle* h) {
    IlsParams p = h->ils.params(h->d_rows, h->d_best_rows, h->cfg.seed, h->cfg.chain_offset);
    p.value_range = (int)h->cfg.n;
    p.restart_is_perm = 1;
    p.do_nothing_first = 0;  // nqueens lib.rs:277-280: ChangeSubset listed first
    p.k_before_shuffle = 0;  // lib.rs:303-306: shuffle, then the subset size
    return p;
}

}  // namespace

extern "C" int32_t cs_nq_ils_init(cs_nq_handle* h, uint32_t cap, uint32_t log_cap) {
    return guarded(h, [&] {
        REQUIRE(!h->is_big, "the ILS shell runs on the shared-memory chain path");
        REQUIRE(cap >= 1 && cap <= ILS_MAX_CAP, "best_solutions_capacity must be 1..64");
        if (!h->scored) throw StateFail{"no solution loaded"};
        const int nc = (int)h->cfg.n_chains;
        h->ils.alloc(nc, (int)h->cfg.n, h->n_pad, (int)cap, (int)log_cap);
        ils_reset_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->ils.d_st, h->ils.d_cur_key,
                                                                    h->ils.d_skip, nc);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h->ils.d_cur, h->d_rows, (size_t)nc * h->n_pad * sizeof(uint16_t),
                           cudaMemcpyDeviceToDevice, h->stream));
        nq_gather_keys_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_st, nullptr,
                                                                         h->ils.d_cur_key, nc);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    });
}

extern "C" int32_t cs_nq_ils_run(cs_nq_handle* h, uint32_t rounds, uint64_t ls_max_iterations,
                                 uint64_t allow, uint32_t stop_when_any_best, cs_ils_stats* stats) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_nq_ils_init first"};
        const int nc = (int)h->cfg.n_chains;
        IlsParams ip = nq_ils_params(h);
        NqParams lp = nq_params(h, 0, nc);
        lp.max_steps = ls_max_iterations;
        lp.allow_no_improve = allow;
        lp.ls_mode = 1;
        lp.skip = h->ils.d_skip;
        const int ig = ils_grid(nc, h->sm_count);
        unsigned launches = 0, run = 0;
        CU(cudaMemsetAsync(h->d_totals, 0, 2 * sizeof(unsigned long long), h->stream));
        CU(cudaEventRecord(h->ev0, h->stream));
        for (uint32_t r = 0; r < rounds; ++r) {
            ils_perturb_kernel<<<ig, ILS_THREADS, h->ils.perturb_smem, h->stream>>>(ip);
            CU(cudaMemsetAsync(h->d_work, 0, sizeof(unsigned int), h->stream));
            nq_launch_step(h, lp, nc);
            nq_gather_keys_kernel<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->d_st, h->ils.d_neu_key,
                                                                             nullptr, nc);
            ils_accept_kernel<<<ig, ILS_THREADS, 0, h->stream>>>(ip);
            CU(cudaGetLastError());
            launches += 4;
            ++run;
            if (stop_when_any_best) {
                ils_summary_kernel<<<1, 1024, 0, h->stream>>>(ip, h->ils.d_sum);
                CU(cudaMemcpyAsync(h->ils.h_sum, h->ils.d_sum, sizeof(IlsSummary),
                                   cudaMemcpyDeviceToHost, h->stream));
                CU(cudaStreamSynchronize(h->stream));
                ++launches;
                if (h->ils.h_sum->chains_done) break;
            }
        }
        ils_summary_kernel<<<1, 1024, 0, h->stream>>>(ip, h->ils.d_sum);
        CU(cudaMemcpyAsync(h->ils.h_sum, h->ils.d_sum, sizeof(IlsSummary), cudaMemcpyDeviceToHost,
                           h->stream));
        // publish the ILS current as the chains' solution so the ordinary getters see it
        CU(cudaMemcpyAsync(h->d_rows, h->ils.d_cur, (size_t)nc * h->n_pad * sizeof(uint16_t),
                           cudaMemcpyDeviceToDevice, h->stream));
        nq_rescore(h, 0, nc);
        nq_refresh_stats(h);
        CU(cudaEventRecord(h->ev1, h->stream));
        CU(cudaMemcpyAsync(h->h_totals, h->d_totals, 2 * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, h->stream));
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

extern "C" int32_t cs_nq_ils_get_best(cs_nq_handle* h, uint32_t chain, int64_t* rows, int64_t* score) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_nq_ils_init first"};
        nq_check_range(h, chain, 1);
        IlsChainState st;
        CU(cudaMemcpyAsync(&st, h->ils.d_st + chain, sizeof st, cudaMemcpyDeviceToHost, h->stream));
        unsigned char slot = 0;
        CU(cudaMemcpyAsync(&slot, h->ils.d_order + (size_t)chain * h->ils.cap, 1, cudaMemcpyDeviceToHost,
                           h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (!st.size) throw StateFail{"no round has run yet (the reference unwrap()s None here)"};
        if (score) {
            CU(cudaMemcpyAsync(score, h->ils.d_bset_key + (size_t)chain * h->ils.cap + slot,
                               sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        }
        if (rows)
            nq_download(h, h->ils.d_bset + ((size_t)chain * h->ils.cap + slot) * h->n_pad -
                               (size_t)0 * h->n_pad, 0, 1, rows);
    });
}

extern "C" int32_t cs_nq_ils_get_log(cs_nq_handle* h, uint32_t chain, int64_t* new_key,
                                     uint32_t* choice, uint64_t cap, uint64_t* n_out) {
    return guarded(h, [&] {
        if (!h->ils.ready) throw StateFail{"call cs_nq_ils_init first"};
        nq_check_range(h, chain, 1);
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
        CU(cudaMemcpyAsync(log.data(), h->ils.d_log + (size_t)chain * h->ils.log_cap,
                           k * sizeof(IlsLogEntry), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (uint64_t q = 0; q < k; ++q) {
            if (new_key) new_key[q] = log[q].new_key;
            if (choice) choice[q] = log[q].choice;
        }
    });
}

#include "es_api.cuh"
