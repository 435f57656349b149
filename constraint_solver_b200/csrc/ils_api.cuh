// Host side of the ILS shell, shared by both plug-ins (included by cs_api.cu).
#pragma once
#include "ils_kernels.cuh"

namespace {

struct IlsHost {
    bool ready = false;
    int cap = 0, log_cap = 0, n_chains = 0, len = 0, stride = 0;
    uint16_t* d_cur = nullptr;
    long long* d_cur_key = nullptr;
    long long* d_neu_key = nullptr;
    uint16_t* d_bset = nullptr;
    long long* d_bset_key = nullptr;
    unsigned char* d_order = nullptr;
    IlsChainState* d_st = nullptr;
    unsigned int* d_skip = nullptr;
    IlsLogEntry* d_log = nullptr;
    IlsSummary* d_sum = nullptr;
    IlsSummary* h_sum = nullptr;  // pinned
    size_t perturb_smem = 0;

    void release() {
        cudaFree(d_cur);
        cudaFree(d_cur_key);
        cudaFree(d_neu_key);
        cudaFree(d_bset);
        cudaFree(d_bset_key);
        cudaFree(d_order);
        cudaFree(d_st);
        cudaFree(d_skip);
        cudaFree(d_log);
        cudaFree(d_sum);
        if (h_sum) cudaFreeHost(h_sum);
        *this = IlsHost();
        cudaGetLastError();
    }

    void alloc(int chains, int len_, int stride_, int cap_, int log_cap_) {
        release();
        n_chains = chains;
        len = len_;
        stride = stride_;
        cap = cap_;
        log_cap = log_cap_;
        const size_t nc = chains;
        CU(cudaMalloc(&d_cur, nc * stride * sizeof(uint16_t)));
        CU(cudaMalloc(&d_cur_key, nc * sizeof(long long)));
        CU(cudaMalloc(&d_neu_key, nc * sizeof(long long)));
        CU(cudaMalloc(&d_bset, nc * cap * stride * sizeof(uint16_t)));
        CU(cudaMalloc(&d_bset_key, nc * cap * sizeof(long long)));
        CU(cudaMalloc(&d_order, nc * cap));
        CU(cudaMalloc(&d_st, nc * sizeof(IlsChainState)));
        CU(cudaMalloc(&d_skip, nc * sizeof(unsigned int)));
        if (log_cap) CU(cudaMalloc(&d_log, nc * log_cap * sizeof(IlsLogEntry)));
        CU(cudaMalloc(&d_sum, sizeof(IlsSummary)));
        CU(cudaMallocHost(&h_sum, sizeof(IlsSummary)));
        perturb_smem = (size_t)(((len + 7) & ~7) + len) * sizeof(uint16_t);
        {   // per function and device, not per handle: always the device's opt-in maximum (see allow_max_smem)
            int dev = 0, optin = 0;
            CU(cudaGetDevice(&dev));
            CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            allow_max_smem(ils_perturb_kernel, optin);
        }
        ready = true;
    }

    IlsParams params(uint16_t* work, const uint16_t* neu, unsigned long long seed,
                     unsigned int chain_offset) const {
        IlsParams p{};
        p.n_chains = n_chains;
        p.len = len;
        p.stride = stride;
        p.best_cap = cap;
        p.cur = d_cur;
        p.cur_key = d_cur_key;
        p.work = work;
        p.neu = neu;
        p.neu_key = d_neu_key;
        p.bset = d_bset;
        p.bset_key = d_bset_key;
        p.order = d_order;
        p.st = d_st;
        p.skip = d_skip;
        p.log = d_log;
        p.log_cap = log_cap;
        p.seed = seed;
        p.chain_offset = chain_offset;
        return p;
    }
};

inline int ils_grid(int chains, int sm) { return chains < sm * 16 ? chains : sm * 16; }

}  // namespace
