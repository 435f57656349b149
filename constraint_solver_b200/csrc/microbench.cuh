// On-chip bandwidth micro-benchmarks: the MEASURED denominators of the roofline block.
//
// The chain kernels are bound by the shared-memory data pipe (n <= 12 096) or by the L2 -> L1 path
// (one n = 10^6 board), not by HBM, and MEASURED_PEAKS.json only holds an HBM figure.  These two
// kernels measure what the hardware sustains for the access shapes that matter here:
//   * shared memory: conflict-free LDS (32-bit, one wavefront per warp instruction, and 128-bit,
//     four wavefronts) streamed by every SM -- the ceiling "wavefronts x 128 B" is compared with;
//   * L2: every SM streaming 16-byte loads over a 32 MB L2-resident buffer with L1 bypassed
//     (ld.global.cg), the shape of the packed big-board scan's gather windows.
// cs_microbench (cs_api.cu) times them with CUDA events and returns GB/s.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csb {

constexpr int MB_SMEM_WORDS = 8192;  // 32 KB per CTA

// WIDE = false: LDS.32, lane-consecutive words (1 wavefront per instruction)
// WIDE = true : LDS.128, lane-consecutive 16-byte words (4 wavefronts per instruction)
template <bool WIDE>
__global__ void __launch_bounds__(1024, 2) mb_smem_kernel(int iters, unsigned int* sink) {
    __shared__ __align__(16) unsigned int buf[MB_SMEM_WORDS];
    for (int k = threadIdx.x; k < MB_SMEM_WORDS; k += blockDim.x) buf[k] = (unsigned)k * 2654435761u;
    __syncthreads();
    unsigned int acc = 0;
    if (WIDE) {
        const uint4* b4 = (const uint4*)buf;
        int idx = threadIdx.x;  // lanes read consecutive 16-byte words: conflict-free, 4 wavefronts per instruction
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {  // eight DISTINCT addresses per thread and iteration
                const uint4 v = b4[(idx + u * 160) & (MB_SMEM_WORDS / 4 - 1)];
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            idx = (idx + 32) & (MB_SMEM_WORDS / 4 - 1);
        }
    } else {
        int idx = threadIdx.x;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += buf[(idx + u * 1056) & (MB_SMEM_WORDS - 1)];
            idx = (idx + 32) & (MB_SMEM_WORDS - 1);
        }
    }
    if (acc == 0x12345678u) *sink = acc;  // keeps the loads alive
}

// every thread streams 16-byte words of a buffer that fits L2 but not L1; .cg = cache in L2 only
__global__ void __launch_bounds__(512) mb_l2_kernel(const uint4* __restrict__ buf, size_t n_vec, int passes,
                                                    unsigned int* sink) {
    unsigned int acc = 0;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t k = tid; k < n_vec; k += nt) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "l"(buf + k));
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

}  // namespace csb
