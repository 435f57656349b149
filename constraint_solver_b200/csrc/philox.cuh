// Philox4x32-10 counter-based RNG, identical on host and device so any chain can be
// replayed on the CPU.  Stands in for the reference's ChaCha20Rng (examples/nqueens/src/
// main.rs:39,66); the stream is replayable, not reference-identical (rand_chacha is an
// un-vendored crates.io dependency).
#pragma once
#include <stdint.h>

namespace csb {

#ifdef __CUDACC__
#define CSB_HD __host__ __device__ __forceinline__
#else
#define CSB_HD inline
#endif

struct Philox4 {
    uint32_t v[4];
};

CSB_HD uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

CSB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                             uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = philox_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = philox_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.v[0] = c0;
    o.v[1] = c1;
    o.v[2] = c2;
    o.v[3] = c3;
    return o;
}

// key = {seed lo, seed hi}; ctr = {counter lo, counter hi, chain, purpose}
CSB_HD Philox4 philox_stream(uint64_t seed, uint32_t chain, uint32_t purpose, uint64_t counter) {
    return philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), chain, purpose,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Sequential 32-bit draws from one stream (block t/4, word t%4).
struct PhiloxDraws {
    uint64_t seed;
    uint32_t chain, purpose;
    uint64_t t;
    Philox4 blk;
    CSB_HD PhiloxDraws(uint64_t s, uint32_t c, uint32_t p, uint64_t t0 = 0)
        : seed(s), chain(c), purpose(p), t(t0) {
        blk = philox_stream(seed, chain, purpose, t >> 2);
    }
    CSB_HD uint32_t next() {
        if ((t & 3) == 0) blk = philox_stream(seed, chain, purpose, t >> 2);
        const uint32_t r = blk.v[t & 3];
        ++t;
        return r;
    }
    // uniform index in [0, range) by multiply-shift (deterministic, mirrored by the oracle)
    CSB_HD uint32_t below(uint32_t range) { return philox_mulhi(next(), range); }
};

}  // namespace csb
