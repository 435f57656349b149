// Fixed-width bitsets (1..4 64-bit words) for the employee-scheduling kernels: slot masks
// (bit t <=> the employee holds slot t), day-window masks and count-occupancy sets.  W = 1 is
// the reference-sized rota (<= 64 scored days) and compiles to plain 64-bit arithmetic; W = 2, 3
// carry rotas of up to 128 / 192 slots (long horizons, several shifts per day).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csb {

typedef unsigned long long u64;

template <int W>
struct Bits {
    u64 w[W];

    __host__ __device__ __forceinline__ static Bits zero() {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) r.w[i] = 0ull;
        return r;
    }
    // bit d; all-zero when d is outside the set (d == 64 * W: a count bin one past the end)
    __host__ __device__ __forceinline__ static Bits bit(int d) {
        Bits r = zero();
        if (W == 1) {
            if (d < 64) r.w[0] = 1ull << d;
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i)
                if ((d >> 6) == i) r.w[i] = 1ull << (d & 63);
        }
        return r;
    }
    // bits 0 .. n-1
    __host__ __device__ __forceinline__ static Bits lowmask(int n) {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const int k = n - 64 * i;
            r.w[i] = k >= 64 ? ~0ull : (k <= 0 ? 0ull : ((1ull << k) - 1ull));
        }
        return r;
    }
    __host__ __device__ __forceinline__ bool test(int d) const {
        if (W == 1) return (w[0] >> d) & 1ull;
        u64 x = 0;
#pragma unroll
        for (int i = 0; i < W; ++i)
            if ((d >> 6) == i) x = w[i];
        return (x >> (d & 63)) & 1ull;
    }
    __host__ __device__ __forceinline__ Bits operator&(const Bits& o) const {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) r.w[i] = w[i] & o.w[i];
        return r;
    }
    __host__ __device__ __forceinline__ Bits operator|(const Bits& o) const {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) r.w[i] = w[i] | o.w[i];
        return r;
    }
    __host__ __device__ __forceinline__ Bits operator^(const Bits& o) const {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) r.w[i] = w[i] ^ o.w[i];
        return r;
    }
    __host__ __device__ __forceinline__ Bits operator~() const {
        Bits r;
#pragma unroll
        for (int i = 0; i < W; ++i) r.w[i] = ~w[i];
        return r;
    }
    // logical shift right by k >= 0 (bit t of the result = bit t + k of *this)
    __host__ __device__ __forceinline__ Bits shr(int k) const {
        Bits r;
        if (W == 1) {
            r.w[0] = k >= 64 ? 0ull : (w[0] >> k);
            return r;
        }
        const int q = k >> 6, b = k & 63;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            u64 lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                if (j == i + q) lo = w[j];
                if (j == i + q + 1) hi = w[j];
            }
            r.w[i] = b ? ((lo >> b) | (hi << (64 - b))) : lo;
        }
        return r;
    }
    __device__ __forceinline__ int popc() const {
        int c = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) c += __popcll(w[i]);
        return c;
    }
    __host__ __device__ __forceinline__ bool any() const {
        u64 x = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) x |= w[i];
        return x != 0;
    }
    // index of the lowest / highest set bit (-1: empty)
    __device__ __forceinline__ int ffs() const {
        int r = -1;
#pragma unroll
        for (int i = W - 1; i >= 0; --i)
            if (w[i]) r = 64 * i + __ffsll((long long)w[i]) - 1;
        return r;
    }
    __device__ __forceinline__ int fls() const {
        int r = -1;
#pragma unroll
        for (int i = 0; i < W; ++i)
            if (w[i]) r = 64 * i + 63 - __clzll((long long)w[i]);
        return r;
    }
    // number of set bits strictly below position f
    __device__ __forceinline__ int rank_below(int f) const { return (*this & lowmask(f)).popc(); }
};

// the low WD words of a wider set known to be empty above them (day-indexed sets of a rota with at most
// 64 * WD days, stored in slot-width words): arithmetic on Bits<WD> instead of Bits<W>
template <int WD, int W>
__host__ __device__ __forceinline__ Bits<WD> bits_lo(const Bits<W>& a) {
    static_assert(WD >= 1 && WD <= W, "WD words of a W-word set");
    Bits<WD> r;
#pragma unroll
    for (int i = 0; i < WD; ++i) r.w[i] = a.w[i];
    return r;
}

// set bit d of a bitset in shared memory: a NATIVE 32-bit atomic on the half-word that holds it (a 64-bit
// shared-memory atomicOr compiles to a compare-and-swap loop; ncu showed it at 8 % of the instructions
// and 16 % of the stall samples of the small-rota step)
__device__ __forceinline__ void bits_atomic_or32(void* base, int d) {
    atomicOr((unsigned int*)base + (d >> 5), 1u << (d & 31));
}
template <int W>
__device__ __forceinline__ void bits_atomic_or(Bits<W>* p, int d) {
    bits_atomic_or32((void*)p, d);
}

}  // namespace csb
