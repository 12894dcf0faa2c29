// philox_map.h -- the contractual site -> random-number map of the reference, restated for
// a layout-free implementation, plus the per-step constant folding used by the kernels.
//
// Contract (reference: optimized/main.cu:112-116, :514-515, :620-621, :637-659; cuRAND
// curand_kernel.h curand_init/skipahead, curand_philox4x32_x.h:160-192):
//   thread id   tid  = ((global_row/16) * (X/2048) + c/32) * 256 + (global_row%16)*16 + (c%16)
//               for the 32-site word c of a colour row (u32 arithmetic),
//   step        s    = 2*it + colour            (it = 0 for the lattice init),
//   call        cl   = 8*((c%32)/16) + q,  q = 0..7,
//   counter          = (lo32(16*s + cl), hi32(16*s + cl), tid, 0),   key = (lo32 seed, hi32 seed),
//   outputs x,y,z,w  -> sites 32*c + {2q, 16+2q, 2q+1, 17+2q}.
//
// Folding: the first two Philox rounds only mix values that are either uniform over the
// launch (counter words, keys) or constant per thread (tid).  StepTables carries the uniform
// halves, pre-multiplied on the host; ThreadConsts the per-thread halves, computed once per
// thread and reused by its 16 calls.  Each call then starts at round 3: 16 wide multiplies
// instead of 20.  philox_fast_*() is compiled for host too so tests can check it against the
// plain 10-round function without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ISING_HD __host__ __device__ __forceinline__
#else
#define ISING_HD inline
#endif

namespace ising {

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

constexpr int REF_BLOCK_ROWS  = 16;  // BLOCK_Y*BMULT_Y, optimized/main.cu:56-58
constexpr int REF_BLOCK_WORDS = 32;  // BLOCK_X*BMULT_X 128-bit words = 32-site words per block row
constexpr int CALLS_PER_TID   = 16;  // 64 draws per reference thread / 4

ISING_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#if defined(__CUDA_ARCH__)
    const uint64_t p = (uint64_t)a * b;   // IMAD.WIDE.U32
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

// Plain Philox4x32-10 (used by host code and as the reference for the folded version).
ISING_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                            uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint32_t h0, l0, h1, l1;
        mulhilo(PHILOX_M0, c0, h0, l0);
        mulhilo(PHILOX_M1, c2, h1, l1);
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Uniform-per-launch part of rounds 1-3 and the round keys of rounds 4-10.
struct StepTables {
    uint32_t x1;                   // hi32(counter) ^ key0[round 1]
    uint32_t k1r3;                 // key1 of round 3 (folded into ThreadConsts::e)
    uint32_t a[CALLS_PER_TID];     // mulhi(M1, c2') ^ key0[round 2]
    uint32_t b[CALLS_PER_TID];     // mullo(M0, ctr_lo) ^ key1[round 2]
    uint32_t d[CALLS_PER_TID];     // mullo(M1, c2') ^ key0[round 3]
    uint32_t k0[7], k1[7];         // keys of rounds 4..10
};

// Tables for the 16 consecutive Philox counters base .. base+15 (the low counter word must not
// carry inside the group: base is a multiple of 16 for the optimized/ map, of 4 -- with 4 calls
// used -- for the tensorcore/ map).  Subsequence (tid) must be below 2^32.
inline StepTables make_step_tables_at(uint64_t seed, uint64_t base)
{
    StepTables t;
    uint32_t k0[10], k1[10];
    k0[0] = (uint32_t)seed; k1[0] = (uint32_t)(seed >> 32);
    for (int r = 1; r < 10; r++) { k0[r] = k0[r - 1] + PHILOX_W0; k1[r] = k1[r - 1] + PHILOX_W1; }
    t.x1 = (uint32_t)(base >> 32) ^ k0[0];
    t.k1r3 = k1[2];
    for (int cl = 0; cl < CALLS_PER_TID; cl++) {
        const uint32_t c0 = (uint32_t)(base + cl);
        uint32_t hi0, lo0, hi1p, lo1p;
        mulhilo(PHILOX_M0, c0, hi0, lo0);
        const uint32_t c2p = hi0 ^ 0u /* tid_hi */ ^ k1[0];
        mulhilo(PHILOX_M1, c2p, hi1p, lo1p);
        t.a[cl] = hi1p ^ k0[1];
        t.b[cl] = lo0 ^ k1[1];
        t.d[cl] = lo1p ^ k0[2];
    }
    for (int r = 0; r < 7; r++) { t.k0[r] = k0[3 + r]; t.k1[r] = k1[3 + r]; }
    return t;
}

// optimized/ map: curand_init offset 64*step draws = 16*step calls
inline StepTables make_step_tables(uint64_t seed, uint64_t step /* 2*it + colour */)
{
    return make_step_tables_at(seed, 16ull * step);
}

// Constant-per-thread part (depends on tid and, through x1, on the step).
struct ThreadConsts {
    uint32_t lo1;   // mullo(M1, tid)                          -> c1 after round 1
    uint32_t hi0p;  // mulhi(M0, c0')                          -> feeds c2 after round 2
    uint32_t e;     // mullo(M0, c0') ^ key1[round 3]
};

ISING_HD ThreadConsts make_thread_consts(uint32_t tid, uint32_t x1, uint32_t k1r3)
{
    ThreadConsts tc;
    uint32_t hi1, lo0p;
    mulhilo(PHILOX_M1, tid, hi1, tc.lo1);
    const uint32_t c0p = hi1 ^ x1;
    mulhilo(PHILOX_M0, c0p, tc.hi0p, lo0p);
    tc.e = lo0p ^ k1r3;
    return tc;
}

// The four outputs of call cl for the thread described by tc.  ta/tb/td = tables.a/b/d[cl].
ISING_HD void philox_fast(const ThreadConsts& tc, uint32_t ta, uint32_t tb, uint32_t td,
                          const uint32_t (&k0)[7], const uint32_t (&k1)[7], uint32_t out[4])
{
    uint32_t c0 = ta ^ tc.lo1;            // state after round 2 (c1 and c3 are folded below)
    uint32_t c2 = tc.hi0p ^ tb;
    uint32_t h0, l0, h1, l1;
    mulhilo(PHILOX_M0, c0, h0, l0);       // round 3
    mulhilo(PHILOX_M1, c2, h1, l1);
    c0 = h1 ^ td; uint32_t c1 = l1; c2 = h0 ^ tc.e; uint32_t c3 = l0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 7; r++) {         // rounds 4..10
        mulhilo(PHILOX_M0, c0, h0, l0);
        mulhilo(PHILOX_M1, c2, h1, l1);
        c0 = h1 ^ c1 ^ k0[r]; c1 = l1; c2 = h0 ^ c3 ^ k1[r]; c3 = l0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------
// B200 only: the high half of a 32x32 multiply on the FP64 pipe.
//
// IMAD.WIDE holds the FMA-heavy pipe of an SM sub-partition for 4 cycles and costs ALU dispatch
// slots on top; the FP64 pipe of the B200 (16 lanes per sub-partition, one
// DFMA per 2 cycles) is otherwise idle in this kernel.  For a 32-bit a held as the low word of
// the double (0x43300000, a) = 2^52 + a,
//     fma_rz(2^52 + a, M * 2^-32, 2^52 - M * 2^20) = RZ(2^52 + a*M / 2^32) = 2^52 + floor(a*M / 2^32):
// every operand is exact, the sum lies in [2^52, 2^52 + 2^32) where one ulp is 1, so
// round-toward-zero is floor() and the low word of the result is mulhi(a, M) -- and its high word
// is 0x43300000 again, so the XOR of the next Philox round can be done in place on the low word
// and the pair fed straight into the next DFMA.  The low half of the product is a plain IMAD.
// ------------------------------------------------------------------------------------------
constexpr uint32_t DBL_2P52_HI = 0x43300000u;

template <uint32_t M>
__device__ __forceinline__ double mulhi_dfma(double a_pair)
{
    constexpr double MS = (double)M * (1.0 / 4294967296.0);               // M * 2^-32, exact
    constexpr double C = 4503599627370496.0 - (double)M * 1048576.0;      // 2^52 - M * 2^20, exact
    return __fma_rz(a_pair, MS, C);
}

// philox_fast with a choice of multiplier per constant: WIDE0/WIDE1 = true keeps IMAD.WIDE for
// the M0 / M1 multiply, false sends its high half to the FP64 pipe.
template <bool WIDE0, bool WIDE1>
__device__ __forceinline__ void philox_fast_mix(const ThreadConsts& tc, uint32_t ta, uint32_t tb, uint32_t td,
                                                const uint32_t (&k0)[7], const uint32_t (&k1)[7], uint32_t out[4],
                                                double* pair0 = nullptr, double* pair2 = nullptr)
{
    uint32_t c0 = ta ^ tc.lo1, c2 = tc.hi0p ^ tb, c1 = 0, c3 = 0;
    double d0 = __hiloint2double((int)DBL_2P52_HI, (int)c0), d2 = __hiloint2double((int)DBL_2P52_HI, (int)c2);
#pragma unroll
    for (int r = -1; r < 7; r++) {        // r = -1 is round 3 (c1, c3 and the keys are folded into td, tc.e)
        uint32_t h0, l0, h1, l1;
        double h0d = 0.0, h1d = 0.0;
        if (WIDE0) mulhilo(PHILOX_M0, c0, h0, l0);
        else { h0d = mulhi_dfma<PHILOX_M0>(d0); l0 = c0 * PHILOX_M0; h0 = (uint32_t)__double2loint(h0d); }
        if (WIDE1) mulhilo(PHILOX_M1, c2, h1, l1);
        else { h1d = mulhi_dfma<PHILOX_M1>(d2); l1 = c2 * PHILOX_M1; h1 = (uint32_t)__double2loint(h1d); }
        if (r < 0) { c0 = h1 ^ td; c2 = h0 ^ tc.e; }
        else       { c0 = h1 ^ c1 ^ k0[r]; c2 = h0 ^ c3 ^ k1[r]; }
        c1 = l1; c3 = l0;
        // the pair that produced h1 / h0 already carries 0x43300000 in its high word
        if (!WIDE0) d0 = __hiloint2double(WIDE1 ? (int)DBL_2P52_HI : __double2hiint(h1d), (int)c0);
        if (!WIDE1) d2 = __hiloint2double(WIDE0 ? (int)DBL_2P52_HI : __double2hiint(h0d), (int)c2);
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    if (pair0) *pair0 = d0;   // (0x43300000, out[0]) and (0x43300000, out[2]): 2^52 + x, ordered like x
    if (pair2) *pair2 = d2;
}
#endif

// Reference thread id of 32-site word c of global row grow (u32 arithmetic, main.cu:514-515).
ISING_HD uint32_t ref_tid(uint32_t grow, uint32_t c, uint32_t grid_x)
{
    return ((grow / REF_BLOCK_ROWS) * grid_x + c / REF_BLOCK_WORDS) * 256u + (grow % REF_BLOCK_ROWS) * 16u + (c % 16u);
}

// curand_uniform of a raw 32-bit output (curand_uniform.h:69-72).
inline float uniform_from_u32(uint32_t x)
{
    volatile float f = (float)x;
    volatile float m = f * 2.3283064e-10f;
    volatile float r = m + (2.3283064e-10f / 2.0f);
    return r;
}

}  // namespace ising
