// host_math.h -- host-side arithmetic of the boundary: the float acceptance table of the
// reference (optimized/main.cu:1683-1697) and its exact integer form.
#pragma once
#include <math.h>
#include <stdint.h>

#include "kernels.cuh"
#include "philox_map.h"

namespace ising {

// exp_h[1][a] of main.cu:1687 for a = number of aligned neighbours; exp_h[0][n] == exp_h[1][4-n]
// bit for bit (the two float products are equal), so one row describes the whole table.
inline void accept_table(float temp, float out[5])
{
    for (int a = 0; a < 5; a++) {
        if (temp > 0) out[a] = expf(-2.0f * (float)(a * 2 - 4) * (1.0f / temp));
        else          out[a] = (a == 2) ? 0.5f : -2.0f * (float)(a * 2 - 4);
    }
}

// #{x in [0,2^32) : uniform(x) <= p}  (strict: < p).  uniform() is monotone non-decreasing in x,
// so the accepted set is a prefix [0, N) and N is found by bisection with the exact float
// arithmetic of curand_uniform.
inline uint64_t accept_count(float p, bool strict)
{
    auto acc = [&](uint64_t x) {
        const float u = uniform_from_u32((uint32_t)x);
        return strict ? (u < p) : (u <= p);
    };
    if (!acc(0)) return 0;
    if (acc(0xffffffffull)) return 1ull << 32;
    uint64_t lo = 0, hi = 0xffffffffull;   // acc(lo) true, acc(hi) false
    while (hi - lo > 1) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (acc(mid)) lo = mid; else hi = mid;
    }
    return hi;
}

// Build the kernel's rule from the table.  Indexing of the rule is by b = 4 - a anti-aligned
// neighbours.  At most two distinct non-trivial thresholds exist (T > 0: a = 3, 4; T <= 0: a = 2).
inline AcceptRule make_rule(float temp)
{
    float tab[5];
    accept_table(temp, tab);
    AcceptRule r{};
    uint64_t slots[2] = {0, 0};
    int nslots = 0;
    for (int b = 0; b < 5; b++) {
        const uint64_t n = accept_count(tab[4 - b], false);
        r.always[b] = r.selA[b] = r.selB[b] = 0;
        if (n == (1ull << 32)) { r.always[b] = 0xffffffffu; continue; }
        if (n == 0) continue;   // never
        int s = -1;
        for (int i = 0; i < nslots; i++) if (slots[i] == n) s = i;
        if (s < 0 && nslots < 2) { slots[nslots] = n; s = nslots++; }
        // more than two distinct thresholds cannot happen for this table (see above)
        if (s == 0) r.selA[b] = 0xffffffffu;
        else if (s == 1) r.selB[b] = 0xffffffffu;
    }
    r.thrA = nslots > 0 ? (uint32_t)(slots[0] - 1) : 0u;
    r.thrB = nslots > 1 ? (uint32_t)(slots[1] - 1) : 0u;
    // fast shape: b>=2 always, b==1 -> slot, b==0 -> slot.  The fast kernel hard-wires
    // b==1 -> A and b==0 -> B, so order the slots accordingly.
    r.fast = 0;
    if (r.always[2] && r.always[3] && r.always[4] && !r.always[1] && !r.always[0] && nslots == 2 &&
        (r.selA[0] | r.selB[0]) && (r.selA[1] | r.selB[1])) {
        if (r.selA[0]) {   // b==0 got slot A: swap so that b==1 is A
            const uint32_t t = r.thrA; r.thrA = r.thrB; r.thrB = t;
            for (int b = 0; b < 5; b++) { const uint32_t s = r.selA[b]; r.selA[b] = r.selB[b]; r.selB[b] = s; }
        }
        r.fast = (r.selA[1] && r.selB[0] && !r.selA[0] && !r.selB[1]) ? 1 : 0;
    }
    return r;
}

}  // namespace ising
