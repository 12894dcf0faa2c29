// kernels.cu -- sm_100a kernels of the checkerboard Metropolis hot path.
//
// Replaces (reference, optimized/main.cu): latticeInit_k :92-151, loadTile :380-461,
// spinUpdateV_2D_k :463-670, getMagn_k :702-734.  Not a port: spins are 1 bit each (32 sites
// per word), the source-colour tile and its halo arrive through TMA (cp.async.bulk.tensor)
// boxes whose coordinates carry the periodic wrap, the Philox counter schedule is folded per
// launch/per thread (philox_map.h), and acceptance is an integer compare on the raw Philox
// output that is shifted into bit masks through the carry flag -- no int->float conversion,
// no table lookup, no per-spin extraction.  What is kept is the contract: which Philox
// output decides which site, and the float acceptance table it is compared against.
#include "kernels.cuh"

namespace ising {

// ------------------------------------------------------------------------------------------
// small PTX wrappers (mbarrier + TMA), sm_90+/sm_100a
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *flag has reached `want` (wrap-safe), but never longer than timeout_ns: a neighbour that
// died or hangs must not hang this GPU.  On timeout the error word is set (ising_sync reports it) and
// the caller goes on with whatever the ghost row holds.
__device__ __noinline__ void bounded_wait_counter(const unsigned int* flag, unsigned int want, unsigned int* err_flag,
                                                  unsigned long long timeout_ns)
{
    if ((int)(ld_acquire_sys(flag) - want) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned int polls = 0;
    while ((int)(ld_acquire_sys(flag) - want) < 0) {
        if ((++polls & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
            if (err_flag) atomicExch_system(err_flag, 1u);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Experiment knob: bit q of ISING_ACC_ALU_MASK sends the accumulation of call q of a word to the ALU pipe (predicated
// LOP3 xor) instead of the FMA-heavy pipe (predicated VIADD).
#ifndef ISING_ACC_ALU_MASK
#define ISING_ACC_ALU_MASK 0x00
#endif
__device__ __forceinline__ void xor_bit_if_gt(uint32_t& acc, uint32_t x, uint32_t thr, int bit)
{
    asm("{\n"
        ".reg .pred p;\n"
        "setp.gt.u32 p, %1, %2;\n"
        "@p xor.b32 %0, %0, %3;\n"
        "}\n"
        : "+r"(acc)
        : "r"(x), "r"(thr), "r"(1u << bit));
}
// acc |= (x > thr) << BIT as one compare and one predicated OR-immediate.
__device__ __forceinline__ void or_bit_if_gt(uint32_t& acc, uint32_t x, uint32_t thr, int bit)
{
    if ((ISING_ACC_ALU_MASK >> ((bit >> 1) & 7)) & 1) { xor_bit_if_gt(acc, x, thr, bit); return; }
    asm("{\n"
        ".reg .pred p;\n"
        "setp.gt.u32 p, %1, %2;\n"
        "@p or.b32 %0, %0, %3;\n"
        "}\n"
        : "+r"(acc)
        : "r"(x), "r"(thr), "r"(1u << bit));   // bit is a literal after unrolling
}

// Which multiplies of call q (0..7 inside a word) take the FP64 route of philox_map.h: bit q of
// ISING_DFMA_MASK0 for the M0 multiply, of ISING_DFMA_MASK1 for the M1 multiply.
#ifndef ISING_DFMA_MASK0
#define ISING_DFMA_MASK0 0x00
#endif
#ifndef ISING_DFMA_MASK1
#define ISING_DFMA_MASK1 0x00
#endif
__device__ __forceinline__ void philox_call(int q, const ThreadConsts& tc, uint32_t ta, uint32_t tb, uint32_t td,
                                            const uint32_t (&k0)[7], const uint32_t (&k1)[7], uint32_t out[4])
{
    const bool f0 = (ISING_DFMA_MASK0 >> q) & 1, f1 = (ISING_DFMA_MASK1 >> q) & 1;   // literals after unrolling
    if (f0 && f1)  philox_fast_mix<false, false>(tc, ta, tb, td, k0, k1, out);
    else if (f0)   philox_fast_mix<false, true>(tc, ta, tb, td, k0, k1, out);
    else if (f1)   philox_fast_mix<true, false>(tc, ta, tb, td, k0, k1, out);
    else           philox_fast(tc, ta, tb, td, k0, k1, out);
}

// Fused halo push of an edge tile (called by the CTAs that own the slab's first / last row only): the
// half-warps holding row 0 / Y-1 store their new words a second time into the neighbour's ghost row
// (peer-mapped), make them visible system-wide and count the tile on the neighbour's arrival counter.
// Out of line so that the code and registers of the push do not weigh on the CTAs that never push.
__device__ __noinline__ void push_edge_rows(const UpdateParams& p, int x0, int y0, uint32_t me0, uint32_t me1)
{
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row = y0 + ty;
    const bool top = (row == 0), bot = (row == p.rows - 1);
    if (top) { p.push_up[x0 + tx] = me0; p.push_up[x0 + tx + 16] = me1; }
    if (bot) { p.push_dn[x0 + tx] = me0; p.push_dn[x0 + tx + 16] = me1; }
    if (top || bot) __threadfence_system();
    __syncwarp();
    // rows 0 and Y-1 are written by the half-warps ty == 0 (threads 0..15) and ty == 15 (240..255)
    if (top && tx == 0) atomicAdd_system(p.sig_up, 1u);
    if (bot && tx == 0) atomicAdd_system(p.sig_dn, 1u);
}

// Reject masks of one 32-site word: 8 Philox calls, outputs x,y,z,w -> bits 2q, 16+2q, 2q+1,
// 17+2q.  One compare + one predicated add of the site's bit per threshold (ISETP on the ALU
// pipe, VIADD on the FMA pipe).  Measured alternatives, all bit-exact and none faster (numbers in
// profiles/r01_occupancy_variants.txt): carry-chain accumulation
// (IADD3 + IADD3.X), predicated LOP3, FP32 accumulators on the FMA-lite pipe, a mix of these,
// and compares on the FP64 pipe (DSETP).
template <int NTHR, int J>
__device__ __forceinline__ void word_reject_masks(const ThreadConsts& tc, const StepTables& st, uint32_t thrA,
                                                  uint32_t thrB, uint32_t& rejA, uint32_t& rejB)
{
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int cl = 8 * J + q;
        uint32_t o[4];
        philox_call(q, tc, st.a[cl], st.b[cl], st.d[cl], st.k0, st.k1, o);
        or_bit_if_gt(a, o[0], thrA, 2 * q);
        or_bit_if_gt(a, o[1], thrA, 16 + 2 * q);
        or_bit_if_gt(a, o[2], thrA, 2 * q + 1);
        or_bit_if_gt(a, o[3], thrA, 17 + 2 * q);
        if (NTHR > 1) {
            or_bit_if_gt(b, o[0], thrB, 2 * q);
            or_bit_if_gt(b, o[1], thrB, 16 + 2 * q);
            or_bit_if_gt(b, o[2], thrB, 2 * q + 1);
            or_bit_if_gt(b, o[3], thrB, 17 + 2 * q);
        }
    }
    rejA = a;
    rejB = b;
}

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// New value of one word from its own spins, the four neighbour words and the reject masks.
template <bool FAST>
__device__ __forceinline__ uint32_t updated_word(uint32_t me, uint32_t up, uint32_t dn, uint32_t ct, uint32_t sd,
                                                 uint32_t rejA, uint32_t rejB, const AcceptRule& r)
{
    if (FAST) {
        // Eight hand-picked LOP3 (the compiler needs 11-12 for the pair-sum formulation below, and every ALU
        // instruction is 2 cycles of the binding resource).  With z1 / z2 = both vertical / both
        // horizontal neighbours aligned with the site, and up ^ dn / ct ^ sd = exactly one of a pair anti-aligned:
        //   keep = z2 & ((z1 & rejB) | ((up ^ dn) & rejA))  |  z1 & (ct ^ sd) & rejA      (dE = +8: slot B, +4: slot A)
        // everything else flips: new = me ^ ~keep, straight out of the last LOP3.
        const uint32_t z1 = lop3<0x81>(up, dn, me);            // ~(up ^ me) & ~(dn ^ me)
        const uint32_t z2 = lop3<0x81>(ct, sd, me);
        const uint32_t a1 = lop3<0x28>(up, dn, rejA);          // (up ^ dn) & rejA
        const uint32_t a2 = lop3<0x28>(ct, sd, rejA);
        const uint32_t u = lop3<0xEA>(z1, rejB, a1);           // (z1 & rejB) | a1
        const uint32_t v = z2 & u;
        const uint32_t x = z1 & a2;
        return lop3<0xE1>(me, v, x);                           // ~(me ^ (v | x))
    } else {
        // d_k = neighbour_k ^ me = 1 where anti-aligned; (g,p) and (g2,p2) are the 2-bit sums of pairs
        const uint32_t p = up ^ dn, g = (up ^ me) & (dn ^ me);
        const uint32_t p2 = ct ^ sd, g2 = (ct ^ me) & (sd ^ me);
        const uint32_t c = p & p2;
        const uint32_t b0 = p ^ p2, b1 = g ^ g2 ^ c, b2 = g & g2;   // b = number anti-aligned
        const uint32_t accA = ~rejA, accB = ~rejB;
        uint32_t cls[5];
        cls[0] = ~(b0 | b1 | b2); cls[1] = b0 & ~b1; cls[2] = ~b0 & b1; cls[3] = b0 & b1; cls[4] = b2;
        uint32_t f = 0;
#pragma unroll
        for (int b = 0; b < 5; b++) f |= cls[b] & (r.always[b] | (r.selA[b] & accA) | (r.selB[b] & accB));
        return me ^ f;
    }
}

// ------------------------------------------------------------------------------------------
// Colour half-sweep.  One CTA = one 16 x 32-word tile, one thread = one reference thread id
// (words tx and tx+16 of row ty).  Five TMA boxes stage the source colour: centre, the row
// above, the row below and a 16-byte box left and right; their coordinates carry the periodic
// wrap.  The tile's own words are private to their thread: loaded straight into registers at
// kernel entry and stored straight back, so that warps never meet at a CTA barrier after the
// prologue (a closing __syncthreads for a TMA store cost 27 % of warp time in barrier stalls).
// ------------------------------------------------------------------------------------------
#ifndef ISING_MIN_BLOCKS
#define ISING_MIN_BLOCKS 4
#endif
template <bool FAST, bool BONDS, bool PUSH>
__global__ void __launch_bounds__(UPDATE_THREADS, ISING_MIN_BLOCKS)
update_color_kernel(const __grid_constant__ CUtensorMap src_tile, const __grid_constant__ CUtensorMap src_row,
                    const __grid_constant__ CUtensorMap src_side, uint32_t* __restrict__ dst /* local row 0 */,
                    const uint32_t* __restrict__ bonds /* 4 planes: up, down, left, right; local row 0 */,
                    const __grid_constant__ UpdateParams p)
{
    __shared__ alignas(128) uint32_t sV[(TH + 2) * TW];   // rows: above, 16 centre rows, below
    __shared__ alignas(128) uint32_t sL[TH * SIDE_W];
    __shared__ alignas(128) uint32_t sR[TH * SIDE_W];
    __shared__ alignas(8) uint64_t bar;

    // tile-row index of this CTA: (y, z) of the grid, because gridDim.y alone stops at 65535
    const int by = (int)(blockIdx.z * gridDim.y + blockIdx.y);
    if (by >= p.launch_tile_rows) return;
    int tile_row;
    if (p.edges_mode) tile_row = (by == 0) ? p.tile_row0 : p.tile_row_alt;
    else if (PUSH)    tile_row = (by == 0) ? 0 : ((by == 1) ? p.launch_tile_rows - 1 : by - 1);
    else              tile_row = p.tile_row0 + by;
    const int x0 = blockIdx.x * TW;       // first word of the tile in the row
    const int y0 = tile_row * TH;         // first local row of the tile
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    // own spins: two 32-site words per thread, in flight while the random numbers are made
    uint32_t* const my = dst + (size_t)(y0 + ty) * p.pitch + x0 + tx;
    uint32_t me[2];
    // +-J bonds of the thread's sites (main.cu:575-618): one bit plane per direction
    uint32_t jb[2][4];
    if (BONDS) {
        const size_t plane = (size_t)p.bond_rows * p.pitch, off = (size_t)(y0 + ty) * p.pitch + x0 + tx;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            jb[0][d] = bonds[d * plane + off];
            jb[1][d] = bonds[d * plane + off + 16];
        }
    }

    // No CTA-wide barrier: warp 0 initialises the mbarrier and then *arrives* on one named barrier
    // per other warp (ids 1..7, 64 threads each); warp w *syncs* on barrier w only just before it
    // first touches the mbarrier, ~1000 instructions later, when warp 0 has long since arrived.
    // (A __syncthreads here held 11 % of all warp-stall samples: every warp waited for warp 0's
    // mbarrier fence, which waited for warp 0's global loads.)
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __threadfence_block();
        __syncwarp();
#pragma unroll
        for (int w = 1; w < UPDATE_THREADS / 32; w++) asm volatile("barrier.cta.arrive %0, 64;" ::"r"(w) : "memory");
    }
    me[0] = my[0];     // issued after the mbarrier fence: a fence behind them would wait for these loads
    me[1] = my[16];
    if (threadIdx.x == 0) {
        if (PUSH) {   // ghost rows of the source colour must have arrived from the neighbours
            if (y0 == 0) bounded_wait_counter(p.wait_top, p.wait_value, p.err_flag, p.spin_timeout_ns);
            if (y0 + TH == p.rows) bounded_wait_counter(p.wait_bot, p.wait_value, p.err_flag, p.spin_timeout_ns);
            asm volatile("fence.proxy.async;" ::: "memory");   // later TMA reads see the peer's stores
        }
        int ytop, ybot;
        if (p.yperiod > 0) {
            ytop = (y0 % p.yperiod == 0) ? y0 + p.yperiod - 1 : y0 - 1;
            ybot = ((y0 + TH) % p.yperiod == 0) ? y0 + TH - p.yperiod : y0 + TH;
        } else {
            ytop = y0 - 1;    // ghost row above the slab when y0 == 0
            ybot = y0 + TH;   // ghost row below the slab when y0 + TH == rows
        }
        const int xl = (x0 % p.xperiod == 0) ? x0 + p.xperiod - 1 : x0 - 1;
        const int xr = ((x0 + TW) % p.xperiod == 0) ? x0 + TW - p.xperiod : x0 + TW;
        constexpr uint32_t bytes = ((TH + 2) * TW + 2 * TH * SIDE_W) * 4;
        mbar_expect_tx(&bar, bytes);
        // allocation row = local row + 1 (row 0 of the allocation is the upper ghost row)
        tma_load_2d(sV + TW, &src_tile, x0, y0 + 1, &bar);
        tma_load_2d(sV, &src_row, x0, ytop + 1, &bar);
        tma_load_2d(sV + (TH + 1) * TW, &src_row, x0, ybot + 1, &bar);
        tma_load_2d(sL, &src_side, xl - (SIDE_W - 1), y0 + 1, &bar);
        tma_load_2d(sR, &src_side, xr, y0 + 1, &bar);
    }

    const uint32_t grow = p.grow0 + (uint32_t)(y0 + ty);
    const uint32_t tid = ref_tid(grow, (uint32_t)(x0 + tx), p.grid_x);
    const ThreadConsts tc = make_thread_consts(tid, p.st.x1, p.st.k1r3);
    // side neighbour is at colour-column m-1 ("readBack", main.cu:542) or m+1
    const bool read_back = (p.color == 0) ? !(grow & 1u) : (grow & 1u);

    uint32_t rejA[2], rejB[2];
    if (FAST || p.rule.selB[0] | p.rule.selB[1] | p.rule.selB[2] | p.rule.selB[3] | p.rule.selB[4]) {
        word_reject_masks<2, 0>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
        word_reject_masks<2, 1>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
    } else {
        word_reject_masks<1, 0>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
        word_reject_masks<1, 1>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
    }

    if (threadIdx.x >= 32) asm volatile("barrier.cta.sync %0, 64;" ::"r"((int)(threadIdx.x >> 5)) : "memory");
    mbar_wait(&bar, 0);

#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int w = tx + 16 * j;
        uint32_t up = sV[ty * TW + w];
        uint32_t ct = sV[(ty + 1) * TW + w];
        uint32_t dn = sV[(ty + 2) * TW + w];
        uint32_t sd;
        if (read_back) {
            const uint32_t prev = (w == 0) ? sL[ty * SIDE_W + SIDE_W - 1] : sV[(ty + 1) * TW + w - 1];
            sd = __funnelshift_l(prev, ct, 1);    // (ct << 1) | (prev >> 31)
        } else {
            const uint32_t next = (w == TW - 1) ? sR[ty * SIDE_W] : sV[(ty + 1) * TW + w + 1];
            sd = __funnelshift_r(ct, next, 1);    // (ct >> 1) | (next << 31)
        }
        if (BONDS) {   // an antiferromagnetic bond flips the neighbour's sign as seen by this site
            up ^= jb[j][0];
            dn ^= jb[j][1];
            if (read_back) { sd ^= jb[j][2]; ct ^= jb[j][3]; }   // sd = left neighbour, ct = right
            else           { ct ^= jb[j][2]; sd ^= jb[j][3]; }   // ct = left neighbour, sd = right
        }
        me[j] = updated_word<FAST>(me[j], up, dn, ct, sd, rejA[j], rejB[j], p.rule);
        my[16 * j] = me[j];
    }
    if (PUSH && (y0 == 0 || y0 + TH == p.rows)) {   // CTA-uniform: only the two edge tile rows push
        push_edge_rows(p, x0, y0, me[0], me[1]);
    }
}

// ------------------------------------------------------------------------------------------
// Same colour half-sweep, two reference thread ids per thread: one CTA = 32 rows x 32 words (two stacked tiles
// of the kernel above), thread (tx, ty) owns words tx, tx+16 of rows ty AND ty+16.  What a CTA pays once --
// parameter and table loads (24 LDCU), barrier set-up and hand-shake, the TMA issue path of thread 0, tile
// addressing, entry / exit -- is now spread over 128 sites per thread instead of 64, and the two Philox streams
// of a call interleave (more independent work per warp between dependent multiplies).  Used for every launch
// that covers an even number of 16-row tile rows without bonds and without the fused push (the interior of a
// slab, or a whole single-slab lattice); everything else takes the one-tile kernel.
// ------------------------------------------------------------------------------------------
constexpr int TH2 = 2 * TH;

// Reject masks of word J of BOTH thread ids of a thread (tc0: row ty, tc1: row ty + 16): the table values of a call
// are fetched once for the two streams.
template <int NTHR, int J>
__device__ __forceinline__ void word_reject_masks2(const ThreadConsts& tc0, const ThreadConsts& tc1, const StepTables& st,
                                                   uint32_t thrA, uint32_t thrB, uint32_t (&rejA)[2], uint32_t (&rejB)[2])
{
    uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int cl = 8 * J + q;
        uint32_t o[4], u[4];
        philox_fast(tc0, st.a[cl], st.b[cl], st.d[cl], st.k0, st.k1, o);
        philox_fast(tc1, st.a[cl], st.b[cl], st.d[cl], st.k0, st.k1, u);
        or_bit_if_gt(a0, o[0], thrA, 2 * q);
        or_bit_if_gt(a0, o[1], thrA, 16 + 2 * q);
        or_bit_if_gt(a0, o[2], thrA, 2 * q + 1);
        or_bit_if_gt(a0, o[3], thrA, 17 + 2 * q);
        or_bit_if_gt(a1, u[0], thrA, 2 * q);
        or_bit_if_gt(a1, u[1], thrA, 16 + 2 * q);
        or_bit_if_gt(a1, u[2], thrA, 2 * q + 1);
        or_bit_if_gt(a1, u[3], thrA, 17 + 2 * q);
        if (NTHR > 1) {
            or_bit_if_gt(b0, o[0], thrB, 2 * q);
            or_bit_if_gt(b0, o[1], thrB, 16 + 2 * q);
            or_bit_if_gt(b0, o[2], thrB, 2 * q + 1);
            or_bit_if_gt(b0, o[3], thrB, 17 + 2 * q);
            or_bit_if_gt(b1, u[0], thrB, 2 * q);
            or_bit_if_gt(b1, u[1], thrB, 16 + 2 * q);
            or_bit_if_gt(b1, u[2], thrB, 2 * q + 1);
            or_bit_if_gt(b1, u[3], thrB, 17 + 2 * q);
        }
    }
    rejA[0] = a0; rejA[1] = a1;
    rejB[0] = b0; rejB[1] = b1;
}

template <bool FAST>
__global__ void __launch_bounds__(UPDATE_THREADS, ISING_MIN_BLOCKS)
update_color_kernel2(const __grid_constant__ CUtensorMap src_tile2 /* box 32 x 32 */,
                     const __grid_constant__ CUtensorMap src_row /* box 32 x 1 */,
                     const __grid_constant__ CUtensorMap src_side2 /* box 4 x 32 */, uint32_t* __restrict__ dst /* local row 0 */,
                     const __grid_constant__ UpdateParams p)
{
    __shared__ alignas(128) uint32_t sV[(TH2 + 2) * TW];   // rows: above, 32 centre rows, below
    __shared__ alignas(128) uint32_t sL[TH2 * SIDE_W];
    __shared__ alignas(128) uint32_t sR[TH2 * SIDE_W];
    __shared__ alignas(8) uint64_t bar;

    const int by = (int)(blockIdx.z * gridDim.y + blockIdx.y);   // 32-row unit of this launch
    if (by >= p.launch_tile_rows) return;
    const int x0 = blockIdx.x * TW;                     // first word of the tile in the row
    const int y0 = p.tile_row0 * TH + by * TH2;         // first local row of the tile (tile_row0 counts 16-row units)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    uint32_t* const my = dst + (size_t)(y0 + ty) * p.pitch + x0 + tx;
    const size_t lower = (size_t)TH * p.pitch;          // same column, 16 rows down
    uint32_t me[2][2];                                  // [row ty / ty + 16][word tx / tx + 16]

    if (threadIdx.x < 32) {                             // see update_color_kernel: no CTA-wide barrier
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __threadfence_block();
        __syncwarp();
#pragma unroll
        for (int w = 1; w < UPDATE_THREADS / 32; w++) asm volatile("barrier.cta.arrive %0, 64;" ::"r"(w) : "memory");
    }
    me[0][0] = my[0];
    me[0][1] = my[16];
    me[1][0] = my[lower];
    me[1][1] = my[lower + 16];
    if (threadIdx.x < 32) {
        // the four wrap tests on four lanes: one division for the warp
        const int lane = threadIdx.x;
        unsigned int n = 0, d = 1;
        if (lane == 0) { n = (unsigned)x0; d = (unsigned)p.xperiod; }
        if (lane == 1) { n = (unsigned)(x0 + TW); d = (unsigned)p.xperiod; }
        if (p.yperiod > 0) {
            if (lane == 2) { n = (unsigned)y0; d = (unsigned)p.yperiod; }
            if (lane == 3) { n = (unsigned)(y0 + TH2); d = (unsigned)p.yperiod; }
        }
        const unsigned int wrap = __ballot_sync(0xffffffffu, (n % d) == 0u);
        if (lane == 0) {
            int ytop = y0 - 1, ybot = y0 + TH2;         // ghost rows when yperiod == 0
            if (p.yperiod > 0) {
                if (wrap & 4u) ytop = y0 + p.yperiod - 1;
                if (wrap & 8u) ybot = y0 + TH2 - p.yperiod;
            }
            const int xl = (wrap & 1u) ? x0 + p.xperiod - 1 : x0 - 1;
            const int xr = (wrap & 2u) ? x0 + TW - p.xperiod : x0 + TW;
            constexpr uint32_t bytes = ((TH2 + 2) * TW + 2 * TH2 * SIDE_W) * 4;
            mbar_expect_tx(&bar, bytes);
            // allocation row = local row + 1 (row 0 of the allocation is the upper ghost row)
            tma_load_2d(sV + TW, &src_tile2, x0, y0 + 1, &bar);
            tma_load_2d(sV, &src_row, x0, ytop + 1, &bar);
            tma_load_2d(sV + (TH2 + 1) * TW, &src_row, x0, ybot + 1, &bar);
            tma_load_2d(sL, &src_side2, xl - (SIDE_W - 1), y0 + 1, &bar);
            tma_load_2d(sR, &src_side2, xr, y0 + 1, &bar);
        }
    }

    const uint32_t grow = p.grow0 + (uint32_t)(y0 + ty);
    const uint32_t tid0 = ref_tid(grow, (uint32_t)(x0 + tx), p.grid_x);
    const uint32_t tid1 = tid0 + p.grid_x * 256u;       // 16 rows down = the next block row of the reference grid
    const ThreadConsts tc0 = make_thread_consts(tid0, p.st.x1, p.st.k1r3);
    const ThreadConsts tc1 = make_thread_consts(tid1, p.st.x1, p.st.k1r3);
    // side neighbour is at colour-column m-1 ("readBack", main.cu:542) or m+1; rows ty and ty + 16 have one parity
    const bool read_back = (p.color == 0) ? !(grow & 1u) : (grow & 1u);

    uint32_t rejA[2][2], rejB[2][2];                    // [word][row]
    if (FAST || p.rule.selB[0] | p.rule.selB[1] | p.rule.selB[2] | p.rule.selB[3] | p.rule.selB[4]) {
        word_reject_masks2<2, 0>(tc0, tc1, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
        word_reject_masks2<2, 1>(tc0, tc1, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
    } else {
        word_reject_masks2<1, 0>(tc0, tc1, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
        word_reject_masks2<1, 1>(tc0, tc1, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
    }

    if (threadIdx.x >= 32) asm volatile("barrier.cta.sync %0, 64;" ::"r"((int)(threadIdx.x >> 5)) : "memory");
    mbar_wait(&bar, 0);

#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int r = ty + TH * t;                      // row inside the 32-row tile
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = tx + 16 * j;
            const uint32_t up = sV[r * TW + w];
            const uint32_t ct = sV[(r + 1) * TW + w];
            const uint32_t dn = sV[(r + 2) * TW + w];
            uint32_t sd;
            if (read_back) {
                const uint32_t prev = (w == 0) ? sL[r * SIDE_W + SIDE_W - 1] : sV[(r + 1) * TW + w - 1];
                sd = __funnelshift_l(prev, ct, 1);    // (ct << 1) | (prev >> 31)
            } else {
                const uint32_t next = (w == TW - 1) ? sR[r * SIDE_W] : sV[(r + 1) * TW + w + 1];
                sd = __funnelshift_r(ct, next, 1);    // (ct >> 1) | (next << 31)
            }
            me[t][j] = updated_word<FAST>(me[t][j], up, dn, ct, sd, rejA[j][t], rejB[j][t], p.rule);
            my[t * lower + 16 * j] = me[t][j];
        }
    }
}

// Wait until the neighbours' fused halo pushes have all arrived: flags[0], flags[2] (top / bottom ghost row
// of colour 0) must have reached want0, flags[1], flags[3] want1.  Put on a stream ahead of anything that
// overwrites the ghost rows (host upload of the next replica).
__global__ void halo_wait_kernel(const unsigned int* flags, unsigned int want0, unsigned int want1,
                                 unsigned int* err_flag, unsigned long long timeout_ns)
{
    if (threadIdx.x < 4) {
        const unsigned int want = (threadIdx.x & 1) ? want1 : want0;
        bounded_wait_counter(flags + threadIdx.x, want, err_flag, timeout_ns);
    }
}

void launch_halo_wait(const unsigned int* flags, unsigned int want0, unsigned int want1, unsigned int* err_flag,
                      unsigned long long timeout_ns, cudaStream_t stream)
{
    halo_wait_kernel<<<1, 32, 0, stream>>>(flags, want0, want1, err_flag, timeout_ns);
}


// ------------------------------------------------------------------------------------------
// Persistent colour half-sweep (the product kernel).  Same tile, same thread <-> reference-thread-id
// map and same arithmetic as update_color_kernel, but one CTA per resident slot (SMs x 4) walks over
// the tiles  blockIdx.x + k * gridDim.x  of the launch:
//   * everything that is per launch or per thread (parameters, read_back, the thread's offset in a
//     tile, barrier set-up) is paid once per CTA lifetime instead of once per 64 sites x 256 threads;
//     the old kernel spent ~27 % of its stall samples in its first 31 instructions;
//   * the tiles (source colour + halo + the tile being updated) arrive through an 8-stage shared-memory
//     ring: one lane of warp k mod 8 issues the six TMA boxes of tile k+4 at the top of iteration k
//     (full[] / empty[] mbarriers, no CTA-wide barrier in the loop, warps drift freely by up to four tiles);
//   * with edges_first the two edge tile rows of a slab are the first tiles of the walk, so the fused
//     halo push (PUSH) leaves in the first iteration of CTAs 0 .. 2*tiles_x-1 of the SAME launch that
//     does the interior -- no separate 64-CTA edge launch in front of the interior any more.
// ------------------------------------------------------------------------------------------
struct alignas(128) TileStage {
    uint32_t v[(TH + 2) * TW];   // source colour, rows: above, 16 centre rows, below
    uint32_t l[TH * SIDE_W];
    uint32_t r[TH * SIDE_W];
    uint32_t own[TH * TW];       // the tile being updated (a global load in the loop was sunk next to its use
                                 // by the compiler and showed up as long-scoreboard stalls)
};

__device__ __forceinline__ int sched_tile_row(int r, const UpdateParams& p)
{
    if (p.edges_first) return (r == 0) ? p.tile_row0 : ((r == 1) ? p.tile_row_alt : r - 1);
    return p.tile_row0 + r;
}

// Warp 0, converged: arm full[stage] and issue the five TMA boxes of the tile (tile_row, col).
// Lanes 0..3 each take one of the four wrap tests so that the warp pays for one division.
template <bool PUSH>
__device__ __forceinline__ void produce_tile(TileStage* st, uint64_t* full, const CUtensorMap* src_tile,
                                             const CUtensorMap* src_row, const CUtensorMap* src_side,
                                             const CUtensorMap* dst_tile, const UpdateParams& p, int tile_row,
                                             int col, int lane)
{
    const int x0 = col * TW, y0 = tile_row * TH;
    unsigned int n = 0, d = 1;
    if (lane == 0) { n = (unsigned)x0; d = (unsigned)p.xperiod; }
    if (lane == 1) { n = (unsigned)(x0 + TW); d = (unsigned)p.xperiod; }
    if (p.yperiod > 0) {
        if (lane == 2) { n = (unsigned)y0; d = (unsigned)p.yperiod; }
        if (lane == 3) { n = (unsigned)(y0 + TH); d = (unsigned)p.yperiod; }
    }
    const unsigned int wrap = __ballot_sync(0xffffffffu, (n % d) == 0u);
    if (lane == 0) {
        if (PUSH) {   // ghost rows of the source colour must have arrived from the neighbours
            if (y0 == 0) bounded_wait_counter(p.wait_top, p.wait_value, p.err_flag, p.spin_timeout_ns);
            if (y0 + TH == p.rows) bounded_wait_counter(p.wait_bot, p.wait_value, p.err_flag, p.spin_timeout_ns);
            asm volatile("fence.proxy.async;" ::: "memory");   // later TMA reads see the peer's stores
        }
        int ytop = y0 - 1, ybot = y0 + TH;     // ghost rows when yperiod == 0
        if (p.yperiod > 0) {
            if (wrap & 4u) ytop = y0 + p.yperiod - 1;
            if (wrap & 8u) ybot = y0 + TH - p.yperiod;
        }
        const int xl = (wrap & 1u) ? x0 + p.xperiod - 1 : x0 - 1;
        const int xr = (wrap & 2u) ? x0 + TW - p.xperiod : x0 + TW;
        constexpr uint32_t bytes = ((TH + 2) * TW + 2 * TH * SIDE_W + TH * TW) * 4;
        mbar_expect_tx(full, bytes);
        // allocation row = local row + 1 (row 0 of the allocation is the upper ghost row)
        tma_load_2d(st->v + TW, src_tile, x0, y0 + 1, full);
        tma_load_2d(st->own, dst_tile, x0, y0 + 1, full);
        tma_load_2d(st->v, src_row, x0, ytop + 1, full);
        tma_load_2d(st->v + (TH + 1) * TW, src_row, x0, ybot + 1, full);
        tma_load_2d(st->l, src_side, xl - (SIDE_W - 1), y0 + 1, full);
        tma_load_2d(st->r, src_side, xr, y0 + 1, full);
    }
    __syncwarp();
}

template <bool FAST, bool BONDS, bool PUSH>
__global__ void __launch_bounds__(UPDATE_THREADS, ISING_MIN_BLOCKS)
update_color_persistent(const __grid_constant__ CUtensorMap src_tile, const __grid_constant__ CUtensorMap src_row,
                        const __grid_constant__ CUtensorMap src_side, const __grid_constant__ CUtensorMap dst_tile,
                        uint32_t* __restrict__ dst /* local row 0 */,
                        const uint32_t* __restrict__ bonds /* 4 planes: up, down, left, right; local row 0 */,
                        const __grid_constant__ UpdateParams p)
{
    __shared__ TileStage stage[UPDATE_STAGES];
    __shared__ alignas(8) uint64_t full[UPDATE_STAGES];    // TMA bytes of the stage have landed
    __shared__ alignas(8) uint64_t empty[UPDATE_STAGES];   // all 8 warps have read the stage

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = (int)gridDim.x;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < UPDATE_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], UPDATE_THREADS / 32); }
        fence_mbar_init();
    }
    __syncthreads();   // once per CTA lifetime

    // consumer walk (tile of iteration k) and producer walk (tile of iteration k + UPDATE_PREFETCH)
    int t = (int)blockIdx.x, r = t / p.tiles_x, col = t - r * p.tiles_x;
    int tp = t, rp = r, colp = col;
#pragma unroll
    for (int i = 0; i < UPDATE_PREFETCH; i++) {   // the first tiles: one warp each
        if (tp < p.n_tiles && warp == i)
            produce_tile<PUSH>(&stage[i], &full[i], &src_tile, &src_row, &src_side, &dst_tile, p, sched_tile_row(rp, p), colp, lane);
        tp += G; rp += p.step_rows; colp += p.step_cols;
        if (colp >= p.tiles_x) { colp -= p.tiles_x; rp++; }
    }

    // per-thread constants of the walk
    const uint32_t tid_thread = threadIdx.x;                         // (grow % 16) * 16 + c % 16
    const uint32_t tile_row_g0 = p.grow0 / (uint32_t)TH;             // global tile row of local tile row 0
    // side neighbour is at colour-column m-1 ("readBack", main.cu:542) or m+1; grow0 and y0 are even
    const bool read_back = (p.color == 0) ? !(ty & 1) : (ty & 1);
    const size_t thread_off = (size_t)ty * p.pitch + tx;
    const bool need_b = FAST || (p.rule.selB[0] | p.rule.selB[1] | p.rule.selB[2] | p.rule.selB[3] | p.rule.selB[4]);

#pragma unroll 1
    for (int k = 0; t < p.n_tiles; k++) {
        const int s = k & (UPDATE_STAGES - 1);
        // ---- producer: tile of iteration k + UPDATE_PREFETCH into the stage read UPDATE_STAGES - UPDATE_PREFETCH
        // iterations ago.  The job rotates over the warps (warp k mod 8 in iteration k): a fixed producer warp
        // executes ~15 % more instructions than the others, falls behind, and the other seven then spin on
        // full[] -- measured: 27 % of all stall samples, 34 polls per wait.
        if (tp < p.n_tiles && warp == (k & (UPDATE_THREADS / 32 - 1))) {
            const int kp = k + UPDATE_PREFETCH, sp = kp & (UPDATE_STAGES - 1);
            if (kp >= UPDATE_STAGES) mbar_wait(&empty[sp], ((kp / UPDATE_STAGES) + 1) & 1);
            produce_tile<PUSH>(&stage[sp], &full[sp], &src_tile, &src_row, &src_side, &dst_tile, p, sched_tile_row(rp, p), colp, lane);
        }
        tp += G; rp += p.step_rows; colp += p.step_cols;
        if (colp >= p.tiles_x) { colp -= p.tiles_x; rp++; }

        // ---- consumer
        const int tile_row = sched_tile_row(r, p);
        const int x0 = col * TW, y0 = tile_row * TH;
        uint32_t* const my = dst + (size_t)y0 * p.pitch + x0 + thread_off;
        uint32_t me[2];
        uint32_t jb[2][4];
        if (BONDS) {
            const size_t plane = (size_t)p.bond_rows * p.pitch, off = (size_t)y0 * p.pitch + x0 + thread_off;
#pragma unroll
            for (int d = 0; d < 4; d++) {
                jb[0][d] = bonds[d * plane + off];
                jb[1][d] = bonds[d * plane + off + 16];
            }
        }
        const uint32_t tid = ((tile_row_g0 + (uint32_t)tile_row) * p.grid_x + (uint32_t)col) * 256u + tid_thread;
        const ThreadConsts tc = make_thread_consts(tid, p.st.x1, p.st.k1r3);

        uint32_t rejA[2], rejB[2];
        if (need_b) {
            word_reject_masks<2, 0>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
            word_reject_masks<2, 1>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
        } else {
            word_reject_masks<1, 0>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[0], rejB[0]);
            word_reject_masks<1, 1>(tc, p.st, p.rule.thrA, p.rule.thrB, rejA[1], rejB[1]);
        }

        mbar_wait(&full[s], (k / UPDATE_STAGES) & 1);
        const TileStage& st = stage[s];
        uint32_t up[2], ct[2], dn[2], sd[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int w = tx + 16 * j;
            me[j] = st.own[ty * TW + w];
            up[j] = st.v[ty * TW + w];
            ct[j] = st.v[(ty + 1) * TW + w];
            dn[j] = st.v[(ty + 2) * TW + w];
            if (read_back) {
                const uint32_t prev = (w == 0) ? st.l[ty * SIDE_W + SIDE_W - 1] : st.v[(ty + 1) * TW + w - 1];
                sd[j] = __funnelshift_l(prev, ct[j], 1);    // (ct << 1) | (prev >> 31)
            } else {
                const uint32_t next = (w == TW - 1) ? st.r[ty * SIDE_W] : st.v[(ty + 1) * TW + w + 1];
                sd[j] = __funnelshift_r(ct[j], next, 1);    // (ct >> 1) | (next << 31)
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // the stage is in registers: the producer may refill it
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (BONDS) {   // an antiferromagnetic bond flips the neighbour's sign as seen by this site
                up[j] ^= jb[j][0];
                dn[j] ^= jb[j][1];
                if (read_back) { sd[j] ^= jb[j][2]; ct[j] ^= jb[j][3]; }   // sd = left neighbour, ct = right
                else           { ct[j] ^= jb[j][2]; sd[j] ^= jb[j][3]; }   // ct = left neighbour, sd = right
            }
            me[j] = updated_word<FAST>(me[j], up[j], dn[j], ct[j], sd[j], rejA[j], rejB[j], p.rule);
            my[16 * j] = me[j];
        }
        if (PUSH && (y0 == 0 || y0 + TH == p.rows)) {   // CTA-uniform: only the two edge tile rows push
            push_edge_rows(p, x0, y0, me[0], me[1]);
        }

        t += G; r += p.step_rows; col += p.step_cols;
        if (col >= p.tiles_x) { col -= p.tiles_x; r++; }
    }
}

void launch_update_persistent(const CUtensorMap& src_tile, const CUtensorMap& src_row, const CUtensorMap& src_side,
                              const CUtensorMap& dst_tile, uint32_t* dst, const uint32_t* bonds, UpdateParams p,
                              int tiles_x, int tile_rows, int sms, cudaStream_t stream)
{
    const long long n_tiles = (long long)tiles_x * tile_rows;
    long long grid = (long long)sms * ISING_MIN_BLOCKS;
    if (grid > n_tiles) grid = n_tiles;
    p.tiles_x = tiles_x;
    p.n_tiles = (int)n_tiles;
    p.launch_tile_rows = tile_rows;
    p.step_rows = (int)(grid / tiles_x);
    p.step_cols = (int)(grid % tiles_x);
#define ISING_LAUNCH_P(F, B, P) \
    update_color_persistent<F, B, P><<<(unsigned)grid, UPDATE_THREADS, 0, stream>>>(src_tile, src_row, src_side, dst_tile, dst, bonds, p)
    const int sel = (p.rule.fast ? 4 : 0) | (bonds ? 2 : 0) | (p.push ? 1 : 0);
    switch (sel) {
        case 0: ISING_LAUNCH_P(false, false, false); break;
        case 1: ISING_LAUNCH_P(false, false, true); break;
        case 2: ISING_LAUNCH_P(false, true, false); break;
        case 3: ISING_LAUNCH_P(false, true, true); break;
        case 4: ISING_LAUNCH_P(true, false, false); break;
        case 5: ISING_LAUNCH_P(true, false, true); break;
        case 6: ISING_LAUNCH_P(true, true, false); break;
        default: ISING_LAUNCH_P(true, true, true); break;
    }
#undef ISING_LAUNCH_P
}

// (tiles_x, tile rows) as a 3-D grid: rows beyond 32768 spill into z
static dim3 tile_grid(int tiles_x, int tile_rows)
{
    const int gy = tile_rows < 32768 ? tile_rows : 32768;
    return dim3((unsigned)tiles_x, (unsigned)gy, (unsigned)((tile_rows + gy - 1) / gy));
}

void launch_update(const CUtensorMap& src_tile, const CUtensorMap& src_row, const CUtensorMap& src_side,
                   uint32_t* dst, const uint32_t* bonds, const UpdateParams& p, int tiles_x, int tiles_y,
                   cudaStream_t stream)
{
    const dim3 grid = tile_grid(tiles_x, tiles_y);
#define ISING_LAUNCH(F, B, P) \
    update_color_kernel<F, B, P><<<grid, UPDATE_THREADS, 0, stream>>>(src_tile, src_row, src_side, dst, bonds, p)
    const int sel = (p.rule.fast ? 4 : 0) | (bonds ? 2 : 0) | (p.push ? 1 : 0);
    switch (sel) {
        case 0: ISING_LAUNCH(false, false, false); break;
        case 1: ISING_LAUNCH(false, false, true); break;
        case 2: ISING_LAUNCH(false, true, false); break;
        case 3: ISING_LAUNCH(false, true, true); break;
        case 4: ISING_LAUNCH(true, false, false); break;
        case 5: ISING_LAUNCH(true, false, true); break;
        case 6: ISING_LAUNCH(true, true, false); break;
        default: ISING_LAUNCH(true, true, true); break;
    }
#undef ISING_LAUNCH
}

void launch_update2(const CUtensorMap& src_tile2, const CUtensorMap& src_row, const CUtensorMap& src_side2, uint32_t* dst,
                    const UpdateParams& p, int tiles_x, int tile_rows32, cudaStream_t stream)
{
    const dim3 grid = tile_grid(tiles_x, tile_rows32);
    if (p.rule.fast) update_color_kernel2<true><<<grid, UPDATE_THREADS, 0, stream>>>(src_tile2, src_row, src_side2, dst, p);
    else             update_color_kernel2<false><<<grid, UPDATE_THREADS, 0, stream>>>(src_tile2, src_row, src_side2, dst, p);
}

const char* update_kernel_name(int fast, int variant)
{
    switch (variant) {
        case 1:
            return fast ? "update_color_kernel<fast: 2 integer thresholds, TMA 16x32-word tiles, 1 bit/spin>"
                        : "update_color_kernel<generic acceptance rule, TMA 16x32-word tiles, 1 bit/spin>";
        case 2:
            return fast ? "update_color_persistent<fast: 2 integer thresholds, 8-stage TMA ring of 16x32-word tiles, 1 bit/spin>"
                        : "update_color_persistent<generic acceptance rule, 8-stage TMA ring of 16x32-word tiles, 1 bit/spin>";
        default:
            return fast ? "update_color_kernel2<fast: 2 integer thresholds, TMA 32x32-word tiles, 2 Philox streams per thread, 1 bit/spin>"
                        : "update_color_kernel2<generic acceptance rule, TMA 32x32-word tiles, 2 Philox streams per thread, 1 bit/spin>";
    }
}

// ------------------------------------------------------------------------------------------
// Lattice init (latticeInit_k): spin = (uniform < 0.5f) <=> x <= thr_half.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UPDATE_THREADS) init_kernel(uint32_t* __restrict__ plane /* local row 0 */,
                                                             const __grid_constant__ InitParams p)
{
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * TW;
    const int by = (int)(blockIdx.z * gridDim.y + blockIdx.y);
    if (by >= p.rows / TH) return;
    const int row = by * TH + ty;
    const uint32_t grow = p.grow0 + (uint32_t)row;
    const uint32_t tid = ref_tid(grow, (uint32_t)(x0 + tx), p.grid_x);
    const ThreadConsts tc = make_thread_consts(tid, p.st.x1, p.st.k1r3);
    uint32_t rej0, rej1, unused;
    word_reject_masks<1, 0>(tc, p.st, p.thr_half, 0u, rej0, unused);
    word_reject_masks<1, 1>(tc, p.st, p.thr_half, 0u, rej1, unused);
    uint32_t* r = plane + (size_t)row * p.pitch + x0 + tx;
    r[0] = ~rej0;
    r[16] = ~rej1;
}

void launch_init(uint32_t* plane, const InitParams& p, cudaStream_t stream)
{
    init_kernel<<<tile_grid(p.pitch / TW, p.rows / TH), UPDATE_THREADS, 0, stream>>>(plane, p);
}

// ------------------------------------------------------------------------------------------
// Observables
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v)
{
    __shared__ unsigned long long sh[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;
}

__global__ void __launch_bounds__(256) count_kernel(const uint4* __restrict__ plane, size_t n4,
                                                   unsigned long long* __restrict__ counts)
{
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = plane[i];
        acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    acc = block_sum_u64(acc);
    if (threadIdx.x == 0) atomicAdd(counts, acc);
}

void launch_count(const uint32_t* plane, int rows, int pitch, unsigned long long* counts, cudaStream_t stream)
{
    const size_t n4 = (size_t)rows * pitch / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    count_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(plane), n4, counts);
}

// Number of unequal nearest-neighbour pairs: every bond joins one black and one white site, so
// visiting the four white neighbours of every black site counts each bond once.
__global__ void __launch_bounds__(256) bond_differ_kernel(const uint32_t* __restrict__ black,
                                                         const uint32_t* __restrict__ white, int rows, int pitch,
                                                         int yperiod, int xperiod, uint32_t grow0,
                                                         unsigned long long* __restrict__ counts)
{
    unsigned long long acc = 0;
    const size_t n = (size_t)rows * pitch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / pitch), w = (int)(i % pitch);
        int ru, rd;
        if (yperiod > 0) {
            ru = (r % yperiod == 0) ? r + yperiod - 1 : r - 1;
            rd = ((r + 1) % yperiod == 0) ? r + 1 - yperiod : r + 1;
        } else {
            ru = r - 1; rd = r + 1;      // ghost rows
        }
        const int wl = (w % xperiod == 0) ? w + xperiod - 1 : w - 1;
        const int wr = ((w + 1) % xperiod == 0) ? w + 1 - xperiod : w + 1;
        const uint32_t me = black[(ptrdiff_t)r * pitch + w];
        const uint32_t up = white[(ptrdiff_t)ru * pitch + w];
        const uint32_t dn = white[(ptrdiff_t)rd * pitch + w];
        const uint32_t ct = white[(ptrdiff_t)r * pitch + w];
        const bool read_back = !((grow0 + (uint32_t)r) & 1u);
        const uint32_t sd = read_back ? __funnelshift_l(white[(ptrdiff_t)r * pitch + wl], ct, 1)
                                      : __funnelshift_r(ct, white[(ptrdiff_t)r * pitch + wr], 1);
        acc += __popc(me ^ up) + __popc(me ^ dn) + __popc(me ^ ct) + __popc(me ^ sd);
    }
    acc = block_sum_u64(acc);
    if (threadIdx.x == 0) atomicAdd(counts, acc);
}

void launch_bond_differ(const uint32_t* black, const uint32_t* white, int rows, int pitch, int yperiod, int xperiod,
                        uint32_t grow0, unsigned long long* counts, cudaStream_t stream)
{
    bond_differ_kernel<<<148 * 8, 256, 0, stream>>>(black, white, rows, pitch, yperiod, xperiod, grow0, counts);
}

// ------------------------------------------------------------------------------------------
// +-J bonds.  bond_random_kernel = hamiltInitB_k (main.cu:153-212): per reference thread 64 Philox
// calls of curand_init(seed+1, tid, 0); draw 128*j + 8*z + 2*l + h decides bit l (0 right, 1 left,
// 2 down, 3 up) of site 16*h + z of word c0 + 16*j; bit set iff uniform < prob.  Rows -1 and Y
// (the neighbouring slabs' edge rows, recomputed from their global row) are filled too so that
// bond_mirror_kernel = hamiltInitW_k (main.cu:214-331) needs no communication.
// Planes: 0 = up, 1 = down, 2 = left, 3 = right, each bond_rows x pitch words.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UPDATE_THREADS) bond_random_kernel(uint32_t* __restrict__ planes /* alloc row 0 */,
                                                                   const __grid_constant__ BondParams p)
{
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int x0 = blockIdx.x * TW;
    const int arow = (int)(blockIdx.z * gridDim.y + blockIdx.y) * TH + ty;   // allocation row 0 .. Y+1
    if (arow >= p.rows + 2) return;
    // global row of allocation row arow (local row arow-1), periodic over the whole lattice
    const long long g = ((long long)p.grow0 + arow - 1 + p.total_rows) % p.total_rows;
    const uint32_t tid = ref_tid((uint32_t)g, (uint32_t)(x0 + tx), p.grid_x);
    const size_t plane = (size_t)(p.rows + 2) * p.pitch;
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        uint32_t acc[4] = {0, 0, 0, 0};                // by bit l
#pragma unroll 1
        for (int z = 0; z < 16; z++) {
#pragma unroll
            for (int lh = 0; lh < 2; lh++) {
                uint32_t o[4];
                philox4x32_10((uint32_t)(32 * j + 2 * z + lh), 0u, tid, 0u, p.key0, p.key1, o);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int l = 2 * lh + (q >> 1), h = q & 1;
                    const bool set = p.always ? true : (!p.never && o[q] <= p.thr);
                    if (set) acc[l] |= 1u << (16 * h + z);
                }
            }
        }
        const size_t off = (size_t)arow * p.pitch + x0 + tx + 16 * j;
        planes[0 * plane + off] = acc[3];   // up
        planes[1 * plane + off] = acc[2];   // down
        planes[2 * plane + off] = acc[1];   // left
        planes[3 * plane + off] = acc[0];   // right
    }
}

__global__ void __launch_bounds__(256) bond_mirror_kernel(const uint32_t* __restrict__ src /* alloc row 0, rows+2 */,
                                                         uint32_t* __restrict__ dst /* local row 0, rows */,
                                                         const __grid_constant__ BondParams p)
{
    const size_t n = (size_t)p.rows * p.pitch;
    const size_t splane = (size_t)(p.rows + 2) * p.pitch, dplane = n;
    const uint32_t* s0 = src + p.pitch;     // local row 0 of the source planes
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / p.pitch), w = (int)(i % p.pitch);
        int ru, rd;
        if (p.yperiod > 0) {
            ru = (r % p.yperiod == 0) ? r + p.yperiod - 1 : r - 1;
            rd = ((r + 1) % p.yperiod == 0) ? r + 1 - p.yperiod : r + 1;
        } else {
            ru = r - 1; rd = r + 1;
        }
        const int wl = (w % p.xperiod == 0) ? w + p.xperiod - 1 : w - 1;
        const int wr = ((w + 1) % p.xperiod == 0) ? w + 1 - p.xperiod : w + 1;
        const bool odd = ((p.grow0 + (uint32_t)r) & 1u) != 0;
        const uint32_t* up_ = s0 + 0 * splane; const uint32_t* dn_ = s0 + 1 * splane;
        const uint32_t* lf_ = s0 + 2 * splane; const uint32_t* rt_ = s0 + 3 * splane;
        const ptrdiff_t row = (ptrdiff_t)r * p.pitch;
        // my up = down bond of the site above, my down = up bond of the site below
        dst[0 * dplane + i] = dn_[(ptrdiff_t)ru * p.pitch + w];
        dst[1 * dplane + i] = up_[(ptrdiff_t)rd * p.pitch + w];
        uint32_t left, right;
        if (!odd) {   // even rows: left partner is site m of the source, right partner is m+1
            left = rt_[row + w];
            right = __funnelshift_r(lf_[row + w], lf_[row + wr], 1);
        } else {      // odd rows: left partner is m-1, right partner is m
            left = __funnelshift_l(rt_[row + wl], rt_[row + w], 1);
            right = lf_[row + w];
        }
        dst[2 * dplane + i] = left;
        dst[3 * dplane + i] = right;
    }
}

void launch_bond_init(uint32_t* random_planes, uint32_t* mirror_planes, const BondParams& p, cudaStream_t stream)
{
    bond_random_kernel<<<tile_grid(p.pitch / TW, (p.rows + 2 + TH - 1) / TH), UPDATE_THREADS, 0, stream>>>(random_planes, p);
    bond_mirror_kernel<<<148 * 8, 256, 0, stream>>>(random_planes, mirror_planes, p);
}

// ------------------------------------------------------------------------------------------
// Two-point correlation (getCorr2D_k / getCorr2DRepl_k, main.cu:870-1070): for every site the
// site j columns to the right and j rows below, j = 1..ncorr.  In the 1-bit colour planes both
// are plain XORs: a shift of j = 2d columns pairs black with black and white with white d bits
// on; j = 2d+1 pairs black m with white m+d+rho and white m with black m+d+1-rho (rho = row
// parity); j rows below pairs equal colours for even j and swapped colours for odd j at the same
// bit.  One CTA per row; counts of EQUAL pairs are accumulated, the host turns them into the
// reference's normalised (+1/-1) sums.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rot_word(const uint32_t* __restrict__ row, int w, int shift, int period)
{
    // word w of `row` rotated left by `shift` sites inside its period-aligned segment:
    // result bit b = site (32*w + b + shift) of the segment (periodic)
    const int seg = (w / period) * period, wl = w - seg;
    const int ws = shift >> 5, bs = shift & 31;
    const uint32_t lo = row[seg + (wl + ws) % period];
    const uint32_t hi = row[seg + (wl + ws + 1) % period];
    return __funnelshift_r(lo, hi, bs);
}

__global__ void __launch_bounds__(128) corr_kernel(const __grid_constant__ CorrParams p, unsigned long long* __restrict__ equal)
{
    extern __shared__ uint32_t sRow[];          // black row, white row
    uint32_t* sB = sRow;
    uint32_t* sW = sRow + p.pitch;
    const int r = blockIdx.x;                     // local row
    const long long g = (long long)p.grow0 + r;   // global row
    const int rho = (int)(g & 1);
    const uint32_t* B = p.black[p.self] + (size_t)r * p.pitch;
    const uint32_t* W = p.white[p.self] + (size_t)r * p.pitch;
    for (int w = threadIdx.x; w < p.pitch; w += blockDim.x) { sB[w] = B[w]; sW[w] = W[w]; }
    __syncthreads();
    for (int j = 1; j <= p.ncorr; j++) {
        unsigned long long acc = 0;
        // row j below, periodic inside the (sub-)lattice
        const long long base = (g / p.yperiod) * p.yperiod;
        const long long gj = base + (g - base + j) % p.yperiod;
        const int slab = (int)(gj / p.rows), rj = (int)(gj % p.rows);
        const uint32_t* Bj = p.black[slab] + (size_t)rj * p.pitch;
        const uint32_t* Wj = p.white[slab] + (size_t)rj * p.pitch;
        const int d = j >> 1;
        for (int w = threadIdx.x; w < p.pitch; w += blockDim.x) {
            const uint32_t b = sB[w], wh = sW[w];
            if (j & 1) {
                acc += __popc(~(b ^ rot_word(sW, w, d + rho, p.xperiod))) + __popc(~(wh ^ rot_word(sB, w, d + 1 - rho, p.xperiod)));
                acc += __popc(~(b ^ Wj[w])) + __popc(~(wh ^ Bj[w]));
            } else {
                acc += __popc(~(b ^ rot_word(sB, w, d, p.xperiod))) + __popc(~(wh ^ rot_word(sW, w, d, p.xperiod)));
                acc += __popc(~(b ^ Bj[w])) + __popc(~(wh ^ Wj[w]));
            }
        }
        acc = block_sum_u64(acc);
        if (threadIdx.x == 0) atomicAdd(equal + (j - 1), acc);
        __syncthreads();
    }
}

void launch_corr(const CorrParams& p, unsigned long long* equal, cudaStream_t stream)
{
    const size_t smem = 2 * (size_t)p.pitch * sizeof(uint32_t);
    if (smem > 48 * 1024) cudaFuncSetAttribute(corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    corr_kernel<<<p.rows, 128, smem, stream>>>(p, equal);
}

// ------------------------------------------------------------------------------------------
// Conversion to / from the reference's 4-bit storage (main.cu:1160-1175): 16 spins per uint64,
// nibble k = site 16*j + k.  One 32-site word <-> two uint64.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spread16(uint32_t h)
{
    unsigned long long x = h & 0xffffu;
    x = (x | (x << 24)) & 0x000000ff000000ffull;
    x = (x | (x << 12)) & 0x000f000f000f000full;
    x = (x | (x << 6)) & 0x0303030303030303ull;
    x = (x | (x << 3)) & 0x1111111111111111ull;
    return x;
}
__device__ __forceinline__ uint32_t gather16(unsigned long long x)
{
    x &= 0x1111111111111111ull;
    x = (x | (x >> 3)) & 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000f000f000f000full;
    x = (x | (x >> 12)) & 0x000000ff000000ffull;
    x = (x | (x >> 24)) & 0xffffull;
    return (uint32_t)x;
}

__global__ void __launch_bounds__(256) to_ref_kernel(const uint32_t* __restrict__ plane, size_t n,
                                                    ulonglong2* __restrict__ out)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = plane[i];
        out[i] = make_ulonglong2(spread16(v), spread16(v >> 16));
    }
}
__global__ void __launch_bounds__(256) from_ref_kernel(const ulonglong2* __restrict__ in, size_t n,
                                                      uint32_t* __restrict__ plane)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const ulonglong2 v = in[i];
        plane[i] = gather16(v.x) | (gather16(v.y) << 16);
    }
}

void launch_to_ref(const uint32_t* plane, int rows, int pitch, unsigned long long* out, cudaStream_t stream)
{
    to_ref_kernel<<<148 * 8, 256, 0, stream>>>(plane, (size_t)rows * pitch, reinterpret_cast<ulonglong2*>(out));
}
void launch_from_ref(const unsigned long long* in, int rows, int pitch, uint32_t* plane, cudaStream_t stream)
{
    from_ref_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const ulonglong2*>(in), (size_t)rows * pitch, plane);
}

}  // namespace ising
