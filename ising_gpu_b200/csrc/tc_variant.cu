// tc_variant.cu -- tensor-core variant of the Metropolis update: B200-native replacement of
// the hot path of the reference's tensorcore/ program (tensorcore/main.cu: init_spins :63-79,
// 4x cublasGemmBatchedEx :252-262/:277-287, add_boundaries :148-226, update_spins :88-146).
//
// The reference materialises the neighbour sums: per colour two batched fp16 GEMMs write
// nn_sums (2 GiB at 32768^2), a third kernel patches the sub-block edges, a fourth reads
// nn_sums back and flips spins.  Here one kernel per colour owns a 128x128 sub-block per CTA,
// forms its neighbour sums on the tensor cores as an INT8 contraction with the bidiagonal 0/1
// matrices, keeps them on chip, adds the edge terms and does the Metropolis step -- nn_sums
// never exists in memory and the lattice is int8 (+-1) instead of fp16.
//
// Contract kept bit for bit: the blocked storage order, the element -> (Philox subsequence,
// offset) map of update_spins/init_spins, the acceptance test `uniform < exp(-2 b nn s)` with
// the DEVICE expf (evaluated once per temperature, turned into integer thresholds on the host),
// and the final.txt format.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ising_b200.h"
#include "host_math.h"
#include "philox_map.h"

using namespace ising;

namespace {

constexpr int SUPN = 256;   // LATTICE_SUP_N
constexpr int SUBN = 128;   // LATTICE_SUB_N
constexpr int TPS = SUBN * SUBN / 16;   // reference threads per sub-block (NLOOPS*SPINSPERTHREAD = 16)

constexpr int MAX_TC_SLABS = 16;

struct TcParams {
    int nbx, nby;            // super-blocks of the WHOLE lattice
    int nby_loc;             // super-block rows per device (slab), nby / ngpus (main.cu:243, :266 split by batch)
    int slab;                // this launch's slab; it owns super-block rows [slab*nby_loc, (slab+1)*nby_loc)
    int8_t* slab_base[MAX_TC_SLABS];   // storage of every slab (peer-mapped), for the terms that cross a slab edge
    uint32_t key0, key1;
    uint64_t call0;          // Philox counter of the first of the 4 calls: 4*(2*iter + colour)
    uint32_t thr[5];         // accept iff x <= thr[k], k = (nn*s+4)/2 ...
    uint32_t always[5];      // ... or always[k] (all ones), and never if never[k]
    uint32_t never[5];
    int is_black;
    int fast;                // v < 0 always flips, v = 0, 2, 4 have thresholds (every alpha > ~0.16)
    StepTables st;           // folded Philox schedule of the 4 calls of this colour step
};

__host__ __device__ __forceinline__ long long sup_offset(long long i, long long j, long long nbx)
{
    return (j * nbx + i) * SUPN * SUPN;
}
__host__ __device__ __forceinline__ long long sub_offset(int a, int b) { return ((long long)b * SUPN + (long long)a * SUBN) * SUBN; }

// spin of sub-block kind (a,b) at element (i,j) of super-block (sx,sy) of the whole lattice, with
// i/j allowed to step one element outside [0,128): continues in the periodic neighbour
// super-block, which may live on another device's slab.
__device__ __forceinline__ int spin_at(const TcParams& p, int sx, int sy, int a, int b, int i, int j)
{
    if (i < 0) { i += SUBN; sx = (sx == 0) ? p.nbx - 1 : sx - 1; }
    if (i >= SUBN) { i -= SUBN; sx = (sx == p.nbx - 1) ? 0 : sx + 1; }
    if (j < 0) { j += SUBN; sy = (sy == 0) ? p.nby - 1 : sy - 1; }
    if (j >= SUBN) { j -= SUBN; sy = (sy == p.nby - 1) ? 0 : sy + 1; }
    const int d = sy / p.nby_loc;
    return p.slab_base[d][sup_offset(sx, sy - d * p.nby_loc, p.nbx) + sub_offset(a, b) + (long long)j * SUBN + i];
}

// which sub-block does CTA `sb` of this colour own (update_spins :99-124)
__device__ __forceinline__ void block_of(int sb, int nbx, bool black, int& bi, int& bj)
{
    bi = sb % (2 * nbx);
    const int h = sb / (2 * nbx);
    bj = black ? ((bi & 1) ? 2 * h + 1 : 2 * h) : ((bi & 1) ? 2 * h : 2 * h + 1);
}

// ------------------------------------------------------------------------------------------
// Validation path: integer stencil.  One CTA per sub-block, thread j owns row j (128 spins).
// ------------------------------------------------------------------------------------------
template <bool WRITE_NN>
__global__ void __launch_bounds__(SUBN) tc_update_stencil_kernel(int8_t* __restrict__ lat, int8_t* __restrict__ nn_out,
                                                                 const __grid_constant__ TcParams p)
{
    const int sb = blockIdx.x + p.slab * 2 * p.nbx * p.nby_loc, j = threadIdx.x;   // global sub-block index
    int bi, bj;
    block_of(sb, p.nbx, p.is_black != 0, bi, bj);
    const int a = bi & 1, b = bj & 1, sx = bi >> 1, sy = bj >> 1;
    const long long base = sup_offset(sx, sy - p.slab * p.nby_loc, p.nbx) + sub_offset(a, b);   // inside this slab
    const int di = (a == 0) ? -1 : +1, dj = (b == 0) ? -1 : +1;

    for (int c = 0; c < 16; c++) {
        const int lin = 16 * j + c, n = lin / TPS, tl = lin % TPS;
        const uint64_t tid = (uint64_t)sb * TPS + (uint64_t)tl;
        int8_t s[8];
        *reinterpret_cast<uint2*>(s) = *reinterpret_cast<const uint2*>(lat + base + (long long)j * SUBN + 8 * c);
        uint32_t o[8];
        for (int h = 0; h < 2; h++) {
            const uint64_t ctr = p.call0 + 2 * n + h;
            philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)tid, (uint32_t)(tid >> 32), p.key0, p.key1,
                          o + 4 * h);
        }
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int i = 8 * c + m;
            const int nn = spin_at(p, sx, sy, 1 - a, b, i, j) + spin_at(p, sx, sy, 1 - a, b, i + di, j) +
                           spin_at(p, sx, sy, a, 1 - b, i, j) + spin_at(p, sx, sy, a, 1 - b, i, j + dj);
            if (WRITE_NN) {
                nn_out[base + (long long)j * SUBN + i] = (int8_t)nn;
            } else {
                const int k = (nn * s[m] + 4) >> 1;
                const bool flip = p.always[k] || (!p.never[k] && o[m] <= p.thr[k]);
                if (flip) s[m] = (int8_t)-s[m];
            }
        }
        if (!WRITE_NN) *reinterpret_cast<uint2*>(lat + base + (long long)j * SUBN + 8 * c) = *reinterpret_cast<uint2*>(s);
    }
}

// ------------------------------------------------------------------------------------------
// Product path: neighbour sums on the 5th-generation tensor cores (tcgen05, INT8 -> INT32 in
// TMEM), fused with the Metropolis step.
//
//   nn[j][i] = sum_k H[j][k] * Kh[k][i]  +  sum_k Kv[j][k] * V[k][i]          (128 x 128 x 128)
//
// H = the sub-block holding the left/right neighbours, V = the one holding the up/down
// neighbours (both int8 +-1 straight from the lattice), Kh/Kv = bidiagonal 0/1 matrices (main
// diagonal + the sub- or super-diagonal, by the parity (a,b) of the target sub-block): the
// reference's "lat x K + K^T x lat" (tensorcore/main.cu:547-569) with int8 instead of fp16.
// Eight tcgen05.mma.kind::i8 (M=128, N=128, K=32) accumulate into one 128-lane x 128-column
// TMEM tile; H is a K-major A operand, V is used in place as an MN-major B operand (no transposed
// copy), and ONE bidiagonal matrix serves as Kh and Kv in all four parities: K-major for the
// "+1" neighbour, MN-major (its transpose) for the "-1" neighbour.  H, V, K and the tile being
// updated are brought by TMA with the 128-byte swizzle; the updated tile goes back with one TMA
// store, so global memory only sees full 128-byte lines.  The terms that cross the sub-block edge
// (add_boundaries, main.cu:148-226) are added in the epilogue from the neighbouring blocks.
// Thread j owns TMEM lane j = row j of the sub-block: tcgen05.ld brings its 128 sums to
// registers, where the 16 runs of 8 spins are updated with the reference's RNG map.
// ------------------------------------------------------------------------------------------
constexpr int TILE_BYTES = SUBN * SUBN;   // 16 KiB per operand tile
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_128B operand descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), the
// 16-byte chunk index of every row XORed with (row % 8) -- exactly what a TMA load with
// CU_TENSOR_MAP_SWIZZLE_128B of a 128-byte-wide box leaves in shared memory.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__host__ __device__ constexpr uint32_t IDESC_I8(int a_mn_major, int b_mn_major, int n = SUBN)
{
    return (2u << 4) /* D = S32 */ | (1u << 7) /* A = INT8 */ | (1u << 10) /* B = INT8 */ |
           ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) /* N */ |
           ((uint32_t)(SUBN >> 4) << 24) /* M */;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

constexpr int MMA_THREADS = 256;   // warps w and w+4 share TMEM lanes 32(w%4).., split the 128 columns

__device__ __forceinline__ void tma_load_rows(void* dst, const CUtensorMap* map, int row, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(dst)),
                 "l"(map), "r"(0), "r"(row), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_parity0(uint64_t* bar)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_addr(bar))
        : "memory");
}

// lat_map: the lattice as (n/128) rows of 128 bytes, box 128 x 128, SWIZZLE_128B -- one box = one
// sub-block.  k_map: the four constant bidiagonal matrices (512 rows of 128 bytes), same box.
template <bool WRITE_NN, bool FAST>
__global__ void __launch_bounds__(MMA_THREADS, 3)
tc_update_mma_kernel(const __grid_constant__ CUtensorMap lat_map, const __grid_constant__ CUtensorMap k_map,
                     int8_t* __restrict__ lat, int8_t* __restrict__ nn_out, const __grid_constant__ TcParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128-byte swizzle atoms must start on 1024-byte boundaries of the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t* sH = smem;
    uint8_t* sV = smem + TILE_BYTES;
    uint8_t* sK = smem + 2 * TILE_BYTES;
    uint8_t* sOwn = smem + 3 * TILE_BYTES;
    __shared__ __align__(16) int8_t sEdge[SUBN];   // the row across the vertical sub-block edge
    __shared__ __align__(8) uint64_t tma_bar;
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_slot;
    // thresholds by (neighbour sum nn, own spin s), addressed without arithmetic on s: the word at
    // byte 16 + 4*nn + off holds thr(v = nn*s) with off = 0 for s = +1 and off = 4 (spin bytes 0, 2
    // of a word) or 1024 (bytes 1, 3) for s = -1 -- "off" is one masked bit of the spin word itself
    __shared__ __align__(16) uint32_t lut[512];

    const int sb = blockIdx.x + p.slab * 2 * p.nbx * p.nby_loc;   // global sub-block index (RNG map)
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    int bi, bj;
    block_of(sb, p.nbx, p.is_black != 0, bi, bj);
    const int a = bi & 1, b = bj & 1, sx = bi >> 1, sy = bj >> 1;
    const long long sup = sup_offset(sx, sy - p.slab * p.nby_loc, p.nbx);   // inside this device's slab
    const long long base = sup + sub_offset(a, b);

    // ---- TMEM allocation (warp 0), barriers, TMA of the four operand tiles (thread 0) -----------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_slot)),
                     "r"((uint32_t)SUBN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&tma_bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mma_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&tma_bar)),
                     "r"((uint32_t)(4 * TILE_BYTES + SUBN))
                     : "memory");
        // source sub-blocks: H = (1-a, b) holds the left/right neighbours, V = (a, 1-b) the up/down ones
        tma_load_rows(sH, &lat_map, (int)((sup + sub_offset(1 - a, b)) >> 7), &tma_bar);
        tma_load_rows(sV, &lat_map, (int)((sup + sub_offset(a, 1 - b)) >> 7), &tma_bar);
        // one bidiagonal matrix K+[r][c] = 1 for c = r, r + 1 serves all four cases: taken K-major it is
        // the "+1" operand, taken MN-major (transposed) the "-1" operand
        tma_load_rows(sK, &k_map, SUBN, &tma_bar);
        // the spins being updated: staged here, updated in place, stored back as one tile
        tma_load_rows(sOwn, &lat_map, (int)(base >> 7), &tma_bar);
        // vertical edge term (add_boundaries): last/first row of sub-block (a, 1-b) of the super-block
        // above (b == 0) or below (b == 1), which may live in the neighbouring device's slab
        {
            const int sy2 = (b == 0) ? (sy == 0 ? p.nby - 1 : sy - 1) : (sy == p.nby - 1 ? 0 : sy + 1);
            const int d2 = sy2 / p.nby_loc;
            const int8_t* edge_row = p.slab_base[d2] + sup_offset(sx, sy2 - d2 * p.nby_loc, p.nbx) + sub_offset(a, 1 - b) +
                                     ((b == 0) ? (SUBN - 1) * SUBN : 0);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(sEdge)),
                         "l"(edge_row), "r"((uint32_t)SUBN), "r"(smem_addr(&tma_bar))
                         : "memory");
        }
    }
    if (t < 5) {          // t = (nn + 4) / 2
        const uint32_t plus = p.always[t] ? 0xffffffffu : p.thr[t], minus = p.always[4 - t] ? 0xffffffffu : p.thr[4 - t];
        lut[2 * t] = plus;
        lut[2 * t + 1] = minus;
        lut[256 + 2 * t] = minus;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;

    // ---- 8 MMAs issued by one thread once the tiles have landed ------------------------------
    if (t == 0) {
        mbar_wait_parity0(&tma_bar);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_addr(sH), aV = smem_addr(sV), aK = smem_addr(sK);
        // k-step of a K-major operand = 32 bytes inside the 128-byte rows, of an MN-major one = 32 rows
        const uint32_t kh_step = (a == 0) ? 4096u : 32u, kv_step = (b == 0) ? 4096u : 32u;
        const uint32_t idesc_h = IDESC_I8(0, a == 0), idesc_v = IDESC_I8(b == 0, 1);
#pragma unroll
        for (int s = 0; s < 4; s++)      // H (K-major A) x Kh (B): nn[j][i] = H[j][i] + H[j][i+di]
            umma_i8(tmem_d, umma_desc_sw128(aH + 32 * s), umma_desc_sw128(aK + kh_step * s), idesc_h, s > 0);
#pragma unroll
        for (int s = 0; s < 4; s++)      // Kv (A) x V (MN-major B, in place): += V[j][i] + V[j+dj][i]
            umma_i8(tmem_d, umma_desc_sw128(aK + kv_step * s), umma_desc_sw128(aV + 4096 * s), idesc_v, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&mma_bar))
                     : "memory");
    }

    // ---- while the tensor core works: own spins, edge terms ----------------------------------
    const int j = 32 * (warp & 3) + lane;        // row = TMEM lane
    const int half = warp >> 2;                  // columns 64*half .. 64*half+63
    // own row in the swizzled tile: 16-byte chunk L of row j sits at chunk L ^ (j % 8)
    uint4* own_row = reinterpret_cast<uint4*>(sOwn + j * SUBN);
    const int swz = j & 7;
    // horizontal edge term: lands on column 0 (a == 0) or 127 (a == 1) of every row
    const int edge_col = (a == 0) ? 0 : SUBN - 1;
    int edge_h = 0;
    if ((edge_col >> 6) == half) edge_h = spin_at(p, sx, sy, 1 - a, b, (a == 0) ? -1 : SUBN, j);
    // vertical edge term: a whole row, only for the first (b == 0) or last (b == 1) row of the block
    const bool has_edge_v = (b == 0) ? (j == 0) : (j == SUBN - 1);
    const int n = j >> 6;                        // which of the reference thread's two runs this row is
    const uint32_t tid0 = (uint32_t)sb * TPS + (uint32_t)(16 * (j & 63));
    const char* lut_c = reinterpret_cast<const char*>(lut) + 16;
    // Philox tables of this thread's two calls (n is warp-uniform: rows 0..63 use calls 0,1, rows 64..127 calls 2,3)
    const uint32_t ta0 = n ? p.st.a[2] : p.st.a[0], tb0 = n ? p.st.b[2] : p.st.b[0], td0 = n ? p.st.d[2] : p.st.d[0];
    const uint32_t ta1 = n ? p.st.a[3] : p.st.a[1], tb1 = n ? p.st.b[3] : p.st.b[1], td1 = n ? p.st.d[3] : p.st.d[1];

    if (t != 0) mbar_wait_parity0(&tma_bar);   // own tile and edge row have landed (thread 0 waited above)
    uint4 own[4];
#pragma unroll
    for (int ch = 0; ch < 4; ch++) own[ch] = own_row[(4 * half + ch) ^ swz];
    // wait for the MMAs
    mbar_wait_parity0(&mma_bar);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

#pragma unroll
    for (int g = 0; g < 2; g++) {               // 2 groups of 32 columns per thread
        uint32_t r[32];
        const int col0 = 64 * half + 32 * g;
        const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr)
            : "memory");
        // terms from across the sub-block edge (add_boundaries)
        if (col0 == 0 && edge_col == 0) r[0] += (uint32_t)edge_h;
        if (col0 == SUBN - 32 && edge_col == SUBN - 1) r[31] += (uint32_t)edge_h;
        if (has_edge_v) {
#pragma unroll
            for (int q4 = 0; q4 < 8; q4++) {
                const uint32_t e4 = *reinterpret_cast<const uint32_t*>(sEdge + col0 + 4 * q4);
#pragma unroll
                for (int q = 0; q < 4; q++) r[4 * q4 + q] += (uint32_t)(int)(int8_t)(e4 >> (8 * q));
            }
        }
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {        // 4 runs of 8 spins per group
            const int c = (col0 >> 3) + cc;     // run index inside the row, 0..15
            uint4& w4 = own[2 * g + (cc >> 1)];
            uint32_t w[2] = {(cc & 1) ? w4.z : w4.x, (cc & 1) ? w4.w : w4.y};   // 8 own spins, one per byte
            if (WRITE_NN) {
#pragma unroll
                for (int h = 0; h < 2; h++)
                    w[h] = (r[8 * cc + 4 * h] & 0xffu) | ((r[8 * cc + 4 * h + 1] & 0xffu) << 8) |
                           ((r[8 * cc + 4 * h + 2] & 0xffu) << 16) | ((r[8 * cc + 4 * h + 3] & 0xffu) << 24);
            } else if (FAST) {
                // one reference thread id per run: fold its constants, then 2 calls of 4 outputs
                const ThreadConsts tc = make_thread_consts(tid0 + (uint32_t)c, p.st.x1, p.st.k1r3);
                uint32_t o[2][4];
                philox_fast(tc, ta0, tb0, td0, p.st.k0, p.st.k1, o[0]);
                philox_fast(tc, ta1, tb1, td1, p.st.k0, p.st.k1, o[1]);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t w0 = w[h], w16 = w[h] >> 16;
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        // byte offset 16 + 4*nn + (s == -1 ? 4 or 1024 : 0)
                        const uint32_t sel = ((m & 2) ? w16 : w0) & ((m & 1) ? 0x400u : 0x4u);
                        const int off = (int)r[8 * cc + 4 * h + m] * 4 + (int)sel;
                        const uint32_t thr = *reinterpret_cast<const uint32_t*>(lut_c + off);
                        asm("{\n.reg .pred q;\nsetp.le.u32 q, %1, %2;\n@q xor.b32 %0, %0, %3;\n}"
                            : "+r"(w[h])
                            : "r"(o[h][m]), "r"(thr), "r"(0xFEu << (8 * m)));   // +1 (0x01) <-> -1 (0xFF)
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t o[4];
                    const uint64_t ctr = p.call0 + 2 * n + h;
                    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), tid0 + (uint32_t)c, 0u, p.key0, p.key1, o);
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const int s = (int)(int8_t)(w[h] >> (8 * m));
                        const int k = ((int)r[8 * cc + 4 * h + m] * s + 4) >> 1;
                        if (p.always[k] || (!p.never[k] && o[m] <= p.thr[k])) w[h] ^= 0xFEu << (8 * m);
                    }
                }
            }
            if (WRITE_NN) {
                *reinterpret_cast<uint2*>(nn_out + base + j * SUBN + 8 * c) = make_uint2(w[0], w[1]);
            } else {
                if (cc & 1) { w4.z = w[0]; w4.w = w[1]; }
                else        { w4.x = w[0]; w4.y = w[1]; }
            }
        }
    }
    if (!WRITE_NN) {
#pragma unroll
        for (int ch = 0; ch < 4; ch++) own_row[(4 * half + ch) ^ swz] = own[ch];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA store
    }

    // ---- store the tile, release TMEM -----------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (!WRITE_NN && t == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&lat_map), "r"(0),
                     "r"((int)(base >> 7)), "r"(smem_addr(sOwn))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released after this
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)SUBN) : "memory");
}

// ------------------------------------------------------------------------------------------
// Product kernel, persistent and warp-specialised.  One CTA per SM walks over the sub-blocks
// blockIdx.x, blockIdx.x + gridDim.x, ... (local tile numbers it = 0, 1, 2, ...).  Warp 16 is the
// producer (one thread: TMA of the H / V operand tiles and the vertical edge row of tile `it`
// into stage it % 4 of a shared-memory ring, then the eight tcgen05.mma into TMEM accumulator
// it % 4 -- all 512 TMEM columns are in use).  Warps 0..15 are two consumer groups of 8 warps;
// group g owns the tiles it = g, g + 2, ... and therefore the stages / accumulators g and g + 2:
// TMEM -> registers, edge terms, Philox, Metropolis step.  The tile being updated never touches
// shared memory: every consumer thread owns 64 consecutive bytes of one row, prefetches them one
// tile ahead with four 16-byte loads and writes them back with four 16-byte stores.
// Hand-shake per stage s: full[s] (TMA bytes landed, seen by the producer only) -> mma[s]
// (tcgen05.commit + a release arrive of the producer) -> free[s] (one arrive per warp of the owning
// group once its TMEM columns and the edge row are in registers; the producer may then overwrite
// the accumulator).  The operand tiles of a stage are refilled as soon as its MMAs have completed,
// so loads run four tiles ahead and a group finds its next tile multiplied when it gets there:
// the load/MMA latency that the one-tile-per-CTA kernel above exposes on every tile is paid once
// per CTA, and there is no CTA-wide barrier inside the loop.
// ------------------------------------------------------------------------------------------
#ifndef ISING_TC_GROUP_WARPS
#define ISING_TC_GROUP_WARPS 8
#endif
#ifndef ISING_TC_GROUPS
#define ISING_TC_GROUPS (16 / ISING_TC_GROUP_WARPS)
#endif
// consumer warps per tile: 8 = 4 TMEM lane quarters x 2 column halves (64 columns per thread), 4 = one warp per
// lane quarter (a thread owns a whole row of 128 columns)
constexpr int PIPE_GROUP_WARPS = ISING_TC_GROUP_WARPS;
constexpr int PIPE_COLS = SUBN / (PIPE_GROUP_WARPS / 4);   // columns per consumer thread
constexpr int PIPE_GROUPS = ISING_TC_GROUPS;            // consumer groups; group g takes the CTA's tiles g, g + G, ...
constexpr int PIPE_CONSUMER_WARPS = PIPE_GROUP_WARPS * PIPE_GROUPS;
// two single-thread service warps behind the consumers: warp PIPE_CONSUMER_WARPS issues the MMAs, the next one the
// TMA loads.  One warp doing both had ~2.2 us of serial work per tile (wait for the MMAs to finish, refill, wait,
// issue) against a tile budget of 2.5 us -- the consumers' 21 % wait at mma[] was a wait for that thread.
constexpr int PIPE_THREADS = 32 * (PIPE_CONSUMER_WARPS + 2);
constexpr int PIPE_STAGES = 4;                          // operand stages = TMEM accumulators (4 x 128 = all 512 columns)
// stage = H tile, V tile, then one 1024-byte swizzle atom whose first row (row 128 of V) receives the edge row of the
// vertically neighbouring block: a fifth k-step of the Kv x V product adds it on the tensor core (selector matrix E)
constexpr int PIPE_STAGE_BYTES = 2 * TILE_BYTES + 1024;
constexpr size_t PIPE_SMEM = (size_t)2 * TILE_BYTES + (size_t)PIPE_STAGES * PIPE_STAGE_BYTES + 4096 + 1024;
static_assert(PIPE_STAGES * SUBN <= 512, "TMEM has 512 columns");
// registers are allocated per 4 warps: 18 warps count as 20 -> 65536 / (20 * 32) = 102 -> 96 per thread;
// 25 warps count as 28 -> 73 -> 72 per thread
constexpr int PIPE_MAXREG = (PIPE_CONSUMER_WARPS <= 16) ? 96 : 72;
// with 96 registers and 64 columns per thread the tile's own spins are prefetched into registers one tile
// ahead; otherwise they are prefetched into L2 and loaded at the start of the tile
// ISING_TC_EARLY_FREE: the sums of the second half of a thread's columns are fetched from TMEM together, at the
// half-way point of the tile, so that the accumulator is handed back after 50 % instead of 75 % of the tile's work
// (more room for the warps of a group to drift).  Needs 16 more registers for the sums; they are taken from the
// register prefetch of the next tile's own spins (L2 prefetch + loads at the start of the tile instead).
#ifndef ISING_TC_EARLY_FREE
#define ISING_TC_EARLY_FREE 0
#endif
constexpr bool PIPE_EARLY_FREE = ISING_TC_EARLY_FREE && (PIPE_COLS == 64);
constexpr bool PIPE_REG_PREFETCH = (PIPE_MAXREG == 96 && PIPE_COLS == 64) && !(PIPE_EARLY_FREE && ISING_TC_EARLY_FREE == 1);
constexpr int PIPE_CHUNKS = PIPE_COLS / 16;             // 16-byte chunks = TMEM column groups per thread
// Build option (off): hand the accumulators over per COLUMN HALF -- every tile multiplied as two M=128, N=64
// products with their own mma[] / free[] barriers, so that a warp only waits for the three other warps that share
// its 64 columns.  Measured SLOWER (1013 vs 1085 updates/ns, profiles/r02_tc_variants.txt): the wait at mma[] was
// never the 8-warp hand-over but an imbalance between the warps (see the fifth k-step of the MMA issuer), and the
// second set of MMAs and barriers only lengthens the issuer's serial path.
#ifndef ISING_TC_SPLIT_HALVES
#define ISING_TC_SPLIT_HALVES 0
#endif
constexpr int PIPE_HALVES = (ISING_TC_SPLIT_HALVES && PIPE_GROUP_WARPS == 8) ? 2 : 1;
constexpr int PIPE_HALF_COLS = SUBN / PIPE_HALVES;      // N of one MMA
constexpr int PIPE_HALF_WARPS = PIPE_GROUP_WARPS / PIPE_HALVES;

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"   // suspend-time hint: sleep in hardware, not in a loop
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

struct TileGeom {
    int sb, a, b, sx, sl;    // global sub-block index, parities, super-block column, super-block row inside the slab
    long long sup, base;     // byte offset of the super-block / of the sub-block inside this device's slab
};
// The tiles tile0, tile0 + step, ... of this colour on this slab, without a division per tile:
// sub-block index sb = h * 2nbx + bi (update_spins :99-124) is kept as the pair (bi, h).
struct TileWalk {
    int bi, h, dbi, dh, two_nbx;
    __device__ __forceinline__ void init(const TcParams& p, int tile0, int step, int ntiles)
    {
        two_nbx = 2 * p.nbx;
        const int sb = tile0 + p.slab * ntiles;
        bi = sb % two_nbx; h = sb / two_nbx; dbi = step % two_nbx; dh = step / two_nbx;
    }
    __device__ __forceinline__ void advance()
    {
        bi += dbi; h += dh;
        if (bi >= two_nbx) { bi -= two_nbx; h++; }
    }
    __device__ __forceinline__ TileGeom geom(const TcParams& p) const
    {
        TileGeom g;
        g.sb = h * two_nbx + bi;
        g.a = bi & 1;
        g.b = p.is_black ? g.a : 1 - g.a;          // block_of(): bj = 2h + (black ? a : 1-a)
        g.sx = bi >> 1;
        g.sl = h - p.slab * p.nby_loc;             // sy = bj >> 1 = h
        g.sup = (long long)(g.sl * p.nbx + g.sx) * (SUPN * SUPN);
        g.base = g.sup + sub_offset(g.a, g.b);
        return g;
    }
};

template <bool FAST>
__global__ void __maxnreg__(PIPE_MAXREG)
tc_update_pipe_kernel(const __grid_constant__ CUtensorMap lat_map, const __grid_constant__ CUtensorMap k_map,
                      int8_t* __restrict__ lat, const __grid_constant__ TcParams p, int ntiles)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t* sK = smem;                                   // the bidiagonal matrix, loaded once
    uint8_t* sE = smem + TILE_BYTES;                      // the edge-row selector matrix, loaded once
    uint8_t* sHV = smem + 2 * TILE_BYTES;                 // stage s at s * PIPE_STAGE_BYTES: H, V, edge row (+ pad)
    __shared__ __align__(8) uint64_t full_bar[PIPE_STAGES], mma_bar[PIPE_STAGES][PIPE_HALVES], free_bar[PIPE_STAGES][PIPE_HALVES], k_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) uint32_t lut[512];           // see tc_update_mma_kernel

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (warp == PIPE_CONSUMER_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_slot)),
                     "r"((uint32_t)(PIPE_STAGES * SUBN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) {
        for (int s = 0; s < PIPE_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full_bar[s])));
            // mma[s]: one arrival from tcgen05.commit, one plain (release) arrival of the producer thread after
            // it has seen full[s] -- the consumers' only wait, which also orders the edge row for them
            for (int hf = 0; hf < PIPE_HALVES; hf++) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_addr(&mma_bar[s][hf])));
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&free_bar[s][hf])), "r"(PIPE_HALF_WARPS));
            }
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&k_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 5) {          // t = (nn + 4) / 2
        const uint32_t plus = p.always[t] ? 0xffffffffu : p.thr[t], minus = p.always[4 - t] ? 0xffffffffu : p.thr[4 - t];
        lut[2 * t] = plus;
        lut[2 * t + 1] = minus;
        lut[256 + 2 * t] = minus;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_slot;
    const int first = blockIdx.x, stride = gridDim.x;

    const int nslabs = p.nby / p.nby_loc;
    if (warp == PIPE_CONSUMER_WARPS + 1) {
        // ---------------- loader: one thread drives TMA ---------------------------------------
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&k_bar)),
                         "r"((uint32_t)(2 * TILE_BYTES))
                         : "memory");
            // one bidiagonal matrix K+[r][c] = 1 for c = r, r + 1 serves all four cases: taken K-major it is
            // the "+1" operand, taken MN-major (transposed) the "-1" operand
            tma_load_rows(sK, &k_map, SUBN, &k_bar);
            tma_load_rows(sE, &k_map, 4 * SUBN, &k_bar);
            // operand tiles of local tile number `it` (geometry g) into stage it % PIPE_STAGES (its previous MMAs
            // have completed)
            auto load_tile = [&](int it, const TileGeom& g) {
                const int s = it & (PIPE_STAGES - 1);
                uint8_t* sH = sHV + s * PIPE_STAGE_BYTES;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&full_bar[s])),
                             "r"((uint32_t)(2 * TILE_BYTES + SUBN))
                             : "memory");
                // source sub-blocks: H = (1-a, b) holds the left/right neighbours, V = (a, 1-b) the up/down ones
                tma_load_rows(sH, &lat_map, (int)((g.sup + sub_offset(1 - g.a, g.b)) >> 7), &full_bar[s]);
                tma_load_rows(sH + TILE_BYTES, &lat_map, (int)((g.sup + sub_offset(g.a, 1 - g.b)) >> 7), &full_bar[s]);
                // vertical edge term (add_boundaries): last/first row of sub-block (a, 1-b) of the super-block
                // above (b == 0) or below (b == 1), which may live in the neighbouring device's slab.  It goes to
                // row 128 of the V tile (swizzle phase 0 = a plain 128-byte copy), where the fifth k-step finds it.
                int d2 = p.slab, sl2;
                if (g.b == 0) {
                    sl2 = g.sl - 1;
                    if (sl2 < 0) { sl2 = p.nby_loc - 1; d2 = (p.slab == 0) ? nslabs - 1 : p.slab - 1; }
                } else {
                    sl2 = g.sl + 1;
                    if (sl2 == p.nby_loc) { sl2 = 0; d2 = (p.slab == nslabs - 1) ? 0 : p.slab + 1; }
                }
                const int8_t* edge_row = p.slab_base[d2] + (long long)(sl2 * p.nbx + g.sx) * (SUPN * SUPN) +
                                         sub_offset(g.a, 1 - g.b) + ((g.b == 0) ? (SUBN - 1) * SUBN : 0);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_addr(sH + 2 * TILE_BYTES)),
                             "l"(edge_row), "r"((uint32_t)SUBN), "r"(smem_addr(&full_bar[s]))
                             : "memory");
            };
            TileWalk lw;
            lw.init(p, first, stride, ntiles);
            for (int k = 0; k < PIPE_STAGES; k++)
                if (first + k * stride < ntiles) { load_tile(k, lw.geom(p)); lw.advance(); }
            // refill stage s with tile it + PIPE_STAGES as soon as the tensor core has read tile it out of it
            int it = 0;
            for (int tile = first; tile + PIPE_STAGES * stride < ntiles; tile += stride, ++it) {
                const int s = it & (PIPE_STAGES - 1);
                const uint32_t ph = (uint32_t)(it / PIPE_STAGES) & 1u;
                for (int hf = 0; hf < PIPE_HALVES; hf++) mbar_wait_parity(&mma_bar[s][hf], ph);
                load_tile(it + PIPE_STAGES, lw.geom(p));
                lw.advance();
            }
        }
        __syncwarp();
    } else if (warp == PIPE_CONSUMER_WARPS) {
        // ---------------- MMA issuer: one thread drives the tensor core ------------------------
        if (lane == 0) {
            const uint32_t aK = smem_addr(sK), aE = smem_addr(sE);
            const int a = first & 1, b = p.is_black ? a : 1 - a;   // parities of every sub-block of this CTA (stride, 2nbx
                                                                   // and ntiles are even, so tile & 1 == first & 1)
            // k-step of a K-major operand = 32 bytes inside the 128-byte rows, of an MN-major one = 32 rows
            const uint32_t kh_step = (a == 0) ? 4096u : 32u, kv_step = (b == 0) ? 4096u : 32u;
            const uint32_t idesc_h = IDESC_I8(0, a == 0, PIPE_HALF_COLS), idesc_v = IDESC_I8(b == 0, 1, PIPE_HALF_COLS);
            const uint32_t idesc_e = IDESC_I8(0, 1, PIPE_HALF_COLS);
            // columns 64*hf .. of the B operands: 64 rows on in a K-major tile, 64 bytes on in an MN-major one
            const uint32_t nh_step = (a == 0) ? (uint32_t)PIPE_HALF_COLS : (uint32_t)PIPE_HALF_COLS * 128u;
            // descriptors: everything that does not depend on the stage is built once; a byte offset inside the
            // 256 KiB window is an add of offset / 16 to the low word of a descriptor
            uint64_t dKh[PIPE_HALVES][4], dKv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                dKv[k] = umma_desc_sw128(aK + kv_step * k);
#pragma unroll
                for (int hf = 0; hf < PIPE_HALVES; hf++) dKh[hf][k] = umma_desc_sw128(aK + kh_step * k + nh_step * hf);
            }
            const uint64_t dE = umma_desc_sw128(aE + (b == 0 ? 0u : 32u));
            const uint64_t dH0 = umma_desc_sw128(smem_addr(sHV));      // stage 0; stage s: + s * PIPE_STAGE_BYTES / 16
            mbar_wait_parity(&k_bar, 0);
            int it = 0;
#pragma unroll 1
            for (int tile = first; tile < ntiles; tile += stride, ++it) {
                const int s = it & (PIPE_STAGES - 1);
                const uint32_t ph = (uint32_t)(it / PIPE_STAGES) & 1u;
                mbar_wait_parity(&full_bar[s], ph);                               // operands landed
                const uint64_t dH = dH0 + (uint64_t)(s * (PIPE_STAGE_BYTES >> 4)), dV = dH + (TILE_BYTES >> 4);
#pragma unroll
                for (int hf = 0; hf < PIPE_HALVES; hf++) {
                    if (it >= PIPE_STAGES) mbar_wait_parity(&free_bar[s][hf], ph ^ 1u);   // accumulator (half) drained
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&mma_bar[s][hf])) : "memory");
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t acc = tmem_d + (uint32_t)(s * SUBN + hf * PIPE_HALF_COLS);
#pragma unroll
                    for (int k = 0; k < 4; k++)      // H (K-major A) x Kh (B): nn[j][i] = H[j][i] + H[j][i+di]
                        umma_i8(acc, dH + 2 * k, dKh[hf][k], idesc_h, k > 0);
#pragma unroll
                    for (int k = 0; k < 4; k++)      // Kv (A) x V (MN-major B, in place): += V[j][i] + V[j+dj][i]
                        umma_i8(acc, dKv[k], dV + (4096 >> 4) * k + ((PIPE_HALF_COLS * hf) >> 4), idesc_v, 1u);
                    // fifth k-step: rows 128.. of V = the edge row (the other 31 rows meet zero columns of E):
                    // E (K-major, slice b) x V''  adds the row above to row 0 (b == 0) or the row below to row 127
                    umma_i8(acc, dE, dV + (4096 >> 4) * 4 + ((PIPE_HALF_COLS * hf) >> 4), idesc_e, 1u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     smem_addr(&mma_bar[s][hf]))
                                 : "memory");
                }
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers: row j = TMEM lane, 64 columns per thread --------------------
        const int grp = warp / PIPE_GROUP_WARPS;     // consumer group: local tiles grp, grp + PIPE_GROUPS, ...
        const int j = 32 * (warp & 3) + lane;        // row = TMEM lane (a warp may only touch lanes 32*(warp%4)..)
        const int half = (warp % PIPE_GROUP_WARPS) >> 2;   // first column = PIPE_COLS * half (0 when a thread owns the row)
        const int colbase = PIPE_COLS * half;
        const int n = j >> 6;                        // which of the reference thread's two runs this row is
        const char* lut_c = reinterpret_cast<const char*>(lut) + 16;
        // Philox tables of this thread's two calls (n is warp-uniform: rows 0..63 use calls 0,1, rows 64..127 calls 2,3)
        const uint32_t ta0 = n ? p.st.a[2] : p.st.a[0], tb0 = n ? p.st.b[2] : p.st.b[0], td0 = n ? p.st.d[2] : p.st.d[0];
        const uint32_t ta1 = n ? p.st.a[3] : p.st.a[1], tb1 = n ? p.st.b[3] : p.st.b[1], td1 = n ? p.st.d[3] : p.st.d[1];

        // own spins and the horizontal edge spin of the NEXT tile are fetched while the current one is worked on
        const int first_g = first + grp * stride, stride_g = PIPE_GROUPS * stride;
        TileWalk cw;
        cw.init(p, first_g, stride_g, ntiles);
        TileGeom gn = cw.geom(p);
        uint4 own_n[PIPE_REG_PREFETCH ? PIPE_CHUNKS : 1];
        int edge_n = 0;
        auto own_ptr = [&](const TileGeom& g) { return lat + g.base + (long long)j * SUBN + colbase; };
        // horizontal edge term: lands on column 0 (a == 0) or 127 (a == 1) of every row; it is column 127 of the
        // block to the left (a == 0) / column 0 of the block to the right, same slab.  nullptr = not this thread's.
        auto edge_ptr = [&](const TileGeom& g) -> const int8_t* {
            const int edge_col = (g.a == 0) ? 0 : SUBN - 1;
            if (edge_col < colbase || edge_col >= colbase + PIPE_COLS) return nullptr;
            const int sx2 = (g.a == 0) ? (g.sx == 0 ? p.nbx - 1 : g.sx - 1) : (g.sx == p.nbx - 1 ? 0 : g.sx + 1);
            return lat + (long long)(g.sl * p.nbx + sx2) * (SUPN * SUPN) + sub_offset(1 - g.a, g.b) + (long long)j * SUBN +
                   ((g.a == 0) ? SUBN - 1 : 0);
        };
        // (volatile asm: the compiler must not sink these loads to their first use)
        auto fetch = [&](const TileGeom& g, uint4* o, int& e) {
            const int8_t* src = own_ptr(g);
#pragma unroll
            for (int ch = 0; ch < PIPE_CHUNKS; ch++)
                asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(o[ch].x), "=r"(o[ch].y), "=r"(o[ch].z), "=r"(o[ch].w)
                             : "l"(src + 16 * ch));
            e = 0;
            if (const int8_t* ep = edge_ptr(g)) asm volatile("ld.global.s8 %0, [%1];" : "=r"(e) : "l"(ep));
        };
        auto l2_prefetch = [&](const TileGeom& g) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(own_ptr(g)));
            if (PIPE_COLS > 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(own_ptr(g) + 64));
            if (const int8_t* ep = edge_ptr(g)) asm volatile("prefetch.global.L2 [%0];" ::"l"(ep));
        };
        if (PIPE_REG_PREFETCH && first_g < ntiles) fetch(gn, own_n, edge_n);

        int it = grp;                                // local tile number as the producer counts them
        for (int tile = first_g; tile < ntiles; tile += stride_g, it += PIPE_GROUPS) {
            const int s = it & (PIPE_STAGES - 1);
            const uint32_t ph = (uint32_t)(it / PIPE_STAGES) & 1u;
            const TileGeom g = gn;
            uint4 own[PIPE_CHUNKS];
            int edge_h;
            if (PIPE_REG_PREFETCH) {
#pragma unroll
                for (int ch = 0; ch < PIPE_CHUNKS; ch++) own[ch] = own_n[ch];
                edge_h = edge_n;
            } else {
                fetch(g, own, edge_h);
            }
            if (tile + stride_g < ntiles) {
                cw.advance();
                gn = cw.geom(p);
                if (PIPE_REG_PREFETCH) fetch(gn, own_n, edge_n);
                else                   l2_prefetch(gn);
            }
            const int edge_col = (g.a == 0) ? 0 : SUBN - 1;
            // (the vertical edge term -- a whole row for the first or last row of the block -- is already in the
            // sums: the tensor core added it, see the fifth k-step of the MMA issuer)
            const uint32_t tid0 = (uint32_t)g.sb * TPS + (uint32_t)(16 * (j & 63));
            int8_t* const dst = lat + g.base + (long long)j * SUBN + colbase;

            // neighbour sums are in TMEM, the edge row in shared memory.  (full[s] itself must not be waited
            // on here: the producer refills the stage right after the MMAs and the barrier may already be
            // two phases ahead of a slow consumer.)
            const int hf = (PIPE_HALVES == 2) ? half : 0;
            mbar_wait_parity(&mma_bar[s][hf], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            uint32_t rr[2][16];
            auto tmem_ld = [&](int gq, uint32_t (&r)[16]) {    // asynchronous: tmem_wait() before the values are used
                const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(s * SUBN + colbase + 16 * gq);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
            };
            auto tmem_wait = [&](uint32_t (&r)[16]) {          // the registers are operands so that no use can move above
                asm volatile("tcgen05.wait::ld.sync.aligned;"
                             : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                               "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                             :
                             : "memory");
            };
#pragma unroll
            for (int gq = 0; gq < PIPE_CHUNKS; gq++) {  // groups of 16 columns
                uint32_t (&r)[16] = rr[gq & 1];
                const int col0 = colbase + 16 * gq;
                bool last_ld;                           // the accumulator has been read completely
                if (PIPE_EARLY_FREE && gq >= PIPE_CHUNKS - 2) {
                    if (gq == PIPE_CHUNKS - 2) {        // both remaining groups at once
                        tmem_ld(gq, rr[gq & 1]);
                        tmem_ld(gq + 1, rr[(gq + 1) & 1]);
                        tmem_wait(rr[0]);
                        tmem_wait(rr[1]);
                    }
                    last_ld = (gq == PIPE_CHUNKS - 2);
                } else {
                    tmem_ld(gq, r);
                    tmem_wait(r);
                    last_ld = (gq == PIPE_CHUNKS - 1);
                }
                // terms from across the sub-block edge (add_boundaries)
                if (col0 == 0 && edge_col == 0) r[0] += (uint32_t)edge_h;
                if (col0 == SUBN - 16 && edge_col == SUBN - 1) r[15] += (uint32_t)edge_h;
                if (last_ld) {
                    // this warp is done with accumulator s: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0)
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&free_bar[s][hf])) : "memory");
                }
                uint4& w4 = own[gq];
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {        // 2 runs of 8 spins per group
                    const int c = (col0 >> 3) + cc;     // run index inside the row, 0..15
                    uint32_t w[2] = {cc ? w4.z : w4.x, cc ? w4.w : w4.y};   // 8 own spins, one per byte
                    if (FAST) {
                        // one reference thread id per run: fold its constants, then 2 calls of 4 outputs
                        const ThreadConsts tc = make_thread_consts(tid0 + (uint32_t)c, p.st.x1, p.st.k1r3);
                        uint32_t o[2][4];
                        philox_fast(tc, ta0, tb0, td0, p.st.k0, p.st.k1, o[0]);
                        philox_fast(tc, ta1, tb1, td1, p.st.k0, p.st.k1, o[1]);
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const uint32_t w0 = w[h], w16 = w[h] >> 16;
#pragma unroll
                            for (int m = 0; m < 4; m++) {
                                // byte offset 16 + 4*nn + (s == -1 ? 4 or 1024 : 0)
                                const uint32_t sel = ((m & 2) ? w16 : w0) & ((m & 1) ? 0x400u : 0x4u);
                                const int off = (int)r[8 * cc + 4 * h + m] * 4 + (int)sel;
                                const uint32_t thr = *reinterpret_cast<const uint32_t*>(lut_c + off);
                                asm("{\n.reg .pred q;\nsetp.le.u32 q, %1, %2;\n@q xor.b32 %0, %0, %3;\n}"
                                    : "+r"(w[h])
                                    : "r"(o[h][m]), "r"(thr), "r"(0xFEu << (8 * m)));   // +1 (0x01) <-> -1 (0xFF)
                            }
                        }
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            uint32_t o[4];
                            const uint64_t ctr = p.call0 + 2 * n + h;
                            philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), tid0 + (uint32_t)c, 0u, p.key0, p.key1, o);
#pragma unroll
                            for (int m = 0; m < 4; m++) {
                                const int sp = (int)(int8_t)(w[h] >> (8 * m));
                                const int k = ((int)r[8 * cc + 4 * h + m] * sp + 4) >> 1;
                                if (p.always[k] || (!p.never[k] && o[m] <= p.thr[k])) w[h] ^= 0xFEu << (8 * m);
                            }
                        }
                    }
                    if (cc) {
                        // 16 bytes finished: store them now, long before their registers are reused
                        asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 16 * gq), "r"(w4.x), "r"(w4.y),
                                     "r"(w[0]), "r"(w[1])
                                     : "memory");
                    } else {
                        w4.x = w[0]; w4.y = w[1];
                    }
                }
            }
        }
    }

    // ---- release TMEM -------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == PIPE_CONSUMER_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(PIPE_STAGES * SUBN))
                     : "memory");
}

// init_spins: element t <- first output of Philox(subsequence t, offset 0); uniform < 0.5 -> -1
__global__ void __launch_bounds__(256) tc_init_kernel(int8_t* __restrict__ lat, long long n, long long t0, uint32_t key0,
                                                     uint32_t key1, uint32_t thr_half)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const unsigned long long g = (unsigned long long)(t0 + t);   // element index in the whole lattice
        uint32_t o[4];
        philox4x32_10(0u, 0u, (uint32_t)g, (uint32_t)(g >> 32), key0, key1, o);
        lat[t] = (o[0] <= thr_half) ? -1 : 1;
    }
}

// exp(-2 * inv_temp * v) for v = -4..4, with the device expf (tensorcore/main.cu:137)
__global__ void tc_accept_kernel(float inv_temp, float* out)
{
    const int k = threadIdx.x;
    if (k < 5) out[k] = exp(-2.0f * inv_temp * (float)(2 * k - 4));
}

__global__ void __launch_bounds__(256) tc_sum_kernel(const int8_t* __restrict__ lat, long long n, long long* out)
{
    long long acc = 0;
    const char4* p4 = reinterpret_cast<const char4*>(lat);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n / 4; t += (long long)gridDim.x * blockDim.x) {
        const char4 v = p4[t];
        acc += v.x + v.y + v.z + v.w;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(out), (unsigned long long)acc);
}

thread_local std::string g_tc_err;
int tc_fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_tc_err = buf;
    return code;
}
#define TC_TRY(call)                                                                                        \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return tc_fail(ISING_ECUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

// the error string is shared with the main API through this hook (ising_api.cu)
extern "C" void ising_internal_set_error(const char* msg);

struct TcDev {
    int device = 0;
    int8_t* lat = nullptr;       // this device's slab: nbx x nby_loc super-blocks
    int8_t* scratch = nullptr;
    int8_t* kmat = nullptr;      // 4 x 128 x 128 bidiagonal 0/1 matrices: Kh(di=-1), Kh(di=+1), Kv(dj=-1), Kv(dj=+1)
    CUtensorMap lat_map, k_map;
    long long* d_sum = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, ev_color = nullptr;
    int sms = 148;
};

struct ising_tc_ctx {
    int nbx = 0, nby = 0, ngpus = 1, nby_loc = 0, nn_path = ISING_TC_NN_MMA;
    int nn_kernel = 0;   // 0 = persistent pipeline (product), 1 = one CTA per sub-block (ISING_TC_KERNEL=tile)
    long long n = 0, n_loc = 0;
    uint64_t seed = 0;
    float alpha = 0, accept[5];
    TcParams base{};
    std::vector<TcDev> dev;
    uint64_t launches = 0;
};

#define TC_RET(expr)                                       \
    do {                                                   \
        int rc__ = (expr);                                 \
        if (rc__) ising_internal_set_error(g_tc_err.c_str()); \
        return rc__;                                       \
    } while (0)

static int tc_create_impl(int nbx, int nby, uint64_t seed, float alpha, int first_device, int ngpus, ising_tc_ctx** out)
{
    if (!out || nbx <= 0 || nby <= 0) return tc_fail(ISING_EINVAL, "nbx and nby must be positive");
    *out = nullptr;
    if (ngpus < 1 || ngpus > MAX_TC_SLABS) return tc_fail(ISING_EINVAL, "ngpus must be 1..%d", MAX_TC_SLABS);
    if (nby % ngpus) return tc_fail(ISING_EINVAL, "Number of super blocks in y dimension must be multiple of number of gpus.");
    int avail = 0;
    TC_TRY(cudaGetDeviceCount(&avail));
    if (first_device < 0 || first_device + ngpus > avail) return tc_fail(ISING_EINVAL, "devices %d..%d requested, %d visible", first_device, first_device + ngpus - 1, avail);
    ising_tc_ctx* c = new ising_tc_ctx();
    c->nbx = nbx; c->nby = nby; c->ngpus = ngpus; c->nby_loc = nby / ngpus; c->seed = seed; c->alpha = alpha;
    if (const char* k = getenv("ISING_TC_KERNEL")) c->nn_kernel = (strcmp(k, "tile") == 0) ? 1 : 0;   // developer A/B switch
    c->n = (long long)nbx * nby * SUPN * SUPN;
    c->n_loc = c->n / ngpus;
    c->dev.resize(ngpus);
    auto bail = [&](int rc) { ising_tc_destroy(c); return rc; };

    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp)
        return bail(tc_fail(ISING_ECUDA, "cuTensorMapEncodeTiled entry point unavailable"));
    EncodeFn enc = (EncodeFn)fp;

    // constant bidiagonal matrices
    // + a fifth tile E for the pipeline kernel's edge-row MMA: E[0][0] = 1 (k-slice 0: the row ABOVE the block
    // lands on row 0), E[127][32] = 1 (k-slice 1: the row BELOW lands on row 127)
    std::vector<int8_t> k(5 * SUBN * SUBN, 0);
    k[(size_t)4 * SUBN * SUBN + 0 * SUBN + 0] = 1;
    k[(size_t)4 * SUBN * SUBN + (SUBN - 1) * SUBN + 32] = 1;
    for (int m = 0; m < 4; m++) {
        const int d = (m & 1) ? +1 : -1;
        for (int r = 0; r < SUBN; r++) {
            k[(size_t)m * SUBN * SUBN + r * SUBN + r] = 1;
            if (r + d >= 0 && r + d < SUBN) k[(size_t)m * SUBN * SUBN + r * SUBN + r + d] = 1;
        }
    }
    for (int i = 0; i < ngpus; i++) {
        TcDev& d = c->dev[i];
        d.device = first_device + i;
        cudaError_t e;
        if ((e = cudaSetDevice(d.device)) != cudaSuccess) return bail(tc_fail(ISING_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e)));
        for (int j = 0; j < ngpus; j++) {   // slab edges read the neighbouring device's rows directly
            if (j == i) continue;
            int ok = 0;
            cudaDeviceCanAccessPeer(&ok, d.device, first_device + j);
            if (!ok) return bail(tc_fail(ISING_ENOPEER, "no direct link between devices %d and %d", d.device, first_device + j));
            e = cudaDeviceEnablePeerAccess(first_device + j, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return bail(tc_fail(ISING_ECUDA, "peer access: %s", cudaGetErrorString(e)));
            cudaGetLastError();
        }
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.device);
        if ((e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaMalloc(&d.lat, (size_t)c->n_loc)) != cudaSuccess || (e = cudaMalloc(&d.d_sum, sizeof(long long))) != cudaSuccess ||
            (e = cudaMalloc(&d.kmat, k.size())) != cudaSuccess ||
            (e = cudaMemcpy(d.kmat, k.data(), k.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaEventCreate(&d.e0)) != cudaSuccess || (e = cudaEventCreate(&d.e1)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.ev_color, cudaEventDisableTiming)) != cudaSuccess)
            return bail(tc_fail(ISING_ECUDA, "tensor-core context: %s", cudaGetErrorString(e)));
        // TMA maps: 128-byte rows, 128 x 128 boxes, 128-byte swizzle
        const cuuint32_t box[2] = {SUBN, SUBN}, estr[2] = {1, 1};
        const cuuint64_t strides[1] = {SUBN};
        const cuuint64_t dl[2] = {SUBN, (cuuint64_t)(c->n_loc / SUBN)}, dk[2] = {SUBN, 5 * SUBN};
        const CUresult r1 = enc(&d.lat_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d.lat, dl, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult r2 = enc(&d.k_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d.kmat, dk, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return bail(tc_fail(ISING_ECUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2));
        c->base.slab_base[i] = d.lat;
    }
    // acceptance values with the device expf, then exact integer thresholds
    {
        TcDev& d = c->dev[0];
        cudaError_t e;
        float* d_acc = nullptr;
        if ((e = cudaSetDevice(d.device)) != cudaSuccess || (e = cudaMalloc(&d_acc, 5 * sizeof(float))) != cudaSuccess)
            return bail(tc_fail(ISING_ECUDA, "%s", cudaGetErrorString(e)));
        const float inv_temp = 1.0f / (alpha * ISING_CRIT_TEMP);   // tensorcore/main.cu:578
        tc_accept_kernel<<<1, 32, 0, d.stream>>>(inv_temp, d_acc);
        if ((e = cudaMemcpyAsync(c->accept, d_acc, 5 * sizeof(float), cudaMemcpyDeviceToHost, d.stream)) != cudaSuccess ||
            (e = cudaStreamSynchronize(d.stream)) != cudaSuccess)
            return bail(tc_fail(ISING_ECUDA, "%s", cudaGetErrorString(e)));
        cudaFree(d_acc);
    }
    c->base.nbx = nbx; c->base.nby = nby; c->base.nby_loc = c->nby_loc;
    c->base.key0 = (uint32_t)seed; c->base.key1 = (uint32_t)(seed >> 32);
    for (int kk = 0; kk < 5; kk++) {
        const uint64_t n = accept_count(c->accept[kk], true);   // uniform < accept (strict, main.cu:138)
        c->base.always[kk] = (n == (1ull << 32)) ? 0xffffffffu : 0u;
        c->base.never[kk] = (n == 0) ? 0xffffffffu : 0u;
        c->base.thr[kk] = (n == 0) ? 0u : (uint32_t)(n - 1);
    }
    // fast shape: v = -4, -2 always flip; v = 0, 2, 4 are plain thresholds (or "always")
    c->base.fast = (c->base.always[0] && c->base.always[1] && !c->base.never[2] && !c->base.never[3] && !c->base.never[4]) ? 1 : 0;
    *out = c;
    return ISING_OK;
}

// One colour on every device.  A device may start only when its ring neighbours have finished the
// previous colour (their edge rows are read directly); the reference synchronises all devices on
// the host instead (sync(), main.cu:228-235, :272, :297).
static int tc_launch_color(ising_tc_ctx* c, bool black, uint64_t iter, int8_t** nn_out)
{
    const int g = c->ngpus;
    for (int i = 0; i < g; i++) {
        TcDev& d = c->dev[i];
        TC_TRY(cudaSetDevice(d.device));
        if (g > 1) {
            TC_TRY(cudaStreamWaitEvent(d.stream, c->dev[(i + g - 1) % g].ev_color, 0));
            TC_TRY(cudaStreamWaitEvent(d.stream, c->dev[(i + 1) % g].ev_color, 0));
        }
        TcParams p = c->base;
        p.slab = i;
        p.is_black = black ? 1 : 0;
        p.call0 = (2ull * iter + (black ? 0 : 1)) * 4ull;   // offset (2*iter [+1])*16 draws, main.cu:269/:294
        const int ctas = 2 * c->nbx * c->nby_loc;
        int8_t* nn = nn_out ? nn_out[i] : nullptr;
        if (c->nn_path == ISING_TC_NN_MMA || c->nn_path == ISING_TC_NN_MMA_TILE) {
            const size_t smem = 4 * TILE_BYTES + 1024;
            TC_TRY(cudaFuncSetAttribute(tc_update_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            TC_TRY(cudaFuncSetAttribute(tc_update_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            TC_TRY(cudaFuncSetAttribute(tc_update_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            p.st = make_step_tables_at(c->seed, p.call0);
            if (nn) {
                tc_update_mma_kernel<true, false><<<ctas, MMA_THREADS, smem, d.stream>>>(d.lat_map, d.k_map, d.lat, nn, p);
            } else if (c->nn_kernel == 1 || c->nn_path == ISING_TC_NN_MMA_TILE) {   // one CTA per sub-block (A/B, validation)
                if (p.fast) tc_update_mma_kernel<false, true><<<ctas, MMA_THREADS, smem, d.stream>>>(d.lat_map, d.k_map, d.lat, nullptr, p);
                else        tc_update_mma_kernel<false, false><<<ctas, MMA_THREADS, smem, d.stream>>>(d.lat_map, d.k_map, d.lat, nullptr, p);
            } else {                          // product: persistent, warp-specialised pipeline
                TC_TRY(cudaFuncSetAttribute(tc_update_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
                TC_TRY(cudaFuncSetAttribute(tc_update_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
                const int grid = ctas < d.sms ? ctas : d.sms;   // one persistent CTA per SM
                if (p.fast) tc_update_pipe_kernel<true><<<grid, PIPE_THREADS, PIPE_SMEM, d.stream>>>(d.lat_map, d.k_map, d.lat, p, ctas);
                else        tc_update_pipe_kernel<false><<<grid, PIPE_THREADS, PIPE_SMEM, d.stream>>>(d.lat_map, d.k_map, d.lat, p, ctas);
            }
        } else {
            if (nn) tc_update_stencil_kernel<true><<<ctas, SUBN, 0, d.stream>>>(d.lat, nn, p);
            else    tc_update_stencil_kernel<false><<<ctas, SUBN, 0, d.stream>>>(d.lat, nullptr, p);
        }
        TC_TRY(cudaGetLastError());
        c->launches++;
    }
    if (g > 1)
        for (int i = 0; i < g; i++) {
            TC_TRY(cudaSetDevice(c->dev[i].device));
            TC_TRY(cudaEventRecord(c->dev[i].ev_color, c->dev[i].stream));
        }
    return ISING_OK;
}

extern "C" {

int ising_tc_create(int nbx, int nby, uint64_t seed, float alpha, int device, ising_tc_ctx** out)
{
    TC_RET(tc_create_impl(nbx, nby, seed, alpha, device, 1, out));
}

int ising_tc_create_multi(int nbx, int nby, uint64_t seed, float alpha, int first_device, int ngpus, ising_tc_ctx** out)
{
    TC_RET(tc_create_impl(nbx, nby, seed, alpha, first_device, ngpus, out));
}

int ising_tc_destroy(ising_tc_ctx* c)
{
    if (!c) return ISING_OK;
    for (TcDev& d : c->dev) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.lat) cudaFree(d.lat);
        if (d.kmat) cudaFree(d.kmat);
        if (d.scratch) cudaFree(d.scratch);
        if (d.d_sum) cudaFree(d.d_sum);
        if (d.e0) cudaEventDestroy(d.e0);
        if (d.e1) cudaEventDestroy(d.e1);
        if (d.ev_color) cudaEventDestroy(d.ev_color);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete c;
    return ISING_OK;
}

int ising_tc_set_nn_path(ising_tc_ctx* c, int path)
{
    if (!c || (path != ISING_TC_NN_MMA && path != ISING_TC_NN_STENCIL && path != ISING_TC_NN_MMA_TILE))
        TC_RET(tc_fail(ISING_EINVAL, "bad nn path"));
    c->nn_path = path;
    return ISING_OK;
}

static int tc_sync_impl(ising_tc_ctx* c)
{
    for (TcDev& d : c->dev) {
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaStreamSynchronize(d.stream));
    }
    return ISING_OK;
}

static int tc_init_impl(ising_tc_ctx* c)
{
    const uint32_t thr_half = (uint32_t)(accept_count(0.5f, true) - 1);
    for (int i = 0; i < c->ngpus; i++) {
        TcDev& d = c->dev[i];
        TC_TRY(cudaSetDevice(d.device));
        tc_init_kernel<<<148 * 8, 256, 0, d.stream>>>(d.lat, c->n_loc, (long long)i * c->n_loc, c->base.key0, c->base.key1, thr_half);
        TC_TRY(cudaGetLastError());
        c->launches++;
    }
    return tc_sync_impl(c);   // every slab is complete before any device reads a neighbour's edge
}
int ising_tc_init(ising_tc_ctx* c)
{
    if (!c) TC_RET(tc_fail(ISING_EINVAL, "null context"));
    TC_RET(tc_init_impl(c));
}

static int tc_update_impl(ising_tc_ctx* c, uint64_t first_iter, int n)
{
    for (int k = 0; k < n; k++) {
        if (int rc = tc_launch_color(c, true, first_iter + k, nullptr)) return rc;
        if (int rc = tc_launch_color(c, false, first_iter + k, nullptr)) return rc;
    }
    return ISING_OK;
}
int ising_tc_update(ising_tc_ctx* c, uint64_t first_iter, int n)
{
    if (!c || n < 0) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_update_impl(c, first_iter, n));
}

static int tc_timed_impl(ising_tc_ctx* c, uint64_t first_iter, int n, float* ms)
{
    if (int rc = tc_sync_impl(c)) return rc;
    for (TcDev& d : c->dev) {
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaEventRecord(d.e0, d.stream));
    }
    if (int rc = tc_update_impl(c, first_iter, n)) return rc;
    float worst = 0.f;
    for (TcDev& d : c->dev) {
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaEventRecord(d.e1, d.stream));
    }
    for (TcDev& d : c->dev) {
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaEventSynchronize(d.e1));
        float t = 0.f;
        TC_TRY(cudaEventElapsedTime(&t, d.e0, d.e1));
        if (t > worst) worst = t;
    }
    *ms = worst;
    return ISING_OK;
}
int ising_tc_update_timed(ising_tc_ctx* c, uint64_t first_iter, int n, float* ms)
{
    if (!c || !ms || n < 0) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_timed_impl(c, first_iter, n, ms));
}

int ising_tc_sync(ising_tc_ctx* c)
{
    if (!c) TC_RET(tc_fail(ISING_EINVAL, "null context"));
    TC_RET(tc_sync_impl(c));
}

static int tc_sum_impl(ising_tc_ctx* c, int64_t* sum)
{
    long long tot = 0;
    for (TcDev& d : c->dev) {
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaMemsetAsync(d.d_sum, 0, sizeof(long long), d.stream));
        tc_sum_kernel<<<148 * 8, 256, 0, d.stream>>>(d.lat, c->n_loc, d.d_sum);
        TC_TRY(cudaGetLastError());
        c->launches++;
    }
    for (TcDev& d : c->dev) {
        long long h = 0;
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaMemcpyAsync(&h, d.d_sum, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
        TC_TRY(cudaStreamSynchronize(d.stream));
        tot += h;
    }
    *sum = tot;
    return ISING_OK;
}
int ising_tc_sum(ising_tc_ctx* c, int64_t* sum)
{
    if (!c || !sum) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_sum_impl(c, sum));
}

static int tc_copy_impl(ising_tc_ctx* c, int8_t* out, const int8_t* in)
{
    for (int i = 0; i < c->ngpus; i++) {   // the slabs are consecutive pieces of the storage order
        TcDev& d = c->dev[i];
        TC_TRY(cudaSetDevice(d.device));
        if (out) TC_TRY(cudaMemcpyAsync(out + (size_t)i * c->n_loc, d.lat, (size_t)c->n_loc, cudaMemcpyDeviceToHost, d.stream));
        if (in) TC_TRY(cudaMemcpyAsync(d.lat, in + (size_t)i * c->n_loc, (size_t)c->n_loc, cudaMemcpyHostToDevice, d.stream));
    }
    return tc_sync_impl(c);
}
int ising_tc_export_storage(ising_tc_ctx* c, int8_t* out)
{
    if (!c || !out) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_copy_impl(c, out, nullptr));
}
int ising_tc_import_storage(ising_tc_ctx* c, const int8_t* in)
{
    if (!c || !in) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_copy_impl(c, nullptr, in));
}

// storage -> true order on the host (write_lattice, main.cu:317-336)
static void to_true_order(const ising_tc_ctx* c, const int8_t* st, int8_t* out)
{
    const long long nx = (long long)c->nbx * SUPN;
    for (int bj = 0; bj < c->nby; bj++)
        for (int bi = 0; bi < c->nbx; bi++)
            for (int b = 0; b < 2; b++)
                for (int a = 0; a < 2; a++) {
                    const int8_t* sub = st + sup_offset(bi, bj, c->nbx) + sub_offset(a, b);
                    for (int j = 0; j < SUBN; j++)
                        for (int i = 0; i < SUBN; i++)
                            out[((long long)bj * SUPN + 2 * j + b) * nx + (long long)bi * SUPN + 2 * i + a] = sub[j * SUBN + i];
                }
}

int ising_tc_export_true(ising_tc_ctx* c, int8_t* out)
{
    if (!c || !out) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    std::vector<int8_t> st((size_t)c->n);
    if (int rc = tc_copy_impl(c, st.data(), nullptr)) TC_RET(rc);
    to_true_order(c, st.data(), out);
    return ISING_OK;
}

int ising_tc_write_lattice(ising_tc_ctx* c, const char* fname)
{
    if (!c || !fname) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    std::vector<int8_t> t((size_t)c->n);
    if (int rc = ising_tc_export_true(c, t.data())) return rc;
    FILE* f = fopen(fname, "w");
    if (!f) TC_RET(tc_fail(ISING_EINVAL, "cannot open %s", fname));
    const long long nx = (long long)c->nbx * SUPN, ny = (long long)c->nby * SUPN;
    std::string line;
    for (long long r = 0; r < ny; r++) {
        line.clear();
        for (long long col = 0; col < nx; col++) line += (t[(size_t)(r * nx + col)] > 0) ? "1 " : "-1 ";
        line += "\n";
        fwrite(line.data(), 1, line.size(), f);
    }
    fclose(f);
    return ISING_OK;
}

int ising_tc_accept_table(ising_tc_ctx* c, float out[5])
{
    if (!c || !out) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    memcpy(out, c->accept, sizeof(c->accept));
    return ISING_OK;
}

static int tc_nn_impl(ising_tc_ctx* c, int is_black, int8_t* out)
{
    int8_t* scr[MAX_TC_SLABS];
    for (int i = 0; i < c->ngpus; i++) {
        TcDev& d = c->dev[i];
        TC_TRY(cudaSetDevice(d.device));
        if (!d.scratch) TC_TRY(cudaMalloc(&d.scratch, (size_t)c->n_loc));
        TC_TRY(cudaMemsetAsync(d.scratch, 0, (size_t)c->n_loc, d.stream));
        scr[i] = d.scratch;
    }
    if (int rc = tc_launch_color(c, is_black != 0, 0, scr)) return rc;
    for (int i = 0; i < c->ngpus; i++) {
        TcDev& d = c->dev[i];
        TC_TRY(cudaSetDevice(d.device));
        TC_TRY(cudaMemcpyAsync(out + (size_t)i * c->n_loc, d.scratch, (size_t)c->n_loc, cudaMemcpyDeviceToHost, d.stream));
    }
    return tc_sync_impl(c);
}
int ising_tc_neighbour_sums(ising_tc_ctx* c, int is_black, int8_t* out)
{
    if (!c || !out) TC_RET(tc_fail(ISING_EINVAL, "bad argument"));
    TC_RET(tc_nn_impl(c, is_black, out));
}

uint64_t ising_tc_launch_count(const ising_tc_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
