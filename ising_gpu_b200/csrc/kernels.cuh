// kernels.cuh -- parameter blocks shared by kernels.cu (device) and ising_api.cu (host).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox_map.h"

namespace ising {

// Tile of one CTA: 16 rows x 32 words (32 sites each) of the colour being updated == one
// thread block of the reference launch (optimized/main.cu:1547-1550), so that one thread
// owns exactly one reference thread id and its 16 Philox calls.
constexpr int TH = 16;
constexpr int TW = 32;
constexpr int SIDE_W = 4;           // 16-byte halo box left/right of the tile (TMA minimum)
constexpr int UPDATE_THREADS = 256;

// Acceptance rule in integer form.  For b = number of ANTI-aligned neighbours (0..4, i.e.
// 4-b aligned) a spin flips iff  always[b] | (selA[b] & x<=thrA) | (selB[b] & x<=thrB),
// x = the site's 32-bit Philox output.  Built on the host from exp_h (main.cu:1683-1697).
struct AcceptRule {
    uint32_t thrA, thrB;
    uint32_t always[5], selA[5], selB[5];   // 0 or 0xffffffff
    int fast;   // 1 when always = {0,0,1,1,1}, selA = b==1, selB = b==0 (every T > ~0.37)
};

struct UpdateParams {
    StepTables st;
    AcceptRule rule;
    uint32_t grid_x;        // X/2048 = reference gridDim.x
    uint32_t grow0;         // global row of this slab's local row 0 (devid*Y)
    int tile_row0;          // first tile row handled by blockIdx.y == 0
    int tile_row_alt;       // edges mode: tile row handled by blockIdx.y == 1
    int edges_mode;
    int launch_tile_rows;   // tile rows covered by this launch (grid y*z may overshoot)
    int rows;               // local rows Y
    int pitch;              // words per row
    int yperiod;            // vertical wrap period in rows inside the slab; 0 = ghost rows
    int xperiod;            // horizontal wrap period in words
    int color;
    int bond_rows;          // rows of one bond plane as passed to the kernel (Y or Y+2)
    // Fused halo push (multi-slab, peer-mapped neighbours): the CTAs owning the slab's first / last
    // row store the new row into the neighbour's ghost row as well and count themselves done on the
    // neighbour's arrival counter; before reading their own ghost row they wait for the neighbour's
    // CTAs of the previous colour step.  No host, copy engine or NCCL call in the sweep loop.
    int push;               // 0 = off; 1 = on; edges are then scheduled first (blockIdx.y 0 and 1)
    uint32_t* push_up;      // neighbour above: its bottom ghost row of the colour being updated
    uint32_t* push_dn;      // neighbour below: its top ghost row
    unsigned int* sig_up;   // arrival counter in the neighbour above (its "bottom ghost ready")
    unsigned int* sig_dn;   // arrival counter in the neighbour below (its "top ghost ready")
    const unsigned int* wait_top;   // own counters for the SOURCE colour's ghost rows
    const unsigned int* wait_bot;
    unsigned int wait_value;        // pushes the neighbours must have completed: steps * tiles_x
    unsigned int* err_flag;         // set to 1 when a bounded wait on a neighbour's counter times out
    unsigned long long spin_timeout_ns;
    // persistent kernel: the launch covers launch_tile_rows * tiles_x tiles, walked by every CTA with
    // stride gridDim.x; step_rows / step_cols = gridDim.x / tiles_x, gridDim.x % tiles_x (no division
    // in the loop).  edges_first: launch row 0 -> tile_row0, 1 -> tile_row_alt, r >= 2 -> r - 1.
    int tiles_x;
    int n_tiles;
    int step_rows, step_cols;
    int edges_first;
};

// Persistent kernel geometry: NSTAGE-deep ring of source-colour tiles in shared memory.
constexpr int UPDATE_STAGES = 8;
constexpr int UPDATE_PREFETCH = 4;   // tiles the TMA producer runs ahead of the consumers

struct InitParams {
    StepTables st;
    uint32_t thr_half;      // spin = 1 iff x <= thr_half  (uniform < 0.5f, main.cu:133)
    uint32_t grid_x;
    uint32_t grow0;
    int rows;
    int pitch;              // words per row
};

struct BondParams {
    uint32_t key0, key1;    // seed + 1
    uint32_t thr;           // bond bit set iff x <= thr (uniform < prob), unless always / never
    int always, never;
    uint32_t grid_x, grow0;
    long long total_rows;   // rows of the whole lattice (all slabs)
    int rows, pitch, yperiod, xperiod;
};

constexpr int MAX_CORR_SLABS = 16;
struct CorrParams {
    const uint32_t* black[MAX_CORR_SLABS];   // local row 0 of every slab of the lattice (peer-mapped)
    const uint32_t* white[MAX_CORR_SLABS];
    int self;                                // index of the slab this launch works on
    int rows, pitch, ncorr;
    long long grow0;
    long long yperiod;                       // rows per periodic unit (sub-lattice or whole lattice)
    int xperiod;                             // words per periodic unit
};

void launch_update(const CUtensorMap& src_tile, const CUtensorMap& src_row, const CUtensorMap& src_side,
                   uint32_t* dst_local_row0, const uint32_t* bond_planes, const UpdateParams& p, int tiles_x,
                   int tiles_y, cudaStream_t stream);
// two stacked tiles per CTA (32 rows), two reference thread ids per thread; tile_row0 of p in 16-row units,
// tile_rows32 = number of 32-row units covered; no bonds, no fused push
void launch_update2(const CUtensorMap& src_tile2, const CUtensorMap& src_row, const CUtensorMap& src_side2,
                    uint32_t* dst_local_row0, const UpdateParams& p, int tiles_x, int tile_rows32, cudaStream_t stream);
void launch_bond_init(uint32_t* random_planes_alloc_row0, uint32_t* mirror_planes_local_row0, const BondParams& p,
                      cudaStream_t stream);
void launch_corr(const CorrParams& p, unsigned long long* equal_counts, cudaStream_t stream);
void launch_init(uint32_t* plane, const InitParams& p, cudaStream_t stream);
// counts[0] += popcount over rows 0..rows-1 of the plane
void launch_count(const uint32_t* plane, int rows, int pitch, unsigned long long* counts, cudaStream_t stream);
// counts[0] += number of unequal nearest-neighbour pairs seen from the black sites
void launch_bond_differ(const uint32_t* black, const uint32_t* white, int rows, int pitch, int yperiod,
                        int xperiod, uint32_t grow0, unsigned long long* counts, cudaStream_t stream);
void launch_to_ref(const uint32_t* plane, int rows, int pitch, unsigned long long* out, cudaStream_t stream);
void launch_from_ref(const unsigned long long* in, int rows, int pitch, uint32_t* plane, cudaStream_t stream);
void launch_halo_wait(const unsigned int* flags, unsigned int want0, unsigned int want1, unsigned int* err_flag,
                      unsigned long long timeout_ns, cudaStream_t stream);
// persistent variant of launch_update: one CTA per resident slot (sms * ctas_per_sm), tiles by stride
void launch_update_persistent(const CUtensorMap& src_tile, const CUtensorMap& src_row, const CUtensorMap& src_side,
                              const CUtensorMap& dst_tile, uint32_t* dst_local_row0, const uint32_t* bond_planes,
                              UpdateParams p, int tiles_x, int tile_rows, int sms, cudaStream_t stream);
// variant: 0 = two tiles per CTA (kernel2), 1 = one tile per CTA, 2 = persistent
const char* update_kernel_name(int fast, int variant);

}  // namespace ising
