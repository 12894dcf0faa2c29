/*
 * ising_b200.h -- C ABI of libising_b200.so, the B200-native (sm_100a) checkerboard
 * Metropolis engine that drops in for the hot path of NVIDIA/ising-gpu's optimized/
 * program (reference files cited as optimized/main.cu:<line>).
 *
 * The reference has no FFI: its hot path sits behind `main()` of the cuIsing binary.
 * The entry points below are what that main() does between option parsing and printing,
 * one call per step of optimized/main.cu:1596-1925, so that a host written in any
 * language (the C++ CLI in ising_gpu_b200/csrc/cuising_main.cpp, the ctypes mirror in
 * ising_gpu_b200/engine.py, a cgo/JNI stub -- see INTEGRATION.md) can drive it.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ISING_E* code otherwise; the message
 *     is available from ising_last_error() (thread-local).  The library never calls exit():
 *     the reference's "message on stderr + exit(EXIT_FAILURE)" policy (cudamacro.h:25-39)
 *     is applied by the CLI, not here.
 *   - plain pointers and sizes only; host buffers are caller-owned, device memory, streams,
 *     events, peer mappings and tensor maps are library-owned.
 *   - a context is driven by one host thread at a time (reference: single host thread,
 *     main.cu:1764-1805).
 *   - results are bit-identical to the reference for the same (X, rows, seed, temperature,
 *     iteration number): the site -> (Philox subsequence, counter, lane) map of
 *     main.cu:112-116, :514-515, :620-660 is kept; the device-internal packing is 1 bit/spin.
 */
#ifndef ISING_B200_H_
#define ISING_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: + ISING_ETIMEOUT, ising_get_temperature, ising_host_alloc/free, ising_run_batch_from_host_n; checkpoint format 2;
 *    ising_update_color_part refused on contexts with the fused halo push */
#define ISING_B200_ABI_VERSION 2

#define ISING_MAX_DEVICES 256          /* MAX_GPU, optimized/main.cu:60 */
#define ISING_MAX_CORR_DEVICES 16      /* ising_correlation: devices of one context                     */
#define ISING_MAX_CORR_ROW_BYTES 204800 /* ising_correlation: X / 8 (one lattice row at 1 bit/spin) limit  */
#define ISING_DEFAULT_SEED 463463564571ull /* SEED_DEF, optimized/main.cu:63 */
#define ISING_CRIT_TEMP 2.26918531421f /* CRIT_TEMP, optimized/main.cu:42 */

enum {
    ISING_OK = 0,
    ISING_EINVAL = -1,  /* bad argument / lattice shape (main.cu:1412-1453)                 */
    ISING_ECUDA = -2,   /* a CUDA runtime/driver call failed (CHECK_CUDA, cudamacro.h:25)    */
    ISING_ENOPEER = -3, /* missing direct link between devices (main.cu:1531-1536)           */
    ISING_ENOMEM = -4,
    ISING_ESTATE = -5,  /* call not valid in this state (e.g. halo access on a 1-slab ctx)   */
    ISING_ETIMEOUT = -6 /* a device-side wait for a neighbouring slab's halo rows gave up    */
};

enum { ISING_BLACK = 0, ISING_WHITE = 1 }; /* enum {C_BLACK, C_WHITE}, main.cu:80 */

typedef struct ising_ctx ising_ctx;

/* Configuration = the scalars main() derives from its options (main.cu:1245-1283,
 * :1395-1471).  Zero-initialise, then fill. */
typedef struct ising_cfg {
    int32_t X;            /* -x: lattice columns, black+white; multiple of 2048 (main.cu:1412) */
    int32_t Y;            /* -y: lattice rows PER DEVICE; multiple of 16 (main.cu:1417)        */
    int32_t ndev;         /* -d: devices driven by this context (>=1)                          */
    int32_t first_device; /* CUDA ordinal of device 0 of the context (reference: always 0)      */
    uint64_t seed;        /* -s (main.cu:1329)                                                  */
    float temp;           /* absolute temperature (main.cu:1465-1471 resolves -a / -t)          */
    int32_t xsl;          /* --xsl sub-lattice columns, 0 = disabled (main.cu:1423-1463)        */
    int32_t ysl;          /* --ysl sub-lattice rows,    0 = disabled                            */
    /* Placement of this context's slabs inside a larger lattice that is spread over several
     * processes (one process per GPU under torchrun).  The reference is single-process
     * (devid = 0..ndev-1, main.cu:1764); slab s of the global lattice is what the reference
     * calls devid.  slab_count == 0 means "this context is the whole lattice". */
    int32_t slab_first;   /* global slab index of this context's device 0                      */
    int32_t slab_count;   /* slabs in the global lattice (0 -> ndev)                            */
    int32_t flags;        /* ISING_F_* */
    float j_prob;         /* -J: probability that a bond is antiferromagnetic (with ISING_F_BONDS) */
    int32_t reserved[4];
} ising_cfg;

enum {
    ISING_F_NO_PEER_COPY = 1, /* multi-slab context whose ghost rows are filled by the caller
                                 (NCCL send/recv or IPC stores), not by the library           */
    ISING_F_BONDS = 2         /* -J: random +-J bonds, hamiltInitB_k/hamiltInitW_k (main.cu:153-331,
                                 applied in the update as main.cu:575-618)                    */
};

/* ---- lifetime --------------------------------------------------------------------------
 * ising_create replaces main.cu:1481-1537 (device checks, peer enabling) and :1596-1706
 * (allocations, acceptance table).  ising_destroy replaces :1900-1925 (without
 * cudaDeviceReset: the process may hold other CUDA state). */
int ising_create(const ising_cfg* cfg, ising_ctx** out);
int ising_destroy(ising_ctx* ctx);
const char* ising_last_error(void);
int ising_abi_version(void);

/* Use caller-provided streams (e.g. torch.cuda.current_stream().cuda_stream) instead of the
 * context's own; stream is a cudaStream_t passed as void*.  dev = index inside the context. */
int ising_set_stream(ising_ctx* ctx, int dev, void* stream);

/* ---- the hot path ----------------------------------------------------------------------*/
/* latticeInit_k x2 (main.cu:1708-1726): random spins from Philox step it=0. */
int ising_init_lattice(ising_ctx* ctx);

/* One colour half-sweep on every device of the context: spinUpdateV_2D_k<COLOR> with
 * iteration number `it` (main.cu:1766-1777 / :1787-1798; it = j+1 for sweep j) followed by
 * the halo push of that colour's boundary rows.  Asynchronous. */
int ising_update_color(ising_ctx* ctx, int color, int it);

/* n full sweeps, iteration numbers first_it .. first_it+n-1 (the reference's loop
 * main.cu:1763-1805 uses first_it = 1).  Asynchronous; ising_sync waits. */
int ising_sweeps(ising_ctx* ctx, int first_it, int n);
int ising_sync(ising_ctx* ctx);

/* Same as ising_sweeps + ising_sync, bracketed by CUDA events on the launching stream(s)
 * (main.cu:1758, :1863, :1879); *ms = elapsed device time, max over the context's devices. */
int ising_sweeps_timed(ising_ctx* ctx, int first_it, int n, float* ms);

/* Rebuild the acceptance table for a new temperature (main.cu:1848-1859). */
int ising_set_temperature(ising_ctx* ctx, float temp);
/* temperature currently in force (after ising_set_temperature or ising_load_checkpoint) */
int ising_get_temperature(const ising_ctx* ctx, float* temp);

/* ---- observables -----------------------------------------------------------------------*/
/* countSpins / getMagn_k (main.cu:831-868): #up and #down over the context's slabs. */
int ising_count_spins(ising_ctx* ctx, uint64_t* up, uint64_t* down);
/* Sum over nearest-neighbour bonds of s_i*s_j (s = +-1) for the context's slabs; energy per
 * site = -bond_sum / (rows*X).  (The reference exposes this only via -c, main.cu:1129-1135.) */
int ising_bond_sum(ising_ctx* ctx, int64_t* bond_sum);

/* Two-point correlation (-c; computeCorr + getCorr2D_k / getCorr2DRepl_k, main.cu:870-1138):
 * out[j-1], j = 1..ncorr (the reference uses 128), = sum over all sites of (+1 if equal else -1)
 * for the site j columns to the right and the site j rows below, periodic inside the
 * (sub-)lattice, divided by 2*X*rows -- the numbers the reference writes to its corr_ file.
 * Single-process contexts only, up to ISING_MAX_CORR_DEVICES devices and X / 8 <= ISING_MAX_CORR_ROW_BYTES
 * (ISING_ESTATE otherwise; the CLI checks both when it parses -c).  For X (or xsl) not a multiple of 4096 the
 * reference's chunk loop (main.cu:905) counts columns twice and reads past the row; this returns the exact
 * periodic two-point function instead (the CLI prints a warning). */
int ising_correlation(ising_ctx* ctx, int ncorr, double* out);
/* +-J bond nibbles of one device's slab in the reference's 4-bit words (bit 3 up, 2 down, 1 left,
 * 0 right): `random` = the buffer hamiltInitB_k fills (used by the WHITE update, main.cu:1796),
 * `mirror` = the one hamiltInitW_k derives (used by the BLACK update, main.cu:1775).  For tests. */
int ising_export_bonds_ref_layout(ising_ctx* ctx, int dev, uint64_t* random, uint64_t* mirror);

/* ---- lattice transfer ------------------------------------------------------------------*/
/* Reference storage format of one device's slab (main.cu:1160-1175): Y*X/32 uint64 words per
 * colour, 16 spins per word, nibble k = colour-column 16*j+k, bit 0 of the nibble = spin. */
int ising_export_ref_layout(ising_ctx* ctx, int dev, uint64_t* black, uint64_t* white);
int ising_import_ref_layout(ising_ctx* ctx, int dev, const uint64_t* black, const uint64_t* white);
/* Device packing: Y*X/64 uint32 words per colour, bit b of word w = colour-column 32*w+b. */
int ising_export_bits(ising_ctx* ctx, int dev, uint32_t* black, uint32_t* white);
int ising_import_bits(ising_ctx* ctx, int dev, const uint32_t* black, const uint32_t* white);
/* Binary checkpoint of every device of the context (the reference has none): a header
 * (shape, slabs, seed, temperature, -J flag and probability, sweeps done) and the bit-packed colour planes.
 * Loading checks the shape, the seed and the Hamiltonian (-J), restores the saved temperature
 * (ising_get_temperature tells the caller) -- the continuation draws from the same Philox stream, so `n1`
 * sweeps, save, load, `n2` sweeps starting at iteration n1 + 1 equal n1 + n2 sweeps -- and refreshes the
 * ghost rows between the context's own devices.  iterations_done may be NULL. */
int ising_save_checkpoint(ising_ctx* ctx, const char* path, uint64_t iterations_done);
int ising_load_checkpoint(ising_ctx* ctx, const char* path, uint64_t* iterations_done);
/* dumpLattice (main.cu:1140-1209): writes "<prefix><dev>.txt" for every device of the context. */
int ising_dump_lattice(ising_ctx* ctx, const char* prefix);

/* End-to-end convenience used by bench.py's e2e leg: upload both colour planes of device 0
 * from host (bit packing), run n sweeps starting at first_it, download the planes and the
 * spin counts.  Everything (H2D, kernels, D2H) happens inside the call. */
int ising_run_from_host(ising_ctx* ctx, const uint32_t* black_in, const uint32_t* white_in,
                        int first_it, int n, uint32_t* black_out, uint32_t* white_out,
                        uint64_t* up, uint64_t* down);

/* Replica batch (ensembles / parameter scans: many independent lattices of the same shape, the
 * way the reference is used from scripts, optimized/README.md:94-148): lattice k is uploaded from
 * black_in[k]/white_in[k] (device bit packing, pinned host memory for full overlap), evolved by n
 * sweeps starting at first_it, counted into up[k]/down[k] and downloaded into black_out[k]/
 * white_out[k] (entries or the arrays themselves may be NULL).  Jobs alternate between the two
 * contexts (same shape, same device; ctx_b may be NULL = no overlap), each on its own stream, so
 * the copies of one job overlap the kernels of the other.  Blocks until the batch is done. */
int ising_run_batch_from_host(ising_ctx* ctx_a, ising_ctx* ctx_b, int nrep, const uint32_t* const* black_in,
                              const uint32_t* const* white_in, int first_it, int n, uint32_t* const* black_out,
                              uint32_t* const* white_out, uint64_t* up, uint64_t* down);

/* The same for one slab of a lattice split over several processes (one per GPU, peers attached with
 * ising_ipc_attach, fused halo push): every rank calls it with its own slab of replica k and the
 * white ghost rows that go with it, ghosts_in[k] = 2 rows of X/64 words: the white row just above
 * the slab and the one just below (the neighbouring slabs' boundary rows of the same replica; the
 * black ones are not needed, a sweep starts with the black update, which reads white only).  The
 * halo of the sweeps travels from kernel to kernel as in ising_sweeps; before a context's ghost rows
 * are overwritten with the next replica's, a device-side wait makes sure the neighbours' last pushes
 * of the previous one have landed.  up[k]/down[k] count this slab only. */
int ising_run_slab_batch_from_host(ising_ctx* ctx_a, ising_ctx* ctx_b, int nrep, const uint32_t* const* black_in,
                                   const uint32_t* const* white_in, const uint32_t* const* ghosts_in, int first_it, int n,
                                   uint32_t* const* black_out, uint32_t* const* white_out, uint64_t* up, uint64_t* down);
/* The same batches over nctx (1..8) contexts of one shape on one device, job k on ctxs[k % nctx].  With two contexts
 * the download of job k and the upload of job k+2 share a stream and must BOTH fit into one job's compute time; with
 * three they have two -- what it takes when many GPUs share the host's memory and PCIe fabric (8 GPUs: bench.py).
 * ghosts_in NULL = single-slab contexts (ising_run_batch_from_host), else slab contexts with attached peers. */
int ising_run_batch_from_host_n(ising_ctx* const* ctxs, int nctx, int nrep, const uint32_t* const* black_in,
                                const uint32_t* const* white_in, const uint32_t* const* ghosts_in, int first_it, int n,
                                uint32_t* const* black_out, uint32_t* const* white_out, uint64_t* up, uint64_t* down);

/* Page-locked host staging memory for the *_from_host calls (cudaHostAlloc, portable).  write_combined != 0 asks for
 * write-combined memory: faster for the device to read over PCIe and it does not pollute the CPU caches, but slow for
 * the CPU to read -- use it for buffers the host only writes (inputs), never for outputs. */
int ising_host_alloc(size_t bytes, int write_combined, void** ptr);
int ising_host_free(void* ptr);

/* ---- multi-process slabs (one process per GPU) ------------------------------------------
 * Device pointer and byte length of one row of a colour plane.  row = -1 is the ghost row
 * above the slab, row = Y the ghost row below it, 0..Y-1 the slab's own rows.  The caller
 * fills ghost rows (from the neighbouring slabs' rows Y-1 / 0) between colour updates. */
int ising_row_ptr(ising_ctx* ctx, int dev, int color, int row, void** dptr, size_t* nbytes);
/* Update only the tile rows that touch the slab edge (first/last 16 rows) or only the rest;
 * lets the caller overlap the halo exchange with the interior (see engine.SlabEngine).
 * Only for contexts whose ghost rows the CALLER fills (ISING_F_NO_PEER_COPY, no ising_ipc_attach):
 * on a context with the fused halo push it returns ISING_ESTATE (the push of a partial launch
 * would not be counted and the ghost-row ordering would be lost). */
enum { ISING_PART_ALL = 0, ISING_PART_EDGES = 1, ISING_PART_INTERIOR = 2 };
int ising_update_color_part(ising_ctx* ctx, int dev, int color, int it, int part, void* stream);

/* ---- device queries used by the CLI banner (main.cu:1483-1536) ---------------------------*/
int ising_device_count(int* count);
/* name (up to name_len bytes), #SMs, max threads per SM, compute capability, ECC flag */
int ising_device_info(int device, char* name, size_t name_len, int* sms, int* max_threads_per_sm,
                      int* cc_major, int* cc_minor, int* ecc);
/* *ok = 1 when `device` can map `peer`'s memory (cudaDeviceCanAccessPeer) */
int ising_can_access_peer(int device, int peer, int* ok);

/* Fused halo push between processes: the update kernel of a slab stores its first/last row
 * straight into the neighbouring slabs' ghost rows (NVLink peer stores) and signals a counter in
 * their memory; no NCCL / copy-engine / host step per colour.  Each rank exports CUDA IPC handles
 * (3 x 64 bytes: black plane, white plane, counters), ships them to its ring neighbours by any
 * means (bench.py: torch.distributed all_gather) and attaches the two neighbours' handles.  After
 * attach, ising_push_edges(color) fills the neighbours' ghost rows once (init / import; follow it
 * with a barrier between the ranks), then ising_update_color / ising_sweeps need no further help.
 * Single-process contexts with ndev > 1 use the same kernel path through peer access
 * automatically (set ISING_B200_HALO=copy to fall back to cudaMemcpyPeerAsync per colour). */
int ising_ipc_export(ising_ctx* ctx, void* handles, size_t nbytes);
int ising_ipc_attach(ising_ctx* ctx, const void* up_handles, const void* down_handles);
int ising_push_edges(ising_ctx* ctx, int color);

/* ---- introspection used by tests and bench ----------------------------------------------*/
/* Number of 32-bit Philox outputs x accepted by `uniform(x) <= p` (strict = 0) or
 * `uniform(x) < p` (strict = 1), uniform(x) = x*2^-32 + 2^-33 in float (curand_uniform).
 * Host-only arithmetic: valid without a GPU.  Range 0 .. 2^32. */
uint64_t ising_accept_count(float p, int strict);
/* The 5 acceptance probabilities by number of ALIGNED neighbours a = 0..4 for temperature
 * temp, i.e. exp_h[1][a] of main.cu:1683-1697 (host expf). */
void ising_accept_table(float temp, float out[5]);
/* The four raw Philox outputs that decide sites 32*word + {2q, 16+2q, 2q+1, 17+2q} of colour
 * `color` in global row `grow` at iteration `it` (0 = lattice init) for an X-column lattice,
 * computed on the host through the same folded schedule the kernels use (philox_map.h).
 * Lets tests check the site -> (subsequence, counter, lane) contract without a GPU. */
void ising_site_random(uint64_t seed, int32_t X, int it, int color, uint32_t grow, uint32_t word, int q,
                       uint32_t out[4]);
/* kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t ising_launch_count(const ising_ctx* ctx);
/* name of the update kernel variant in use, for logs */
const char* ising_kernel_name(const ising_ctx* ctx);

/* ==========================================================================================
 * Tensor-core variant: drop-in for the hot path of the reference's tensorcore/ program
 * (tensorcore/main.cu): +-1 lattice in 256x256 super-blocks of four 128x128 sub-blocks
 * (:37-44), neighbour sums as matrix products with a bidiagonal 0/1 matrix (:252-262, :547-569)
 * plus block-edge terms (add_boundaries :148-226), Metropolis with its own RNG map
 * (update_spins :88-146, init_spins :63-79).  Here the sums are an INT8 tensor-core
 * contraction fused with the update in one kernel; nn_sums is never written to memory.
 * ========================================================================================== */
typedef struct ising_tc_ctx ising_tc_ctx;

enum { ISING_TC_NN_MMA = 0,      /* INT8 tensor-core contraction, persistent warp-specialised kernel (product path) */
       ISING_TC_NN_STENCIL = 1,  /* plain integer 4-neighbour stencil (validation of the MMA)                       */
       ISING_TC_NN_MMA_TILE = 2  /* INT8 tensor-core contraction, one CTA per 128x128 sub-block (A/B, validation)   */ };

/* nbx, nby: lattice columns / rows in units of 256 (-x, -y of ising_tensorcore, main.cu:402-451);
 * alpha: temperature in T_crit units (-a); seed (-s, default 1234). */
int ising_tc_create(int nbx, int nby, uint64_t seed, float alpha, int device, ising_tc_ctx** out);
/* -g NGPUS (main.cu:447-451): the super-block rows are split over ngpus devices of this process
 * (nby % ngpus == 0), each device updates its slab and reads the neighbouring slabs' edge rows
 * through peer access; devices first_device .. first_device+ngpus-1. */
int ising_tc_create_multi(int nbx, int nby, uint64_t seed, float alpha, int first_device, int ngpus, ising_tc_ctx** out);
int ising_tc_destroy(ising_tc_ctx* ctx);
int ising_tc_set_nn_path(ising_tc_ctx* ctx, int path);
/* init_spins (main.cu:492-497) */
int ising_tc_init(ising_tc_ctx* ctx);
/* n calls of update() (main.cu:237-298) with iteration numbers first_iter .. first_iter+n-1
 * (the reference numbers warm-up and trial iterations consecutively from 1, :582-593). Async. */
int ising_tc_update(ising_tc_ctx* ctx, uint64_t first_iter, int n);
int ising_tc_update_timed(ising_tc_ctx* ctx, uint64_t first_iter, int n, float* ms);
int ising_tc_sync(ising_tc_ctx* ctx);
/* sum of all spins (the reference's final CUB reduce, main.cu:611-636) */
int ising_tc_sum(ising_tc_ctx* ctx, int64_t* sum);
/* lattice as int8 +-1: storage (blocked) order, or true row-major order ny x nx (main.cu:317-336) */
int ising_tc_export_storage(ising_tc_ctx* ctx, int8_t* out);
int ising_tc_import_storage(ising_tc_ctx* ctx, const int8_t* in);
int ising_tc_export_true(ising_tc_ctx* ctx, int8_t* out);
/* write_lattice (main.cu:300-352): "<spin> " per site, one line per row */
int ising_tc_write_lattice(ising_tc_ctx* ctx, const char* fname);
/* the five acceptance values exp(-2*inv_temp*v), v = nn*s = -4,-2,0,2,4, evaluated on the device
 * with the same expf the reference's update_spins uses (main.cu:137) */
int ising_tc_accept_table(ising_tc_ctx* ctx, float out[5]);
/* neighbour sums of one colour in storage order (0 elsewhere) as the selected nn path computes
 * them -- what main.cu:252-266 leaves in nn_sums; for tests */
int ising_tc_neighbour_sums(ising_tc_ctx* ctx, int is_black, int8_t* out);
uint64_t ising_tc_launch_count(const ising_tc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ISING_B200_H_ */
