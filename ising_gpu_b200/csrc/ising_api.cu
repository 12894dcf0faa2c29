// ising_api.cu -- C ABI of libising_b200.so (include/ising_b200.h): contexts, device memory,
// tensor maps, streams, peer halo copies, host<->device transfers.  Host-side counterpart of
// optimized/main.cu:1481-1925 of the reference, minus option parsing and printing (those live
// in cuising_main.cpp).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ising_b200.h"
#include "host_math.h"
#include "kernels.cuh"

using namespace ising;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(call)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ISING_ECUDA, "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

struct DevState {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    uint32_t* plane[2] = {nullptr, nullptr};   // (Y+2) x pitch words; row 0 / Y+1 are ghost rows
    CUtensorMap map_tile[2], map_row[2], map_side[2];
    CUtensorMap map_tile2[2], map_side2[2];    // 32-row boxes of the two-tiles-per-CTA kernel
    uint32_t* bond_random = nullptr;           // 4 planes x (Y+2) x pitch (hamiltInitB_k), with edge rows
    uint32_t* bond_mirror = nullptr;           // 4 planes x Y x pitch (hamiltInitW_k)
    unsigned long long* d_corr = nullptr;
    unsigned long long* d_counts = nullptr;    // [0] up-count, [1] unequal bonds
    // fused halo push: arrival counters written by the neighbours (recv_top[2], recv_bot[2]) and
    // the neighbours' planes / counters as seen from this device (peer access or CUDA IPC)
    unsigned int* d_flags = nullptr;
    uint32_t* up_plane[2] = {nullptr, nullptr};
    uint32_t* dn_plane[2] = {nullptr, nullptr};
    unsigned int* up_flags = nullptr;
    unsigned int* dn_flags = nullptr;
    void* ipc_opened[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int sms = 148;                             // multiprocessors: the persistent update kernel runs sms x 4 CTAs
    // fused push: the two edge tile rows run on their own high-priority stream next to the interior launch
    cudaStream_t stream_edge = nullptr;
    cudaEvent_t ev_pre = nullptr, ev_edge = nullptr;
    cudaEvent_t ev_color = nullptr;            // kernel + halo pushes of the last colour issued
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
};

}  // namespace

struct ising_ctx {
    ising_cfg cfg;
    int pitch = 0;        // 32-site words per colour row = X/64
    int slab_count = 1;
    bool bonds = false;
    bool fused = false;       // halo rows are pushed by the update kernel itself (peer stores + counters)
    bool persistent = false;  // ISING_B200_KERNEL=persistent selects the persistent kernel (A/B; measured 3 % slower)
    bool two_tiles = true;    // ISING_B200_KERNEL=tile1 forces the one-tile-per-CTA kernel everywhere (A/B)
    bool split_edges = true;  // tile kernel, fused push: edge tile rows in their own launch, not inside the interior's
    bool edge_stream = true;  // ... and that launch on a second, high-priority stream, concurrent with the interior
    unsigned long long spin_timeout_ns = 30ull * 1000000000ull;   // bound of every device-side wait on a neighbour
    uint64_t pushes_done[2] = {0, 0};
    bool peer_halo = false;   // library copies boundary rows between its own devices
    bool ghost_rows = false;  // vertical neighbours of the slab edge come from ghost rows
    int yperiod = 0, xperiod = 0;
    AcceptRule rule;
    uint32_t thr_half = 0;
    uint64_t launches = 0;
    // replica batches: per-job up-counts on the device and their pinned landing zone, kept between calls (allocating
    // and freeing page-locked memory inside every batch call is a device-wide synchronisation in the timed path)
    unsigned long long* batch_h_counts = nullptr;
    unsigned long long* batch_d_counts = nullptr;
    int batch_counts_cap = 0;
    std::vector<DevState> dev;
};

namespace {

constexpr int FLAG_WORDS = 16;      // d_flags: [0..3] arrival counters, [ERR_FLAG_WORD] error word
constexpr int ERR_FLAG_WORD = 8;

uint32_t* local_row(const ising_ctx* c, const DevState& d, int color, int row)
{
    return d.plane[color] + (size_t)(row + 1) * c->pitch;
}

int make_map(CUtensorMap* m, uint32_t* base, int pitch, int rows_alloc, int bw, int bh)
{
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(ISING_ECUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows_alloc};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ISING_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ISING_OK;
}

int validate(const ising_cfg* c)
{
    // main.cu:1412-1421
    if (c->X <= 0 || (c->X % 2048)) return fail(ISING_EINVAL, "X must be a positive multiple of 2048 (got %d)", c->X);
    if (c->Y <= 0 || (c->Y % 16)) return fail(ISING_EINVAL, "Y must be a positive multiple of 16 (got %d)", c->Y);
    if (c->ndev < 1 || c->ndev > ISING_MAX_DEVICES) return fail(ISING_EINVAL, "ndev out of range (%d)", c->ndev);
    // main.cu:1440-1453
    if (c->xsl || c->ysl) {
        if (c->xsl <= 0 || (c->X % c->xsl) || (c->xsl % 2048))
            return fail(ISING_EINVAL, "xsl must be a multiple of 2048 and a divisor of X (got %d)", c->xsl);
        if (c->ysl <= 0 || (c->Y % c->ysl) || (c->ysl % 16))
            return fail(ISING_EINVAL, "ysl must be a multiple of 16 and a divisor of Y (got %d)", c->ysl);
    }
    if (c->slab_count < 0 || c->slab_first < 0) return fail(ISING_EINVAL, "negative slab placement");
    if (c->slab_count && c->slab_first + c->ndev > c->slab_count)
        return fail(ISING_EINVAL, "slab_first + ndev exceeds slab_count");
    return ISING_OK;
}

UpdateParams base_params(const ising_ctx* c, int dev, int color, int it)
{
    UpdateParams p{};
    p.st = make_step_tables(c->cfg.seed, 2ull * (uint64_t)it + (uint64_t)color);
    p.rule = c->rule;
    p.grid_x = (uint32_t)(c->cfg.X / 2048);
    p.grow0 = (uint32_t)(c->cfg.slab_first + dev) * (uint32_t)c->cfg.Y;
    p.rows = c->cfg.Y;
    p.pitch = c->pitch;
    p.yperiod = c->yperiod;
    p.xperiod = c->xperiod;
    p.color = color;
    p.bond_rows = (color == 0) ? c->cfg.Y : c->cfg.Y + 2;   // black uses the mirror planes, white the random ones
    return p;
}

int launch_part(ising_ctx* c, int dev, int color, int it, int part, cudaStream_t stream)
{
    DevState& d = c->dev[dev];
    UpdateParams p = base_params(c, dev, color, it);
    const int tiles_x = c->pitch / TW, tiles_y = c->cfg.Y / TH;
    const int src = color ^ 1;
    int gy = tiles_y;
    if (part == ISING_PART_EDGES) {
        p.edges_mode = 1; p.tile_row0 = 0; p.tile_row_alt = tiles_y - 1;
        gy = tiles_y > 1 ? 2 : 1;
    } else if (part == ISING_PART_INTERIOR) {
        p.tile_row0 = 1; gy = tiles_y - 2;
        if (gy <= 0) return ISING_OK;
    }
    p.launch_tile_rows = gy;
    p.err_flag = d.d_flags ? d.d_flags + ERR_FLAG_WORD : nullptr;
    p.spin_timeout_ns = c->spin_timeout_ns;
    if (c->persistent) {
        // edge tile rows are the first tiles of the walk whenever the launch owns them: the fused halo
        // push leaves in the first iteration of the same launch that does the interior
        p.edges_mode = 0;
        p.edges_first = (part == ISING_PART_EDGES) || (part == ISING_PART_ALL && c->fused);
        if (p.edges_first) { p.tile_row0 = 0; p.tile_row_alt = tiles_y - 1; }
    }
    if (c->fused && part != ISING_PART_INTERIOR) {   // the launch that owns the edge rows pushes them
        const int Y = c->cfg.Y;
        p.push = 1;
        p.push_up = d.up_plane[color] + (size_t)(Y + 1) * c->pitch;   // neighbour above: its bottom ghost row
        p.push_dn = d.dn_plane[color];                                // neighbour below: its top ghost row
        p.sig_up = d.up_flags + 2 + color;                            // its recv_bot[color]
        p.sig_dn = d.dn_flags + 0 + color;                            // its recv_top[color]
        p.wait_top = d.d_flags + 0 + src;
        p.wait_bot = d.d_flags + 2 + src;
        p.wait_value = (unsigned int)(c->pushes_done[src] * (uint64_t)tiles_x);
    }
    CU_TRY(cudaSetDevice(d.device));
    // the reference hands the BLACK update the hamiltInitW_k buffer and the WHITE update the
    // hamiltInitB_k one (main.cu:1775, :1796)
    const uint32_t* bonds = nullptr;
    if (c->bonds) bonds = (color == 0) ? d.bond_mirror : d.bond_random + c->pitch;   // local row 0
    // two stacked tiles per CTA wherever the launch covers whole 32-row units of one periodic domain and needs
    // neither bonds nor the push code (the interior of a slab, or a whole single-slab lattice)
    const bool two = c->two_tiles && !c->persistent && !bonds && !p.push && part != ISING_PART_EDGES && (gy % 2) == 0 &&
                     (c->yperiod == 0 || (c->yperiod % (2 * TH) == 0 && p.tile_row0 % 2 == 0));
    if (two) p.launch_tile_rows = gy / 2;   // counted in 32-row units by that kernel
    if (two)
        launch_update2(d.map_tile2[src], d.map_row[src], d.map_side2[src], local_row(c, d, color, 0), p, tiles_x, gy / 2, stream);
    else if (c->persistent)
        launch_update_persistent(d.map_tile[src], d.map_row[src], d.map_side[src], d.map_tile[color],
                                 local_row(c, d, color, 0), bonds, p, tiles_x, gy, d.sms, stream);
    else
        launch_update(d.map_tile[src], d.map_row[src], d.map_side[src], local_row(c, d, color, 0), bonds, p, tiles_x, gy,
                      stream);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return ISING_OK;
}

// Copy the freshly updated colour's first/last row into the neighbouring slabs' ghost rows
// (replaces the implicit remote loads of main.cu:413-428 + the host barrier :1779-1784).
int push_halo(ising_ctx* c, int color)
{
    const int n = (int)c->dev.size();
    const size_t row_bytes = (size_t)c->pitch * 4;
    for (int i = 0; i < n; i++) {
        DevState& d = c->dev[i];
        DevState& up = c->dev[(i + n - 1) % n];
        DevState& dn = c->dev[(i + 1) % n];
        CU_TRY(cudaSetDevice(d.device));
        // my row 0 -> ghost row below the slab above me; my row Y-1 -> ghost row above the slab below me
        CU_TRY(cudaMemcpyPeerAsync(local_row(c, up, color, c->cfg.Y), up.device, local_row(c, d, color, 0), d.device,
                                   row_bytes, d.stream));
        CU_TRY(cudaMemcpyPeerAsync(local_row(c, dn, color, -1), dn.device, local_row(c, d, color, c->cfg.Y - 1),
                                   d.device, row_bytes, d.stream));
        CU_TRY(cudaEventRecord(d.ev_color, d.stream));
    }
    for (int i = 0; i < n; i++) {
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaStreamWaitEvent(d.stream, c->dev[(i + n - 1) % n].ev_color, 0));
        CU_TRY(cudaStreamWaitEvent(d.stream, c->dev[(i + 1) % n].ev_color, 0));
    }
    return ISING_OK;
}

}  // namespace

extern "C" {

// shared by the tensor-core variant's translation unit (not part of the public header)
void ising_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

const char* ising_last_error(void) { return g_err.c_str(); }
int ising_abi_version(void) { return ISING_B200_ABI_VERSION; }
uint64_t ising_accept_count(float p, int strict) { return accept_count(p, strict != 0); }
void ising_accept_table(float temp, float out[5]) { accept_table(temp, out); }
uint64_t ising_launch_count(const ising_ctx* ctx) { return ctx ? ctx->launches : 0; }
// name of the kernel that updates the bulk of this context's lattice (launch_part picks per launch)
const char* ising_kernel_name(const ising_ctx* ctx)
{
    if (!ctx) return update_kernel_name(1, 0);
    const int tiles_y = ctx->cfg.Y / TH;
    const int bulk_rows = ctx->fused ? tiles_y - 2 : tiles_y;     // fused: the interior launch
    const bool two = ctx->two_tiles && !ctx->persistent && !ctx->bonds && bulk_rows > 0 && (bulk_rows % 2) == 0 &&
                     (ctx->yperiod == 0 || ctx->yperiod % (2 * TH) == 0);
    return update_kernel_name(ctx->rule.fast, two ? 0 : (ctx->persistent ? 2 : 1));
}

int ising_device_count(int* count)
{
    if (!count) return fail(ISING_EINVAL, "null argument");
    CU_TRY(cudaGetDeviceCount(count));
    return ISING_OK;
}

int ising_device_info(int device, char* name, size_t name_len, int* sms, int* max_threads_per_sm, int* cc_major,
                      int* cc_minor, int* ecc)
{
    cudaDeviceProp p;
    CU_TRY(cudaGetDeviceProperties(&p, device));
    if (name && name_len) { strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
    if (sms) *sms = p.multiProcessorCount;
    if (max_threads_per_sm) *max_threads_per_sm = p.maxThreadsPerMultiProcessor;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (ecc) *ecc = p.ECCEnabled;
    return ISING_OK;
}

int ising_can_access_peer(int device, int peer, int* ok)
{
    if (!ok) return fail(ISING_EINVAL, "null argument");
    CU_TRY(cudaDeviceCanAccessPeer(ok, device, peer));
    return ISING_OK;
}

void ising_site_random(uint64_t seed, int32_t X, int it, int color, uint32_t grow, uint32_t word, int q,
                       uint32_t out[4])
{
    const StepTables st = make_step_tables(seed, 2ull * (uint64_t)it + (uint64_t)color);
    const uint32_t tid = ref_tid(grow, word, (uint32_t)(X / 2048));
    const ThreadConsts tc = make_thread_consts(tid, st.x1, st.k1r3);
    const int cl = 8 * (int)((word % 32u) / 16u) + q;
    philox_fast(tc, st.a[cl], st.b[cl], st.d[cl], st.k0, st.k1, out);
}

int ising_create(const ising_cfg* cfg, ising_ctx** out)
{
    if (!cfg || !out) return fail(ISING_EINVAL, "null argument");
    *out = nullptr;
    if (int rc = validate(cfg)) return rc;
    int ndev_avail = 0;
    CU_TRY(cudaGetDeviceCount(&ndev_avail));
    if (cfg->first_device < 0 || cfg->first_device + cfg->ndev > ndev_avail)
        return fail(ISING_EINVAL, "devices %d..%d requested, %d visible", cfg->first_device,
                    cfg->first_device + cfg->ndev - 1, ndev_avail);

    ising_ctx* c = new ising_ctx();
    c->cfg = *cfg;
    c->pitch = cfg->X / 64;
    c->slab_count = cfg->slab_count ? cfg->slab_count : cfg->ndev;
    const bool sub = cfg->xsl || cfg->ysl;
    c->xperiod = (cfg->xsl ? cfg->xsl : cfg->X) / 64;
    if (sub) {                       // sub-lattices never cross a slab (ysl divides Y)
        c->yperiod = cfg->ysl ? cfg->ysl : cfg->Y;
    } else if (c->slab_count == 1) { // one slab: periodic inside it, no ghost rows needed
        c->yperiod = cfg->Y;
    } else {
        c->yperiod = 0; c->ghost_rows = true;
        c->peer_halo = (c->slab_count == cfg->ndev) && !(cfg->flags & ISING_F_NO_PEER_COPY);
    }
    c->rule = make_rule(cfg->temp);
    c->thr_half = (uint32_t)(accept_count(0.5f, true) - 1);
    c->dev.resize(cfg->ndev);
    if (const char* k = getenv("ISING_B200_KERNEL")) {
        c->persistent = strcmp(k, "persistent") == 0;
        c->two_tiles = strcmp(k, "tile1") != 0;
    }
    if (const char* k = getenv("ISING_B200_SPLIT_EDGES")) c->split_edges = atoi(k) != 0;
    if (const char* k = getenv("ISING_B200_EDGE_STREAM")) c->edge_stream = atoi(k) != 0;
    if (const char* t = getenv("ISING_B200_SPIN_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) c->spin_timeout_ns = (unsigned long long)ms * 1000000ull;
    }

    auto cleanup = [&](int rc) { ising_destroy(c); return rc; };
    for (int i = 0; i < cfg->ndev; i++) c->dev[i].device = cfg->first_device + i;

    c->bonds = (cfg->flags & ISING_F_BONDS) != 0;
    if (cfg->ndev > 1) {   // all pairs, as main.cu:1513-1536 (halo ring + correlation reads)
        for (int i = 0; i < cfg->ndev; i++) {
            if (cudaSetDevice(c->dev[i].device) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "cudaSetDevice failed"));
            for (int j = 0; j < cfg->ndev; j++) {
                if (j == i) continue;
                int ok = 0;
                cudaDeviceCanAccessPeer(&ok, c->dev[i].device, c->dev[j].device);
                if (!ok) return cleanup(fail(ISING_ENOPEER, "no direct link between devices %d and %d", c->dev[i].device, c->dev[j].device));
                const cudaError_t e = cudaDeviceEnablePeerAccess(c->dev[j].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return cleanup(fail(ISING_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)));
                cudaGetLastError();
            }
        }
    }

    const size_t plane_bytes = (size_t)(cfg->Y + 2) * c->pitch * 4;
    for (int i = 0; i < cfg->ndev; i++) {
        DevState& d = c->dev[i];
        cudaError_t e;
        if ((e = cudaSetDevice(d.device)) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "cudaSetDevice(%d): %s", d.device, cudaGetErrorString(e)));
        cudaDeviceProp prop;
        if ((e = cudaGetDeviceProperties(&prop, d.device)) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "%s", cudaGetErrorString(e)));
        d.sms = prop.multiProcessorCount;
        if (prop.major < 10) return cleanup(fail(ISING_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", d.device, prop.major, prop.minor));
        if ((e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking)) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "stream: %s", cudaGetErrorString(e)));
        for (int col = 0; col < 2; col++) {
            if ((e = cudaMalloc(&d.plane[col], plane_bytes)) != cudaSuccess) return cleanup(fail(ISING_ENOMEM, "cudaMalloc(%zu): %s", plane_bytes, cudaGetErrorString(e)));
            if ((e = cudaMemsetAsync(d.plane[col], 0, plane_bytes, d.stream)) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "%s", cudaGetErrorString(e)));
            int rc;
            if ((rc = make_map(&d.map_tile[col], d.plane[col], c->pitch, cfg->Y + 2, TW, TH))) return cleanup(rc);
            if ((rc = make_map(&d.map_row[col], d.plane[col], c->pitch, cfg->Y + 2, TW, 1))) return cleanup(rc);
            if ((rc = make_map(&d.map_side[col], d.plane[col], c->pitch, cfg->Y + 2, SIDE_W, TH))) return cleanup(rc);
            if ((rc = make_map(&d.map_tile2[col], d.plane[col], c->pitch, cfg->Y + 2, TW, 2 * TH))) return cleanup(rc);
            if ((rc = make_map(&d.map_side2[col], d.plane[col], c->pitch, cfg->Y + 2, SIDE_W, 2 * TH))) return cleanup(rc);
        }
        if ((e = cudaMalloc(&d.d_counts, 2 * sizeof(unsigned long long))) != cudaSuccess) return cleanup(fail(ISING_ENOMEM, "%s", cudaGetErrorString(e)));
        if (c->ghost_rows) {
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            if ((e = cudaStreamCreateWithPriority(&d.stream_edge, cudaStreamNonBlocking, prio_hi)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&d.ev_pre, cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&d.ev_edge, cudaEventDisableTiming)) != cudaSuccess)
                return cleanup(fail(ISING_ECUDA, "edge stream: %s", cudaGetErrorString(e)));
            if ((e = cudaMalloc(&d.d_flags, FLAG_WORDS * sizeof(unsigned int))) != cudaSuccess ||
                (e = cudaMemsetAsync(d.d_flags, 0, FLAG_WORDS * sizeof(unsigned int), d.stream)) != cudaSuccess)
                return cleanup(fail(ISING_ENOMEM, "flags: %s", cudaGetErrorString(e)));
        }
        if (c->bonds) {
            if ((e = cudaMalloc(&d.bond_random, 4 * plane_bytes)) != cudaSuccess ||
                (e = cudaMalloc(&d.bond_mirror, 4 * (size_t)cfg->Y * c->pitch * 4)) != cudaSuccess)
                return cleanup(fail(ISING_ENOMEM, "bond planes: %s", cudaGetErrorString(e)));
        }
        if ((e = cudaEventCreateWithFlags(&d.ev_color, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreate(&d.ev_t0)) != cudaSuccess || (e = cudaEventCreate(&d.ev_t1)) != cudaSuccess)
            return cleanup(fail(ISING_ECUDA, "event: %s", cudaGetErrorString(e)));
        if ((e = cudaStreamSynchronize(d.stream)) != cudaSuccess) return cleanup(fail(ISING_ECUDA, "%s", cudaGetErrorString(e)));
    }
    if (c->peer_halo) {   // single process: neighbours are reachable through peer access
        const char* mode = getenv("ISING_B200_HALO");
        const int n = cfg->ndev;
        for (int i = 0; i < n; i++) {
            DevState& d = c->dev[i];
            DevState& u = c->dev[(i + n - 1) % n];
            DevState& w = c->dev[(i + 1) % n];
            for (int col = 0; col < 2; col++) { d.up_plane[col] = u.plane[col]; d.dn_plane[col] = w.plane[col]; }
            d.up_flags = u.d_flags; d.dn_flags = w.d_flags;
        }
        c->fused = !(mode && strcmp(mode, "copy") == 0);   // "copy": cudaMemcpyPeerAsync + events per colour
    }
    *out = c;
    return ISING_OK;
}

int ising_destroy(ising_ctx* c)
{
    if (!c) return ISING_OK;
    if (!c->dev.empty() && (c->batch_h_counts || c->batch_d_counts)) {
        cudaSetDevice(c->dev[0].device);
        if (c->batch_h_counts) cudaFreeHost(c->batch_h_counts);
        if (c->batch_d_counts) cudaFree(c->batch_d_counts);
    }
    for (DevState& d : c->dev) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        for (int col = 0; col < 2; col++) if (d.plane[col]) cudaFree(d.plane[col]);
        if (d.d_counts) cudaFree(d.d_counts);
        if (d.bond_random) cudaFree(d.bond_random);
        if (d.bond_mirror) cudaFree(d.bond_mirror);
        if (d.d_corr) cudaFree(d.d_corr);
        if (d.d_flags) cudaFree(d.d_flags);
        for (void* q : d.ipc_opened) if (q) cudaIpcCloseMemHandle(q);
        if (d.stream_edge) { cudaStreamSynchronize(d.stream_edge); cudaStreamDestroy(d.stream_edge); }
        if (d.ev_pre) cudaEventDestroy(d.ev_pre);
        if (d.ev_edge) cudaEventDestroy(d.ev_edge);
        if (d.ev_color) cudaEventDestroy(d.ev_color);
        if (d.ev_t0) cudaEventDestroy(d.ev_t0);
        if (d.ev_t1) cudaEventDestroy(d.ev_t1);
        if (d.stream && d.own_stream) cudaStreamDestroy(d.stream);
    }
    delete c;
    return ISING_OK;
}

int ising_set_stream(ising_ctx* c, int dev, void* stream)
{
    if (!c || dev < 0 || dev >= (int)c->dev.size()) return fail(ISING_EINVAL, "bad device index");
    DevState& d = c->dev[dev];
    CU_TRY(cudaSetDevice(d.device));
    CU_TRY(cudaStreamSynchronize(d.stream));
    if (d.own_stream) CU_TRY(cudaStreamDestroy(d.stream));
    d.stream = (cudaStream_t)stream;
    d.own_stream = false;
    return ISING_OK;
}

int ising_set_temperature(ising_ctx* c, float temp)
{
    if (!c) return fail(ISING_EINVAL, "null context");
    c->cfg.temp = temp;
    c->rule = make_rule(temp);
    return ISING_OK;
}

int ising_get_temperature(const ising_ctx* c, float* temp)
{
    if (!c || !temp) return fail(ISING_EINVAL, "bad argument");
    *temp = c->cfg.temp;
    return ISING_OK;
}

int ising_init_lattice(ising_ctx* c)
{
    if (!c) return fail(ISING_EINVAL, "null context");
    for (int i = 0; i < (int)c->dev.size(); i++) {
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        for (int col = 0; col < 2; col++) {
            InitParams p{};
            p.st = make_step_tables(c->cfg.seed, (uint64_t)col);   // it = 0
            p.thr_half = c->thr_half;
            p.grid_x = (uint32_t)(c->cfg.X / 2048);
            p.grow0 = (uint32_t)(c->cfg.slab_first + i) * (uint32_t)c->cfg.Y;
            p.rows = c->cfg.Y;
            p.pitch = c->pitch;
            launch_init(local_row(c, d, col, 0), p, d.stream);
            CU_TRY(cudaGetLastError());
            c->launches++;
        }
    }
    if (c->bonds) {   // main.cu:1728-1742
        for (int i = 0; i < (int)c->dev.size(); i++) {
            DevState& d = c->dev[i];
            CU_TRY(cudaSetDevice(d.device));
            BondParams b{};
            const uint64_t s1 = c->cfg.seed + 1ull;
            b.key0 = (uint32_t)s1; b.key1 = (uint32_t)(s1 >> 32);
            const uint64_t n = accept_count(c->cfg.j_prob, true);   // uniform < prob
            b.always = (n == (1ull << 32)); b.never = (n == 0); b.thr = n ? (uint32_t)(n - 1) : 0u;
            b.grid_x = (uint32_t)(c->cfg.X / 2048);
            b.grow0 = (uint32_t)(c->cfg.slab_first + i) * (uint32_t)c->cfg.Y;
            b.total_rows = (long long)c->slab_count * c->cfg.Y;
            b.rows = c->cfg.Y; b.pitch = c->pitch; b.yperiod = c->yperiod; b.xperiod = c->xperiod;
            launch_bond_init(d.bond_random, d.bond_mirror, b, d.stream);
            CU_TRY(cudaGetLastError());
            c->launches += 2;
        }
    }
    if (c->peer_halo) {
        if (int rc = push_halo(c, 0)) return rc;
        if (int rc = push_halo(c, 1)) return rc;
    }
    return ISING_OK;
}

int ising_correlation(ising_ctx* c, int ncorr, double* out)
{
    if (!c || !out || ncorr < 1 || ncorr > 4096) return fail(ISING_EINVAL, "bad argument");
    const int n = (int)c->dev.size();
    if (c->slab_count != n) return fail(ISING_ESTATE, "ising_correlation needs the whole lattice in one context");
    static_assert(MAX_CORR_SLABS == ISING_MAX_CORR_DEVICES, "header limit out of sync");
    if (n > MAX_CORR_SLABS) return fail(ISING_ESTATE, "ising_correlation supports up to %d devices", MAX_CORR_SLABS);
    if (2 * (size_t)c->pitch * 4 > ISING_MAX_CORR_ROW_BYTES) return fail(ISING_ESTATE, "X too large for the correlation kernel");
    if (int rc = ising_sync(c)) return rc;
    CorrParams p{};
    for (int i = 0; i < n; i++) { p.black[i] = local_row(c, c->dev[i], 0, 0); p.white[i] = local_row(c, c->dev[i], 1, 0); }
    p.rows = c->cfg.Y; p.pitch = c->pitch; p.ncorr = ncorr;
    p.yperiod = (c->cfg.xsl || c->cfg.ysl) ? (long long)(c->cfg.ysl ? c->cfg.ysl : c->cfg.Y) : (long long)n * c->cfg.Y;
    p.xperiod = c->xperiod;
    std::vector<unsigned long long> tot(ncorr, 0ull), h(ncorr);
    for (int i = 0; i < n; i++) {
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        if (!d.d_corr) CU_TRY(cudaMalloc(&d.d_corr, 4096 * sizeof(unsigned long long)));
        CU_TRY(cudaMemsetAsync(d.d_corr, 0, ncorr * sizeof(unsigned long long), d.stream));
        p.self = i; p.grow0 = (long long)i * c->cfg.Y;
        launch_corr(p, d.d_corr, d.stream);
        CU_TRY(cudaGetLastError());
        c->launches++;
    }
    for (int i = 0; i < n; i++) {
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemcpyAsync(h.data(), d.d_corr, ncorr * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream));
        CU_TRY(cudaStreamSynchronize(d.stream));
        for (int j = 0; j < ncorr; j++) tot[j] += h[j];
    }
    // equal pairs E of 2N pairs (N sites, right + below): sum of +-1 = 2E - 2N, normalised by 2N
    const double N = (double)n * (double)c->cfg.Y * (double)c->cfg.X;
    for (int j = 0; j < ncorr; j++) out[j] = (2.0 * (double)tot[j] - 2.0 * N) / (2.0 * N);
    return ISING_OK;
}

int ising_export_bonds_ref_layout(ising_ctx* c, int dev, uint64_t* random, uint64_t* mirror)
{
    if (!c || dev < 0 || dev >= (int)c->dev.size()) return fail(ISING_EINVAL, "bad device index");
    if (!c->bonds) return fail(ISING_ESTATE, "context created without ISING_F_BONDS");
    DevState& d = c->dev[dev];
    const size_t words = (size_t)c->cfg.Y * c->pitch;
    std::vector<uint32_t> pl(4 * words);
    CU_TRY(cudaSetDevice(d.device));
    CU_TRY(cudaStreamSynchronize(d.stream));
    for (int which = 0; which < 2; which++) {
        uint64_t* out = which ? mirror : random;
        if (!out) continue;
        for (int k = 0; k < 4; k++) {
            const uint32_t* src = which ? d.bond_mirror + k * words
                                        : d.bond_random + k * (size_t)(c->cfg.Y + 2) * c->pitch + c->pitch;
            CU_TRY(cudaMemcpy(pl.data() + k * words, src, words * 4, cudaMemcpyDeviceToHost));
        }
        // plane k: 0 up (bit 3), 1 down (bit 2), 2 left (bit 1), 3 right (bit 0)
        for (size_t w = 0; w < words; w++)
            for (int hlf = 0; hlf < 2; hlf++) {
                uint64_t v = 0;
                for (int z = 0; z < 16; z++) {
                    const int b = 16 * hlf + z;
                    const uint64_t nib = (((pl[0 * words + w] >> b) & 1u) << 3) | (((pl[1 * words + w] >> b) & 1u) << 2) |
                                         (((pl[2 * words + w] >> b) & 1u) << 1) | ((pl[3 * words + w] >> b) & 1u);
                    v |= nib << (4 * z);
                }
                out[2 * w + hlf] = v;
            }
    }
    return ISING_OK;
}

int ising_update_color(ising_ctx* c, int color, int it)
{
    if (!c || (color != 0 && color != 1)) return fail(ISING_EINVAL, "bad argument");
    if (c->fused) {
        // Tile kernel: the two edge tile rows (64 CTAs at X = 65536) run in the kernel variant that stores them into
        // the neighbours' ghost rows; the interior runs in the plain variant (one kernel carrying the push code for
        // every CTA measured 1.2 % slower per launch).  Both read only what the previous colour step wrote, so they
        // are independent of each other: the edge launch goes to a second, high-priority stream and the interior
        // grid fills the SMs next to it -- launched back to back on one stream the GPU idled at < 11 % occupancy
        // for one CTA lifetime (~10 us) twice per sweep (round-1 verdict: the whole 1.6 % lost at 8 GPUs).
        // The persistent kernel walks the edge tile rows first inside ONE launch.
        const bool split = !c->persistent && c->split_edges && c->cfg.Y / TH > 2;
        for (int i = 0; i < (int)c->dev.size(); i++) {
            DevState& d = c->dev[i];
            if (!split) {
                if (int rc = launch_part(c, i, color, it, ISING_PART_ALL, d.stream)) return rc;
            } else if (!c->edge_stream || !d.stream_edge) {
                if (int rc = launch_part(c, i, color, it, ISING_PART_EDGES, d.stream)) return rc;
                if (int rc = launch_part(c, i, color, it, ISING_PART_INTERIOR, d.stream)) return rc;
            } else {
                CU_TRY(cudaSetDevice(d.device));
                CU_TRY(cudaEventRecord(d.ev_pre, d.stream));                 // everything issued so far, incl. the previous
                CU_TRY(cudaStreamWaitEvent(d.stream_edge, d.ev_pre, 0));     // colour step's interior
                if (int rc = launch_part(c, i, color, it, ISING_PART_EDGES, d.stream_edge)) return rc;
                CU_TRY(cudaEventRecord(d.ev_edge, d.stream_edge));
                if (int rc = launch_part(c, i, color, it, ISING_PART_INTERIOR, d.stream)) return rc;
                CU_TRY(cudaStreamWaitEvent(d.stream, d.ev_edge, 0));         // whatever follows on the main stream sees both
            }
        }
        c->pushes_done[color]++;
        return ISING_OK;
    }
    for (int i = 0; i < (int)c->dev.size(); i++)
        if (int rc = launch_part(c, i, color, it, ISING_PART_ALL, c->dev[i].stream)) return rc;
    if (c->peer_halo) return push_halo(c, color);
    return ISING_OK;
}

int ising_ipc_export(ising_ctx* c, void* handles, size_t nbytes)
{
    if (!c || c->dev.size() != 1 || !handles || nbytes < 3 * sizeof(cudaIpcMemHandle_t))
        return fail(ISING_EINVAL, "ising_ipc_export: 1-device context and a buffer of %zu bytes needed", 3 * sizeof(cudaIpcMemHandle_t));
    if (!c->ghost_rows) return fail(ISING_ESTATE, "context is not a slab of a larger lattice");
    DevState& d = c->dev[0];
    CU_TRY(cudaSetDevice(d.device));
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles;
    CU_TRY(cudaIpcGetMemHandle(&h[0], d.plane[0]));
    CU_TRY(cudaIpcGetMemHandle(&h[1], d.plane[1]));
    CU_TRY(cudaIpcGetMemHandle(&h[2], d.d_flags));
    return ISING_OK;
}

int ising_ipc_attach(ising_ctx* c, const void* up_handles, const void* dn_handles)
{
    if (!c || c->dev.size() != 1 || !up_handles || !dn_handles) return fail(ISING_EINVAL, "bad argument");
    if (!c->ghost_rows) return fail(ISING_ESTATE, "context is not a slab of a larger lattice");
    DevState& d = c->dev[0];
    CU_TRY(cudaSetDevice(d.device));
    const cudaIpcMemHandle_t* hu = (const cudaIpcMemHandle_t*)up_handles;
    const cudaIpcMemHandle_t* hd = (const cudaIpcMemHandle_t*)dn_handles;
    void* pu[3];
    void* pd[3];
    for (int k = 0; k < 3; k++) {
        CU_TRY(cudaIpcOpenMemHandle(&pu[k], hu[k], cudaIpcMemLazyEnablePeerAccess));
        d.ipc_opened[k] = pu[k];
    }
    const bool same = memcmp(hu, hd, 3 * sizeof(cudaIpcMemHandle_t)) == 0;   // 2 slabs: one neighbour is both
    for (int k = 0; k < 3; k++) {
        if (same) pd[k] = pu[k];
        else { CU_TRY(cudaIpcOpenMemHandle(&pd[k], hd[k], cudaIpcMemLazyEnablePeerAccess)); d.ipc_opened[3 + k] = pd[k]; }
    }
    d.up_plane[0] = (uint32_t*)pu[0]; d.up_plane[1] = (uint32_t*)pu[1]; d.up_flags = (unsigned int*)pu[2];
    d.dn_plane[0] = (uint32_t*)pd[0]; d.dn_plane[1] = (uint32_t*)pd[1]; d.dn_flags = (unsigned int*)pd[2];
    c->fused = true;
    return ISING_OK;
}

// Store this slab's first/last row of `color` into the neighbours' ghost rows with peer stores
// (initial fill after init/import; the sweep loop pushes from inside the update kernel).
int ising_push_edges(ising_ctx* c, int color)
{
    if (!c || (color != 0 && color != 1)) return fail(ISING_EINVAL, "bad argument");
    if (!c->fused) return fail(ISING_ESTATE, "fused halo push is not enabled on this context");
    const size_t row_bytes = (size_t)c->pitch * 4;
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemcpyAsync(d.up_plane[color] + (size_t)(c->cfg.Y + 1) * c->pitch, local_row(c, d, color, 0), row_bytes,
                               cudaMemcpyDefault, d.stream));
        CU_TRY(cudaMemcpyAsync(d.dn_plane[color], local_row(c, d, color, c->cfg.Y - 1), row_bytes, cudaMemcpyDefault,
                               d.stream));
        CU_TRY(cudaStreamSynchronize(d.stream));
    }
    return ISING_OK;
}

int ising_update_color_part(ising_ctx* c, int dev, int color, int it, int part, void* stream)
{
    if (!c || dev < 0 || dev >= (int)c->dev.size() || (color != 0 && color != 1) || part < 0 || part > 2)
        return fail(ISING_EINVAL, "bad argument");
    // a partial launch on a fused context would push without being counted in pushes_done: every later
    // wait value would be too small and the ghost-row ordering silently lost
    if (c->fused) return fail(ISING_ESTATE, "ising_update_color_part is not valid on a context with the fused halo push; use ising_update_color");
    return launch_part(c, dev, color, it, part, stream ? (cudaStream_t)stream : c->dev[dev].stream);
}

int ising_sweeps(ising_ctx* c, int first_it, int n)
{
    if (!c || n < 0) return fail(ISING_EINVAL, "bad argument");
    for (int k = 0; k < n; k++) {
        if (int rc = ising_update_color(c, ISING_BLACK, first_it + k)) return rc;
        if (int rc = ising_update_color(c, ISING_WHITE, first_it + k)) return rc;
    }
    return ISING_OK;
}

int ising_sync(ising_ctx* c)
{
    if (!c) return fail(ISING_EINVAL, "null context");
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaStreamSynchronize(d.stream));
    }
    if (c->fused)   // a bounded wait on a neighbour's arrival counter gave up: the lattice is not valid any more
        for (DevState& d : c->dev) {
            unsigned int err = 0;
            CU_TRY(cudaSetDevice(d.device));
            CU_TRY(cudaMemcpy(&err, d.d_flags + ERR_FLAG_WORD, sizeof(err), cudaMemcpyDeviceToHost));
            if (err) return fail(ISING_ETIMEOUT, "device %d: halo rows of a neighbouring slab did not arrive within %.1f s",
                                 d.device, (double)c->spin_timeout_ns * 1e-9);
        }
    return ISING_OK;
}

int ising_sweeps_timed(ising_ctx* c, int first_it, int n, float* ms)
{
    if (!c || !ms) return fail(ISING_EINVAL, "bad argument");
    if (int rc = ising_sync(c)) return rc;
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaEventRecord(d.ev_t0, d.stream));
    }
    if (int rc = ising_sweeps(c, first_it, n)) return rc;
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaEventRecord(d.ev_t1, d.stream));
    }
    float worst = 0.f;
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaEventSynchronize(d.ev_t1));
        float t = 0.f;
        CU_TRY(cudaEventElapsedTime(&t, d.ev_t0, d.ev_t1));
        if (t > worst) worst = t;
    }
    *ms = worst;
    return ISING_OK;
}

int ising_count_spins(ising_ctx* c, uint64_t* up, uint64_t* down)
{
    if (!c || !up || !down) return fail(ISING_EINVAL, "bad argument");
    uint64_t tot = 0;
    for (DevState& d : c->dev) {
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemsetAsync(d.d_counts, 0, sizeof(unsigned long long), d.stream));
        for (int col = 0; col < 2; col++) {
            launch_count(local_row(c, d, col, 0), c->cfg.Y, c->pitch, d.d_counts, d.stream);
            CU_TRY(cudaGetLastError());
            c->launches++;
        }
    }
    for (DevState& d : c->dev) {
        unsigned long long h = 0;
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemcpyAsync(&h, d.d_counts, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
        CU_TRY(cudaStreamSynchronize(d.stream));
        tot += h;
    }
    *up = tot;
    *down = (uint64_t)c->dev.size() * (uint64_t)c->cfg.Y * (uint64_t)c->cfg.X - tot;
    return ISING_OK;
}

int ising_bond_sum(ising_ctx* c, int64_t* bond_sum)
{
    if (!c || !bond_sum) return fail(ISING_EINVAL, "bad argument");
    // Ghost rows: bond_differ_kernel reads the WHITE ghost rows, which the neighbouring slabs write.
    //   fused push  : their stores are ordered by the arrival counters only -> wait for the counters on
    //                 this device's stream first (covers peer-access and IPC neighbours alike);
    //   copy mode   : push_halo() already made this stream wait for the neighbours' copies;
    //   caller-managed (ISING_F_NO_PEER_COPY without attach): current after the caller's own exchange.
    uint64_t differ = 0;
    for (int i = 0; i < (int)c->dev.size(); i++) {
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        if (c->ghost_rows && c->fused) {
            const unsigned int tiles_x = (unsigned int)(c->pitch / TW);
            launch_halo_wait(d.d_flags, (unsigned int)(c->pushes_done[0] * tiles_x), (unsigned int)(c->pushes_done[1] * tiles_x),
                             d.d_flags + ERR_FLAG_WORD, c->spin_timeout_ns, d.stream);
            CU_TRY(cudaGetLastError());
            c->launches++;
        }
        CU_TRY(cudaMemsetAsync(d.d_counts + 1, 0, sizeof(unsigned long long), d.stream));
        launch_bond_differ(local_row(c, d, 0, 0), local_row(c, d, 1, 0), c->cfg.Y, c->pitch, c->yperiod, c->xperiod,
                           (uint32_t)(c->cfg.slab_first + i) * (uint32_t)c->cfg.Y, d.d_counts + 1, d.stream);
        CU_TRY(cudaGetLastError());
        c->launches++;
    }
    for (DevState& d : c->dev) {
        unsigned long long h = 0;
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemcpyAsync(&h, d.d_counts + 1, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
        CU_TRY(cudaStreamSynchronize(d.stream));
        differ += h;
    }
    const uint64_t bonds = 2ull * (uint64_t)c->dev.size() * (uint64_t)c->cfg.Y * (uint64_t)c->cfg.X;
    *bond_sum = (int64_t)bonds - 2 * (int64_t)differ;
    return ISING_OK;
}

static int check_dev(ising_ctx* c, int dev)
{
    if (!c || dev < 0 || dev >= (int)c->dev.size()) return fail(ISING_EINVAL, "bad device index %d", dev);
    return ISING_OK;
}

int ising_export_bits(ising_ctx* c, int dev, uint32_t* black, uint32_t* white)
{
    if (int rc = check_dev(c, dev)) return rc;
    DevState& d = c->dev[dev];
    const size_t bytes = (size_t)c->cfg.Y * c->pitch * 4;
    CU_TRY(cudaSetDevice(d.device));
    if (black) CU_TRY(cudaMemcpyAsync(black, local_row(c, d, 0, 0), bytes, cudaMemcpyDeviceToHost, d.stream));
    if (white) CU_TRY(cudaMemcpyAsync(white, local_row(c, d, 1, 0), bytes, cudaMemcpyDeviceToHost, d.stream));
    CU_TRY(cudaStreamSynchronize(d.stream));
    return ISING_OK;
}

static int refresh_ghosts_after_import(ising_ctx* c)
{
    if (!c->peer_halo) return ISING_OK;
    if (int rc = push_halo(c, 0)) return rc;
    return push_halo(c, 1);
}

int ising_import_bits(ising_ctx* c, int dev, const uint32_t* black, const uint32_t* white)
{
    if (int rc = check_dev(c, dev)) return rc;
    DevState& d = c->dev[dev];
    const size_t bytes = (size_t)c->cfg.Y * c->pitch * 4;
    CU_TRY(cudaSetDevice(d.device));
    if (black) CU_TRY(cudaMemcpyAsync(local_row(c, d, 0, 0), black, bytes, cudaMemcpyHostToDevice, d.stream));
    if (white) CU_TRY(cudaMemcpyAsync(local_row(c, d, 1, 0), white, bytes, cudaMemcpyHostToDevice, d.stream));
    CU_TRY(cudaStreamSynchronize(d.stream));
    return refresh_ghosts_after_import(c);
}

// ---- binary checkpoint (the reference has none) ---------------------------------
namespace {
struct CkptHeader {
    char magic[8];          // "ISB200CK"
    uint32_t version;       // 2 (1 had no flags / j_prob)
    int32_t X, Y, ndev, xsl, ysl, slab_first, slab_count;
    uint64_t seed;
    float temp;             // temperature in force when the checkpoint was written (restored on load)
    float j_prob;           // -J probability (the bond planes are regenerated from seed + 1 and this)
    uint64_t iterations;    // sweeps done; the next sweep uses it = iterations + 1
    int32_t flags;          // ISING_F_BONDS of the writing context
    int32_t pad;
};
constexpr uint32_t CKPT_VERSION = 2;
}  // namespace

int ising_save_checkpoint(ising_ctx* c, const char* path, uint64_t iterations_done)
{
    if (!c || !path) return fail(ISING_EINVAL, "bad argument");
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(ISING_EINVAL, "cannot open %s for writing", path);
    CkptHeader h{};
    memcpy(h.magic, "ISB200CK", 8);
    h.version = CKPT_VERSION;
    h.flags = c->cfg.flags & ISING_F_BONDS; h.j_prob = c->bonds ? c->cfg.j_prob : 0.f;
    h.X = c->cfg.X; h.Y = c->cfg.Y; h.ndev = (int32_t)c->dev.size(); h.xsl = c->cfg.xsl; h.ysl = c->cfg.ysl;
    h.slab_first = c->cfg.slab_first; h.slab_count = c->slab_count; h.seed = c->cfg.seed; h.temp = c->cfg.temp;
    h.iterations = iterations_done;
    const size_t words = (size_t)c->cfg.Y * c->pitch;
    std::vector<uint32_t> b(words), w(words);
    bool ok = fwrite(&h, sizeof(h), 1, fp) == 1;
    for (int i = 0; ok && i < (int)c->dev.size(); i++) {
        if (int rc = ising_export_bits(c, i, b.data(), w.data())) { fclose(fp); return rc; }
        ok = fwrite(b.data(), 4, words, fp) == words && fwrite(w.data(), 4, words, fp) == words;
    }
    ok = (fclose(fp) == 0) && ok;
    return ok ? ISING_OK : fail(ISING_EINVAL, "short write to %s", path);
}

int ising_load_checkpoint(ising_ctx* c, const char* path, uint64_t* iterations_done)
{
    if (!c || !path) return fail(ISING_EINVAL, "bad argument");
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(ISING_EINVAL, "cannot open %s", path);
    CkptHeader h{};
    if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, "ISB200CK", 8) != 0) {
        fclose(fp);
        return fail(ISING_EINVAL, "%s is not an ising_b200 checkpoint", path);
    }
    if (h.version != CKPT_VERSION) {
        fclose(fp);
        return fail(ISING_EINVAL, "%s is a version %u checkpoint, this library reads version %u", path, h.version, CKPT_VERSION);
    }
    // the Hamiltonian must be the one the run was started with: same -J flag and probability
    if (h.flags != (c->cfg.flags & ISING_F_BONDS) || (c->bonds && h.j_prob != c->cfg.j_prob)) {
        fclose(fp);
        return fail(ISING_EINVAL, "checkpoint %s was written %s random bonds (p = %g), this context runs %s (p = %g)", path,
                    h.flags ? "with" : "without", (double)h.j_prob, c->bonds ? "with" : "without",
                    (double)(c->bonds ? c->cfg.j_prob : 0.f));
    }
    if (h.X != c->cfg.X || h.Y != c->cfg.Y || h.ndev != (int32_t)c->dev.size() || h.slab_first != c->cfg.slab_first ||
        h.slab_count != c->slab_count || h.xsl != c->cfg.xsl || h.ysl != c->cfg.ysl) {
        fclose(fp);
        return fail(ISING_EINVAL, "checkpoint %s is %d x %d x %d devices (slabs %d+%d, sub-lattice %d x %d): does not match this context",
                    path, h.Y, h.X, h.ndev, h.slab_first, h.slab_count, h.ysl, h.xsl);
    }
    if (h.seed != c->cfg.seed) {   // the random-number stream of the continuation depends on it
        fclose(fp);
        return fail(ISING_EINVAL, "checkpoint %s was written with seed %llu, this context uses %llu", path,
                    (unsigned long long)h.seed, (unsigned long long)c->cfg.seed);
    }
    const size_t words = (size_t)c->cfg.Y * c->pitch;
    std::vector<uint32_t> b(words), w(words);
    for (int i = 0; i < (int)c->dev.size(); i++) {
        if (fread(b.data(), 4, words, fp) != words || fread(w.data(), 4, words, fp) != words) {
            fclose(fp);
            return fail(ISING_EINVAL, "checkpoint %s is truncated", path);
        }
        DevState& d = c->dev[i];
        CU_TRY(cudaSetDevice(d.device));
        CU_TRY(cudaMemcpyAsync(local_row(c, d, 0, 0), b.data(), words * 4, cudaMemcpyHostToDevice, d.stream));
        CU_TRY(cudaMemcpyAsync(local_row(c, d, 1, 0), w.data(), words * 4, cudaMemcpyHostToDevice, d.stream));
        CU_TRY(cudaStreamSynchronize(d.stream));
    }
    fclose(fp);
    if (iterations_done) *iterations_done = h.iterations;
    // the run continues under the temperature it was saved at (a -u ramp may have moved it since the start)
    if (h.temp != c->cfg.temp)
        if (int rc = ising_set_temperature(c, h.temp)) return rc;
    return refresh_ghosts_after_import(c);   // all devices are loaded: boundary rows -> neighbours' ghost rows
}

int ising_export_ref_layout(ising_ctx* c, int dev, uint64_t* black, uint64_t* white)
{
    if (int rc = check_dev(c, dev)) return rc;
    DevState& d = c->dev[dev];
    const size_t words = (size_t)c->cfg.Y * c->pitch * 2;   // uint64 words per colour
    unsigned long long* tmp = nullptr;
    CU_TRY(cudaSetDevice(d.device));
    CU_TRY(cudaMalloc(&tmp, words * 8));
    uint64_t* dst[2] = {black, white};
    for (int col = 0; col < 2; col++) {
        if (!dst[col]) continue;
        launch_to_ref(local_row(c, d, col, 0), c->cfg.Y, c->pitch, tmp, d.stream);
        c->launches++;
        cudaError_t e = cudaMemcpyAsync(dst[col], tmp, words * 8, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(ISING_ECUDA, "export: %s", cudaGetErrorString(e)); }
    }
    CU_TRY(cudaFree(tmp));
    return ISING_OK;
}

int ising_import_ref_layout(ising_ctx* c, int dev, const uint64_t* black, const uint64_t* white)
{
    if (int rc = check_dev(c, dev)) return rc;
    DevState& d = c->dev[dev];
    const size_t words = (size_t)c->cfg.Y * c->pitch * 2;
    unsigned long long* tmp = nullptr;
    CU_TRY(cudaSetDevice(d.device));
    CU_TRY(cudaMalloc(&tmp, words * 8));
    const uint64_t* src[2] = {black, white};
    for (int col = 0; col < 2; col++) {
        if (!src[col]) continue;
        cudaError_t e = cudaMemcpyAsync(tmp, src[col], words * 8, cudaMemcpyHostToDevice, d.stream);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(ISING_ECUDA, "import: %s", cudaGetErrorString(e)); }
        launch_from_ref(tmp, c->cfg.Y, c->pitch, local_row(c, d, col, 0), d.stream);
        c->launches++;
        e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(ISING_ECUDA, "import: %s", cudaGetErrorString(e)); }
    }
    CU_TRY(cudaFree(tmp));
    return refresh_ghosts_after_import(c);
}

int ising_dump_lattice(ising_ctx* c, const char* prefix)
{
    if (!c || !prefix) return fail(ISING_EINVAL, "bad argument");
    const size_t words = (size_t)c->cfg.Y * c->pitch;
    std::vector<uint32_t> b(words), w(words);
    std::vector<char> line((size_t)c->cfg.X + 1);
    for (int i = 0; i < (int)c->dev.size(); i++) {
        if (int rc = ising_export_bits(c, i, b.data(), w.data())) return rc;
        char fname[512];
        snprintf(fname, sizeof(fname), "%s%d.txt", prefix, i);   // main.cu:1157, :1185
        FILE* fp = fopen(fname, "w");
        if (!fp) return fail(ISING_EINVAL, "cannot open %s for writing", fname);
        for (int r = 0; r < c->cfg.Y; r++) {
            const uint32_t* rb = b.data() + (size_t)r * c->pitch;
            const uint32_t* rw = w.data() + (size_t)r * c->pitch;
            // the reference tests the LOCAL row index for the interleave (main.cu:1166, :1194);
            // Y is even, so local and global parity agree
            const bool odd = r & 1;
            for (int m = 0; m < c->cfg.X / 2; m++) {
                const char cb = (char)('0' + ((rb[m >> 5] >> (m & 31)) & 1u));
                const char cw = (char)('0' + ((rw[m >> 5] >> (m & 31)) & 1u));
                line[2 * (size_t)m] = odd ? cw : cb;
                line[2 * (size_t)m + 1] = odd ? cb : cw;
            }
            line[(size_t)c->cfg.X] = '\n';
            fwrite(line.data(), 1, line.size(), fp);
        }
        fclose(fp);
    }
    return ISING_OK;
}

int ising_run_from_host(ising_ctx* c, const uint32_t* black_in, const uint32_t* white_in, int first_it, int n,
                        uint32_t* black_out, uint32_t* white_out, uint64_t* up, uint64_t* down)
{
    if (!c || c->dev.size() != 1) return fail(ISING_ESTATE, "ising_run_from_host needs a 1-device context");
    DevState& d = c->dev[0];
    const size_t bytes = (size_t)c->cfg.Y * c->pitch * 4;
    CU_TRY(cudaSetDevice(d.device));
    if (black_in) CU_TRY(cudaMemcpyAsync(local_row(c, d, 0, 0), black_in, bytes, cudaMemcpyHostToDevice, d.stream));
    if (white_in) CU_TRY(cudaMemcpyAsync(local_row(c, d, 1, 0), white_in, bytes, cudaMemcpyHostToDevice, d.stream));
    if (int rc = ising_sweeps(c, first_it, n)) return rc;
    if (black_out) CU_TRY(cudaMemcpyAsync(black_out, local_row(c, d, 0, 0), bytes, cudaMemcpyDeviceToHost, d.stream));
    if (white_out) CU_TRY(cudaMemcpyAsync(white_out, local_row(c, d, 1, 0), bytes, cudaMemcpyDeviceToHost, d.stream));
    if (up && down) return ising_count_spins(c, up, down);
    CU_TRY(cudaStreamSynchronize(d.stream));
    return ISING_OK;
}

// Replica batch: job k runs on ctx[k & 1]'s own stream as H2D -> n sweeps -> count -> D2H, everything
// asynchronous.  The kernels of job k+1 wait (event) for the kernels of job k, so the jobs compute
// one after the other at full speed while the upload of job k+1 and the download of job k-1 ride
// on the copy engines -- without that the two streams run in lock-step and the SMs idle during
// the copies.  A context is reused by job k+2 in stream order.
static int run_batch(ising_ctx* const* ctxs, int nctx, int nrep, const uint32_t* const* black_in,
                     const uint32_t* const* white_in, const uint32_t* const* ghosts_in, int first_it, int n,
                     uint32_t* const* black_out, uint32_t* const* white_out, uint64_t* up, uint64_t* down)
{
    if (!ctxs || nctx < 1 || nctx > 8 || !ctxs[0] || nrep < 0 || n < 0 || !black_in || !white_in)
        return fail(ISING_EINVAL, "bad argument");
    std::vector<ising_ctx*> cs;
    for (int i = 0; i < nctx; i++) cs.push_back(ctxs[i] ? ctxs[i] : ctxs[0]);
    ising_ctx* const ctx_a = cs[0];
    for (ising_ctx* c : cs) {
        if (c->dev.size() != 1) return fail(ISING_ESTATE, "replica batches need 1-device contexts");
        if (!ghosts_in && c->ghost_rows) return fail(ISING_ESTATE, "ising_run_batch_from_host needs 1-slab contexts");
        if (ghosts_in && !(c->ghost_rows && c->fused && c->dev[0].up_flags))
            return fail(ISING_ESTATE, "ising_run_slab_batch_from_host needs slab contexts with attached peers (ising_ipc_attach)");
        if (c->cfg.X != ctx_a->cfg.X || c->cfg.Y != ctx_a->cfg.Y || c->dev[0].device != ctx_a->dev[0].device)
            return fail(ISING_EINVAL, "the two contexts must have the same shape and device");
    }
    const size_t bytes = (size_t)ctx_a->cfg.Y * ctx_a->pitch * 4;
    const uint64_t spins = (uint64_t)ctx_a->cfg.Y * (uint64_t)ctx_a->cfg.X;
    CU_TRY(cudaSetDevice(ctx_a->dev[0].device));
    unsigned long long* h_counts = nullptr;      // pinned landing zone of the per-job up-counts
    unsigned long long* d_counts = nullptr;
    if (up && down && nrep > 0) {
        if (ctx_a->batch_counts_cap < nrep) {        // (re)allocate the counters owned by the first context
            if (ctx_a->batch_h_counts) cudaFreeHost(ctx_a->batch_h_counts);
            if (ctx_a->batch_d_counts) cudaFree(ctx_a->batch_d_counts);
            ctx_a->batch_h_counts = nullptr; ctx_a->batch_d_counts = nullptr; ctx_a->batch_counts_cap = 0;
            const int cap = nrep < 64 ? 64 : nrep;
            CU_TRY(cudaHostAlloc((void**)&ctx_a->batch_h_counts, sizeof(unsigned long long) * cap, cudaHostAllocDefault));
            CU_TRY(cudaMalloc((void**)&ctx_a->batch_d_counts, sizeof(unsigned long long) * cap));
            ctx_a->batch_counts_cap = cap;
        }
        h_counts = ctx_a->batch_h_counts;
        d_counts = ctx_a->batch_d_counts;
        // zeroed on the stream of the first job; the count kernels of later jobs follow it through the event chain
        CU_TRY(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * nrep, cs[0]->dev[0].stream));
    }
    int rc = ISING_OK;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < nrep && rc == ISING_OK && e == cudaSuccess; k++) {
        ising_ctx* c = cs[k % nctx];
        DevState& d = c->dev[0];
        if (ghosts_in) {
            // the neighbours' last pushes of the previous replica on this context must have landed before
            // the ghost rows are overwritten from the host
            const unsigned int tiles_x = (unsigned int)(c->pitch / TW);
            launch_halo_wait(d.d_flags, (unsigned int)(c->pushes_done[0] * tiles_x), (unsigned int)(c->pushes_done[1] * tiles_x),
                             d.d_flags + ERR_FLAG_WORD, c->spin_timeout_ns, d.stream);
            c->launches++;
            e = cudaGetLastError();
            // Only the WHITE ghost rows come from the host: the first colour step (black) reads nothing else,
            // and the black ghost rows are written by the neighbours' first black step -- which may run ahead
            // of this upload, so a host copy of them could overwrite fresher data.
            const size_t row_bytes = (size_t)c->pitch * 4;
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(local_row(c, d, 1, -1), ghosts_in[k], row_bytes, cudaMemcpyHostToDevice, d.stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(local_row(c, d, 1, c->cfg.Y), ghosts_in[k] + c->pitch, row_bytes, cudaMemcpyHostToDevice,
                                    d.stream);
            if (e != cudaSuccess) break;
        }
        e = cudaMemcpyAsync(local_row(c, d, 0, 0), black_in[k], bytes, cudaMemcpyHostToDevice, d.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(local_row(c, d, 1, 0), white_in[k], bytes, cudaMemcpyHostToDevice, d.stream);
        // compute of job k starts when compute of job k-1 (other context) is done
        if (e == cudaSuccess && k > 0 && cs[(k - 1) % nctx] != c)
            e = cudaStreamWaitEvent(d.stream, cs[(k - 1) % nctx]->dev[0].ev_color, 0);
        if (e != cudaSuccess) break;
        rc = ising_sweeps(c, first_it, n);
        if (rc != ISING_OK) break;
        if (d_counts) {
            for (int col = 0; col < 2; col++) {
                launch_count(local_row(c, d, col, 0), c->cfg.Y, c->pitch, d_counts + k, d.stream);
                c->launches++;
            }
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventRecord(d.ev_color, d.stream);
        if (e == cudaSuccess && d_counts) e = cudaMemcpyAsync(h_counts + k, d_counts + k, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess && black_out && black_out[k]) e = cudaMemcpyAsync(black_out[k], local_row(c, d, 0, 0), bytes, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess && white_out && white_out[k]) e = cudaMemcpyAsync(white_out[k], local_row(c, d, 1, 0), bytes, cudaMemcpyDeviceToHost, d.stream);
    }
    for (ising_ctx* c : cs) {   // always drain every stream before the staging buffers go away
        cudaError_t es = cudaStreamSynchronize(c->dev[0].stream);
        if (e == cudaSuccess) e = es;
    }
    if (h_counts && rc == ISING_OK && e == cudaSuccess)
        for (int k = 0; k < nrep; k++) { up[k] = h_counts[k]; down[k] = spins - h_counts[k]; }
    if (rc != ISING_OK) return rc;
    if (e != cudaSuccess) return fail(ISING_ECUDA, "ising_run_batch_from_host: %s", cudaGetErrorString(e));
    return ISING_OK;
}

int ising_run_batch_from_host(ising_ctx* ctx_a, ising_ctx* ctx_b, int nrep, const uint32_t* const* black_in,
                              const uint32_t* const* white_in, int first_it, int n, uint32_t* const* black_out,
                              uint32_t* const* white_out, uint64_t* up, uint64_t* down)
{
    ising_ctx* cs[2] = {ctx_a, ctx_b ? ctx_b : ctx_a};
    return run_batch(cs, 2, nrep, black_in, white_in, nullptr, first_it, n, black_out, white_out, up, down);
}

int ising_run_batch_from_host_n(ising_ctx* const* ctxs, int nctx, int nrep, const uint32_t* const* black_in,
                                const uint32_t* const* white_in, const uint32_t* const* ghosts_in, int first_it, int n,
                                uint32_t* const* black_out, uint32_t* const* white_out, uint64_t* up, uint64_t* down)
{
    return run_batch(ctxs, nctx, nrep, black_in, white_in, ghosts_in, first_it, n, black_out, white_out, up, down);
}

int ising_run_slab_batch_from_host(ising_ctx* ctx_a, ising_ctx* ctx_b, int nrep, const uint32_t* const* black_in,
                                   const uint32_t* const* white_in, const uint32_t* const* ghosts_in, int first_it, int n,
                                   uint32_t* const* black_out, uint32_t* const* white_out, uint64_t* up, uint64_t* down)
{
    if (!ghosts_in) return fail(ISING_EINVAL, "ghosts_in is required");
    ising_ctx* cs[2] = {ctx_a, ctx_b ? ctx_b : ctx_a};
    return run_batch(cs, 2, nrep, black_in, white_in, ghosts_in, first_it, n, black_out, white_out, up, down);
}

int ising_host_alloc(size_t bytes, int write_combined, void** ptr)
{
    if (!ptr || !bytes) return fail(ISING_EINVAL, "bad argument");
    *ptr = nullptr;
    const unsigned int flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
    const cudaError_t e = cudaHostAlloc(ptr, bytes, flags);
    if (e != cudaSuccess) return fail(ISING_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return ISING_OK;
}

int ising_host_free(void* ptr)
{
    if (ptr) CU_TRY(cudaFreeHost(ptr));
    return ISING_OK;
}

int ising_row_ptr(ising_ctx* c, int dev, int color, int row, void** dptr, size_t* nbytes)
{
    if (int rc = check_dev(c, dev)) return rc;
    if ((color != 0 && color != 1) || row < -1 || row > c->cfg.Y || !dptr) return fail(ISING_EINVAL, "bad argument");
    *dptr = local_row(c, c->dev[dev], color, row);
    if (nbytes) *nbytes = (size_t)c->pitch * 4;
    return ISING_OK;
}

}  // extern "C"
