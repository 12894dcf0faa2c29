/*
 * ising_oracle.c -- CPU restatement of the reference's checkerboard Metropolis
 * hot path (NVIDIA/ising-gpu, optimized/main.cu).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may
 * load it.  The product (libising_b200.so) never links or calls anything here.
 *
 * Parity status: PINNED.  The functions below reproduce the known-answer spin
 * counts printed in the reference's optimized/README.md (lines 128-137, 187-196,
 * 240, 307) -- see tests/test_oracle_kat.py and oracle/pin_kat.py.
 *
 * Representation: one byte per spin (0 = down, 1 = up), two colour planes of
 * rows x (X/2) bytes.  Plain scalar loops, float acceptance exactly as the
 * reference does it (u <= exp_table[s][n]); no integer-threshold shortcut here,
 * so that the shortcut used by the CUDA kernels is checked against the real
 * float path.
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- geometry of the reference launch (optimized/main.cu:55-58, 1547-1550) ---- */
enum {
    REF_BDIM_X = 16,           /* threads per block in x                                  */
    REF_BDIM_Y = 16,           /* threads per block in y == rows per block                */
    REF_LOOP_X = 2,            /* 128-bit words per thread                                */
    REF_SPW    = 16,           /* spins per 64-bit word at 4 bit/spin (main.cu:40, 1243)  */
    REF_TID_SPINS = 2 * REF_SPW * REF_LOOP_X   /* 64 spins per reference thread           */
};

typedef struct {
    int       X;        /* lattice columns, all colours (main.cu:1245)                    */
    long long rows;     /* total rows over all devices = ndev * Y (main.cu:1591)          */
    int       xsl;      /* sub-lattice width in lattice columns  (== X when disabled)     */
    long long ysl;      /* sub-lattice height in rows            (== rows when disabled)  */
    long long half;     /* X/2 = sites of one colour per row                              */
    uint8_t  *plane[2]; /* [0]=black, [1]=white, rows*half bytes each                     */
    uint8_t  *bond[2];  /* optional +-J bond nibbles per site (NULL when -J unused)       */
    /* slab mode (multi-process tests): this object holds rows [grow0, grow0+rows) of a taller
     * lattice; the rows just outside come from ghost buffers instead of the periodic wrap.    */
    long long grow0;
    uint8_t  *ghost_top[2], *ghost_bot[2];   /* per colour, half bytes each; NULL = periodic    */
} ora_lattice;

/* ---- Philox4x32-10: public algorithm (Salmon et al. 2011), as shipped in the CUDA
 * toolkit's curand_philox4x32_x.h:88-91 (constants), :160-192 (round + 10 rounds).
 * cuRAND is a toolkit dependency of the reference, not vendored; container has 10.3.10. */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* curand_uniform for Philox (toolkit curand_uniform.h:69-72): x*2^-32 + 2^-33,
 * evaluated in float.  (float)x rounds to nearest; the multiply by 2^-32 is exact,
 * so a fused or unfused add gives the same float.  Range (0, 1]. */
float ora_uniform(uint32_t x)
{
    volatile float f = (float)x;           /* keep it a true float, no excess precision */
    volatile float m = f * 2.3283064e-10f;
    return m + (2.3283064e-10f / 2.0f);
}

/* Acceptance table, optimized/main.cu:1683-1697 (host, glibc expf). */
void ora_exp_table(float temp, float tab[2][5])
{
    for (int i = 0; i < 2; i++) {
        for (int j = 0; j < 5; j++) {
            if (temp > 0) {
                tab[i][j] = expf((i ? -2.0f : 2.0f) * (float)(j * 2 - 4) * (1.0f / temp));
            } else {
                tab[i][j] = (j == 2) ? 0.5f : (i ? -2.0f : 2.0f) * (float)(j * 2 - 4);
            }
        }
    }
}

/* ---- lattice lifetime ---- */
ora_lattice *ora_create(int X, long long rows, int xsl, long long ysl)
{
    if (X <= 0 || rows <= 0 || (X % 2048) || (rows % 16)) return NULL;   /* main.cu:1412-1421 */
    if (xsl <= 0) xsl = X;
    if (ysl <= 0) ysl = rows;
    if ((X % xsl) || (xsl % 2048) || (rows % ysl) || (ysl % 16)) return NULL; /* :1440-1453 */
    ora_lattice *L = (ora_lattice *)calloc(1, sizeof(*L));
    if (!L) return NULL;
    L->X = X; L->rows = rows; L->xsl = xsl; L->ysl = ysl; L->half = X / 2;
    for (int c = 0; c < 2; c++) {
        L->plane[c] = (uint8_t *)calloc((size_t)rows * (size_t)L->half, 1);
        if (!L->plane[c]) { free(L->plane[0]); free(L); return NULL; }
    }
    return L;
}

/* Turn the object into a slab starting at global row grow0, with ghost rows (zeroed). */
int ora_make_slab(ora_lattice *L, long long grow0)
{
    L->grow0 = grow0;
    for (int c = 0; c < 2; c++) {
        L->ghost_top[c] = (uint8_t *)calloc((size_t)L->half, 1);
        L->ghost_bot[c] = (uint8_t *)calloc((size_t)L->half, 1);
        if (!L->ghost_top[c] || !L->ghost_bot[c]) return -1;
    }
    return 0;
}
uint8_t *ora_ghost(ora_lattice *L, int color, int bottom) { return bottom ? L->ghost_bot[color] : L->ghost_top[color]; }

void ora_destroy(ora_lattice *L)
{
    if (!L) return;
    free(L->plane[0]); free(L->plane[1]);
    for (int c = 0; c < 2; c++) { free(L->ghost_top[c]); free(L->ghost_bot[c]); }
    free(L->bond[0]);  free(L->bond[1]);
    free(L);
}

uint8_t *ora_plane(ora_lattice *L, int color) { return L->plane[color]; }

/* Reference thread id of the thread that owns (row r, 128-bit word c) -- main.cu:107-113
 * and :511-515.  Evaluated in unsigned 32-bit arithmetic like the reference expression
 * (int * unsigned -> unsigned), then widened. */
static inline uint32_t ref_tid(const ora_lattice *L, long long r, long long c)
{
    const uint32_t gridX = (uint32_t)(L->half / 32 / (REF_BDIM_X * REF_LOOP_X));
    r += L->grow0;                                            /* global row of a slab's local row */
    const uint32_t brow  = (uint32_t)(r / REF_BDIM_Y);        /* devid*gridDim.y + blockIdx.y */
    const uint32_t bx    = (uint32_t)(c / (REF_BDIM_X * REF_LOOP_X));
    const uint32_t tx    = (uint32_t)(c % REF_BDIM_X);
    const uint32_t ty    = (uint32_t)(r % REF_BDIM_Y);
    return (brow * gridX + bx) * (REF_BDIM_X * REF_BDIM_Y) + ty * REF_BDIM_X + tx;
}

/* The four uniforms' raw 32-bit values of Philox call `call` (0..15) made by reference
 * thread `tid` at step `s = 2*it + color`.  curand_init(seed, tid, 64*s) -- main.cu:115-116,
 * :620-621 -> counter = (lo,hi of 16*s + call ; lo,hi of tid), key = seed
 * (toolkit curand_kernel.h:1022-1037, :971-981). */
static inline void ref_draw4(unsigned long long seed, uint32_t tid, unsigned long long s,
                             int call, uint32_t out[4])
{
    const unsigned long long ctr_lo = 16ull * s + (unsigned long long)call;
    const uint32_t ctr[4] = { (uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), tid, 0u };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    ora_philox4x32_10(ctr, key, out);
}

/* Site order of the 4 outputs of call (j, q): .x nibble 2q, .y nibble 2q, .x nibble 2q+1,
 * .y nibble 2q+1 of 128-bit word c0 + 16*j  (draw loop main.cu:127-140, :637-659).
 * In colour-column units: m = 32*c + {2q, 16+2q, 2q+1, 17+2q}. */
static const int LANE_OFF[4] = { 0, 16, 1, 17 };

/* latticeInit_k, main.cu:92-151: spin = (uniform < 0.5f), step s = color (it = 0). */
void ora_init_color(ora_lattice *L, int color, unsigned long long seed)
{
    const long long nwords = L->half / 32;           /* 128-bit words per colour row */
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *row = L->plane[color] + (size_t)r * L->half;
        for (long long c = 0; c < nwords; c++) {
            /* word c is loop index j of thread with tx = c%16 in block c/32 */
            const int j = (int)((c % 32) / 16);
            const uint32_t tid = ref_tid(L, r, c);
            for (int q = 0; q < 8; q++) {
                uint32_t x[4];
                ref_draw4(seed, tid, (unsigned long long)color, 8 * j + q, x);
                for (int l = 0; l < 4; l++)
                    row[32 * c + 2 * q + LANE_OFF[l]] = (ora_uniform(x[l]) < 0.5f) ? 1 : 0;
            }
        }
    }
}

/* Periodic neighbours inside a sub-lattice (loadTile wrap, main.cu:414, :422, :433, :441).
 * Rows wrap at ysl, colour-columns wrap at xsl/2. */
static inline long long wrap_row(const ora_lattice *L, long long r, int d)
{
    const long long base = (r / L->ysl) * L->ysl;
    long long o = r - base + d;
    if (o < 0) o += L->ysl;
    if (o >= L->ysl) o -= L->ysl;
    return base + o;
}
static inline long long wrap_col(const ora_lattice *L, long long m, int d)
{
    const long long w = L->xsl / 2;
    const long long base = (m / w) * w;
    long long o = m - base + d;
    if (o < 0) o += w;
    if (o >= w) o -= w;
    return base + o;
}

/* spinUpdateV_2D_k, main.cu:463-670.  Neighbours of (r, m) in the other colour plane:
 * (r-1,m), (r+1,m), (r,m) and (r,m-1) when readBack else (r,m+1); readBack = !(r&1) for
 * black, (r&1) for white (main.cu:542-573).  n = number of up neighbours (nibble sums
 * main.cu:623-635).  Flip iff uniform <= tab[s][n] (main.cu:652-657).
 * With bonds (main.cu:575-618): nibble bits <up,down,left,right> = <8,4,2,1> XOR the
 * neighbour spin before summing. */
void ora_update_color(ora_lattice *L, int color, unsigned long long seed, int it,
                      const float tab[2][5])
{
    const long long nwords = L->half / 32;
    const unsigned long long s = 2ull * (unsigned long long)it + (unsigned long long)color;
    const uint8_t *src = L->plane[color ^ 1];
    /* the reference hands the BLACK update the buffer filled by hamiltInitW_k and the WHITE
     * update the one filled by hamiltInitB_k (main.cu:1775, :1796): kept as is */
    const uint8_t *bnd = L->bond[color ^ 1];
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *row = L->plane[color] + (size_t)r * L->half;
        const uint8_t *up = src + (size_t)wrap_row(L, r, -1) * L->half;
        const uint8_t *ct = src + (size_t)r * L->half;
        const uint8_t *dw = src + (size_t)wrap_row(L, r, +1) * L->half;
        if (r == 0 && L->ghost_top[color ^ 1]) up = L->ghost_top[color ^ 1];
        if (r == L->rows - 1 && L->ghost_bot[color ^ 1]) dw = L->ghost_bot[color ^ 1];
        const int readBack = (color == 0) ? !(r & 1) : (int)(r & 1);
        for (long long c = 0; c < nwords; c++) {
            const int j = (int)((c % 32) / 16);
            const uint32_t tid = ref_tid(L, r, c);
            for (int q = 0; q < 8; q++) {
                uint32_t x[4];
                ref_draw4(seed, tid, s, 8 * j + q, x);
                for (int l = 0; l < 4; l++) {
                    const long long m = 32 * c + 2 * q + LANE_OFF[l];
                    const long long ms = wrap_col(L, m, readBack ? -1 : +1);
                    int nu = up[m], nd = dw[m], nc = ct[m], ns = ct[ms];
                    if (bnd) {
                        const int J = bnd[(size_t)r * L->half + m];
                        nu ^= (J >> 3) & 1;
                        nd ^= (J >> 2) & 1;
                        if (readBack) { ns ^= (J >> 1) & 1; nc ^= J & 1; }   /* sd=left, ct=right */
                        else          { nc ^= (J >> 1) & 1; ns ^= J & 1; }   /* ct=left, sd=right */
                    }
                    const int n = nu + nd + nc + ns;
                    const int sp = row[m];
                    if (ora_uniform(x[l]) <= tab[sp][n]) row[m] = (uint8_t)(sp ^ 1);
                }
            }
        }
    }
}

/* One full sweep j (0-based): black then white with it = j+1 (main.cu:1763-1799). */
void ora_sweep(ora_lattice *L, unsigned long long seed, int j, float temp)
{
    float tab[2][5];
    ora_exp_table(temp, tab);
    ora_update_color(L, 0, seed, j + 1, tab);
    ora_update_color(L, 1, seed, j + 1, tab);
}

/* countSpins / getMagn_k, main.cu:702-734, :831-868. */
void ora_count(const ora_lattice *L, unsigned long long *up, unsigned long long *down)
{
    unsigned long long u = 0;
    const size_t n = (size_t)L->rows * (size_t)L->half;
    for (int c = 0; c < 2; c++) {
        const uint8_t *p = L->plane[c];
        unsigned long long uc = 0;
        #pragma omp parallel for reduction(+:uc) schedule(static)
        for (long long i = 0; i < (long long)n; i++) uc += p[i];
        u += uc;
    }
    *up = u;
    *down = 2ull * n - u;
}

/* Nearest-neighbour energy per site, e = -(1/N) sum_i s_i (s_right + s_down), s = +-1,
 * periodic inside each sub-lattice (observable added by the new framework; the reference
 * exposes it only through column 1 of the -c file, main.cu:1129-1135). */
double ora_energy(const ora_lattice *L)
{
    long long acc = 0;
    #pragma omp parallel for reduction(+:acc) schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        const long long rd = wrap_row(L, r, +1);
        for (int color = 0; color < 2; color++) {
            const uint8_t *me  = L->plane[color]     + (size_t)r  * L->half;
            const uint8_t *oth = L->plane[color ^ 1] + (size_t)r  * L->half;
            const uint8_t *dwn = L->plane[color ^ 1] + (size_t)rd * L->half;
            /* right neighbour of (r,m) of this colour: other colour, same row, at m when the
             * site sits on the even lattice column of the pair, else at m+1 */
            const int right_is_same_m = (color == 0) ? !(r & 1) : (int)(r & 1);
            for (long long m = 0; m < L->half; m++) {
                const int s  = 2 * me[m] - 1;
                const int sr = 2 * oth[right_is_same_m ? m : wrap_col(L, m, +1)] - 1;
                const int sd = 2 * dwn[m] - 1;
                acc += s * (sr + sd);
            }
        }
    }
    return -(double)acc / ((double)L->rows * (double)L->X);
}

/* Export one colour plane in the reference's storage format: 64-bit words, 16 spins per
 * word, nibble k = bits 4k..4k+3, bit 0 of the nibble = spin (main.cu:1160-1175 decode;
 * :131-138 encode).  out must hold rows * X/32 words. */
void ora_export_ref(const ora_lattice *L, int color, unsigned long long *out)
{
    const long long lld = L->half / REF_SPW;
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        const uint8_t *row = L->plane[color] + (size_t)r * L->half;
        for (long long w = 0; w < lld; w++) {
            unsigned long long v = 0;
            for (int k = 0; k < REF_SPW; k++)
                v |= (unsigned long long)(row[w * REF_SPW + k] & 1) << (4 * k);
            out[(size_t)r * lld + w] = v;
        }
    }
}

void ora_import_ref(ora_lattice *L, int color, const unsigned long long *in)
{
    const long long lld = L->half / REF_SPW;
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *row = L->plane[color] + (size_t)r * L->half;
        for (long long w = 0; w < lld; w++) {
            const unsigned long long v = in[(size_t)r * lld + w];
            for (int k = 0; k < REF_SPW; k++) row[w * REF_SPW + k] = (uint8_t)((v >> (4 * k)) & 1);
        }
    }
}

/* Export one colour plane bit-packed, 32 sites per uint32 (bit b of word w = site 32w+b):
 * the framework's device layout, used by tests to compare buffers directly. */
void ora_export_bits(const ora_lattice *L, int color, uint32_t *out)
{
    const long long nw = L->half / 32;
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        const uint8_t *row = L->plane[color] + (size_t)r * L->half;
        for (long long w = 0; w < nw; w++) {
            uint32_t v = 0;
            for (int b = 0; b < 32; b++) v |= (uint32_t)(row[32 * w + b] & 1) << b;
            out[(size_t)r * nw + w] = v;
        }
    }
}

void ora_import_bits(ora_lattice *L, int color, const uint32_t *in)
{
    const long long nw = L->half / 32;
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *row = L->plane[color] + (size_t)r * L->half;
        for (long long w = 0; w < nw; w++) {
            const uint32_t v = in[(size_t)r * nw + w];
            for (int b = 0; b < 32; b++) row[32 * w + b] = (uint8_t)((v >> b) & 1);
        }
    }
}

/* dumpLattice text format, main.cu:1160-1176: one line per row, X hex digits (0/1), colours
 * interleaved by row parity (odd rows start with white).  Rows [row0, row0+nrows). */
int ora_dump(const ora_lattice *L, long long row0, long long nrows, const char *fname)
{
    FILE *fp = fopen(fname, "w");
    if (!fp) return -1;
    char *line = (char *)malloc((size_t)L->X + 2);
    for (long long r = row0; r < row0 + nrows; r++) {
        const uint8_t *b = L->plane[0] + (size_t)r * L->half;
        const uint8_t *w = L->plane[1] + (size_t)r * L->half;
        for (long long m = 0; m < L->half; m++) {
            const char cb = (char)('0' + b[m]), cw = (char)('0' + w[m]);
            line[2 * m]     = (r & 1) ? cw : cb;
            line[2 * m + 1] = (r & 1) ? cb : cw;
        }
        line[L->X] = '\n';
        fwrite(line, 1, (size_t)L->X + 1, fp);
    }
    free(line);
    fclose(fp);
    return 0;
}



/* +-J bonds.  hamiltInitB_k (main.cu:153-212): every black site gets 4 random bits
 * <up,down,left,right> = <8,4,2,1> from curand_init(seed+1, tid, 0); draw order per 128-bit
 * word: nibble z = 0..15, bit l = 0..3, .x then .y -> draw index 128*j + 8*z + 2*l + h; bit set
 * iff uniform < prob.  hamiltInitW_k (main.cu:214-331): a white site's bits mirror the bonds
 * of its four black neighbours (its up bond is the down bond of the black site above, ...),
 * periodic inside the sub-lattice.  The seed handed in is the run's seed; +1 is applied here. */
int ora_init_bonds(ora_lattice *L, unsigned long long seed, float prob)
{
    const size_t n = (size_t)L->rows * (size_t)L->half;
    for (int c = 0; c < 2; c++) {
        free(L->bond[c]);
        L->bond[c] = (uint8_t *)calloc(n, 1);
        if (!L->bond[c]) return -1;
    }
    const long long nwords = L->half / 32;
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *row = L->bond[0] + (size_t)r * L->half;
        for (long long c = 0; c < nwords; c++) {
            const int j = (int)((c % 32) / 16);
            const uint32_t tid = ref_tid(L, r, c);
            for (int z = 0; z < 16; z++)
                for (int lh = 0; lh < 2; lh++) {
                    uint32_t x[4];
                    const unsigned long long call = (unsigned long long)(32 * j + 2 * z + lh);
                    const uint32_t ctr[4] = { (uint32_t)call, 0u, tid, 0u };
                    const unsigned long long s1 = seed + 1ull;
                    const uint32_t key[2] = { (uint32_t)s1, (uint32_t)(s1 >> 32) };
                    ora_philox4x32_10(ctr, key, x);
                    for (int q = 0; q < 4; q++) {          /* lane q: bit l = 2*lh + q/2, half h = q%2 */
                        const int l = 2 * lh + q / 2, h = q % 2;
                        if (ora_uniform(x[q]) < prob) row[32 * c + 16 * h + z] |= (uint8_t)(1u << l);
                    }
                }
        }
    }
    /* white = mirror of the black neighbours */
    #pragma omp parallel for schedule(static)
    for (long long r = 0; r < L->rows; r++) {
        uint8_t *w = L->bond[1] + (size_t)r * L->half;
        const uint8_t *bu = L->bond[0] + (size_t)wrap_row(L, r, -1) * L->half;
        const uint8_t *bd = L->bond[0] + (size_t)wrap_row(L, r, +1) * L->half;
        const uint8_t *bc = L->bond[0] + (size_t)r * L->half;
        const int odd = (int)((r + L->grow0) & 1);
        for (long long m = 0; m < L->half; m++) {
            /* even rows: white (col 2m+1) has black m on its left and black m+1 on its right;
             * odd rows: white (col 2m) has black m-1 on its left and black m on its right */
            const long long ml = odd ? wrap_col(L, m, -1) : m, mr = odd ? m : wrap_col(L, m, +1);
            uint8_t v = 0;
            v |= (uint8_t)(((bu[m] >> 2) & 1) << 3);    /* my up    = down bond of the black above  */
            v |= (uint8_t)(((bd[m] >> 3) & 1) << 2);    /* my down  = up bond of the black below    */
            v |= (uint8_t)(((bc[ml] >> 0) & 1) << 1);   /* my left  = right bond of the black left  */
            v |= (uint8_t)(((bc[mr] >> 1) & 1) << 0);   /* my right = left bond of the black right  */
            w[m] = v;
        }
    }
    return 0;
}

uint8_t *ora_bonds(ora_lattice *L, int color) { return L->bond[color]; }

/* Two-point correlation of getCorr2D_k / getCorr2DRepl_k (main.cu:870-1070) as a plain
 * definition: out[j-1] = sum over all sites of (+1 if equal else -1) for the site j columns
 * to the right and the site j rows below, j = 1..ncorr, periodic inside the sub-lattice,
 * divided by 2*X*rows (computeCorr, main.cu:1129-1135).  The reference kernel equals this
 * definition when X (or xsl) is a multiple of 4096; for other widths its chunk loop counts
 * columns twice (chunkDimX = 4096 regardless of X) -- that quirk is not restated. */
void ora_correlation(const ora_lattice *L, int ncorr, double *out)
{
    const long long X = L->X;
    for (int j = 1; j <= ncorr; j++) {
        long long acc = 0;
        #pragma omp parallel for reduction(+:acc) schedule(static)
        for (long long r = 0; r < L->rows; r++) {
            const long long sl_r0 = (r / L->ysl) * L->ysl;
            const long long rj = sl_r0 + (r - sl_r0 + j) % L->ysl;
            for (long long c = 0; c < X; c++) {
                const long long sl_c0 = (c / L->xsl) * L->xsl;
                const long long cj = sl_c0 + (c - sl_c0 + j) % L->xsl;
                #define SPIN(rr, cc) (L->plane[(((rr) + (cc)) & 1) ? 1 : 0][(size_t)(rr) * L->half + ((cc) >> 1)])
                const int s = SPIN(r, c);
                acc += (s == SPIN(r, cj)) ? 1 : -1;
                acc += (s == SPIN(rj, c)) ? 1 : -1;
                #undef SPIN
            }
        }
        out[j - 1] = (double)acc / (2.0 * (double)X * (double)L->rows);
    }
}

/* Streaming variant of ora_init_color + ora_count for lattices too large to hold as bytes:
 * number of up spins the reference's latticeInit_k (main.cu:92-151) produces for both
 * colours of an X x rows lattice.  Used to pin README KAT-2/KAT-3 initial counts. */
unsigned long long ora_init_count(int X, long long rows, unsigned long long seed)
{
    ora_lattice G; memset(&G, 0, sizeof(G));
    G.X = X; G.rows = rows; G.half = X / 2; G.xsl = X; G.ysl = rows;
    const long long nwords = G.half / 32;
    unsigned long long up = 0;
    #pragma omp parallel for reduction(+:up) schedule(static)
    for (long long r = 0; r < rows; r++) {
        for (int color = 0; color < 2; color++) {
            for (long long c = 0; c < nwords; c++) {
                const int j = (int)((c % 32) / 16);
                const uint32_t tid = ref_tid(&G, r, c);
                for (int q = 0; q < 8; q++) {
                    uint32_t x[4];
                    ref_draw4(seed, tid, (unsigned long long)color, 8 * j + q, x);
                    for (int l = 0; l < 4; l++) up += (ora_uniform(x[l]) < 0.5f) ? 1 : 0;
                }
            }
        }
    }
    return up;
}

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
