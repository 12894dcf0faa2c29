/*
 * tensorcore_oracle.c -- CPU restatement of the reference's tensorcore/ variant
 * (NVIDIA/ising-gpu, tensorcore/main.cu).  TEST INFRASTRUCTURE, NOT THE PRODUCT (same rules as
 * ising_oracle.c).
 *
 * Parity status: pinned against the reference binary (tensorcore/main.cu built for sm_100a with
 * only its final CUB reduce replaced, oracle/build_ref_tensorcore.py) run on a B200:
 * tests/golden/reference_tc_b200.json.  The reference evaluates exp() on the DEVICE
 * (tensorcore/main.cu:137, no fast-math); a CPU cannot reproduce CUDA's expf bit for bit in
 * general, so the acceptance values are an INPUT here (tc_set_accept): tests feed the five
 * values computed by the GPU library, CPU-only tests use glibc expf (self-consistent).
 *
 * Storage follows the reference (tensorcore/main.cu:37-44, :300-336): super-blocks of 256x256,
 * four 128x128 sub-blocks (a,b) at SUB_OFFSET(a,b) = (b*256 + a*128)*128, element (i,j) at
 * j*128+i  <->  true row bj*256 + 2j + b, true column bi*256 + 2i + a.  Black = (0,0),(1,1).
 * Spins are +1/-1 stored as int8 (the reference uses fp16 +-1).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   /* ising_oracle.c */
float ora_uniform(uint32_t x);

enum { SUP = 256, SUB = 128 };

typedef struct {
    int nbx, nby;                /* super-blocks along x (columns) and y (rows)            */
    long long nx, ny;
    unsigned long long seed;
    float accept[5];             /* by v = nn*s in {-4,-2,0,2,4} -> index (v+4)/2            */
    int8_t *lat;                 /* blocked storage order                                    */
} tc_lattice;

static inline long long sup_offset(long long i, long long j, long long nbx) { return (j * nbx + i) * SUP * SUP; }
static inline long long sub_offset(int a, int b) { return ((long long)b * SUP + (long long)a * SUB) * SUB; }

tc_lattice *tc_create(int nbx, int nby, unsigned long long seed, float alpha)
{
    if (nbx <= 0 || nby <= 0) return NULL;
    tc_lattice *L = (tc_lattice *)calloc(1, sizeof(*L));
    L->nbx = nbx; L->nby = nby; L->nx = (long long)nbx * SUP; L->ny = (long long)nby * SUP; L->seed = seed;
    L->lat = (int8_t *)malloc((size_t)(L->nx * L->ny));
    /* tensorcore/main.cu:578 inv_temp = 1/(alpha*TCRIT); :137 accept = exp(-2*inv_temp*nn*s) */
    const float inv_temp = 1.0f / (alpha * 2.26918531421f);
    for (int k = 0; k < 5; k++) L->accept[k] = expf(-2.0f * inv_temp * (float)(2 * k - 4));
    return L;
}
void tc_destroy(tc_lattice *L) { if (L) { free(L->lat); free(L); } }
void tc_set_accept(tc_lattice *L, const float a[5]) { memcpy(L->accept, a, sizeof(L->accept)); }
void tc_get_accept(const tc_lattice *L, float a[5]) { memcpy(a, L->accept, sizeof(L->accept)); }
int8_t *tc_storage(tc_lattice *L) { return L->lat; }

/* storage index of the spin at true (row, col) -- inverse of write_lattice, main.cu:323-334 */
static inline long long storage_index(const tc_lattice *L, long long row, long long col)
{
    const long long bj = row / SUP, bi = col / SUP;
    const int rr = (int)(row % SUP), cc = (int)(col % SUP);
    return sup_offset(bi, bj, L->nbx) + sub_offset(cc & 1, rr & 1) + (long long)(rr >> 1) * SUB + (cc >> 1);
}

static inline void draw4(unsigned long long seed, unsigned long long subseq, unsigned long long call, uint32_t out[4])
{
    const uint32_t ctr[4] = { (uint32_t)call, (uint32_t)(call >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32) };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    ora_philox4x32_10(ctr, key, out);
}

/* init_spins, main.cu:63-79: element tid of the storage gets curand_init(seed, tid, 0), first
 * uniform < 0.5 -> -1 else +1. */
void tc_init(tc_lattice *L)
{
    const long long n = L->nx * L->ny;
    #pragma omp parallel for schedule(static)
    for (long long t = 0; t < n; t++) {
        uint32_t o[4];
        draw4(L->seed, (unsigned long long)t, 0, o);
        L->lat[t] = (ora_uniform(o[0]) < 0.5f) ? -1 : 1;
    }
}

/* One colour of update(), main.cu:237-298 = neighbour sums (4 GEMMs + add_boundaries, i.e. the
 * periodic 4-neighbour sum) + update_spins (main.cu:88-146).  iter = n+1 counted over warm-up
 * and trial iterations (main.cu:582-593).  RNG offset = (2*iter + colour)*16 draws (main.cu:269,
 * :294), 16 sequential draws per thread: 2 runs of 8 consecutive storage elements. */
void tc_update_color(tc_lattice *L, int is_black, unsigned long long iter)
{
    const long long tps = (long long)SUB * SUB / 16;                    /* threads per sub-block */
    const long long nthreads = 2LL * L->nbx * L->nby * tps;
    const unsigned long long call0 = (2ULL * iter + (is_black ? 0 : 1)) * 4ULL;   /* 16 draws = 4 calls */
    #pragma omp parallel for schedule(static)
    for (long long tid = 0; tid < nthreads; tid++) {
        const int bi = (int)((tid / tps) % (2 * L->nbx));
        int bj = (int)(tid / (tps * 2 * L->nbx));
        const long long tl = tid % tps;
        if (is_black) bj = (bi % 2) ? 2 * bj + 1 : 2 * bj;
        else          bj = (bi % 2) ? 2 * bj : 2 * bj + 1;
        const int a = bi % 2, b = bj % 2;
        const long long base = sup_offset(bi / 2, bj / 2, L->nbx) + sub_offset(a, b);
        uint32_t o[4];
        for (int n = 0; n < 2; n++) {
            for (int m = 0; m < 8; m++) {
                if ((m & 3) == 0) draw4(L->seed, (unsigned long long)tid, call0 + 2 * n + (m >> 2), o);
                const long long e = (tl + n * tps) * 8 + m;
                const int i = (int)(e % SUB), j = (int)(e / SUB);
                const long long row = (long long)(bj / 2) * SUP + 2 * j + b, col = (long long)(bi / 2) * SUP + 2 * i + a;
                const long long ru = (row + L->ny - 1) % L->ny, rd = (row + 1) % L->ny;
                const long long cl = (col + L->nx - 1) % L->nx, cr = (col + 1) % L->nx;
                const int nn = L->lat[storage_index(L, ru, col)] + L->lat[storage_index(L, rd, col)] +
                               L->lat[storage_index(L, row, cl)] + L->lat[storage_index(L, row, cr)];
                const int s = L->lat[base + e];
                if (ora_uniform(o[m & 3]) < L->accept[(nn * s + 4) / 2]) L->lat[base + e] = (int8_t)-s;
            }
        }
    }
}

void tc_update(tc_lattice *L, unsigned long long iter)
{
    tc_update_color(L, 1, iter);
    tc_update_color(L, 0, iter);
}

long long tc_sum(const tc_lattice *L)
{
    long long s = 0;
    const long long n = L->nx * L->ny;
    for (long long t = 0; t < n; t++) s += L->lat[t];
    return s;
}

/* true-order lattice (ny rows x nx columns), as write_lattice lays it out (main.cu:317-336) */
void tc_export_true(const tc_lattice *L, int8_t *out)
{
    for (long long r = 0; r < L->ny; r++)
        for (long long c = 0; c < L->nx; c++) out[r * L->nx + c] = L->lat[storage_index(L, r, c)];
}

/* 4-neighbour sums of one colour in storage order (0 at the other colour's elements): what the
 * GEMMs + add_boundaries leave in nn_sums (main.cu:252-266).  Used to test the MMA kernel. */
void tc_neighbour_sums(const tc_lattice *L, int is_black, int8_t *out)
{
    memset(out, 0, (size_t)(L->nx * L->ny));
    for (long long r = 0; r < L->ny; r++)
        for (long long c = 0; c < L->nx; c++) {
            if ((((r + c) & 1) == 0) != (is_black != 0)) continue;
            const long long ru = (r + L->ny - 1) % L->ny, rd = (r + 1) % L->ny;
            const long long cl = (c + L->nx - 1) % L->nx, cr = (c + 1) % L->nx;
            out[storage_index(L, r, c)] = (int8_t)(L->lat[storage_index(L, ru, c)] + L->lat[storage_index(L, rd, c)] +
                                                  L->lat[storage_index(L, r, cl)] + L->lat[storage_index(L, r, cr)]);
        }
}

/* final.txt, main.cu:338-348: "<value> " per spin, values printed as floats (1 / -1) */
int tc_write_lattice(const tc_lattice *L, const char *fname)
{
    FILE *f = fopen(fname, "w");
    if (!f) return -1;
    for (long long r = 0; r < L->ny; r++) {
        for (long long c = 0; c < L->nx; c++) fputs(L->lat[storage_index(L, r, c)] > 0 ? "1 " : "-1 ", f);
        fputc('\n', f);
    }
    fclose(f);
    return 0;
}
