// pipe_map.cu -- which SM pipe executes the integer instructions the Metropolis kernel is made
// of, and at what rate.  Run under ncu:
//   ncu --metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,\
//       sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fp64.sum,\
//       sm__inst_executed_pipe_xu.sum,sm__inst_executed.sum,smsp__cycles_active.avg,\
//       smsp__issue_active.avg.pct_of_peak_sustained_active --csv ./pipe_map
// Each kernel = 8 independent dependency chains per thread of ONE PTX instruction, 4096 x 4
// repetitions, 148*4 CTAs of 256 threads, so per-SMSP issue is saturated (8 warps x 8 chains).
// It also prints achieved warp-instructions/clk/SMSP from clock64 (kernels are ~ms long so the
// clock has ramped).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8

#define KERNEL(NAME, DECL, BODY)                                                            \
    __global__ void __launch_bounds__(256, 4) NAME(uint32_t* out, uint32_t s, long long* cyc) \
    {                                                                                       \
        uint32_t a[CH], b[CH];                                                              \
        uint64_t w[CH];                                                                     \
        double d[CH];                                                                       \
        float f[CH];                                                                        \
        _Pragma("unroll") for (int i = 0; i < CH; i++) {                                    \
            a[i] = s * (i + 3) + threadIdx.x; b[i] = s ^ (i * 0x9E3779B9u);                  \
            w[i] = ((uint64_t)a[i] << 32) | b[i]; d[i] = (double)a[i]; f[i] = (float)b[i];    \
        }                                                                                   \
        DECL;                                                                               \
        __syncthreads();                                                                    \
        const long long t0 = clock64();                                                     \
        _Pragma("unroll 1") for (int it = 0; it < ITERS; it++) {                            \
            _Pragma("unroll") for (int r = 0; r < 4; r++) {                                 \
                _Pragma("unroll") for (int i = 0; i < CH; i++) { BODY; }                    \
            }                                                                               \
        }                                                                                   \
        const long long t1 = clock64();                                                     \
        uint32_t x = 0;                                                                     \
        _Pragma("unroll") for (int i = 0; i < CH; i++)                                      \
            x ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i] ^ (uint32_t)f[i]; \
        out[blockIdx.x * 256 + threadIdx.x] = x;                                            \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                    \
    }

KERNEL(k_imad_wide, const uint32_t M = s | 0xD2511F53u,
       asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[i]) : "r"(M)))
KERNEL(k_mul_wide, const uint32_t M = s | 0xD2511F53u,
       asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, hi, %1; }" : "+l"(w[i]) : "r"(M)))
KERNEL(k_imad_hi, const uint32_t M = s | 0xD2511F53u,
       asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(M), "r"(b[i])))
KERNEL(k_imad_lo, const uint32_t M = s | 0xD2511F53u,
       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(M), "r"(b[i])))
KERNEL(k_lop3, const uint32_t M = s | 0xD2511F53u,
       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(M), "r"(b[i])))
KERNEL(k_iadd3, const uint32_t M = s | 0xD2511F53u,
       asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])))
KERNEL(k_viadd, const uint32_t M = s,
       asm volatile("add.u32 %0, %0, 0x40;" : "+r"(a[i])))
KERNEL(k_iadd3_3, const uint32_t M = s,
       asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(M)))
KERNEL(k_carry_pair, const uint32_t M = s | 0xD2511F53u,
       asm volatile("{ .reg .u32 t; add.cc.u32 t, %1, %2; addc.u32 %0, %0, %0; }" : "+r"(a[i]) : "r"(a[(i + 1) % CH]), "r"(M)))
KERNEL(k_setp_por, const uint32_t M = s | 0xD2511F53u,
       asm volatile("{ .reg .pred p; setp.gt.u32 p, %1, %2; @p or.b32 %0, %0, 0x40; }" : "+r"(b[i]) : "r"(a[(i + 1) % CH] + it), "r"(M)))
KERNEL(k_shf, const uint32_t M = s | 0xD2511F53u,
       asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a[i]) : "r"(b[i])))
KERNEL(k_prmt, const uint32_t M = s | 0xD2511F53u,
       asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a[i]) : "r"(b[i])))
KERNEL(k_ffma, const float M = (float)s,
       asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(M)))
KERNEL(k_dfma, const double M = (double)s,
       asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(M)))
KERNEL(k_popc, const uint32_t M = s,
       asm volatile("{ .reg .u32 t; popc.b32 t, %0; add.u32 %0, t, %1; }" : "+r"(a[i]) : "r"(b[i])))
KERNEL(k_imnmx, const uint32_t M = s,
       asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i] + it)))
KERNEL(k_sel, const uint32_t M = s,
       asm volatile("{ .reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a[i]) : "r"(b[i]), "r"(M + it)))
// philox round as the kernel does it: 2 wide multiplies + 2 three-input xors
#define PHILOX_BODY                                                                                   \
    do {                                                                                              \
        const uint64_t p0 = (uint64_t)a[i] * 0xD2511F53u;                                             \
        const uint64_t p1 = (uint64_t)(uint32_t)w[i] * 0xCD9E8D57u;                                   \
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b[i] ^ M;                                          \
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ (uint32_t)(w[i] >> 32) ^ (M * 9u);                 \
        a[i] = n0; b[i] = (uint32_t)p1; w[i] = ((uint64_t)(uint32_t)p0 << 32) | n2;                   \
    } while (0)
KERNEL(k_philox_round, const uint32_t M = s * 5u, PHILOX_BODY)

typedef void (*kern_t)(uint32_t*, uint32_t, long long*);

static void run(const char* name, kern_t k, int per_step, uint32_t* out, long long* cyc, long long* h, int grid)
{
    k<<<grid, 256>>>(out, 12345u, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, 256>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)h[i]; avg /= grid;
    // 4 CTAs x 8 warps per SM = 8 warps per SMSP, all resident (grid = 148*4)
    const double warp_instr_per_smsp = 8.0 * ITERS * 4 * CH * per_step;
    printf("%-16s %8.3f ms  cycles/CTA %10.0f  intended warp-instr/clk/SMSP %.3f  (cycles per warp-instr %.2f)\n", name,
           ms, avg, warp_instr_per_smsp / avg, avg / warp_instr_per_smsp);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 4;
    uint32_t* out; long long* cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    long long* h = (long long*)malloc(grid * 8);
    printf("device %s, %d SMs, grid %d x 256\n", p.name, p.multiProcessorCount, grid);
    run("imad_wide", k_imad_wide, 1, out, cyc, h, grid);
    run("mul_wide", k_mul_wide, 1, out, cyc, h, grid);
    run("imad_hi", k_imad_hi, 1, out, cyc, h, grid);
    run("imad_lo", k_imad_lo, 1, out, cyc, h, grid);
    run("lop3", k_lop3, 1, out, cyc, h, grid);
    run("iadd3", k_iadd3, 1, out, cyc, h, grid);
    run("viadd", k_viadd, 1, out, cyc, h, grid);
    run("iadd3_3", k_iadd3_3, 1, out, cyc, h, grid);
    run("carry_pair", k_carry_pair, 2, out, cyc, h, grid);
    run("setp_por", k_setp_por, 2, out, cyc, h, grid);
    run("shf", k_shf, 1, out, cyc, h, grid);
    run("prmt", k_prmt, 1, out, cyc, h, grid);
    run("ffma", k_ffma, 1, out, cyc, h, grid);
    run("dfma", k_dfma, 1, out, cyc, h, grid);
    run("popc", k_popc, 2, out, cyc, h, grid);
    run("imnmx", k_imnmx, 1, out, cyc, h, grid);
    run("sel", k_sel, 2, out, cyc, h, grid);
    run("philox_round", k_philox_round, 4, out, cyc, h, grid);
    return 0;
}
