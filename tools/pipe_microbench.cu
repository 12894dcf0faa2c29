// Pipe-throughput microbenchmark for the instruction mix of the Metropolis update kernel
// (Philox4x32 = IMAD.WIDE + LOP3; acceptance = carry-compare + shift-in).  Prints achieved
// thread-instructions / clk / SM for each mix so that DESIGN.md's issue-rate roofline uses
// measured, not assumed, pipe widths.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench pipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 256
#define CH 8

enum { M_WIDE, M_HI, M_LO, M_LOP3, M_IADD3, M_CARRY, M_MIX11, M_MIX12, M_PHILOX, M_SETP_OR,
       M_SHF, M_WIDECMP, M_FULLMIX, M_COUNT };

static const char* NAMES[M_COUNT] = {
    "IMAD.WIDE.U32", "IMAD.HI.U32", "IMAD(lo)", "LOP3", "IADD3", "IADD3.cc+IADD3.X pair",
    "WIDE:LOP3 1:1", "WIDE:LOP3 1:2", "philox round (2 WIDE+2 LOP3)", "ISETP+@P LOP3 pair",
    "SHF funnel", "WIDE-compare+LOP3 pair", "philox round + 1 carry pair + 1 widecmp pair" };
// intended SASS instructions per inner step (per chain)
static const int INSTR[M_COUNT] = { 1, 1, 1, 1, 1, 2, 2, 3, 4, 2, 1, 2, 8 };

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cyc)
{
    uint32_t a[CH], b[CH], c[CH], d[CH];
    uint64_t w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        a[i] = seed + threadIdx.x * 7 + i; b[i] = seed ^ (i * 0x9E3779B9u); c[i] = i + seed; d[i] = seed * i;
        w[i] = ((uint64_t)a[i] << 32) | b[i];
    }
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t T = seed * 3u, K0 = seed * 5u, K1 = seed * 9u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == M_WIDE) {
                asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }"
                             : "+l"(w[i]) : "r"(M0));
            } else if (MODE == M_HI) {
                a[i] = __umulhi(a[i], M0) + b[i];
            } else if (MODE == M_LO) {
                a[i] = a[i] * M0 + b[i];
            } else if (MODE == M_LOP3) {
                a[i] = (a[i] ^ b[i]) | (c[i] & a[i]);
            } else if (MODE == M_IADD3) {
                a[i] = a[i] + b[i] + c[i];
            } else if (MODE == M_CARRY) {
                asm volatile("{ .reg .u32 t; sub.cc.u32 t, %2, %1; addc.u32 %0, %0, %0; }"
                             : "+r"(c[i]) : "r"(c[(i + 1) % CH]), "r"(T));
            } else if (MODE == M_MIX11) {
                asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }"
                             : "+l"(w[i]) : "r"(M0));
                a[i] = (a[i] ^ b[i]) | (c[i] & a[i]);
            } else if (MODE == M_MIX12) {
                asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }"
                             : "+l"(w[i]) : "r"(M0));
                a[i] = (a[i] ^ b[i]) | (c[i] & a[i]);
                d[i] = (d[i] ^ c[i]) | (b[i] & d[i]);
            } else if (MODE == M_PHILOX) {
                const uint64_t p0 = (uint64_t)a[i] * M0, p1 = (uint64_t)c[i] * M1;
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b[i] ^ K0, n2 = (uint32_t)(p0 >> 32) ^ d[i] ^ K1;
                a[i] = n0; b[i] = (uint32_t)p1; c[i] = n2; d[i] = (uint32_t)p0;
            } else if (MODE == M_SETP_OR) {
                if (c[(i + 1) % CH] > T) c[i] += c[i];
            } else if (MODE == M_SHF) {
                a[i] = __funnelshift_l(b[i], a[i], 1);
            } else if (MODE == M_WIDECMP) {
                // hi word of x*2^k + ((~T)<<k) has the carry [x > T] at bit k
                const uint64_t addend = (uint64_t)(~T) << 7;
                const uint64_t s = (uint64_t)c[(i + 1) % CH] * 128u + addend;
                c[i] = (c[i] << 1) | ((uint32_t)(s >> 32) & 128u);
            } else if (MODE == M_FULLMIX) {
                const uint64_t p0 = (uint64_t)a[i] * M0, p1 = (uint64_t)c[i] * M1;
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b[i] ^ K0, n2 = (uint32_t)(p0 >> 32) ^ d[i] ^ K1;
                a[i] = n0; b[i] = (uint32_t)p1; c[i] = n2; d[i] = (uint32_t)p0;
                uint32_t acc = (uint32_t)w[i], acc2 = (uint32_t)(w[i] >> 32);
                asm volatile("{ .reg .u32 t; sub.cc.u32 t, %2, %1; addc.u32 %0, %0, %0; }"
                             : "+r"(acc) : "r"(n0), "r"(T));
                const uint64_t s = (uint64_t)n2 * 128u + ((uint64_t)(~K0) << 7);
                acc2 |= (uint32_t)(s >> 32) & 128u;
                w[i] = ((uint64_t)acc2 << 32) | acc;
            }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) r ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(int sms, int ctas_per_sm, uint32_t* out, long long* cyc, long long* hcyc)
{
    const int grid = sms * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, 12345u, cyc);   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hcyc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)hcyc[i]; avg /= grid;
    const double thr_instr_per_cta = 256.0 * ITERS * CH * INSTR[MODE];
    const double per_clk_sm = thr_instr_per_cta * ctas_per_sm / avg;
    printf("%-48s ctas/SM %d  cyc/iter %7.2f  cyc/CTA %9.0f  thread-instr/clk/SM %7.1f  (warp-instr/clk/SMSP %5.3f)  %.3f ms -> %.2f GHz-equivalent\n",
           NAMES[MODE], ctas_per_sm, avg / ITERS, avg, per_clk_sm, per_clk_sm / 128.0, ms, avg / (ms * 1e6));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    uint32_t* out; long long* cyc; cudaMalloc(&out, sms * 8 * 256 * 4); cudaMalloc(&cyc, sms * 8 * 8);
    long long* hcyc = (long long*)malloc(sms * 8 * 8);
    for (int cps = 2; cps <= 8; cps *= 2) {
        run<M_WIDE>(sms, cps, out, cyc, hcyc);   run<M_HI>(sms, cps, out, cyc, hcyc);
        run<M_LO>(sms, cps, out, cyc, hcyc);     run<M_LOP3>(sms, cps, out, cyc, hcyc);
        run<M_IADD3>(sms, cps, out, cyc, hcyc);  run<M_CARRY>(sms, cps, out, cyc, hcyc);
        run<M_MIX11>(sms, cps, out, cyc, hcyc);  run<M_MIX12>(sms, cps, out, cyc, hcyc);
        run<M_PHILOX>(sms, cps, out, cyc, hcyc); run<M_SETP_OR>(sms, cps, out, cyc, hcyc);
        run<M_SHF>(sms, cps, out, cyc, hcyc);    run<M_WIDECMP>(sms, cps, out, cyc, hcyc);
        run<M_FULLMIX>(sms, cps, out, cyc, hcyc);
        printf("\n");
    }
    return 0;
}
