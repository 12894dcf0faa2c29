// pipe_map2.cu -- does the operand kind of LOP3 (register / uniform register / constant bank)
// change its issue cost, alone and mixed with IMAD.WIDE?  Same harness as pipe_map.cu.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
struct Keys { uint32_t k[32]; };

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(uint32_t* out, const __grid_constant__ Keys K, uint32_t s, long long* cyc)
{
    uint32_t a[CH], b[CH], c[CH], d[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = s * (i + 3) + threadIdx.x; b[i] = s ^ (i * 0x9E3779B9u); c[i] = a[i] * 7; d[i] = b[i] * 5; }
    uint32_t kr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) kr[i] = K.k[i] + threadIdx.x;   // per-thread -> lives in R registers
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) a[i] = a[i] ^ b[i] ^ kr[(i + r) & 7];                 // LOP3 R,R,R
                if (MODE == 1) a[i] = a[i] ^ b[i] ^ K.k[(i + 4 * r) & 31];           // LOP3 R,R,UR/const
                if (MODE == 2 || MODE == 3) {                                        // philox round, keys in R / in UR
                    const uint64_t p0 = (uint64_t)a[i] * 0xD2511F53u, p1 = (uint64_t)c[i] * 0xCD9E8D57u;
                    const uint32_t k0 = (MODE == 2) ? kr[(i + r) & 7] : K.k[(i + 4 * r) & 31];
                    const uint32_t k1 = (MODE == 2) ? kr[(i + r + 1) & 7] : K.k[(i + 4 * r + 1) & 31];
                    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b[i] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ d[i] ^ k1;
                    a[i] = n0; b[i] = (uint32_t)p1; c[i] = n2; d[i] = (uint32_t)p0;
                }
                if (MODE == 4 || MODE == 5) {                                        // philox round + 4 compare/accumulate pairs
                    const uint64_t p0 = (uint64_t)a[i] * 0xD2511F53u, p1 = (uint64_t)c[i] * 0xCD9E8D57u;
                    const uint32_t k0 = K.k[(i + 4 * r) & 31], k1 = K.k[(i + 4 * r + 1) & 31];
                    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b[i] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ d[i] ^ k1;
                    a[i] = n0; b[i] = (uint32_t)p1; c[i] = n2; d[i] = (uint32_t)p0;
                    if (MODE == 4) {   // carry pairs (IADD3 + IMAD.X)
                        asm("{ .reg .u32 t; add.cc.u32 t, %1, %2; addc.u32 %0, %0, %0; }" : "+r"(kr[i & 3]) : "r"(n0), "r"(K.k[30]));
                        asm("{ .reg .u32 t; add.cc.u32 t, %1, %2; addc.u32 %0, %0, %0; }" : "+r"(kr[4 + (i & 3)]) : "r"(n2), "r"(K.k[31]));
                    } else {           // setp + predicated lop3
                        asm("{ .reg .pred p; setp.gt.u32 p, %1, %2; @p xor.b32 %0, %0, 0x40; }" : "+r"(kr[i & 3]) : "r"(n0), "r"(K.k[30]));
                        asm("{ .reg .pred p; setp.gt.u32 p, %1, %2; @p xor.b32 %0, %0, 0x80; }" : "+r"(kr[4 + (i & 3)]) : "r"(n2), "r"(K.k[31]));
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) x ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ kr[i];
    out[blockIdx.x * 256 + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, uint32_t* out, long long* cyc, long long* h, int grid)
{
    Keys K; for (int i = 0; i < 32; i++) K.k[i] = 0x9E3779B9u * (i + 1);
    k<MODE><<<grid, 256>>>(out, K, 12345u, cyc);
    k<MODE><<<grid, 256>>>(out, K, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)h[i]; avg /= grid;
    printf("%-28s cycles per inner step per warp with 8 warps/SMSP: %.2f  (=> per SMSP per step: %.2f)\n", name, avg / (ITERS * 4.0 * CH), avg / (ITERS * 4.0 * CH) / 8);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 4;
    uint32_t* out; long long* cyc; cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    long long* h = (long long*)malloc(grid * 8);
    run<0>("lop3 R,R,R", out, cyc, h, grid);
    run<1>("lop3 R,R,UR", out, cyc, h, grid);
    run<2>("philox round keys in R", out, cyc, h, grid);
    run<3>("philox round keys in UR", out, cyc, h, grid);
    run<4>("philox + 2 carry pairs", out, cyc, h, grid);
    run<5>("philox + 2 setp/@p xor", out, cyc, h, grid);
    return 0;
}
