// Micro-benchmark of the per-SMSP issue rate of the instruction classes the decoder's row body is made of
// (sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ubench scripts/pipe_ubench.cu
// Output: warp-instructions per clock per SM sub-partition for each class (4 SMSPs per SM, 8 warps per SMSP resident).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCHAIN 8
#define UNROLL 32
#define ITERS 512

template <int OP>
__device__ __forceinline__ void step(uint32_t& x, uint32_t a, uint32_t b)
{
    if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(a));                 // FADD
    if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));      // FFMA
    if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));      // IMAD
    if (OP == 3) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));      // IMAD.HI
    if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(a), "r"(b));  // LOP3
    if (OP == 5) asm volatile("min.f32 %0, %0, %1;" : "+r"(x) : "r"(a));                     // FMNMX
    if (OP == 6) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, t, %0;}" : "+r"(x) : "r"(a));   // VIADDMNMX
    if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(x) : "r"(a));           // SHF
    if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));                     // IADD3
    if (OP == 9) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(x) : "r"(a), "r"(b));   // FSETP+SEL
    if (OP == 10) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(a));                 // FMUL
    if (OP == 11) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));                              // shift by constant (IMAD.SHL or SHF?)
    if (OP == 12) asm volatile("min.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));        // FMNMX3
    if (OP == 14) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(x) : "r"(a), "r"(b));   // ISETP+SEL
    if (OP == 15) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+r"(x) : "r"(a));    // FFMA imm
    if (OP == 13) asm volatile("{.reg .u32 t; mad.lo.u32 t, %0, %1, %2; mad.hi.u32 %0, t, %1, %2;}" : "+r"(x) : "r"(a), "r"(b));   // IMAD + IMAD.HI pair
}

// MIX: alternate one ALU-pipe op (LOP3) with NF FMA-pipe ops (FADD) on independent chains
template <int OPA, int OPB, int NB>
__global__ void __launch_bounds__(256) mix_kernel(uint32_t* out, uint32_t a, uint32_t b, long long* cyc)
{
    uint32_t x[NCHAIN], y[NCHAIN];
    for (int i = 0; i < NCHAIN; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL / NCHAIN; u++) {
#pragma unroll
            for (int i = 0; i < NCHAIN; i++) {
                step<OPA>(x[i], a, b);
#pragma unroll
                for (int k = 0; k < NB; k++) step<OPB>(y[(i + k) % NCHAIN], a, b);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < NCHAIN; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed fp32x2 add / fma (FADD2 / FFMA2): two chains per instruction
template <int OP>
__global__ void __launch_bounds__(256) pk_kernel(uint32_t* out, uint32_t a, uint32_t b, long long* cyc)
{
    uint32_t x[NCHAIN], y[NCHAIN];
    for (int i = 0; i < NCHAIN; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * 3 + i; }
    __syncthreads();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL / NCHAIN; u++) {
#pragma unroll
            for (int i = 0; i < NCHAIN; i++) {
                if (OP == 0)
                    asm volatile("{.reg .b64 p, q; mov.b64 p, {%0, %1}; mov.b64 q, {%2, %3}; add.rn.f32x2 p, p, q; mov.b64 {%0, %1}, p;}"
                                 : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
                else
                    asm volatile("{.reg .b64 p, q; mov.b64 p, {%0, %1}; mov.b64 q, {%2, %3}; fma.rn.f32x2 p, p, q, q; mov.b64 {%0, %1}, p;}"
                                 : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
            }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < NCHAIN; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run_pk(const char* name, uint32_t* out, long long* cyc)
{
    const int blocks = 148 * 16, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pk_kernel<OP><<<blocks, threads>>>(out, 0x3f800001u, 0x00000003u, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pk_kernel<OP><<<blocks, threads>>>(out, 0x3f800001u, 0x00000003u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double instr = (double)blocks * (threads / 32) * ITERS * UNROLL;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f warp-instr/clk/SMSP  (%.3f ms)\n", name, instr / cycles / 592.0, ms);
}

template <int OPA, int OPB, int NB>
void run(const char* name, uint32_t* out, long long* cyc, int perIter)
{
    // whole-grid timing: total warp-instructions / (elapsed * SM clock * 592 SMSPs); independent of residency
    const int blocks = 148 * 16, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mix_kernel<OPA, OPB, NB><<<blocks, threads>>>(out, 0x3f800001u, 0x00000003u, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    mix_kernel<OPA, OPB, NB><<<blocks, threads>>>(out, 0x3f800001u, 0x00000003u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mix_kernel<OPA, OPB, NB>, threads, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double instr = (double)blocks * (threads / 32) * ITERS * UNROLL * perIter;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %.3f warp-instr/clk/SMSP  (%.3f ms, occ %d blocks/SM, clk %d kHz)\n", name, instr / cycles / 592.0, ms, occ, clk);
}

int main()
{
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 16 * 256 * 4);
    cudaMalloc(&cyc, 148 * 16 * 8);
    run<0, 0, 0>("FADD", out, cyc, 1);
    run<1, 0, 0>("FFMA", out, cyc, 1);
    run<10, 0, 0>("FMUL", out, cyc, 1);
    run<2, 0, 0>("IMAD", out, cyc, 1);
    run<3, 0, 0>("IMAD.HI", out, cyc, 1);
    run<13, 0, 0>("IMAD+IMAD.HI pair", out, cyc, 2);
    run<4, 0, 0>("LOP3", out, cyc, 1);
    run<5, 0, 0>("FMNMX", out, cyc, 1);
    run<12, 0, 0>("FMNMX3", out, cyc, 1);
    run<6, 0, 0>("VIADDMNMX", out, cyc, 1);
    run<7, 0, 0>("SHF", out, cyc, 1);
    run<9, 0, 0>("FSETP+SEL (2 instr)", out, cyc, 2);
    run<4, 0, 1>("LOP3 + 1 FADD", out, cyc, 2);
    run<4, 0, 2>("LOP3 + 2 FADD", out, cyc, 3);
    run<4, 2, 1>("LOP3 + 1 IMAD", out, cyc, 2);
    run<4, 2, 2>("LOP3 + 2 IMAD", out, cyc, 3);
    run<4, 3, 1>("LOP3 + 1 IMAD.HI", out, cyc, 2);
    run<0, 2, 1>("FADD + 1 IMAD", out, cyc, 2);
    run<0, 2, 2>("FADD + 2 IMAD", out, cyc, 3);
    run<5, 4, 1>("FMNMX + 1 LOP3", out, cyc, 2);
    run<7, 4, 1>("SHF + 1 LOP3", out, cyc, 2);
    run<15, 0, 0>("FFMA imm", out, cyc, 1);
    run<14, 0, 0>("ISETP+SEL (2 instr)", out, cyc, 2);
    run<5, 0, 1>("FMNMX + 1 FADD", out, cyc, 2);
    run<5, 0, 2>("FMNMX + 2 FADD", out, cyc, 3);
    run<12, 0, 1>("FMNMX3 + 1 FADD", out, cyc, 2);
    run<4, 1, 1>("LOP3 + 1 FFMA", out, cyc, 2);
    run<4, 1, 2>("LOP3 + 2 FFMA", out, cyc, 3);
    run_pk<0>("FADD2 (f32x2)", out, cyc);
    run_pk<1>("FFMA2 (f32x2)", out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
