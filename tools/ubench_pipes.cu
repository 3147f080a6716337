// ubench_pipes.cu -- issue-rate microbenchmark of the integer / min-max instructions the ORB
// kernels lean on (B200, sm_100a).  Each variant runs ILP-8 independent chains per thread, 8 warps
// per SM sub-partition; prints warp-instructions per cycle per SM.  Build and run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

enum Op { V3 = 0, H2, F3, MIX_V3_H2, IMADOP, PRMTOP, LOP3OP, POPCOP, MIX_POPC_LOP3, V2, MIX_V3_IMAD, MIX_V3_F3, NOPS };
static const char* kNames[] = { "VIMNMX3.S16x2", "HMNMX2", "FMNMX3", "VIMNMX3+HMNMX2 (1:1)", "IMAD", "PRMT", "LOP3", "POPC",
                                "POPC+LOP3 (1:1)", "VIMNMX.S16x2 (2-in)", "VIMNMX3+IMAD (1:1)", "VIMNMX3+FMNMX3 (1:1)" };

template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned seed)
{
    unsigned a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed * (threadIdx.x + 1 + i); b[i] = seed ^ (i * 0x9e3779b9u + threadIdx.x); }
    const unsigned c = seed | 0x00010001u;
    #pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == V3) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
            if (OP == V2) a[i] = (it & 1) ? __vmaxs2(a[i], b[i]) : __vmins2(a[i], c);
            if (OP == H2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == F3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i]), "f"(*(float*)&c));
            if (OP == MIX_V3_H2) {
                if (i & 1) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
                else asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            }
            if (OP == MIX_V3_F3) {
                if (i & 1) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
                else asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i]), "f"(*(float*)&c));
            }
            if (OP == IMADOP) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == MIX_V3_IMAD) {
                if (i & 1) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            }
            if (OP == PRMTOP) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == LOP3OP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == POPCOP) { unsigned t; asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(a[i])); a[i] = t + b[i]; }
            if (OP == MIX_POPC_LOP3) {
                unsigned t, u;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(u) : "r"(a[i]), "r"(b[i]), "r"(c));
                asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(u));
                a[i] = t + b[i];
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(unsigned* d, int sms, double mhz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sms, 1024>>>(d, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<sms, 1024>>>(d, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops_per_thread = (double)ITERS * ILP * ((OP == POPCOP) ? 2 : (OP == MIX_POPC_LOP3) ? 3 : 1);
    const double warp_inst = ops_per_thread * 32;   // 32 warps per SM
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-26s %8.3f ms  %6.2f warp-inst/clk/SM (counting the listed ops%s)\n", kNames[OP], ms, warp_inst / cycles,
           OP == POPCOP ? " + 1 IADD each" : OP == MIX_POPC_LOP3 ? " + 1 IADD each" : "");
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (nominal; rates assume this clock)\n", p.name, p.multiProcessorCount, mhz);
    unsigned* d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 1024 * 4);
    run<V3>(d, p.multiProcessorCount, mhz);
    run<V2>(d, p.multiProcessorCount, mhz);
    run<H2>(d, p.multiProcessorCount, mhz);
    run<F3>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_H2>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_F3>(d, p.multiProcessorCount, mhz);
    run<IMADOP>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_IMAD>(d, p.multiProcessorCount, mhz);
    run<PRMTOP>(d, p.multiProcessorCount, mhz);
    run<LOP3OP>(d, p.multiProcessorCount, mhz);
    run<POPCOP>(d, p.multiProcessorCount, mhz);
    run<MIX_POPC_LOP3>(d, p.multiProcessorCount, mhz);
    cudaFree(d);
    return 0;
}
