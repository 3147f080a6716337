// ubench_pipes.cu -- issue-rate microbenchmark of the integer / min-max instructions the ORB
// kernels lean on (B200, sm_100a).  Each variant runs ILP-8 independent chains per thread, 8 warps
// per SM sub-partition; prints warp-instructions per cycle per SM.  Build and run on the GPU box:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

enum Op { V3 = 0, H2, F3, MIX_V3_H2, IMADOP, PRMTOP, LOP3OP, POPCOP, MIX_POPC_LOP3, V2, MIX_V3_IMAD, MIX_V3_F3, DP4A, DP2A, MIX_V3_DP4A, FFMAOP, MIX_V3_FFMA, LDS32, LDS64, LDS128, MIX_V3_LDS32, SHFL, MIX_V3_SHFL, VIADDMN, MIX_IMAD_DP4A, MIX_PRMT_IMAD, NOPS };
static const char* kNames[] = { "VIMNMX3.S16x2", "HMNMX2", "FMNMX3", "VIMNMX3+HMNMX2 (1:1)", "IMAD", "PRMT", "LOP3", "POPC",
                                "POPC+LOP3 (1:1)", "VIMNMX.S16x2 (2-in)", "VIMNMX3+IMAD (1:1)", "VIMNMX3+FMNMX3 (1:1)", "IDP.4A", "IDP.2A", "VIMNMX3+IDP.4A (1:1)", "FFMA", "VIMNMX3+FFMA (1:1)", "LDS.32", "LDS.64", "LDS.128", "VIMNMX3+LDS.32 (4:1)", "SHFL.IDX", "VIMNMX3+SHFL (4:1)", "VIADDMNMX.S16x2", "IMAD+IDP.4A (1:1)", "PRMT+IMAD (1:1)" };

template <int OP>
__global__ void __launch_bounds__(1024) k(unsigned* out, unsigned seed)
{
    __shared__ unsigned sm[1024];
    sm[threadIdx.x] = threadIdx.x * 4u;
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
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
            if (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == DP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            if (OP == MIX_V3_DP4A) {
                if (i & 1) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
                else asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            }
            if (OP == MIX_IMAD_DP4A) {
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            }
            if (OP == MIX_PRMT_IMAD) {
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
                else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            }
            if (OP == FFMAOP) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i]), "f"(*(float*)&c));
            if (OP == MIX_V3_FFMA) {
                if (i & 1) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i]), "f"(*(float*)&c));
            }
            if (OP == VIADDMN) a[i] = __viaddmax_s16x2(a[i], b[i], c);
            if (OP == SHFL) a[i] = __shfl_sync(0xffffffffu, a[i], (threadIdx.x + 1) & 31);
            if (OP == MIX_V3_SHFL) {
                if (i == 0) a[i] = __shfl_sync(0xffffffffu, a[i], (threadIdx.x + 1) & 31);
                else if (i < 5) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
            }
            if (OP == LDS32) { unsigned t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(sbase + ((a[i] & 0x3fcu)))); a[i] ^= t; }
            if (OP == LDS64) { unsigned t, u; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(t), "=r"(u) : "r"(sbase + ((a[i] & 0x3f8u)))); a[i] ^= t ^ u; }
            if (OP == LDS128) { unsigned t, u, v, w; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t), "=r"(u), "=r"(v), "=r"(w) : "r"(sbase + ((a[i] & 0x3f0u)))); a[i] ^= t ^ u ^ v ^ w; }
            if (OP == MIX_V3_LDS32) {
                if (i == 0) { unsigned t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(sbase + 4 * threadIdx.x % 1024)); b[1] ^= t; }
                else if (i < 5) a[i] = (it & 1) ? __vimax3_s16x2(a[i], b[i], c) : __vimin3_s16x2(a[i], b[i], c);
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
    const double ops_per_thread = (double)ITERS * ((OP == MIX_V3_SHFL || OP == MIX_V3_LDS32) ? 5 : ILP) * ((OP == POPCOP) ? 2 : (OP == MIX_POPC_LOP3) ? 3 : (OP == LDS32 || OP == LDS64 || OP == LDS128) ? 3 : 1);
    const double warp_inst = ops_per_thread * 32;   // 32 warps per SM
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-26s %8.3f ms  %6.2f warp-inst/clk/SM (counting the listed ops%s)\n", kNames[OP], ms, warp_inst / cycles,
           OP == POPCOP ? " + 1 IADD each" : OP == MIX_POPC_LOP3 ? " + 1 IADD each" : (OP == LDS32 || OP == LDS64 || OP == LDS128) ? " + LOP3 addr + XOR each: 3 inst per load" : "");
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
    run<DP4A>(d, p.multiProcessorCount, mhz);
    run<DP2A>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_DP4A>(d, p.multiProcessorCount, mhz);
    run<MIX_IMAD_DP4A>(d, p.multiProcessorCount, mhz);
    run<MIX_PRMT_IMAD>(d, p.multiProcessorCount, mhz);
    run<FFMAOP>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_FFMA>(d, p.multiProcessorCount, mhz);
    run<VIADDMN>(d, p.multiProcessorCount, mhz);
    run<SHFL>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_SHFL>(d, p.multiProcessorCount, mhz);
    run<LDS32>(d, p.multiProcessorCount, mhz);
    run<LDS64>(d, p.multiProcessorCount, mhz);
    run<LDS128>(d, p.multiProcessorCount, mhz);
    run<MIX_V3_LDS32>(d, p.multiProcessorCount, mhz);
    cudaFree(d);
    return 0;
}
