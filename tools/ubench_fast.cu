// ubench_fast.cu -- pipe-rate and inner-loop microbenchmarks behind the FAST kernel design (B200, sm_100a).
//
// Part 1: issue rate of the f16x2 min/max and add/fma instructions, alone and mixed (which pipe, which rate),
//         with the SASS mnemonic each PTX line becomes noted beside it (check with cuobjdump -sass).
// Part 2: the FAST score network (orb_fast.cu) fed from a shared-memory tile by conflict-free 32-bit loads,
//         in its pipe-placement variants, at several warps per SM sub-partition: cycles per pixel-pair row.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench_fast tools/ubench_fast.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

enum Op { H2 = 0, H3, HADD, HFMAR, MIX_H2_HADD, MIX_H3_HADD, MIX_H3_HFMA12, MIX_H2_PRMT, MIX_H2_H3, IADD3OP, MIX_H2_IADD3, MIX_HADD_IADD3, MIX_H2_LOP, NOPS };
static const char* kNames[] = { "HMNMX2 (2-in)", "VHMNMX (3-in)", "HADD2", "HFMA2.RELU", "HMNMX2+HADD2 1:1", "VHMNMX+HADD2 1:1", "VHMNMX+HFMA2 1:2",
                                "HMNMX2+PRMT 1:1", "HMNMX2+VHMNMX 1:1", "IADD3", "HMNMX2+IADD3 1:1", "HADD2+IADD3 1:1", "HMNMX2+LOP3 1:1" };

// min then max: ptxas fuses two chained max (or two chained min) into one 3-input VHMNMX, a min/max alternation stays 2-input
__device__ __forceinline__ void op_h2(unsigned& a, unsigned b, int it) { if (it & 1) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); else asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); }
__device__ __forceinline__ void op_h3(unsigned& a, unsigned b, unsigned c)
{
    asm volatile("{ .reg .b32 t; max.f16x2 t, %0, %1; max.f16x2 t, t, %2; min.f16x2 t, t, %1; min.f16x2 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));   // two VHMNMX
}
__device__ __forceinline__ void op_hadd(unsigned& a, unsigned b) { asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); }
__device__ __forceinline__ void op_hfmar(unsigned& a, unsigned b) { asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(0xbc00bc00u)); }
__device__ __forceinline__ void op_prmt(unsigned& a, unsigned b, unsigned c) { asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c)); }
__device__ __forceinline__ void op_iadd3(unsigned& a, unsigned b, unsigned c) { asm volatile("{ .reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c)); }
__device__ __forceinline__ void op_lop(unsigned& a, unsigned b, unsigned c) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c)); }

template <int OP>
__global__ void __launch_bounds__(1024) k_rate(unsigned* out, unsigned seed)
{
    unsigned a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = (seed * (threadIdx.x + 1 + i)) & 0x00ff00ffu; b[i] = (seed ^ (i * 0x9e3779b9u + threadIdx.x)) & 0x00ff00ffu; }
    const unsigned c = (seed | 0x00010001u) & 0x00ff00ffu;
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == H2) op_h2(a[i], b[i], it);
            if (OP == H3) op_h3(a[i], b[i], c);
            if (OP == HADD) op_hadd(a[i], b[i]);
            if (OP == HFMAR) op_hfmar(a[i], b[i]);
            if (OP == MIX_H2_HADD) { if (i & 1) op_h2(a[i], b[i], it); else op_hadd(a[i], b[i]); }
            if (OP == MIX_H3_HADD) { if (i & 1) op_h3(a[i], b[i], c); else op_hadd(a[i], b[i]); }
            if (OP == MIX_H3_HFMA12) { if (i % 3 == 0) op_h3(a[i], b[i], c); else op_hfmar(a[i], b[i]); }
            if (OP == MIX_H2_PRMT) { if (i & 1) op_h2(a[i], b[i], it); else op_prmt(a[i], b[i], c); }
            if (OP == MIX_H2_H3) { if (i & 1) op_h2(a[i], b[i], it); else op_h3(a[i], b[i], c); }
            if (OP == IADD3OP) op_iadd3(a[i], b[i], c);
            if (OP == MIX_H2_IADD3) { if (i & 1) op_h2(a[i], b[i], it); else op_iadd3(a[i], b[i], c); }
            if (OP == MIX_HADD_IADD3) { if (i & 1) op_hadd(a[i], b[i]); else op_iadd3(a[i], b[i], c); }
            if (OP == MIX_H2_LOP) { if (i & 1) op_h2(a[i], b[i], it); else op_lop(a[i], b[i], c); }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
static void run_rate(unsigned* d, int sms, double mhz, int threads = 1024)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rate<OP><<<sms, threads>>>(d, 12345u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_rate<OP><<<sms, threads>>>(d, 12345u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // SASS instructions per loop body (checked with cuobjdump): op_h3 is two VHMNMX; op_iadd3 one IADD3
    double per = ILP;
    if (OP == H3) per = 2 * ILP;
    if (OP == MIX_H3_HADD || OP == MIX_H2_H3) per = ILP / 2 * 3;
    if (OP == MIX_H3_HFMA12) per = 0; 
    for (int i = 0; OP == MIX_H3_HFMA12 && i < ILP; ++i) per += (i % 3 == 0) ? 2 : 1;
    const double warp_inst = (double)ITERS * per * (threads / 32);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-22s %2d warps/SM %8.3f ms  %6.2f warp-inst/clk/SM\n", kNames[OP], threads / 32, ms, warp_inst / cycles);
}

// ---------------------------------------------------------------- part 2: the score network from a tile
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ void hminmax_fma(uint32_t a, uint32_t b, uint32_t& mn, uint32_t& mx)
{
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0xbc00bc00u), "r"(a));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(mx) : "r"(b), "r"(r));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(mn) : "r"(a), "r"(r));
}
// single min / max on the FMA pipe: 2 instructions each
__device__ __forceinline__ uint32_t hmin_fma(uint32_t a, uint32_t b)
{
    uint32_t r, d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0xbc00bc00u), "r"(a));   // relu(a-b)
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(r));
    return d;
}
__device__ __forceinline__ uint32_t hmax_fma(uint32_t a, uint32_t b)
{
    uint32_t r, d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(0xbc00bc00u), "r"(a));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(b), "r"(r));
    return d;
}
template <bool F> __device__ __forceinline__ void hminmax(uint32_t a, uint32_t b, uint32_t& mn, uint32_t& mx)
{
    if (F) hminmax_fma(a, b, mn, mx); else { mn = hmin2(a, b); mx = hmax2(a, b); }
}
// MODE bit 0: first-stage pairs on the FMA pipe, bit 1: lo/hi pairs on the FMA pipe, bit 2: the Q stage as single FMA-pipe min/max
template <int MODE>
__device__ __forceinline__ void fast_network(const uint32_t* E, uint32_t& M1, uint32_t& M2)
{
    uint32_t Bn[8], Bx[8], Qn[8], Qx[8], Y[8], Z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hminmax<(MODE & 1) != 0>(E[2 * j], E[2 * j + 1], Bn[j], Bx[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (MODE & 4) { Qn[j] = hmin_fma(Bn[j], Bn[(j + 1) & 7]); Qx[j] = hmax_fma(Bx[j], Bx[(j + 1) & 7]); }
        else { Qn[j] = hmin2(Bn[j], Bn[(j + 1) & 7]); Qx[j] = hmax2(Bx[j], Bx[(j + 1) & 7]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t Fn = hmin2(Qn[j], Qn[(j + 2) & 7]);
        const uint32_t Fx = hmax2(Qx[j], Qx[(j + 2) & 7]);
        uint32_t ln, lx;
        hminmax<(MODE & 2) != 0>(E[(2 * j + 15) & 15], E[(2 * j + 8) & 15], ln, lx);
        Y[j] = hmin2(Fn, lx);
        Z[j] = hmax2(Fx, ln);
    }
    M1 = hmax2(hmax2(hmax2(Y[0], Y[1]), hmax2(Y[2], Y[3])), hmax2(hmax2(Y[4], Y[5]), hmax2(Y[6], Y[7])));
    M2 = hmin2(hmin2(hmin2(Z[0], Z[1]), hmin2(Z[2], Z[3])), hmin2(hmin2(Z[4], Z[5]), hmin2(Z[6], Z[7])));
}

#define RS 72       // words per tile row: copy A 36 words, copy B 36 words
#define TROWS 22    // 16 scored rows + 6
#define NCHUNK 64   // chunks (of 16 rows) each warp scores

// One warp = one band: lane p owns pixel pair p; tile rows in shared memory, 17 conflict-free loads per pair row.
// LDSMODE 0: 17 loads per row; 1: the dx=+-3 words of the three middle rows slide through registers (13 loads per row)
template <int MODE, int UNROLL, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_net(uint32_t* out, int nchunk, unsigned seed)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tile = smem + warp * (TROWS * RS);
    for (int i = lane; i < TROWS * RS; i += 32) {
        unsigned h = (i + 1) * 2654435761u ^ seed ^ (blockIdx.x * 7919u);
        tile[i] = (h >> 7 & 0xffu) | ((h >> 19 & 0xffu) << 16);
    }
    __syncwarp();
    const uint32_t* a0 = tile + lane;            // odd dx: words a[0..3] = dx -3, -1, +1, +3
    const uint32_t* b0 = tile + 36 + lane;       // even dx: words b[0..2] = dx -2, 0, +2
    uint32_t acc = 0;
    for (int ch = 0; ch < nchunk; ++ch) {
        const uint32_t* a = a0;
        const uint32_t* b = b0;
#pragma unroll (UNROLL)
        for (int ly = 0; ly < 16; ++ly, a += RS, b += RS) {
            uint32_t E[16];
            E[0] = b[6 * RS + 1];
            E[1] = a[6 * RS + 2];
            E[2] = b[5 * RS + 2];
            E[3] = a[4 * RS + 3];
            E[4] = a[3 * RS + 3];
            E[5] = a[2 * RS + 3];
            E[6] = b[1 * RS + 2];
            E[7] = a[0 * RS + 2];
            E[8] = b[0 * RS + 1];
            E[9] = a[0 * RS + 1];
            E[10] = b[1 * RS + 0];
            E[11] = a[2 * RS + 0];
            E[12] = a[3 * RS + 0];
            E[13] = a[4 * RS + 0];
            E[14] = b[5 * RS + 0];
            E[15] = a[6 * RS + 1];
            const uint32_t v = b[3 * RS + 1];
            uint32_t M1, M2;
            fast_network<MODE>(E, M1, M2);
            const uint32_t br = M1 + (0x01000100u - v), dk = (v + 0x01000100u) - M2;
            const uint32_t t = hmax2(br, dk);
            const uint32_t o = __viaddmax_s16x2_relu(t, 0xfef8fef8u, 0u);
            acc ^= o + ly;
        }
        if (acc == 0x12345u) tile[lane] ^= acc;   // keeps the loads inside the chunk loop
        __syncwarp();
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x)] = acc;
}

template <int MODE, int UNROLL, int WPB>
static void run_net(uint32_t* d, int sms, double mhz, int blocks_per_sm)
{
    const size_t smem = (size_t)WPB * TROWS * RS * 4;
    cudaFuncSetAttribute(k_net<MODE, UNROLL, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // pad the dynamic size so that exactly blocks_per_sm blocks fit
    size_t pad = (size_t)(220 * 1024) / blocks_per_sm - 1024;
    if (pad < smem) pad = smem;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * blocks_per_sm;
    k_net<MODE, UNROLL, WPB><<<grid, 32 * WPB, pad>>>(d, NCHUNK, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_net<MODE, UNROLL, WPB><<<grid, 32 * WPB, pad>>>(d, NCHUNK, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_net<MODE, UNROLL, WPB>, 32 * WPB, pad);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_net<MODE, UNROLL, WPB>);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    const double rows_per_smsp = (double)NCHUNK * 16 * blocks_per_sm * WPB / 4.0;   // warp pair-rows per sub-partition
    printf("net mode %d unroll %d: %2d warps/SM (occ %d blocks), %3d regs: %7.3f ms  %6.1f clk per warp pair-row per SMSP  (%s)\n",
           MODE, UNROLL, blocks_per_sm * WPB, occ, fa.numRegs, ms, cycles / rows_per_smsp, cudaGetErrorString(cudaGetLastError()));
}

// ---------------------------------------------------------------- part 3: scoring + fused NMS + survivor record
// The loop of part 2 with (a) the score finished on the FMA pipe in signed fp16 arithmetic, (b) the strict 3x3 maximum
// of the previous row taken from registers and two shuffles, (c) the survivor of a row pair written to a stash word.
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { uint32_t d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hfma2_relu(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) { uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hsetgt2(uint32_t a, uint32_t b) { uint32_t d; asm("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax3(uint32_t a, uint32_t b, uint32_t c) { return hmax2(hmax2(a, b), c); }

template <int MODE>
__device__ __forceinline__ uint32_t score_row(const uint32_t* a, const uint32_t* b, const uint32_t lm, const uint32_t bias)
{
    uint32_t E[16];
    E[0] = b[6 * RS + 1];
    E[1] = a[6 * RS + 2];
    E[2] = b[5 * RS + 2];
    E[3] = a[4 * RS + 3];
    E[4] = a[3 * RS + 3];
    E[5] = a[2 * RS + 3];
    E[6] = b[1 * RS + 2];
    E[7] = a[0 * RS + 2];
    E[8] = b[0 * RS + 1];
    E[9] = a[0 * RS + 1];
    E[10] = b[1 * RS + 0];
    E[11] = a[2 * RS + 0];
    E[12] = a[3 * RS + 0];
    E[13] = a[4 * RS + 0];
    E[14] = b[5 * RS + 0];
    E[15] = a[6 * RS + 1];
    const uint32_t v = b[3 * RS + 1];
    uint32_t M1, M2;
    fast_network<MODE>(E, M1, M2);
    const uint32_t t = hmax2(hsub2(M1, v), hsub2(v, M2));     // score + 1, signed
    return hfma2_relu(t, lm, bias);                           // max(score - minTh + 1, 0), 0 in masked lanes
}
// strict 3x3 maximum of row m between rows u and d; returns m where it survives, else 0 (per 16-bit lane)
__device__ __forceinline__ uint32_t nms_row(const uint32_t u, const uint32_t m, const uint32_t d, const uint32_t selL, const uint32_t selR)
{
    const uint32_t Cw = hmax3(u, m, d), Uw = hmax2(u, d);
    const uint32_t Lw = __shfl_up_sync(0xffffffffu, Cw, 1), Rw = __shfl_down_sync(0xffffffffu, Cw, 1);
    const uint32_t nb = hmax3(__byte_perm(Lw, Cw, selL), __byte_perm(Cw, Rw, selR), Uw);
    return hmul2(m, hsetgt2(m, nb));
}

template <int MODE, int WPB>
__global__ void __launch_bounds__(32 * WPB) k_net2(uint32_t* out, int nchunk, unsigned seed)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* tile = smem + warp * (TROWS * RS + 8 * 32);
    uint32_t* stash = tile + TROWS * RS;
    for (int i = lane; i < TROWS * RS; i += 32) {
        unsigned h = (i + 1) * 2654435761u ^ seed ^ (blockIdx.x * 7919u);
        tile[i] = (h >> 7 & 0xffu) | ((h >> 19 & 0xffu) << 16);
    }
    __syncwarp();
    const uint32_t* a0 = tile + lane;
    const uint32_t* b0 = tile + 36 + lane;
    const uint32_t selL = lane == 0 ? 0x5411u : 0x5432u, selR = lane == 31 ? 0x1132u : 0x5432u;
    const uint32_t lm = lane < 31 ? 0x3c003c00u : 0x00003c00u;
    const uint32_t bias = 0x80078007u;           // -7 * 2^-24 per lane
    uint32_t acc = 0;
    uint32_t u = 0, m = 0;
    for (int ch = 0; ch < nchunk; ++ch) {
        const uint32_t* a = a0;
        const uint32_t* b = b0;
#pragma unroll 1
        for (int ly = 0; ly < 16; ly += 2, a += 2 * RS, b += 2 * RS) {
            const uint32_t d0 = score_row<MODE>(a, b, lm, bias);
            const uint32_t d1 = score_row<MODE>(a + RS, b + RS, lm, bias);
            const uint32_t s0 = nms_row(u, m, d0, selL, selR);       // survivors of the row before d0
            const uint32_t s1 = nms_row(m, d0, d1, selL, selR);      // survivors of row d0
            stash[(ly >> 1) * 32 + lane] = hfma2(s1, 0x5c005c00u, s0);   // s0 + 256 * s1: at most one of the four lanes is non-zero
            u = d0; m = d1;
        }
        __syncwarp();
        // scan: the rare survivors
        for (int i = 0; i < 8; ++i) {
            const uint32_t w = stash[i * 32 + lane];
            if (w) acc += w + i;
        }
        if (acc == 0x12345u) tile[lane] ^= acc;
        __syncwarp();
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x)] = acc ^ u ^ m;
}

template <int MODE, int WPB>
static void run_net2(uint32_t* d, int sms, double mhz, int blocks_per_sm)
{
    const size_t smem = (size_t)WPB * (TROWS * RS + 8 * 32) * 4;
    cudaFuncSetAttribute(k_net2<MODE, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    size_t pad = (size_t)(220 * 1024) / blocks_per_sm - 1024;
    if (pad < smem) pad = smem;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * blocks_per_sm;
    k_net2<MODE, WPB><<<grid, 32 * WPB, pad>>>(d, NCHUNK, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_net2<MODE, WPB><<<grid, 32 * WPB, pad>>>(d, NCHUNK, 1u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_net2<MODE, WPB>);
    const double cycles = ms * 1e-3 * mhz * 1e6;
    const double rows_per_smsp = (double)NCHUNK * 16 * blocks_per_sm * WPB / 4.0;
    printf("net2 (score + NMS + record) mode %d: %2d warps/SM, %3d regs: %7.3f ms  %6.1f clk per warp pair-row per SMSP  (%s)\n",
           MODE, blocks_per_sm * WPB, fa.numRegs, ms, cycles / rows_per_smsp, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.0f MHz (nominal; rates assume this clock)\n", p.name, sms, mhz);
    unsigned* d;
    cudaMalloc(&d, (size_t)sms * 64 * 1024 * 4);
    run_rate<H2>(d, sms, mhz);
    run_rate<H3>(d, sms, mhz);
    run_rate<HADD>(d, sms, mhz);
    run_rate<HFMAR>(d, sms, mhz);
    run_rate<MIX_H2_HADD>(d, sms, mhz);
    run_rate<MIX_H3_HADD>(d, sms, mhz);
    run_rate<MIX_H3_HFMA12>(d, sms, mhz);
    run_rate<MIX_H2_PRMT>(d, sms, mhz);
    run_rate<MIX_H2_H3>(d, sms, mhz);
    run_rate<IADD3OP>(d, sms, mhz);
    run_rate<MIX_H2_IADD3>(d, sms, mhz);
    run_rate<MIX_HADD_IADD3>(d, sms, mhz);
    run_rate<MIX_H2_LOP>(d, sms, mhz);
    for (int t = 128; t <= 1024; t *= 2) {   // does a mixed stream issue faster with fewer warps per scheduler?
        run_rate<H2>(d, sms, mhz, t);
        run_rate<MIX_H2_HADD>(d, sms, mhz, t);
        run_rate<MIX_HADD_IADD3>(d, sms, mhz, t);
        run_rate<MIX_H3_HFMA12>(d, sms, mhz, t);
    }
    for (int bps = 4; bps <= 12; bps += 2) {   // 2-warp blocks: 8 .. 24 warps per SM
        run_net<0, 2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net<1, 2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net<2, 2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net<3, 2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net<7, 2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net<6, 2, 2>((uint32_t*)d, sms, mhz, bps);
    }
    for (int bps = 6; bps <= 12; bps += 2) {
        run_net2<0, 2>((uint32_t*)d, sms, mhz, bps);
        run_net2<1, 2>((uint32_t*)d, sms, mhz, bps);
        run_net2<2, 2>((uint32_t*)d, sms, mhz, bps);
        run_net2<3, 2>((uint32_t*)d, sms, mhz, bps);
    }
    run_net<2, 1, 2>((uint32_t*)d, sms, mhz, 8);
    run_net<3, 1, 2>((uint32_t*)d, sms, mhz, 8);
    run_net<2, 4, 2>((uint32_t*)d, sms, mhz, 8);
    run_net<3, 4, 2>((uint32_t*)d, sms, mhz, 8);
    cudaFree(d);
    return 0;
}
