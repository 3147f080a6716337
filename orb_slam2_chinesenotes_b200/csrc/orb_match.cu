// orb_match.cu -- Hamming matching kernels (sm_100a) and their C entry points.
//   k_hamming_bf   ORBmatcher::DescriptorDistance over all pairs, src/ORBmatcher.cc:46-63,
//                  with the best / second-best bookkeeping of :129-141 (strict <, first wins)
// 256-bit descriptors live in registers (query) and shared memory (train tile); the distance
// is 8 x (LOP3 xor + POPC).  Integer pipe only.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"

#define BF_NT 128     // threads per block = queries per block
#define BF_TILE 128   // train descriptors per shared-memory tile

__global__ void __launch_bounds__(BF_NT) k_hamming_bf(const uint4* __restrict__ q, const int nq,
                                                     const uint4* __restrict__ t, const int nt,
                                                     int* __restrict__ best_idx, int* __restrict__ best_dist,
                                                     int* __restrict__ second_dist)
{
    __shared__ uint4 s_t[BF_TILE * 2];
    const int prob = blockIdx.y;
    const int qi = blockIdx.x * BF_NT + threadIdx.x;
    const uint4* qp = q + (size_t)prob * nq * 2;
    const uint4* tp = t + (size_t)prob * nt * 2;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (qi < nq) { a0 = __ldg(qp + 2 * qi); a1 = __ldg(qp + 2 * qi + 1); }
    int bd = 256, bd2 = 256, bi = -1;   // initial values of src/ORBmatcher.cc:101-105
    for (int j0 = 0; j0 < nt; j0 += BF_TILE) {
        const int m = min(BF_TILE, nt - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * m; i += BF_NT) s_t[i] = __ldg(tp + 2 * j0 + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (d < bd) { bd2 = bd; bd = d; bi = j0 + j; }
            else if (d < bd2) bd2 = d;
        }
    }
    if (qi < nq) {
        const size_t o = (size_t)prob * nq + qi;
        best_idx[o] = bi; best_dist[o] = bd; second_dist[o] = bd2;
    }
}

namespace {
bool dev_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
struct Tmp {
    void* p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
};
} // namespace

extern "C" int orbm_hamming_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, int nprob,
                               int* best_idx, int* best_dist, int* second_dist, int device)
{
    if (!q || !t || !best_idx || !best_dist || !second_dist || nq <= 0 || nt < 0 || nprob <= 0) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    const size_t qb = (size_t)nprob * nq * 32, tb = (size_t)nprob * nt * 32, ob = (size_t)nprob * nq * 4;
    Tmp dq, dt, d0, d1, d2;
    const uint8_t* pq = q; const uint8_t* pt = t;
    int* o0 = best_idx; int* o1 = best_dist; int* o2 = second_dist;
#define CK(x) do { if ((x) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; } } while (0)
    if (!dev_ptr(q)) { CK(cudaMalloc(&dq.p, qb)); CK(cudaMemcpy(dq.p, q, qb, cudaMemcpyHostToDevice)); pq = (const uint8_t*)dq.p; }
    if (!dev_ptr(t)) { CK(cudaMalloc(&dt.p, tb + 32)); CK(cudaMemcpy(dt.p, t, tb, cudaMemcpyHostToDevice)); pt = (const uint8_t*)dt.p; }
    const bool h0 = !dev_ptr(best_idx), h1 = !dev_ptr(best_dist), h2 = !dev_ptr(second_dist);
    if (h0) { CK(cudaMalloc(&d0.p, ob)); o0 = (int*)d0.p; }
    if (h1) { CK(cudaMalloc(&d1.p, ob)); o1 = (int*)d1.p; }
    if (h2) { CK(cudaMalloc(&d2.p, ob)); o2 = (int*)d2.p; }
    if (((uintptr_t)pq | (uintptr_t)pt) & 15) return ORBX_E_ARG; // descriptors are read as 128-bit words
    k_hamming_bf<<<dim3((nq + BF_NT - 1) / BF_NT, nprob), BF_NT>>>((const uint4*)pq, nq, (const uint4*)pt, nt, o0, o1, o2);
    CK(cudaGetLastError());
    if (h0) CK(cudaMemcpy(best_idx, o0, ob, cudaMemcpyDeviceToHost));
    if (h1) CK(cudaMemcpy(best_dist, o1, ob, cudaMemcpyDeviceToHost));
    if (h2) CK(cudaMemcpy(second_dist, o2, ob, cudaMemcpyDeviceToHost));
    CK(cudaDeviceSynchronize());
#undef CK
    return ORBX_OK;
}
