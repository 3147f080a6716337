// orb_device.cuh -- device-side plumbing shared by the extraction kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "orb_plan.h"
#include "../../include/orb_b200.h"

// Per-launch view of one batch (all pointers device memory).
struct OrbBatch {
    const uint8_t* img0;      // level 0 = the caller's frames (or their staged copy)
    size_t img0_stride;       // bytes between frames
    int img0_pitch;           // bytes between rows
    uint8_t* pyr;             // [batch][plan.pyr_bytes]   levels 1..
    uint8_t* blur;            // [batch][plan.blur_bytes]  levels 0..
    uint32_t* cand;           // [batch][plan.cand_per_frame] packed FAST candidates
    uint16_t* node_of;        // [batch][plan.cand_per_frame] quadtree scratch: node of each candidate
    int* cand_count;          // [batch][ORB_MAX_LEVELS]
    uint32_t* lkp;            // [batch][plan.kp_per_frame] packed keypoints per level, list order
    int* lkp_count;           // [batch][ORB_MAX_LEVELS]
    orbx_kp* kps;             // [batch][cap]
    uint8_t* desc;            // [batch][cap][32]
    int* n_out;               // [batch]
    int cap;
    const OrbTap* taps;       // resize tables
};

__device__ __forceinline__ const uint8_t* orb_level_ptr(const OrbPlan& plan, const OrbBatch& io, int frame, int l, int* pitch)
{
    if (l == 0) { *pitch = io.img0_pitch; return io.img0 + (size_t)frame * io.img0_stride; }
    *pitch = plan.lv[l].pitch;
    return io.pyr + (size_t)frame * plan.pyr_bytes + plan.lv[l].img_off;
}

// 4 bytes at an arbitrary byte address through aligned 32-bit loads.  The caller guarantees
// that p+7 is readable whenever p is not 4-byte aligned.
__device__ __forceinline__ uint32_t orb_ld_u32_unaligned(const uint8_t* p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t s = (uint32_t)(a & 3);
    const uint32_t* q = (const uint32_t*)(a - s);
    const uint32_t lo = __ldg(q);
    if (s == 0) return lo;
    const uint32_t hi = __ldg(q + 1);
    return __funnelshift_r(lo, hi, s * 8);
}

__device__ __forceinline__ int orb_refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// Block-wide inclusive scan, in place, of a[0..n) in shared memory.  All NT threads call it;
// scratch must hold NT/32 ints.  Returns the total.  (n up to a few thousand.)
template <int NT>
__device__ int orb_block_scan_incl(int* a, int n, int* scratch)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ipt = (n + NT - 1) / NT;
    const int b = tid * ipt;
    int sum = 0;
    for (int i = 0; i < ipt; ++i) if (b + i < n) sum += a[b + i];
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { int s = scratch[w]; if (w < wid) woff += s; total += s; }
    int run = woff + inc - sum;
    for (int i = 0; i < ipt; ++i) if (b + i < n) { run += a[b + i]; a[b + i] = run; }
    __syncthreads();
    return total;
}
