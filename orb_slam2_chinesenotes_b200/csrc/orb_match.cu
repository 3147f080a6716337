// orb_match.cu -- Hamming matching kernels (sm_100a) and their C entry points.
//   k_hamming_bf        ORBmatcher::DescriptorDistance over all pairs, src/ORBmatcher.cc:46-63, with the
//                       best / second-best bookkeeping of :129-141 (strict <, first wins)
//   single-problem entry points with HOST arrays for every matcher: they stage the inputs in a per-thread
//                       arena and run ONE problem through the block-per-problem kernels of orb_match_batch.cu /
//                       orb_match_bow.cu / orb_mappoint.cu
//   k_grid_build, k_window_candidates, k_resolve_points / k_resolve_frame
//                       the candidate-list form of the window matchers (src/ORBmatcher.cc:73-157, :160-300 and
//                       the other best-candidate overloads): only for query sets too large for the shared memory
//                       of the block-per-problem kernel (more than ~23 000 queries)
//   (Frame::ComputeStereoMatches, src/Frame.cc:513-699: kernels in orb_stereo.cu, entry point here)
// 256-bit descriptors as 8 x u32.  Integer pipes only.
//
// The candidate-list form, why two phases: a query's candidate set and distances do not depend on other queries, but the
// reference skips candidates that an EARLIER query already claimed (:115-117, :234-236, :1094), so the
// choice of the best candidate must follow query order.  Phase 1 lists, per query and in the
// reference's candidate order (grid column, grid row, insertion), every candidate that passes the
// order-independent filters together with its distance; phase 2 walks the queries in order.  The
// reference's running best/second-best update equals "the two smallest by (distance, position)", so a
// warp reduces 32 candidates at a time.
#include <cuda_runtime.h>
#include <new>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/orb_b200.h"
#include "orb_launch.h"
#include "orb_stereo.h"

#define GRID_COLS 64    // FRAME_GRID_COLS, include/Frame.h:38
#define GRID_ROWS 48    // FRAME_GRID_ROWS, include/Frame.h:37
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define TH_HIGH 100     // src/ORBmatcher.cc:37
#define TH_LOW 50       // :38
#include "orb_match_common.cuh"   // HISTO_LENGTH, orb_three_maxima
#define MAX_KP 8192     // keypoints per frame the grid kernel sorts in shared memory

// ------------------------------------------------------------------------------ brute force
#ifndef ORB_BF_CSA
#define ORB_BF_CSA 2  // brute force: carry-save compression before POPC (0 = plain 8 POPC per pair; 1, 2, 3 see k_hamming_bf)
#endif
#define BF_NT 128     // threads per block = queries per block
#define BF_TILE 128   // train descriptors per shared-memory tile

#define BF_KEY_SHIFT 22   // packed (distance << 22 | train index) keys: train sets below 2^22 descriptors
template <bool PACKED>
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
    uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
    for (int j0 = 0; j0 < nt; j0 += BF_TILE) {
        const int m = min(BF_TILE, nt - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * m; i += BF_NT) s_t[i] = __ldg(tp + 2 * j0 + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
#if ORB_BF_CSA
            // POPC issues at a quarter of the LOP3 rate and bounds the plain form (8 per pair).  Carry-save adders
            // (sum = x^y^z, carry = majority, one LOP3 each) compress three words of equal weight into two, so POPC
            // work can be traded for LOP3 work until the two pipes are balanced:
            //   ORB_BF_CSA = 1: two CSAs           -> 6 POPC + 4 LOP3
            //   ORB_BF_CSA = 2: three CSAs         -> 5 POPC + 6 LOP3   (balanced on sm_100: POPC 0.5, LOP3 ~2 per clock and SM)
            //   ORB_BF_CSA = 4: four CSAs          -> 4 POPC + 8 LOP3
            //   ORB_BF_CSA = 3: full Harley-Seal   -> 4 POPC + 14 LOP3
            const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w,
                           x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
            const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
            const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
#if ORB_BF_CSA == 1
            const int d = (__popc(s1) + __popc(s2)) + (__popc(x6) + __popc(x7)) + 2 * (__popc(c1) + __popc(c2));
#elif ORB_BF_CSA == 2
            const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
            const int d = (__popc(s3) + __popc(x7)) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
#elif ORB_BF_CSA == 4
            const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
            const uint32_t t1 = c1 ^ c2 ^ c3, e1 = (c1 & c2) | (c3 & (c1 | c2));
            const int d = (__popc(s3) + __popc(x7)) + 2 * __popc(t1) + 4 * __popc(e1);
#else
            const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
            const uint32_t ones = s3 ^ x7, c4 = s3 & x7;
            const uint32_t t1 = c1 ^ c2 ^ c3, e1 = (c1 & c2) | (c3 & (c1 | c2));
            const uint32_t twos = t1 ^ c4, e2 = t1 & c4;
            const uint32_t fours = e1 ^ e2, eights = e1 & e2;
            const int d = __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
#endif
#else
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
#endif
            if (PACKED) {
                // the running best / second-best update of :129-141 (strict <, first wins) keeps the two smallest
                // (distance, index) pairs; as one key per candidate that is three min/max (the pack is an IMAD on the
                // FMA pipe)
                const uint32_t key = (uint32_t)d * (1u << BF_KEY_SHIFT) + (uint32_t)(j0 + j);
                k2 = min(k2, max(k1, key));
                k1 = min(k1, key);
            } else {
                if (d < bd) { bd2 = bd; bd = d; bi = j0 + j; }
                else if (d < bd2) bd2 = d;
            }
        }
    }
    if (PACKED) {
        // a candidate at distance 256 is never accepted (the initial bestDist is 256 and the test is strict)
        if ((k1 >> BF_KEY_SHIFT) < 256u) { bd = (int)(k1 >> BF_KEY_SHIFT); bi = (int)(k1 & ((1u << BF_KEY_SHIFT) - 1u)); }
        bd2 = min((int)(k2 >> BF_KEY_SHIFT), 256);
    }
    if (qi < nq) {
        const size_t o = (size_t)prob * nq + qi;
        best_idx[o] = bi; best_dist[o] = bd; second_dist[o] = bd2;
    }
}

// ------------------------------------------------------------------------------ grid
struct DevFrame {
    int n;
    const orbx_kp* kps;
    const uint32_t* desc;    // n x 8 words
    const float* u_right;    // or nullptr
    float min_x, min_y, inv_w, inv_h;
    int* cell_start;         // [GRID_CELLS + 1]
    uint16_t* items;         // keypoint indices grouped by cell, index order inside a cell
};

// One block: cell of every keypoint (round(), not floor: src/Frame.cc:414-415), keys (cell << 16 | index)
// sorted in shared memory, so a cell's members come out in insertion (index) order.
__global__ void __launch_bounds__(1024) k_grid_build(DevFrame F)
{
    extern __shared__ uint32_t keys[];
    int sn = 32; while (sn < F.n) sn <<= 1;
    for (int i = threadIdx.x; i < sn; i += blockDim.x) {
        uint32_t key = 0xffffffffu;
        if (i < F.n) {
            const int px = (int)roundf((F.kps[i].x - F.min_x) * F.inv_w);
            const int py = (int)roundf((F.kps[i].y - F.min_y) * F.inv_h);
            if (!(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS)) key = ((uint32_t)(px * GRID_ROWS + py) << 16) | (uint32_t)i;
        }
        keys[i] = key;
    }
    for (int i = threadIdx.x; i <= GRID_CELLS; i += blockDim.x) F.cell_start[i] = 0;
    __syncthreads();
    for (int k = 2; k <= sn; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < sn; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = keys[i], b = keys[ixj];
                    if (((i & k) == 0) ? (a > b) : (a < b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // cell_start[c] = first position whose cell >= c
    for (int i = threadIdx.x; i < sn; i += blockDim.x) {
        const uint32_t key = keys[i];
        const int c = key == 0xffffffffu ? GRID_CELLS : (int)(key >> 16);
        const int cprev = i == 0 ? -1 : (keys[i - 1] == 0xffffffffu ? GRID_CELLS : (int)(keys[i - 1] >> 16));
        for (int cc = cprev + 1; cc <= c; ++cc) F.cell_start[cc] = i;
        if (key != 0xffffffffu) F.items[i] = (uint16_t)(key & 0xffffu);
        if (i == sn - 1) for (int cc = c + 1; cc <= GRID_CELLS; ++cc) F.cell_start[cc] = sn;
    }
}

// ------------------------------------------------------------------------------ phase 1
struct WinQuery {
    float u, v, r;          // window centre and half-size (Frame::GetFeaturesInArea x, y, r)
    int min_level, max_level;
    float ur, er_max;       // right-image check: skip if uRight[idx] > 0 and |ur - uRight[idx]| > er_max
    int valid;              // 0: the reference `continue`s before the window query
};

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint4 b0, const uint4 b1)
{
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// One warp per query.  list[q*cap + k] = candidate index | distance << 16, in the reference's order.
__global__ void __launch_bounds__(256) k_window_candidates(const DevFrame F, const WinQuery* __restrict__ queries, const uint32_t* __restrict__ qdesc,
                                                          const int nq, uint32_t* __restrict__ list, int* __restrict__ count, const int cap)
{
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const WinQuery Q = queries[q];
    int n = 0;
    if (Q.valid) {
        // src/Frame.cc:355-372
        const int nMinCellX = max(0, (int)floorf((Q.u - F.min_x - Q.r) * F.inv_w));
        const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf((Q.u - F.min_x + Q.r) * F.inv_w));
        const int nMinCellY = max(0, (int)floorf((Q.v - F.min_y - Q.r) * F.inv_h));
        const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf((Q.v - F.min_y + Q.r) * F.inv_h));
        if (nMinCellX < GRID_COLS && nMaxCellX >= 0 && nMinCellY < GRID_ROWS && nMaxCellY >= 0) {
            const bool check_levels = Q.min_level > 0 || Q.max_level >= 0;       // :375
            uint32_t d[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = __ldg(qdesc + (size_t)q * 8 + i);
            uint32_t* out = list + (size_t)q * cap;
            for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
                // the cells (ix, iy0..iy1) are adjacent in the sorted item list
                const int s0 = F.cell_start[ix * GRID_ROWS + nMinCellY], s1 = F.cell_start[ix * GRID_ROWS + nMaxCellY + 1];
                for (int base = s0; base < s1; base += 32) {
                    const int j = base + lane;
                    bool ok = j < s1;
                    int idx = 0, dist = 0;
                    if (ok) {
                        idx = F.items[j];
                        const orbx_kp kp = F.kps[idx];
                        if (check_levels) {
                            if (kp.octave < Q.min_level) ok = false;
                            if (Q.max_level >= 0 && kp.octave > Q.max_level) ok = false;
                        }
                        if (!(fabsf(kp.x - Q.u) < Q.r && fabsf(kp.y - Q.v) < Q.r)) ok = false;   // :402
                        if (ok && F.u_right) {
                            const float ur = F.u_right[idx];
                            if (ur > 0 && fabsf(Q.ur - ur) > Q.er_max) ok = false;
                        }
                        if (ok) {
                            const uint4* p = (const uint4*)(F.desc + (size_t)idx * 8);
                            dist = hamming256(d, __ldg(p), __ldg(p + 1));
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (ok) out[n + __popc(m & ((1u << lane) - 1u))] = (uint32_t)idx | ((uint32_t)dist << 16);
                    n += __popc(m);
                }
            }
        }
    }
    if (lane == 0) count[q] = n;
}

// ------------------------------------------------------------------------------ phase 2 helpers
// Two smallest (distance, position) keys of one query's list among the candidates `usable` admits.
// key = dist << 20 | position (lists are shorter than 2^20); 0xffffffff = none.
template <typename Usable>
__device__ __forceinline__ void warp_top2(const uint32_t* lst, const int n, const int lane, Usable usable,
                                         uint32_t& k1, uint32_t& k2)
{
    k1 = 0xffffffffu; k2 = 0xffffffffu;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        uint32_t key = 0xffffffffu;
        if (j < n) {
            const uint32_t e = lst[j];
            if (usable((int)(e & 0xffffu), (int)(e >> 16))) key = ((e >> 16) << 20) | (uint32_t)j;
        }
        const uint32_t m1 = __reduce_min_sync(0xffffffffu, key);
        const uint32_t m2 = __reduce_min_sync(0xffffffffu, key == m1 ? 0xffffffffu : key);
        // merge (m1 <= m2) into (k1 <= k2); keys are unique or "none"
        if (m1 < k1) { k2 = min(k1, m2); k1 = m1; }
        else k2 = min(k2, m1);
    }
}

__global__ void k_debug_three_maxima(const int* __restrict__ sizes, const int n, int* __restrict__ ind)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a, b, c;
    orb_three_maxima(sizes + (size_t)i * HISTO_LENGTH, a, b, c);
    ind[3 * i] = a; ind[3 * i + 1] = b; ind[3 * i + 2] = c;
}

// rotation-histogram bin, src/ORBmatcher.cc:263-268 (factor = 1.0f/HISTO_LENGTH, as in the reference)
__device__ __forceinline__ int rot_bin(const float a1, const float a2)
{
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// ------------------------------------------------------------------------------ phase 2: map points -> frame
// src/ORBmatcher.cc:99-153.  assign[k]: map point attached to keypoint k (-1 none).  One warp.
__global__ void __launch_bounds__(32) k_resolve_points(const DevFrame F, const int nq, const WinQuery* __restrict__ queries,
                                                      const uint32_t* __restrict__ list, const int* __restrict__ count, const int cap,
                                                      const int* __restrict__ observations, int* __restrict__ assign, const float nnratio,
                                                      int* __restrict__ nmatches_out)
{
    const int lane = threadIdx.x;
    int nmatches = 0;
    for (int q = 0; q < nq; ++q) {
        const int n = queries[q].valid ? count[q] : 0;
        if (n == 0) continue;
        const uint32_t* lst = list + (size_t)q * cap;
        uint32_t k1, k2;
        warp_top2(lst, n, lane, [&](int idx, int) { const int a = assign[idx]; return !(a >= 0 && a < nq && observations[a] > 0); }, k1, k2);   // :115-117
        if (k1 == 0xffffffffu) continue;
        const int bestDist = (int)(k1 >> 20), bestIdx = (int)(lst[k1 & 0xfffffu] & 0xffffu);
        if (bestDist <= TH_HIGH) {
            const int bestLevel = F.kps[bestIdx].octave;
            int bestDist2 = 256, bestLevel2 = -1;
            if (k2 != 0xffffffffu) { bestDist2 = (int)(k2 >> 20); bestLevel2 = F.kps[lst[k2 & 0xfffffu] & 0xffffu].octave; }
            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)) continue;   // :146-149
            __syncwarp();
            if (lane == 0) assign[bestIdx] = q;
            __syncwarp();
            ++nmatches;
        }
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// ------------------------------------------------------------------------------ phase 2: last frame -> current frame
// src/ORBmatcher.cc:187-297.  obs[k]: Observations() of the point attached to current keypoint k (-1 free).
__global__ void __launch_bounds__(32) k_resolve_frame(const DevFrame F, const int nq, const WinQuery* __restrict__ queries,
                                                     const uint32_t* __restrict__ list, const int* __restrict__ count, const int cap,
                                                     const int* __restrict__ last_obs, const float* __restrict__ last_angle,
                                                     int* __restrict__ obs, int* __restrict__ assign, const int check_ori,
                                                     int* __restrict__ hist_entry /* [nq] kp index */, int* __restrict__ hist_bin /* [nq] */,
                                                     const int th_accept, int* __restrict__ nmatches_out)
{
    __shared__ int sizes[HISTO_LENGTH];
    const int lane = threadIdx.x;
    if (lane < HISTO_LENGTH) sizes[lane] = 0;
    __syncwarp();
    int nmatches = 0, nh = 0;
    for (int q = 0; q < nq; ++q) {
        const int n = queries[q].valid ? count[q] : 0;
        if (n == 0) continue;
        const uint32_t* lst = list + (size_t)q * cap;
        uint32_t k1, k2;
        warp_top2(lst, n, lane, [&](int idx, int) { return !(obs[idx] > 0); }, k1, k2);          // :234-236
        if (k1 == 0xffffffffu) continue;
        const int bestDist = (int)(k1 >> 20), bestIdx = (int)(lst[k1 & 0xfffffu] & 0xffffu);
        if (bestDist <= th_accept) {                                                              // :256 (TH_HIGH), :381 (ORBdist), :530 (TH_LOW)
            __syncwarp();
            if (lane == 0) {
                assign[bestIdx] = q; obs[bestIdx] = last_obs[q];
                if (check_ori) {
                    const int bin = rot_bin(last_angle[q], F.kps[bestIdx].angle);
                    hist_entry[nh] = bestIdx; hist_bin[nh] = bin; sizes[bin] += 1;
                }
            }
            __syncwarp();
            ++nmatches; ++nh;
        }
    }
    if (check_ori) {
        int i1, i2, i3;
        orb_three_maxima(sizes, i1, i2, i3);
        // every entry of a rejected bin clears its keypoint and decrements, duplicates included (:286-296)
        int dec = 0;
        for (int e = lane; e < nh; e += 32) {
            const int b = hist_bin[e];
            if (b != i1 && b != i2 && b != i3) { assign[hist_entry[e]] = -1; ++dec; }
        }
        for (int d = 16; d > 0; d >>= 1) dec += __shfl_xor_sync(0xffffffffu, dec, d);
        nmatches -= dec;
    }
    if (lane == 0) *nmatches_out = nmatches;
}

// ================================================================================ host side
namespace {
bool dev_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Per-call arena over per-THREAD cached memory: a single-problem entry point uploads a dozen small arrays, and a
// cudaMalloc + synchronous cudaMemcpy for each of them cost more than the kernels (0.33 ms per SearchByProjection
// call, 0.13 ms of it on the GPU).  Uploads are staged in a pinned twin of the device block at the same offset and go
// out as ONE asynchronous copy per flush(); pure scratch comes from a second, device-only block list.  The blocks
// persist between calls of the same thread (entry points are re-entrant across threads: thread_local) and are
// released when the thread ends.  Every entry point finishes with a synchronous copy back or a device synchronise, so
// the staging memory is free again when the next call starts.
// With ORB_B200_DEBUG set in the environment a failing CUDA call is reported on stderr (the matcher entry points have
// no context object to keep an error string in).
static void report_cuda(cudaError_t e, const char* what, int line)
{
    static const bool on = getenv("ORB_B200_DEBUG") != nullptr;
    if (on) fprintf(stderr, "orb_b200: %s (orb_match.cu:%d): %s\n", what, line, cudaGetErrorString(e));
}
struct ArenaBlock { char* dev = nullptr; char* pin = nullptr; size_t cap = 0, used = 0, flushed = 0; };
struct ThreadArena {
    int device = -1;
    bool inflight = false;              // flush() queued host-to-device copies that no synchronisation has covered yet
    cudaEvent_t flushed_ev = nullptr;   // recorded behind the last of them
    std::vector<ArenaBlock> up, scratch;
    void release()
    {
        for (ArenaBlock& b : up) { cudaFree(b.dev); cudaFreeHost(b.pin); }
        for (ArenaBlock& b : scratch) cudaFree(b.dev);
        up.clear(); scratch.clear();
        cudaGetLastError();
    }
    ~ThreadArena() { release(); }
};
thread_local ThreadArena t_arena;

struct Scratch {
    bool ok = true;
    Scratch()
    {
        int dev = -1;
        if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); ok = false; return; }
        if (dev != t_arena.device) { t_arena.release(); t_arena.device = dev; }
        // an entry point that failed after flush() may have left copies from the pinned twins in flight: they must land
        // before the bytes are reused (every successful call ends synchronised, the wait is then free)
        if (t_arena.inflight) { if (cudaEventSynchronize(t_arena.flushed_ev) != cudaSuccess) cudaGetLastError(); t_arena.inflight = false; }
        for (ArenaBlock& b : t_arena.up) b.used = b.flushed = 0;
        for (ArenaBlock& b : t_arena.scratch) b.used = 0;
    }
    ArenaBlock* take(std::vector<ArenaBlock>& blocks, size_t bytes, bool pinned, size_t* off)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        for (ArenaBlock& b : blocks)
            if (b.cap - b.used >= bytes) { *off = b.used; b.used += bytes; return &b; }
        ArenaBlock nb;
        nb.cap = bytes > ((size_t)4 << 20) ? bytes : ((size_t)4 << 20);
        { const cudaError_t e_ = cudaMalloc((void**)&nb.dev, nb.cap); if (e_ != cudaSuccess) { report_cuda(e_, "cudaMalloc", __LINE__); cudaGetLastError(); ok = false; return nullptr; } }
        if (pinned) { const cudaError_t e_ = cudaHostAlloc((void**)&nb.pin, nb.cap, cudaHostAllocDefault); if (e_ != cudaSuccess) { report_cuda(e_, "cudaHostAlloc", __LINE__); cudaGetLastError(); cudaFree(nb.dev); ok = false; return nullptr; } }
        nb.used = bytes;
        blocks.push_back(nb);
        *off = 0;
        return &blocks.back();
    }
    void* alloc(size_t bytes)          // device-only scratch
    {
        size_t off = 0;
        ArenaBlock* b = take(t_arena.scratch, bytes ? bytes : 4, false, &off);
        return b ? b->dev + off : nullptr;
    }
    template <typename T> T* up(const T* h, size_t count)      // staged for the next flush(); h may be null (no data)
    {
        size_t off = 0;
        ArenaBlock* b = take(t_arena.up, sizeof(T) * (count ? count : 1), true, &off);
        if (!b) return nullptr;
        if (h && count) memcpy(b->pin + off, h, sizeof(T) * count);
        return (T*)(b->dev + off);
    }
    void* pinned_of(const void* d)      // the staging twin of a pointer returned by up()
    {
        for (ArenaBlock& b : t_arena.up)
            if ((const char*)d >= b.dev && (const char*)d < b.dev + b.cap) return b.pin + ((const char*)d - b.dev);
        return nullptr;
    }
    bool flush()                        // everything staged so far -> device, one copy per block; before any launch
    {
        for (ArenaBlock& b : t_arena.up)
            if (b.used > b.flushed) {
                { const cudaError_t e_ = cudaMemcpyAsync(b.dev + b.flushed, b.pin + b.flushed, b.used - b.flushed, cudaMemcpyHostToDevice, 0); if (e_ != cudaSuccess) { report_cuda(e_, "flush", __LINE__); cudaGetLastError(); ok = false; return false; } }
                b.flushed = b.used;
                if (!t_arena.flushed_ev && cudaEventCreateWithFlags(&t_arena.flushed_ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); t_arena.flushed_ev = nullptr; }
                if (t_arena.flushed_ev && cudaEventRecord(t_arena.flushed_ev, 0) == cudaSuccess) t_arena.inflight = true;
            }
        return ok;
    }
};

bool make_frame(Scratch& S, const orbm_frame* F, DevFrame* D)
{
    if (!F || F->n < 0 || F->n > MAX_KP || (F->n > 0 && (!F->kps || !F->desc))) return false;
    D->n = F->n;
    // a frame that is already on the device (orbm_frame_upload / orbm_frame_view) is used where it lies
    D->kps = dev_ptr(F->kps) ? F->kps : S.up(F->kps, (size_t)F->n);
    D->desc = (const uint32_t*)(dev_ptr(F->desc) ? F->desc : S.up(F->desc, (size_t)F->n * 32));
    D->u_right = F->u_right ? (dev_ptr(F->u_right) ? F->u_right : S.up(F->u_right, (size_t)F->n)) : nullptr;
    D->min_x = F->min_x; D->min_y = F->min_y;
    D->inv_w = (float)GRID_COLS / (F->max_x - F->min_x);     // src/Frame.cc:108
    D->inv_h = (float)GRID_ROWS / (F->max_y - F->min_y);     // :109
    D->cell_start = (int*)S.alloc(sizeof(int) * (GRID_CELLS + 1));
    D->items = (uint16_t*)S.alloc(sizeof(uint16_t) * (size_t)(F->n + 32));
    if (!S.ok) return false;
    int sn = 32; while (sn < F->n) sn <<= 1;
    if (!S.flush()) return false;
    k_grid_build<<<1, 1024, (size_t)sn * 4>>>(*D);
    return cudaGetLastError() == cudaSuccess;
}

#define CKM(x) do { const cudaError_t e_ = (x); if (e_ != cudaSuccess) { report_cuda(e_, #x, __LINE__); cudaGetLastError(); return ORBX_E_CUDA; } } while (0)

int run_candidates(Scratch& S, const DevFrame& D, const std::vector<WinQuery>& hq, const uint8_t* qdesc_host, int nq,
                   WinQuery** dq, uint32_t** list, int** count, int* cap)
{
    *cap = D.n > 0 ? D.n : 1;
    *dq = S.up(hq.data(), (size_t)nq);
    const uint32_t* dqd = (const uint32_t*)S.up(qdesc_host, (size_t)nq * 32);
    *list = (uint32_t*)S.alloc(sizeof(uint32_t) * (size_t)nq * (size_t)*cap);
    *count = (int*)S.alloc(sizeof(int) * (size_t)nq);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    k_window_candidates<<<(nq + 7) / 8, 256>>>(D, *dq, dqd, nq, *list, *count, *cap);
    CKM(cudaGetLastError());
    return ORBX_OK;
}
// phase 2 of every "best candidate only" window matcher (frame-to-frame, relocalisation, loop closing)
int resolve_best(Scratch& S, const DevFrame& D, int nq, const WinQuery* dq, const uint32_t* list, const int* count, int cap,
                 const int* h_qobs, const float* h_angle, const int* h_obs, int* assign_out, int check_ori, int th_accept, int* nmatches)
{
    int* d_qobs = S.up(h_qobs, (size_t)nq);
    float* d_angle = S.up(h_angle, (size_t)nq);
    int* d_obs = S.up(h_obs, (size_t)D.n);
    int* d_assign = S.up(assign_out, (size_t)D.n);
    int* d_he = (int*)S.alloc(sizeof(int) * (size_t)nq);
    int* d_hb = (int*)S.alloc(sizeof(int) * (size_t)nq);
    int* d_nm = (int*)S.alloc(4);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    k_resolve_frame<<<1, 32>>>(D, nq, dq, list, count, cap, d_qobs, d_angle, d_obs, d_assign, check_ori, d_he, d_hb, th_accept, d_nm);
    CKM(cudaGetLastError());
    CKM(cudaMemcpy(assign_out, d_assign, sizeof(int) * (size_t)D.n, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(nmatches, d_nm, 4, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}
} // namespace

// assign_out [n] and the match count come back in ONE copy (d_assign has n + 1 entries, the count last): a synchronous
// device-to-host copy costs ~10 us whatever its size
static int fetch_assign(const int* d_assign, int n, int* assign_out, int* nmatches)
{
    thread_local std::vector<int> tmp;
    tmp.resize((size_t)n + 1);
    const cudaError_t e = cudaMemcpy(tmp.data(), d_assign, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { report_cuda(e, "fetch_assign", __LINE__); cudaGetLastError(); return ORBX_E_CUDA; }
    memcpy(assign_out, tmp.data(), sizeof(int) * (size_t)n);
    *nmatches = tmp[(size_t)n];
    return ORBX_OK;
}

bool orb_match_batch_fits(int kp_stride, int nq_stride);   // orb_match_batch.cu

namespace {
// One problem through the block-per-problem kernels of orb_match_batch.cu (the low-latency path: one launch,
// queries resolved in parallel).  Frame already uploaded in D (kps / desc / u_right).
struct OneFrame {
    orbm_frames F;
    int* d_n;
};
bool one_frame(Scratch& S, const orbm_frame* F, OneFrame* O)
{
    if (!F || F->n <= 0 || F->n > MAX_KP || !F->kps || !F->desc) return false;
    O->F.nprob = 1;
    // a frame that is already on the device (orbm_frame_upload / orbm_frame_view) is used where it lies
    O->F.kps = dev_ptr(F->kps) ? F->kps : S.up(F->kps, (size_t)F->n);
    O->F.desc = dev_ptr(F->desc) ? F->desc : S.up(F->desc, (size_t)F->n * 32);
    O->F.u_right = F->u_right ? (dev_ptr(F->u_right) ? F->u_right : S.up(F->u_right, (size_t)F->n)) : nullptr;
    O->d_n = S.up(&F->n, 1);
    O->F.n = O->d_n;
    O->F.kp_stride = F->n;
    O->F.min_x = F->min_x; O->F.max_x = F->max_x; O->F.min_y = F->min_y; O->F.max_y = F->max_y;
    O->F.max_n = 0;
    return S.ok;
}

// best-candidate-only search of already projected queries (hq) through orbm_window_search_best_batch
int best_via_batch(Scratch& S, const orbm_frame* F, bool use_ur, const std::vector<WinQuery>& hq, const uint8_t* qdesc, const float* h_angle,
                   const int* h_qobs, const int* h_init_obs, int* assign_out, int check_ori, int th_accept, int* nmatches)
{
    const int nq = (int)hq.size(), n = F->n;
    OneFrame O;
    if (!one_frame(S, F, &O)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    std::vector<float> uvr((size_t)nq * 3), ur((size_t)nq), er((size_t)nq);
    std::vector<int> minl((size_t)nq), maxl((size_t)nq);
    std::vector<uint8_t> valid((size_t)nq);
    for (int i = 0; i < nq; ++i) {
        const WinQuery& Q = hq[(size_t)i];
        valid[(size_t)i] = (uint8_t)(Q.valid != 0);
        uvr[3 * (size_t)i] = Q.valid ? Q.u : 0.f; uvr[3 * (size_t)i + 1] = Q.valid ? Q.v : 0.f; uvr[3 * (size_t)i + 2] = Q.valid ? Q.r : 0.f;
        minl[(size_t)i] = Q.valid ? Q.min_level : 0; maxl[(size_t)i] = Q.valid ? Q.max_level : 0;
        ur[(size_t)i] = Q.valid ? Q.ur : 0.f; er[(size_t)i] = Q.valid ? Q.er_max : 0.f;
    }
    orbm_windows W;
    int* d_nq = S.up(&nq, 1);
    W.nq = d_nq; W.nq_stride = nq;
    W.uvr = S.up(uvr.data(), uvr.size()); W.min_level = S.up(minl.data(), minl.size()); W.max_level = S.up(maxl.data(), maxl.size());
    W.ur = use_ur ? S.up(ur.data(), ur.size()) : nullptr; W.er_max = use_ur ? S.up(er.data(), er.size()) : nullptr;
    W.valid = S.up(valid.data(), valid.size());
    W.qdesc = S.up(qdesc, (size_t)nq * 32);
    W.q_angle = h_angle ? S.up(h_angle, (size_t)nq) : nullptr;
    W.q_obs = h_qobs ? S.up(h_qobs, (size_t)nq) : nullptr;
    int* d_init = h_init_obs ? S.up(h_init_obs, (size_t)n) : nullptr;
    int* d_assign = (int*)S.alloc(sizeof(int) * ((size_t)n + 1));
    int* d_nm = d_assign ? d_assign + n : nullptr;
    if (!S.ok) return ORBX_E_CUDA;
    if (!use_ur) O.F.u_right = nullptr;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_window_search_best_batch(&O.F, &W, d_init, d_assign, th_accept, check_ori, d_nm, nullptr, nullptr);
    if (rc) return rc;
    return fetch_assign(d_assign, n, assign_out, nmatches);
}
} // namespace

extern "C" {

// ---- frame handles: a Frame's keypoints / descriptors / mvuRight uploaded once, reused by every search of that frame
struct orbm_frame_handle {
    int device, n;
    char* block;             // one allocation: kps | desc | u_right
    const orbx_kp* kps; const uint8_t* desc; const float* u_right;
    float min_x, max_x, min_y, max_y;
};

int orbm_frame_upload(const orbm_frame* F, int device, orbm_frame_handle** handle)
{
    if (!handle) return ORBX_E_ARG;
    *handle = nullptr;
    if (!F || F->n < 0 || F->n > MAX_KP || (F->n > 0 && (!F->kps || !F->desc))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    orbm_frame_handle* H = new (std::nothrow) orbm_frame_handle();
    if (!H) return ORBX_E_CUDA;
    H->device = device; H->n = F->n; H->block = nullptr; H->kps = nullptr; H->desc = nullptr; H->u_right = nullptr;
    H->min_x = F->min_x; H->max_x = F->max_x; H->min_y = F->min_y; H->max_y = F->max_y;
    if (F->n > 0) {
        const size_t n = (size_t)F->n;
        const size_t o_desc = (n * sizeof(orbx_kp) + 255) & ~(size_t)255, o_ur = o_desc + ((n * 32 + 255) & ~(size_t)255);
        const size_t bytes = o_ur + (F->u_right ? n * 4 : 0);
        if (cudaMalloc((void**)&H->block, bytes) != cudaSuccess) { cudaGetLastError(); delete H; return ORBX_E_CUDA; }
        // pageable sources: cudaMemcpy returns once the bytes are staged, so the caller's vectors may go away at once
        bool ok = cudaMemcpy(H->block, F->kps, n * sizeof(orbx_kp), cudaMemcpyDefault) == cudaSuccess &&
                  cudaMemcpy(H->block + o_desc, F->desc, n * 32, cudaMemcpyDefault) == cudaSuccess &&
                  (!F->u_right || cudaMemcpy(H->block + o_ur, F->u_right, n * 4, cudaMemcpyDefault) == cudaSuccess) &&
                  cudaStreamSynchronize(0) == cudaSuccess;   // a copy from pageable memory may still be in flight when cudaMemcpy returns; the view may be used on any stream
        if (!ok) { cudaGetLastError(); cudaFree(H->block); delete H; return ORBX_E_CUDA; }
        H->kps = (const orbx_kp*)H->block; H->desc = (const uint8_t*)(H->block + o_desc);
        H->u_right = F->u_right ? (const float*)(H->block + o_ur) : nullptr;
    }
    *handle = H;
    return ORBX_OK;
}

int orbm_frame_view(const orbm_frame_handle* H, orbm_frame* view)
{
    if (!H || !view) return ORBX_E_ARG;
    view->n = H->n; view->kps = H->kps; view->desc = H->desc; view->u_right = H->u_right;
    view->min_x = H->min_x; view->max_x = H->max_x; view->min_y = H->min_y; view->max_y = H->max_y;
    return ORBX_OK;
}

int orbm_frame_device(const orbm_frame_handle* H) { return H ? H->device : -1; }

void orbm_frame_release(orbm_frame_handle* H)
{
    if (!H) return;
    if (H->block) {
        int prev = -1;
        const bool sw = cudaGetDevice(&prev) == cudaSuccess && prev != H->device && cudaSetDevice(H->device) == cudaSuccess;
        cudaFree(H->block);
        if (sw) cudaSetDevice(prev);
        cudaGetLastError();
    }
    delete H;
}

int orbm_hamming_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, int nprob,
                    int* best_idx, int* best_dist, int* second_dist, int device)
{
    if (!q || !t || !best_idx || !best_dist || !second_dist || nq <= 0 || nt < 0 || nprob <= 0) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    const size_t qb = (size_t)nprob * nq * 32, tb = (size_t)nprob * nt * 32, ob = (size_t)nprob * nq * 4;
    Scratch S;
    const uint8_t* pq = q; const uint8_t* pt = t;
    int* o0 = best_idx; int* o1 = best_dist; int* o2 = second_dist;
    if (!dev_ptr(q)) pq = S.up(q, qb);
    if (!dev_ptr(t)) { uint8_t* d = S.up((const uint8_t*)nullptr, tb + 32); if (d && tb) memcpy((void*)S.pinned_of(d), t, tb); pt = d; }
    const bool h0 = !dev_ptr(best_idx), h1 = !dev_ptr(best_dist), h2 = !dev_ptr(second_dist);
    if (h0) o0 = (int*)S.alloc(ob);
    if (h1) o1 = (int*)S.alloc(ob);
    if (h2) o2 = (int*)S.alloc(ob);
    if (!S.ok) return ORBX_E_CUDA;
    if (((uintptr_t)pq | (uintptr_t)pt) & 15) return ORBX_E_ARG; // descriptors are read as 128-bit words
    if (!S.flush()) return ORBX_E_CUDA;
    if (nt < (1 << BF_KEY_SHIFT))
        k_hamming_bf<true><<<dim3((nq + BF_NT - 1) / BF_NT, nprob), BF_NT>>>((const uint4*)pq, nq, (const uint4*)pt, nt, o0, o1, o2);
    else
        k_hamming_bf<false><<<dim3((nq + BF_NT - 1) / BF_NT, nprob), BF_NT>>>((const uint4*)pq, nq, (const uint4*)pt, nt, o0, o1, o2);
    CKM(cudaGetLastError());
    if (h0) CKM(cudaMemcpy(best_idx, o0, ob, cudaMemcpyDeviceToHost));
    if (h1) CKM(cudaMemcpy(best_dist, o1, ob, cudaMemcpyDeviceToHost));
    if (h2) CKM(cudaMemcpy(second_dist, o2, ob, cudaMemcpyDeviceToHost));
    CKM(cudaDeviceSynchronize());
    return ORBX_OK;
}

int orbm_hamming_bf_async(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int nprob,
                          int* d_best_idx, int* d_best_dist, int* d_second_dist, void* cuda_stream)
{
    if (!d_q || !d_t || !d_best_idx || !d_best_dist || !d_second_dist || nq <= 0 || nt < 0 || nprob <= 0) return ORBX_E_ARG;
    if (!dev_ptr(d_q) || !dev_ptr(d_t) || !dev_ptr(d_best_idx) || !dev_ptr(d_best_dist) || !dev_ptr(d_second_dist)) return ORBX_E_ARG;
    if (((uintptr_t)d_q | (uintptr_t)d_t) & 15) return ORBX_E_ARG;
    cudaPointerAttributes a;
    int prev = -1;
    if (cudaPointerGetAttributes(&a, d_q) != cudaSuccess || cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(a.device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    const dim3 grid((nq + BF_NT - 1) / BF_NT, nprob);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (nt < (1 << BF_KEY_SHIFT)) k_hamming_bf<true><<<grid, BF_NT, 0, st>>>((const uint4*)d_q, nq, (const uint4*)d_t, nt, d_best_idx, d_best_dist, d_second_dist);
    else k_hamming_bf<false><<<grid, BF_NT, 0, st>>>((const uint4*)d_q, nq, (const uint4*)d_t, nt, d_best_idx, d_best_dist, d_second_dist);
    const cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    return e == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
}

int orbm_search_by_projection_points(const orbm_frame* F, const float* scale, int nlevels,
                                     int nq, const float* proj_xyxr, const int* level, const float* view_cos,
                                     const uint8_t* in_view, const uint8_t* bad, const int* observations,
                                     const uint8_t* qdesc, const int* init_assign, int* assign_out,
                                     float th, float nnratio, int* nmatches, int device)
{
    if (!F || !scale || nq < 0 || (F->n > 0 && !assign_out) || !nmatches || (nq > 0 && (!proj_xyxr || !level || !view_cos || !in_view || !bad || !observations || !qdesc)))
        return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    for (int k = 0; k < F->n; ++k) assign_out[k] = init_assign ? init_assign[k] : -1;
    if (nq == 0 || F->n == 0) return ORBX_OK;
    Scratch S;
    if (nlevels <= 32 && orb_match_batch_fits(F->n, nq)) {
        // one launch of the block-per-problem kernel (orb_match_batch.cu): queries are built and resolved on the device
        OneFrame O;
        if (!one_frame(S, F, &O)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
        orbm_points Qp;
        int* d_nq = S.up(&nq, 1);
        Qp.nq = d_nq; Qp.nq_stride = nq;
        Qp.proj_xyxr = S.up(proj_xyxr, (size_t)nq * 3); Qp.level = S.up(level, (size_t)nq); Qp.view_cos = S.up(view_cos, (size_t)nq);
        Qp.in_view = S.up(in_view, (size_t)nq); Qp.bad = S.up(bad, (size_t)nq); Qp.observations = S.up(observations, (size_t)nq);
        Qp.qdesc = S.up(qdesc, (size_t)nq * 32);
        int* d_init = init_assign ? S.up(init_assign, (size_t)F->n) : nullptr;
        int* d_assign = (int*)S.alloc(sizeof(int) * ((size_t)F->n + 1));
        int* d_nm = d_assign ? d_assign + F->n : nullptr;
        if (!S.ok) return ORBX_E_CUDA;
        if (!S.flush()) return ORBX_E_CUDA;
        const int rc = orbm_search_by_projection_points_batch(&O.F, scale, nlevels, &Qp, d_init, d_assign, th, nnratio, d_nm, nullptr, nullptr);
        if (rc) return rc;
        return fetch_assign(d_assign, F->n, assign_out, nmatches);
    }
    DevFrame D;
    if (!make_frame(S, F, &D)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    // the per-query part of the reference loop that does not depend on other queries (:82-98, :119-124)
    std::vector<WinQuery> hq((size_t)nq);
    const bool bFactor = th != 1.0;
    for (int i = 0; i < nq; ++i) {
        WinQuery& Q = hq[(size_t)i];
        Q.valid = in_view[i] && !bad[i] && level[i] >= 0 && level[i] < nlevels;
        if (!Q.valid) continue;
        float r = ((double)view_cos[i] > 0.998) ? 2.5f : 4.0f;              // RadiusByViewingCos, :1653-1660
        if (bFactor) r *= th;
        Q.u = proj_xyxr[3 * i]; Q.v = proj_xyxr[3 * i + 1];
        Q.r = r * scale[level[i]];
        Q.min_level = level[i] - 1; Q.max_level = level[i];
        Q.ur = proj_xyxr[3 * i + 2]; Q.er_max = r * scale[level[i]];
    }
    WinQuery* dq; uint32_t* list; int* count; int cap;
    int rc = run_candidates(S, D, hq, qdesc, nq, &dq, &list, &count, &cap);
    if (rc) return rc;
    int* d_obs = S.up(observations, (size_t)nq);
    int* d_assign = S.up(assign_out, (size_t)F->n);
    int* d_nm = (int*)S.alloc(4);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    k_resolve_points<<<1, 32>>>(D, nq, dq, list, count, cap, d_obs, d_assign, nnratio, d_nm);
    CKM(cudaGetLastError());
    CKM(cudaMemcpy(assign_out, d_assign, sizeof(int) * (size_t)F->n, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(nmatches, d_nm, 4, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbm_search_by_projection_frame(const orbm_frame* cur, int n_last, const orbx_kp* kps_last,
                                    const uint8_t* last_mp, const uint8_t* last_outlier, const float* last_xyz,
                                    const uint8_t* last_mp_desc, const int* last_mp_obs,
                                    const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                    const float* scale, int nlevels, const int* cur_init_obs, int* assign_out,
                                    float th, int bMono, int checkOri, int* nmatches, int device)
{
    if (!cur || n_last < 0 || (cur->n > 0 && !assign_out) || !nmatches || !Tcw_cur || !Tcw_last || !K || !scale ||
        (n_last > 0 && (!kps_last || !last_mp || !last_xyz || !last_mp_desc))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    const int n_cur = cur->n;
    std::vector<int> h_obs((size_t)(n_cur > 0 ? n_cur : 1));
    for (int k = 0; k < n_cur; ++k) { h_obs[(size_t)k] = cur_init_obs ? cur_init_obs[k] : -1; assign_out[k] = (cur_init_obs && cur_init_obs[k] >= 0) ? -2 : -1; }
    if (n_last == 0 || n_cur == 0) return ORBX_OK;
    Scratch S;
    DevFrame D;
    const bool via_batch = orb_match_batch_fits(n_cur, n_last);
    if (!via_batch && !make_frame(S, cur, &D)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    // projection of the last frame's map points with the current pose: the float arithmetic of
    // cv::gemm for CV_32F (products accumulated left to right, translation added last), :172-208
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const float mb = bf / fx;
    float twc[3], tlc[3];
    for (int i = 0; i < 3; ++i) {
        float s = (-Tcw_cur[0 * 4 + i]) * Tcw_cur[3];
        s = s + (-Tcw_cur[1 * 4 + i]) * Tcw_cur[7];
        s = s + (-Tcw_cur[2 * 4 + i]) * Tcw_cur[11];
        twc[i] = s;
    }
    for (int i = 0; i < 3; ++i) {
        float s = Tcw_last[4 * i] * twc[0];
        s = s + Tcw_last[4 * i + 1] * twc[1];
        s = s + Tcw_last[4 * i + 2] * twc[2];
        tlc[i] = s + Tcw_last[4 * i + 3];
    }
    const bool bForward = tlc[2] > mb && !bMono, bBackward = -tlc[2] > mb && !bMono;
    std::vector<WinQuery> hq((size_t)n_last);
    std::vector<float> h_angle((size_t)n_last);
    std::vector<int> h_lobs((size_t)n_last);
    for (int i = 0; i < n_last; ++i) {
        WinQuery& Q = hq[(size_t)i];
        h_angle[(size_t)i] = kps_last[i].angle;
        h_lobs[(size_t)i] = last_mp_obs ? last_mp_obs[i] : 1;
        Q.valid = 0;
        if (!last_mp[i] || (last_outlier && last_outlier[i])) continue;
        const float* x = last_xyz + 3 * i;
        float pc[3];
        for (int r = 0; r < 3; ++r) {
            float s = Tcw_cur[4 * r] * x[0];
            s = s + Tcw_cur[4 * r + 1] * x[1];
            s = s + Tcw_cur[4 * r + 2] * x[2];
            pc[r] = s + Tcw_cur[4 * r + 3];
        }
        const float invzc = (float)(1.0 / (double)pc[2]);                       // :199
        if (invzc < 0) continue;
        const float u = fx * pc[0] * invzc + cx, v = fy * pc[1] * invzc + cy;
        if (u < cur->min_x || u > cur->max_x) continue;
        if (v < cur->min_y || v > cur->max_y) continue;
        const int oct = kps_last[i].octave;
        if (oct < 0 || oct >= nlevels) continue;
        Q.valid = 1; Q.u = u; Q.v = v; Q.r = th * scale[oct];
        if (bForward) { Q.min_level = oct; Q.max_level = -1; }
        else if (bBackward) { Q.min_level = 0; Q.max_level = oct; }
        else { Q.min_level = oct - 1; Q.max_level = oct + 1; }
        Q.ur = u - bf * invzc; Q.er_max = Q.r;                                  // :241-244
    }
    if (via_batch)
        return best_via_batch(S, cur, cur->u_right != nullptr, hq, last_mp_desc, h_angle.data(), h_lobs.data(), cur_init_obs ? h_obs.data() : nullptr,
                              assign_out, checkOri, TH_HIGH, nmatches);
    WinQuery* dq; uint32_t* list; int* count; int cap;
    int rc = run_candidates(S, D, hq, last_mp_desc, n_last, &dq, &list, &count, &cap);
    if (rc) return rc;
    return resolve_best(S, D, n_last, dq, list, count, cap, h_lobs.data(), h_angle.data(), h_obs.data(), assign_out, checkOri, TH_HIGH, nmatches);
}

int orbm_window_search_best(const orbm_frame* F, int nq, const float* uvr, const int* min_level, const int* max_level,
                            const float* ur, const float* er_max, const uint8_t* valid, const uint8_t* qdesc,
                            const float* q_angle, const int* q_obs, const int* init_obs, int* assign_out,
                            int th_accept, int check_ori, int* nmatches, int device)
{
    if (!F || nq < 0 || (F->n > 0 && !assign_out) || !nmatches || (nq > 0 && (!uvr || !min_level || !max_level || !qdesc))) return ORBX_E_ARG;
    if (check_ori && nq > 0 && !q_angle) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    const int n = F->n;
    std::vector<int> h_obs((size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) { h_obs[(size_t)k] = init_obs ? init_obs[k] : -1; assign_out[k] = (init_obs && init_obs[k] >= 0) ? -2 : -1; }
    if (nq == 0 || n == 0) return ORBX_OK;
    Scratch S;
    DevFrame D;
    const bool via_batch = orb_match_batch_fits(n, nq);
    if (!via_batch && !make_frame(S, F, &D)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    std::vector<WinQuery> hq((size_t)nq);
    std::vector<float> h_angle((size_t)nq, 0.f);
    std::vector<int> h_qobs((size_t)nq, 1);
    for (int i = 0; i < nq; ++i) {
        WinQuery& Q = hq[(size_t)i];
        Q.valid = valid ? (valid[i] != 0) : 1;
        Q.u = uvr[3 * i]; Q.v = uvr[3 * i + 1]; Q.r = uvr[3 * i + 2];
        Q.min_level = min_level[i]; Q.max_level = max_level[i];
        Q.ur = ur ? ur[i] : 0.f; Q.er_max = er_max ? er_max[i] : 3.0e38f;
        if (q_angle) h_angle[(size_t)i] = q_angle[i];
        if (q_obs) h_qobs[(size_t)i] = q_obs[i];
    }
    if (via_batch)
        return best_via_batch(S, F, ur != nullptr && F->u_right != nullptr, hq, qdesc, h_angle.data(), h_qobs.data(), init_obs ? h_obs.data() : nullptr,
                              assign_out, check_ori, th_accept, nmatches);
    if (!ur) D.u_right = nullptr;
    WinQuery* dq; uint32_t* list; int* count; int cap;
    int rc = run_candidates(S, D, hq, qdesc, nq, &dq, &list, &count, &cap);
    if (rc) return rc;
    return resolve_best(S, D, nq, dq, list, count, cap, h_qobs.data(), h_angle.data(), h_obs.data(), assign_out, check_ori, th_accept, nmatches);
}

// One Fuse / SearchBySim3 search (orbm_window_best_free_batch with one problem): host arrays in, one launch, the two
// result arrays and the count come back in one copy.
int orbm_window_best_free(const orbm_frame* F, int nq, const float* uvr, const int* level, const float* ur, const uint8_t* valid,
                          const uint8_t* qdesc, const float* inv_sigma2, int nlevels, int th_accept,
                          int* best_idx, int* best_dist, int* nfound, int device)
{
    if (!F || nq < 0 || !nfound || (nq > 0 && (!uvr || !level || !qdesc || !best_idx || !best_dist))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nfound = 0;
    for (int i = 0; i < nq; ++i) { best_idx[i] = -1; best_dist[i] = 256; }
    if (nq == 0 || F->n == 0) return ORBX_OK;
    Scratch S;
    OneFrame O;
    if (!one_frame(S, F, &O)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    orbm_free_windows W;
    int* d_nq = S.up(&nq, 1);
    W.nq = d_nq; W.nq_stride = nq;
    W.uvr = S.up(uvr, (size_t)nq * 3); W.level = S.up(level, (size_t)nq);
    W.ur = ur ? S.up(ur, (size_t)nq) : nullptr;
    W.valid = valid ? S.up(valid, (size_t)nq) : nullptr;
    W.qdesc = S.up(qdesc, (size_t)nq * 32);
    int* d_out = (int*)S.alloc(sizeof(int) * ((size_t)nq * 2 + 1));
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_window_best_free_batch(&O.F, &W, inv_sigma2, nlevels, th_accept, d_out, d_out + nq, d_out + 2 * (size_t)nq, nullptr);
    if (rc) return rc;
    thread_local std::vector<int> tmp;
    tmp.resize((size_t)nq * 2 + 1);
    const cudaError_t e = cudaMemcpy(tmp.data(), d_out, sizeof(int) * ((size_t)nq * 2 + 1), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { report_cuda(e, "orbm_window_best_free", __LINE__); cudaGetLastError(); return ORBX_E_CUDA; }
    if (tmp[(size_t)nq * 2] < 0) return ORBX_E_ARG;          // more than 8192 keypoints
    memcpy(best_idx, tmp.data(), sizeof(int) * (size_t)nq);
    memcpy(best_dist, tmp.data() + nq, sizeof(int) * (size_t)nq);
    *nfound = tmp[(size_t)nq * 2];
    return ORBX_OK;
}

int orbm_search_for_initialization(const orbm_frame* F1, const orbm_frame* F2, float* prev_matched, int* matches12,
                                   int windowSize, float nnratio, int checkOri, int* nmatches, int device)
{
    if (!F1 || !F2 || !prev_matched || !matches12 || !nmatches || F1->n < 0 || (F1->n > 0 && (!F1->kps || !F1->desc))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    const int n1 = F1->n, n2 = F2->n;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    Scratch S;
    if (n1 <= MAX_KP && n2 <= MAX_KP && orb_match_batch_fits(n2, n1)) {   // always true up to 8192 x 8192
        // one launch of the block-per-problem kernel (orb_match_batch.cu k_init_fixpoint)
        OneFrame O1, O2;
        if (!one_frame(S, F1, &O1) || !one_frame(S, F2, &O2)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
        float* d_prev = S.up(prev_matched, (size_t)n1 * 2);
        int* d_m12 = (int*)S.alloc(sizeof(int) * (size_t)n1);
        int* d_nm = (int*)S.alloc(4);
        if (!S.ok || !S.flush()) return ORBX_E_CUDA;
        const int rc = orbm_search_for_initialization_batch(&O1.F, &O2.F, d_prev, d_m12, windowSize, nnratio, checkOri, d_nm, nullptr, nullptr);
        if (rc) return rc;
        CKM(cudaMemcpy(matches12, d_m12, sizeof(int) * (size_t)n1, cudaMemcpyDeviceToHost));
        CKM(cudaMemcpy(prev_matched, d_prev, sizeof(float) * (size_t)n1 * 2, cudaMemcpyDeviceToHost));
        CKM(cudaMemcpy(nmatches, d_nm, 4, cudaMemcpyDeviceToHost));
        return *nmatches < 0 ? ORBX_E_ARG : ORBX_OK;
    }
    return ORBX_E_ARG;     // more than 8192 keypoints on a side
}

int orbm_stereo_matches(orbx_ctx* ex_left, int frame_l, orbx_ctx* ex_right, int frame_r,
                        int nl, const orbx_kp* kps_l, const uint8_t* desc_l,
                        int nr, const orbx_kp* kps_r, const uint8_t* desc_r,
                        float bf, float fx, float* u_right, float* depth, int* nmatched)
{
    if (!ex_left || !ex_right || nl < 0 || nr < 0 || !u_right || !depth || (nl > 0 && (!kps_l || !desc_l)) || (nr > 0 && (!kps_r || !desc_r)) || nr > 65535)
        return ORBX_E_ARG;
    OrbStereoView V;
    memset(&V, 0, sizeof(V));
    int dev_l = 0, dev_r = 0;
    int wr[ORB_MAX_LEVELS], hr[ORB_MAX_LEVELS], nlr = 0;
    if (orb_ctx_levels(ex_left, frame_l, V.l, V.lpitch, V.w, V.h, V.scale, V.inv_scale, &V.nlevels, &dev_l) ||
        orb_ctx_levels(ex_right, frame_r, V.r, V.rpitch, wr, hr, nullptr, nullptr, &nlr, &dev_r) || dev_l != dev_r || nlr != V.nlevels)
        return ORBX_E_ARG;
    for (int l = 0; l < V.nlevels; ++l) if (wr[l] != V.w[l] || hr[l] != V.h[l]) return ORBX_E_ARG;   // both extractors must have seen the same shape
    if (cudaSetDevice(dev_l) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    if (nmatched) *nmatched = 0;
    for (int i = 0; i < nl; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    if (nl == 0 || nr == 0) return ORBX_OK;
    Scratch S;
    const int cap = nl > nr ? nl : nr;
    const int counts[2] = { nl, nr };
    V.kl = S.up(kps_l, (size_t)nl);
    V.kr = S.up(kps_r, (size_t)nr);
    V.dl = (const uint32_t*)S.up(desc_l, (size_t)nl * 32);
    V.dr = (const uint32_t*)S.up(desc_r, (size_t)nr * 32);
    const int* dcounts = S.up(counts, 2);
    V.nl = dcounts; V.nr = dcounts + 1; V.nstride = 0; V.kstride = 0; V.cap = cap;
    V.bf = bf; V.mb = bf / fx;                                                   // src/Frame.cc:121
    V.u_right = (float*)S.alloc(sizeof(float) * (size_t)nl);
    V.depth = (float*)S.alloc(sizeof(float) * (size_t)nl);
    V.ostride = 0;
    V.n_stereo = (int*)S.alloc(4);
    V.sad = (int*)S.alloc(sizeof(int) * (size_t)cap);
    V.rec = (uint4*)S.alloc(sizeof(uint4) * (size_t)cap);
    V.row_start = (int*)S.alloc(sizeof(int) * (size_t)(V.h[0] + 2));
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    CKM(orb_launch_stereo(V, 1, nl, 0));
    CKM(cudaMemcpy(u_right, V.u_right, sizeof(float) * (size_t)nl, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(depth, V.depth, sizeof(float) * (size_t)nl, cudaMemcpyDeviceToHost));
    int n = 0;
    CKM(cudaMemcpy(&n, V.n_stereo, 4, cudaMemcpyDeviceToHost));
    if (nmatched) *nmatched = n;
    return ORBX_OK;
}


// ---- single problems with HOST arrays over the device-resident kernels (what the C++ forwarders call)

int orbm_search_by_bow(const orbm_frame* A, const uint8_t* a_valid, int nn_a, const int* node_id_a, const int* node_off_a, const int* feat_a,
                       const orbm_frame* B, const uint8_t* b_valid, int nn_b, const int* node_id_b, const int* node_off_b, const int* feat_b,
                       int kf_kf, float nnratio, int check_ori, int* match12, int* nmatches, int device)
{
    if (!A || !B || !a_valid || !match12 || !nmatches || nn_a < 0 || nn_b < 0 || A->n < 0 || B->n < 0) return ORBX_E_ARG;
    if ((nn_a > 0 && (!node_id_a || !node_off_a || !feat_a)) || (nn_b > 0 && (!node_id_b || !node_off_b || !feat_b))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    for (int i = 0; i < A->n; ++i) match12[i] = -1;
    if (A->n == 0 || B->n == 0 || nn_a == 0 || nn_b == 0) return ORBX_OK;
    Scratch S;
    OneFrame OA, OB;
    if (!one_frame(S, A, &OA) || !one_frame(S, B, &OB)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    OA.F.u_right = nullptr; OB.F.u_right = nullptr;
    const int ta = node_off_a[nn_a], tb = node_off_b[nn_b];
    if (ta < 0 || ta > A->n || tb < 0 || tb > B->n) return ORBX_E_ARG;
    std::vector<int> fa((size_t)A->n, 0), fb((size_t)B->n, 0);
    memcpy(fa.data(), feat_a, sizeof(int) * (size_t)ta);
    memcpy(fb.data(), feat_b, sizeof(int) * (size_t)tb);
    orbm_featvec VA, VB;
    VA.node_id = S.up(node_id_a, (size_t)nn_a); VA.node_off = S.up(node_off_a, (size_t)nn_a + 1); VA.n_nodes = S.up(&nn_a, 1);
    VA.feat = S.up(fa.data(), fa.size()); VA.node_stride = nn_a;
    VB.node_id = S.up(node_id_b, (size_t)nn_b); VB.node_off = S.up(node_off_b, (size_t)nn_b + 1); VB.n_nodes = S.up(&nn_b, 1);
    VB.feat = S.up(fb.data(), fb.size()); VB.node_stride = nn_b;
    const uint8_t* d_av = S.up(a_valid, (size_t)A->n);
    const uint8_t* d_bv = b_valid ? S.up(b_valid, (size_t)B->n) : nullptr;
    int* d_m = (int*)S.alloc(sizeof(int) * (size_t)A->n);
    int* d_nm = (int*)S.alloc(4);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_search_by_bow_batch(&OA.F, &VA, d_av, &OB.F, &VB, d_bv, kf_kf, nnratio, check_ori, d_m, nullptr, d_nm, nullptr, nullptr);
    if (rc) return rc;
    CKM(cudaMemcpy(match12, d_m, sizeof(int) * (size_t)A->n, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(nmatches, d_nm, 4, cudaMemcpyDeviceToHost));
    return *nmatches < 0 ? ORBX_E_ARG : ORBX_OK;
}

// ORBmatcher::SearchForTriangulation for one key-frame pair with HOST arrays (orbm_search_for_triangulation_batch with one problem)
int orbm_search_for_triangulation(const orbm_frame* A, const uint8_t* a_valid, int nn_a, const int* node_id_a, const int* node_off_a, const int* feat_a,
                                  const orbm_frame* B, const uint8_t* b_valid, int nn_b, const int* node_id_b, const int* node_off_b, const int* feat_b,
                                  const float* F12, const float* epipole, const float* scale, const float* sigma2, int nlevels,
                                  int check_ori, int* match12, int* nmatches, int device)
{
    if (!A || !B || !a_valid || !b_valid || !match12 || !nmatches || nn_a < 0 || nn_b < 0 || A->n < 0 || B->n < 0 || !F12 || !epipole || !scale ||
        !sigma2 || nlevels <= 0 || nlevels > 32)
        return ORBX_E_ARG;
    if ((nn_a > 0 && (!node_id_a || !node_off_a || !feat_a)) || (nn_b > 0 && (!node_id_b || !node_off_b || !feat_b))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *nmatches = 0;
    for (int i = 0; i < A->n; ++i) match12[i] = -1;
    if (A->n == 0 || B->n == 0 || nn_a == 0 || nn_b == 0) return ORBX_OK;
    Scratch S;
    OneFrame OA, OB;
    if (!one_frame(S, A, &OA) || !one_frame(S, B, &OB)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    const int ta = node_off_a[nn_a], tb = node_off_b[nn_b];
    if (ta < 0 || ta > A->n || tb < 0 || tb > B->n) return ORBX_E_ARG;
    std::vector<int> fa((size_t)A->n, 0), fb((size_t)B->n, 0);
    memcpy(fa.data(), feat_a, sizeof(int) * (size_t)ta);
    memcpy(fb.data(), feat_b, sizeof(int) * (size_t)tb);
    orbm_featvec VA, VB;
    VA.node_id = S.up(node_id_a, (size_t)nn_a); VA.node_off = S.up(node_off_a, (size_t)nn_a + 1); VA.n_nodes = S.up(&nn_a, 1);
    VA.feat = S.up(fa.data(), fa.size()); VA.node_stride = nn_a;
    VB.node_id = S.up(node_id_b, (size_t)nn_b); VB.node_off = S.up(node_off_b, (size_t)nn_b + 1); VB.n_nodes = S.up(&nn_b, 1);
    VB.feat = S.up(fb.data(), fb.size()); VB.node_stride = nn_b;
    const uint8_t* d_av = S.up(a_valid, (size_t)A->n);
    const uint8_t* d_bv = S.up(b_valid, (size_t)B->n);
    float fe[12];
    memcpy(fe, F12, sizeof(float) * 9); fe[9] = epipole[0]; fe[10] = epipole[1]; fe[11] = 0.f;
    const float* d_fe = S.up(fe, 12);
    int* d_m = (int*)S.alloc(sizeof(int) * ((size_t)A->n + 1));
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_search_for_triangulation_batch(&OA.F, &VA, d_av, &OB.F, &VB, d_bv, d_fe, d_fe + 9, scale, sigma2, nlevels, check_ori,
                                                       d_m, d_m + A->n, nullptr);
    if (rc) return rc;
    const int rf = fetch_assign(d_m, A->n, match12, nmatches);
    if (rf) return rf;
    return *nmatches < 0 ? ORBX_E_ARG : ORBX_OK;
}

int orbm_project_points(const float* Tcw, const float* K, float bf, float min_x, float max_x, float min_y, float max_y,
                        float scale_factor, int nlevels, float viewing_cos_limit, int n, const float* xyz, const float* normal,
                        const float* max_distance, const float* min_distance, uint8_t* in_view, float* proj_xyxr, int* level,
                        float* view_cos, int device)
{
    if (!Tcw || !K || n < 0 || (n > 0 && (!xyz || !normal || !max_distance || !min_distance || !in_view || !proj_xyxr || !level || !view_cos))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    if (n == 0) return ORBX_OK;
    Scratch S;
    const float* d_T = S.up(Tcw, 16);
    const int* d_n = S.up(&n, 1);
    const float* d_x = S.up(xyz, (size_t)n * 3); const float* d_nrm = S.up(normal, (size_t)n * 3);
    const float* d_mx = S.up(max_distance, (size_t)n); const float* d_mn = S.up(min_distance, (size_t)n);
    uint8_t* d_iv = (uint8_t*)S.alloc((size_t)n);
    float* d_p = S.up(proj_xyxr, (size_t)n * 3);               // entries of points not in view keep the caller's values
    int* d_l = S.up(level, (size_t)n);
    float* d_vc = S.up(view_cos, (size_t)n);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_project_points_batch(1, d_T, K, bf, min_x, max_x, min_y, max_y, scale_factor, nlevels, viewing_cos_limit, d_n, n, 0,
                                             d_x, d_nrm, d_mx, d_mn, d_iv, d_p, d_l, d_vc, nullptr, nullptr);
    if (rc) return rc;
    CKM(cudaMemcpy(in_view, d_iv, (size_t)n, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(proj_xyxr, d_p, sizeof(float) * (size_t)n * 3, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(level, d_l, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(view_cos, d_vc, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbm_distinctive_descriptor(const uint8_t* desc, int n, const uint8_t* bad, int* best_idx, int* best_median, int device)
{
    if (n < 0 || !best_idx || (n > 0 && !desc)) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    *best_idx = -1;
    if (best_median) *best_median = -1;
    if (n == 0) return ORBX_OK;
    Scratch S;
    const int off[2] = { 0, n };
    const uint8_t* d_d = S.up(desc, (size_t)n * 32);
    const uint8_t* d_b = bad ? S.up(bad, (size_t)n) : nullptr;
    const int* d_o = S.up(off, 2);
    int* d_r = (int*)S.alloc(8);
    if (!S.ok) return ORBX_E_CUDA;
    if (!S.flush()) return ORBX_E_CUDA;
    const int rc = orbm_distinctive_descriptors(d_d, d_o, 1, d_b, d_r, d_r + 1, nullptr);
    if (rc) return rc;
    int r[2];
    CKM(cudaMemcpy(r, d_r, 8, cudaMemcpyDeviceToHost));
    *best_idx = r[0];
    if (best_median) *best_median = r[1];
    return r[0] == -2 ? ORBX_E_ARG : ORBX_OK;
}


// ---- test taps (tests/test_match_gpu.py): the pieces of the matchers that otherwise only show through match indices
int orbm_debug_three_maxima(const int* sizes, int n, int* ind, int device)
{
    if (!sizes || !ind || n < 0) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    if (n == 0) return ORBX_OK;
    Scratch S;
    const int* d_sizes = S.up(sizes, (size_t)n * HISTO_LENGTH);
    int* d_ind = (int*)S.alloc(sizeof(int) * 3 * (size_t)n);
    if (!S.ok || !S.flush()) return ORBX_E_CUDA;
    k_debug_three_maxima<<<(n + 127) / 128, 128>>>(d_sizes, n, d_ind);
    CKM(cudaGetLastError());
    CKM(cudaMemcpy(ind, d_ind, sizeof(int) * 3 * (size_t)n, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbm_debug_features_in_area(const orbm_frame* F, int nq, const float* xyr, const int* min_level, const int* max_level,
                                int cap, int* idx_out, int* count_out, int device)
{
    if (!F || nq < 0 || cap <= 0 || !idx_out || !count_out || (nq > 0 && (!xyr || !min_level || !max_level))) return ORBX_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    for (int q = 0; q < nq; ++q) count_out[q] = 0;
    if (nq == 0 || F->n == 0) return ORBX_OK;
    Scratch S;
    DevFrame D;
    if (!make_frame(S, F, &D)) return S.ok ? ORBX_E_ARG : ORBX_E_CUDA;
    std::vector<WinQuery> hq((size_t)nq);
    for (int q = 0; q < nq; ++q) {
        WinQuery w;
        w.u = xyr[3 * q]; w.v = xyr[3 * q + 1]; w.r = xyr[3 * q + 2];
        w.min_level = min_level[q]; w.max_level = max_level[q]; w.ur = 0.f; w.er_max = 0.f; w.valid = 1;
        hq[(size_t)q] = w;
    }
    std::vector<uint32_t> zero((size_t)nq * 8, 0u);
    const WinQuery* dq = S.up(hq.data(), (size_t)nq);
    const uint32_t* dd = S.up(zero.data(), (size_t)nq * 8);
    uint32_t* list = (uint32_t*)S.alloc(sizeof(uint32_t) * (size_t)nq * cap);
    int* count = (int*)S.alloc(sizeof(int) * (size_t)nq);
    if (!S.ok || !S.flush()) return ORBX_E_CUDA;
    int sn = 32; while (sn < D.n) sn <<= 1;
    k_grid_build<<<1, 1024, sizeof(uint32_t) * (size_t)sn>>>(D);
    k_window_candidates<<<(nq + 7) / 8, 256>>>(D, dq, dd, nq, list, count, cap);
    CKM(cudaGetLastError());
    std::vector<uint32_t> hl((size_t)nq * cap);
    CKM(cudaMemcpy(hl.data(), list, sizeof(uint32_t) * hl.size(), cudaMemcpyDeviceToHost));
    CKM(cudaMemcpy(count_out, count, sizeof(int) * (size_t)nq, cudaMemcpyDeviceToHost));
    for (int q = 0; q < nq; ++q)
        for (int k = 0; k < count_out[q] && k < cap; ++k) idx_out[(size_t)q * cap + k] = (int)(hl[(size_t)q * cap + k] & 0xffffu);
    return ORBX_OK;
}

} // extern "C"
