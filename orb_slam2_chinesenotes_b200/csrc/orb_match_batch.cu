// orb_match_batch.cu -- batched, device-resident window matchers (sm_100a): one thread block per
// (frame, query set) problem, many problems per launch, nothing leaves the GPU between extraction
// and matching.
//   MODE_POINTS  ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)   src/ORBmatcher.cc:73-157
//   MODE_BEST    the best-candidate-only overloads once the caller has projected its points
//                (motion model :160-300, relocalisation :303-431, loop closing :434-549)
//   k_init_fixpoint   ORBmatcher::SearchForInitialization                                :1055-1180
//   + Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:243-259, 348-409) and the rotation
//     histogram with ComputeThreeMaxima (src/ORBmatcher.cc:1663-1707)
//
// How the reference's query ORDER is kept without walking the queries one after another.  A query skips a
// keypoint k when the map point attached to k has Observations() > 0 (:115-117, :234-236).  Attached points only
// change by earlier queries claiming k, and once a claimant with Observations() > 0 holds k nobody later can take
// it, so "k is taken when query q runs"  <=>  blocker[k] < q, where blocker[k] is the FIRST query that claimed k
// with Observations() > 0 (-1 when the point attached before the call has Observations() > 0).  Given every
// query's decision, blocker[] follows; given blocker[], every query's decision follows independently of the
// others.  The kernel iterates the two steps to their fixpoint: after round r the decisions of the first r
// queries are the sequential ones (a query only reads blockers of earlier queries), so the fixpoint is reached,
// is unique and equals the sequential result; in practice 2-4 rounds.  After the first round a query is only
// recomputed when one of its two best candidates became blocked, or when it had skipped a blocked candidate and
// some decision changed -- nothing else can alter its result.
// Measured dead end (round 2): compacting the queries that need recomputation into a list between rounds (ballot +
// prefix into the dead sort-key storage), so that the warps of rounds 1.. work on 32 needy queries at a time instead of
// the one or two among the 32 they own, costs two more block barriers and a pass over the queries per round and
// measures 0.45-0.49 ms per 444 problems against 0.405 ms: the later rounds are cheap already, the time is in the grid
// sort and round 0.
// Measured dead end (round 2, last session): a "two-speed" window walk -- every lane runs ahead over its share of the window's
// candidates (positions sub, sub + 4, ... of the concatenated column runs) with only the cheap tests (level, exact window, taken,
// right coordinate) until one passes, then the lanes that hold one compute their distances together -- is bit-exact and 2.8x
// SLOWER (0.97 vs 0.35 ms per 444 problems): the warp waits for its slowest scanner before EVERY distance step, so it pays the
// sum over steps of the longest gap instead of the longest lane's total.  (Queueing the survivors per 4-lane group and computing
// distances on full groups would bound that, but a scan step costs ~36 instructions against ~58 for the plain iteration, the
// distance part being only ~30 of them: at the measured ~30 % pass rate it would save ~15 %.)
//
// Layout per problem (built once per launch in the block): keypoints sorted by (grid column, grid row, index) --
// the order GetFeaturesInArea returns them in -- as 16-byte records {x, y, octave | index << 8 | taken << 31,
// uRight} plus their descriptors in the same order, so the candidates of one grid column of a window are
// CONTIGUOUS: a warp reads 32 records and 32 descriptors with unit stride, no index indirection.  Position in
// that order doubles as the tie-break of the reference's running best/second-best ("two smallest by
// (distance, position)").
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

#include <climits>
#include <cstdio>
#include <cstdlib>

#include "../../include/orb_b200.h"

#define GRID_COLS 64    // FRAME_GRID_COLS, include/Frame.h:38
#define GRID_ROWS 48    // FRAME_GRID_ROWS, include/Frame.h:37
#define GRID_CELLS (GRID_COLS * GRID_ROWS)
#define TH_HIGH 100     // src/ORBmatcher.cc:37
#include "orb_match_common.cuh"   // HISTO_LENGTH, orb_three_maxima
#define MB_MAX_KP 8192
#define MB_MAX_LEVELS 32
#ifndef MB_NT
#define MB_NT 1024
#endif
#define MB_NONE 0xffffffffu
#define MB_CL 6            // claimants per keypoint kept for the parallel rounds of SearchForInitialization
#define TH_LOW 50         // src/ORBmatcher.cc:38

enum { MODE_POINTS = 0, MODE_BEST = 1 };

struct MbParams {
    // frames
    const orbx_kp* kps; const uint32_t* desc; const float* u_right; const int* n; int kp_stride, n_bound;
    float min_x, min_y, max_x, max_y, inv_w, inv_h;
    // queries
    const int* nq; int nq_stride;
    const float* f0;          // POINTS: proj_xyxr [.,3]            BEST: uvr [.,3]
    const int* i0;            // POINTS: level                      BEST: min_level
    const int* i1;            //                                    BEST: max_level
    const float* f1;          // POINTS: view_cos                   BEST: ur (or null)
    const float* f2;          //                                    BEST: er_max (or null)
    const uint8_t* b0;        // POINTS: in_view                    BEST: valid (or null)
    const uint8_t* b1;        // POINTS: bad
    const int* qobs;          // Observations() of the queries' points (BEST: null = 1)
    const uint32_t* qdesc;
    const float* q_angle;     // BEST + check_ori
    const int* init;          // POINTS: init_assign [.,kp_stride]  BEST: init_obs [.,kp_stride]   (or null)
    int* assign_out; int* nmatches;
    float th, nnratio; int nlevels, th_accept, check_ori;
    float scale[MB_MAX_LEVELS];
    // workspace
    uint4* rec; uint32_t* sdesc;     // [nprob][kp_stride], [nprob][kp_stride][8]
    int sn_max, nq_max;              // shared-memory sizing
    int rec_in_smem, desc_in_smem;   // the sorted records / descriptors live in shared memory when they fit
    // SearchForInitialization only
    const orbx_kp* kps1;             // F1.mvKeysUn [nprob][nq_stride]
    float* prev;                     // vbPrevMatched [nprob][nq_stride][2], updated in place
    int* matches12;                  // vnMatches12 [nprob][nq_stride]
    uint32_t* cl;                    // workspace [nprob][kp_stride][MB_CL]: claimants (query << 9 | distance) of every position
    float window;
    int* rounds;                     // [nprob] or null: fixpoint rounds used (profiling / tests)
    // order-free searches only (k_window_best_free)
    int* best_dist;                  // [nprob][nq_stride]
    int split, chi2;                 // blocks per problem; Fuse's reprojection-error test with scale[] = mvInvLevelSigma2
};

struct WinQ { float u, v, r, ur, er_max; int min_level, max_level, valid; };

template <int MODE>
__device__ __forceinline__ WinQ load_query(const MbParams& P, const size_t qo)
{
    WinQ Q;
    if (MODE == MODE_POINTS) {
        const int level = P.i0[qo];
        Q.valid = P.b0[qo] && !P.b1[qo] && level >= 0 && level < P.nlevels;     // :82-85
        float r = ((double)P.f1[qo] > 0.998) ? 2.5f : 4.0f;                     // RadiusByViewingCos, :1653-1660
        if (P.th != 1.0f) r = __fmul_rn(r, P.th);                               // :76, :90-91
        const float rs = Q.valid ? __fmul_rn(r, P.scale[level]) : 0.f;
        Q.u = P.f0[3 * qo]; Q.v = P.f0[3 * qo + 1]; Q.ur = P.f0[3 * qo + 2];
        Q.r = rs; Q.er_max = rs;                                                // :93-95, :121-123
        Q.min_level = level - 1; Q.max_level = level;
    } else {
        Q.valid = P.b0 ? (P.b0[qo] != 0) : 1;
        Q.u = P.f0[3 * qo]; Q.v = P.f0[3 * qo + 1]; Q.r = P.f0[3 * qo + 2];
        Q.min_level = P.i0[qo]; Q.max_level = P.i1[qo];
        Q.ur = P.f1 ? P.f1[qo] : 0.f; Q.er_max = P.f2 ? P.f2[qo] : 3.0e38f;
    }
    return Q;
}

__device__ __forceinline__ int mb_rot_bin(const float a1, const float a2)   // src/ORBmatcher.cc:263-268
{
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// One query, one GROUP of MB_G lanes (a window holds a few candidates per grid column, far fewer than 32): the two
// smallest (distance << 16 | position) keys among the candidates that are not taken, and whether a taken
// candidate was skipped.  gmask = the lanes of the group; all groups of a warp run this together.
#ifndef MB_G
#define MB_G 4
#endif
// taken(j): the candidate at position j is held by an earlier query (tested before the distance is computed);
// beaten(j, dist): an earlier query matched it at a distance <= dist (SearchForInitialization only).
template <typename Taken, typename Beaten>
__device__ __forceinline__ void query_top2(const MbParams& P, const WinQ& Q, const uint32_t (&qd)[8],
                                           const uint4* rec, const uint4* sdesc, const int* cell_start,
                                           Taken taken, Beaten beaten, const bool use_ur, const int sub, const unsigned gmask,
                                           uint32_t& k1, uint32_t& k2, bool& skipped)
{
    uint32_t a1 = MB_NONE, a2 = MB_NONE;
    int nb = 0;
    // src/Frame.cc:355-372
    const int nMinCellX = max(0, (int)floorf((Q.u - P.min_x - Q.r) * P.inv_w));
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf((Q.u - P.min_x + Q.r) * P.inv_w));
    const int nMinCellY = max(0, (int)floorf((Q.v - P.min_y - Q.r) * P.inv_h));
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf((Q.v - P.min_y + Q.r) * P.inv_h));
    if (Q.valid && !(nMinCellX >= GRID_COLS || nMaxCellX < 0 || nMinCellY >= GRID_ROWS || nMaxCellY < 0)) {
        const bool check_levels = Q.min_level > 0 || Q.max_level >= 0;               // :375
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
            const int s0 = cell_start[ix * GRID_ROWS + nMinCellY], s1 = cell_start[ix * GRID_ROWS + nMaxCellY + 1];
            for (int j = s0 + sub; j < s1; j += MB_G) {
                ORB_CHECK(j >= 0 && j < cell_start[GRID_CELLS]);
                const uint4 r = rec[j];
                const int oct = (int)(r.z & 0xffu);
                if (check_levels) {
                    if (oct < Q.min_level) continue;
                    if (Q.max_level >= 0 && oct > Q.max_level) continue;
                }
                if (!(fabsf(__uint_as_float(r.x) - Q.u) < Q.r && fabsf(__uint_as_float(r.y) - Q.v) < Q.r)) continue;   // :402
                if (taken(j)) { ++nb; continue; }                                     // src/ORBmatcher.cc:115-117 / :234-236
                if (use_ur) {
                    const float ur = __uint_as_float(r.w);
                    if (ur > 0 && fabsf(Q.ur - ur) > Q.er_max) continue;             // :119-124 / :238-244
                }
                const uint4 b0 = sdesc[2 * j], b1 = sdesc[2 * j + 1];
                const uint32_t dist = __popc(qd[0] ^ b0.x) + __popc(qd[1] ^ b0.y) + __popc(qd[2] ^ b0.z) + __popc(qd[3] ^ b0.w) +
                                      __popc(qd[4] ^ b1.x) + __popc(qd[5] ^ b1.y) + __popc(qd[6] ^ b1.z) + __popc(qd[7] ^ b1.w);
                if (beaten(j, (int)dist)) continue;                                   // :1094
                const uint32_t key = (dist << 16) | (uint32_t)j;
                if (key < a1) { a2 = a1; a1 = key; }
                else if (key < a2) a2 = key;
            }
        }
    }
    // group reductions with xor shuffles (redux.sync with a different member mask per group is serialised per mask)
    k1 = a1;
#pragma unroll
    for (int d = MB_G / 2; d > 0; d >>= 1) k1 = min(k1, __shfl_xor_sync(0xffffffffu, k1, d));
    if (a1 == k1) a1 = a2;                        // keys are unique (position) or "none"
    k2 = a1;
#pragma unroll
    for (int d = MB_G / 2; d > 0; d >>= 1) k2 = min(k2, __shfl_xor_sync(0xffffffffu, k2, d));
    skipped = (__ballot_sync(0xffffffffu, nb > 0) & gmask) != 0;
}

// The blocked-candidate order matters: in the reference the "taken" test comes BEFORE the right-image test, and
// both only `continue`, so their order does not change the candidate set.

#ifdef ORB_MATCH_CLOCKS
__device__ long long g_clk[16];
#define CLK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_clk[i] = clock64(); } while (0)
#else
#define CLK(i)
#endif

// Frame::AssignFeaturesToGrid (src/Frame.cc:243-259) for one frame, by the whole block: the cell of every keypoint
// (round(), :414-415), keys = cell << 16 | index in the order (grid column, grid row, index), cell_start[c] = first
// position of cell c (cell_start[GRID_CELLS] = number of keypoints inside the grid).  Ends with a barrier.
// A counting sort -- cell histogram with shared-memory atomics, exclusive scan over the 3072 cells, scatter, and an
// insertion sort inside every cell (a cell holds a handful of keypoints) -- with a dozen block barriers where the bitonic
// sort of 2048 keys needs 66 (measured on one 2000-keypoint frame: 15.1 -> 5.4 us); a frame that crowds more than 32
// keypoints into one cell takes the bitonic sort.  tmp: sn ints of scratch.
template <int NT>
__device__ __forceinline__ void mb_sort_frame(const MbParams& P, const orbx_kp* kps, const int n, const int sn, uint32_t* keys, int* cell_start, int* tmp)
{
    __shared__ int s_scan[NT / 32], s_crowded;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int c = tid; c <= GRID_CELLS; c += NT) cell_start[c] = 0;
    if (tid == 0) s_crowded = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const int px = (int)roundf((kps[i].x - P.min_x) * P.inv_w);
        const int py = (int)roundf((kps[i].y - P.min_y) * P.inv_h);
        const int c = (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
        tmp[i] = c;
        if (c >= 0 && atomicAdd(&cell_start[c], 1) >= 32) s_crowded = 1;
    }
    __syncthreads();
    if (!s_crowded) {
        constexpr int IPT = (GRID_CELLS + NT) / NT;                       // cells per thread, GRID_CELLS + 1 entries in all
        int v[IPT], sum = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int c = tid * IPT + k; v[k] = c <= GRID_CELLS ? cell_start[c] : 0; sum += v[k]; }
        int inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_scan[wid] = inc;
        __syncthreads();
        int run = inc - sum;
        for (int w = 0; w < wid; ++w) run += s_scan[w];
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int c = tid * IPT + k; if (c <= GRID_CELLS) cell_start[c] = run; run += v[k]; }
        __syncthreads();
        // scatter with the start of each cell as its cursor: afterwards cell_start[c] is the END of cell c
        for (int i = tid; i < n; i += NT) {
            const int c = tmp[i];
            if (c >= 0) { const int pos = atomicAdd(&cell_start[c], 1); ORB_CHECK(pos >= 0 && pos < sn && c < GRID_CELLS); keys[pos] = ((uint32_t)c << 16) | (uint32_t)i; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int c = tid * IPT + k; v[k] = (c > 0 && c <= GRID_CELLS) ? cell_start[c - 1] : 0; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int c = tid * IPT + k; if (c <= GRID_CELLS) cell_start[c] = v[k]; }
        __syncthreads();
        for (int c = tid; c < GRID_CELLS; c += NT) {                       // insertion order inside the cell (:256)
            const int s0 = cell_start[c], s1 = cell_start[c + 1];
            for (int i = s0 + 1; i < s1; ++i) {
                const uint32_t key = keys[i];
                int j = i - 1;
                while (j >= s0 && keys[j] > key) { keys[j + 1] = keys[j]; --j; }
                keys[j + 1] = key;
            }
        }
        __syncthreads();
        return;
    }
    for (int i = tid; i < sn; i += NT) keys[i] = (i < n && tmp[i] >= 0) ? (((uint32_t)tmp[i] << 16) | (uint32_t)i) : MB_NONE;
    __syncthreads();
    for (int k = 2; k <= sn; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (sn >> 1); t += NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;
                const uint32_t a = keys[i], b = keys[ixj];
                if (((i & k) == 0) ? (a > b) : (a < b)) { keys[i] = b; keys[ixj] = a; }
            }
            __syncthreads();
        }
    for (int i = tid; i < sn; i += NT) {
        const uint32_t key = keys[i];
        const int c = key == MB_NONE ? GRID_CELLS : (int)(key >> 16);
        const int cprev = i == 0 ? -1 : (keys[i - 1] == MB_NONE ? GRID_CELLS : (int)(keys[i - 1] >> 16));
        for (int cc = cprev + 1; cc <= c; ++cc) cell_start[cc] = i;
        if (i == sn - 1) for (int cc = c + 1; cc <= GRID_CELLS; ++cc) cell_start[cc] = sn;
    }
    __syncthreads();
}

// LOC: where the sorted records / descriptors live -- 0: both in shared memory, 1: records in shared memory,
// descriptors in the global workspace, 2: both global.  A template parameter so that the loads of the window walk
// are LDS, not generic loads.
// CL: blocks per problem.  1: many problems per launch, one block each.  MB_CLUSTER (LOC 0 only): ONE problem spread over a
// thread-block cluster -- what a tracker's single SearchByProjection call is, where one block leaves 147 SMs idle while it
// walks 2000 windows.  Every block of the cluster sorts its own copy of the frame (redundant, but parallel); the queries are
// dealt to the warps of all blocks (32 / MB_G per warp and pass, i.e. one per lane group); a query's state lives in the
// shared memory of the block that owns it and the others read it over distributed shared memory when they rebuild their
// copy of the blocker array; two cluster barriers per round; block 0 collects the decisions and writes the results.
#ifndef MB_CLUSTER
#define MB_CLUSTER 8
#endif
template <int MODE, int LOC, int CL>
__global__ void __launch_bounds__(MB_NT) k_match_fixpoint(const __grid_constant__ MbParams P)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_flag[3], s_cnt[2], s_sizes[HISTO_LENGTH], s_ind[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MB_NT >> 5;
    const int prob = blockIdx.x / CL, crank = CL > 1 ? (int)cg::this_cluster().block_rank() : 0;
    constexpr int QW = CL > 1 ? 32 / MB_G : 32;            // queries a warp takes per pass
    const int gw = crank * nwarps + warp, gwn = CL * nwarps;
    auto owner = [&](const int q) { return ((q / QW) % gwn) / nwarps; };   // cluster rank of the block that holds query q's state
    const int n = P.n[prob], nq = P.nq[prob];
    int* const nm_out = P.nmatches + prob;
    if (n < 0 || n > P.sn_max || n > P.kp_stride || nq < 0 || nq > P.nq_stride) { if (tid == 0 && crank == 0) *nm_out = -1; return; }
    const size_t ko = (size_t)prob * P.kp_stride, qo = (size_t)prob * P.nq_stride;
    const orbx_kp* kps = P.kps + ko;
    const int* init = P.init ? P.init + ko : nullptr;
    int* assign_out = P.assign_out + ko;
    const int* qobs = P.qobs ? P.qobs + qo : nullptr;
    int sn = 32; while (sn < n) sn <<= 1;
    uint32_t* keys = smem;                                  // [sn_max]
    int* cell_start = (int*)(smem + P.sn_max);              // [GRID_CELLS + 1] (+3 pad)
    int* const blk_a = cell_start + GRID_CELLS + 4;         // [sn_max] blocker of every position, double buffered
    int* const blk_b = blk_a + P.sn_max;                    // [sn_max]
    int* blk = blk_a;
    uint32_t* st_top = (uint32_t*)(blk_b + P.sn_max);       // [nq_max]  k1 position | k2 position << 16 (0xffff none)
    uint32_t* st_best = st_top + P.nq_max;                  // [nq_max]  (accepted position + 1) | skipped << 16 | bin << 24
    // sorted records and descriptors: shared memory when the problem fits (the window walk is a chain of dependent
    // loads, so their latency is what a query costs), else the global workspace
    uint4* const sm16 = (uint4*)(st_best + P.nq_max);
    uint4* rec; uint4* sdesc;
    if (LOC == 0) { rec = sm16; sdesc = sm16 + P.sn_max; }
    else if (LOC == 1) { rec = sm16; sdesc = (uint4*)(P.sdesc + ko * 8); }
    else { rec = P.rec + ko; sdesc = (uint4*)(P.sdesc + ko * 8); }

    CLK(0);
    mb_sort_frame<MB_NT>(P, kps, n, sn, keys, cell_start, blk_a);
    const int nvalid = cell_start[GRID_CELLS];
    CLK(2);
    const bool use_ur = P.u_right != nullptr && (MODE == MODE_POINTS || P.f1 != nullptr);
    for (int j = tid; j < nvalid; j += MB_NT) {
        const int idx = (int)(keys[j] & 0xffffu);
        const orbx_kp kp = kps[idx];
        uint32_t taken = 0;
        if (init) {
            const int a = init[idx];
            if (MODE == MODE_POINTS) taken = (a >= 0 && a < nq && qobs[a] > 0) ? 1u : 0u;   // attached point with Observations() > 0 (an index outside the query set counts as free)
            else taken = a > 0 ? 1u : 0u;
        }
        rec[j] = make_uint4(__float_as_uint(kp.x), __float_as_uint(kp.y), (uint32_t)(kp.octave & 0xff) | ((uint32_t)idx << 8) | (taken << 31),
                            __float_as_uint(use_ur ? P.u_right[ko + idx] : -1.0f));
        blk[j] = taken ? -1 : INT_MAX;
    }
    // descriptors into position order: 8 threads move one descriptor (coalesced 32-byte rows on both sides)
    for (int t = tid; t < nvalid * 8; t += MB_NT) {
        const int j = t >> 3, wd = t & 7;
        ((uint32_t*)sdesc)[(size_t)j * 8 + wd] = P.desc[(ko + (keys[j] & 0xffffu)) * 8 + wd];
    }
    for (int q = tid; q < nq; q += MB_NT) { st_top[q] = MB_NONE; st_best[q] = 0; }
    if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
    if (tid < HISTO_LENGTH) s_sizes[tid] = 0;
    __syncthreads();       // rec / sdesc are read back by this block only

    CLK(3);
    // ---- rounds
    int round = 0;
    for (;; ++round) {
        const int par = round & 1;
        const bool unblocked = round > 0 && s_flag[2] != 0;     // some keypoint became free again in the last rebuild
        for (int base = gw * QW; base < nq; base += gwn * QW) {
            const int q = base + lane;
            bool need = false;
            if (lane < QW && q < nq) {
                if (round == 0) need = true;
                else {
                    const uint32_t top = st_top[q], sb = st_best[q];
                    const uint32_t p1 = top & 0xffffu, p2 = top >> 16;
                    need = (p1 != 0xffffu && blk[p1] < q) || (p2 != 0xffffu && blk[p2] < q) || (((sb >> 16) & 1u) && unblocked);
                }
            }
            // every lane fetches ITS query (unit-stride loads, one memory round trip per 32 queries); the warp
            // then works through the queries one at a time with the parameters broadcast by shuffles
            WinQ myQ;
            uint4 myd0 = make_uint4(0, 0, 0, 0), myd1 = myd0;
            myQ.valid = 0;
            if (need) {
                myQ = load_query<MODE>(P, qo + q);
                if (myQ.valid) {
                    const uint4* qd = (const uint4*)(P.qdesc + (qo + q) * 8);
                    myd0 = __ldg(qd); myd1 = __ldg(qd + 1);
                } else if (st_best[q] & 0xffffu) { st_best[q] = 0; st_top[q] = MB_NONE; s_flag[par] = 1; }   // cannot happen: validity is fixed
            }
            // groups of MB_G lanes; the queries of this pass that need work are dealt to the groups in turn (the i-th of them
            // to group i % (32 / MB_G)), so that after round 0, when only a few of the 32 do, the warp makes one or two trips
            // instead of MB_G nearly empty ones
            const bool go = need && myQ.valid;
            const unsigned gomask = __ballot_sync(0xffffffffu, go);
            if (!gomask) continue;
            const int ngo = __popc(gomask);
            const int sub = lane & (MB_G - 1), g0 = lane & ~(MB_G - 1);
            const unsigned gmask = ((1u << MB_G) - 1u) << g0;
            for (int i0 = 0; i0 < ngo; i0 += 32 / MB_G) {
                // lane of this group's query = position of the k-th set bit of gomask (binary search with popc; __fns is a loop)
                int k = i0 + lane / MB_G, src = 0;
                const bool have = k < ngo;
                {
                    unsigned m = gomask;
                    int c = __popc(m & 0xffffu); if (k >= c) { k -= c; src += 16; m >>= 16; }
                    c = __popc(m & 0xffu);       if (k >= c) { k -= c; src += 8; m >>= 8; }
                    c = __popc(m & 0xfu);        if (k >= c) { k -= c; src += 4; m >>= 4; }
                    c = __popc(m & 0x3u);        if (k >= c) { k -= c; src += 2; m >>= 2; }
                    c = (int)(m & 1u);           if (k >= c) src += 1;
                }
                if (!have) src = 0;
                const int qq = base + src;
                WinQ Q;
                Q.valid = have;
                Q.u = __shfl_sync(0xffffffffu, myQ.u, src); Q.v = __shfl_sync(0xffffffffu, myQ.v, src); Q.r = __shfl_sync(0xffffffffu, myQ.r, src);
                Q.ur = __shfl_sync(0xffffffffu, myQ.ur, src); Q.er_max = __shfl_sync(0xffffffffu, myQ.er_max, src);
                Q.min_level = __shfl_sync(0xffffffffu, myQ.min_level, src); Q.max_level = __shfl_sync(0xffffffffu, myQ.max_level, src);
                uint32_t d[8];
                d[0] = __shfl_sync(0xffffffffu, myd0.x, src); d[1] = __shfl_sync(0xffffffffu, myd0.y, src);
                d[2] = __shfl_sync(0xffffffffu, myd0.z, src); d[3] = __shfl_sync(0xffffffffu, myd0.w, src);
                d[4] = __shfl_sync(0xffffffffu, myd1.x, src); d[5] = __shfl_sync(0xffffffffu, myd1.y, src);
                d[6] = __shfl_sync(0xffffffffu, myd1.z, src); d[7] = __shfl_sync(0xffffffffu, myd1.w, src);
                uint32_t k1 = MB_NONE, k2 = MB_NONE;
                bool skipped = false;
                query_top2(P, Q, d, rec, sdesc, cell_start, [&](const int j) { return blk[j] < qq; }, [](int, int) { return false; },
                           use_ur, sub, gmask, k1, k2, skipped);
                if (sub == 0 && Q.valid) {
                    int best = -1;
                    if (k1 != MB_NONE) {
                        const int bestDist = (int)(k1 >> 16);
                        if (MODE == MODE_POINTS) {
                            if (bestDist <= TH_HIGH) {                                                  // :143
                                const int bestLevel = (int)(rec[k1 & 0xffffu].z & 0xffu);
                                int bestDist2 = 256, bestLevel2 = -1;
                                if (k2 != MB_NONE) { bestDist2 = (int)(k2 >> 16); bestLevel2 = (int)(rec[k2 & 0xffffu].z & 0xffu); }
                                if (!(bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(P.nnratio, (float)bestDist2)))   // :146-149
                                    best = (int)(k1 & 0xffffu);
                            }
                        } else if (bestDist <= P.th_accept) best = (int)(k1 & 0xffffu);                // :256 / :381 / :530
                    }
                    const uint32_t nb = (uint32_t)(best + 1) | (skipped ? 0x10000u : 0u);
                    if ((st_best[qq] & 0xffffu) != (uint32_t)(best + 1)) s_flag[par] = 1;
                    st_best[qq] = nb;
                    st_top[qq] = (k1 == MB_NONE ? 0xffffu : (k1 & 0xffffu)) | ((k2 == MB_NONE ? 0xffffu : (k2 & 0xffffu)) << 16);
                }
            }
        }
        bool changed;
        if (CL > 1) {
            cg::this_cluster().sync();                       // every block's decisions of this round are in place
            changed = false;
#pragma unroll
            for (int r = 0; r < CL; ++r) changed |= cg::this_cluster().map_shared_rank(s_flag, r)[par] != 0;
        } else {
            __syncthreads();
            changed = s_flag[par] != 0;
        }
        if (round < 8) CLK(4 + round);
        if (!changed) break;
        // blocker[] from the decisions, into the other buffer; a keypoint whose blocker moved to a LATER query (or
        // went away) is the only thing that can change the result of a query that skipped taken candidates
        int* const nblk = blk == blk_a ? blk_b : blk_a;
        for (int j = tid; j < nvalid; j += MB_NT) nblk[j] = (rec[j].z >> 31) ? -1 : INT_MAX;
        if (tid == 0) { s_flag[par ^ 1] = 0; s_flag[2] = 0; }
        __syncthreads();
        for (int q = tid; q < nq; q += MB_NT) {
            const uint32_t sb = CL > 1 ? cg::this_cluster().map_shared_rank(st_best, owner(q))[q] : st_best[q];
            const int b = (int)(sb & 0xffffu) - 1;
            if (b >= 0 && (qobs ? qobs[q] : 1) > 0) atomicMin(&nblk[b], q);
        }
        __syncthreads();
        bool freed = false;
        for (int j = tid; j < nvalid; j += MB_NT) freed |= nblk[j] > blk[j];
        if (freed) s_flag[2] = 1;
        blk = nblk;
        if (CL > 1) cg::this_cluster().sync();               // nobody still reads the decisions the next round overwrites
        else __syncthreads();
    }
    if (CL > 1) {
        // block 0 collects every query's decision and finishes alone; the others stay until it has read them
        if (crank == 0)
            for (int q = tid; q < nq; q += MB_NT) {
                const int o = owner(q);
                if (o != 0) st_best[q] = cg::this_cluster().map_shared_rank(st_best, o)[q];
            }
        cg::this_cluster().sync();
        if (crank != 0) return;
    }
    if (P.rounds && tid == 0) P.rounds[prob] = round + 1;

    // ---- results.  Last writer of a keypoint = its largest claimant (every claimant precedes the blocker or is it).
    for (int k = tid; k < n; k += MB_NT) {
        int a = -1;
        if (init) a = MODE == MODE_POINTS ? init[k] : (init[k] >= 0 ? -2 : -1);
        assign_out[k] = a;
    }
    for (int j = tid; j < nvalid; j += MB_NT) blk[j] = -1;
    __syncthreads();
    int mine = 0;
    for (int q = tid; q < nq; q += MB_NT) {
        const int b = (int)(st_best[q] & 0xffffu) - 1;
        if (b < 0) continue;
        ++mine;
        atomicMax(&blk[b], q);
        if (MODE == MODE_BEST && P.check_ori) {
            const int bin = mb_rot_bin(P.q_angle[qo + q], kps[rec[b].z >> 8 & 0x7fffffu].angle);
            st_top[q] = (uint32_t)bin;
            atomicAdd(&s_sizes[bin], 1);
        }
    }
    if (mine) atomicAdd(&s_cnt[0], mine);
    __syncthreads();
    if (MODE == MODE_BEST && P.check_ori) {
        if (tid == 0) {   // ComputeThreeMaxima, src/ORBmatcher.cc:1663-1707
            int ind1, ind2, ind3;
            orb_three_maxima(s_sizes, ind1, ind2, ind3);
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
        // every entry of a rejected bin clears its keypoint and decrements, duplicates included (:286-296)
        int dec = 0;
        for (int q = tid; q < nq; q += MB_NT) {
            const int b = (int)(st_best[q] & 0xffffu) - 1;
            if (b < 0) continue;
            const int bin = (int)st_top[q];
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { blk[b] = -3; ++dec; }
        }
        if (dec) atomicAdd(&s_cnt[1], dec);
        __syncthreads();
    }
    for (int j = tid; j < nvalid; j += MB_NT) {
        const int w = blk[j];
        if (w >= 0) assign_out[rec[j].z >> 8 & 0x7fffffu] = w;
        else if (w == -3) assign_out[rec[j].z >> 8 & 0x7fffffu] = -1;
    }
    if (tid == 0) *nm_out = s_cnt[0] - s_cnt[1];
    CLK(12);
}

// ------------------------------------------------------------------------------------------ monocular initialisation
// ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:1055-1180, one block per (F1, F2) pair.  Queries = the level-0
// keypoints of F1 in order; candidates = level-0 keypoints of F2 inside the window around vbPrevMatched.  A candidate is
// skipped when an EARLIER query already matched it at a distance <= this one's (:1094); an accepted match replaces the
// previous owner (:1115-1122).  The same fixpoint idea as above with a richer state: per F2 keypoint the list of its
// claimants (query, distance).  "Beaten when query q runs" <=> some claimant q' < q has distance <= dist.  Every round
// recomputes all queries against the lists of the previous round (after round r the first r decisions are final); a
// keypoint with more than MB_CL claimants in some round switches the block to the plain in-order walk by one warp.
template <int LOC>
__global__ void __launch_bounds__(MB_NT) k_init_fixpoint(const __grid_constant__ MbParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_flag[3], s_cnt[2], s_sizes[HISTO_LENGTH], s_ind[3];
    const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MB_NT >> 5;
    const int n = P.n[prob], nq = P.nq[prob];
    int* const nm_out = P.nmatches + prob;
    if (n < 0 || n > P.sn_max || n > P.kp_stride || nq < 0 || nq > P.nq_stride || nq > MB_MAX_KP) { if (tid == 0) *nm_out = -1; return; }
    const size_t ko = (size_t)prob * P.kp_stride, qo = (size_t)prob * P.nq_stride;
    const orbx_kp* kps = P.kps + ko;
    const orbx_kp* kps1 = P.kps1 + qo;
    float* prev = P.prev + 2 * qo;
    int* matches12 = P.matches12 + qo;
    uint32_t* cl = P.cl + ko * MB_CL;
    int sn = 32; while (sn < n) sn <<= 1;
    uint32_t* keys = smem;
    int* cell_start = (int*)(smem + P.sn_max);
    int* const cnt = cell_start + GRID_CELLS + 4;           // [sn_max] claimants of every position; last writer at the end
    int* const md = cnt + P.sn_max;                         // [sn_max] vMatchedDistance of the in-order walk
    uint32_t* st_dist = (uint32_t*)(md + P.sn_max);         // [nq_max] accepted distance (| bin << 16 at the end)
    uint32_t* st_best = st_dist + P.nq_max;                 // [nq_max] accepted position + 1, 0 = none
    uint4* const sm16 = (uint4*)(st_best + P.nq_max);
    uint4* rec; uint4* sdesc;
    if (LOC == 0) { rec = sm16; sdesc = sm16 + P.sn_max; }
    else if (LOC == 1) { rec = sm16; sdesc = (uint4*)(P.sdesc + ko * 8); }
    else { rec = P.rec + ko; sdesc = (uint4*)(P.sdesc + ko * 8); }

    mb_sort_frame<MB_NT>(P, kps, n, sn, keys, cell_start, cnt);
    const int nvalid = cell_start[GRID_CELLS];
    for (int j = tid; j < nvalid; j += MB_NT) {
        const int idx = (int)(keys[j] & 0xffffu);
        const orbx_kp kp = kps[idx];
        rec[j] = make_uint4(__float_as_uint(kp.x), __float_as_uint(kp.y), (uint32_t)(kp.octave & 0xff) | ((uint32_t)idx << 8), __float_as_uint(-1.0f));
        cnt[j] = 0;
    }
    for (int t = tid; t < nvalid * 8; t += MB_NT) {
        const int j = t >> 3, wd = t & 7;
        ((uint32_t*)sdesc)[(size_t)j * 8 + wd] = P.desc[(ko + (keys[j] & 0xffffu)) * 8 + wd];
    }
    for (int q = tid; q < nq; q += MB_NT) { st_dist[q] = 0; st_best[q] = 0; }
    if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
    if (tid < HISTO_LENGTH) s_sizes[tid] = 0;
    __syncthreads();

    auto load_q = [&](const int q) {                         // :1071-1080
        WinQ Q;
        const orbx_kp k1 = kps1[q];
        Q.valid = !(k1.octave > 0);
        Q.u = prev[2 * q]; Q.v = prev[2 * q + 1]; Q.r = P.window;
        Q.min_level = k1.octave; Q.max_level = k1.octave;
        Q.ur = 0.f; Q.er_max = 0.f;
        return Q;
    };
    auto decide = [&](const int q, const uint32_t k1, const uint32_t k2) -> bool {      // :1107-1113; returns "changed"
        uint32_t nb = 0, nd = 0;
        if (k1 != MB_NONE) {
            const int bestDist = (int)(k1 >> 16);
            const float second = k2 == MB_NONE ? 2147483648.0f : (float)(int)(k2 >> 16);   // INT_MAX when there is none
            if (bestDist <= TH_LOW && (float)bestDist < __fmul_rn(second, P.nnratio)) { nb = (k1 & 0xffffu) + 1; nd = (uint32_t)bestDist; }
        }
        const bool changed = st_best[q] != nb || st_dist[q] != nd;
        st_best[q] = nb; st_dist[q] = nd;
        return changed;
    };

    int round = 0;
    bool in_order = false;
    for (;; ++round) {
        const int par = round & 1;
        for (int base = warp * 32; base < nq; base += nwarps * 32) {
            const int q = base + lane;
            WinQ myQ;
            myQ.valid = 0;
            uint4 myd0 = make_uint4(0, 0, 0, 0), myd1 = myd0;
            if (q < nq) {
                myQ = load_q(q);
                if (myQ.valid) { const uint4* qd = (const uint4*)(P.qdesc + (qo + q) * 8); myd0 = __ldg(qd); myd1 = __ldg(qd + 1); }
            }
            const bool go = myQ.valid != 0;
            if (!__any_sync(0xffffffffu, go)) continue;
            const int sub = lane & (MB_G - 1), g0 = lane & ~(MB_G - 1);
            const unsigned gmask = ((1u << MB_G) - 1u) << g0;
            for (int i = 0; i < MB_G; ++i) {
                const int src = g0 + i, qq = base + src;
                WinQ Q;
                Q.valid = __shfl_sync(0xffffffffu, (int)go, src);
                if (!__any_sync(0xffffffffu, Q.valid)) continue;
                Q.u = __shfl_sync(0xffffffffu, myQ.u, src); Q.v = __shfl_sync(0xffffffffu, myQ.v, src); Q.r = P.window;
                Q.ur = 0.f; Q.er_max = 0.f;
                Q.min_level = __shfl_sync(0xffffffffu, myQ.min_level, src); Q.max_level = Q.min_level;
                uint32_t d[8];
                d[0] = __shfl_sync(0xffffffffu, myd0.x, src); d[1] = __shfl_sync(0xffffffffu, myd0.y, src);
                d[2] = __shfl_sync(0xffffffffu, myd0.z, src); d[3] = __shfl_sync(0xffffffffu, myd0.w, src);
                d[4] = __shfl_sync(0xffffffffu, myd1.x, src); d[5] = __shfl_sync(0xffffffffu, myd1.y, src);
                d[6] = __shfl_sync(0xffffffffu, myd1.z, src); d[7] = __shfl_sync(0xffffffffu, myd1.w, src);
                uint32_t k1 = MB_NONE, k2 = MB_NONE;
                bool skipped = false;
                query_top2(P, Q, d, rec, sdesc, cell_start, [](int) { return false; },
                           [&](const int j, const int dist) {
                               const int c = min(cnt[j], MB_CL);
                               for (int e = 0; e < c; ++e) {
                                   const uint32_t ent = cl[(size_t)j * MB_CL + e];
                                   if ((int)(ent >> 9) < qq && (int)(ent & 0x1ffu) <= dist) return true;
                               }
                               return false;
                           }, false, sub, gmask, k1, k2, skipped);
                if (sub == 0 && Q.valid && decide(qq, k1, k2)) s_flag[par] = 1;
            }
        }
        __syncthreads();
        if (!s_flag[par]) break;
        // claimant lists from the decisions
        for (int j = tid; j < nvalid; j += MB_NT) cnt[j] = 0;
        if (tid == 0) s_flag[par ^ 1] = 0;
        __syncthreads();
        for (int q = tid; q < nq; q += MB_NT) {
            const int b = (int)st_best[q] - 1;
            if (b < 0) continue;
            const int e = atomicAdd(&cnt[b], 1);
            if (e < MB_CL) cl[(size_t)b * MB_CL + e] = ((uint32_t)q << 9) | st_dist[q];
            else s_flag[2] = 1;
        }
        __syncthreads();
        if (s_flag[2] || round > nq) { in_order = true; break; }
    }
    if (in_order) {
        // the reference's own order on one warp (group 0 works, the other lanes only take part in the shuffles)
        for (int j = tid; j < nvalid; j += MB_NT) md[j] = INT_MAX;
        for (int q = tid; q < nq; q += MB_NT) { st_dist[q] = 0; st_best[q] = 0; }
        __syncthreads();
        if (warp == 0) {
            const int sub = lane & (MB_G - 1);
            const unsigned gmask = ((1u << MB_G) - 1u) << (lane & ~(MB_G - 1));
            for (int q = 0; q < nq; ++q) {
                WinQ Q = load_q(q);
                const bool valid = Q.valid != 0;
                Q.valid = valid && lane < MB_G;
                uint32_t d[8];
#pragma unroll
                for (int w = 0; w < 8; ++w) d[w] = __ldg(P.qdesc + (qo + q) * 8 + w);
                uint32_t k1 = MB_NONE, k2 = MB_NONE;
                bool skipped = false;
                query_top2(P, Q, d, rec, sdesc, cell_start, [](int) { return false; }, [&](const int j, const int dist) { return md[j] <= dist; },
                           false, sub, gmask, k1, k2, skipped);
                if (lane == 0 && valid) { decide(q, k1, k2); if (st_best[q]) md[st_best[q] - 1] = (int)st_dist[q]; }
                __syncwarp();
            }
        }
        __syncthreads();
    }
    if (P.rounds && tid == 0) P.rounds[prob] = in_order ? -(round + 1) : round + 1;

    // ---- results: the owner of a keypoint is its last claimant; replaced queries keep their histogram entry (:1124-1134)
    for (int j = tid; j < nvalid; j += MB_NT) cnt[j] = -1;
    __syncthreads();
    for (int q = tid; q < nq; q += MB_NT) {
        const int b = (int)st_best[q] - 1;
        if (b < 0) continue;
        atomicMax(&cnt[b], q);
        if (P.check_ori) {
            const int bin = mb_rot_bin(kps1[q].angle, kps[rec[b].z >> 8 & 0x7fffffu].angle);
            st_dist[q] |= (uint32_t)bin << 16;
            atomicAdd(&s_sizes[bin], 1);
        }
    }
    __syncthreads();
    int owned = 0;
    for (int j = tid; j < nvalid; j += MB_NT) owned += cnt[j] >= 0 ? 1 : 0;
    if (owned) atomicAdd(&s_cnt[0], owned);
    if (P.check_ori && tid == 0) {   // ComputeThreeMaxima, src/ORBmatcher.cc:1663-1707
        int ind1, ind2, ind3;
        orb_three_maxima(s_sizes, ind1, ind2, ind3);
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int dec = 0;
    for (int q = tid; q < nq; q += MB_NT) {
        int m = -1;
        const int b = (int)st_best[q] - 1;
        if (b >= 0 && cnt[b] == q) {                                          // still the owner
            m = (int)(rec[b].z >> 8 & 0x7fffffu);
            if (P.check_ori) {
                const int bin = (int)(st_dist[q] >> 16);
                if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { m = -1; ++dec; }     // :1163-1167
            }
        }
        matches12[q] = m;
        if (m >= 0) { prev[2 * q] = kps[m].x; prev[2 * q + 1] = kps[m].y; }  // :1175-1177
    }
    if (dec) atomicAdd(&s_cnt[1], dec);
    __syncthreads();
    if (tid == 0) *nm_out = s_cnt[0] - s_cnt[1];
}

// ------------------------------------------------------------------------------------------ order-free searches
// The inner search that ORBmatcher::Fuse (src/ORBmatcher.cc:1364-1513 and :1516-1633) and ORBmatcher::SearchBySim3
// (:836-1052) run for every projected map point: the closest descriptor among the keypoints KeyFrame::GetFeaturesInArea
// (src/KeyFrame.cc:637-676) returns for the window, at the predicted level or the one below (:1436-1437, :907-908,
// :1593-1594), first candidate on ties (strict <, :1478), accepted at TH_LOW (Fuse) or TH_HIGH (SearchBySim3).  Fuse(pKF,
// vpMapPoints, th) also drops candidates whose reprojection error fails the chi-square test (:1440-1469): three degrees of
// freedom (7.8) when the keypoint has a right coordinate (mvuRight >= 0), two (5.99) otherwise.
// Unlike the searches above no query depends on another one -- what the reference does with a match (Replace,
// AddObservation, the mutual check of SearchBySim3) happens after the search and stays with the caller -- so there is no
// order to resolve: P.split blocks share a problem, each builds the position-ordered frame of k_match_fixpoint for itself
// (redundant, but a single Fuse call then uses split SMs instead of one) and takes every split-th group of queries, 4 lanes
// per query; per query the smallest (distance << 16 | position) key is the reference's answer.
// LOC 0: records and descriptors in shared memory, 1: descriptors in the global workspace (frames above ~2700 keypoints).
template <int LOC>
__global__ void __launch_bounds__(MB_NT) k_window_best_free(const __grid_constant__ MbParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MB_NT >> 5;
    const int prob = blockIdx.x / P.split, part = blockIdx.x % P.split;
    const int n = P.n[prob], nq = P.nq[prob];
    if (n < 0 || n > P.sn_max || n > P.kp_stride || nq < 0 || nq > P.nq_stride) { if (tid == 0 && part == 0) P.nmatches[prob] = -1; return; }
    const size_t ko = (size_t)prob * P.kp_stride, qo = (size_t)prob * P.nq_stride;
    const orbx_kp* kps = P.kps + ko;
    int sn = 32; while (sn < n) sn <<= 1;
    uint32_t* keys = smem;                                  // [sn_max]
    int* cell_start = (int*)(smem + P.sn_max);              // [GRID_CELLS + 1] (+3 pad)
    int* tmp = cell_start + GRID_CELLS + 4;                 // [sn_max]
    uint4* rec = (uint4*)(tmp + P.sn_max);                  // [sn_max]
    uint4* sdesc = LOC == 0 ? rec + P.sn_max : (uint4*)(P.sdesc + ((size_t)blockIdx.x * P.kp_stride) * 8);
    mb_sort_frame<MB_NT>(P, kps, n, sn, keys, cell_start, tmp);
    const int nvalid = cell_start[GRID_CELLS];
    for (int j = tid; j < nvalid; j += MB_NT) {
        const int idx = (int)(keys[j] & 0xffffu);
        ORB_CHECK(idx >= 0 && idx < n && j < P.sn_max);
        const orbx_kp kp = kps[idx];
        rec[j] = make_uint4(__float_as_uint(kp.x), __float_as_uint(kp.y), (uint32_t)(kp.octave & 0xff) | ((uint32_t)idx << 8),
                            __float_as_uint(P.u_right ? P.u_right[ko + idx] : -1.0f));
    }
    for (int t = tid; t < nvalid * 8; t += MB_NT) {
        const int j = t >> 3, wd = t & 7;
        ((uint32_t*)sdesc)[(size_t)j * 8 + wd] = P.desc[(ko + (keys[j] & 0xffffu)) * 8 + wd];
    }
    __syncthreads();
    constexpr int QW = 32 / MB_G;
    const int sub = lane & (MB_G - 1);
    int found = 0;
    // a group's query data (a dozen words from global memory) is fetched one trip ahead of the window walk that uses it
    struct QD { float u, v, r, ur; int level; bool live; uint4 q0, q1; };
    auto fetch = [&](const int q) {
        QD d;
        d.live = q < nq && (!P.b0 || P.b0[qo + q]);
        d.u = d.v = d.r = d.ur = 0.f; d.level = 0; d.q0 = d.q1 = make_uint4(0, 0, 0, 0);
        if (d.live) {
            d.u = P.f0[3 * (qo + q)]; d.v = P.f0[3 * (qo + q) + 1]; d.r = P.f0[3 * (qo + q) + 2];
            d.level = P.i0[qo + q];
            d.ur = P.f1 ? P.f1[qo + q] : 0.f;
            d.q0 = __ldg((const uint4*)(P.qdesc + (qo + q) * 8)); d.q1 = __ldg((const uint4*)(P.qdesc + (qo + q) * 8) + 1);
        }
        return d;
    };
    const int step = P.split * nwarps * QW;
    int base = (part * nwarps + warp) * QW;
    QD nxt = fetch(base + lane / MB_G);
    for (; base < nq; base += step) {
        const int q = base + lane / MB_G;
        const QD cur = nxt;
        nxt = fetch(q + step);
        const float u = cur.u, v = cur.v, r = cur.r, ur = cur.ur;
        const int level = cur.level;
        const uint4 q0 = cur.q0, q1 = cur.q1;
        uint32_t best = MB_NONE;
        if (cur.live) {
            // src/KeyFrame.cc:642-656
            const int nMinCellX = max(0, (int)floorf((u - P.min_x - r) * P.inv_w));
            const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf((u - P.min_x + r) * P.inv_w));
            const int nMinCellY = max(0, (int)floorf((v - P.min_y - r) * P.inv_h));
            const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf((v - P.min_y + r) * P.inv_h));
            if (!(nMinCellX >= GRID_COLS || nMaxCellX < 0 || nMinCellY >= GRID_ROWS || nMaxCellY < 0))
                for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
                    ORB_CHECK(ix >= 0 && ix < GRID_COLS && nMinCellY >= 0 && nMaxCellY < GRID_ROWS);
                    const int s0 = cell_start[ix * GRID_ROWS + nMinCellY], s1 = cell_start[ix * GRID_ROWS + nMaxCellY + 1];
                    ORB_CHECK(s0 >= 0 && s0 <= s1 && s1 <= nvalid);
                    for (int j = s0 + sub; j < s1; j += MB_G) {
                        const uint4 rc = rec[j];
                        ORB_CHECK((int)(rc.z & 0xffu) < MB_MAX_LEVELS || !P.chi2);
                        const int oct = (int)(rc.z & 0xffu);
                        if (oct < level - 1 || oct > level) continue;                                        // :1436-1437
                        const float ex = __fsub_rn(u, __uint_as_float(rc.x)), ey = __fsub_rn(v, __uint_as_float(rc.y));
                        if (!(fabsf(ex) < r && fabsf(ey) < r)) continue;                                      // KeyFrame.cc:668
                        if (P.chi2) {
                            float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                            const float kr = __uint_as_float(rc.w);
                            if (kr >= 0) {                                                                    // :1442-1455
                                const float er = __fsub_rn(ur, kr);
                                e2 = __fadd_rn(e2, __fmul_rn(er, er));
                                if ((double)__fmul_rn(e2, P.scale[oct]) > 7.8) continue;
                            } else if ((double)__fmul_rn(e2, P.scale[oct]) > 5.99) continue;                  // :1457-1469
                        }
                        const uint4 b0 = sdesc[2 * j], b1 = sdesc[2 * j + 1];
                        const uint32_t dist = __popc(q0.x ^ b0.x) + __popc(q0.y ^ b0.y) + __popc(q0.z ^ b0.z) + __popc(q0.w ^ b0.w) +
                                              __popc(q1.x ^ b1.x) + __popc(q1.y ^ b1.y) + __popc(q1.z ^ b1.z) + __popc(q1.w ^ b1.w);
                        best = min(best, (dist << 16) | (uint32_t)j);
                    }
                }
        }
#pragma unroll
        for (int d = MB_G / 2; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
        if (sub == 0 && q < nq) {
            const int dist = best == MB_NONE ? 256 : (int)(best >> 16);
            const bool ok = best != MB_NONE && dist <= P.th_accept;                                           // :1483 / :920 / :1607
            ORB_CHECK(best == MB_NONE || (int)(best & 0xffffu) < nvalid);
            ORB_CHECK(!ok || (int)(rec[best & 0xffffu].z >> 8) < n);
            P.assign_out[qo + q] = ok ? (int)(rec[best & 0xffffu].z >> 8) : -1;
            P.best_dist[qo + q] = dist;
            found += ok;
        }
    }
    found = __reduce_add_sync(0xffffffffu, found);
    if (lane == 0 && found) atomicAdd(P.nmatches + prob, found);
}

// ================================================================================ host side
namespace {
struct DevGuard {   // restores the caller's current device
    int prev = -1;
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int device_of(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    return a.device;
}

size_t base_smem(int kp_stride, int nq_stride, int* sn_out, int* nq_out)
{
    int sn = 32; while (sn < kp_stride && sn < MB_MAX_KP) sn <<= 1;
    const int nqm = (nq_stride + 3) & ~3;
    if (sn_out) *sn_out = sn;
    if (nq_out) *nq_out = nqm;
    return ((size_t)sn * 3 + GRID_CELLS + 4 + (size_t)nqm * 2) * 4;
}
const size_t kSmemMax = 224 * 1024;

template <int MODE, int LOC>
cudaError_t launch_loc(const MbParams& P, int nprob, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaSuccess;
    // a handful of problems with everything in shared memory: one thread-block cluster per problem
    static const bool cluster_on = []{ const char* v = getenv("ORB_MATCH_CLUSTER"); return !(v && v[0] == '0'); }();
    if (LOC == 0 && MB_CLUSTER > 1 && cluster_on && nprob * MB_CLUSTER <= 144) {
        // a constant: the attribute is per-function state shared by all host threads (a per-launch value races)
        e = cudaFuncSetAttribute(k_match_fixpoint<MODE, 0, MB_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(nprob * MB_CLUSTER)); cfg.blockDim = dim3(MB_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = MB_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, k_match_fixpoint<MODE, 0, MB_CLUSTER>, P);
            if (e == cudaSuccess) return cudaSuccess;
        }
        cudaGetLastError();                                  // the cluster could not be placed: one block per problem instead
        e = cudaSuccess;
    }
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_match_fixpoint<MODE, LOC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return e;
    k_match_fixpoint<MODE, LOC, 1><<<nprob, MB_NT, smem, st>>>(P);
    return cudaGetLastError();
}

template <int MODE>
int launch(MbParams& P, int nprob, cudaStream_t st)
{
    const int dev = device_of(P.kps);
    if (dev < 0 || device_of(P.assign_out) != dev || device_of(P.nmatches) != dev || device_of(P.qdesc) != dev) return ORBX_E_ARG;
    DevGuard g;
    if (cudaGetDevice(&g.prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    size_t smem = base_smem(P.n_bound, P.nq_stride, &P.sn_max, &P.nq_max);
    const size_t smem_max = kSmemMax;
    if (smem > smem_max) return ORBX_E_ARG;
    P.rec_in_smem = smem + (size_t)P.sn_max * 16 <= smem_max;
    if (P.rec_in_smem) smem += (size_t)P.sn_max * 16;
    P.desc_in_smem = P.rec_in_smem && smem + (size_t)P.sn_max * 32 <= smem_max;
    // (measured at 2000 x 2000: everything in shared memory 0.25 ms per 296 problems, descriptors in global memory
    // 0.30 ms, records too 0.34 ms -- although the smaller footprints would let two blocks share an SM)
    if (P.desc_in_smem) smem += (size_t)P.sn_max * 32;
    void* ws = nullptr;
    const size_t rec_bytes = P.rec_in_smem ? 0 : (size_t)nprob * P.kp_stride * sizeof(uint4);
    const size_t ws_bytes = rec_bytes + (P.desc_in_smem ? 0 : (size_t)nprob * P.kp_stride * 32) + 16;
    { const cudaError_t em = cudaMallocAsync(&ws, ws_bytes, st);
      if (em != cudaSuccess) { if (getenv("ORB_B200_DEBUG")) fprintf(stderr, "orb_b200: cudaMallocAsync: %s\n", cudaGetErrorString(em)); cudaGetLastError(); return ORBX_E_CUDA; } }
    P.rec = (uint4*)ws;
    P.sdesc = (uint32_t*)((char*)ws + rec_bytes);
    const cudaError_t e = P.desc_in_smem ? launch_loc<MODE, 0>(P, nprob, smem, st)
                        : P.rec_in_smem ? launch_loc<MODE, 1>(P, nprob, smem, st) : launch_loc<MODE, 2>(P, nprob, smem, st);
    cudaFreeAsync(ws, st);
    if (e != cudaSuccess) { if (getenv("ORB_B200_DEBUG")) fprintf(stderr, "orb_b200: k_match_fixpoint launch: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return ORBX_E_CUDA; }
    return ORBX_OK;
}

bool fill_frames(MbParams& P, const orbm_frames* F)
{
    if (!F || F->nprob <= 0 || !F->kps || !F->desc || !F->n || F->kp_stride <= 0) return false;
    if (((uintptr_t)F->desc & 15) || ((uintptr_t)F->kps & 3)) return false;
    P.kps = F->kps; P.desc = (const uint32_t*)F->desc; P.u_right = F->u_right; P.n = F->n; P.kp_stride = F->kp_stride;
    P.n_bound = (F->max_n > 0 && F->max_n < F->kp_stride) ? F->max_n : F->kp_stride;
    P.min_x = F->min_x; P.min_y = F->min_y; P.max_x = F->max_x; P.max_y = F->max_y;
    P.inv_w = (float)GRID_COLS / (F->max_x - F->min_x);     // src/Frame.cc:108
    P.inv_h = (float)GRID_ROWS / (F->max_y - F->min_y);     // :109
    return true;
}
} // namespace

// Internal (orb_match.cu): can one problem of this size run in the block-per-problem kernel?
bool orb_match_batch_fits(int kp_stride, int nq_stride)
{
    return kp_stride > 0 && nq_stride > 0 && nq_stride <= MB_MAX_KP * 4 && base_smem(kp_stride, nq_stride, nullptr, nullptr) <= kSmemMax;
}

extern "C" {

#ifdef ORB_MATCH_CLOCKS
int orbm_debug_clocks(long long* out) { return cudaMemcpyFromSymbol(out, g_clk, sizeof(long long) * 16) == cudaSuccess ? 0 : 4; }
#endif

int orbm_search_by_projection_points_batch(const orbm_frames* F, const float* scale, int nlevels, const orbm_points* Q,
                                           const int* init_assign, int* assign_out, float th, float nnratio,
                                           int* nmatches, int* rounds, void* cuda_stream)
{
    MbParams P = {};
    if (!fill_frames(P, F) || !scale || nlevels <= 0 || nlevels > MB_MAX_LEVELS || !Q || !Q->nq || Q->nq_stride <= 0 || Q->nq_stride > MB_MAX_KP * 4 ||
        !Q->proj_xyxr || !Q->level || !Q->view_cos || !Q->in_view || !Q->bad || !Q->observations || !Q->qdesc || !assign_out || !nmatches)
        return ORBX_E_ARG;
    if ((uintptr_t)Q->qdesc & 15) return ORBX_E_ARG;
    P.nq = Q->nq; P.nq_stride = Q->nq_stride;
    P.f0 = Q->proj_xyxr; P.i0 = Q->level; P.f1 = Q->view_cos; P.b0 = Q->in_view; P.b1 = Q->bad; P.qobs = Q->observations;
    P.qdesc = (const uint32_t*)Q->qdesc;
    P.init = init_assign; P.assign_out = assign_out; P.nmatches = nmatches; P.rounds = rounds;
    P.th = th; P.nnratio = nnratio; P.nlevels = nlevels;
    for (int i = 0; i < nlevels; ++i) P.scale[i] = scale[i];
    return launch<MODE_POINTS>(P, F->nprob, (cudaStream_t)cuda_stream);
}

int orbm_window_search_best_batch(const orbm_frames* F, const orbm_windows* Q, const int* init_obs, int* assign_out,
                                  int th_accept, int check_ori, int* nmatches, int* rounds, void* cuda_stream)
{
    MbParams P = {};
    if (!fill_frames(P, F) || !Q || !Q->nq || Q->nq_stride <= 0 || Q->nq_stride > MB_MAX_KP * 4 || !Q->uvr || !Q->min_level || !Q->max_level ||
        !Q->qdesc || !assign_out || !nmatches || (check_ori && !Q->q_angle))
        return ORBX_E_ARG;
    if ((uintptr_t)Q->qdesc & 15) return ORBX_E_ARG;
    P.nq = Q->nq; P.nq_stride = Q->nq_stride;
    P.f0 = Q->uvr; P.i0 = Q->min_level; P.i1 = Q->max_level; P.f1 = Q->ur; P.f2 = Q->er_max; P.b0 = Q->valid; P.qobs = Q->q_obs;
    P.qdesc = (const uint32_t*)Q->qdesc; P.q_angle = Q->q_angle;
    P.init = init_obs; P.assign_out = assign_out; P.nmatches = nmatches; P.rounds = rounds;
    P.th_accept = th_accept; P.check_ori = check_ori;
    return launch<MODE_BEST>(P, F->nprob, (cudaStream_t)cuda_stream);
}

int orbm_window_best_free_batch(const orbm_frames* F, const orbm_free_windows* Q, const float* inv_sigma2, int nlevels,
                                int th_accept, int* best_idx, int* best_dist, int* nfound, void* cuda_stream)
{
    MbParams P = {};
    if (!fill_frames(P, F) || !Q || !Q->nq || Q->nq_stride <= 0 || !Q->uvr || !Q->level || !Q->qdesc || !best_idx || !best_dist || !nfound ||
        (inv_sigma2 && (nlevels <= 0 || nlevels > MB_MAX_LEVELS)))
        return ORBX_E_ARG;
    if ((uintptr_t)Q->qdesc & 15) return ORBX_E_ARG;
    const int nprob = F->nprob;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    P.nq = Q->nq; P.nq_stride = Q->nq_stride;
    P.f0 = Q->uvr; P.i0 = Q->level; P.f1 = Q->ur; P.b0 = Q->valid; P.qdesc = (const uint32_t*)Q->qdesc;
    P.assign_out = best_idx; P.best_dist = best_dist; P.nmatches = nfound; P.th_accept = th_accept;
    P.chi2 = inv_sigma2 != nullptr;
    // a keypoint's octave indexes the table: levels the caller did not supply get 0 (every error passes), as no
    // keypoint of the key frame can lie there
    for (int i = 0; i < MB_MAX_LEVELS; ++i) P.scale[i] = (inv_sigma2 && i < nlevels) ? inv_sigma2[i] : 0.f;
    const int dev = device_of(P.kps);
    if (dev < 0 || device_of(best_idx) != dev || device_of(best_dist) != dev || device_of(nfound) != dev || device_of(P.qdesc) != dev) return ORBX_E_ARG;
    DevGuard g;
    if (cudaGetDevice(&g.prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    int sn = 32; while (sn < P.n_bound && sn < MB_MAX_KP) sn <<= 1;
    P.sn_max = sn;
    size_t smem = ((size_t)sn * 2 + GRID_CELLS + 4) * 4 + (size_t)sn * 16;
    const bool desc_in_smem = smem + (size_t)sn * 32 <= kSmemMax;
    if (desc_in_smem) smem += (size_t)sn * 32;
    if (smem > kSmemMax) return ORBX_E_ARG;
    // blocks per problem: enough to cover the GPU when the launch holds a handful of problems, at most one block per
    // trip of 256 queries
    int split = 148 / nprob;
    const int trips = (Q->nq_stride + MB_NT / MB_G - 1) / (MB_NT / MB_G);
    if (split > trips) split = trips;
    if (split > 16) split = 16;
    if (split < 1) split = 1;
    P.split = split;
    void* ws = nullptr;
    if (!desc_in_smem) {
        if (cudaMallocAsync(&ws, (size_t)nprob * split * P.kp_stride * 32 + 16, st) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
        P.sdesc = (uint32_t*)ws;
    }
    cudaError_t e = cudaMemsetAsync(nfound, 0, sizeof(int) * (size_t)nprob, st);
    if (e == cudaSuccess) {
        if (desc_in_smem) {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_window_best_free<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e == cudaSuccess) { k_window_best_free<0><<<nprob * split, MB_NT, smem, st>>>(P); e = cudaGetLastError(); }
        } else {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_window_best_free<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e == cudaSuccess) { k_window_best_free<1><<<nprob * split, MB_NT, smem, st>>>(P); e = cudaGetLastError(); }
        }
    }
    if (ws) cudaFreeAsync(ws, st);
    if (e != cudaSuccess) { if (getenv("ORB_B200_DEBUG")) fprintf(stderr, "orb_b200: k_window_best_free: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return ORBX_E_CUDA; }
    return ORBX_OK;
}

int orbm_search_for_initialization_batch(const orbm_frames* F1, const orbm_frames* F2, float* prev_matched, int* matches12,
                                         int windowSize, float nnratio, int checkOri, int* nmatches, int* rounds, void* cuda_stream)
{
    MbParams P = {};
    if (!fill_frames(P, F2) || !F1 || F1->nprob != F2->nprob || !F1->kps || !F1->desc || !F1->n || F1->kp_stride <= 0 ||
        ((uintptr_t)F1->desc & 15) || !prev_matched || !matches12 || !nmatches)
        return ORBX_E_ARG;
    const int nprob = F2->nprob;
    P.u_right = nullptr;
    P.nq = F1->n; P.nq_stride = F1->kp_stride; P.kps1 = F1->kps; P.qdesc = (const uint32_t*)F1->desc;
    P.prev = prev_matched; P.matches12 = matches12; P.nmatches = nmatches; P.rounds = rounds;
    P.window = (float)windowSize; P.nnratio = nnratio; P.check_ori = checkOri;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int dev = device_of(P.kps);
    if (dev < 0 || device_of(F1->kps) != dev || device_of(prev_matched) != dev || device_of(matches12) != dev || device_of(nmatches) != dev) return ORBX_E_ARG;
    DevGuard g;
    if (cudaGetDevice(&g.prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    size_t smem = base_smem(P.n_bound, P.nq_stride, &P.sn_max, &P.nq_max);
    if (smem > kSmemMax || P.nq_stride > MB_MAX_KP) return ORBX_E_ARG;
    P.rec_in_smem = smem + (size_t)P.sn_max * 16 <= kSmemMax;
    if (P.rec_in_smem) smem += (size_t)P.sn_max * 16;
    P.desc_in_smem = P.rec_in_smem && smem + (size_t)P.sn_max * 32 <= kSmemMax;
    if (P.desc_in_smem) smem += (size_t)P.sn_max * 32;
    const size_t rec_bytes = P.rec_in_smem ? 0 : (size_t)nprob * P.kp_stride * sizeof(uint4);
    const size_t desc_bytes = P.desc_in_smem ? 0 : (size_t)nprob * P.kp_stride * 32;
    const size_t cl_bytes = (size_t)nprob * P.kp_stride * MB_CL * 4;
    char* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, rec_bytes + desc_bytes + cl_bytes + 16, st) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    P.rec = (uint4*)ws; P.sdesc = (uint32_t*)(ws + rec_bytes); P.cl = (uint32_t*)(ws + rec_bytes + desc_bytes);
    cudaError_t e = cudaSuccess;
    if (P.desc_in_smem) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_init_fixpoint<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) { k_init_fixpoint<0><<<nprob, MB_NT, smem, st>>>(P); e = cudaGetLastError(); }
    } else if (P.rec_in_smem) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_init_fixpoint<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) { k_init_fixpoint<1><<<nprob, MB_NT, smem, st>>>(P); e = cudaGetLastError(); }
    } else {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_init_fixpoint<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
        if (e == cudaSuccess) { k_init_fixpoint<2><<<nprob, MB_NT, smem, st>>>(P); e = cudaGetLastError(); }
    }
    cudaFreeAsync(ws, st);
    if (e != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    return ORBX_OK;
}

} // extern "C"
