// orb_match_bow.cu -- vocabulary-guided matching, batched and device resident (sm_100a):
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)      src/ORBmatcher.cc:552-697
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)   src/ORBmatcher.cc:700-832
//   ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, F12, pairs, bOnlyStereo)   src/ORBmatcher.cc:1183-1361  (TRI)
// One thread block per (side 1, side 2) problem.  DBoW2::FeatureVector (a std::map NodeId -> feature indices) arrives
// in CSR form: node_id ascending, node_off, feat.  The reference walks the two maps in lockstep and, inside a shared
// node, takes side-1 features in order; a side-2 feature that an earlier side-1 feature matched is skipped by the
// later ones (:603-604 / :747).  The order of the side-1 features is simply their POSITION in feat1, and the
// candidates of one of them are one contiguous run of feat2, so the same fixpoint over "first claimant of each
// candidate" as in orb_match_batch.cu applies: every claim blocks, blockers only ever move to earlier queries, and a
// query is recomputed only when one of its two best candidates became blocked.  Groups of 4 lanes work on one
// query each (a node holds a handful of features).
#include <cuda_runtime.h>
#include <stdint.h>

#include <climits>

#include "../../include/orb_b200.h"

#define TH_LOW 50        // src/ORBmatcher.cc:38
#include "orb_match_common.cuh"   // HISTO_LENGTH, orb_three_maxima
#define BW_NT 512
#ifndef BW_G
#define BW_G 4
#endif
#define BW_MAX_KP 8192
#define BW_NONE 0xffffffffu

struct BowSide {
    const orbx_kp* kps; const uint32_t* desc; const int* n; int kp_stride;
    const int* node_id; const int* node_off; const int* n_nodes; const int* feat; int node_stride;
    const uint8_t* valid;   // side 1: has a usable map point; side 2: NULL = every feature may be matched
};
struct BowParams {
    BowSide A, B;
    float nnratio; int check_ori, strict;
    int* match12; int* match21; int* nmatches; int* rounds;
    int n1_max, n2_max, desc_in_smem;
    // SearchForTriangulation only
    const float* ur1; const float* ur2;     // mvuRight of the two key frames [nprob][kp_stride] (NULL = monocular)
    const float* F12; const float* epipole; // [nprob][9] row-major, [nprob][2]
    float scale[32], sigma2[32];            // pKF2->mvScaleFactors, mvLevelSigma2
};

__device__ __forceinline__ int bw_rot_bin(const float a1, const float a2)   // src/ORBmatcher.cc:627-633
{
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / HISTO_LENGTH));
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// TRI = ORBmatcher::SearchForTriangulation: the same walk over shared vocabulary nodes, but (1) a candidate must lie within
// TH_LOW, away from the epipole when neither feature is a stereo one (:1272-1279) and on the epipolar line of the side-1
// feature (CheckDistEpipolarLine, :1636-1650); (2) of equally close candidates the LAST one wins (`dist>bestDist` continues,
// :1260); (3) the reference tests vbMatched2 (:1254) but never sets it, so NO feature depends on another one: one round, no
// blockers, several side-1 features may end up with the same side-2 feature.  The keypoints of the few candidates that pass
// the distance test are read from global memory.
template <bool DSM, bool TRI>    // DSM: side-2 descriptors staged in shared memory (LDS in the candidate loop) or read from global memory
__global__ void __launch_bounds__(BW_NT) k_bow_fixpoint(const __grid_constant__ BowParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_flag[2], s_cnt[2], s_sizes[HISTO_LENGTH], s_ind[3];
    const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BW_NT >> 5;
    const int n1 = P.A.n[prob], n2 = P.B.n[prob], nn1 = P.A.n_nodes[prob], nn2 = P.B.n_nodes[prob];
    int* const nm_out = P.nmatches + prob;
    if (n1 < 0 || n1 > P.A.kp_stride || n1 > P.n1_max || n2 < 0 || n2 > P.B.kp_stride || n2 > P.n2_max ||
        nn1 < 0 || nn1 > P.A.node_stride || nn2 < 0 || nn2 > P.B.node_stride) { if (tid == 0) *nm_out = -1; return; }
    const size_t ka = (size_t)prob * P.A.kp_stride, kb = (size_t)prob * P.B.kp_stride;
    const int* id1 = P.A.node_id + (size_t)prob * P.A.node_stride; const int* off1 = P.A.node_off + (size_t)prob * (P.A.node_stride + 1);
    const int* id2 = P.B.node_id + (size_t)prob * P.B.node_stride; const int* off2 = P.B.node_off + (size_t)prob * (P.B.node_stride + 1);
    const int* feat1 = P.A.feat + ka; const int* feat2 = P.B.feat + kb;
    const int t1 = nn1 > 0 ? min(off1[nn1], n1) : 0, t2 = nn2 > 0 ? min(off2[nn2], n2) : 0;   // features listed in the vectors
    int* match12 = P.match12 + ka;
    int* match21 = P.match21 ? P.match21 + kb : nullptr;

    int* blk = (int*)smem;                                  // [n2_max] first claimant of every side-2 position
    uint32_t* seg = (uint32_t*)(blk + P.n2_max);            // [n1_max] candidate run of every side-1 position: start | end << 16; 0 = none
    uint32_t* st_top = seg + P.n1_max;                      // [n1_max] k1 position | k2 position << 16
    uint32_t* st_best = st_top + P.n1_max;                  // [n1_max] accepted position + 1 (0 = none); bin afterwards
    uint8_t* use2 = (uint8_t*)(st_best + P.n1_max);         // [n2_max] side-2 position may be matched
    uint4* sdesc = (uint4*)(use2 + P.n2_max);               // [n2_max][2] side-2 descriptors in position order (if they fit)
    const bool dsm = DSM;

    // ---- lockstep walk of the two maps (:576-: equal NodeIds; lower_bound otherwise) = intersection of the id lists
    for (int p = tid; p < t1; p += BW_NT) seg[p] = 0;
    for (int p = tid; p < P.n1_max; p += BW_NT) { st_top[p] = BW_NONE; st_best[p] = 0; }
    __syncthreads();
    for (int a = tid; a < nn1; a += BW_NT) {
        const int id = id1[a];
        int lo = 0, hi = nn2;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (id2[mid] < id) lo = mid + 1; else hi = mid; }
        if (lo < nn2 && id2[lo] == id) {
            const int c0 = min(off2[lo], t2), c1 = min(off2[lo + 1], t2);
            if (c1 > c0) for (int p = off1[a]; p < min(off1[a + 1], t1); ++p) seg[p] = (uint32_t)c0 | ((uint32_t)c1 << 16);
        }
    }
    for (int j = tid; j < t2; j += BW_NT) {
        const int idx = feat2[j];
        use2[j] = (idx >= 0 && idx < n2 && (!P.B.valid || P.B.valid[kb + idx])) ? 1 : 0;
        blk[j] = INT_MAX;
    }
    if (dsm)
        for (int t = tid; t < t2 * 8; t += BW_NT) {
            const int j = t >> 3, wd = t & 7, idx = feat2[j];
            ((uint32_t*)sdesc)[(size_t)j * 8 + wd] = (idx >= 0 && idx < n2) ? P.B.desc[(kb + idx) * 8 + wd] : 0u;
        }
    if (tid == 0) { s_flag[0] = 0; s_flag[1] = 0; s_cnt[0] = 0; s_cnt[1] = 0; }
    if (tid < HISTO_LENGTH) s_sizes[tid] = 0;
    __syncthreads();

    // ---- rounds
    int round = 0;
    for (;; ++round) {
        const int par = round & 1;
        for (int base = warp * 32; base < t1; base += nwarps * 32) {
            const int q = base + lane;
            bool go = false;
            uint32_t myseg = 0;
            uint4 myd0 = make_uint4(0, 0, 0, 0), myd1 = myd0;
            float myla = 0.f, mylb = 0.f, mylc = 0.f;
            int myst1 = 0;
            if (q < t1) {
                myseg = seg[q];
                const int idx1 = feat1[q];
                go = myseg != 0 && idx1 >= 0 && idx1 < n1 && P.A.valid[ka + idx1];             // :587-590 / :733-736 / :1222-1232
                if (TRI && go) {
                    const float* F = P.F12 + (size_t)prob * 9;
                    const float x1 = P.A.kps[ka + idx1].x, y1 = P.A.kps[ka + idx1].y;
                    myla = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[0]), __fmul_rn(y1, F[3])), F[6]);   // :1640-1642
                    mylb = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[1]), __fmul_rn(y1, F[4])), F[7]);
                    mylc = __fadd_rn(__fadd_rn(__fmul_rn(x1, F[2]), __fmul_rn(y1, F[5])), F[8]);
                    myst1 = P.ur1 && P.ur1[ka + idx1] >= 0;
                }
                if (go && round > 0) {
                    const uint32_t top = st_top[q];
                    const uint32_t p1 = top & 0xffffu, p2 = top >> 16;
                    go = (p1 != 0xffffu && blk[p1] < q) || (p2 != 0xffffu && blk[p2] < q);
                }
                if (go) {
                    const uint4* d = (const uint4*)(P.A.desc + (ka + idx1) * 8);
                    myd0 = __ldg(d); myd1 = __ldg(d + 1);
                }
            }
            if (!__any_sync(0xffffffffu, go)) continue;
            const int sub = lane & (BW_G - 1), g0 = lane & ~(BW_G - 1);
            for (int i = 0; i < BW_G; ++i) {
                const int src = g0 + i, qq = base + src;
                const int act = __shfl_sync(0xffffffffu, (int)go, src);
                if (!__any_sync(0xffffffffu, act)) continue;
                const uint32_t sg = __shfl_sync(0xffffffffu, myseg, src);
                const uint32_t d0 = __shfl_sync(0xffffffffu, myd0.x, src), d1 = __shfl_sync(0xffffffffu, myd0.y, src),
                               d2 = __shfl_sync(0xffffffffu, myd0.z, src), d3 = __shfl_sync(0xffffffffu, myd0.w, src),
                               d4 = __shfl_sync(0xffffffffu, myd1.x, src), d5 = __shfl_sync(0xffffffffu, myd1.y, src),
                               d6 = __shfl_sync(0xffffffffu, myd1.z, src), d7 = __shfl_sync(0xffffffffu, myd1.w, src);
                float la = 0.f, lb = 0.f, lc = 0.f, ex = 0.f, ey = 0.f;
                int st1 = 0;
                if (TRI) {
                    la = __shfl_sync(0xffffffffu, myla, src); lb = __shfl_sync(0xffffffffu, mylb, src); lc = __shfl_sync(0xffffffffu, mylc, src);
                    st1 = __shfl_sync(0xffffffffu, myst1, src);
                    ex = P.epipole[2 * prob]; ey = P.epipole[2 * prob + 1];
                }
                uint32_t a1 = BW_NONE, a2 = BW_NONE;
                if (act) {
                    const int c1 = (int)(sg >> 16);
                    for (int j = (int)(sg & 0xffffu) + sub; j < c1; j += BW_G) {
                        if (!use2[j] || (!TRI && blk[j] < qq)) continue;                        // :603-604 / :747-751 / :1254-1262
                        uint4 b0, b1;
                        if (dsm) { b0 = sdesc[2 * j]; b1 = sdesc[2 * j + 1]; }
                        else { const uint4* b = (const uint4*)(P.B.desc + (kb + feat2[j]) * 8); b0 = __ldg(b); b1 = __ldg(b + 1); }
                        const uint32_t dist = __popc(d0 ^ b0.x) + __popc(d1 ^ b0.y) + __popc(d2 ^ b0.z) + __popc(d3 ^ b0.w) +
                                              __popc(d4 ^ b1.x) + __popc(d5 ^ b1.y) + __popc(d6 ^ b1.z) + __popc(d7 ^ b1.w);
                        uint32_t key = (dist << 16) | (uint32_t)j;
                        if (TRI) {
                            if (dist > TH_LOW) continue;                                        // :1268
                            const int idx2 = feat2[j];
                            ORB_CHECK(j >= 0 && j < t2 && idx2 >= 0 && idx2 < n2);
                            const orbx_kp k2 = P.B.kps[kb + idx2];
                            const int oct2 = k2.octave & 31;
                            if (!st1 && !(P.ur2 && P.ur2[kb + idx2] >= 0)) {                     // :1272-1279
                                const float dex = __fsub_rn(ex, k2.x), dey = __fsub_rn(ey, k2.y);
                                if (__fadd_rn(__fmul_rn(dex, dex), __fmul_rn(dey, dey)) < __fmul_rn(100.0f, P.scale[oct2])) continue;
                            }
                            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);   // :1644-1649
                            const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
                            if (den == 0) continue;
                            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                            if (!((double)dsqr < 3.84 * (double)P.sigma2[oct2])) continue;
                            key = (dist << 16) | (0xffffu - (uint32_t)j);                       // ties: the last candidate
                        }
                        if (key < a1) { a2 = a1; a1 = key; }
                        else if (key < a2) a2 = key;
                    }
                }
                // group reductions with xor shuffles (redux.sync with one member mask per group is serialised per mask)
                uint32_t k1 = a1;
#pragma unroll
                for (int d = BW_G / 2; d > 0; d >>= 1) k1 = min(k1, __shfl_xor_sync(0xffffffffu, k1, d));
                if (a1 == k1) a1 = a2;
                uint32_t k2 = a1;
#pragma unroll
                for (int d = BW_G / 2; d > 0; d >>= 1) k2 = min(k2, __shfl_xor_sync(0xffffffffu, k2, d));
                if (sub == 0 && act) {
                    int best = -1;
                    const int bestDist1 = k1 == BW_NONE ? 256 : (int)(k1 >> 16), bestDist2 = k2 == BW_NONE ? 256 : (int)(k2 >> 16);
                    if (TRI) { if (k1 != BW_NONE) best = (int)(0xffffu - (k1 & 0xffffu)); ORB_CHECK(best < t2); }     // :1288
                    else if ((P.strict ? bestDist1 < TH_LOW : bestDist1 <= TH_LOW) &&                             // :618 / :768
                             (float)bestDist1 < __fmul_rn(P.nnratio, (float)bestDist2)) best = (int)(k1 & 0xffffu);   // :620 / :770
                    if (!TRI && st_best[qq] != (uint32_t)(best + 1)) s_flag[par] = 1;
                    st_best[qq] = (uint32_t)(best + 1);
                    st_top[qq] = (k1 == BW_NONE ? 0xffffu : (k1 & 0xffffu)) | ((k2 == BW_NONE ? 0xffffu : (k2 & 0xffffu)) << 16);
                }
            }
        }
        __syncthreads();
        if (!s_flag[par]) break;
        for (int j = tid; j < t2; j += BW_NT) blk[j] = INT_MAX;
        if (tid == 0) s_flag[par ^ 1] = 0;
        __syncthreads();
        for (int q = tid; q < t1; q += BW_NT) {
            const int b = (int)st_best[q] - 1;
            if (b >= 0) atomicMin(&blk[b], q);
        }
        __syncthreads();
    }
    if (P.rounds && tid == 0) P.rounds[prob] = round + 1;

    // ---- results: every accepted query holds its own side-2 feature (claims always block), then the rotation histogram
    for (int k = tid; k < n1; k += BW_NT) match12[k] = -1;
    if (match21) for (int k = tid; k < n2; k += BW_NT) match21[k] = -1;
    __syncthreads();
    int mine = 0;
    for (int q = tid; q < t1; q += BW_NT) {
        const int b = (int)st_best[q] - 1;
        if (b < 0) continue;
        ++mine;
        if (P.check_ori) {
            const int bin = bw_rot_bin(P.A.kps[ka + feat1[q]].angle, P.B.kps[kb + feat2[b]].angle);
            st_top[q] = (uint32_t)bin;
            atomicAdd(&s_sizes[bin], 1);
        }
    }
    if (mine) atomicAdd(&s_cnt[0], mine);
    __syncthreads();
    if (P.check_ori && tid == 0) {   // ComputeThreeMaxima, src/ORBmatcher.cc:1663-1707
        int ind1, ind2, ind3;
            orb_three_maxima(s_sizes, ind1, ind2, ind3);
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int dec = 0;
    for (int q = tid; q < t1; q += BW_NT) {
        const int b = (int)st_best[q] - 1;
        if (b < 0) continue;
        if (P.check_ori) {
            const int bin = (int)st_top[q];
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { ++dec; continue; }     // :678-683 / :822-827
        }
        const int i1 = feat1[q], i2 = feat2[b];
        match12[i1] = i2;
        if (match21) match21[i2] = i1;
    }
    if (dec) atomicAdd(&s_cnt[1], dec);
    __syncthreads();
    if (tid == 0) *nm_out = s_cnt[0] - s_cnt[1];
}

namespace {
int dev_of(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    return a.device;
}
bool fill_side(BowSide& S, const orbm_frames* F, const orbm_featvec* V, const uint8_t* valid)
{
    if (!F || !V || !F->kps || !F->desc || !F->n || F->kp_stride <= 0 || !V->node_id || !V->node_off || !V->n_nodes || !V->feat || V->node_stride <= 0)
        return false;
    if ((uintptr_t)F->desc & 15) return false;
    S.kps = F->kps; S.desc = (const uint32_t*)F->desc; S.n = F->n; S.kp_stride = F->kp_stride;
    S.node_id = V->node_id; S.node_off = V->node_off; S.n_nodes = V->n_nodes; S.feat = V->feat; S.node_stride = V->node_stride;
    S.valid = valid;
    return true;
}
} // namespace

namespace {
template <bool TRI>
int launch_bow(BowParams& P, const orbm_frames* A, const orbm_frames* B, const void* const* same_device, int n_same, cudaStream_t st)
{
    const int dev = dev_of(A->kps);
    if (dev < 0) return ORBX_E_ARG;
    for (int i = 0; i < n_same; ++i) if (dev_of(same_device[i]) != dev) return ORBX_E_ARG;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    const int b1 = (A->max_n > 0 && A->max_n < A->kp_stride) ? A->max_n : A->kp_stride, b2 = (B->max_n > 0 && B->max_n < B->kp_stride) ? B->max_n : B->kp_stride;
    P.n1_max = (b1 < BW_MAX_KP ? b1 : BW_MAX_KP) + 3 & ~3;
    P.n2_max = (b2 < BW_MAX_KP ? b2 : BW_MAX_KP) + 15 & ~15;
    size_t smem = (size_t)P.n2_max * 4 + (size_t)P.n1_max * 12 + (size_t)P.n2_max;
    const size_t smem_max = 224 * 1024;
    int rc = ORBX_OK;
    if (smem > smem_max) rc = ORBX_E_ARG;
    else {
        P.desc_in_smem = smem + (size_t)P.n2_max * 32 <= smem_max;
        if (P.desc_in_smem) smem += (size_t)P.n2_max * 32;
        cudaError_t e = cudaSuccess;
        if (P.desc_in_smem) {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_bow_fixpoint<true, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
            if (e == cudaSuccess) { k_bow_fixpoint<true, TRI><<<A->nprob, BW_NT, smem, st>>>(P); e = cudaGetLastError(); }
        } else {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(k_bow_fixpoint<false, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
            if (e == cudaSuccess) { k_bow_fixpoint<false, TRI><<<A->nprob, BW_NT, smem, st>>>(P); e = cudaGetLastError(); }
        }
        if (e != cudaSuccess) { cudaGetLastError(); rc = ORBX_E_CUDA; }
    }
    cudaSetDevice(prev);
    return rc;
}
} // namespace

extern "C" int orbm_search_by_bow_batch(const orbm_frames* A, const orbm_featvec* VA, const uint8_t* a_valid,
                                        const orbm_frames* B, const orbm_featvec* VB, const uint8_t* b_valid,
                                        int kf_kf, float nnratio, int check_ori, int* match12, int* match21,
                                        int* nmatches, int* rounds, void* cuda_stream)
{
    BowParams P = {};
    if (!fill_side(P.A, A, VA, a_valid) || !fill_side(P.B, B, VB, b_valid) || !a_valid || !match12 || !nmatches || A->nprob <= 0 || A->nprob != B->nprob)
        return ORBX_E_ARG;
    P.nnratio = nnratio; P.check_ori = check_ori; P.strict = kf_kf ? 1 : 0;
    P.match12 = match12; P.match21 = match21; P.nmatches = nmatches; P.rounds = rounds;
    const void* same[] = { B->kps, match12, nmatches, VA->feat, VB->feat };
    return launch_bow<false>(P, A, B, same, 5, (cudaStream_t)cuda_stream);
}

extern "C" int orbm_search_for_triangulation_batch(const orbm_frames* A, const orbm_featvec* VA, const uint8_t* a_valid,
                                                   const orbm_frames* B, const orbm_featvec* VB, const uint8_t* b_valid,
                                                   const float* F12, const float* epipole, const float* scale, const float* sigma2, int nlevels,
                                                   int check_ori, int* match12, int* nmatches, void* cuda_stream)
{
    BowParams P = {};
    if (!fill_side(P.A, A, VA, a_valid) || !fill_side(P.B, B, VB, b_valid) || !a_valid || !b_valid || !match12 || !nmatches || A->nprob <= 0 ||
        A->nprob != B->nprob || !F12 || !epipole || !scale || !sigma2 || nlevels <= 0 || nlevels > 32)
        return ORBX_E_ARG;
    P.check_ori = check_ori;
    P.match12 = match12; P.match21 = nullptr; P.nmatches = nmatches; P.rounds = nullptr;
    P.ur1 = A->u_right; P.ur2 = B->u_right; P.F12 = F12; P.epipole = epipole;
    for (int i = 0; i < 32; ++i) { P.scale[i] = i < nlevels ? scale[i] : 0.f; P.sigma2[i] = i < nlevels ? sigma2[i] : 0.f; }
    const void* same[] = { B->kps, match12, nmatches, VA->feat, VB->feat, F12, epipole };
    return launch_bow<true>(P, A, B, same, 7, (cudaStream_t)cuda_stream);
}
