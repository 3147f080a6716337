// orb_sparse.cu -- the per-keypoint kernels of the extractor, sm_100a:
//   k_octree    ORBextractor::DistributeOctTree + ExtractorNode::DivideNode
//               src/ORBextractor.cc:436-495, 562-792
//   k_describe  IC_Angle / computeOrientation (:78-115), computeOrbDescriptor (:120-161),
//               output assembly of operator() (:1100-1148)
#include "orb_device.cuh"
#include "orb_launch.h"

// =============================================================================== quadtree
// One block per (frame, level).  The reference grows a std::list of nodes; here the list is an
// array in LIST ORDER (node id == list position) that is rebuilt every pass:
//
//   new list = [children of the node processed last (n4,n3,n2,n1, empties dropped)] ...
//              [children of the node processed first] ++ [unsplit nodes in old order]
//
// which is what push_front + erase produce (:650-691).  A full pass processes every node with
// more than one key in list order.  Once size + 3*nToExpand > N the reference switches to
// one-at-a-time splitting of the expandable nodes sorted by (key count, node address) from the
// back (:701-766); with the bump-arena allocator the address order is the creation order,
// which is the reverse of the list position of freshly created nodes, so the processing order
// is (count descending, list position ascending).  The break index is found with a prefix sum
// of the gains.  The retained keypoint of a node is its maximum response, first in the
// reference's candidate order on ties (:773-789); candidates arrive UNORDERED here, so that
// order is rebuilt from the coordinates: (cell row, cell col, y, x).
//
// List size bound: a full pass runs only while size + 3*nToExpand <= N, so it ends with at most
// N nodes (the very first pass: at most 4*nIni); a partial round stops at the first split that
// reaches N, i.e. at most N+2.  plan.lv[l].kp_cap = max(N+3, 4*nIni).
#ifndef ORB_OCT_NT
#define ORB_OCT_NT 256         // threads per block for batches (throughput: 128 / 256 / 512 / 1024 threads = 1.21 / 0.89 / 1.08 / 2.28 ms per 1024 frames)
#endif
#ifndef ORB_OCT_NT_SMALL
#define ORB_OCT_NT_SMALL 1024  // ... and for a handful of frames, where one block per level is all the parallelism there is (latency)
#endif
#ifndef ORB_OCT_NT_MID
#define ORB_OCT_NT_MID 512     // ... and for launches whose blocks fit the GPU in one wave (148 SMs x 4 blocks of 512 threads)
#endif
#ifndef ORB_OCT_MID_BLOCKS
#define ORB_OCT_MID_BLOCKS 592
#endif
#ifndef OCT_U
#define OCT_U 4                // candidates a thread takes per trip of the two per-pass candidate loops
#endif
#ifndef ORB_OCT_SMALL_BATCH
#define ORB_OCT_SMALL_BATCH 8
#endif

struct OctShared {
    short4* box[2];   // (ulx, urx, uly, bry)
    int* cnt[2];
    int* cc;          // [4*cap] child key counts, then new positions of the children
    int* kk;          // [cap] non-empty children of a node (0 if not expandable)
    int* rk;          // [cap] processing rank of a node, -1 if it is not split this pass
    int* gr;          // [sortn] k in processing order -> inclusive prefix
    int* keep;        // [cap] exclusive count of unsplit nodes before a node
    uint32_t* sk;     // [sortn] sort keys of the partial rounds
};

__device__ __forceinline__ int oct_child(const short4 b, const int x, const int y)
{
    const int midx = b.x + ((b.y - b.x + 1) >> 1);   // UL.x + ceil((UR.x-UL.x)/2)  (:438)
    const int midy = b.z + ((b.w - b.z + 1) >> 1);   // UL.y + ceil((BR.y-UL.y)/2)  (:439)
    return (x < midx ? 0 : 1) + (y < midy ? 0 : 2);  // n1,n2,n3,n4 (:467-484)
}

template <int OCT_NT>
__global__ void __launch_bounds__(OCT_NT) k_octree(const __grid_constant__ OrbPlan plan, const OrbBatch io, const int level_lo)
{
    extern __shared__ __align__(16) unsigned char oct_smem[];
    __shared__ int s_scratch[OCT_NT / 32];
    __shared__ int s_nexp, s_m, s_L;

    const int l = blockIdx.x + level_lo, frame = blockIdx.y, tid = threadIdx.x;
    const OrbLevel& LV = plan.lv[l];
    const int cap = (plan.max_nodes + 31) & ~31;
    int sortn = 32; while (sortn < cap) sortn <<= 1;

    OctShared S;
    {
        unsigned char* p = oct_smem;
        S.box[0] = (short4*)p; p += sizeof(short4) * cap;
        S.box[1] = (short4*)p; p += sizeof(short4) * cap;
        S.cnt[0] = (int*)p; p += 4 * cap;
        S.cnt[1] = (int*)p; p += 4 * cap;
        S.cc = (int*)p; p += 16 * cap;
        S.kk = (int*)p; p += 4 * cap;
        S.rk = (int*)p; p += 4 * cap;
        S.keep = (int*)p; p += 4 * cap;
        S.gr = (int*)p; p += 4 * sortn;
        S.sk = (uint32_t*)p;
    }

    const int n = min(io.cand_count[frame * ORB_MAX_LEVELS + l], LV.cand_cap);
    const uint32_t* keys = io.cand + (size_t)frame * plan.cand_per_frame + LV.cand_off;
    uint16_t* node = io.node_of + (size_t)frame * plan.cand_per_frame + LV.cand_off;
    const int N = LV.quota;
    int cur = 0;

    // ---- roots (:567-615): key -> root (int)(x / hX); empty roots are dropped
    const int nIni = LV.nIni;
    for (int i = tid; i < nIni; i += OCT_NT) S.cc[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += OCT_NT) {
        int r = (int)__fdiv_rn((float)ORB_PX(keys[i]), LV.hX);
        r = min(r, nIni - 1);
        node[i] = (uint16_t)r;
        atomicAdd(&S.cc[r], 1);
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += OCT_NT) S.keep[i] = S.cc[i] > 0 ? 1 : 0;
    __syncthreads();
    int L = orb_block_scan_incl<OCT_NT>(S.keep, nIni, s_scratch);
    for (int i = tid; i < nIni; i += OCT_NT) {
        if (S.cc[i] > 0) {
            const int pos = S.keep[i] - 1;
            S.box[0][pos] = make_short4((short)(int)__fmul_rn(LV.hX, (float)i), (short)(int)__fmul_rn(LV.hX, (float)(i + 1)), 0, (short)LV.H);
            S.cnt[0][pos] = S.cc[i];
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += OCT_NT) node[i] = (uint16_t)(S.keep[node[i]] - 1);
    __syncthreads();

    bool partial = false;
    for (int iter = 0; iter < 64 && L > 0; ++iter) {
        short4* box = S.box[cur]; int* cnt = S.cnt[cur];
        short4* nbox = S.box[cur ^ 1]; int* ncnt = S.cnt[cur ^ 1];
        // A. child key counts of every expandable node
        for (int i = tid; i < 4 * L; i += OCT_NT) S.cc[i] = 0;
        if (tid == 0) { s_nexp = 0; s_m = 0; }
        __syncthreads();
        // (four candidates per trip, their loads issued together: the loop waits on global memory, not on arithmetic)
        for (int i0 = tid; i0 < n; i0 += OCT_U * OCT_NT) {
            int pv[OCT_U]; uint32_t kv[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; ++u) { const int i = i0 + u * OCT_NT; pv[u] = i < n ? (int)(node[i] & 0xfff) : -1; kv[u] = i < n ? keys[i] : 0u; }
#pragma unroll
            for (int u = 0; u < OCT_U; ++u) {
                const int p = pv[u];
                if (p >= 0 && cnt[p] > 1) {
                    const int c = oct_child(box[p], ORB_PX(kv[u]), ORB_PY(kv[u]));
                    atomicAdd(&S.cc[4 * p + c], 1);
                    node[i0 + u * OCT_NT] = (uint16_t)(p | (c << 12));
                }
            }
        }
        __syncthreads();
        // B. per node: number of non-empty children; who is expandable
        for (int p = tid; p < L; p += OCT_NT) {
            int k = 0;
            if (cnt[p] > 1) k = (S.cc[4 * p] > 0) + (S.cc[4 * p + 1] > 0) + (S.cc[4 * p + 2] > 0) + (S.cc[4 * p + 3] > 0);
            S.kk[p] = k;
            S.rk[p] = -1;
            S.keep[p] = k > 0 ? 1 : 0;   // expandable flag, scanned below
        }
        __syncthreads();
        // C. processing order of the expandable nodes and how many of them are split (m)
        int nexp, m;
        if (!partial) {
            nexp = orb_block_scan_incl<OCT_NT>(S.keep, L, s_scratch);   // rank = list order (:636)
            for (int p = tid; p < L; p += OCT_NT)
                if (S.kk[p] > 0) { const int r = S.keep[p] - 1; S.rk[p] = r; S.gr[r] = S.kk[p]; }
            m = nexp;
            __syncthreads();
        } else {
            // sort by (count desc, list position asc): key = count << 12 | (4095 - pos), descending
            for (int i = tid; i < sortn; i += OCT_NT) S.sk[i] = 0;
            __syncthreads();
            for (int p = tid; p < L; p += OCT_NT)
                if (S.kk[p] > 0) S.sk[atomicAdd(&s_nexp, 1)] = ((uint32_t)cnt[p] << 12) | (uint32_t)(4095 - p);
            __syncthreads();
            nexp = s_nexp;
            int sn = 32; while (sn < nexp) sn <<= 1;
            for (int k = 2; k <= sn; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < sn; i += OCT_NT) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const uint32_t a = S.sk[i], b = S.sk[ixj];
                            const bool desc = (i & k) == 0;
                            if (desc ? (a < b) : (a > b)) { S.sk[i] = b; S.sk[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int r = tid; r < nexp; r += OCT_NT) {
                const int p = 4095 - (int)(S.sk[r] & 0xfffu);
                S.rk[p] = r;
                S.gr[r] = S.kk[p];
            }
            m = nexp; // refined after the prefix sum
            __syncthreads();
        }
        if (nexp == 0) break;                                           // nothing can be split (:696 size==prevSize)
        orb_block_scan_incl<OCT_NT>(S.gr, nexp, s_scratch);             // G[r] = children created by ranks 0..r
        if (partial) {
            // split in order until size >= N (:759): first r with L + G[r] - (r+1) >= N
            int below = 0;
            for (int r = tid; r < nexp; r += OCT_NT) below += (L + S.gr[r] - (r + 1) < N) ? 1 : 0;
            atomicAdd(&s_m, below);
            __syncthreads();
            m = min(s_m + 1, nexp);
        }
        const int totalNew = S.gr[m - 1];
        // unsplit nodes keep their relative order behind the new children
        for (int p = tid; p < L; p += OCT_NT) S.keep[p] = (S.rk[p] >= 0 && S.rk[p] < m) ? 0 : 1;
        __syncthreads();
        const int nunsplit = orb_block_scan_incl<OCT_NT>(S.keep, L, s_scratch);
        const int newL = totalNew + nunsplit;
        if (tid == 0) s_nexp = 0;
        __syncthreads();
        // D. build the new list
        for (int p = tid; p < L; p += OCT_NT) {
            const int r = S.rk[p];
            if (r >= 0 && r < m) {
                int pos = totalNew - S.gr[r];                           // children of later-processed nodes come first
                const short4 b = box[p];
                const int midx = b.x + ((b.y - b.x + 1) >> 1), midy = b.z + ((b.w - b.z + 1) >> 1);
                int nexp_local = 0;
#pragma unroll
                for (int c = 3; c >= 0; --c) {                          // n4 ends up frontmost (:650-688)
                    const int kc = S.cc[4 * p + c];
                    if (kc > 0) {
                        nbox[pos] = make_short4((short)((c & 1) ? midx : b.x), (short)((c & 1) ? b.y : midx),
                                                (short)((c & 2) ? midy : b.z), (short)((c & 2) ? b.w : midy));
                        ncnt[pos] = kc;
                        S.cc[4 * p + c] = pos;
                        nexp_local += kc > 1;
                        ++pos;
                    } else {
                        S.cc[4 * p + c] = -1;
                    }
                }
                if (nexp_local) atomicAdd(&s_nexp, nexp_local);
            } else {
                const int pos = totalNew + S.keep[p] - 1;
                nbox[pos] = box[p];
                ncnt[pos] = cnt[p];
                S.keep[p] = pos;
                if (cnt[p] > 1) atomicAdd(&s_nexp, 1);
            }
        }
        __syncthreads();
        // E. move the keys
        for (int i0 = tid; i0 < n; i0 += OCT_U * OCT_NT) {
            int vv[OCT_U];
#pragma unroll
            for (int u = 0; u < OCT_U; ++u) { const int i = i0 + u * OCT_NT; vv[u] = i < n ? (int)node[i] : -1; }
#pragma unroll
            for (int u = 0; u < OCT_U; ++u) {
                if (vv[u] < 0) continue;
                const int v = vv[u], p = v & 0xfff;
                const int r = S.rk[p];
                node[i0 + u * OCT_NT] = (uint16_t)((r >= 0 && r < m) ? S.cc[4 * p + (v >> 12)] : S.keep[p]);
            }
        }
        const int nToExpand = s_nexp;
        const int prevL = L;
        L = newL;
        cur ^= 1;
        __syncthreads();
        // F. termination (:696-702, :762-763)
        if (L >= N || L == prevL) break;
        if (!partial && L + 3 * nToExpand > N) partial = true;
    }

    // ---- retained keypoint per node (:773-789): max response, first in reference order on ties
    uint32_t* best = (uint32_t*)S.cc;
    for (int p = tid; p < L; p += OCT_NT) best[p] = 0;
    __syncthreads();
    const int wc = LV.wCell, hc = LV.hCell;
    // x / wCell and y / hCell by multiplication: floor(n / d) = (n * (2^20 / d + 1)) >> 20 for d <= 64 and n < 16384 (coordinates are < 4128)
    const uint32_t mwc = (1u << 20) / (uint32_t)wc + 1u, mhc = (1u << 20) / (uint32_t)hc + 1u;
    for (int i = tid; i < n; i += OCT_NT) {
        const uint32_t k = keys[i];
        const int ax = ORB_PX(k) - 3, ay = ORB_PY(k) - 3;
        const int cj = (int)(((uint32_t)ax * mwc) >> 20), ci = (int)(((uint32_t)ay * mhc) >> 20);
        const uint32_t order = (uint32_t)(((ci * LV.ncx + cj) * hc + (ay - ci * hc)) * wc + (ax - cj * wc));
        atomicMax(&best[node[i] & 0xfff], ((uint32_t)ORB_PS(k) << 24) | (0xffffffu - order));
    }
    __syncthreads();
    uint32_t* out = io.lkp + (size_t)frame * plan.kp_per_frame + LV.kp_off;
    for (int i = tid; i < n; i += OCT_NT) {
        const uint32_t k = keys[i];
        const int ax = ORB_PX(k) - 3, ay = ORB_PY(k) - 3;
        const int cj = (int)(((uint32_t)ax * mwc) >> 20), ci = (int)(((uint32_t)ay * mhc) >> 20);
        const uint32_t order = (uint32_t)(((ci * LV.ncx + cj) * hc + (ay - ci * hc)) * wc + (ax - cj * wc));
        const int p = node[i] & 0xfff;
        if (best[p] == (((uint32_t)ORB_PS(k) << 24) | (0xffffffu - order)) && p < LV.kp_cap) out[p] = k;
    }
    if (tid == 0) io.lkp_count[frame * ORB_MAX_LEVELS + l] = min(L, LV.kp_cap);
    (void)s_L;
}

size_t orb_octree_smem_bytes(const OrbPlan& plan)
{
    const size_t cap = (size_t)((plan.max_nodes + 31) & ~31);
    size_t sortn = 32; while (sortn < cap) sortn <<= 1;
    return cap * (8 * 2 + 4 * 2 + 16 + 4 * 3) + sortn * 8 + 64;
}

cudaError_t orb_launch_octree(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st, int level_lo, int level_hi)
{
    if (level_hi > plan.nlevels) level_hi = plan.nlevels;
    if (level_lo >= level_hi) return cudaSuccess;
    const size_t smem = orb_octree_smem_bytes(plan);
    const bool small = batch <= ORB_OCT_SMALL_BATCH;
    // a chunk of the host-buffer pipeline (64 frames): the blocks of all levels are resident at once, so the launch lasts as
    // long as its slowest block; twice the threads per block shorten that
    const bool mid = !small && (long long)batch * (level_hi - level_lo) <= ORB_OCT_MID_BLOCKS;
    if (mid) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_octree<ORB_OCT_NT_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, ORB_SMEM_OPTIN);
            if (e != cudaSuccess) return e;
        }
        k_octree<ORB_OCT_NT_MID><<<dim3(level_hi - level_lo, batch), ORB_OCT_NT_MID, smem, st>>>(plan, io, level_lo);
        return cudaGetLastError();
    }
    if (smem > 48 * 1024) {
        // a constant: the attribute is per-function state shared by all host threads (a per-launch value races)
        cudaError_t e = small ? cudaFuncSetAttribute(k_octree<ORB_OCT_NT_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, ORB_SMEM_OPTIN)
                              : cudaFuncSetAttribute(k_octree<ORB_OCT_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ORB_SMEM_OPTIN);
        if (e != cudaSuccess) return e;
    }
    if (small) k_octree<ORB_OCT_NT_SMALL><<<dim3(level_hi - level_lo, batch), ORB_OCT_NT_SMALL, smem, st>>>(plan, io, level_lo);
    else k_octree<ORB_OCT_NT><<<dim3(level_hi - level_lo, batch), ORB_OCT_NT, smem, st>>>(plan, io, level_lo);
    return cudaGetLastError();
}

// =============================================================================== describe
// rBRIEF sampling pattern (512 points, x,y int8; one descriptor byte = 16 points = 32 bytes).
// Lanes read different bytes, so it lives in global memory behind L1, not in __constant__.
__device__ __align__(32) const signed char d_pattern[1024] = {
#include "orb_pattern.inc"
};

// cv::fastAtan2 (scalar path of OpenCV's atanImpl<float>), degrees; every operation rounds to
// float on its own (the file is compiled with -fmad=false; the intrinsics say so explicitly).
__device__ __forceinline__ float orb_fast_atan2(const float y, const float x)
{
    const float S = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, S), p3 = __fmul_rn(-0.3258083974640975f, S);
    const float p5 = __fmul_rn(0.1555786518463281f, S), p7 = __fmul_rn(-0.04432655554792128f, S);
    const float eps = 2.2204460492503131e-16f; // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/sincosf.h): reduction by 2/pi*2^24 and the
// double-precision polynomials, each operation rounded separately.
// CUDA's own sinf/cosf differ from glibc in the last bit for ~1% of angles.
__device__ __forceinline__ float orb_sincos_poly(const double x, const double x2, const bool negcos, const int n)
{
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                 C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
        const double x7 = __dmul_rn(x3, x2);
        const double s = __dadd_rn(x, __dmul_rn(x3, S1));
        return (float)__dadd_rn(s, __dmul_rn(x7, s1));
    }
    const double sg = negcos ? -1.0 : 1.0;
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(sg * C3, __dmul_rn(x2, sg * C4));
    const double c1 = __dadd_rn(sg * C0, __dmul_rn(x2, sg * C1));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(c1, __dmul_rn(x4, sg * C2));
    return (float)__dadd_rn(c, __dmul_rn(x6, c2));
}

__device__ __forceinline__ void orb_sincosf(const float y, float* sn, float* cs)
{
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top < 0x3f4) { // |y| < pi/4
        if (top < 0x398) { *sn = y; *cs = 1.0f; return; }
        const double x2 = __dmul_rn(x, x);
        *sn = orb_sincos_poly(x, x2, false, 0);
        *cs = orb_sincos_poly(x, x2, false, 1);
        return;
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = ((int)r + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double xs = x * sgn, x2 = __dmul_rn(x, x);
    const bool neg = (n & 2) != 0;
    *sn = orb_sincos_poly(xs, x2, neg, n);
    *cs = orb_sincos_poly(xs, x2, neg, n ^ 1);
}

#ifndef DESC_NT
#define DESC_NT 128
#endif
// One warp per output keypoint.  Lanes are the 31 columns of the orientation patch, then the
// 32 descriptor bytes (16 rotated taps each).
// Measured dead end (round 2): issuing fastAtan2 + the FP64 sincos once per keypoint instead of in all 32 lanes -- a
// block takes 8..32 keypoints in three phases (moments per warp / one LANE per keypoint for angle, sin, cos / descriptor
// per warp, two block barriers) -- removes ~70 of ~420 instructions per keypoint but measures 1.17..1.45 ms per 512
// frames against 0.99 ms for this version (72 instead of 39 registers, phases that wait for each other): the kernel
// follows the uncoalesced tap gathers, not the issue rate.
// Measured dead end (round 2, second attempt, all inside one warp): a warp takes K = 1 / 2 / 4 / 8 consecutive keypoints,
// computes their moments as an 8 x 4 grid of lanes over the patch (8 row words per lane, byte masks of the disc, two
// IDP.4A per word: ~100 instead of ~290 instructions), runs fastAtan2 + sincos ONCE with keypoint k in lane k, then the K
// descriptors: bit-exact, ~25 % fewer instructions per keypoint, and 1.19 / 1.59 / 1.46 / 1.40 ms per 512 frames against
// 0.99 ms.  The word loads of the grid touch four image rows per instruction (more L1 sectors than the one-row byte loads
// of the column walk), 72 registers halve the resident warps, and a warp that walks K keypoints in turn has one
// keypoint's loads in flight where K warps had K: the kernel is bound by L1 sectors and load latency, not by issue.
__global__ void __launch_bounds__(DESC_NT) k_describe(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    const int frame = blockIdx.y, lane = threadIdx.x & 31;
    const int s = blockIdx.x * (DESC_NT / 32) + (threadIdx.x >> 5);   // output slot within the frame
    // level of slot s: prefix over the per-level keypoint counts (operator() appends level by level, :1118-1148)
    int cnt = lane < plan.nlevels ? io.lkp_count[frame * ORB_MAX_LEVELS + lane] : 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    if (s == 0 && lane == 0) io.n_out[frame] = total;
    if (s >= total || s >= io.cap) return;
    const unsigned ball = __ballot_sync(0xffffffffu, inc > s);
    const int l = __ffs(ball) - 1;
    const int first = __shfl_sync(0xffffffffu, inc - cnt, l);
    const OrbLevel& L = plan.lv[l];
    const uint32_t k = io.lkp[(size_t)frame * plan.kp_per_frame + L.kp_off + (s - first)];
    const int cx = ORB_PX(k) + ORB_BORDER0, cy = ORB_PY(k) + ORB_BORDER0;   // :892-893, level frame

    // ---- IC_Angle (:78-105): integer moments over the 749-px disc
    int pitch;
    const uint8_t* img = orb_level_ptr(plan, io, frame, l, &pitch);
    // lane = column u of the disc.  The disc is symmetric (the reference forces it, :552-558), so column u holds
    // the rows |v| <= umax[|u|]; u is constant per lane, so m10 = u * (column sum) and only m01 needs a multiply
    // per row; the row pointer advances by the pitch.
    const int u = lane - ORB_HALF_PATCH;
    int m10 = 0, m01 = 0;
    if (lane < ORB_PATCH) {
        // rows +v and -v of a column are in or out of the disc together: one predicate, one 3-input add for the column sum
        // and v * (below - above) for m01 per pair
        const int vmax = plan.umax[u < 0 ? -u : u];
        const uint8_t* pd = img + (size_t)cy * pitch + cx + u;
        const uint8_t* pu = pd;
        int colsum = __ldg(pd);
#pragma unroll
        for (int v = 1; v <= ORB_HALF_PATCH; ++v) {
            pd += pitch; pu -= pitch;
            if (v <= vmax) {
                const int below = __ldg(pd), above = __ldg(pu);
                colsum += below + above;
                m01 += v * (below - above);
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, d);
        m01 += __shfl_xor_sync(0xffffffffu, m01, d);
    }
    const float angle = orb_fast_atan2((float)m01, (float)m10);

    // ---- computeOrbDescriptor (:120-161)
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float a, b;
    orb_sincosf(__fmul_rn(angle, factorPI), &b, &a);                   // a = cos, b = sin (:125)
    int blur_pitch = L.pitch;
    const uint8_t* bc = io.blur + (size_t)frame * plan.blur_bytes + L.blur_off + (size_t)cy * blur_pitch + cx;
    __align__(16) signed char pat[32];                                  // 16 points = 8 pairs per byte
    ((int4*)pat)[0] = __ldg((const int4*)(d_pattern + lane * 32));
    ((int4*)pat)[1] = __ldg((const int4*)(d_pattern + lane * 32) + 1);
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float px = (float)pat[4 * j], py = (float)pat[4 * j + 1];
        const float qx = (float)pat[4 * j + 2], qy = (float)pat[4 * j + 3];
        // GET_VALUE (:132-134): row = cvRound(x*b + y*a), col = cvRound(x*a - y*b)
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(qx, b), __fmul_rn(qy, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(qx, a), __fmul_rn(qy, b)));
        const int t0 = __ldg(bc + r0 * blur_pitch + c0), t1 = __ldg(bc + r1 * blur_pitch + c1);
        val |= (t0 < t1) << j;
    }
    io.desc[((size_t)frame * io.cap + s) * 32 + lane] = (uint8_t)val;

    // ---- keypoint record (:892-896, :1140-1146): 7 lanes write the 7 fields
    if (lane < 7) {
        float x = (float)cx, y = (float)cy;
        if (l != 0) { x = __fmul_rn(x, L.scale); y = __fmul_rn(y, L.scale); }
        uint32_t w;
        switch (lane) {
            case 0: w = __float_as_uint(x); break;
            case 1: w = __float_as_uint(y); break;
            case 2: w = __float_as_uint(L.size); break;
            case 3: w = __float_as_uint(angle); break;
            case 4: w = __float_as_uint((float)ORB_PS(k)); break;
            case 5: w = (uint32_t)l; break;
            default: w = 0xffffffffu; break;                              // class_id = -1
        }
        ((uint32_t*)(io.kps + (size_t)frame * io.cap + s))[lane] = w;
    }
}

cudaError_t orb_launch_describe(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    int slots = plan.kp_per_frame < io.cap ? plan.kp_per_frame : io.cap;
    if (slots < 1) slots = 1;
    k_describe<<<dim3((slots + DESC_NT / 32 - 1) / (DESC_NT / 32), batch), DESC_NT, 0, st>>>(plan, io);
    return cudaGetLastError();
}
