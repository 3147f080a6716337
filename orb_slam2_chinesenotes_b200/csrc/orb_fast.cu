// orb_fast.cu -- per-cell FAST-9 with NMS and the iniThFAST -> minThFAST retry, sm_100a.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:826-875
// (cv::FAST(cell, kps, iniThFAST, true), and again with minThFAST when that returns nothing).
//
// One block per processed 30-px cell.  Everything is done on PACKED PIXEL PAIRS: a thread owns
// 4 horizontally adjacent evaluated pixels, fetches each of the 16 ring offsets as one
// (funnel-shifted) 32-bit window of shared memory, widens it to two 16x2 registers and runs
// the sliding min/max network with the 3-input DPX instructions (VIMNMX3.S16x2), so one
// instruction advances two pixels and there is no data-dependent branch in the scoring.
//
// Score (OpenCV cornerScore<16>, threshold independent), d[k] = v - ring[k]:
//     score = max( max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m] ) - 1
// With the biased E[k] = ring[k] + (255 - v) = 255 - d[k]  (0..510, no carry between halves):
//     max_k min9(-d) = M1 - 255,  M1 = max_k min_{m<9} E[k+m]
//     max_k min9( d) = 255 - M2,  M2 = min_k max_{m<9} E[k+m]
// A pixel is a FAST corner at threshold t iff score >= t.  NMS is the strict 3x3 maximum of the
// score map INSIDE the cell (FAST runs on the cropped cell image, so neighbours outside the
// cell's evaluated rectangle count as 0); both thresholds read the same map, hence the
// reference's retry is a per-cell choice of cut-off: iniThFAST if any NMS survivor reaches it,
// else minThFAST.
#include "orb_device.cuh"
#include "orb_launch.h"

#define FAST_NT 128

// window of 4 bytes starting dx bytes right of the middle word of (w0,w1,w2)
template <int DX>
__device__ __forceinline__ uint32_t fast_win(const uint32_t w0, const uint32_t w1, const uint32_t w2)
{
    if (DX < 0) return __funnelshift_r(w0, w1, 8 * (4 + DX));
    if (DX > 0) return __funnelshift_r(w1, w2, 8 * DX);
    return w1;
}

// sliding-window network over the 16 biased ring values of a pixel pair -> packed (M1, M2)
__device__ __forceinline__ void fast_network(const uint32_t* E, uint32_t& M1, uint32_t& M2)
{
    uint32_t A[16], B[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        A[j] = __vimin3_s16x2(E[j], E[(j + 1) & 15], E[(j + 2) & 15]);
        B[j] = __vimax3_s16x2(E[j], E[(j + 1) & 15], E[(j + 2) & 15]);
    }
    uint32_t mn[16], mx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        mn[i] = __vimin3_s16x2(A[i], A[(i + 3) & 15], A[(i + 6) & 15]);   // min E[i..i+8]
        mx[i] = __vimax3_s16x2(B[i], B[(i + 3) & 15], B[(i + 6) & 15]);   // max E[i..i+8]
    }
    uint32_t a = __vimax3_s16x2(mn[0], mn[1], mn[2]), b = __vimax3_s16x2(mn[3], mn[4], mn[5]);
    uint32_t c = __vimax3_s16x2(mn[6], mn[7], mn[8]), d = __vimax3_s16x2(mn[9], mn[10], mn[11]);
    uint32_t e = __vimax3_s16x2(mn[12], mn[13], mn[14]);
    M1 = __vimax3_s16x2(__vimax3_s16x2(a, b, c), __vimax3_s16x2(d, e, mn[15]), a);
    a = __vimin3_s16x2(mx[0], mx[1], mx[2]); b = __vimin3_s16x2(mx[3], mx[4], mx[5]);
    c = __vimin3_s16x2(mx[6], mx[7], mx[8]); d = __vimin3_s16x2(mx[9], mx[10], mx[11]);
    e = __vimin3_s16x2(mx[12], mx[13], mx[14]);
    M2 = __vimin3_s16x2(__vimin3_s16x2(a, b, c), __vimin3_s16x2(d, e, mx[15]), a);
}

__device__ __forceinline__ int fast_score_of(const int m1, const int m2, const int minTh)
{
    const int s = max(m1 - 255, 255 - m2) - 1;
    return s >= minTh ? s : 0;
}

__global__ void __launch_bounds__(FAST_NT) k_fast_cells(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    extern __shared__ uint32_t smem[];
    __shared__ int s_nA, s_nB, s_base;
    const int frame = blockIdx.y;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= plan.lv[l + 1].cell_first) ++l;
    const OrbLevel& L = plan.lv[l];
    const int cell = blockIdx.x - L.cell_first;
    const int ci = cell / L.ncx, cj = cell - ci * L.ncx;
    const int x0 = ORB_BORDER0 + cj * L.wCell, y0 = ORB_BORDER0 + ci * L.hCell;
    const int x1 = min(x0 + L.wCell + 6, L.w - ORB_BORDER0), y1 = min(y0 + L.hCell + 6, L.h - ORB_BORDER0);
    const int ew = x1 - x0 - 6, eh = y1 - y0 - 6;   // evaluated rectangle: x0+3.., y0+3..
    if (ew <= 0 || eh <= 0) return;
    const int gq = (ew + 3) >> 2;                   // 4-pixel groups per row
    const int tw = gq + 2;                          // words per staged row: bytes x0-1 .. (evaluated lx at byte lx+4)
    const int th = eh + 6;
    uint32_t* tile = smem;
    uint32_t* score = smem + plan.fast_tile_words;  // (eh+2) rows x tw words, pixel lx at byte lx+4 of row ly+1

    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    for (int i = threadIdx.x; i < th * tw; i += FAST_NT) {
        const int r = i / tw, k = i - r * tw;
        tile[i] = orb_ld_u32_unaligned(src + (size_t)(y0 + r) * pitch + (x0 - 1) + 4 * k);
    }
    for (int i = threadIdx.x; i < (eh + 2) * tw; i += FAST_NT) score[i] = 0;
    if (threadIdx.x == 0) { s_nA = 0; s_nB = 0; }
    __syncthreads();

    // ---- scores: one item = 4 pixels (ly, 4q..4q+3)
    for (int i = threadIdx.x; i < eh * gq; i += FAST_NT) {
        const int ly = i / gq, q = i - ly * gq;
        const uint32_t* t0 = tile + ly * tw + q;   // ring row dy=-3 is tile row ly, centre row is ly+3
        uint32_t W[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) { W[r][0] = t0[r * tw]; W[r][1] = t0[r * tw + 1]; W[r][2] = t0[r * tw + 2]; }
        // ring windows in OpenCV order; row index = dy + 3
        uint32_t win[16];
        win[0] = fast_win<0>(W[6][0], W[6][1], W[6][2]);    //  ( 0, 3)
        win[1] = fast_win<1>(W[6][0], W[6][1], W[6][2]);    //  ( 1, 3)
        win[2] = fast_win<2>(W[5][0], W[5][1], W[5][2]);    //  ( 2, 2)
        win[3] = fast_win<3>(W[4][0], W[4][1], W[4][2]);    //  ( 3, 1)
        win[4] = fast_win<3>(W[3][0], W[3][1], W[3][2]);    //  ( 3, 0)
        win[5] = fast_win<3>(W[2][0], W[2][1], W[2][2]);    //  ( 3,-1)
        win[6] = fast_win<2>(W[1][0], W[1][1], W[1][2]);    //  ( 2,-2)
        win[7] = fast_win<1>(W[0][0], W[0][1], W[0][2]);    //  ( 1,-3)
        win[8] = fast_win<0>(W[0][0], W[0][1], W[0][2]);    //  ( 0,-3)
        win[9] = fast_win<-1>(W[0][0], W[0][1], W[0][2]);   //  (-1,-3)
        win[10] = fast_win<-2>(W[1][0], W[1][1], W[1][2]);  //  (-2,-2)
        win[11] = fast_win<-3>(W[2][0], W[2][1], W[2][2]);  //  (-3,-1)
        win[12] = fast_win<-3>(W[3][0], W[3][1], W[3][2]);  //  (-3, 0)
        win[13] = fast_win<-3>(W[4][0], W[4][1], W[4][2]);  //  (-3, 1)
        win[14] = fast_win<-2>(W[5][0], W[5][1], W[5][2]);  //  (-2, 2)
        win[15] = fast_win<-1>(W[6][0], W[6][1], W[6][2]);  //  (-1, 3)
        const uint32_t C = W[3][1];
        uint32_t out = 0;
#pragma unroll
        for (int hpair = 0; hpair < 2; ++hpair) {
            const uint32_t sel = hpair ? 0x4342u : 0x4140u;
            const uint32_t up = __byte_perm(C, 0, sel) ^ 0x00ff00ffu;   // (255 - v) per half
            uint32_t E[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) E[k] = __byte_perm(win[k], 0, sel) + up;
            uint32_t M1, M2;
            fast_network(E, M1, M2);
            const int s0 = fast_score_of((int)(M1 & 0xffffu), (int)(M2 & 0xffffu), plan.minTh);
            const int s1 = fast_score_of((int)(M1 >> 16), (int)(M2 >> 16), plan.minTh);
            out |= ((uint32_t)s0 | ((uint32_t)s1 << 8)) << (16 * hpair);
        }
        // pixels past the evaluated width must stay 0 (they are "outside the cell image" for the NMS)
        const int valid = ew - 4 * q;
        if (valid < 4) out &= (1u << (8 * valid)) - 1u;
        score[(ly + 1) * tw + q + 1] = out;
    }
    __syncthreads();

    // ---- strict 3x3 maximum inside the cell, survivors split by iniThFAST
    // A-list (score >= iniThFAST) grows from the front of the free tile area, B-list from the back
    uint32_t* stage = tile;
    const int stage_cap = th * tw;
    for (int i = threadIdx.x; i < eh * gq; i += FAST_NT) {
        const int ly = i / gq, q = i - ly * gq;
        const uint32_t* sc = score + (ly + 1) * tw + q;
        const uint32_t cw = sc[1];
        if (cw == 0) continue;
        uint32_t nb[2] = { 0, 0 };                                 // neighbour maxima of pairs (0,1) and (2,3)
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
            const uint32_t a0 = sc[r * tw], a1 = sc[r * tw + 1], a2 = sc[r * tw + 2];
            const uint32_t wl = __funnelshift_r(a0, a1, 24);       // bytes 3..6 of the 12-byte span
            const uint32_t wr = __funnelshift_r(a1, a2, 8);        // bytes 5..8
            const uint32_t l1 = __byte_perm(wl, 0, 0x4140), l2 = __byte_perm(wl, 0, 0x4241), l3 = __byte_perm(wl, 0, 0x4342);
            const uint32_t r1 = __byte_perm(wr, 0, 0x4140), r2 = __byte_perm(wr, 0, 0x4241), r3 = __byte_perm(wr, 0, 0x4342);
            if (r == 0) { nb[0] = __vimax3_s16x2(nb[0], l1, l3); nb[1] = __vimax3_s16x2(nb[1], r1, r3); }
            else { nb[0] = __vimax3_s16x2(nb[0], __vmaxs2(l1, l2), l3); nb[1] = __vimax3_s16x2(nb[1], __vmaxs2(r1, r2), r3); }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int s = (int)((cw >> (8 * j)) & 0xffu);
            const int m = (int)((nb[j >> 1] >> (16 * (j & 1))) & 0xffffu);
            if (s > m) {
                // border-frame coordinates (src/ORBextractor.cc:868-869): cell-local + (j*wCell, i*hCell)
                const uint32_t rec = orb_pack(4 * q + j + 3 + cj * L.wCell, ly + 3 + ci * L.hCell, s);
                if (s >= plan.iniTh) stage[atomicAdd(&s_nA, 1)] = rec;
                else stage[stage_cap - 1 - atomicAdd(&s_nB, 1)] = rec;
            }
        }
    }
    __syncthreads();
    const int nA = s_nA, nB = s_nB;
    const int nout = nA ? nA : nB;                                 // retry at minThFAST only if the ini pass is empty
    if (nout == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&io.cand_count[frame * ORB_MAX_LEVELS + l], nout);
    __syncthreads();
    const int base = s_base;
    uint32_t* out = io.cand + (size_t)frame * plan.cand_per_frame + L.cand_off;
    for (int i = threadIdx.x; i < nout; i += FAST_NT)
        if (base + i < L.cand_cap) out[base + i] = nA ? stage[i] : stage[stage_cap - 1 - i];
}

size_t orb_fast_smem_bytes(const OrbPlan& plan)
{
    return ((size_t)plan.fast_tile_words + plan.fast_score_words) * 4 + 16;
}

cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    if (plan.total_cells == 0) return cudaSuccess;
    const size_t smem = orb_fast_smem_bytes(plan);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_fast_cells<<<dim3(plan.total_cells, batch), FAST_NT, smem, st>>>(plan, io);
    return cudaGetLastError();
}
