// orb_fast.cu -- per-cell FAST-9 with NMS and the iniThFAST -> minThFAST retry, sm_100a.
// Replaces the cell loop of ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:826-875
// (cv::FAST(cell, kps, iniThFAST, true), and again with minThFAST when that returns nothing).
//
// One block per STRIP: a run of up to ORB_FAST_STRIP horizontally adjacent 30-px cells of one
// cell row.  Every cell of the strip is staged into shared memory as its own word-aligned
// image (evaluated rectangle + 3-px ring apron), so all later accesses are aligned 32-bit
// words and the NMS is cell-local by construction.  Everything is done on PACKED PIXEL PAIRS:
// a work item is 4 horizontally adjacent evaluated pixels; each of the 16 ring offsets is one
// funnel-shifted 32-bit window, widened to two 16x2 registers, and the sliding min/max network
// runs on the 3-input DPX instructions (VIMNMX3.S16x2): one instruction advances two pixels and
// there is no data-dependent branch in the scoring.
//
// Score (OpenCV cornerScore<16>, threshold independent), d[k] = v - ring[k]:
//     score = max( max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m] ) - 1
// v is constant over the ring, so the network runs on the RAW ring values:
//     max_k min9(-d) = M1 - v,  M1 = max_k min_{m<9} ring[k+m]
//     max_k min9( d) = v - M2,  M2 = min_k max_{m<9} ring[k+m]
// A pixel is a FAST corner at threshold t iff score >= t.  NMS is the strict 3x3 maximum of the
// score map INSIDE the cell (FAST runs on the cropped cell image, so neighbours outside the
// cell's evaluated rectangle count as 0); both thresholds read the same map, hence the
// reference's retry is a per-cell choice of cut-off: iniThFAST if any NMS survivor of the cell
// reaches it, else minThFAST.
#include "orb_device.cuh"
#include "orb_launch.h"

#define FAST_NT 256

// window of 4 bytes starting DX bytes right of the middle word of (w0,w1,w2)
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

// n / d for 0 <= n < 2^16, 1 <= d < 2^8: (n + 0.5) / d is never within float rounding of an integer
__device__ __forceinline__ int fast_div(const int n, const int d)
{
    return __float2int_rz(__fdividef((float)n + 0.5f, (float)d));
}

// bright = max_k min9(ring) - v, dark = v - min_k max9(ring)
__device__ __forceinline__ int fast_score_of(const int bright, const int dark, const int minTh)
{
    const int s = max(bright, dark) - 1;
    return s >= minTh ? s : 0;
}

__global__ void __launch_bounds__(FAST_NT) k_fast_strips(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    extern __shared__ uint32_t smem[];
    __shared__ int s_nq, s_nsurv, s_nkeep, s_nwr, s_base, s_any[ORB_FAST_STRIP];
    const int frame = blockIdx.y, tid = threadIdx.x;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= plan.lv[l + 1].strip_first) ++l;
    const OrbLevel& L = plan.lv[l];
    const int strip = blockIdx.x - L.strip_first;
    const int ci = strip / L.spr, cj0 = (strip - ci * L.spr) * ORB_FAST_STRIP;
    const int ncs = min(ORB_FAST_STRIP, L.ncx - cj0);
    const int wc = L.wCell, maxBX = L.w - ORB_BORDER0;
    const int y0 = ORB_BORDER0 + ci * L.hCell, y1 = min(y0 + L.hCell + 6, L.h - ORB_BORDER0);
    const int eh = y1 - y0 - 6;                      // evaluated rows y0+3 .. y1-4
    if (eh <= 0) return;
    const int gq = (wc + 3) >> 2;                    // 4-pixel groups per cell row
    const int tw = gq + 2;                           // words per staged cell row; evaluated lx sits at byte lx+4
    const int TW = ncs * tw;                         // words per staged strip row
    const int th = eh + 6;
    uint32_t* tile = smem;                                            // th x TW
    uint32_t* score = smem + plan.fast_tile_words;                    // (eh+2) x TW, pixel (ly,lx) at row ly+1, byte lx+4
    uint16_t* queue = (uint16_t*)(score + plan.fast_score_words);     // items with a non-zero score
    uint32_t* surv = tile;                                            // NMS survivors (tile is dead by then)
    uint16_t* surv_tag = queue + ((plan.fast_items_max + 1) & ~1);    // cell | isA << 15

    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    const int w = L.w;
    // staging: a lane owns a (cell, word) column and walks down the rows, so the index math and the
    // in-range test are paid once per column
    for (int idx = tid & 31; idx < TW; idx += 32) {
        const int c = fast_div(idx, tw), k = idx - c * tw;
        const int x = ORB_BORDER0 + (cj0 + c) * wc - 1 + 4 * k;
        const uint8_t* col = src + (size_t)y0 * pitch + x;
        if (x + 7 < w) {
            for (int r = tid >> 5; r < th; r += FAST_NT / 32) tile[r * TW + idx] = orb_ld_u32_unaligned(col + (size_t)r * pitch);
        } else {
            for (int r = tid >> 5; r < th; r += FAST_NT / 32) {
                uint32_t v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) if (x + b < w) v |= (uint32_t)__ldg(col + (size_t)r * pitch + b) << (8 * b);
                tile[r * TW + idx] = v;
            }
        }
    }
    for (int i = tid; i < (eh + 2) * TW; i += FAST_NT) score[i] = 0;
    if (tid < ORB_FAST_STRIP) s_any[tid] = 0;
    if (tid == 0) { s_nq = 0; s_nsurv = 0; s_nkeep = 0; s_nwr = 0; }
    __syncthreads();

    // ---- scores: one item = 4 pixels (ly, cell c, 4q..4q+3)
    const int ipr = ncs * gq;                        // items per row
    for (int i = tid; i < eh * ipr; i += FAST_NT) {
        const int ly = fast_div(i, ipr), rem = i - ly * ipr, c = fast_div(rem, gq), q = rem - c * gq;
        const int ew = min(wc, maxBX - 6 - (ORB_BORDER0 + (cj0 + c) * wc));   // evaluated width of this cell
        const int valid = ew - 4 * q;
        if (valid <= 0) continue;
        const uint32_t* t0 = tile + ly * TW + c * tw + q;   // ring row dy=-3 is tile row ly, centre row is ly+3
        uint32_t W[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) { W[r][0] = t0[r * TW]; W[r][1] = t0[r * TW + 1]; W[r][2] = t0[r * TW + 2]; }
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
            uint32_t E[16];                                              // raw ring values of the pixel pair
#pragma unroll
            for (int k = 0; k < 16; ++k) E[k] = __byte_perm(win[k], 0, sel);
            uint32_t M1, M2;
            fast_network(E, M1, M2);
            const int v0 = (int)((C >> (16 * hpair)) & 0xffu), v1 = (int)((C >> (16 * hpair + 8)) & 0xffu);
            const int s0 = fast_score_of((int)(M1 & 0xffffu) - v0, v0 - (int)(M2 & 0xffffu), plan.minTh);
            const int s1 = fast_score_of((int)(M1 >> 16) - v1, v1 - (int)(M2 >> 16), plan.minTh);
            out |= ((uint32_t)s0 | ((uint32_t)s1 << 8)) << (16 * hpair);
        }
        // pixels past the evaluated width must stay 0 (they are "outside the cell image" for the NMS)
        if (valid < 4) out &= (1u << (8 * valid)) - 1u;
        if (out) {
            score[(ly + 1) * TW + c * tw + q + 1] = out;
            queue[atomicAdd(&s_nq, 1)] = (uint16_t)(ly | (c << 6) | (q << 9));
        }
    }
    __syncthreads();

    // ---- strict 3x3 maximum inside each cell, over the items that scored at all
    const int nq = s_nq;
    for (int j = tid; j < nq; j += FAST_NT) {
        const int e = queue[j];
        const int ly = e & 63, c = (e >> 6) & 7, q = e >> 9;
        const uint32_t* sc = score + (ly + 1) * TW + c * tw + q;
        const uint32_t cw = sc[1];
        uint32_t nb[2] = { 0, 0 };                                 // neighbour maxima of pairs (0,1) and (2,3)
#pragma unroll
        for (int r = -1; r <= 1; ++r) {
            const uint32_t a0 = sc[r * TW], a1 = sc[r * TW + 1], a2 = sc[r * TW + 2];
            const uint32_t wl = __funnelshift_r(a0, a1, 24);       // bytes 3..6 of the 12-byte span
            const uint32_t wr = __funnelshift_r(a1, a2, 8);        // bytes 5..8
            const uint32_t l1 = __byte_perm(wl, 0, 0x4140), l2 = __byte_perm(wl, 0, 0x4241), l3 = __byte_perm(wl, 0, 0x4342);
            const uint32_t r1 = __byte_perm(wr, 0, 0x4140), r2 = __byte_perm(wr, 0, 0x4241), r3 = __byte_perm(wr, 0, 0x4342);
            if (r == 0) { nb[0] = __vimax3_s16x2(nb[0], l1, l3); nb[1] = __vimax3_s16x2(nb[1], r1, r3); }
            else { nb[0] = __vimax3_s16x2(nb[0], __vmaxs2(l1, l2), l3); nb[1] = __vimax3_s16x2(nb[1], __vmaxs2(r1, r2), r3); }
        }
        // s > m per byte; nb holds 16-bit lanes (m0,m1),(m2,m3)
        uint32_t keep = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int s = (int)((cw >> (8 * p)) & 0xffu);
            const int m = (int)((nb[p >> 1] >> (16 * (p & 1))) & 0xffffu);
            keep |= (s > m ? 1u : 0u) << p;
        }
        while (keep) {
            const int p = __ffs(keep) - 1;
            keep &= keep - 1;
            const int s = (int)((cw >> (8 * p)) & 0xffu);
            const int slot = atomicAdd(&s_nsurv, 1);
            // border-frame coordinates (src/ORBextractor.cc:868-869): cell-local + (j*wCell, i*hCell)
            surv[slot] = orb_pack(4 * q + p + 3 + (cj0 + c) * wc, ly + 3 + ci * L.hCell, s);
            const bool isA = s >= plan.iniTh;
            surv_tag[slot] = (uint16_t)(c | (isA ? 0x8000 : 0));
            if (isA) s_any[c] = 1;
        }
    }
    __syncthreads();
    // ---- per-cell cut-off: keep the iniThFAST survivors, or all of them if the cell has none (:857-861)
    const int nsurv = s_nsurv;
    int mykeep = 0;
    for (int j = tid; j < nsurv; j += FAST_NT) {
        const int t = surv_tag[j];
        if ((t & 0x8000) || !s_any[t & 0x7fff]) ++mykeep;
    }
    if (mykeep) atomicAdd(&s_nkeep, mykeep);
    __syncthreads();
    const int nkeep = s_nkeep;
    if (nkeep == 0) return;
    if (tid == 0) s_base = atomicAdd(&io.cand_count[frame * ORB_MAX_LEVELS + l], nkeep);
    __syncthreads();
    const int base = s_base;
    uint32_t* out = io.cand + (size_t)frame * plan.cand_per_frame + L.cand_off;
    for (int j = tid; j < nsurv; j += FAST_NT) {
        const int t = surv_tag[j];
        if ((t & 0x8000) || !s_any[t & 0x7fff]) {
            const int slot = base + atomicAdd(&s_nwr, 1);
            if (slot < L.cand_cap) out[slot] = surv[j];
        }
    }
}

size_t orb_fast_smem_bytes(const OrbPlan& plan)
{
    return ((size_t)plan.fast_tile_words + plan.fast_score_words) * 4 + ((size_t)((plan.fast_items_max + 1) & ~1)) * 2 +
           (size_t)plan.fast_surv_max * 2 + 32;
}

cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    if (plan.total_strips == 0) return cudaSuccess;
    const size_t smem = orb_fast_smem_bytes(plan);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_fast_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_fast_strips<<<dim3(plan.total_strips, batch), FAST_NT, smem, st>>>(plan, io);
    return cudaGetLastError();
}
