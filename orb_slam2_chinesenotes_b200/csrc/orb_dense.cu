// orb_dense.cu -- the dense (per-pixel) kernels of the extractor, sm_100a:
//   k_pyr_resize   ComputePyramid            src/ORBextractor.cc:1153-1180 (cv::resize INTER_LINEAR 8U)
//   k_blur7        GaussianBlur 7x7 sigma 2  src/ORBextractor.cc:1129-1130
//   k_fast_cells   per-cell FAST-9 + NMS + iniThFAST->minThFAST retry  src/ORBextractor.cc:826-875
//   k_border       REFLECT_101 border of mvImagePyramid               src/ORBextractor.cc:1168-1174
// Integer / byte arithmetic only; results are bit-identical to OpenCV 4.13.
#include "orb_device.cuh"
#include "orb_launch.h"

#define BLUR_TW 128
#define BLUR_TH 32
#define BLUR_NT 256

// ------------------------------------------------------------------------------ pyramid
// One thread = 4 horizontally adjacent destination pixels (one aligned 32-bit store).
// dst(y,x) = ((b0*(H0>>4))>>16 + (b1*(H1>>4))>>16 + 2) >> 2, H = s[ofs]*a0 + s[ofs+1]*a1
// (OpenCV HResizeLinear / VResizeLinear<uchar,int,short>, weights 11-bit).
__global__ void __launch_bounds__(256) k_pyr_resize(const __grid_constant__ OrbPlan plan, const OrbBatch io, const int l)
{
    const int frame = blockIdx.z;
    const OrbLevel& D = plan.lv[l];
    const OrbLevel& S = plan.lv[l - 1];
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= D.h || x4 >= D.pitch) return;
    int spitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l - 1, &spitch);
    uint8_t* dst = io.pyr + (size_t)frame * plan.pyr_bytes + D.img_off;
    const int2 tyv = __ldg((const int2*)&io.taps[D.ytab + y]);
    OrbTap ty; ty.ofs = tyv.x; ty.c01 = (uint32_t)tyv.y;
    const int sy0 = ty.ofs, sy1 = min(sy0 + 1, S.h - 1);
    const int b0 = (int)(ty.c01 & 0xffffu), b1 = (int)(ty.c01 >> 16);
    const uint8_t* r0 = src + (size_t)sy0 * spitch;
    const uint8_t* r1 = src + (size_t)sy1 * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x4 + i;
        if (x < D.w) {
            const int2 txv = __ldg((const int2*)&io.taps[D.xtab + x]);
            OrbTap tx; tx.ofs = txv.x; tx.c01 = (uint32_t)txv.y;
            const int sx0 = tx.ofs, sx1 = min(sx0 + 1, S.w - 1);
            const int a0 = (int)(tx.c01 & 0xffffu), a1 = (int)(tx.c01 >> 16);
            const int H0 = (int)__ldg(r0 + sx0) * a0 + (int)__ldg(r0 + sx1) * a1;
            const int H1 = (int)__ldg(r1 + sx0) * a0 + (int)__ldg(r1 + sx1) * a1;
            const int v = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xff) << (8 * i);
        }
    }
    *(uint32_t*)(dst + (size_t)y * D.pitch + x4) = out;
}

// ------------------------------------------------------------------------------ blur
// Separable fixed-point Gaussian, kernel {18,34,48,56,48,34,18}/256 per axis, one rounding
// (acc + 2^15) >> 16 after the column pass; REFLECT_101 at the level's own edges (the
// reference blurs a clone of the ROI).  A block produces a 128x32 tile: rows are staged as
// 32-bit words in shared memory, the row pass leaves 16-bit sums, the column pass emits
// 4 pixels per thread as one 32-bit store.
__global__ void __launch_bounds__(BLUR_NT) k_blur7(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    __shared__ uint32_t s_in[(BLUR_TH + 6) * (BLUR_TW / 4 + 2)];
    __shared__ uint32_t s_h[(BLUR_TH + 6) * (BLUR_TW / 2)]; // two 16-bit sums per word
    const int frame = blockIdx.y;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= plan.lv[l + 1].blur_tile_first) ++l;
    const OrbLevel& L = plan.lv[l];
    const int t = blockIdx.x - L.blur_tile_first;
    const int tx0 = (t % L.blur_tiles_x) * BLUR_TW, ty0 = (t / L.blur_tiles_x) * BLUR_TH;
    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    const int w = L.w, h = L.h;
    constexpr int IW = BLUR_TW / 4 + 2; // input words per staged row: pixels tx0-4 .. tx0+131

    // stage (BLUR_TH+6) x IW words
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * IW; i += BLUR_NT) {
        const int r = i / IW, k = i - r * IW;
        const int y = orb_refl101(ty0 - 3 + r, h);
        const int x = tx0 - 4 + 4 * k;
        const uint8_t* row = src + (size_t)y * pitch;
        uint32_t v;
        if (x >= 0 && x + 7 < w) {
            v = orb_ld_u32_unaligned(row + x);
        } else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)__ldg(row + orb_refl101(x + b, w)) << (8 * b);
        }
        s_in[i] = v;
    }
    __syncthreads();
    // row pass: each item = 4 adjacent outputs of one staged row (bytes 1..10 of 3 words)
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += BLUR_NT) {
        const int r = i / (BLUR_TW / 4), q = i - r * (BLUR_TW / 4);
        const uint32_t w0 = s_in[r * IW + q], w1 = s_in[r * IW + q + 1], w2 = s_in[r * IW + q + 2];
        int p[10];
        p[0] = (w0 >> 8) & 0xff; p[1] = (w0 >> 16) & 0xff; p[2] = w0 >> 24;
        p[3] = w1 & 0xff; p[4] = (w1 >> 8) & 0xff; p[5] = (w1 >> 16) & 0xff; p[6] = w1 >> 24;
        p[7] = w2 & 0xff; p[8] = (w2 >> 8) & 0xff; p[9] = (w2 >> 16) & 0xff;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = 18 * (p[j] + p[j + 6]) + 34 * (p[j + 1] + p[j + 5]) + 48 * (p[j + 2] + p[j + 4]) + 56 * p[j + 3];
        s_h[r * (BLUR_TW / 2) + 2 * q] = o[0] | (o[1] << 16);
        s_h[r * (BLUR_TW / 2) + 2 * q + 1] = o[2] | (o[3] << 16);
    }
    __syncthreads();
    // column pass
    uint8_t* dst = io.blur + (size_t)frame * plan.blur_bytes + L.blur_off;
    for (int i = threadIdx.x; i < BLUR_TH * (BLUR_TW / 4); i += BLUR_NT) {
        const int r = i / (BLUR_TW / 4), q = i - r * (BLUR_TW / 4);
        const int y = ty0 + r, x = tx0 + 4 * q;
        if (y >= h || x >= L.pitch) continue;
        uint32_t a0 = 32768u, a1 = 32768u, a2 = 32768u, a3 = 32768u;
        const uint32_t K[7] = { 18, 34, 48, 56, 48, 34, 18 };
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const uint2 hv = *(const uint2*)&s_h[(r + j) * (BLUR_TW / 2) + 2 * q];
            a0 += K[j] * (hv.x & 0xffffu); a1 += K[j] * (hv.x >> 16);
            a2 += K[j] * (hv.y & 0xffffu); a3 += K[j] * (hv.y >> 16);
        }
        *(uint32_t*)(dst + (size_t)y * L.pitch + x) = (a0 >> 16) | ((a1 >> 16) << 8) | ((a2 >> 16) << 16) | ((a3 >> 16) << 24);
    }
}

// ------------------------------------------------------------------------------ FAST
// Threshold-independent FAST-9 score (OpenCV cornerScore<16>):
//   score = max( max_k min_{m<9} d[k+m], max_k min_{m<9} -d[k+m] ) - 1,  d[k] = v - ring[k]
// computed on packed signed 16-bit pairs P[i] = (d[i], d[i+8]) with the 3-input DPX min/max.
__device__ __forceinline__ int orb_fast_score(const int v, const int* r)
{
    uint32_t E[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) E[i] = __byte_perm((uint32_t)(v - r[i]), (uint32_t)(v - r[i + 8]), 0x5410);
#pragma unroll
    for (int i = 0; i < 8; ++i) E[8 + i] = __byte_perm(E[i], 0, 0x1032); // halves swapped: (d[i+8], d[i])
    uint32_t A[14], B[14];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        A[j] = __vimin3_s16x2(E[j], E[j + 1], E[j + 2]);
        B[j] = __vimax3_s16x2(E[j], E[j + 1], E[j + 2]);
    }
#pragma unroll
    for (int j = 8; j < 14; ++j) { A[j] = __byte_perm(A[j - 8], 0, 0x1032); B[j] = __byte_perm(B[j - 8], 0, 0x1032); }
    uint32_t mn[8], mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mn[i] = __vimin3_s16x2(A[i], A[i + 3], A[i + 6]);   // (min d[i..i+8], min d[i+8..i+16])
        mx[i] = __vimax3_s16x2(B[i], B[i + 3], B[i + 6]);
    }
    uint32_t p = __vimax3_s16x2(__vimax3_s16x2(mn[0], mn[1], mn[2]), __vimax3_s16x2(mn[3], mn[4], mn[5]), __vmaxs2(mn[6], mn[7]));
    uint32_t n = __vimin3_s16x2(__vimin3_s16x2(mx[0], mx[1], mx[2]), __vimin3_s16x2(mx[3], mx[4], mx[5]), __vmins2(mx[6], mx[7]));
    const int best_pos = max((int)(short)(p & 0xffffu), (int)(short)(p >> 16));
    const int best_neg = min((int)(short)(n & 0xffffu), (int)(short)(n >> 16));
    return max(best_pos, -best_neg) - 1;
}

#define FAST_NT 128
// One block per processed 30-px cell (src/ORBextractor.cc:826-850).  The cell image (evaluated
// rectangle + 3-px ring apron) is staged in shared memory with vectorised row loads, every
// evaluated pixel gets the antipodal pre-test at minThFAST, survivors are compacted into a
// queue and scored, then the 3x3 strict NMS runs CELL-LOCALLY (neighbours outside the cell's
// evaluated rectangle count as 0, which is what FAST on a cropped cell image does).  The
// reference re-runs FAST at minThFAST only when the iniThFAST pass returns nothing; both
// passes read the same score map, so the retry is a per-cell choice of cut-off.
__global__ void __launch_bounds__(FAST_NT) k_fast_cells(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    extern __shared__ uint32_t smem[];
    __shared__ int s_nq, s_any_ini, s_nout, s_base;
    const int frame = blockIdx.y;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= plan.lv[l + 1].cell_first) ++l;
    const OrbLevel& L = plan.lv[l];
    const int cell = blockIdx.x - L.cell_first;
    const int ci = cell / L.ncx, cj = cell - ci * L.ncx;
    const int x0 = ORB_BORDER0 + cj * L.wCell, y0 = ORB_BORDER0 + ci * L.hCell;
    const int x1 = min(x0 + L.wCell + 6, L.w - ORB_BORDER0), y1 = min(y0 + L.hCell + 6, L.h - ORB_BORDER0);
    const int ew = x1 - x0 - 6, eh = y1 - y0 - 6;   // evaluated rectangle
    if (ew <= 0 || eh <= 0) return;
    const int tw = (x1 - x0 + 3) >> 2;              // tile words per row
    const int tp = tw * 4;                          // tile pitch in bytes
    const int th = y1 - y0;
    uint8_t* tile = (uint8_t*)smem;
    const int sp = ew + 2;                          // score pitch (1-px zero frame)
    uint8_t* score = (uint8_t*)(smem + plan.fast_tile_words);
    const int score_words = (sp * (eh + 2) + 3) >> 2;
    uint16_t* queue = (uint16_t*)(smem + plan.fast_tile_words + plan.fast_score_words);

    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    // stage: tile rows y0..y1, columns x0..x0+tp (x0+tp+3 <= w-1: the row continues 16 px past x1)
    for (int i = threadIdx.x; i < th * tw; i += FAST_NT) {
        const int r = i / tw, k = i - r * tw;
        smem[r * tw + k] = orb_ld_u32_unaligned(src + (size_t)(y0 + r) * pitch + x0 + 4 * k);
    }
    for (int i = threadIdx.x; i < score_words; i += FAST_NT) ((uint32_t*)score)[i] = 0;
    if (threadIdx.x == 0) { s_nq = 0; s_any_ini = 0; s_nout = 0; }
    __syncthreads();

    const int t = plan.minTh;
    const int o1 = tp, o2 = 2 * tp, o3 = 3 * tp;
    // antipodal pre-test: a 9-arc contains one pixel of each opposite pair
    for (int ly = threadIdx.x >> 5; ly < eh; ly += FAST_NT / 32) {
        for (int lx = threadIdx.x & 31; lx < ew; lx += 32) {
            const uint8_t* c = tile + (ly + 3) * tp + lx + 3;
            const int v = c[0], hi = v + t, lo = v - t;
            int a = c[o3], b = c[-o3];                                   // ring 0 (0,3), 8 (0,-3)
            bool br = (a > hi) | (b > hi), dk = (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = c[3]; b = c[-3];                                         // ring 4 (3,0), 12 (-3,0)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = c[o2 + 2]; b = c[-o2 - 2];                               // ring 2 (2,2), 10 (-2,-2)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = c[-o2 + 2]; b = c[o2 - 2];                               // ring 6 (2,-2), 14 (-2,2)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            a = c[o3 + 1]; b = c[-o3 - 1];                               // ring 1 (1,3), 9 (-1,-3)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = c[o1 + 3]; b = c[-o1 - 3];                               // ring 3 (3,1), 11 (-3,-1)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = c[-o1 + 3]; b = c[o1 - 3];                               // ring 5 (3,-1), 13 (-3,1)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            a = c[-o3 + 1]; b = c[o3 - 1];                               // ring 7 (1,-3), 15 (-1,3)
            br &= (a > hi) | (b > hi); dk &= (a < lo) | (b < lo);
            if (!(br | dk)) continue;
            queue[atomicAdd(&s_nq, 1)] = (uint16_t)(ly * ew + lx);
        }
    }
    __syncthreads();
    const int nq = s_nq;
    // exact score of the survivors; corners at minThFAST are those with score >= minThFAST
    for (int i = threadIdx.x; i < nq; i += FAST_NT) {
        const int e = queue[i], ly = e / ew, lx = e - ly * ew;
        const uint8_t* c = tile + (ly + 3) * tp + lx + 3;
        int r[16];
        r[0] = c[o3];       r[1] = c[o3 + 1];   r[2] = c[o2 + 2];   r[3] = c[o1 + 3];
        r[4] = c[3];        r[5] = c[-o1 + 3];  r[6] = c[-o2 + 2];  r[7] = c[-o3 + 1];
        r[8] = c[-o3];      r[9] = c[-o3 - 1];  r[10] = c[-o2 - 2]; r[11] = c[-o1 - 3];
        r[12] = c[-3];      r[13] = c[o1 - 3];  r[14] = c[o2 - 2];  r[15] = c[o3 - 1];
        const int s = orb_fast_score(c[0], r);
        if (s >= t) score[(ly + 1) * sp + lx + 1] = (uint8_t)s;
    }
    __syncthreads();
    // strict 3x3 maximum inside the cell; remember survivors in the queue entry's top bit
    for (int i = threadIdx.x; i < nq; i += FAST_NT) {
        const int e = queue[i], ly = e / ew, lx = e - ly * ew;
        const uint8_t* p = score + (ly + 1) * sp + lx + 1;
        const int s = p[0];
        bool keep = s > 0;
        keep = keep && s > p[-1] && s > p[1] && s > p[-sp - 1] && s > p[-sp] && s > p[-sp + 1] &&
               s > p[sp - 1] && s > p[sp] && s > p[sp + 1];
        if (keep) {
            queue[i] = (uint16_t)(e | 0x8000);
            if (s >= plan.iniTh) s_any_ini = 1;
        }
    }
    __syncthreads();
    const int cut = s_any_ini ? plan.iniTh : plan.minTh;
    // survivors at the chosen cut-off: packed into the (now free) tile area, then one
    // reservation in the level's candidate list per cell and a coalesced copy out
    uint32_t* stage = smem;
    for (int i = threadIdx.x; i < nq; i += FAST_NT) {
        const int e = queue[i];
        if (!(e & 0x8000)) continue;
        const int ee = e & 0x7fff, ly = ee / ew, lx = ee - ly * ew;
        const int s = score[(ly + 1) * sp + lx + 1];
        // border-frame coordinates (src/ORBextractor.cc:868-869): cell-local + (j*wCell, i*hCell)
        if (s >= cut) stage[atomicAdd(&s_nout, 1)] = orb_pack(lx + 3 + cj * L.wCell, ly + 3 + ci * L.hCell, s);
    }
    __syncthreads();
    const int nout = s_nout;
    if (nout == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&io.cand_count[frame * ORB_MAX_LEVELS + l], nout);
    __syncthreads();
    const int base = s_base;
    uint32_t* out = io.cand + (size_t)frame * plan.cand_per_frame + L.cand_off;
    for (int i = threadIdx.x; i < nout; i += FAST_NT)
        if (base + i < L.cand_cap) out[base + i] = stage[i];
}

// ------------------------------------------------------------------------------ border
// (w+2b) x (h+2b) REFLECT_101-padded copy of one level (mvImagePyramid's parent buffer).
__global__ void k_border(const uint8_t* __restrict__ src, int w, int h, int spitch, uint8_t* __restrict__ dst, int dpitch, int b)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w + 2 * b || y >= h + 2 * b) return;
    dst[(size_t)y * dpitch + x] = src[(size_t)orb_refl101(y - b, h) * spitch + orb_refl101(x - b, w)];
}

// ------------------------------------------------------------------------------ launchers
size_t orb_fast_smem_bytes(const OrbPlan& plan)
{
    return ((size_t)plan.fast_tile_words + plan.fast_score_words) * 4 + (size_t)plan.fast_eval_max * 2 + 16;
}

cudaError_t orb_launch_pyramid(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    for (int l = 1; l < plan.nlevels; ++l) {
        const OrbLevel& D = plan.lv[l];
        dim3 blk(32, 8), grd((D.pitch / 4 + 31) / 32, (D.h + 7) / 8, batch);
        k_pyr_resize<<<grd, blk, 0, st>>>(plan, io, l);
    }
    return cudaGetLastError();
}

cudaError_t orb_launch_blur(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    k_blur7<<<dim3(plan.total_blur_tiles, batch), BLUR_NT, 0, st>>>(plan, io);
    return cudaGetLastError();
}

cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    if (plan.total_cells == 0) return cudaSuccess;
    const size_t smem = orb_fast_smem_bytes(plan);
    static bool attr_done = false;
    if (smem > 48 * 1024 && !attr_done) {
        cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    k_fast_cells<<<dim3(plan.total_cells, batch), FAST_NT, smem, st>>>(plan, io);
    return cudaGetLastError();
}

cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st)
{
    dim3 blk(32, 8), grd((w + 2 * b + 31) / 32, (h + 2 * b + 7) / 8);
    k_border<<<grd, blk, 0, st>>>(src, w, h, spitch, dst, dpitch, b);
    return cudaGetLastError();
}
