// orb_dense.cu -- the dense (per-pixel) kernels of the extractor, sm_100a:
//   k_pyr_resize   ComputePyramid            src/ORBextractor.cc:1153-1180 (cv::resize INTER_LINEAR 8U)
//   k_blur7        GaussianBlur 7x7 sigma 2  src/ORBextractor.cc:1129-1130
//   k_border       REFLECT_101 border of mvImagePyramid               src/ORBextractor.cc:1168-1174
// Integer / byte arithmetic only; results are bit-identical to OpenCV 4.13.
// Both filters are separable and staged through shared memory; on this part they are bound by
// instruction issue (ncu: ALU pipe ~70 %), not by HBM, so the kernels are organised to spend as
// few instructions per pixel as possible: index math is hoisted per thread, the 8-bit pass of
// the blur runs on packed 16x2 lanes (two pixels per IMAD), and every store is a 32-bit word.
#include "orb_device.cuh"
#include "orb_launch.h"

// ------------------------------------------------------------------------------ pyramid
// cv::resize INTER_LINEAR on 8U (OpenCV HResizeLinear / VResizeLinear<uchar,int,short>):
//   H(r,x)   = s[r][ofs]*a0 + s[r][ofs+1]*a1            (11-bit weights, int32)
//   dst(y,x) = ( ((b0*(H(r0,x)>>4))>>16) + ((b1*(H(r1,x)>>4))>>16) + 2 ) >> 2
// A block makes a PYR_TW x PYR_TH tile: first the horizontal pass for the source rows the tile
// needs (a thread owns one destination column, so its tap is loaded once), stored as H>>4 in
// 16 bits; then the vertical pass, 4 pixels per thread, one 32-bit store.
#define PYR_TW 128
#define PYR_TH 16
#define PYR_NT 256
#define PYR_ROWS (2 * PYR_TH + 3)   // source rows per tile for scale factors up to 2

__global__ void __launch_bounds__(PYR_NT) k_pyr_resize(const __grid_constant__ OrbPlan plan, const OrbBatch io, const int l)
{
    __shared__ __align__(8) uint32_t s_h[PYR_ROWS * (PYR_TW / 2)];   // (H >> 4) as 16-bit, two per word
    const int frame = blockIdx.z, tid = threadIdx.x;
    const OrbLevel& D = plan.lv[l];
    const OrbLevel& S = plan.lv[l - 1];
    const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;
    int spitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l - 1, &spitch);
    const int2* xtab = (const int2*)io.taps + D.xtab;
    const int2* ytab = (const int2*)io.taps + D.ytab;
    const int ylast = min(ty0 + PYR_TH, D.h) - 1;
    const int rs0 = __ldg(&ytab[ty0]).x;
    const int rs1 = min(__ldg(&ytab[ylast]).x + 1, S.h - 1);
    const int nrows = rs1 - rs0 + 1;                     // <= PYR_ROWS (checked by the launcher)
    {   // horizontal pass
        const int c = tid & (PYR_TW - 1);
        const int2 t = __ldg(&xtab[min(tx0 + c, D.w - 1)]);
        const int sx0 = t.x, sx1 = min(t.x + 1, S.w - 1);
        const int a0 = t.y & 0xffff, a1 = (int)((uint32_t)t.y >> 16);
        const uint8_t* p = src + (size_t)rs0 * spitch;
        uint16_t* hs = (uint16_t*)s_h;
        for (int r = tid / PYR_TW; r < nrows; r += PYR_NT / PYR_TW) {
            const uint8_t* row = p + (size_t)r * spitch;
            const int H = (int)__ldg(row + sx0) * a0 + (int)__ldg(row + sx1) * a1;
            hs[r * PYR_TW + c] = (uint16_t)(H >> 4);
        }
    }
    __syncthreads();
    uint8_t* dst = io.pyr + (size_t)frame * plan.pyr_bytes + D.img_off;
    const int g = tid & 31;                              // 4-pixel group
    const int x = tx0 + 4 * g;
    if (x >= D.pitch) return;
    for (int yy = tid >> 5; yy < PYR_TH; yy += PYR_NT / 32) {
        const int y = ty0 + yy;
        if (y >= D.h) break;
        const int2 t = __ldg(&ytab[y]);
        const int r0 = t.x - rs0, r1 = min(t.x + 1, S.h - 1) - rs0;
        const int b0 = t.y & 0xffff, b1 = (int)((uint32_t)t.y >> 16);
        const uint2 h0 = *(const uint2*)&s_h[r0 * (PYR_TW / 2) + 2 * g];
        const uint2 h1 = *(const uint2*)&s_h[r1 * (PYR_TW / 2) + 2 * g];
        const int v0 = (((b0 * (int)(h0.x & 0xffffu)) >> 16) + ((b1 * (int)(h1.x & 0xffffu)) >> 16) + 2) >> 2;
        const int v1 = (((b0 * (int)(h0.x >> 16)) >> 16) + ((b1 * (int)(h1.x >> 16)) >> 16) + 2) >> 2;
        const int v2 = (((b0 * (int)(h0.y & 0xffffu)) >> 16) + ((b1 * (int)(h1.y & 0xffffu)) >> 16) + 2) >> 2;
        const int v3 = (((b0 * (int)(h0.y >> 16)) >> 16) + ((b1 * (int)(h1.y >> 16)) >> 16) + 2) >> 2;
        *(uint32_t*)(dst + (size_t)y * D.pitch + x) = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
    }
}

// Generic fallback (any scale factor): one thread = 4 destination pixels straight from global.
__global__ void __launch_bounds__(256) k_pyr_resize_generic(const __grid_constant__ OrbPlan plan, const OrbBatch io, const int l)
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
    const int2 ty = __ldg((const int2*)&io.taps[D.ytab + y]);
    const int sy0 = ty.x, sy1 = min(sy0 + 1, S.h - 1);
    const int b0 = ty.y & 0xffff, b1 = (int)((uint32_t)ty.y >> 16);
    const uint8_t* r0 = src + (size_t)sy0 * spitch;
    const uint8_t* r1 = src + (size_t)sy1 * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x4 + i;
        if (x < D.w) {
            const int2 tx = __ldg((const int2*)&io.taps[D.xtab + x]);
            const int sx0 = tx.x, sx1 = min(sx0 + 1, S.w - 1);
            const int a0 = tx.y & 0xffff, a1 = (int)((uint32_t)tx.y >> 16);
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
// (acc + 2^15) >> 16 at the end; REFLECT_101 at the level's own edges (the reference blurs a
// clone of the ROI).  A block produces an ORB_BLUR_TW x ORB_BLUR_TH tile:
//   stage   (TH+6) rows x 32 words (pixels tx0-4 .. tx0+123), a lane owns one word column;
//   column pass on the BYTES first: sums <= 255*256 fit 16 bits, so two pixels share one
//           register lane pair and one IMAD advances both (no carry between the halves);
//   row pass on the 16-bit sums with 32-bit accumulators, 4 pixels per thread from a sliding
//           window of 12 values, one 32-bit store.
// (Column-then-row equals row-then-column: the sums are exact integers.)
#define BLUR_NT 256
#define BLUR_IW 32   // staged words per row

__global__ void __launch_bounds__(BLUR_NT) k_blur7(const __grid_constant__ OrbPlan plan, const OrbBatch io)
{
    __shared__ uint32_t s_in[(ORB_BLUR_TH + 6) * BLUR_IW];
    __shared__ __align__(8) uint32_t s_v[ORB_BLUR_TH * 2 * BLUR_IW];   // column-pass sums, 16-bit, two per word
    const int frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int l = 0;
    while (l + 1 < plan.nlevels && (int)blockIdx.x >= plan.lv[l + 1].blur_tile_first) ++l;
    const OrbLevel& L = plan.lv[l];
    const int t = blockIdx.x - L.blur_tile_first;
    const int tyi = t / L.blur_tiles_x;
    const int tx0 = (t - tyi * L.blur_tiles_x) * ORB_BLUR_TW, ty0 = tyi * ORB_BLUR_TH;
    int pitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &pitch);
    const int w = L.w, h = L.h;

    {   // stage: lane = word column (pixels x .. x+3), warps walk down the rows.  Columns / rows more than
        // 3 px outside the image only feed outputs that are never stored, so indices are clamped to
        // [-3, n+2] first and ONE reflection is enough (n >= 4 always holds here).
        const int x = tx0 - 4 + 4 * lane;
        const bool fast = x >= 0 && x + 7 < w;
        int xr[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = min(max(x + b, -3), w + 2);
            i = i < 0 ? -i : i;
            xr[b] = i >= w ? 2 * (w - 1) - i : i;
        }
        for (int r = warp; r < ORB_BLUR_TH + 6; r += BLUR_NT / 32) {
            int y = min(max(ty0 - 3 + r, -3), h + 2);
            y = y < 0 ? -y : y;
            y = y >= h ? 2 * (h - 1) - y : y;
            const uint8_t* row = src + (size_t)y * pitch;
            uint32_t v;
            if (fast) v = orb_ld_u32_unaligned(row + x);
            else v = (uint32_t)__ldg(row + xr[0]) | ((uint32_t)__ldg(row + xr[1]) << 8) | ((uint32_t)__ldg(row + xr[2]) << 16) | ((uint32_t)__ldg(row + xr[3]) << 24);
            s_in[r * BLUR_IW + lane] = v;
        }
    }
    __syncthreads();
    {   // column pass: (word column = lane, block of 4 output rows = warp)
        uint32_t lo[10], hi[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const uint32_t v = s_in[(4 * warp + i) * BLUR_IW + lane];
            lo[i] = __byte_perm(v, 0, 0x4140);   // (p0, p1) as 16-bit lanes
            hi[i] = __byte_perm(v, 0, 0x4342);   // (p2, p3)
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            uint2 r;
            r.x = 18u * (lo[o] + lo[o + 6]) + 34u * (lo[o + 1] + lo[o + 5]) + 48u * (lo[o + 2] + lo[o + 4]) + 56u * lo[o + 3];
            r.y = 18u * (hi[o] + hi[o + 6]) + 34u * (hi[o + 1] + hi[o + 5]) + 48u * (hi[o + 2] + hi[o + 4]) + 56u * hi[o + 3];
            *(uint2*)&s_v[(4 * warp + o) * (2 * BLUR_IW) + 2 * lane] = r;
        }
    }
    __syncthreads();
    // row pass: group g = output pixels tx0+4g .. +3 = staged columns 4g+4 .. 4g+7; window = columns 4g+1 .. 4g+10
    uint8_t* dst = io.blur + (size_t)frame * plan.blur_bytes + L.blur_off;
    // lane = group (30 of 32 lanes busy), warps walk down the rows: no division, addresses advance by a pitch
    if (lane < ORB_BLUR_TW / 4 && tx0 + 4 * lane < L.pitch) {
        const int g = lane;
        uint8_t* out = dst + (size_t)(ty0 + warp) * L.pitch + tx0 + 4 * g;
        const int rmax = min(ORB_BLUR_TH, h - ty0);
        for (int r = warp; r < rmax; r += BLUR_NT / 32, out += (size_t)(BLUR_NT / 32) * L.pitch) {
            const uint2* pv = (const uint2*)&s_v[r * (2 * BLUR_IW) + 2 * g];
            const uint2 q0 = pv[0], q1 = pv[1], q2 = pv[2];
            uint32_t v[12];
            v[0] = q0.x & 0xffffu; v[1] = q0.x >> 16; v[2] = q0.y & 0xffffu; v[3] = q0.y >> 16;
            v[4] = q1.x & 0xffffu; v[5] = q1.x >> 16; v[6] = q1.y & 0xffffu; v[7] = q1.y >> 16;
            v[8] = q2.x & 0xffffu; v[9] = q2.x >> 16; v[10] = q2.y & 0xffffu; v[11] = q2.y >> 16;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                o[j] = (18u * (v[j + 1] + v[j + 7]) + 34u * (v[j + 2] + v[j + 6]) + 48u * (v[j + 3] + v[j + 5]) + 56u * v[j + 4] + 32768u) >> 16;
            *(uint32_t*)out = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        }
    }
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
cudaError_t orb_launch_pyramid(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    for (int l = 1; l < plan.nlevels; ++l) {
        const OrbLevel& D = plan.lv[l];
        const OrbLevel& S = plan.lv[l - 1];
        // source rows spanned by PYR_TH destination rows: <= PYR_TH * (S.h / D.h) + 2
        const bool tiled = (long long)PYR_TH * S.h + 3LL * D.h <= (long long)PYR_ROWS * D.h;
        if (tiled) {
            dim3 grd((D.pitch + PYR_TW - 1) / PYR_TW, (D.h + PYR_TH - 1) / PYR_TH, batch);
            k_pyr_resize<<<grd, PYR_NT, 0, st>>>(plan, io, l);
        } else {
            dim3 blk(32, 8), grd((D.pitch / 4 + 31) / 32, (D.h + 7) / 8, batch);
            k_pyr_resize_generic<<<grd, blk, 0, st>>>(plan, io, l);
        }
    }
    return cudaGetLastError();
}

cudaError_t orb_launch_blur(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    k_blur7<<<dim3(plan.total_blur_tiles, batch), BLUR_NT, 0, st>>>(plan, io);
    return cudaGetLastError();
}

cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st)
{
    dim3 blk(32, 8), grd((w + 2 * b + 31) / 32, (h + 2 * b + 7) / 8);
    k_border<<<grd, blk, 0, st>>>(src, w, h, spitch, dst, dpitch, b);
    return cudaGetLastError();
}
