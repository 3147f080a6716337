// orb_dense.cu -- the dense (per-pixel) kernels of the extractor, sm_100a:
//   k_pyr_resize   ComputePyramid            src/ORBextractor.cc:1153-1180 (cv::resize INTER_LINEAR 8U)
//   k_blur7        GaussianBlur 7x7 sigma 2  src/ORBextractor.cc:1129-1130
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

cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st)
{
    dim3 blk(32, 8), grd((w + 2 * b + 31) / 32, (h + 2 * b + 7) / 8);
    k_border<<<grd, blk, 0, st>>>(src, w, h, spitch, dst, dpitch, b);
    return cudaGetLastError();
}
