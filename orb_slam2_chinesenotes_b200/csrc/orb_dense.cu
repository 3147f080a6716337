// orb_dense.cu -- the dense (per-pixel) kernels of the extractor, sm_100a:
//   k_pyr_resize   ComputePyramid            src/ORBextractor.cc:1153-1180 (cv::resize INTER_LINEAR 8U)
//   k_blur7        GaussianBlur 7x7 sigma 2  src/ORBextractor.cc:1129-1130
//   k_border       REFLECT_101 border of mvImagePyramid               src/ORBextractor.cc:1168-1174
// Integer / byte arithmetic only; results are bit-identical to OpenCV 4.13.
// Both filters are separable; on this part they are bound by instruction issue before HBM, so
// the kernels are organised to spend as few instructions per pixel as possible: index math is
// hoisted per thread, the 8-bit pass of the blur runs on packed 16x2 lanes (two pixels per
// IMAD), its 16-bit pass on IDP.2A, and every store is a 32-bit word.
#include "orb_device.cuh"
#include "orb_launch.h"

__device__ __forceinline__ uint32_t dp2a_lo(const uint32_t a, const uint32_t b, const uint32_t c)
{
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(const uint32_t a, const uint32_t b, const uint32_t c)
{
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ------------------------------------------------------------------------------ pyramid
// cv::resize INTER_LINEAR on 8U (OpenCV HResizeLinear / VResizeLinear<uchar,int,short>):
//   H(r,x)   = s[r][ofs]*a0 + s[r][ofs+1]*a1            (11-bit weights, int32)
//   dst(y,x) = ( ((b0*(H(r0,x)>>4))>>16) + ((b1*(H(r1,x)>>4))>>16) + 2 ) >> 2
// No shared memory and no barriers: a WARP owns a column of 128 destination pixels (lane = 4
// adjacent pixels) and walks down PYR_TH destination rows.
//   horizontal pass of ONE source row: the lane's 8 taps lie in 8 consecutive source bytes (scale <= 1.25);
//           they are fetched as aligned words (one source row AHEAD of their use), funnel-shifted to the lane's first
//           tap, permuted into two (p0,p0+1,p1,p1+1) byte quads (selectors fixed per lane), and each H is one IDP.2A
//           with the (a0,a1) pair as the 16-bit operand;
//   vertical pass: the walk goes over SOURCE rows r: H(r) is computed exactly once, and the destination row whose
//           upper tap is r-1 (at most one, the vertical scale is >= 1) is emitted from H(r-1), H(r) -- no routing of
//           rows to register slots, one warp-uniform compare per source row.  H is kept as H & ~15 and the row weights
//           are pre-shifted by 12, so ((H >> 4) * b) >> 16 is one IMAD.HI and the sum of both rows + 2 is two.
#ifndef PYR_NT
#define PYR_NT 64
#endif
#define PYR_TW 128
#ifndef PYR_MIN_WARPS
#define PYR_MIN_WARPS 64   // warps per SM a launch should at least have before its tiles get taller
#endif
#define PYR_TH 64

// Everything a launch needs about its level, as plain kernel parameters (constant bank, fixed offsets).
struct PyrJob {
    const uint8_t* src; size_t src_stride; int spitch;      // source level: first frame, bytes between frames / rows
    uint8_t* dst; size_t dst_stride; int dpitch;            // destination level
    int sw, sh, dw, dh;
    const int2* xtab;                                       // OrbTap table of the x axis
    const int4* ytab;                                       // y axis: { source row, c0 << 12, c1 << 12, - }
    int th, tiles_x, tiles_y;                               // destination rows per warp, warp tiles per frame
};

#ifndef PYR_STAGES
#define PYR_STAGES 4     // source rows in flight per warp
#endif
#ifndef PYR_LD64
#define PYR_LD64 1       // 1: a lane's window travels as two 8-byte words (four MIO instructions per source row), 0: as three 4-byte words (six)
#endif
#if PYR_LD64
typedef uint2 PyrSlot;
#define PYR_WORDS 2
#else
typedef uint32_t PyrSlot;
#define PYR_WORDS 3
#endif
__device__ __forceinline__ void cp_async4(const uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(const uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(const uint32_t dst, const void* src, const uint32_t src_bytes)   // the rest of the 8 bytes is zero-filled
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(const uint32_t dst, const void* src, const uint32_t src_bytes)   // src_bytes 0: nothing is read
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// The row walk of one lane.  The kernel waits on memory latency, not on instruction issue (ncu: long scoreboard), and
// registers spent on prefetching cost resident warps, so the source rows travel through a small shared-memory ring
// instead: each lane posts the three aligned words of its own taps with cp.async (LDGSTS) PYR_STAGES rows ahead and reads
// them back when the row's turn comes -- lane-private slots, no barrier, only cp.async.wait_group.
// ALLFAST: no lane of the warp touches the right edge of the source.  Otherwise words that start behind the row's last
// pixel are zero-filled instead of read (they only ever meet the zero weight of the clamped tap).
template <bool ALLFAST>
__device__ __forceinline__ void pyr_walk(const PyrJob& J, PyrSlot (*ring)[PYR_WORDS][32], const int frame, const int lane, const int d0, const int sx0, const bool live,
                                         const uint32_t (&C)[4], const uint32_t selA, const uint32_t selB, const int yb, const int ye)
{
    const int sh = J.sh, spitch = J.spitch, dpitch = J.dpitch;
    const int room = J.sw - sx0;                                           // source bytes from the lane's first tap to the row end
    int4 ty = __ldg(&J.ytab[yb]);                                          // taps of the next destination row to emit
    const int r_last = min(__ldg(&J.ytab[ye - 1]).x + 1, sh - 1);
    int r = ty.x;
    const uint8_t* a = J.src + (size_t)frame * J.src_stride + (size_t)r * spitch + sx0;   // the lane's first tap in the row being POSTED
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(&ring[0][0][lane]);
    auto post = [&](const int slot, const bool on) {                        // one commit group per row, empty once the rows are used up
        if (on) {
#if PYR_LD64
            // the 8 tap bytes lie in the 16 bytes from the 8-byte boundary at or below the first one
            const uint8_t* q = (const uint8_t*)((uintptr_t)a & ~(uintptr_t)7);
            const uint32_t d = ring_s + (uint32_t)slot * (2 * 32 * 8);
            if (ALLFAST) { cp_async8(d, q); cp_async8(d + 256, q + 8); }
            else {
                const int al = (int)((uintptr_t)a & 7);
                cp_async8_zfill(d, q, (uint32_t)min(al + room, 8));         // starts at or before the first tap; only what the row still holds
                cp_async8_zfill(d + 256, q + 8, (uint32_t)min(max(room - (8 - al), 0), 8));
            }
#else
            const uint8_t* q = (const uint8_t*)((uintptr_t)a & ~(uintptr_t)3);
            const uint32_t d = ring_s + (uint32_t)slot * (3 * 32 * 4);
            if (ALLFAST) { cp_async4(d, q); cp_async4(d + 128, q + 4); cp_async4(d + 256, q + 8); }
            else {
                const int al = (int)((uintptr_t)a & 3);
                cp_async4(d, q);
                cp_async4_zfill(d + 128, q + 4, 4 - al < room ? 4u : 0u);
                cp_async4_zfill(d + 256, q + 8, 8 - al < room ? 4u : 0u);
            }
#endif
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        a += spitch;
    };
#pragma unroll
    for (int s = 0; s < PYR_STAGES; ++s) post(s, r + s <= r_last);
    const uint32_t back = (uint32_t)(PYR_STAGES * spitch);                 // `a` runs this far ahead of the row being consumed
    uint8_t* __restrict__ out = J.dst + (size_t)frame * J.dst_stride + (size_t)yb * dpitch + d0;
    int y = yb;
    // destination row y from the H rows of its two taps; the weights arrive pre-shifted by 12
    auto emit = [&](const uint32_t* __restrict__ H0, const uint32_t b0, const uint32_t* __restrict__ H1, const uint32_t b1) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __umulhi(H0[j], b0) + __umulhi(H1[j], b1) + 2u;
        // (v0 | v1 << 16) >> 2 leaks two bits of v1 into bits 14-15, which the byte selection below skips
        const uint32_t px = __byte_perm(__byte_perm(v[0], v[1], 0x5410) >> 2, __byte_perm(v[2], v[3], 0x5410) >> 2, 0x6420);
        if (live) *(uint32_t*)out = px;
        out += dpitch;
        ++y;
        if (y < ye) ty = __ldg(&J.ytab[y]); else ty.x = -2;               // -2: no source row ends it
    };
    // one source row: H(r) into Hn; emits the destination row whose taps are (r-1, r).  Returns true after the last source
    // row, when the rows of the bottom clamp (upper tap sh-1, lower weight 0) are emitted as well.
    auto advance = [&](const int slot, uint32_t* __restrict__ Hn, const uint32_t* __restrict__ Ho) -> bool {
        asm volatile("cp.async.wait_group %0;" :: "n"(PYR_STAGES - 1) : "memory");
#if PYR_LD64
        const uint2 lo = ring[slot][0][lane], hi = ring[slot][1][lane];
        const uint32_t ac = (uint32_t)(uintptr_t)a - back;                 // low address bits of the row being consumed
        const bool up = (ac & 4u) != 0;                                     // the window starts in the second word
        const uint32_t q0 = up ? lo.y : lo.x, q1 = up ? hi.x : lo.y, q2 = up ? hi.y : hi.x;
        const uint32_t sft = ac << 3;                                       // SHF.W takes the amount modulo 32
#else
        const uint32_t q0 = ring[slot][0][lane], q1 = ring[slot][1][lane], q2 = ring[slot][2][lane];
        const uint32_t sft = ((uint32_t)(uintptr_t)a - back) << 3;        // SHF.W takes the amount modulo 32
#endif
        const uint32_t w0 = __funnelshift_r(q0, q1, sft), w1 = __funnelshift_r(q1, q2, sft);
        const uint32_t A = __byte_perm(w0, w1, selA), B = __byte_perm(w0, w1, selB);
        post(slot, r + PYR_STAGES <= r_last);                              // after the words have been consumed
        Hn[0] = dp2a_lo(C[0], A, 0u) & 0xfffffff0u;
        Hn[1] = dp2a_hi(C[1], A, 0u) & 0xfffffff0u;
        Hn[2] = dp2a_lo(C[2], B, 0u) & 0xfffffff0u;
        Hn[3] = dp2a_hi(C[3], B, 0u) & 0xfffffff0u;
        if (ty.x == r - 1) emit(Ho, (uint32_t)ty.y, Hn, (uint32_t)ty.z);
        if (r == r_last) {
            while (y < ye) emit(Hn, (uint32_t)ty.y, Hn, 0u);
            return true;
        }
        ++r;
        return false;
    };
    uint32_t HA[4] = { 0u, 0u, 0u, 0u }, HB[4] = { 0u, 0u, 0u, 0u };
    static_assert(PYR_STAGES % 2 == 0, "the slot walk below pairs the H registers with the slots");
    for (;;) {
#pragma unroll
        for (int s = 0; s < PYR_STAGES; s += 2) {
            if (advance(s, HA, HB)) return;
            if (advance(s + 1, HB, HA)) return;
        }
    }
}

__global__ void __launch_bounds__(PYR_NT) k_pyr_resize(const __grid_constant__ PyrJob J)
{
    const int frame = blockIdx.z, lane = threadIdx.x & 31;
    const int sw = J.sw, dw = J.dw;
    const int wt = blockIdx.x * (PYR_NT / 32) + (threadIdx.x >> 5);
    if (wt >= J.tiles_x * J.tiles_y) return;
    const int tyi = __float2int_rz(__fmul_rn((float)wt + 0.5f, __frcp_rn((float)J.tiles_x)));
    const int d0 = (wt - tyi * J.tiles_x) * PYR_TW + 4 * lane;            // first destination column of the lane
    const int yb = tyi * J.th, ye = min(yb + J.th, J.dh);
    const bool live = d0 < dw;                                             // lanes right of the level walk along on column 0 and store nothing
    const int dc = live ? d0 : 0;
    // ---- column setup: taps of the 4 destination pixels relative to the first one
    uint32_t C[4];
    int rel[4];
    const int2 t0 = __ldg(&J.xtab[min(dc, dw - 1)]);
    const int sx0 = t0.x;
    C[0] = (uint32_t)t0.y; rel[0] = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        const int2 t = __ldg(&J.xtab[min(dc + j, dw - 1)]);
        C[j] = (uint32_t)t.y; rel[j] = min(t.x - sx0, 6);
    }
    const uint32_t selA = (uint32_t)(rel[0] | ((rel[0] + 1) << 4) | (rel[1] << 8) | ((rel[1] + 1) << 12));
    const uint32_t selB = (uint32_t)(rel[2] | ((rel[2] + 1) << 4) | (rel[3] << 8) | ((rel[3] + 1) << 12));
    // bytes sx0 .. sx0+7 are fetched through three aligned words
    __shared__ __align__(8) PyrSlot ring[PYR_NT / 32][PYR_STAGES][PYR_WORDS][32];
    const bool fast = sx0 + (PYR_LD64 ? 16 : 12) <= sw;                    // the lane's whole fetch window lies inside the row
    if (__all_sync(0xffffffffu, fast)) pyr_walk<true>(J, ring[threadIdx.x >> 5], frame, lane, d0, sx0, live, C, selA, selB, yb, ye);
    else pyr_walk<false>(J, ring[threadIdx.x >> 5], frame, lane, d0, sx0, live, C, selA, selB, yb, ye);
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
// clone of the ROI).  No shared memory and no barriers: a WARP owns a column of ORB_BLUR_TW
// output pixels (lane = 4 adjacent pixels; lanes 0 and 31 only carry the 3-px apron) and walks
// down ORB_BLUR_TH rows:
//   load    one 32-bit word per lane and row, issued 7 rows ahead of its use;
//   column pass on the BYTES first: the 7 rows slide through registers as packed 16x2 lanes
//           (sums <= 255*256 fit 16 bits, no carry between the halves), one IMAD advances two pixels;
//   row pass on the 16-bit column sums: the neighbours' sums arrive by shuffle, and each output is
//           four IDP.2A (two 16-bit sums x two 8-bit weights + accumulate, rounding constant folded in);
//   store   one 32-bit word per lane and row.
// (Column-then-row equals row-then-column: the sums are exact integers.)
#ifndef BLUR_NT
#define BLUR_NT 32       // one warp per block; measured per 1024 frames: 32 threads 1.33 ms, 64 1.52, 128 1.54, 256 1.66
#endif

// Filter taps as bytes for IDP.2A (.lo uses bytes 0-1, .hi bytes 2-3) and the rounding constant, handed over as kernel
// parameters so that they are constant-bank operands of the instructions (as literals they are re-materialised into
// uniform registers in every row of the loop).
struct BlurTaps { uint32_t w0, w1, w2, w3, rnd; };

// The row walk of one warp.  Like the pyramid, the kernel waits on memory before it runs out of issue slots, so the rows
// travel through a per-warp shared-memory ring filled by cp.async seven rows ahead (lane-private slots, no barrier) and the
// prefetch costs no registers.  Every lane fetches the one or two aligned words that hold its four pixels -- for the lanes at
// the level's left / right edge the four REFLECT_101 columns, which span at most four bytes as well -- and one PRMT whose
// selector is (column offsets + the row's misalignment) puts them in order: no lane takes a different path.
// ONEWORD: every row of the level starts on a 4-byte boundary and no lane of the warp is an edge lane, so the four
// pixels ARE one aligned word.
template <bool ONEWORD>
__device__ __forceinline__ void blur_walk(const BlurTaps& T, uint32_t (*ring)[2][32], const int lane, const uint8_t* __restrict__ src, const int pitch, const int w, const int h,
                                          uint8_t* __restrict__ out, const int opitch, const int x, const int yb, const int ye, const bool writer)
{
    // columns more than 3 px outside the image only feed outputs that are never stored, so indices are clamped to
    // [-3, w+2] first and ONE reflection is enough (w >= 4 holds here)
    int xr[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int i = min(max(x + b, -3), w + 2);
        i = i < 0 ? -i : i;
        xr[b] = i >= w ? 2 * (w - 1) - i : i;
    }
    const int xmin = min(min(xr[0], xr[1]), min(xr[2], xr[3])), xmax = max(max(xr[0], xr[1]), max(xr[2], xr[3]));
    const uint32_t sel0 = (uint32_t)((xr[0] - xmin) | ((xr[1] - xmin) << 4) | ((xr[2] - xmin) << 8) | ((xr[3] - xmin) << 12));
    const int span = xmax - xmin + 1;                                      // <= 4
    const int h2 = 2 * (h - 1);
    const uint8_t* const col = src + xmin;
    auto row_ptr = [&](int y) -> const uint8_t* {                          // y in [-3, h+2]
        y = abs(y);
        y = min(y, h2 - y);
        return col + (size_t)(unsigned)y * (unsigned)pitch;
    };
    auto load_row = [&](const int y) -> uint32_t {                         // direct loads: the rows above the first output row
        const uint8_t* p = row_ptr(y);
        if (ONEWORD) return __ldg((const uint32_t*)p);
        const int al = (int)((uintptr_t)p & 3);
        const uint32_t* q = (const uint32_t*)(p - al);
        const uint32_t w0 = __ldg(q), w1 = al + span > 4 ? __ldg(q + 1) : 0u;
        return __byte_perm(w0, w1, sel0 + 0x1111u * (uint32_t)al);
    };
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(&ring[0][0][lane]);
    auto post = [&](const int slot, const int y, const bool on) {          // one commit group per row, empty once the rows are used up
        if (on) {
            const uint8_t* p = row_ptr(y);
            const uint32_t d = ring_s + (uint32_t)slot * (2 * 32 * 4);
            if (ONEWORD) cp_async4(d, p);
            else {
                const int al = (int)((uintptr_t)p & 3);
                cp_async4(d, p - al);
                cp_async4_zfill(d + 128, p - al + 4, al + span > 4 ? 4u : 0u);   // never reads a word that starts behind the lane's last pixel
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto take = [&](const int slot, const int y) -> uint32_t {             // row y out of the ring
        asm volatile("cp.async.wait_group 6;" ::: "memory");
        if (ONEWORD) return ring[slot][0][lane];
        const uint32_t al = (uint32_t)(uintptr_t)row_ptr(y) & 3u;
        return __byte_perm(ring[slot][0][lane], ring[slot][1][lane], sel0 + 0x1111u * al);
    };
    const uint32_t W0 = T.w0, W1 = T.w1, W2 = T.w2, W3 = T.w3, RND = T.rnd;
    const int ylast = ye + 2;                                              // last row any output of this tile reads
#pragma unroll
    for (int i = 0; i < 7; ++i) post(i, yb + 3 + i, yb + 3 + i <= ylast);  // rows yb+3 .. yb+9
    uint32_t lo[7], hi[7];
#pragma unroll
    for (int i = 0; i < 6; ++i) {                                          // rows yb-3 .. yb+2
        const uint32_t v = load_row(yb - 3 + i);
        lo[i] = __byte_perm(v, 0, 0x4140);
        hi[i] = __byte_perm(v, 0, 0x4342);
    }
    for (int y7 = yb; y7 < ye; y7 += 7) {
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            const int y = y7 + u;
            if (y >= ye) break;
#define SL(i) ((u + (i)) % 7)
            // row y-3+i sits in slot SL(i); the new row y+3 was posted 7 iterations ago
            const uint32_t v = take(u, y + 3);
            post(u, y + 10, y + 10 <= ylast);
            lo[SL(6)] = __byte_perm(v, 0, 0x4140);
            hi[SL(6)] = __byte_perm(v, 0, 0x4342);
            const uint32_t vlo = 18u * (lo[SL(0)] + lo[SL(6)]) + 34u * (lo[SL(1)] + lo[SL(5)]) + 48u * (lo[SL(2)] + lo[SL(4)]) + 56u * lo[SL(3)];
            const uint32_t vhi = 18u * (hi[SL(0)] + hi[SL(6)]) + 34u * (hi[SL(1)] + hi[SL(5)]) + 48u * (hi[SL(2)] + hi[SL(4)]) + 56u * hi[SL(3)];
#undef SL
            // column sums of pixels x-4..x-1 (left lane) and x+4..x+7 (right lane)
            const uint32_t llo = __shfl_up_sync(0xffffffffu, vlo, 1), lhi = __shfl_up_sync(0xffffffffu, vhi, 1);
            const uint32_t rlo = __shfl_down_sync(0xffffffffu, vlo, 1), rhi = __shfl_down_sync(0xffffffffu, vhi, 1);
            // out(x+j) = 18 s[j-3] + 34 s[j-2] + 48 s[j-1] + 56 s[j] + 48 s[j+1] + 34 s[j+2] + 18 s[j+3] + 2^15
            const uint32_t o0 = dp2a_hi(vhi, W1, dp2a_lo(vlo, W1, dp2a_hi(lhi, W0, dp2a_lo(llo, W0, RND))));
            const uint32_t o1 = dp2a_hi(rlo, W3, dp2a_lo(vhi, W3, dp2a_hi(vlo, W2, dp2a_lo(lhi, W2, RND))));
            const uint32_t o2 = dp2a_hi(rlo, W1, dp2a_lo(vhi, W1, dp2a_hi(vlo, W0, dp2a_lo(lhi, W0, RND))));
            const uint32_t o3 = dp2a_hi(rhi, W3, dp2a_lo(rlo, W3, dp2a_hi(vhi, W2, dp2a_lo(vlo, W2, RND))));
            if (writer) *(uint32_t*)out = __byte_perm(__byte_perm(o0, o1, 0x0062), __byte_perm(o2, o3, 0x0062), 0x5410);
            out += opitch;
        }
    }
}

__global__ void __launch_bounds__(BLUR_NT) k_blur7(const __grid_constant__ OrbPlan plan, const OrbBatch io, const BlurTaps T)
{
    __shared__ uint32_t ring[BLUR_NT / 32][7][2][32];
    const int frame = blockIdx.y, lane = threadIdx.x & 31;
    const int wt = blockIdx.x * (BLUR_NT / 32) + (threadIdx.x >> 5);       // this warp's tile within the frame
    if (wt >= plan.total_blur_tiles) return;
    int l = 0;
    while (l + 1 < plan.nlevels && wt >= plan.lv[l + 1].blur_tile_first) ++l;
    const int w = plan.lv[l].w, h = plan.lv[l].h, opitch = plan.lv[l].pitch, btx = plan.lv[l].blur_tiles_x;
    const int t = wt - plan.lv[l].blur_tile_first;
    const int tyi = __float2int_rz(__fmul_rn((float)t + 0.5f, __frcp_rn((float)btx)));
    const int tx0 = (t - tyi * btx) * ORB_BLUR_TW, yb = tyi * ORB_BLUR_TH;
    int pitch;
    const uint8_t* __restrict__ src = orb_level_ptr(plan, io, frame, l, &pitch);
    const int ye = min(yb + ORB_BLUR_TH, h);
    const int x = tx0 - 4 + 4 * lane;                                      // lane owns pixels x .. x+3
    uint8_t* __restrict__ out = io.blur + (size_t)frame * plan.blur_bytes + plan.lv[l].blur_off + (size_t)yb * opitch + x;
    const bool writer = lane >= 1 && lane <= ORB_BLUR_TW / 4 && x < w;
    const bool inner = __all_sync(0xffffffffu, x >= 0 && x + 3 < w);       // no lane reflects a column
    const bool aligned = ((((uintptr_t)src) | (unsigned)pitch) & 3) == 0;  // x is a multiple of 4
    if (inner && aligned) blur_walk<true>(T, ring[threadIdx.x >> 5], lane, src, pitch, w, h, out, opitch, x, yb, ye, writer);
    else blur_walk<false>(T, ring[threadIdx.x >> 5], lane, src, pitch, w, h, out, opitch, x, yb, ye, writer);
}

// ------------------------------------------------------------------------------ border
// (w+2b) x (h+2b) REFLECT_101-padded copy of one level (mvImagePyramid's parent buffer).
__global__ void k_border(const uint8_t* __restrict__ src, int w, int h, int spitch, uint8_t* __restrict__ dst, int dpitch, int b)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w + 2 * b || y >= h + 2 * b) return;
    dst[(size_t)y * dpitch + x] = src[(size_t)orb_refl101(y - b, h) * spitch + orb_refl101(x - b, w)];
}

// All levels of `frames` frames at once (the single-call path: one launch, one device-to-host copy): level l of frame f
// goes to dst + f * job.frame_bytes + job.off[l] with row pitch job.pitch[l]; one thread writes 4 adjacent bytes.
__global__ void __launch_bounds__(256) k_border_levels(const __grid_constant__ OrbPlan plan, const OrbBatch io, const __grid_constant__ OrbBorderJob job, uint8_t* __restrict__ dst)
{
    const int l = blockIdx.z % plan.nlevels, frame = blockIdx.z / plan.nlevels;
    const OrbLevel& L = plan.lv[l];
    const int b = ORB_EDGE, bw = L.w + 2 * b, bh = L.h + 2 * b;
    const int x4 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x4 >= bw || y >= bh) return;
    int spitch;
    const uint8_t* src = orb_level_ptr(plan, io, frame, l, &spitch);
    const uint8_t* row = src + (size_t)orb_refl101(y - b, L.h) * spitch;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) v |= (uint32_t)__ldg(row + orb_refl101(min(x4 + i, bw - 1) - b, L.w)) << (8 * i);
    *(uint32_t*)(dst + (size_t)frame * job.frame_bytes + job.off[l] + (size_t)y * job.pitch[l] + x4) = v;   // pitch is a multiple of 4, the tail bytes are padding
}

cudaError_t orb_launch_border_levels(const OrbPlan& plan, const OrbBatch& io, const OrbBorderJob& job, uint8_t* dst, int frames, cudaStream_t st)
{
    const int bw = plan.lv[0].w + 2 * ORB_EDGE, bh = plan.lv[0].h + 2 * ORB_EDGE;
    dim3 blk(32, 8), grd(((bw + 3) / 4 + 31) / 32, (bh + 7) / 8, plan.nlevels * frames);
    k_border_levels<<<grd, blk, 0, st>>>(plan, io, job, dst);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------ launchers
cudaError_t orb_launch_pyramid_level(const OrbPlan& plan, const OrbBatch& io, int batch, int l, cudaStream_t st)
{
    const OrbLevel& D = plan.lv[l];
    const OrbLevel& S = plan.lv[l - 1];
    // the walker needs the 4 destination pixels of a lane to span at most 7 source pixels
    // ... and every source row to end at most one destination row (no upsampling)
    const bool walker = 4LL * S.w <= 5LL * D.w && S.w >= 12 && S.h >= D.h;
    if (walker) {
        // rows per warp: as many as still leave every SM a few dozen warps (each row waits for its loads)
        int th = PYR_TH, tiles;
        for (;; th >>= 1) {
            tiles = ((D.w + PYR_TW - 1) / PYR_TW) * ((D.h + th - 1) / th);
            if (th <= 8 || (long long)tiles * batch >= 148LL * PYR_MIN_WARPS) break;
        }
        PyrJob J;
        if (l == 1) { J.src = io.img0; J.src_stride = io.img0_stride; J.spitch = io.img0_pitch; }
        else { J.src = io.pyr + S.img_off; J.src_stride = plan.pyr_bytes; J.spitch = S.pitch; }
        J.dst = io.pyr + D.img_off; J.dst_stride = plan.pyr_bytes; J.dpitch = D.pitch;
        J.sw = S.w; J.sh = S.h; J.dw = D.w; J.dh = D.h;
        J.xtab = (const int2*)io.taps + D.xtab; J.ytab = (const int4*)(io.taps + D.ytab4);
        J.th = th; J.tiles_x = (D.w + PYR_TW - 1) / PYR_TW; J.tiles_y = (D.h + th - 1) / th;
        k_pyr_resize<<<dim3((tiles + PYR_NT / 32 - 1) / (PYR_NT / 32), 1, batch), PYR_NT, 0, st>>>(J);
    } else {
        dim3 blk(32, 8), grd((D.pitch / 4 + 31) / 32, (D.h + 7) / 8, batch);
        k_pyr_resize_generic<<<grd, blk, 0, st>>>(plan, io, l);
    }
    return cudaGetLastError();
}

cudaError_t orb_launch_pyramid(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    for (int l = 1; l < plan.nlevels; ++l) {
        const cudaError_t e = orb_launch_pyramid_level(plan, io, batch, l, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int orb_pyramid_launch_count(const OrbPlan& plan) { return plan.nlevels - 1; }   // one launch per level: a level is resized from the ROUNDED level above

cudaError_t orb_launch_blur(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st)
{
    const BlurTaps T = { 0x30221200u, 0x12223038u, 0x38302212u, 0x00122230u, 32768u };
    k_blur7<<<dim3((plan.total_blur_tiles + BLUR_NT / 32 - 1) / (BLUR_NT / 32), batch), BLUR_NT, 0, st>>>(plan, io, T);
    return cudaGetLastError();
}

cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st)
{
    dim3 blk(32, 8), grd((w + 2 * b + 31) / 32, (h + 2 * b + 7) / 8);
    k_border<<<grd, blk, 0, st>>>(src, w, h, spitch, dst, dpitch, b);
    return cudaGetLastError();
}
