// orb_stereo.cu -- Frame::ComputeStereoMatches (src/Frame.cc:513-699) for a batch of rectified pairs, sm_100a.
// The pairs never leave the GPU between extraction and depth: the kernels read the keypoints, descriptors and
// pyramids exactly where the extractor left them.
//
//   k_stereo_rows   one block per pair: right keypoints bucketed by integer row (counting sort in shared
//                   memory) into 16-byte records {x, band rows minr/maxr (:535-539), index, octave}, so the band
//                   scan reads them with one coalesced load per lane.  Stands in for vRowIndices (:529-544): the reference lists every right keypoint
//                   under each row of its band [floor(y-r), ceil(y+r)], r = 2*scale[octave]; here a left
//                   keypoint scans the buckets of rows rowL-band .. rowL+band (a superset) and applies the
//                   exact band test, so the candidate SET is the reference's.  Candidate order does not
//                   matter: the running "dist < bestDist" (:585-589) keeps the smallest distance and, among
//                   equals, the smallest iR -- the minimum of dist<<16 | iR.
//   k_stereo_match  one warp per left keypoint: band scan with 256-bit Hamming (:546-596), 11x11 SAD over 11
//                   shifts on the level images (:599-648), parabola (:650-663), disparity/depth (:666-679).
//                   The left patch and the 11 x 21 right strip are staged once in shared memory (per warp); the 11
//                   SADs read them from there.
//   k_stereo_cut    one block per pair: median of the SADs by two 256-bin histogram passes (SAD < 2^16),
//                   threshold 1.5f*1.4f*median (:685-698).
// Integer work except the sub-pixel parabola, which is evaluated operation by operation (_rn).
#include "orb_device.cuh"
#include "orb_stereo.h"

#define TH_HIGH 100     // src/ORBmatcher.cc:37
#define TH_LOW 50       // :38

// REFLECT_101 for indices at most n-1 outside [0, n): one reflection, branch free
__device__ __forceinline__ int st_refl1(const int i, const int n)
{
    const int a = abs(i);
    return min(a, 2 * (n - 1) - a);
}

__device__ __forceinline__ int st_hamming256(const uint32_t* a, const uint4 b0, const uint4 b1)
{
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

__global__ void __launch_bounds__(256) k_stereo_rows(const __grid_constant__ OrbStereoView V)
{
    extern __shared__ int s_cnt[];          // [h0 + 1] counts, then running cursors
    __shared__ int s_scan[8];
    const int pair = blockIdx.x, tid = threadIdx.x, h0 = V.h[0];
    const int nr = min(V.nr[(size_t)pair * V.nstride], V.cap);
    const orbx_kp* kr = V.kr + (size_t)pair * V.kstride;
    int* row_start = V.row_start + (size_t)pair * (h0 + 2);
    uint4* rec = V.rec + (size_t)pair * V.cap;
    for (int i = tid; i <= h0; i += 256) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < nr; i += 256) atomicAdd(&s_cnt[min(max((int)kr[i].y, 0), h0 - 1)], 1);
    __syncthreads();
    orb_block_scan_incl<256>(s_cnt, h0 + 1, s_scan);                    // s_cnt[r] = keypoints in rows 0..r
    for (int i = tid; i <= h0; i += 256) row_start[i + 1] = s_cnt[i];
    if (tid == 0) row_start[0] = 0;
    __syncthreads();
    // scatter: a row's slots are filled from its end (inclusive count) downwards
    for (int i = tid; i < nr; i += 256) {
        const orbx_kp k = kr[i];
        const float r = __fmul_rn(2.0f, V.scale[k.octave]);
        const int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));       // :535-539
        const int lo = min(max(minr, -32768), 32767), hi = min(max(maxr, -32768), 32767);
        rec[atomicSub(&s_cnt[min(max((int)k.y, 0), h0 - 1)], 1) - 1] =
            make_uint4(__float_as_uint(k.x), (uint32_t)(lo & 0xffff) | ((uint32_t)hi << 16), (uint32_t)i | ((uint32_t)k.octave << 16), 0u);
    }
}

#ifndef ST_NT
#define ST_NT 64         // threads per block of k_stereo_match (a warp per left keypoint); per 1024 frames: 64 threads 0.87 ms, 128 0.88, 256 0.92
#endif
__global__ void __launch_bounds__(256) k_stereo_match(const __grid_constant__ OrbStereoView V)
{
    __shared__ __align__(16) uint8_t s_patch[8][384];      // per warp: left patch [121] at 0, right strip [11][21] at 128
    const int pair = blockIdx.y, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nl = min(V.nl[(size_t)pair * V.nstride], V.cap);
    if (iL >= nl) return;
    const int nr = min(V.nr[(size_t)pair * V.nstride], V.cap);
    const orbx_kp* kr = V.kr + (size_t)pair * V.kstride;
    const uint32_t* dr = V.dr + (size_t)pair * V.kstride * 8;
    float out_u = -1.0f, out_d = -1.0f; int out_s = -1;
    const orbx_kp kp = V.kl[(size_t)pair * V.kstride + iL];
    const int levelL = kp.octave, rowL = (int)kp.y;
    const float uL = kp.x;
    const float minD = 0.f, maxD = __fdiv_rn(V.bf, V.mb);                       // :523-524
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);        // :560-561
    uint32_t best = 0xffffffffu;     // dist << 16 | iR
    if (!(maxU < 0) && nr > 0) {                                               // :563
        uint32_t d[8];
        const uint32_t* dl = V.dl + ((size_t)pair * V.kstride + iL) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = __ldg(dl + i);
        const int h0 = V.h[0];
        const int* row_start = V.row_start + (size_t)pair * (h0 + 2);
        const uint4* rec = V.rec + (size_t)pair * V.cap;
        // candidates sit at most one level away: their band is at most 2 * (largest scale of the three levels) rows
        const int band = orb_stereo_band(fmaxf(V.scale[levelL], fmaxf(V.scale[max(levelL - 1, 0)], V.scale[min(levelL + 1, V.nlevels - 1)])));
        const int r_lo = min(max(rowL - band, 0), h0 - 1), r_hi = min(max(rowL + band, 0), h0 - 1);
        const int c0 = row_start[r_lo], c1 = row_start[r_hi + 1];
        for (int base = c0; base < c1; base += 32) {
            const int c = base + lane;
            if (c < c1) {
                const uint4 R = __ldg(rec + c);
                const float x = __uint_as_float(R.x);
                const int minr = (int)(short)(R.y & 0xffffu), maxr = (int)R.y >> 16;
                const int iR = (int)(R.z & 0xffffu), oct = (int)(R.z >> 16);
                if (rowL >= minr && rowL <= maxr && !(oct < levelL - 1 || oct > levelL + 1) && x >= minU && x <= maxU) {
                    const uint4* p = (const uint4*)(dr + (size_t)iR * 8);
                    const int dist = st_hamming256(d, __ldg(p), __ldg(p + 1));
                    if (dist < TH_HIGH) best = min(best, ((uint32_t)dist << 16) | (uint32_t)iR);   // bestDist starts at TH_HIGH (:568)
                }
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
    }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                              // :517
    if (best != 0xffffffffu && (int)(best >> 16) < thOrbDist) {
        const int bestIdxR = (int)(best & 0xffffu);
        const float uR0 = kr[bestIdxR].x;
        const float sf = V.inv_scale[levelL];
        const float scaleduL = roundf(__fmul_rn(kp.x, sf)), scaledvL = roundf(__fmul_rn(kp.y, sf)), scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, L = 5;
        const int lw = V.w[levelL], lh = V.h[levelL];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;    // :624-625
        if (!(iniu < 0 || endu >= (float)lw) && lw > 16 && lh > 16) {   // (levels are never that small: orbx_shape_supported)
            const uint8_t* IL = V.l[levelL] + (size_t)pair * V.lstride[levelL];
            const uint8_t* IR = V.r[levelL] + (size_t)pair * V.rstride[levelL];
            const int lp = V.lpitch[levelL], rp = V.rpitch[levelL];
            const int r0 = (int)(scaledvL - w), cL0 = (int)(scaleduL - w), cRm = (int)(scaleduR0 - w);
            // Reads may leave the level image by up to 10 px on the left: that is the REFLECT_101 border of
            // mvImagePyramid (never materialised here).  Stage the 11x11 left patch and the 11x21 right strip
            // (columns scaleduR0-10 .. scaleduR0+10) once; lane owns patch positions p = lane + 32k.
            uint8_t* sL = s_patch[threadIdx.x >> 5];
            uint8_t* sR = sL + 128;
            __syncwarp();
            // (almost every patch lies inside the level: then no index needs reflecting -- the same bytes with a third of the instructions)
            const bool inside = r0 >= 0 && r0 + 10 < lh && cL0 >= 0 && cL0 + 10 < lw && cRm - L >= 0 && cRm - L + 20 < lw;
            if (inside) {
                const uint8_t* pl = IL + (size_t)r0 * lp + cL0;
                const uint8_t* pr = IR + (size_t)r0 * rp + (cRm - L);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int p = lane + 32 * k;
                    if (p < 121) { const int yy = p / 11, xx = p - yy * 11; sL[p] = pl[yy * lp + xx]; }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = lane + 32 * k;
                    if (q < 231) { const int yy = q / 21, xx = q - yy * 21; sR[q] = pr[yy * rp + xx]; }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int p = lane + 32 * k;
                    if (p < 121) { const int yy = p / 11, xx = p - yy * 11; sL[p] = IL[(size_t)st_refl1(r0 + yy, lh) * lp + st_refl1(cL0 + xx, lw)]; }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int q = lane + 32 * k;
                    if (q < 231) { const int yy = q / 21, xx = q - yy * 21; sR[q] = IR[(size_t)st_refl1(r0 + yy, lh) * rp + st_refl1(cRm - L + xx, lw)]; }
                }
            }
            __syncwarp();
            const int cL = sL[5 * 11 + 5];
            int a[4], ro[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int p = min(lane + 32 * k, 120), yy = p / 11, xx = p - yy * 11;
                a[k] = (int)sL[p] - cL;
                ro[k] = yy * 21 + xx;
            }
            const bool last = lane + 96 < 121;                                 // k = 3 exists for lanes 0..24
            int bestDist = 2147483647, bestincR = 0;
            float vDists[11];
#pragma unroll
            for (int i = 0; i < 11; ++i) {
                const int cR = sR[5 * 21 + 5 + i];
                int s = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k)                                   // |(l - cL) - (r - cR)| = |(l - cL + cR) - r|
                    if (k < 3 || last) s = (int)__sad(a[k] + cR, (int)sR[ro[k] + i], (unsigned)s);
                s = __reduce_add_sync(0xffffffffu, s);
                const float dist = (float)s;
                if (dist < (float)bestDist) { bestDist = (int)dist; bestincR = i - L; }   // :641-645
                vDists[i] = dist;
            }
            if (!(bestincR == -L || bestincR == L)) {                          // :650
                float dist1 = 0, dist2 = 0, dist3 = 0;
#pragma unroll
                for (int i = 0; i < 11; ++i) {   // (indexing with a runtime value would spill the array)
                    if (i == L + bestincR - 1) dist1 = vDists[i];
                    if (i == L + bestincR) dist2 = vDists[i];
                    if (i == L + bestincR + 1) dist3 = vDists[i];
                }
                const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(V.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)bestincR), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= minD && disparity < maxD) {
                        if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)((double)uL - 0.01); }
                        out_d = __fdiv_rn(V.bf, disparity); out_u = bestuR; out_s = bestDist;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        const size_t o = (size_t)pair * V.ostride + iL;
        V.u_right[o] = out_u; V.depth[o] = out_d; V.sad[(size_t)pair * V.cap + iL] = out_s;
    }
}

// Median cut (:685-698): threshold = 1.5f*1.4f*median of the SADs, median = element nd/2 of the sorted list.
__global__ void __launch_bounds__(256) k_stereo_cut(const __grid_constant__ OrbStereoView V)
{
    __shared__ int s_hist[256], s_nd, s_bin, s_rank;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int nl = min(V.nl[(size_t)pair * V.nstride], V.cap);
    const int* sad = V.sad + (size_t)pair * V.cap;
    float* u_right = V.u_right + (size_t)pair * V.ostride;
    float* depth = V.depth + (size_t)pair * V.ostride;
    s_hist[tid] = 0;
    if (tid == 0) s_nd = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nl; i += 256) { const int v = sad[i]; if (v >= 0) { ++mine; atomicAdd(&s_hist[(v >> 8) & 255], 1); } }
    if (mine) atomicAdd(&s_nd, mine);
    __syncthreads();
    const int nd = s_nd;
    if (nd == 0) { if (tid == 0) V.n_stereo[pair] = 0; return; }
    if (tid == 0) {                                    // bin of the element with rank nd/2
        int k = nd / 2, b = 0;
        while (k >= s_hist[b]) { k -= s_hist[b]; ++b; }
        s_bin = b; s_rank = k;
    }
    __syncthreads();
    const int hi = s_bin, k2 = s_rank;
    __syncthreads();
    s_hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += 256) { const int v = sad[i]; if (v >= 0 && ((v >> 8) & 255) == hi) atomicAdd(&s_hist[v & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int k = k2, b = 0;
        while (k >= s_hist[b]) { k -= s_hist[b]; ++b; }
        s_bin = (hi << 8) | b;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)s_bin);
    for (int i = tid; i < nl; i += 256)
        if (sad[i] >= 0 && !((float)sad[i] < thDist)) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    if (tid == 0) V.n_stereo[pair] = nd;
}

cudaError_t orb_launch_stereo(const OrbStereoView& V, int pairs, int max_left, cudaStream_t st)
{
    if (pairs <= 0 || max_left <= 0) return cudaSuccess;
    const size_t smem = (size_t)(V.h[0] + 1) * sizeof(int);
    k_stereo_rows<<<pairs, 256, smem, st>>>(V);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_stereo_match<<<dim3((max_left + ST_NT / 32 - 1) / (ST_NT / 32), pairs), ST_NT, 0, st>>>(V);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_stereo_cut<<<pairs, 256, 0, st>>>(V);
    return cudaGetLastError();
}
