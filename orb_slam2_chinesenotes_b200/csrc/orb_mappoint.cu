// orb_mappoint.cu -- the map-point side of the matching path, sm_100a:
//   k_project_points   Frame::isInFrustum (src/Frame.cc:288-345) with Frame::UpdatePoseMatrices' camera centre
//                      (:280-285) and MapPoint::PredictScale (src/MapPoint.cc:459-475) for every (frame, map point):
//                      writes the tracking fields SearchByProjection(Frame, MapPoints) reads (mbTrackInView,
//                      mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos) in the layout of orbm_points, so the
//                      projection feeds orbm_search_by_projection_points_batch without leaving the GPU
//   k_distinctive      MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:275-340): all-pairs Hamming
//                      distances inside each map point's observation set, median per row, first smallest median
// Float semantics follow OpenCV as the reference uses it: 3x3 * 3x1 products accumulated left to right in float,
// translation added last; cv::norm and Mat::dot accumulate in double; glibc logf restated in FP64.  No FMA
// contraction anywhere (-fmad=false and explicit _rn intrinsics).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "../../include/orb_b200.h"

// ------------------------------------------------------------------------------------------ logf (glibc 2.39)
// sysdeps/ieee754/flt-32/e_logf.c: 16-entry table, degree-3 polynomial, double arithmetic in the source's order.
__constant__ double c_logf_tab[16][2] = {
    { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
    { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
    { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
    { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
    { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
    { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
    { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
    { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },
};

__device__ __forceinline__ float glibc_logf(float x)
{
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return NAN;
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));            // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    const double invc = c_logf_tab[i][0], logc = c_logf_tab[i][1], z = (double)__uint_as_float(iz);
    const double r = __dsub_rn(__dmul_rn(z, invc), 1.0);
    const double y0 = __dadd_rn(logc, __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
    y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return (float)y;
}

// ------------------------------------------------------------------------------------------ projection
struct ProjParams {
    const float* Tcw;                  // [nprob][16] row-major 4x4
    float fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y, log_scale, cos_limit;
    int nlevels;
    const int* nq; int nq_stride; size_t pt_stride;      // pt_stride = 0: one point set shared by all frames
    const float* xyz; const float* normal; const float* max_d; const float* min_d;
    uint8_t* in_view; float* proj; int* level; float* view_cos; int* n_in_view;
};

__global__ void __launch_bounds__(256) k_project_points(const ProjParams P)
{
    __shared__ float T[12], Ow[3];
    __shared__ int s_cnt;
    const int prob = blockIdx.y;
    if (threadIdx.x < 12) T[threadIdx.x] = P.Tcw[(size_t)prob * 16 + threadIdx.x];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (threadIdx.x < 3) {       // mOw = -mRcw.t() * mtcw, src/Frame.cc:284: negate, then products left to right
        const int i = threadIdx.x;
        float s = __fmul_rn(-T[0 * 4 + i], T[3]);
        s = __fadd_rn(s, __fmul_rn(-T[1 * 4 + i], T[7]));
        s = __fadd_rn(s, __fmul_rn(-T[2 * 4 + i], T[11]));
        Ow[i] = s;
    }
    __syncthreads();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int nq = min(P.nq[prob], P.nq_stride);
    bool ok = q < nq;
    if (ok) {
        const size_t pi = (size_t)prob * P.pt_stride + q, oi = (size_t)prob * P.nq_stride + q;
        const float X = P.xyz[3 * pi], Y = P.xyz[3 * pi + 1], Z = P.xyz[3 * pi + 2];
        float Pc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {                                       // :295
            float s = __fmul_rn(T[4 * r], X);
            s = __fadd_rn(s, __fmul_rn(T[4 * r + 1], Y));
            s = __fadd_rn(s, __fmul_rn(T[4 * r + 2], Z));
            Pc[r] = __fadd_rn(s, T[4 * r + 3]);
        }
        float u = 0, v = 0, invz = 0, dist = 0, vc = 0;
        ok = !(Pc[2] < 0.0f);                                               // :301
        if (ok) {
            invz = __fdiv_rn(1.0f, Pc[2]);
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, Pc[0]), invz), P.cx);   // :305-306
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, Pc[1]), invz), P.cy);
            ok = !(u < P.min_x || u > P.max_x) && !(v < P.min_y || v > P.max_y);
        }
        if (ok) {
            const float maxD = __fmul_rn(1.2f, P.max_d[pi]), minD = __fmul_rn(0.8f, P.min_d[pi]);   // src/MapPoint.cc:424-435
            const float PO0 = __fsub_rn(X, Ow[0]), PO1 = __fsub_rn(Y, Ow[1]), PO2 = __fsub_rn(Z, Ow[2]);
            double s2 = __dmul_rn((double)PO0, (double)PO0);
            s2 = __dadd_rn(s2, __dmul_rn((double)PO1, (double)PO1));
            s2 = __dadd_rn(s2, __dmul_rn((double)PO2, (double)PO2));
            dist = (float)__dsqrt_rn(s2);                                   // :318
            ok = !(dist < minD || dist > maxD);
            if (ok) {
                double dot = __dmul_rn((double)PO0, (double)P.normal[3 * pi]);
                dot = __dadd_rn(dot, __dmul_rn((double)PO1, (double)P.normal[3 * pi + 1]));
                dot = __dadd_rn(dot, __dmul_rn((double)PO2, (double)P.normal[3 * pi + 2]));
                vc = (float)__ddiv_rn(dot, (double)dist);                   // :324
                ok = !(vc < P.cos_limit);
            }
        }
        P.in_view[oi] = ok ? 1 : 0;
        if (ok) {
            // MapPoint::PredictScale(dist, Frame*), src/MapPoint.cc:459-475
            const float ratio = __fdiv_rn(P.max_d[pi], dist);
            int nScale = (int)ceilf(__fdiv_rn(glibc_logf(ratio), P.log_scale));
            if (nScale < 0) nScale = 0;
            else if (nScale >= P.nlevels) nScale = P.nlevels - 1;
            P.proj[3 * oi] = u; P.proj[3 * oi + 1] = v; P.proj[3 * oi + 2] = __fsub_rn(u, __fmul_rn(P.bf, invz));
            P.level[oi] = nScale;
            P.view_cos[oi] = vc;
        }
    }
    if (P.n_in_view) {
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
        __syncthreads();
        if (threadIdx.x == 0 && s_cnt) atomicAdd(P.n_in_view + prob, s_cnt);
    }
}

// ------------------------------------------------------------------------------------------ motion-model projection
// The per-point part of ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:172-244:
// last-frame map points into the current image, search radius by the last keypoint's octave, level range by the
// direction of motion.  Writes the orbm_windows arrays the best-candidate kernel reads.
struct LastParams {
    const float* Tcw_cur; const float* Tcw_last;      // [nprob][16]
    float fx, fy, cx, cy, bf, mb, min_x, max_x, min_y, max_y, th;
    int mono, nlevels;
    float scale[32];
    const int* n_last; int stride;
    const orbx_kp* kps_last; const uint8_t* last_mp; const uint8_t* last_outlier; const float* last_xyz;
    float* uvr; int* min_level; int* max_level; float* ur; float* er_max; uint8_t* valid; float* q_angle;
};

__global__ void __launch_bounds__(256) k_project_last_frame(const LastParams P)
{
    __shared__ float T[12], tlcz;
    const int prob = blockIdx.y;
    if (threadIdx.x < 12) T[threadIdx.x] = P.Tcw_cur[(size_t)prob * 16 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        // twc = -Rcw.t() * tcw (:177), tlc = Rlw * twc + tlw (:181): cv::gemm order, products left to right
        const float* L = P.Tcw_last + (size_t)prob * 16;
        float twc[3];
        for (int i = 0; i < 3; ++i) {
            float s = __fmul_rn(-T[0 * 4 + i], T[3]);
            s = __fadd_rn(s, __fmul_rn(-T[1 * 4 + i], T[7]));
            s = __fadd_rn(s, __fmul_rn(-T[2 * 4 + i], T[11]));
            twc[i] = s;
        }
        float s = __fmul_rn(L[8], twc[0]);
        s = __fadd_rn(s, __fmul_rn(L[9], twc[1]));
        s = __fadd_rn(s, __fmul_rn(L[10], twc[2]));
        tlcz = __fadd_rn(s, L[11]);
    }
    __syncthreads();
    const bool bForward = tlcz > P.mb && !P.mono, bBackward = -tlcz > P.mb && !P.mono;     // :184-185
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(P.n_last[prob], P.stride)) return;
    const size_t o = (size_t)prob * P.stride + i;
    const orbx_kp kp = P.kps_last[o];
    P.q_angle[o] = kp.angle;
    bool ok = P.last_mp[o] && !(P.last_outlier && P.last_outlier[o]);                       // :190-193
    float u = 0, v = 0, invzc = 0;
    if (ok) {
        const float X = P.last_xyz[3 * o], Y = P.last_xyz[3 * o + 1], Z = P.last_xyz[3 * o + 2];
        float pc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float s = __fmul_rn(T[4 * r], X);
            s = __fadd_rn(s, __fmul_rn(T[4 * r + 1], Y));
            s = __fadd_rn(s, __fmul_rn(T[4 * r + 2], Z));
            pc[r] = __fadd_rn(s, T[4 * r + 3]);
        }
        invzc = (float)__ddiv_rn(1.0, (double)pc[2]);                                       // :199: 1.0 / float
        ok = !(invzc < 0);
        if (ok) {
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, pc[0]), invzc), P.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, pc[1]), invzc), P.cy);
            ok = !(u < P.min_x || u > P.max_x) && !(v < P.min_y || v > P.max_y) && kp.octave >= 0 && kp.octave < P.nlevels;
        }
    }
    P.valid[o] = ok ? 1 : 0;
    if (!ok) { P.uvr[3 * o] = 0; P.uvr[3 * o + 1] = 0; P.uvr[3 * o + 2] = 0; P.min_level[o] = 0; P.max_level[o] = 0; P.ur[o] = 0; P.er_max[o] = 0; return; }
    const int oct = kp.octave;
    const float r = __fmul_rn(P.th, P.scale[oct]);                                           // :215
    P.uvr[3 * o] = u; P.uvr[3 * o + 1] = v; P.uvr[3 * o + 2] = r;
    if (bForward) { P.min_level[o] = oct; P.max_level[o] = -1; }                             // :219-224
    else if (bBackward) { P.min_level[o] = 0; P.max_level[o] = oct; }
    else { P.min_level[o] = oct - 1; P.max_level[o] = oct + 1; }
    P.ur[o] = __fsub_rn(u, __fmul_rn(P.bf, invzc));                                          // :241-244
    P.er_max[o] = r;
}

// ------------------------------------------------------------------------------------------ distinctive descriptor
#define DD_NT 128
#define DD_MAX 4096      // observations per map point the kernel handles

__global__ void __launch_bounds__(DD_NT) k_distinctive(const uint4* __restrict__ desc, const int* __restrict__ offsets,
                                                      const uint8_t* __restrict__ bad, int* __restrict__ best_idx, int* __restrict__ best_median)
{
    __shared__ uint16_t s_idx[DD_MAX];
    __shared__ int s_hist[DD_NT / 32][288];
    __shared__ int s_n;
    __shared__ unsigned s_best;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o0 = offsets[p], n = offsets[p + 1] - o0;
    if (tid == 0) { s_n = 0; s_best = 0xffffffffu; }
    __syncthreads();
    if (n > DD_MAX) { if (tid == 0) { best_idx[p] = -2; if (best_median) best_median[p] = -1; } return; }
    // usable observations in the caller's order (KeyFrame::isBad ones are skipped, :299): ordered compaction
    for (int base = 0; base < n; base += DD_NT) {
        const int i = base + tid;
        const bool ok = i < n && !(bad && bad[o0 + i]);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        __shared__ int s_w[DD_NT / 32];
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = s_n;
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (ok) s_idx[before + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < DD_NT / 32; ++w) t += s_w[w]; s_n += t; }
        __syncthreads();
    }
    const int N = s_n;
    if (N == 0) { if (tid == 0) { best_idx[p] = -1; if (best_median) best_median[p] = -1; } return; }
    const int k = (N - 1) >> 1;                       // vDists[0.5*(N-1)], :327
    int* hist = s_hist[warp];
    unsigned mine = 0xffffffffu;
    for (int i = warp; i < N; i += DD_NT / 32) {
        for (int b = lane; b < 288; b += 32) hist[b] = 0;
        __syncwarp();
        const uint4 a0 = desc[2 * (size_t)(o0 + s_idx[i])], a1 = desc[2 * (size_t)(o0 + s_idx[i]) + 1];
        for (int j = lane; j < N; j += 32) {
            const uint4 b0 = desc[2 * (size_t)(o0 + s_idx[j])], b1 = desc[2 * (size_t)(o0 + s_idx[j]) + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            atomicAdd(&hist[d], 1);
        }
        __syncwarp();
        // (k+1)-th smallest: lane l owns bins 9l .. 9l+8
        int c = 0;
#pragma unroll
        for (int b = 0; b < 9; ++b) c += hist[9 * lane + b];
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        const unsigned owner = __ballot_sync(0xffffffffu, incl >= k + 1);
        const int ol = __ffs(owner) - 1;
        int median = 0;
        if (lane == ol) {
            int acc = incl - c;
            for (int b = 0; b < 9; ++b) { acc += hist[9 * lane + b]; if (acc >= k + 1) { median = 9 * lane + b; break; } }
        }
        median = __shfl_sync(0xffffffffu, median, ol);
        const unsigned key = ((unsigned)median << 16) | (unsigned)i;          // first smallest median wins (:329)
        mine = min(mine, key);
        __syncwarp();
    }
    if (lane == 0) atomicMin(&s_best, mine);
    __syncthreads();
    if (tid == 0) { best_idx[p] = (int)s_idx[s_best & 0xffffu]; if (best_median) best_median[p] = (int)(s_best >> 16); }
}

// ================================================================================ host side
// ------------------------------------------------------------------------------------------ projection for Fuse / SearchBySim3
// The per-point prologue of ORBmatcher::Fuse (src/ORBmatcher.cc:1388-1426, :1546-1584) and of one direction of SearchBySim3
// (:886-909 / :937-960) for every (key frame, map point): what orbm_window_best_free_batch searches, written in its layout, so
// projection -> search stays on the GPU.  Differences from Frame::isInFrustum above, all kept: x = Pc.x * invz first, then
// u = fx * x + cx (:1395-1399); KeyFrame::IsInImage with the key frame's int bounds (u >= min && u < max); the viewing-angle
// test in double, PO.dot(Pn) < 0.5 * dist3D (:1420), and no such test in SearchBySim3, whose distance is the norm of the point in
// the target camera (:899); MapPoint::PredictScale(dist, KeyFrame*) (src/MapPoint.cc:442-457) is the same arithmetic as the
// Frame overload.  pose [nprob][24]: Fuse R(9) t(3) Ow(3); SearchBySim3 R_a(9) t_a(3) sR(9) tt(3) (p = sR * (R_a * x + t_a) + tt).
struct FuseProjParams {
    const float* pose; int sim3;
    float fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y, log_scale, th;
    int nlevels;
    float scale[32];
    const int* nq; int nq_stride; size_t pt_stride;
    const float* xyz; const float* normal; const float* max_d; const float* min_d; const uint8_t* skip;
    float* uvr; int* level; float* ur; uint8_t* valid;
};

__global__ void __launch_bounds__(256) k_fuse_project(const FuseProjParams P)
{
    __shared__ float T[24];
    const int prob = blockIdx.y;
    if (threadIdx.x < 24) T[threadIdx.x] = P.pose[(size_t)prob * 24 + threadIdx.x];
    __syncthreads();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P.nq_stride) return;
    const size_t pi = (size_t)prob * P.pt_stride + q, oi = (size_t)prob * P.nq_stride + q;
    bool ok = q < P.nq[prob] && !(P.skip && P.skip[pi]);
    float u = 0.f, v = 0.f, r = 0.f, urv = 0.f;
    int lvl = 0;
    if (ok) {
        const float X = P.xyz[3 * pi], Y = P.xyz[3 * pi + 1], Z = P.xyz[3 * pi + 2];
        float Pc[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s = __fmul_rn(T[3 * k], X);
            s = __fadd_rn(s, __fmul_rn(T[3 * k + 1], Y));
            s = __fadd_rn(s, __fmul_rn(T[3 * k + 2], Z));
            Pc[k] = __fadd_rn(s, T[9 + k]);
        }
        if (P.sim3) {                                                       // through the similarity into the other camera (:888 / :939)
            float Pb[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float s = __fmul_rn(T[12 + 3 * k], Pc[0]);
                s = __fadd_rn(s, __fmul_rn(T[12 + 3 * k + 1], Pc[1]));
                s = __fadd_rn(s, __fmul_rn(T[12 + 3 * k + 2], Pc[2]));
                Pb[k] = __fadd_rn(s, T[21 + k]);
            }
            Pc[0] = Pb[0]; Pc[1] = Pb[1]; Pc[2] = Pb[2];
        }
        ok = !(Pc[2] < 0.0f);
        float invz = 0.f, dist = 0.f;
        if (ok) {
            invz = __fdiv_rn(1.0f, Pc[2]);                                  // 1/z and (float)(1.0/z) are the same float
            u = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(Pc[0], invz)), P.cx);
            v = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(Pc[1], invz)), P.cy);
            ok = u >= P.min_x && u < P.max_x && v >= P.min_y && v < P.max_y;                  // KeyFrame::IsInImage
        }
        if (ok) {
            const float maxD = __fmul_rn(1.2f, P.max_d[pi]), minD = __fmul_rn(0.8f, P.min_d[pi]);   // src/MapPoint.cc:424-435
            float d0, d1, d2;
            if (P.sim3) { d0 = Pc[0]; d1 = Pc[1]; d2 = Pc[2]; }
            else { d0 = __fsub_rn(X, T[12]); d1 = __fsub_rn(Y, T[13]); d2 = __fsub_rn(Z, T[14]); }   // PO = p3Dw - Ow
            double s2 = __dmul_rn((double)d0, (double)d0);
            s2 = __dadd_rn(s2, __dmul_rn((double)d1, (double)d1));
            s2 = __dadd_rn(s2, __dmul_rn((double)d2, (double)d2));
            dist = (float)__dsqrt_rn(s2);
            ok = !(dist < minD || dist > maxD);
            if (ok && !P.sim3) {
                double dot = __dmul_rn((double)d0, (double)P.normal[3 * pi]);
                dot = __dadd_rn(dot, __dmul_rn((double)d1, (double)P.normal[3 * pi + 1]));
                dot = __dadd_rn(dot, __dmul_rn((double)d2, (double)P.normal[3 * pi + 2]));
                ok = !(dot < __dmul_rn(0.5, (double)dist));                  // :1420
            }
        }
        if (ok) {
            const float ratio = __fdiv_rn(P.max_d[pi], dist);               // PredictScale, src/MapPoint.cc:442-457
            int nScale = (int)ceilf(__fdiv_rn(glibc_logf(ratio), P.log_scale));
            if (nScale < 0) nScale = 0;
            else if (nScale >= P.nlevels) nScale = P.nlevels - 1;
            lvl = nScale;
            r = __fmul_rn(P.th, P.scale[lvl & 31]);
            urv = __fsub_rn(u, __fmul_rn(P.bf, invz));
        } else { u = 0.f; v = 0.f; }
    }
    P.uvr[3 * oi] = u; P.uvr[3 * oi + 1] = v; P.uvr[3 * oi + 2] = r;
    P.level[oi] = lvl;
    if (P.ur) P.ur[oi] = urv;
    P.valid[oi] = ok ? 1 : 0;
}

namespace {
int dev_of(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    return a.device;
}
struct DevScope {
    int prev = -1;
    bool enter(int dev) { return cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
    ~DevScope() { if (prev >= 0) cudaSetDevice(prev); }
};
} // namespace

extern "C" {

int orbm_project_points_batch(int nprob, const float* Tcw, const float* K, float bf, float min_x, float max_x, float min_y,
                              float max_y, float scale_factor, int nlevels, float viewing_cos_limit,
                              const int* nq, int nq_stride, int points_shared, const float* xyz, const float* normal,
                              const float* max_distance, const float* min_distance, uint8_t* in_view, float* proj_xyxr,
                              int* level, float* view_cos, int* n_in_view, void* cuda_stream)
{
    if (nprob <= 0 || !Tcw || !K || !nq || nq_stride <= 0 || nlevels <= 0 || !xyz || !normal || !max_distance || !min_distance ||
        !in_view || !proj_xyxr || !level || !view_cos)
        return ORBX_E_ARG;
    const int dev = dev_of(xyz);
    if (dev < 0 || dev_of(Tcw) != dev || dev_of(nq) != dev || dev_of(in_view) != dev) return ORBX_E_ARG;
    DevScope g;
    if (!g.enter(dev)) { cudaGetLastError(); return ORBX_E_CUDA; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    ProjParams P;
    P.Tcw = Tcw; P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = bf;
    P.min_x = min_x; P.max_x = max_x; P.min_y = min_y; P.max_y = max_y;
    P.log_scale = logf(scale_factor);                       // mfLogScaleFactor = log(mfScaleFactor), src/Frame.cc:75 (host libm)
    P.cos_limit = viewing_cos_limit; P.nlevels = nlevels;
    P.nq = nq; P.nq_stride = nq_stride; P.pt_stride = points_shared ? 0 : (size_t)nq_stride;
    P.xyz = xyz; P.normal = normal; P.max_d = max_distance; P.min_d = min_distance;
    P.in_view = in_view; P.proj = proj_xyxr; P.level = level; P.view_cos = view_cos; P.n_in_view = n_in_view;
    if (n_in_view && cudaMemsetAsync(n_in_view, 0, sizeof(int) * (size_t)nprob, st) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    k_project_points<<<dim3((nq_stride + 255) / 256, nprob), 256, 0, st>>>(P);
    if (cudaGetLastError() != cudaSuccess) return ORBX_E_CUDA;
    return ORBX_OK;
}

int orbm_fuse_project_batch(int nprob, int sim3, const float* pose, const float* K, float bf, float min_x, float max_x, float min_y,
                            float max_y, float scale_factor, const float* scale, int nlevels, float th,
                            const int* nq, int nq_stride, int points_shared, const float* xyz, const float* normal,
                            const float* max_distance, const float* min_distance, const uint8_t* skip,
                            float* uvr, int* level, float* ur, uint8_t* valid, void* cuda_stream)
{
    if (nprob <= 0 || !pose || !K || !scale || !nq || nq_stride <= 0 || nlevels <= 0 || nlevels > 32 || !xyz || (!sim3 && !normal) ||
        !max_distance || !min_distance || !uvr || !level || !valid)
        return ORBX_E_ARG;
    const int dev = dev_of(xyz);
    if (dev < 0 || dev_of(pose) != dev || dev_of(nq) != dev || dev_of(uvr) != dev || dev_of(level) != dev || dev_of(valid) != dev) return ORBX_E_ARG;
    DevScope g;
    if (!g.enter(dev)) { cudaGetLastError(); return ORBX_E_CUDA; }
    FuseProjParams P = {};
    P.pose = pose; P.sim3 = sim3 ? 1 : 0;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = bf;
    P.min_x = min_x; P.max_x = max_x; P.min_y = min_y; P.max_y = max_y;
    P.log_scale = logf(scale_factor);                       // mfLogScaleFactor, src/KeyFrame.cc:38 (copied from the Frame's host logf)
    P.th = th; P.nlevels = nlevels;
    for (int i = 0; i < 32; ++i) P.scale[i] = i < nlevels ? scale[i] : 0.f;
    P.nq = nq; P.nq_stride = nq_stride; P.pt_stride = points_shared ? 0 : (size_t)nq_stride;
    P.xyz = xyz; P.normal = normal; P.max_d = max_distance; P.min_d = min_distance; P.skip = skip;
    P.uvr = uvr; P.level = level; P.ur = ur; P.valid = valid;
    k_fuse_project<<<dim3((nq_stride + 255) / 256, nprob), 256, 0, (cudaStream_t)cuda_stream>>>(P);
    if (cudaGetLastError() != cudaSuccess) return ORBX_E_CUDA;
    return ORBX_OK;
}

int orbm_search_by_projection_frame_batch(const orbm_frames* cur, const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                          const float* scale, int nlevels, const int* n_last, int last_stride,
                                          const orbx_kp* kps_last, const uint8_t* last_mp, const uint8_t* last_outlier,
                                          const float* last_xyz, const uint8_t* last_mp_desc, const int* last_mp_obs,
                                          const int* cur_init_obs, int* assign_out, float th, int bMono, int checkOri,
                                          int* nmatches, void* cuda_stream)
{
    if (!cur || cur->nprob <= 0 || !Tcw_cur || !Tcw_last || !K || !scale || nlevels <= 0 || nlevels > 32 || !n_last || last_stride <= 0 ||
        !kps_last || !last_mp || !last_xyz || !last_mp_desc || !assign_out || !nmatches)
        return ORBX_E_ARG;
    const int dev = dev_of(kps_last);
    if (dev < 0 || dev_of(Tcw_cur) != dev || dev_of(Tcw_last) != dev || dev_of(last_xyz) != dev || dev_of(n_last) != dev) return ORBX_E_ARG;
    DevScope g;
    if (!g.enter(dev)) { cudaGetLastError(); return ORBX_E_CUDA; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int nprob = cur->nprob;
    const size_t nq = (size_t)nprob * last_stride;
    // workspace: uvr (3 f), min/max level (2 i), ur, er_max, q_angle (3 f), valid (1 b)
    const size_t bytes = nq * (3 + 2 + 3) * 4 + nq;
    char* ws = nullptr;
    if (cudaMallocAsync((void**)&ws, bytes + 64, st) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    LastParams P;
    P.Tcw_cur = Tcw_cur; P.Tcw_last = Tcw_last;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = bf; P.mb = bf / K[0];      // mb = mbf / fx, src/Frame.cc:121
    P.min_x = cur->min_x; P.max_x = cur->max_x; P.min_y = cur->min_y; P.max_y = cur->max_y; P.th = th;
    P.mono = bMono ? 1 : 0; P.nlevels = nlevels;
    for (int i = 0; i < nlevels; ++i) P.scale[i] = scale[i];
    P.n_last = n_last; P.stride = last_stride;
    P.kps_last = kps_last; P.last_mp = last_mp; P.last_outlier = last_outlier; P.last_xyz = last_xyz;
    P.uvr = (float*)ws; P.min_level = (int*)(P.uvr + 3 * nq); P.max_level = P.min_level + nq;
    P.ur = (float*)(P.max_level + nq); P.er_max = P.ur + nq; P.q_angle = P.er_max + nq; P.valid = (uint8_t*)(P.q_angle + nq);
    k_project_last_frame<<<dim3((last_stride + 255) / 256, nprob), 256, 0, st>>>(P);
    int rc = cudaGetLastError() == cudaSuccess ? ORBX_OK : ORBX_E_CUDA;
    if (rc == ORBX_OK) {
        orbm_windows W;
        W.nq = n_last; W.nq_stride = last_stride; W.uvr = P.uvr; W.min_level = P.min_level; W.max_level = P.max_level;
        W.ur = cur->u_right ? P.ur : nullptr; W.er_max = cur->u_right ? P.er_max : nullptr;
        W.valid = P.valid; W.qdesc = last_mp_desc; W.q_angle = P.q_angle; W.q_obs = last_mp_obs;
        rc = orbm_window_search_best_batch(cur, &W, cur_init_obs, assign_out, 100 /* TH_HIGH, :256 */, checkOri, nmatches, nullptr, cuda_stream);
    }
    cudaFreeAsync(ws, st);
    return rc;
}

int orbm_distinctive_descriptors(const uint8_t* desc, const int* offsets, int npoints, const uint8_t* bad,
                                 int* best_idx, int* best_median, void* cuda_stream)
{
    if (!desc || !offsets || npoints <= 0 || !best_idx || ((uintptr_t)desc & 15)) return ORBX_E_ARG;
    const int dev = dev_of(desc);
    if (dev < 0 || dev_of(offsets) != dev || dev_of(best_idx) != dev) return ORBX_E_ARG;
    DevScope g;
    if (!g.enter(dev)) { cudaGetLastError(); return ORBX_E_CUDA; }
    k_distinctive<<<npoints, DD_NT, 0, (cudaStream_t)cuda_stream>>>((const uint4*)desc, offsets, bad, best_idx, best_median);
    if (cudaGetLastError() != cudaSuccess) return ORBX_E_CUDA;
    return ORBX_OK;
}

} // extern "C"
