// orb_frame.cu -- the two per-keypoint steps between ORBextractor::operator() and the matchers in the reference's monocular
// and RGB-D Frame constructors (src/Frame.cc:127-240), device resident, so that such a frame stays on the GPU like a stereo one:
//   Frame::UndistortKeyPoints       src/Frame.cc:436-468   (cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK))
//   Frame::ComputeStereoFromRGBD    src/Frame.cc:702-727
// One thread per keypoint, one launch per batch of frames.  cv::undistortPoints works in double: normalise with the camera
// matrix, five fixed-point iterations of the inverse distortion model (OpenCV's default TermCriteria(MAX_ITER, 5, 0.01) counts
// only), project with the new camera matrix, narrow to float.  The operation order below reproduces OpenCV 4.13 bit for bit
// (tests/test_frame_steps.py checks the CPU restatement against cv2 live and the kernel against the restatement); the library is
// compiled with -fmad=false, so no product is fused into an addition.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_match_common.cuh"   // ORB_CHECK

namespace {
struct UndistortParams {
    const orbx_kp* in; orbx_kp* out; const int* n; int cap, batch;
    double fx, fy, cx, cy, ifx, ify;
    double k[12];          // k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 (absent ones 0)
    int identity;          // mDistCoef[0] == 0: mvKeysUn = mvKeys (:438-442)
};

__global__ void k_undistort_keypoints(const UndistortParams P)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.cap || i >= P.n[f]) return;
    orbx_kp kp = P.in[(size_t)f * P.cap + i];
    if (!P.identity) {
        const double* k = P.k;
        double x = ((double)kp.x - P.cx) * P.ifx, y = ((double)kp.y - P.cy) * P.ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        kp.x = (float)(P.fx * x + P.cx);            // only the coordinates change (:461-466)
        kp.y = (float)(P.fy * y + P.cy);
    }
    P.out[(size_t)f * P.cap + i] = kp;
}

struct RgbdParams {
    const orbx_kp* kps; const orbx_kp* kps_un; const int* n; int cap, batch;
    const float* depth; size_t pitch, frame_stride; int w, h;
    float bf; float* u_right; float* depth_out;
};

__global__ void k_stereo_from_rgbd(const RgbdParams P)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.cap) return;
    const size_t o = (size_t)f * P.cap + i;
    float ur = -1.0f, dz = -1.0f;                                                 // :704-705
    if (i < P.n[f]) {
        const orbx_kp kp = P.kps[o];
        const int v = (int)kp.y, u = (int)kp.x;                                   // imDepth.at<float>(v, u) with float arguments (:715)
        if (u >= 0 && u < P.w && v >= 0 && v < P.h) {
            ORB_CHECK((size_t)v * P.pitch + (size_t)u * 4 + 4 <= (size_t)(P.h - 1) * P.pitch + (size_t)P.w * 4);
            const float d = *(const float*)((const char*)P.depth + (size_t)f * P.frame_stride + (size_t)v * P.pitch + (size_t)u * 4);
            if (d > 0) {
                dz = d;
                ur = __fsub_rn((P.kps_un ? P.kps_un[o] : kp).x, __fdiv_rn(P.bf, d));   // :720 (the UNDISTORTED x)
            }
        }
    }
    P.u_right[o] = ur;
    P.depth_out[o] = dz;
}

int dev_of_ptr(const void* p)
{
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
    return a.device;
}
struct Guard { int prev = -1; ~Guard() { if (prev >= 0) cudaSetDevice(prev); } };
} // namespace

extern "C" {

int orbx_undistort_keypoints_batch(const orbx_kp* d_kps, orbx_kp* d_kps_un, const int* d_n, int cap_per_frame, int batch,
                                   const float* K, const float* dist_coef, int n_dist, void* cuda_stream)
{
    if (!d_kps || !d_kps_un || !d_n || cap_per_frame <= 0 || batch <= 0 || !K || n_dist < 0 || n_dist > 12 || (n_dist > 0 && !dist_coef)) return ORBX_E_ARG;
    if (n_dist != 0 && n_dist != 4 && n_dist != 5 && n_dist != 8 && n_dist != 12) return ORBX_E_ARG;   // the sizes cv::undistortPoints accepts
    const int dev = dev_of_ptr(d_kps);
    if (dev < 0 || dev_of_ptr(d_kps_un) != dev || dev_of_ptr(d_n) != dev) return ORBX_E_ARG;
    Guard g;
    if (cudaGetDevice(&g.prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    UndistortParams P = {};
    P.in = d_kps; P.out = d_kps_un; P.n = d_n; P.cap = cap_per_frame; P.batch = batch;
    P.fx = (double)K[0]; P.fy = (double)K[1]; P.cx = (double)K[2]; P.cy = (double)K[3];
    P.ifx = 1.0 / P.fx; P.ify = 1.0 / P.fy;
    for (int i = 0; i < n_dist; ++i) P.k[i] = (double)dist_coef[i];
    P.identity = n_dist == 0 || dist_coef[0] == 0.0f;
    const dim3 grid((unsigned)((cap_per_frame + 255) / 256), (unsigned)batch);
    k_undistort_keypoints<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(P);
    if (cudaGetLastError() != cudaSuccess) return ORBX_E_CUDA;
    return ORBX_OK;
}

int orbx_stereo_from_rgbd_batch(const orbx_kp* d_kps, const orbx_kp* d_kps_un, const int* d_n, int cap_per_frame, int batch,
                                const float* d_depth, size_t depth_pitch, size_t depth_frame_stride, int w, int h, float bf,
                                float* d_u_right, float* d_depth_out, void* cuda_stream)
{
    if (!d_kps || !d_n || cap_per_frame <= 0 || batch <= 0 || !d_depth || w <= 0 || h <= 0 || depth_pitch < (size_t)w * 4 || (depth_pitch & 3) ||
        (depth_frame_stride & 3) || !d_u_right || !d_depth_out)
        return ORBX_E_ARG;
    const int dev = dev_of_ptr(d_kps);
    if (dev < 0 || dev_of_ptr(d_depth) != dev || dev_of_ptr(d_n) != dev || dev_of_ptr(d_u_right) != dev || dev_of_ptr(d_depth_out) != dev ||
        (d_kps_un && dev_of_ptr(d_kps_un) != dev))
        return ORBX_E_ARG;
    Guard g;
    if (cudaGetDevice(&g.prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return ORBX_E_CUDA; }
    RgbdParams P = { d_kps, d_kps_un, d_n, cap_per_frame, batch, d_depth, depth_pitch, depth_frame_stride, w, h, bf, d_u_right, d_depth_out };
    const dim3 grid((unsigned)((cap_per_frame + 255) / 256), (unsigned)batch);
    k_stereo_from_rgbd<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(P);
    if (cudaGetLastError() != cudaSuccess) return ORBX_E_CUDA;
    return ORBX_OK;
}

} // extern "C"
