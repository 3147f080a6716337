// oracle/ref_matcher_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the reference's UNMODIFIED src/ORBmatcher.cc and src/Frame.cc (compiled
// from /root/reference against oracle/cvshim with the mocks of oracle/mock/mock_slam.hpp).
// Frames are the reference's real ORB_SLAM2::Frame objects: either built by its own stereo
// constructor (extraction x2 + ComputeStereoMatches + AssignFeaturesToGrid, src/Frame.cc:62-124)
// or filled from caller arrays and gridded with the reference's AssignFeaturesToGrid.
#define private public      // this TU only: reach Frame::AssignFeaturesToGrid / ORBmatcher internals
#define protected public
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

#include <chrono>
#include <cstddef>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "ref_arena.hpp"

using namespace ORB_SLAM2;

namespace {
struct RefKp { float x, y, size, angle, response; int octave, class_id; };

// A Frame filled from arrays.  Statics (image bounds, grid scale) are set like
// Frame::ComputeImageBounds / the first-frame block of the constructors do (src/Frame.cc:100-118).
void fill_frame(Frame& F, int n, const RefKp* kps, const unsigned char* desc, const float* uRight,
                const float* scale, int nlevels, float minX, float maxX, float minY, float maxY)
{
    F.N = n;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; ++i) F.mvKeys[i] = cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
    F.mvKeysUn = F.mvKeys;
    F.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(F.mDescriptors.data, desc, (size_t)n * 32);
    F.mvuRight.assign(n, -1.0f);
    F.mvDepth.assign(n, -1.0f);
    if (uRight) for (int i = 0; i < n; ++i) F.mvuRight[i] = uRight[i];
    F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(n, false);
    F.mnScaleLevels = nlevels;
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvScaleFactors.resize(nlevels); F.mvLevelSigma2.resize(nlevels); F.mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {
        F.mvInvScaleFactors[i] = 1.0f / scale[i];
        F.mvLevelSigma2[i] = scale[i] * scale[i];
        F.mvInvLevelSigma2[i] = 1.0f / F.mvLevelSigma2[i];
    }
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);    // src/Frame.cc:108
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);   // :109
    Frame::mbInitialComputations = false;
    F.AssignFeaturesToGrid();                                                                                // :243-259
}

cv::Mat mat_from(const float* p, int r, int c)
{
    cv::Mat m(r, c, CV_32F);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = p[i * c + j];
    return m;
}
} // namespace

// A stereo Frame built in storage whose `mb` member already holds mbf/fx: the constructor reads mb (through
// ComputeStereoMatches, src/Frame.cc:93 -> :523) before it assigns it (:121); in the running system the storage is
// the previous Frame's (Tracking builds every Frame at the same address), which is the steady state modelled here.
struct SteadyStereoFrame {
    alignas(Frame) unsigned char buf[sizeof(Frame)];
    Frame* F;
    SteadyStereoFrame(const cv::Mat& imL, const cv::Mat& imR, ORBextractor* exl, ORBextractor* exr, ORBVocabulary* voc,
                      cv::Mat& K, cv::Mat& dist, float bf, float thDepth)
    {
        std::memset(buf, 0, sizeof(buf));
        const float mb = bf / K.at<float>(0, 0);
        std::memcpy(buf + offsetof(Frame, mb), &mb, sizeof(mb));
        F = new (buf) Frame(imL, imR, 0.0, exl, exr, voc, K, dist, bf, thDepth);
    }
    ~SteadyStereoFrame() { F->~Frame(); }
    SteadyStereoFrame(const SteadyStereoFrame&) = delete;
};

extern "C" {

// ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:46-63
int orbref_descriptor_distance(const unsigned char* a, const unsigned char* b)
{
    cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(A, B);
}

// Frame::GetFeaturesInArea on a frame built from arrays (src/Frame.cc:348-409).  Returns the count.
int orbref_features_in_area(int n, const RefKp* kps, const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                            float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    ref_arena::Scope scope;
    int cnt;
    {
        std::vector<unsigned char> desc((size_t)(n > 0 ? n : 1) * 32, 0);
        Frame F;
        fill_frame(F, n, kps, desc.data(), NULL, scale, nlevels, minX, maxX, minY, maxY);
        std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
        cnt = (int)v.size();
        for (int i = 0; i < cnt && i < cap; ++i) out[i] = (int)v[i];
    }
    return cnt;
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:1055-1180).
// prev_matched [n1][2] is updated in place like vbPrevMatched; matches12 [n1] receives vnMatches12.
int orbref_search_for_initialization(int n1, const RefKp* kps1, const unsigned char* desc1,
                                     int n2, const RefKp* kps2, const unsigned char* desc2,
                                     const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                                     float* prev_matched, int* matches12, int windowSize, float nnratio, int checkOri)
{
    ref_arena::Scope scope;
    int nm;
    {
        Frame F1, F2;
        fill_frame(F1, n1, kps1, desc1, NULL, scale, nlevels, minX, maxX, minY, maxY);
        fill_frame(F2, n2, kps2, desc2, NULL, scale, nlevels, minX, maxX, minY, maxY);
        std::vector<cv::Point2f> prev(n1);
        for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
        std::vector<int> m12;
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
        for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    }
    return nm;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:73-157).
// Map points come as arrays [nq]: proj (x, y, xr), level, viewCos, track_in_view, bad, observations, descriptor.
// u_right [n] may be NULL.  init_assign [n]: index of a map point already attached to a keypoint or -1.
// assign_out [n] receives the index of the map point attached to each keypoint afterwards.
int orbref_search_by_projection_points(int n, const RefKp* kps, const unsigned char* desc, const float* u_right,
                                       const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                                       int nq, const float* proj_xyxr, const int* level, const float* view_cos,
                                       const unsigned char* in_view, const unsigned char* bad, const int* observations,
                                       const unsigned char* qdesc, const int* init_assign, int* assign_out,
                                       float th, float nnratio)
{
    ref_arena::Scope scope;
    int nm;
    {
        Frame F;
        fill_frame(F, n, kps, desc, u_right, scale, nlevels, minX, maxX, minY, maxY);
        std::vector<MapPoint> mps(nq);
        std::vector<MapPoint*> ptrs(nq);
        for (int i = 0; i < nq; ++i) {
            MapPoint& m = mps[i];
            m.mTrackProjX = proj_xyxr[3 * i]; m.mTrackProjY = proj_xyxr[3 * i + 1]; m.mTrackProjXR = proj_xyxr[3 * i + 2];
            m.mnTrackScaleLevel = level[i]; m.mTrackViewCos = view_cos[i];
            m.mbTrackInView = in_view[i] != 0; m.bad = bad[i] != 0; m.nObs = observations[i];
            m.descriptor.create(1, 32, CV_8U);
            std::memcpy(m.descriptor.data, qdesc + (size_t)i * 32, 32);
            ptrs[i] = &m;
        }
        if (init_assign) for (int k = 0; k < n; ++k) if (init_assign[k] >= 0) F.mvpMapPoints[k] = ptrs[init_assign[k]];
        ORBmatcher matcher(nnratio, true);
        nm = matcher.SearchByProjection(F, ptrs, th);
        for (int k = 0; k < n; ++k) assign_out[k] = F.mvpMapPoints[k] ? (int)(F.mvpMapPoints[k] - &mps[0]) : -1;
    }
    return nm;
}

// ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono) (src/ORBmatcher.cc:160-300).
// last_mp [n_last]: 1 if the last-frame keypoint has a map point (world position in last_xyz, its
// descriptor = last_mp_desc); last_outlier [n_last]; poses are row-major 4x4 Tcw.  K = fx,fy,cx,cy.
// cur_init_obs [n_cur]: -1 = keypoint free, otherwise Observations() of a map point already attached.
// assign_out [n_cur]: index i of the last-frame keypoint whose map point got attached, -2 for a
// pre-attached point that was kept, -1 for none.
int orbref_search_by_projection_frame(int n_cur, const RefKp* kps_cur, const unsigned char* desc_cur, const float* u_right_cur,
                                      int n_last, const RefKp* kps_last, const unsigned char* last_mp, const unsigned char* last_outlier,
                                      const float* last_xyz, const unsigned char* last_mp_desc, const int* last_mp_obs,
                                      const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                      const float* scale, int nlevels, float minX, float maxX, float minY, float maxY,
                                      const int* cur_init_obs, int* assign_out, float th, int bMono, float nnratio, int checkOri)
{
    ref_arena::Scope scope;
    int nm;
    {
        std::vector<unsigned char> zero((size_t)(n_last > 0 ? n_last : 1) * 32, 0);
        Frame cur, last;
        fill_frame(last, n_last, kps_last, zero.data(), NULL, scale, nlevels, minX, maxX, minY, maxY);
        fill_frame(cur, n_cur, kps_cur, desc_cur, u_right_cur, scale, nlevels, minX, maxX, minY, maxY);
        Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
        Frame::invfx = 1.0f / K[0]; Frame::invfy = 1.0f / K[1];
        cur.mbf = last.mbf = bf; cur.mb = last.mb = bf / K[0];                      // src/Frame.cc:121
        cur.mTcw = mat_from(Tcw_cur, 4, 4); last.mTcw = mat_from(Tcw_last, 4, 4);
        std::vector<MapPoint> mps(n_last), pre(n_cur);
        for (int i = 0; i < n_last; ++i) {
            if (!last_mp[i]) continue;
            MapPoint& m = mps[i];
            m.worldPos = mat_from(last_xyz + 3 * i, 3, 1);
            m.descriptor.create(1, 32, CV_8U);
            std::memcpy(m.descriptor.data, last_mp_desc + (size_t)i * 32, 32);
            m.nObs = last_mp_obs ? last_mp_obs[i] : 1;
            last.mvpMapPoints[i] = &m;
            last.mvbOutlier[i] = last_outlier && last_outlier[i];
        }
        if (cur_init_obs) for (int k = 0; k < n_cur; ++k) if (cur_init_obs[k] >= 0) { pre[k].nObs = cur_init_obs[k]; cur.mvpMapPoints[k] = &pre[k]; }
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchByProjection(cur, last, th, bMono != 0);
        for (int k = 0; k < n_cur; ++k) {
            MapPoint* p = cur.mvpMapPoints[k];
            if (!p) assign_out[k] = -1;
            else if (p >= &mps[0] && p < &mps[0] + n_last) assign_out[k] = (int)(p - &mps[0]);
            else assign_out[k] = -2;
        }
    }
    return nm;
}

// The reference's stereo Frame constructor (src/Frame.cc:62-124): two extractors on two threads,
// ComputeStereoMatches (:513-699), AssignFeaturesToGrid.  Returns N (left keypoints).
// Outputs (each may be NULL): left/right keypoints + descriptors, mvuRight, mvDepth.
int orbref_stereo_frame(const unsigned char* left, const unsigned char* right, int w, int h,
                        int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                        float fx, float fy, float cx, float cy, float bf, float thDepth,
                        RefKp* kps_l, unsigned char* desc_l, int cap_l, RefKp* kps_r, unsigned char* desc_r, int cap_r, int* n_right,
                        float* u_right, float* depth)
{
    ref_arena::Scope scope;
    ref_arena::reset_parked();
    int n;
    {
        ORBextractor exl(nfeatures, scaleFactor, nlevels, iniTh, minTh), exr(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)right, (size_t)w);
        float Kd[9] = { fx, 0, cx, 0, fy, cy, 0, 0, 1 };
        cv::Mat K = mat_from(Kd, 3, 3);
        float Dd[4] = { 0, 0, 0, 0 };
        cv::Mat dist = mat_from(Dd, 4, 1);
        Frame::mbInitialComputations = true;     // recompute the image bounds / grid scale for this shape
        ORBVocabulary voc;
        SteadyStereoFrame SF(imL, imR, &exl, &exr, &voc, K, dist, bf, thDepth);
        Frame& F = *SF.F;
        n = F.N;
        for (int i = 0; i < n && i < cap_l; ++i) {
            const cv::KeyPoint& k = F.mvKeys[i];
            if (kps_l) { RefKp r = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id }; kps_l[i] = r; }
            if (desc_l) std::memcpy(desc_l + (size_t)i * 32, F.mDescriptors.ptr(i), 32);
            if (u_right) u_right[i] = F.mvuRight[i];
            if (depth) depth[i] = F.mvDepth[i];
        }
        const int nr = (int)F.mvKeysRight.size();
        if (n_right) *n_right = nr;
        for (int i = 0; i < nr && i < cap_r; ++i) {
            const cv::KeyPoint& k = F.mvKeysRight[i];
            if (kps_r) { RefKp r = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id }; kps_r[i] = r; }
            if (desc_r) std::memcpy(desc_r + (size_t)i * 32, F.mDescriptorsRight.ptr(i), 32);
        }
    }
    return n;
}

// CPU baseline for BASELINE.json configs[1]: npairs rectified pairs (L0,R0,L1,R1,..., pitch == w) through the
// reference's stereo Frame constructor, round-robin over nworkers threads; every Frame runs its two extractors on
// two threads of its own (src/Frame.cc:82-85), so 2*nworkers threads are busy.  Returns the wall time in seconds;
// total_depth receives the number of keypoints that got a depth.
// Reference quirk: the constructor calls ComputeStereoMatches() (:93) BEFORE it sets mb = mbf/fx (:121), so minZ = mb
// (:523) is whatever the member's storage held -- the previous Frame's mb when Frames are built at the same address
// (what Tracking does), zero/garbage for the very first one (maxD = inf then).  The oracle and the CUDA path
// implement the steady state, mb = mbf/fx; SteadyStereoFrame above pins the reference to it.
double orbref_stereo_bench(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
                           const unsigned char* frames, int npairs, int w, int h, int nworkers,
                           float fx, float fy, float cx, float cy, float bf, float thDepth, long long* total_depth,
                           float* depth_out /* [npairs][cap] or NULL */, int cap)
{
    if (nworkers < 1) nworkers = 1;
    std::vector<long long> counts((size_t)nworkers, 0);
    std::vector<std::thread> workers;
    ref_arena::reset_parked();
    Frame::mbInitialComputations = true;
    {   // the static image bounds / grid scale are computed by the first Frame: do that before the threads start
        ORBextractor exl(nfeatures, scaleFactor, nlevels, iniTh, minTh), exr(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        ref_arena::Scope scope;
        float Kd[9] = { fx, 0, cx, 0, fy, cy, 0, 0, 1 };
        float Dd[4] = { 0, 0, 0, 0 };
        cv::Mat imL(h, w, CV_8UC1, (void*)frames, (size_t)w), imR(h, w, CV_8UC1, (void*)(frames + (size_t)w * h), (size_t)w);
        cv::Mat K = mat_from(Kd, 3, 3), dist = mat_from(Dd, 4, 1);
        ORBVocabulary voc;
        SteadyStereoFrame SF(imL, imR, &exl, &exr, &voc, K, dist, bf, thDepth);
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < nworkers; ++t)
        workers.emplace_back([&, t]() {
            ORBextractor exl(nfeatures, scaleFactor, nlevels, iniTh, minTh), exr(nfeatures, scaleFactor, nlevels, iniTh, minTh);
            float Kd[9] = { fx, 0, cx, 0, fy, cy, 0, 0, 1 };
            float Dd[4] = { 0, 0, 0, 0 };
            for (int p = t; p < npairs; p += nworkers) {
                ref_arena::Scope scope;
                const unsigned char* left = frames + (size_t)(2 * p) * w * h;
                cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)(left + (size_t)w * h), (size_t)w);
                cv::Mat K = mat_from(Kd, 3, 3), dist = mat_from(Dd, 4, 1);
                ORBVocabulary voc;
                SteadyStereoFrame SF(imL, imR, &exl, &exr, &voc, K, dist, bf, thDepth);
                Frame& F = *SF.F;
                for (int i = 0; i < F.N; ++i) counts[(size_t)t] += F.mvDepth[i] > 0;
                if (depth_out) for (int i = 0; i < F.N && i < cap; ++i) depth_out[(size_t)p * cap + i] = F.mvDepth[i];
            }
        });
    for (auto& th : workers) th.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (total_depth) { *total_depth = 0; for (long long c : counts) *total_depth += c; }
    return dt;
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:702-727) on a frame built from arrays: kps = mvKeys (distorted), kps_un =
// mvKeysUn, depth = the h x w float image.  u_right / depth_out [n] receive mvuRight / mvDepth.
int orbref_stereo_from_rgbd(int n, const RefKp* kps, const RefKp* kps_un, const float* depth, int w, int h, float bf,
                            float* u_right, float* depth_out)
{
    ref_arena::Scope scope;
    {
        Frame F;
        F.N = n;
        F.mvKeys.resize(n); F.mvKeysUn.resize(n);
        for (int i = 0; i < n; ++i) {
            F.mvKeys[i] = cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
            F.mvKeysUn[i] = cv::KeyPoint(kps_un[i].x, kps_un[i].y, kps_un[i].size, kps_un[i].angle, kps_un[i].response, kps_un[i].octave, kps_un[i].class_id);
        }
        F.mbf = bf;
        cv::Mat D(h, w, CV_32F);
        for (int r = 0; r < h; ++r) std::memcpy(D.ptr(r), depth + (size_t)r * w, sizeof(float) * (size_t)w);
        F.ComputeStereoFromRGBD(D);
        for (int i = 0; i < n; ++i) { u_right[i] = F.mvuRight[i]; depth_out[i] = F.mvDepth[i]; }
    }
    return 0;
}

} // extern "C"
