// oracle/ref_fuse_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points around the reference's UNMODIFIED
//   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, const float th)                              src/ORBmatcher.cc:1364-1513
//   ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, float th, vector<MapPoint*>&)   src/ORBmatcher.cc:1516-1633
//   ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vector<MapPoint*>&, s12, R12, t12, th)              src/ORBmatcher.cc:836-1052
// Key frames are the mock of oracle/mock/mock_slam.hpp with the grid copied from a real ORB_SLAM2::Frame that the reference's
// own Frame::AssignFeaturesToGrid filled (what the KeyFrame constructor does, src/KeyFrame.cc:49-55).  Map points are rows of
// plain arrays; a point is named by its row, -1 = NULL.
#define private public
#define protected public
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

#include "ref_arena.hpp"

using namespace ORB_SLAM2;

namespace {
struct RefKp { float x, y, size, angle, response; int octave, class_id; };

void fill_kf(KeyFrame& KF, int n, const RefKp* kps, const unsigned char* desc, const float* u_right,
             float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2, int nlevels, const float* K, float bf)
{
    Frame F;
    F.N = n;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; ++i) F.mvKeys[i] = cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
    F.mvKeysUn = F.mvKeys;
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    F.AssignFeaturesToGrid();
    KF.N = n;
    KF.mvKeysUn = F.mvKeysUn;
    KF.mvuRight.assign(n, -1.0f);
    if (u_right) KF.mvuRight.assign(u_right, u_right + n);
    KF.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(KF.mDescriptors.data, desc, (size_t)n * 32);
    KF.mvScaleFactors.assign(scale, scale + nlevels);
    if (inv_sigma2) KF.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
    KF.fx = K[0]; KF.fy = K[1]; KF.cx = K[2]; KF.cy = K[3]; KF.mbf = bf;
    KF.mnMinX = (int)Frame::mnMinX; KF.mnMaxX = (int)Frame::mnMaxX; KF.mnMinY = (int)Frame::mnMinY; KF.mnMaxY = (int)Frame::mnMaxY;   // include/KeyFrame.h: ints
    KF.mfGridElementWidthInv = Frame::mfGridElementWidthInv; KF.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    KF.mGrid.resize(FRAME_GRID_COLS);
    for (int i = 0; i < FRAME_GRID_COLS; ++i) {
        KF.mGrid[i].resize(FRAME_GRID_ROWS);
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) KF.mGrid[i][j] = F.mGrid[i][j];
    }
    KF.mapPoints.assign(n, static_cast<MapPoint*>(NULL));
}

cv::Mat mat_of(const float* v, int rows, int cols)
{
    cv::Mat m(rows, cols, CV_32F);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(r, c) = v[r * cols + c];
    return m;
}

void fill_points(std::vector<MapPoint>& mps, int npts, const unsigned char* bad, const float* xyz, const float* normal, const unsigned char* mp_desc,
                 const int* pred_level, const float* min_dist, const float* max_dist, const int* nobs)
{
    for (int i = 0; i < npts; ++i) {
        MapPoint& m = mps[i];
        m.mnId = (unsigned long)i;
        m.bad = bad[i] != 0;
        m.worldPos.create(3, 1, CV_32F); m.normal.create(3, 1, CV_32F);
        for (int r = 0; r < 3; ++r) { m.worldPos.at<float>(r) = xyz[3 * i + r]; m.normal.at<float>(r) = normal ? normal[3 * i + r] : 0.f; }
        m.descriptor.create(1, 32, CV_8U);
        std::memcpy(m.descriptor.data, mp_desc + (size_t)i * 32, 32);
        m.mnTrackScaleLevel = pred_level[i];         // what the mock's PredictScale returns
        m.minDist = min_dist[i]; m.maxDist = max_dist[i];
        m.nObs = nobs ? nobs[i] : 0;
    }
}
}

// State arrays in and out: bad, nobs, kf_idx (GetIndexInKeyFrame(pKF)) [npts]; kf_mp [n] (pKF->GetMapPoint); replaced_by [npts]
// receives Replace's argument (-1 = never replaced).  sim3 = 0: Fuse(pKF, vpMapPoints, th) with the key frame's pose
// (Rcw, tcw, Ow); sim3 = 1: Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) with pose16 = Scw (row-major 4x4), replace_out [nlist].
extern "C" int orbref_fuse(int sim3, int n, const RefKp* kps, const unsigned char* desc, const float* u_right,
                           float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2, int nlevels,
                           const float* K, float bf, const float* Rcw, const float* tcw, const float* Ow, const float* Scw,
                           int npts, unsigned char* bad, const float* xyz, const float* normal, const unsigned char* mp_desc,
                           const int* pred_level, const float* min_dist, const float* max_dist, int* nobs, int* kf_idx, int* replaced_by,
                           int nlist, const int* list, int* kf_mp, int* replace_out, float th)
{
    ref_arena::Scope scope;
    int nf;
    {
        KeyFrame KF;
        fill_kf(KF, n, kps, desc, u_right, minX, maxX, minY, maxY, scale, inv_sigma2, nlevels, K, bf);
        if (!sim3) { KF.R = mat_of(Rcw, 3, 3); KF.t = mat_of(tcw, 3, 1); KF.Ow = mat_of(Ow, 3, 1); }
        std::vector<MapPoint> mps(npts > 0 ? npts : 1);
        fill_points(mps, npts, bad, xyz, normal, mp_desc, pred_level, min_dist, max_dist, nobs);
        for (int i = 0; i < npts; ++i) if (kf_idx[i] >= 0) mps[i].obs[&KF] = (size_t)kf_idx[i];
        for (int k = 0; k < n; ++k) if (kf_mp[k] >= 0) KF.mapPoints[k] = &mps[kf_mp[k]];
        std::vector<MapPoint*> pts(nlist);
        for (int i = 0; i < nlist; ++i) pts[i] = list[i] >= 0 ? &mps[list[i]] : static_cast<MapPoint*>(NULL);
        ORBmatcher matcher(0.6f, true);
        if (sim3) {
            std::vector<MapPoint*> rep(nlist, static_cast<MapPoint*>(NULL));
            nf = matcher.Fuse(&KF, mat_of(Scw, 4, 4), pts, th, rep);
            for (int i = 0; i < nlist; ++i) replace_out[i] = rep[i] ? (int)(rep[i] - &mps[0]) : -1;
        } else {
            nf = matcher.Fuse(&KF, pts, th);
        }
        for (int i = 0; i < npts; ++i) {
            bad[i] = mps[i].bad ? 1 : 0;
            nobs[i] = mps[i].nObs;
            kf_idx[i] = mps[i].GetIndexInKeyFrame(&KF);
            replaced_by[i] = mps[i].replacedBy ? (int)(mps[i].replacedBy - &mps[0]) : -1;
        }
        for (int k = 0; k < n; ++k) kf_mp[k] = KF.mapPoints[k] ? (int)(KF.mapPoints[k] - &mps[0]) : -1;
    }
    return nf;
}

// mp1 [n1] / mp2 [n2]: GetMapPointMatches() of the two key frames; matches12 [n1]: vpMatches12 in and out; idx_in_kf2 [npts]:
// GetIndexInKeyFrame(pKF2).
extern "C" int orbref_search_by_sim3(int n1, const RefKp* kps1, const unsigned char* desc1, const int* mp1,
                                     int n2, const RefKp* kps2, const unsigned char* desc2, const int* mp2,
                                     float minX, float maxX, float minY, float maxY, const float* scale, int nlevels, const float* K,
                                     const float* R1w, const float* t1w, const float* R2w, const float* t2w, float s12, const float* R12, const float* t12,
                                     int npts, const unsigned char* bad, const float* xyz, const unsigned char* mp_desc, const int* pred_level,
                                     const float* min_dist, const float* max_dist, const int* idx_in_kf2, int* matches12, float th)
{
    ref_arena::Scope scope;
    int nf;
    {
        KeyFrame KF1, KF2;
        fill_kf(KF1, n1, kps1, desc1, NULL, minX, maxX, minY, maxY, scale, NULL, nlevels, K, 0.f);
        fill_kf(KF2, n2, kps2, desc2, NULL, minX, maxX, minY, maxY, scale, NULL, nlevels, K, 0.f);
        KF1.R = mat_of(R1w, 3, 3); KF1.t = mat_of(t1w, 3, 1);
        KF2.R = mat_of(R2w, 3, 3); KF2.t = mat_of(t2w, 3, 1);
        std::vector<MapPoint> mps(npts > 0 ? npts : 1);
        fill_points(mps, npts, bad, xyz, NULL, mp_desc, pred_level, min_dist, max_dist, NULL);
        for (int i = 0; i < npts; ++i) if (idx_in_kf2[i] >= 0) mps[i].obs[&KF2] = (size_t)idx_in_kf2[i];
        for (int k = 0; k < n1; ++k) if (mp1[k] >= 0) KF1.mapPoints[k] = &mps[mp1[k]];
        for (int k = 0; k < n2; ++k) if (mp2[k] >= 0) KF2.mapPoints[k] = &mps[mp2[k]];
        std::vector<MapPoint*> m12(n1, static_cast<MapPoint*>(NULL));
        for (int k = 0; k < n1; ++k) if (matches12[k] >= 0) m12[k] = &mps[matches12[k]];
        ORBmatcher matcher(0.75f, true);
        nf = matcher.SearchBySim3(&KF1, &KF2, m12, s12, mat_of(R12, 3, 3), mat_of(t12, 3, 1), th);
        for (int k = 0; k < n1; ++k) matches12[k] = m12[k] ? (int)(m12[k] - &mps[0]) : -1;
    }
    return nf;
}
