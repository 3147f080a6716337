// oracle/ref_sim3_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry point around the reference's UNMODIFIED loop-closing overload
//   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, vector<MapPoint*>&, int th)
//   src/ORBmatcher.cc:434-549
// The key frame is the mock of oracle/mock/mock_slam.hpp with its grid copied from a real ORB_SLAM2::Frame that the
// reference's own Frame::AssignFeaturesToGrid filled (what the KeyFrame constructor does, src/KeyFrame.cc:49-55).
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
}

// matched_in [n]: -1 free, k >= 0: candidate point k is already matched to this keypoint, -2: some other point is.
// assign_out [n]: index of the candidate point in vpMatched afterwards, -2 for the foreign point, -1 for none.
extern "C" int orbref_search_by_projection_sim3(int n, const RefKp* kps, const unsigned char* desc,
                                                float minX, float maxX, float minY, float maxY, const float* scale, int nlevels,
                                                const float* K, const float* Scw, int npts, const unsigned char* bad, const float* xyz,
                                                const float* normal, const unsigned char* mp_desc, const int* pred_level,
                                                const float* min_dist, const float* max_dist, const int* matched_in, int* assign_out, int th)
{
    ref_arena::Scope scope;
    int nm;
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
        KeyFrame KF;
        KF.N = n;
        KF.mvKeysUn = F.mvKeysUn;
        KF.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
        if (n > 0) std::memcpy(KF.mDescriptors.data, desc, (size_t)n * 32);
        KF.mvScaleFactors.assign(scale, scale + nlevels);
        KF.fx = K[0]; KF.fy = K[1]; KF.cx = K[2]; KF.cy = K[3];
        KF.mnMinX = (int)Frame::mnMinX; KF.mnMaxX = (int)Frame::mnMaxX; KF.mnMinY = (int)Frame::mnMinY; KF.mnMaxY = (int)Frame::mnMaxY;   // include/KeyFrame.h: ints
        KF.mfGridElementWidthInv = Frame::mfGridElementWidthInv; KF.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
        KF.mGrid.resize(FRAME_GRID_COLS);
        for (int i = 0; i < FRAME_GRID_COLS; ++i) {
            KF.mGrid[i].resize(FRAME_GRID_ROWS);
            for (int j = 0; j < FRAME_GRID_ROWS; ++j) KF.mGrid[i][j] = F.mGrid[i][j];
        }
        cv::Mat S(4, 4, CV_32F);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) S.at<float>(r, c) = Scw[4 * r + c];
        std::vector<MapPoint> mps(npts > 0 ? npts : 1);
        std::vector<MapPoint*> pts(npts);
        for (int i = 0; i < npts; ++i) {
            MapPoint& m = mps[i];
            m.bad = bad[i] != 0;
            m.worldPos.create(3, 1, CV_32F); m.normal.create(3, 1, CV_32F);
            for (int r = 0; r < 3; ++r) { m.worldPos.at<float>(r) = xyz[3 * i + r]; m.normal.at<float>(r) = normal[3 * i + r]; }
            m.descriptor.create(1, 32, CV_8U);
            std::memcpy(m.descriptor.data, mp_desc + (size_t)i * 32, 32);
            m.mnTrackScaleLevel = pred_level[i];         // what the mock's PredictScale returns
            m.minDist = min_dist[i]; m.maxDist = max_dist[i];
            pts[i] = &m;
        }
        MapPoint foreign;
        std::vector<MapPoint*> matched(n, static_cast<MapPoint*>(NULL));
        for (int k = 0; k < n; ++k) matched[k] = matched_in[k] >= 0 ? &mps[matched_in[k]] : matched_in[k] == -2 ? &foreign : static_cast<MapPoint*>(NULL);
        ORBmatcher matcher(0.75f, true);
        nm = matcher.SearchByProjection(&KF, S, pts, matched, th);
        for (int k = 0; k < n; ++k) {
            MapPoint* p = matched[k];
            assign_out[k] = !p ? -1 : p == &foreign ? -2 : (int)(p - &mps[0]);
        }
    }
    return nm;
}
