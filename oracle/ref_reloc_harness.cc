// oracle/ref_reloc_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry point around the reference's UNMODIFIED relocalisation overload
//   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)  src/ORBmatcher.cc:303-431
// on a real ORB_SLAM2::Frame (filled from arrays, gridded by the reference's AssignFeaturesToGrid) and the
// mock KeyFrame / MapPoint of oracle/mock/mock_slam.hpp.
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

extern "C" int orbref_search_by_projection_reloc(int n_cur, const RefKp* kps_cur, const unsigned char* desc_cur,
                                                 float minX, float maxX, float minY, float maxY, const float* scale, int nlevels,
                                                 int nkf, const unsigned char* has_mp, const unsigned char* bad, const unsigned char* already_found,
                                                 const float* xyz, const unsigned char* mp_desc, const int* pred_level,
                                                 const float* min_dist, const float* max_dist, const float* kf_angle,
                                                 const float* Tcw, const float* K, const unsigned char* cur_taken, int* assign_out,
                                                 float th, int ORBdist, int check_ori)
{
    ref_arena::Scope scope;
    int nm;
    {
        Frame F;
        F.N = n_cur;
        F.mvKeys.resize(n_cur);
        for (int i = 0; i < n_cur; ++i) F.mvKeys[i] = cv::KeyPoint(kps_cur[i].x, kps_cur[i].y, kps_cur[i].size, kps_cur[i].angle, kps_cur[i].response, kps_cur[i].octave, kps_cur[i].class_id);
        F.mvKeysUn = F.mvKeys;
        F.mDescriptors.create(n_cur > 0 ? n_cur : 1, 32, CV_8U);
        if (n_cur > 0) std::memcpy(F.mDescriptors.data, desc_cur, (size_t)n_cur * 32);
        F.mvuRight.assign(n_cur, -1.0f);
        F.mvpMapPoints.assign(n_cur, static_cast<MapPoint*>(NULL));
        F.mvScaleFactors.assign(scale, scale + nlevels);
        Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
        Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
        F.AssignFeaturesToGrid();
        F.mTcw.create(4, 4, CV_32F);
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) F.mTcw.at<float>(r, c) = Tcw[4 * r + c];
        MapPoint taken;
        if (cur_taken) for (int k = 0; k < n_cur; ++k) if (cur_taken[k]) F.mvpMapPoints[k] = &taken;
        KeyFrame KF;
        KF.N = nkf;
        KF.mvKeysUn.resize(nkf);
        std::vector<MapPoint> mps(nkf);
        KF.mapPoints.assign(nkf, static_cast<MapPoint*>(NULL));
        std::set<MapPoint*> found;
        for (int i = 0; i < nkf; ++i) {
            KF.mvKeysUn[i].angle = kf_angle[i];
            if (!has_mp[i]) continue;
            MapPoint& m = mps[i];
            m.bad = bad[i] != 0;
            m.worldPos.create(3, 1, CV_32F);
            for (int r = 0; r < 3; ++r) m.worldPos.at<float>(r) = xyz[3 * i + r];
            m.descriptor.create(1, 32, CV_8U);
            std::memcpy(m.descriptor.data, mp_desc + (size_t)i * 32, 32);
            m.mnTrackScaleLevel = pred_level[i];         // what the mock's PredictScale returns
            m.minDist = min_dist[i]; m.maxDist = max_dist[i];
            KF.mapPoints[i] = &m;
            if (already_found[i]) found.insert(&m);
        }
        ORBmatcher matcher(0.9f, check_ori != 0);
        nm = matcher.SearchByProjection(F, &KF, found, th, ORBdist);
        for (int k = 0; k < n_cur; ++k) {
            MapPoint* p = F.mvpMapPoints[k];
            if (!p) assign_out[k] = -1;
            else if (p == &taken) assign_out[k] = -2;
            else assign_out[k] = (int)(p - &mps[0]);
        }
    }
    return nm;
}
