// oracle/ref_frustum_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry point around the reference's UNMODIFIED Frame::isInFrustum (src/Frame.cc:288-345) and
// Frame::SetPose / UpdatePoseMatrices (:271-285): the projection step of Tracking::SearchLocalPoints that feeds
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th).  MapPoint is the mock of mock/mock_slam.hpp.
#define private public
#define protected public
#include "Frame.h"
#undef private
#undef protected

#include <cmath>
#include <cstring>
#include <vector>

using namespace ORB_SLAM2;

// The mock's PredictScale(dist, Frame*): the statements of src/MapPoint.cc:459-475 on the raw mfMaxDistance.
int MapPoint::PredictScaleReal(const float& currentDist, Frame* pF)
{
    float ratio = maxDistRaw / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

extern "C" {

// Tcw row-major 4x4; K = fx, fy, cx, cy; bounds = mnMinX, mnMaxX, mnMinY, mnMaxY; per point: world position,
// normal (mNormalVector), mfMaxDistance, mfMinDistance.  Outputs per point: mbTrackInView, (mTrackProjX,
// mTrackProjY, mTrackProjXR), mnTrackScaleLevel, mTrackViewCos; entries of points not in view are left untouched.
int orbref_is_in_frustum(const float* Tcw, const float* K, float bf, float minX, float maxX, float minY, float maxY,
                         float scale_factor, int nlevels, float viewing_cos_limit, int n, const float* xyz, const float* normal,
                         const float* max_distance, const float* min_distance, unsigned char* in_view, float* proj_xyxr,
                         int* level, float* view_cos)
{
    Frame F;
    F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3]; F.mbf = bf;
    F.mnScaleLevels = nlevels;
    F.mfScaleFactor = scale_factor;
    F.mfLogScaleFactor = log(F.mfScaleFactor);                         // src/Frame.cc:75
    Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
    cv::Mat T(4, 4, CV_32F);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.at<float>(i, j) = Tcw[4 * i + j];
    F.SetPose(T);                                                      // :271-285
    int count = 0;
    for (int i = 0; i < n; ++i) {
        MapPoint mp;
        mp.worldPos = cv::Mat(3, 1, CV_32F);
        mp.normal = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) { mp.worldPos.at<float>(k) = xyz[3 * i + k]; mp.normal.at<float>(k) = normal[3 * i + k]; }
        mp.maxDistRaw = max_distance[i];
        mp.maxDist = 1.2f * max_distance[i];                           // GetMaxDistanceInvariance, src/MapPoint.cc:431-435
        mp.minDist = 0.8f * min_distance[i];                           // GetMinDistanceInvariance, :424-428
        mp.realPredict = true;
        const bool ok = F.isInFrustum(&mp, viewing_cos_limit);
        in_view[i] = ok ? 1 : 0;
        if (ok) {
            proj_xyxr[3 * i] = mp.mTrackProjX; proj_xyxr[3 * i + 1] = mp.mTrackProjY; proj_xyxr[3 * i + 2] = mp.mTrackProjXR;
            level[i] = mp.mnTrackScaleLevel; view_cos[i] = mp.mTrackViewCos;
            ++count;
        }
    }
    return count;
}

} // extern "C"
