// tests/cpp/matcher_compile_test.cc -- instantiates the ORBmatcher forwarders
// (host/ORBmatcher_b200.hpp) against Frame / MapPoint types that expose the same members as the
// reference's classes, and exposes them to pytest.  Where /root/reference exists the Makefile target
// `matcher_ref` compiles the same header against the reference's REAL include/Frame.h instead.
#ifdef WITH_REFERENCE_HEADERS
#include "Frame.h"
typedef ORB_SLAM2::Frame FrameT;
typedef ORB_SLAM2::MapPoint MapPointT;
typedef ORB_SLAM2::KeyFrame KeyFrameT;
#else
#include "cvshim.hpp"
#include <vector>
struct MapPointT {
    bool mbTrackInView, bad; int mnTrackScaleLevel, nObs; float mTrackViewCos, mTrackProjX, mTrackProjY, mTrackProjXR;
    cv::Mat descriptor, pos;
    bool isBad() { return bad; }
    float GetMinDistanceInvariance() { return 0.f; }
    float GetMaxDistanceInvariance() { return 1e9f; }
    template <class F> int PredictScale(const float&, F*) { return mnTrackScaleLevel; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return descriptor.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return pos.clone(); }
    float GetMaxDistance() { return 1e9f; }
    float GetMinDistance() { return 0.f; }
};
#include <map>
struct FeatVecT : std::map<unsigned int, std::vector<unsigned int> > {};
struct FrameT {
    int N, mnScaleLevels; float mbf, mb, mfScaleFactor;
    FeatVecT mFeatVec;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    cv::Mat mDescriptors, mTcw;
    std::vector<MapPointT*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY;
};
struct KeyFrameT {
    float fx, fy, cx, cy; int mnMinX, mnMinY, mnMaxX, mnMaxY;
    cv::Mat mDescriptors;
    FeatVecT mFeatVec;
    std::vector<float> mvScaleFactors;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPointT*> pts;
    std::vector<MapPointT*> GetMapPointMatches() { return pts; }
};
float FrameT::fx, FrameT::fy, FrameT::cx, FrameT::cy, FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY;
#endif
#include "ORBmatcher_b200.hpp"

extern "C" int matcher_forwarders_instantiate(int run)
{
    if (!run) return 0;   // instantiation is the test; running needs a GPU and populated frames
    FrameT a, b;
    std::vector<MapPointT*> pts;
    std::vector<cv::Point2f> prev;
    std::vector<int> m12;
    int n = ORB_SLAM2::b200::SearchByProjection(a, pts, 3.0f, 0.8f);
    n += ORB_SLAM2::b200::SearchByProjection(a, b, 7.0f, false, true);
    n += ORB_SLAM2::b200::SearchForInitialization(a, b, prev, m12, 100, 0.9f, true);
    KeyFrameT kf;
    std::set<MapPointT*> found;
    n += ORB_SLAM2::b200::SearchByProjection(a, &kf, found, 10.0f, 100, true);
    std::vector<MapPointT*> matched;
    n += ORB_SLAM2::b200::SearchByProjection(&kf, cv::Mat(), pts, matched, 10);
    n += ORB_SLAM2::b200::SearchByBoW(&kf, a, matched, 0.7f, true);
    n += ORB_SLAM2::b200::SearchByBoW(&kf, &kf, matched, 0.75f, true);
#ifndef WITH_REFERENCE_HEADERS      // the reference's MapPoint lacks the two raw-distance accessors until patched
    std::vector<bool> inView;
    n += ORB_SLAM2::b200::IsInFrustum(a, pts, 0.5f, inView);
#endif
    std::vector<cv::Mat> vd;
    n += ORB_SLAM2::b200::DistinctiveDescriptor(vd);
    return n;
}
