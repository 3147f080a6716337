// oracle/mock/mock_slam.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Force-included (-include) when the Makefile compiles the reference's UNMODIFIED
// src/ORBmatcher.cc and src/Frame.cc.  The real include/Frame.h, ORBmatcher.h and
// ORBextractor.h are used as they are; only the types whose implementation is out of scope
// for the hot path (KeyFrame, MapPoint, ORBVocabulary, Converter) are replaced by minimal
// mocks, by pre-defining their include guards (include/KeyFrame.h:21, MapPoint.h:21,
// ORBVocabulary.h:22, Converter.h:21).  SURVEY.md App. D.
#ifndef ORACLE_MOCK_SLAM_HPP
#define ORACLE_MOCK_SLAM_HPP

#include <cassert>
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "cvshim.hpp"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

using namespace std;   // the reference headers use bare vector<> / pair<> (Frame.h:100, ORBmatcher.h:80)

#define KEYFRAME_H
#define MAPPOINT_H
#define ORBVOCABULARY_H
#define CONVERTER_H

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;
class Map;

class MapPoint {
public:
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0),
                 mTrackViewCos(0), mnTrackReferenceForFrame(0), mnLastFrameSeen(0), mnId(0), nObs(0), bad(false),
                 maxDist(1e9f), minDist(0.f) {}
    // tracking variables read by ORBmatcher::SearchByProjection (src/ORBmatcher.cc:82-124)
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    long unsigned int mnTrackReferenceForFrame, mnLastFrameSeen, mnId;
    // state
    int nObs;
    bool bad;
    cv::Mat descriptor, worldPos, normal;
    float maxDist, minDist;
    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return descriptor.clone(); }
    cv::Mat GetWorldPos() { return worldPos.clone(); }
    cv::Mat GetNormal() { return normal.clone(); }
    float GetMaxDistanceInvariance() { return maxDist; }   // the harness stores the invariance bounds directly
    float GetMinDistanceInvariance() { return minDist; }
    int PredictScale(const float&, KeyFrame*) { return mnTrackScaleLevel; }
    int PredictScale(const float&, Frame*) { return mnTrackScaleLevel; }
    int GetIndexInKeyFrame(KeyFrame*) { return -1; }
    bool IsInKeyFrame(KeyFrame*) { return false; }
    void Replace(MapPoint*) {}
    void AddObservation(KeyFrame*, size_t) {}
    void IncreaseVisible(int = 1) {}
    void IncreaseFound(int = 1) {}
};

class KeyFrame {
public:
    KeyFrame() : N(0), fx(0), fy(0), cx(0), cy(0), invfx(0), invfy(0), mbf(0), mb(0), mThDepth(0), mnId(0) {}
    int N;
    float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    long unsigned int mnId;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<MapPoint*> mapPoints;
    cv::Mat R, t, Ow;
    std::vector<size_t> GetFeaturesInArea(const float&, const float&, const float&) const { return std::vector<size_t>(); }
    bool IsInImage(const float&, const float&) const { return true; }
    std::vector<MapPoint*> GetMapPointMatches() { return mapPoints; }
    std::set<MapPoint*> GetMapPoints() { return std::set<MapPoint*>(mapPoints.begin(), mapPoints.end()); }
    MapPoint* GetMapPoint(const size_t& i) { return mapPoints[i]; }
    void AddMapPoint(MapPoint* p, const size_t& i) { mapPoints[i] = p; }
    cv::Mat GetRotation() { return R.clone(); }
    cv::Mat GetTranslation() { return t.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    cv::Mat GetPose() { return cv::Mat(); }
    bool isBad() { return false; }
};

// ORBVocabulary / Converter are only reached from Frame::ComputeBoW (src/Frame.cc:425-433)
class ORBVocabulary {
public:
    void transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) { std::abort(); }
};
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat&) { std::abort(); return std::vector<cv::Mat>(); }
};

} // namespace ORB_SLAM2

#endif
