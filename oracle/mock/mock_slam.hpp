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

#include <algorithm>
#include <cassert>
#include <cmath>
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
                 maxDist(1e9f), minDist(0.f), maxDistRaw(0.f), realPredict(false), replacedBy(NULL) {}
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
    // realPredict: the arithmetic of src/MapPoint.cc:459-475 on maxDistRaw (= mfMaxDistance), defined in
    // ref_frustum_harness.cc where Frame is complete; pinned against the reference's own MapPoint.cc through
    // oracle/_ref/libmappointref.so (tests/test_projection_oracle.py).  Otherwise the preset level.
    float maxDistRaw;
    bool realPredict;
    int PredictScale(const float& d, Frame* pF) { return realPredict ? PredictScaleReal(d, pF) : mnTrackScaleLevel; }
    int PredictScaleReal(const float& currentDist, Frame* pF);
    // The observation bookkeeping that the loops of ORBmatcher::Fuse / SearchBySim3 can see (src/ORBmatcher.cc:1385, 1487-1502,
    // 869-875): which key frames observe the point at which keypoint, and what MapPoint::AddObservation (src/MapPoint.cc:93-105)
    // and MapPoint::Replace (:204-258) do to it.  Empty unless a harness fills it, so the other harnesses see the old
    // "never in a key frame" behaviour.  Defined below, where KeyFrame is complete.
    std::map<KeyFrame*, size_t> obs;
    MapPoint* replacedBy;
    int GetIndexInKeyFrame(KeyFrame* kf) { std::map<KeyFrame*, size_t>::iterator it = obs.find(kf); return it == obs.end() ? -1 : (int)it->second; }
    bool IsInKeyFrame(KeyFrame* kf) { return obs.count(kf) != 0; }
    inline void Replace(MapPoint* p);
    inline void AddObservation(KeyFrame* kf, size_t idx);
    void IncreaseVisible(int = 1) {}
    void IncreaseFound(int = 1) {}
};

class KeyFrame {
public:
    KeyFrame() : N(0), fx(0), fy(0), cx(0), cy(0), invfx(0), invfy(0), mbf(0), mb(0), mThDepth(0), mnId(0), mnFrameId(0), mTimeStamp(0),
                 mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0), mfGridElementWidthInv(0), mfGridElementHeightInv(0) {}
    int N;
    float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    long unsigned int mnId;
    long unsigned int mnFrameId;      // include/KeyFrame.h:126: id of the Frame the key frame was made from
    double mTimeStamp;                // include/KeyFrame.h:128
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<MapPoint*> mapPoints;
    cv::Mat R, t, Ow;
    // The key frame's copy of the frame grid (src/KeyFrame.cc:33-55 copies F.mGrid and the cell sizes); the harness
    // fills it from a real ORB_SLAM2::Frame gridded by the reference's own AssignFeaturesToGrid.
    int mnMinX, mnMinY, mnMaxX, mnMaxY;
    float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::vector<std::vector<size_t> > > mGrid;      // [col][row] -> keypoint indices
    // Restatement of KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:637-676 (KeyFrame.cc itself drags in Map,
    // KeyFrameDatabase and the vocabulary, so it is not compiled): cells floor/ceil of the window, clamped;
    // columns outer, rows inner, insertion order inside a cell; strict |dx| < r and |dy| < r.  Empty grid = no hits.
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const
    {
        std::vector<size_t> hits;
        const int ncols = (int)mGrid.size(), nrows = ncols ? (int)mGrid[0].size() : 0;
        if (!ncols || !nrows) return hits;
        const int c0 = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
        const int c1 = std::min(ncols - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
        const int r0 = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
        const int r1 = std::min(nrows - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (c0 >= ncols || c1 < 0 || r0 >= nrows || r1 < 0) return hits;
        for (int c = c0; c <= c1; ++c)
            for (int q = r0; q <= r1; ++q)
                for (size_t j = 0; j < mGrid[c][q].size(); ++j) {
                    const cv::KeyPoint& kp = mvKeysUn[mGrid[c][q][j]];
                    if (std::fabs(kp.pt.x - x) < r && std::fabs(kp.pt.y - y) < r) hits.push_back(mGrid[c][q][j]);
                }
        return hits;
    }
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }   // :679-682
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

// src/MapPoint.cc:93-105: a key frame is recorded once; a stereo observation counts twice
inline void MapPoint::AddObservation(KeyFrame* kf, size_t idx)
{
    if (obs.count(kf)) return;
    obs[kf] = idx;
    nObs += (idx < kf->mvuRight.size() && kf->mvuRight[idx] >= 0) ? 2 : 1;
}
// src/MapPoint.cc:204-258: this point goes bad; every key frame that observed it now observes p at the same keypoint, unless
// it observes p already, in which case the keypoint loses its point
inline void MapPoint::Replace(MapPoint* p)
{
    if (p == this) return;
    std::map<KeyFrame*, size_t> o = obs;
    obs.clear();
    bad = true;
    replacedBy = p;
    for (std::map<KeyFrame*, size_t>::iterator it = o.begin(); it != o.end(); ++it) {
        KeyFrame* kf = it->first;
        if (!p->IsInKeyFrame(kf)) { kf->mapPoints[it->second] = p; p->AddObservation(kf, it->second); }
        else kf->mapPoints[it->second] = NULL;
    }
}

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
