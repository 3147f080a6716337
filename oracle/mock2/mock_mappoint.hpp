// oracle/mock2/mock_mappoint.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Force-included when the Makefile compiles the reference's UNMODIFIED src/MapPoint.cc into
// oracle/_ref/libmappointref.so.  The real include/MapPoint.h, Frame.h, ORBmatcher.h and ORBextractor.h are used
// as they are; KeyFrame and Map (whose implementations drag in the whole SLAM system) are replaced by the
// minimal members MapPoint.cc touches, by pre-defining their include guards (include/KeyFrame.h:21, Map.h:21).
// A separate library from liborbref.so because that one mocks MapPoint itself.
#ifndef ORACLE_MOCK_MAPPOINT_HPP
#define ORACLE_MOCK_MAPPOINT_HPP

#include <algorithm>
#include <cassert>
#include <climits>
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

using namespace std;

#define KEYFRAME_H
#define MAP_H
#define ORBVOCABULARY_H
#define CONVERTER_H

namespace ORB_SLAM2 {

class MapPoint;
class Frame;

class Map {
public:
    std::mutex mMutexPointCreation;                 // include/Map.h
    void EraseMapPoint(MapPoint*) {}
};

class KeyFrame {
public:
    KeyFrame() : mnId(0), mnFrameId(0), mnScaleLevels(0), mfLogScaleFactor(0), bad(false) {}
    long unsigned int mnId, mnFrameId;
    std::vector<float> mvuRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels;
    float mfLogScaleFactor;
    cv::Mat mDescriptors, Ow;
    bool bad;
    bool isBad() { return bad; }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    void EraseMapPointMatch(const size_t&) {}
    void ReplaceMapPointMatch(const size_t&, MapPoint*) {}
};

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
