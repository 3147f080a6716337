// oracle/ref_mappoint_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the reference's UNMODIFIED src/MapPoint.cc (compiled where it lies under /root/reference
// into oracle/_ref/libmappointref.so, see Makefile): MapPoint::ComputeDistinctiveDescriptors (:275-340) and
// MapPoint::PredictScale (:442-475).  KeyFrame and Map are the mocks of mock2/mock_mappoint.hpp.
// The observations map is keyed by KeyFrame POINTER (include/MapPoint.h:126), so the order of the descriptors --
// and with it the winner among equal medians -- follows the key frames' addresses; the harness keeps the key
// frames in one vector, i.e. address order = input order.
#include <cstdint>
#include <cstring>
#include <vector>

#include "MapPoint.h"

using namespace ORB_SLAM2;

extern "C" {

// desc [n][32]; kf_bad [n] or NULL.  Returns 1 and the chosen descriptor in out32, or 0 when the reference leaves
// mDescriptor untouched (no usable observation).
int orbref_mp_distinctive(const uint8_t* desc, int n, const uint8_t* kf_bad, uint8_t* out32)
{
    Map map;
    std::vector<KeyFrame> kfs((size_t)(n > 0 ? n : 1));
    for (size_t i = 0; i < kfs.size(); ++i) {
        KeyFrame& k = kfs[i];
        k.mnId = i;
        k.mvuRight.assign(1, -1.0f);
        k.mvKeysUn.resize(1);
        k.mvScaleFactors.assign(1, 1.0f);
        k.mnScaleLevels = 1;
        k.Ow = cv::Mat::zeros(3, 1, CV_32F);
        k.mDescriptors.create(1, 32, CV_8U);
        if ((int)i < n) memcpy(k.mDescriptors.ptr<uint8_t>(0), desc + 32 * i, 32);
        k.bad = kf_bad && (int)i < n && kf_bad[i];
    }
    cv::Mat pos = cv::Mat::zeros(3, 1, CV_32F);
    MapPoint mp(pos, &kfs[0], &map);
    for (int i = 0; i < n; ++i) mp.AddObservation(&kfs[(size_t)i], 0);
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    if (d.empty()) return 0;
    memcpy(out32, d.ptr<uint8_t>(0), 32);
    return 1;
}

// PredictScale(currentDist, KeyFrame*) for a point whose mfMaxDistance is exactly max_distance (set through
// UpdateNormalAndDepth with the reference key frame at the origin, level 0 and the point at (0, 0, max_distance)).
// invariance2 (or NULL) receives GetMinDistanceInvariance(), GetMaxDistanceInvariance().
int orbref_mp_predict_scale(float max_distance, const float* cur_dist, int n, float scale_factor, int nlevels, int* out,
                            float* invariance2)
{
    Map map;
    KeyFrame kf;
    kf.mvuRight.assign(1, -1.0f);
    kf.mvKeysUn.resize(1);
    kf.mvKeysUn[0].octave = 0;
    kf.mvScaleFactors.resize((size_t)nlevels);
    kf.mvScaleFactors[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) kf.mvScaleFactors[(size_t)i] = kf.mvScaleFactors[(size_t)i - 1] * scale_factor;
    kf.mnScaleLevels = nlevels;
    kf.mfLogScaleFactor = log(scale_factor);                    // as src/Frame.cc:75
    kf.Ow = cv::Mat::zeros(3, 1, CV_32F);
    cv::Mat pos = cv::Mat::zeros(3, 1, CV_32F);
    pos.at<float>(2) = max_distance;
    MapPoint mp(pos, &kf, &map);
    mp.AddObservation(&kf, 0);
    mp.UpdateNormalAndDepth();
    if (invariance2) { invariance2[0] = mp.GetMinDistanceInvariance(); invariance2[1] = mp.GetMaxDistanceInvariance(); }
    for (int i = 0; i < n; ++i) out[i] = mp.PredictScale(cur_dist[i], &kf);
    return 0;
}

} // extern "C"
