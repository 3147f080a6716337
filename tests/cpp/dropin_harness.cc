// tests/cpp/dropin_harness.cc -- test-only C wrapper around the DROP-IN C++ class
// (orb_slam2_chinesenotes_b200/host/ORBextractor.{h,cc}) so pytest can call it the way
// Frame::ExtractORB does (src/Frame.cc:262-268).  Built against oracle/cvshim, which stands in
// for OpenCV in this image (test infrastructure; a SLAM build uses the real OpenCV).
#include "ORBextractor.h"

#include <chrono>
#include <cstring>

struct Kp { float x, y, size, angle, response; int octave, class_id; };

extern "C" {

void* dropin_create(int nfeatures, float scaleFactor, int nlevels, int ini, int mn)
{
    try { return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, ini, mn); }
    catch (...) { return 0; }
}
void dropin_destroy(void* h) { delete static_cast<ORB_SLAM2::ORBextractor*>(h); }

int dropin_accessors(void* h, float* scale, float* inv, float* s2, float* is2, float* scale_factor)
{
    ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(h);
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->GetLevels(); ++i) { scale[i] = a[i]; inv[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; }
    *scale_factor = e->GetScaleFactor();
    return e->GetLevels();
}

// (*extractor)(image, cv::Mat(), keypoints, descriptors); returns n, or -1 when the outputs were left untouched
int dropin_call(void* h, const unsigned char* img, int w, int hgt, size_t step, Kp* kps, unsigned char* desc, int cap)
{
    ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(h);
    cv::Mat image = (img && w > 0 && hgt > 0) ? cv::Mat(hgt, w, CV_8UC1, (void*)img, step) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    keys.push_back(cv::KeyPoint(-12345.f, 0.f, 0.f));
    cv::Mat descriptors;
    (*e)(image, cv::Mat(), keys, descriptors);
    if (keys.size() == 1 && keys[0].pt.x == -12345.f) return -1;
    const int n = (int)keys.size();
    for (int i = 0; i < n && i < cap; ++i) {
        Kp k = { keys[i].pt.x, keys[i].pt.y, keys[i].size, keys[i].angle, keys[i].response, keys[i].octave, keys[i].class_id };
        kps[i] = k;
        std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    return n;
}

// Wall-clock milliseconds per (*extractor)(image, Mat(), keypoints, descriptors) call, measured inside C++ (no ctypes in
// the loop): what Frame::ExtractORB costs, mvImagePyramid included when the download is on.
double dropin_time_ms(void* h, const unsigned char* img, int w, int hgt, size_t step, int reps, int with_pyramid)
{
    ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(h);
    e->SetPyramidDownload(with_pyramid != 0);
    cv::Mat image(hgt, w, CV_8UC1, (void*)img, step);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    for (int i = 0; i < 5; ++i) (*e)(image, cv::Mat(), keys, descriptors);
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) (*e)(image, cv::Mat(), keys, descriptors);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
    e->SetPyramidDownload(true);
    return keys.empty() ? -1.0 : ms;
}

// mvImagePyramid[level]: ROI size and, like Frame.cc does, pixels addressed relative to the ROI (the
// 19-px border is reachable with negative offsets).
int dropin_pyramid(void* h, int level, int border, unsigned char* dst, size_t dst_step, int* w, int* hgt)
{
    ORB_SLAM2::ORBextractor* e = static_cast<ORB_SLAM2::ORBextractor*>(h);
    const cv::Mat& m = e->mvImagePyramid[level];
    if (m.empty()) return -1;
    *w = m.cols; *hgt = m.rows;
    if (dst)
        for (int r = -border; r < m.rows + border; ++r)
            std::memcpy(dst + (size_t)(r + border) * dst_step, m.data + (ptrdiff_t)r * (ptrdiff_t)(size_t)m.step - border, (size_t)(m.cols + 2 * border));
    return 0;
}

}
