// ORBextractor.h -- the class a SLAM build includes INSTEAD of the reference's include/ORBextractor.h.
//
// The interface is the reference's (include/ORBextractor.h:47-111: constructor, operator(), the six Get*
// accessors, the public mvImagePyramid) because Frame.cc and Tracking.cc are compiled against it unchanged; every
// body forwards to liborb_b200.so through the C ABI of include/orb_b200.h (ORBextractor.cc next to this file).
// INTEGRATION.md has the recipe.  No CPU fallback exists: the constructor throws std::runtime_error when no CUDA
// device can be used.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

struct orbx_ctx;   // include/orb_b200.h

namespace ORB_SLAM2 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // Keypoints and 32-byte descriptors of an 8-bit single-channel image.  As in the reference the mask is ignored,
    // an empty image leaves the outputs untouched and zero keypoints release the descriptor matrix.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    // Pyramid geometry, copied out by every Frame constructor (src/Frame.cc:73-79).
    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Levels of the last image as ROIs inside their REFLECT_101-padded buffers, which is how
    // Frame::ComputeStereoMatches reads them (src/Frame.cc:520, 611-633).  Refreshed by operator() unless the
    // download is switched off below.
    std::vector<cv::Mat> mvImagePyramid;

    // ---- not in the reference ------------------------------------------------------------------------------
    static void SetDevice(int device);                      // CUDA ordinal of extractors built afterwards (default 0 or $ORB_B200_DEVICE)
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }   // off: the pyramid stays on the GPU (orbm_stereo_matches reads it there)
    orbx_ctx* Context() { return mpCtx; }                   // the C-ABI context, for the batched / stereo entry points

protected:
    // the reference's members (include/ORBextractor.h:98-111); scaleFactor really is a double there
    int nfeatures;
    double scaleFactor;
    int nlevels, iniThFAST, minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    orbx_ctx* mpCtx;
    bool mbDownloadPyramid;

private:
    ORBextractor(const ORBextractor&);              // one CUDA context per instance: not copyable
    ORBextractor& operator=(const ORBextractor&);
};

}  // namespace ORB_SLAM2

#endif
