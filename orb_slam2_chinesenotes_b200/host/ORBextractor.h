// ORBextractor.h -- drop-in replacement of the reference's include/ORBextractor.h.
//
// Same class, same public interface (constructor, operator(), Get* accessors, public
// mvImagePyramid: reference include/ORBextractor.h:47-111); the bodies forward to the B200 library
// through the C ABI of include/orb_b200.h.  A SLAM build replaces the reference's ORBextractor.h/.cc
// with this pair and links liborb_b200.so; nothing else changes.
// There is no CPU fallback: construction throws std::runtime_error when no CUDA device is usable.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>

struct orbx_ctx;

namespace ORB_SLAM2
{

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels,
                 int iniThFAST, int minThFAST);

    ~ORBextractor();

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int inline GetLevels(){
        return nlevels;}

    float inline GetScaleFactor(){
        return scaleFactor;}

    std::vector<float> inline GetScaleFactors(){
        return mvScaleFactor;
    }

    std::vector<float> inline GetInverseScaleFactors(){
        return mvInvScaleFactor;
    }

    std::vector<float> inline GetScaleSigmaSquares(){
        return mvLevelSigma2;
    }

    std::vector<float> inline GetInverseScaleSigmaSquares(){
        return mvInvLevelSigma2;
    }

    // Filled after every call when pyramid download is on (default): level ROIs inside
    // REFLECT_101-padded buffers, exactly what Frame::ComputeStereoMatches reads (src/Frame.cc:520,611-633).
    std::vector<cv::Mat> mvImagePyramid;

    // --- additions (not in the reference) ---------------------------------------------------------
    // CUDA ordinal for extractors constructed afterwards (default 0, or $ORB_B200_DEVICE).
    static void SetDevice(int device);
    // Skip the device->host copy of the pyramid (use orbm_stereo_matches on the device instead).
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }
    // The C-ABI context (for orbm_stereo_matches and batched calls).
    orbx_ctx* Context() { return mpCtx; }

protected:

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbx_ctx* mpCtx;
    bool mbDownloadPyramid;

private:
    ORBextractor(const ORBextractor&);
    ORBextractor& operator=(const ORBextractor&);
};

} //namespace ORB_SLAM

#endif
