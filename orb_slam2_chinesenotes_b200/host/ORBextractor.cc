// ORBextractor.cc -- host forwarder of the drop-in ORBextractor (see ORBextractor.h).
// Replaces the bodies of the reference's src/ORBextractor.cc; every step of operator()
// (pyramid, per-cell FAST, quadtree, orientation, blur, rBRIEF: src/ORBextractor.cc:1084-1150)
// runs on the GPU behind orbx_extract().
#include "ORBextractor.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "orb_b200.h"

namespace ORB_SLAM2
{

static int gDevice = -1;

void ORBextractor::SetDevice(int device) { gDevice = device; }

static int CurrentDevice()
{
    if (gDevice >= 0) return gDevice;
    const char* e = std::getenv("ORB_B200_DEVICE");
    return e ? std::atoi(e) : 0;
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels,
                           int _iniThFAST, int _minThFAST) :
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST), mpCtx(NULL), mbDownloadPyramid(true)
{
    const int rc = orbx_create(&mpCtx, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, CurrentDevice());
    if (rc != ORBX_OK)
        throw std::runtime_error("ORBextractor(B200): orbx_create failed with status " + std::to_string(rc) +
                                 " (no usable CUDA device? there is no CPU fallback)");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orbx_tables(mpCtx, &mvScaleFactor[0], &mvInvScaleFactor[0], &mvLevelSigma2[0], &mvInvLevelSigma2[0], &mnFeaturesPerLevel[0]);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor()
{
    orbx_destroy(mpCtx);
}

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                               cv::OutputArray _descriptors)
{
    if(_image.empty())
        return;                                             // outputs untouched, as in the reference

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );

    // One call: upload, every kernel and the download of keypoints, descriptors and (if wanted) all padded pyramid
    // levels replay as one CUDA graph; `out` points into pinned memory owned by the context (include/orb_b200.h).
    orbx_frame_out out;
    const int rc = orbx_extract_frame(mpCtx, image.data, image.cols, image.rows, (size_t)image.step, mbDownloadPyramid ? 1 : 0, &out);
    if(rc != ORBX_OK)
        throw std::runtime_error(std::string("ORBextractor(B200): ") + orbx_last_error(mpCtx));
    const int n = out.n;

    if( n == 0 )
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        if(descriptors.isContinuous())
            std::memcpy(descriptors.ptr(0), out.desc, (size_t)n*32);
        else
            for(int i=0; i<n; i++)
                std::memcpy(descriptors.ptr(i), out.desc + (size_t)i*32, 32);
    }

    _keypoints.clear();
    _keypoints.reserve(n);
    for(int i=0; i<n; i++)
        _keypoints.push_back(cv::KeyPoint(out.kps[i].x, out.kps[i].y, out.kps[i].size, out.kps[i].angle, out.kps[i].response,
                                          out.kps[i].octave, out.kps[i].class_id));

    if(mbDownloadPyramid)
    {
        // As in the reference (src/ORBextractor.cc:1157-1178) mvImagePyramid[level] is a view inside its REFLECT_101-padded
        // buffer.  The buffers live in the context's pinned memory and are overwritten by the next operator() call on this
        // extractor -- which is when the reference replaces them too.
        const int EDGE_THRESHOLD = 19;
        for(int level=0; level<nlevels; ++level)
        {
            const int w = out.level_w[level], h = out.level_h[level];
            unsigned char* padded = const_cast<unsigned char*>(out.level[level]) - (size_t)EDGE_THRESHOLD*out.level_pitch[level] - EDGE_THRESHOLD;
            cv::Mat temp(h + 2*EDGE_THRESHOLD, w + 2*EDGE_THRESHOLD, CV_8UC1, padded, out.level_pitch[level]);
            mvImagePyramid[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
        }
    }
}

} //namespace ORB_SLAM
