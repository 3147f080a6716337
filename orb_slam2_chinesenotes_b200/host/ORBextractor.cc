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

    // DistributeOctTree keeps at least its quota per level and at most a few more: start with
    // nfeatures plus slack and grow on the (rare) capacity status.
    int capacity = nfeatures + 4*nlevels + 64;
    std::vector<orbx_kp> kps;
    std::vector<unsigned char> desc;
    int n = 0;
    for(;;)
    {
        kps.resize(capacity);
        desc.resize((size_t)capacity*32);
        const int rc = orbx_extract(mpCtx, image.data, image.cols, image.rows, (size_t)image.step,
                                    &kps[0], &desc[0], capacity, &n);
        if(rc == ORBX_OK)
            break;
        if(rc == ORBX_E_CAPACITY) { capacity = n + 64; continue; }
        throw std::runtime_error(std::string("ORBextractor(B200): ") + orbx_last_error(mpCtx));
    }

    if( n == 0 )
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        for(int i=0; i<n; i++)
            std::memcpy(descriptors.ptr(i), &desc[(size_t)i*32], 32);
    }

    _keypoints.clear();
    _keypoints.reserve(n);
    for(int i=0; i<n; i++)
        _keypoints.push_back(cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response,
                                          kps[i].octave, kps[i].class_id));

    if(mbDownloadPyramid)
    {
        const int EDGE_THRESHOLD = 19;
        for(int level=0; level<nlevels; ++level)
        {
            int w=0, h=0;
            orbx_pyramid_level(mpCtx, 0, level, 1, NULL, 0, &w, &h);
            cv::Mat temp(h, w, CV_8UC1);
            if(orbx_pyramid_level(mpCtx, 0, level, 1, temp.data, (size_t)temp.step, NULL, NULL) != ORBX_OK)
                throw std::runtime_error(std::string("ORBextractor(B200): ") + orbx_last_error(mpCtx));
            mvImagePyramid[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w-2*EDGE_THRESHOLD, h-2*EDGE_THRESHOLD));
        }
    }
}

} //namespace ORB_SLAM
