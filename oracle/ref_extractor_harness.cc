// oracle/ref_extractor_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the reference's UNMODIFIED ORBextractor
// (/root/reference/include/ORBextractor.h, /root/reference/src/ORBextractor.cc), which the
// Makefile compiles from where it lies against oracle/cvshim and links with the bump
// arena (ref_arena.cc).  Output: oracle/_ref/liborbref.so.  Used by tests/ to pin the
// restatement in orb_oracle.c and by bench.py's reference arm as the CPU baseline.
#include <chrono>
#include <thread>
#include <vector>

#include "ORBextractor.h"
#include "ref_arena.hpp"

namespace {
// reach the protected stages and tables (include/ORBextractor.h:88-111)
class Probe : public ORB_SLAM2::ORBextractor {
public:
    Probe(int n, float s, int l, int ini, int mn) : ORBextractor(n, s, l, ini, mn) {}
    using ORBextractor::ComputePyramid;
    using ORBextractor::ComputeKeyPointsOctTree;
    using ORBextractor::DistributeOctTree;
    using ORBextractor::pattern;
    using ORBextractor::umax;
    using ORBextractor::mnFeaturesPerLevel;
    using ORBextractor::mvScaleFactor;
    using ORBextractor::mvInvScaleFactor;
    using ORBextractor::mvLevelSigma2;
    using ORBextractor::mvInvLevelSigma2;
    using ORBextractor::nlevels;
};

struct RefKp { float x, y, size, angle, response; int octave, class_id; };

inline RefKp to_c(const cv::KeyPoint& k)
{
    RefKp r = { k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id };
    return r;
}
} // namespace

extern "C" {

void* orbref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    // constructor allocations (tables, pattern) stay live: no ArenaScope here
    return new Probe(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}

void orbref_extractor_destroy(void* h) { delete static_cast<Probe*>(h); }

int orbref_extractor_levels(void* h) { return static_cast<Probe*>(h)->GetLevels(); }

// Tables through the PUBLIC accessors (ORBextractor.h:64-84) plus the protected ones.
void orbref_extractor_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                             int* per_level, int* umax16, int* pattern1024)
{
    ref_arena::Scope scope;
    Probe* p = static_cast<Probe*>(h);
    {
        std::vector<float> a = p->GetScaleFactors(), b = p->GetInverseScaleFactors(),
                           c = p->GetScaleSigmaSquares(), d = p->GetInverseScaleSigmaSquares();
        for (int i = 0; i < p->nlevels; ++i) {
            if (scale) scale[i] = a[i];
            if (inv_scale) inv_scale[i] = b[i];
            if (sigma2) sigma2[i] = c[i];
            if (inv_sigma2) inv_sigma2[i] = d[i];
            if (per_level) per_level[i] = p->mnFeaturesPerLevel[i];
        }
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = p->umax[i];
    if (pattern1024) for (int i = 0; i < 512; ++i) { pattern1024[2 * i] = p->pattern[i].x; pattern1024[2 * i + 1] = p->pattern[i].y; }
}

// ORBextractor::operator() (ORBextractor.cc:1084).  Returns the keypoint count n (which may
// exceed cap; at most cap entries are written), or -1 if the outputs were left untouched
// (empty image).
int orbref_extract(void* h, const unsigned char* img, int w, int h_, size_t step,
                   RefKp* kps, unsigned char* desc, int cap)
{
    ref_arena::Scope scope;
    Probe* p = static_cast<Probe*>(h);
    int n = -1;
    {
        cv::Mat image = (img && w > 0 && h_ > 0) ? cv::Mat(h_, w, CV_8UC1, (void*)img, step) : cv::Mat();
        std::vector<cv::KeyPoint> out;
        cv::Mat descriptors;
        out.push_back(cv::KeyPoint(-12345.f, 0.f, 0.f)); // sentinel: detects "outputs untouched"
        (*p)(image, cv::Mat(), out, descriptors);
        if (!(out.size() == 1 && out[0].pt.x == -12345.f)) {
            n = (int)out.size();
            for (int i = 0; i < n && i < cap; ++i) {
                if (kps) kps[i] = to_c(out[i]);
                if (desc) std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
            }
        }
    }
    return n;
}

// mvImagePyramid[level] after a call (ORBextractor.h:86).  with_border: copy the whole
// (w+38)x(h+38) parent buffer the ROI lives in (the ROI sits at (19,19)).
int orbref_pyramid_level(void* h, int level, int with_border, unsigned char* dst, size_t dst_step, int* w, int* hgt)
{
    Probe* p = static_cast<Probe*>(h);
    if (level < 0 || level >= p->nlevels) return -1;
    const cv::Mat& m = p->mvImagePyramid[level];
    if (m.empty()) return -1;
    const int b = with_border ? 19 : 0;
    if (w) *w = m.cols + 2 * b;
    if (hgt) *hgt = m.rows + 2 * b;
    if (dst)
        for (int r = -b; r < m.rows + b; ++r)
            std::memcpy(dst + (size_t)(r + b) * dst_step, m.data + (ptrdiff_t)r * (ptrdiff_t)(size_t)m.step - b, (size_t)(m.cols + 2 * b));
    return 0;
}

// ComputePyramid + ComputeKeyPointsOctTree (ORBextractor.cc:1094-1097): per-level keypoints
// in level coordinates with orientation, before blur / descriptors / rescaling.
int orbref_keypoints_octtree(void* h, const unsigned char* img, int w, int h_, size_t step,
                             RefKp* kps, int cap, int* per_level_counts)
{
    ref_arena::Scope scope;
    Probe* p = static_cast<Probe*>(h);
    int n = 0;
    {
        cv::Mat image(h_, w, CV_8UC1, (void*)img, step);
        p->ComputePyramid(image);
        std::vector<std::vector<cv::KeyPoint> > all;
        p->ComputeKeyPointsOctTree(all);
        for (int l = 0; l < p->nlevels; ++l) {
            if (per_level_counts) per_level_counts[l] = (int)all[l].size();
            for (size_t i = 0; i < all[l].size(); ++i, ++n)
                if (n < cap && kps) kps[n] = to_c(all[l][i]);
        }
    }
    return n;
}

// DistributeOctTree (ORBextractor.cc:562) on caller-supplied candidates.
int orbref_distribute(void* h, const RefKp* in, int n_in, int minX, int maxX, int minY, int maxY, int N, int level,
                      RefKp* out, int cap)
{
    ref_arena::Scope scope;
    Probe* p = static_cast<Probe*>(h);
    int n = 0;
    {
        std::vector<cv::KeyPoint> v;
        v.reserve((size_t)n_in);
        for (int i = 0; i < n_in; ++i)
            v.push_back(cv::KeyPoint(in[i].x, in[i].y, in[i].size, in[i].angle, in[i].response, in[i].octave, in[i].class_id));
        std::vector<cv::KeyPoint> r = p->DistributeOctTree(v, minX, maxX, minY, maxY, N, level);
        n = (int)r.size();
        for (int i = 0; i < n && i < cap; ++i) out[i] = to_c(r[i]);
    }
    return n;
}

// Push extra allocations through the arena before a call: purity probe (results must not
// depend on where in the arena a call starts).
void orbref_arena_skew(size_t bytes) { (void)ref_arena::alloc(bytes); }

// CPU baseline: nframes frames (contiguous, pitch == w) round-robin over nthreads worker
// threads, one extractor instance per thread (the class is stateful).  Returns the wall
// time in seconds of the whole pass; total_kps receives the number of keypoints found.
double orbref_extract_bench(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                            const unsigned char* frames, int nframes, int w, int h_, int nthreads,
                            long long* total_kps)
{
    if (nthreads < 1) nthreads = 1;
    std::vector<long long> counts((size_t)nthreads, 0);
    std::vector<std::thread> workers;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < nthreads; ++t)
        workers.emplace_back([&, t]() {
            Probe ex(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
            for (int f = t; f < nframes; f += nthreads) {
                ref_arena::Scope scope;
                cv::Mat image(h_, w, CV_8UC1, (void*)(frames + (size_t)f * w * h_), (size_t)w);
                std::vector<cv::KeyPoint> out;
                cv::Mat descriptors;
                ex(image, cv::Mat(), out, descriptors);
                counts[(size_t)t] += (long long)out.size();
            }
        });
    for (auto& th : workers) th.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (total_kps) { *total_kps = 0; for (long long c : counts) *total_kps += c; }
    return dt;
}

int orbref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
