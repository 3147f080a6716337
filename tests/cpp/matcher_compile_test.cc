// tests/cpp/matcher_compile_test.cc -- instantiates the ORBmatcher forwarders
// (host/ORBmatcher_b200.hpp) against Frame / MapPoint types that expose the same members as the
// reference's classes, and exposes them to pytest.  Where /root/reference exists the Makefile target
// `matcher_ref` compiles the same header against the reference's REAL include/Frame.h instead.
#ifdef WITH_REFERENCE_HEADERS
#include "Frame.h"
typedef ORB_SLAM2::Frame FrameT;
typedef ORB_SLAM2::MapPoint MapPointT;
typedef ORB_SLAM2::KeyFrame KeyFrameT;
#else
#include "cvshim.hpp"
#include <map>
#include <set>
#include <vector>
struct KeyFrameT;
struct MapPointT {
    bool mbTrackInView, bad; int mnTrackScaleLevel, nObs; float mTrackViewCos, mTrackProjX, mTrackProjY, mTrackProjXR;
    cv::Mat descriptor, pos, normalv;
    float minD, maxD; bool haveRange;                         // zero-initialised by the vectors that hold the points: no range set
    bool isBad() { return bad; }
    float GetMinDistanceInvariance() { return haveRange ? minD : 0.f; }
    float GetMaxDistanceInvariance() { return haveRange ? maxD : 1e9f; }
    // observation bookkeeping as the reference's MapPoint keeps it (src/MapPoint.cc:93-105, 204-258), for Fuse / SearchBySim3
    std::map<KeyFrameT*, size_t> obs;
    MapPointT* replacedBy;
    bool IsInKeyFrame(KeyFrameT* kf) { return obs.count(kf) != 0; }
    int GetIndexInKeyFrame(KeyFrameT* kf) { std::map<KeyFrameT*, size_t>::iterator it = obs.find(kf); return it == obs.end() ? -1 : (int)it->second; }
    inline void AddObservation(KeyFrameT* kf, size_t idx);
    inline void Replace(MapPointT* p);
    template <class F> int PredictScale(const float&, F*) { return mnTrackScaleLevel; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return descriptor.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
    cv::Mat GetNormal() { return normalv.empty() ? pos.clone() : normalv.clone(); }
    float GetMaxDistance() { return 1e9f; }
    float GetMinDistance() { return 0.f; }
};
#include <map>
struct FeatVecT : std::map<unsigned int, std::vector<unsigned int> > {};
static unsigned long NextFrameId() { static unsigned long n = 0; return n++; }   // Frame::nNextId++ (src/Frame.cc:213)
struct FrameT {
    long unsigned int mnId; double mTimeStamp;                                     // include/Frame.h:120,131
    FrameT() : mnId(NextFrameId()), mTimeStamp(0.0) {}
    int N, mnScaleLevels; float mbf, mb, mfScaleFactor;
    FeatVecT mFeatVec;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    cv::Mat mDescriptors, mTcw;
    std::vector<MapPointT*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY;
};
struct KeyFrameT {
    long unsigned int mnFrameId; double mTimeStamp;                                // include/KeyFrame.h:128,130: the id of the Frame it was made from
    KeyFrameT() : mnFrameId(NextFrameId()), mTimeStamp(0.0) {}
    float fx, fy, cx, cy, mbf; int mnMinX, mnMinY, mnMaxX, mnMaxY;
    cv::Mat mDescriptors, Rcw, tcw, Ow;
    FeatVecT mFeatVec;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2, mvLevelSigma2, mvuRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPointT*> pts;
    std::vector<MapPointT*> GetMapPointMatches() { return pts; }
    std::set<MapPointT*> GetMapPoints() { std::set<MapPointT*> s; for (size_t i = 0; i < pts.size(); ++i) if (pts[i] && !pts[i]->isBad()) s.insert(pts[i]); return s; }   // src/KeyFrame.cc:378-392
    MapPointT* GetMapPoint(const size_t& i) { return pts[i]; }
    void AddMapPoint(MapPointT* p, const size_t& i) { pts[i] = p; }
    cv::Mat GetRotation() { return Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
};
inline void MapPointT::AddObservation(KeyFrameT* kf, size_t idx)
{
    if (obs.count(kf)) return;
    obs[kf] = idx;
    nObs += (idx < kf->mvuRight.size() && kf->mvuRight[idx] >= 0) ? 2 : 1;
}
inline void MapPointT::Replace(MapPointT* p)
{
    if (p == this) return;
    std::map<KeyFrameT*, size_t> o = obs;
    obs.clear();
    bad = true;
    replacedBy = p;
    for (std::map<KeyFrameT*, size_t>::iterator it = o.begin(); it != o.end(); ++it) {
        if (!p->IsInKeyFrame(it->first)) { it->first->pts[it->second] = p; p->AddObservation(it->first, it->second); }
        else it->first->pts[it->second] = 0;
    }
}
float FrameT::fx, FrameT::fy, FrameT::cx, FrameT::cy, FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY;
#endif
#include "ORBmatcher_b200.hpp"

extern "C" int matcher_forwarders_instantiate(int run)
{
    if (!run) return 0;   // instantiation is the test; running needs a GPU and populated frames
    FrameT a, b;
    std::vector<MapPointT*> pts;
    std::vector<cv::Point2f> prev;
    std::vector<int> m12;
    int n = ORB_SLAM2::b200::SearchByProjection(a, pts, 3.0f, 0.8f);
    n += ORB_SLAM2::b200::SearchByProjection(a, b, 7.0f, false, true);
    n += ORB_SLAM2::b200::SearchForInitialization(a, b, prev, m12, 100, 0.9f, true);
    KeyFrameT kf;
    std::set<MapPointT*> found;
    n += ORB_SLAM2::b200::SearchByProjection(a, &kf, found, 10.0f, 100, true);
    std::vector<MapPointT*> matched;
    n += ORB_SLAM2::b200::SearchByProjection(&kf, cv::Mat(), pts, matched, 10);
    n += ORB_SLAM2::b200::SearchByBoW(&kf, a, matched, 0.7f, true);
    n += ORB_SLAM2::b200::SearchByBoW(&kf, &kf, matched, 0.75f, true);
    n += ORB_SLAM2::b200::Fuse(&kf, pts, 3.0f);
    n += ORB_SLAM2::b200::Fuse(&kf, cv::Mat(), pts, 4.0f, matched);
    const float s12 = 1.f;
    n += ORB_SLAM2::b200::SearchBySim3(&kf, &kf, matched, s12, cv::Mat(), cv::Mat(), 7.5f);
    std::vector<std::pair<size_t, size_t> > pairs;
    n += ORB_SLAM2::b200::SearchForTriangulation(&kf, &kf, cv::Mat(), pairs, false, true);
#ifndef WITH_REFERENCE_HEADERS      // the reference's MapPoint lacks the two raw-distance accessors until patched
    std::vector<bool> inView;
    n += ORB_SLAM2::b200::IsInFrustum(a, pts, 0.5f, inView);
#endif
    std::vector<cv::Mat> vd;
    n += ORB_SLAM2::b200::DistinctiveDescriptor(vd);
    return n;
}

#ifndef WITH_REFERENCE_HEADERS
// ---- runtime entry points for pytest (GPU): objects of the stand-in types are filled from arrays and handed to the
// forwarders exactly as the patched ORBmatcher.cc / Tracking.cc / MapPoint.cc would hand over the SLAM system's own.
#include <chrono>
#include <cstring>
namespace {
struct Kp { float x, y, size, angle, response; int octave, class_id; };
void fill_keys(std::vector<cv::KeyPoint>& v, const Kp* k, int n)
{
    v.resize(n);
    for (int i = 0; i < n; ++i) v[i] = cv::KeyPoint(k[i].x, k[i].y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id);
}
void fill_desc(cv::Mat& m, const unsigned char* d, int n)
{
    m.create(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(m.data, d, (size_t)n * 32);
}
void fill_featvec(FeatVecT& fv, int nn, const int* id, const int* off, const int* feat)
{
    for (int k = 0; k < nn; ++k) for (int j = off[k]; j < off[k + 1]; ++j) fv[(unsigned)id[k]].push_back((unsigned)feat[j]);
}
} // namespace

extern "C" {

// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) through b200::SearchByProjection
int fwd_search_by_projection_points(int n, const Kp* kps, const unsigned char* desc, const float* uright, const float* scale, int nlevels,
                                    float minX, float maxX, float minY, float maxY, int nq, const float* proj, const int* level,
                                    const float* view_cos, const unsigned char* in_view, const unsigned char* bad, const int* obs,
                                    const unsigned char* qdesc, const int* init_assign, int* assign_out, float th, float nnratio)
{
    FrameT F;
    F.N = n;
    fill_keys(F.mvKeysUn, kps, n); F.mvKeys = F.mvKeysUn;
    fill_desc(F.mDescriptors, desc, n);
    F.mvuRight.assign(n, -1.f);
    if (uright) F.mvuRight.assign(uright, uright + n);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    FrameT::mnMinX = minX; FrameT::mnMaxX = maxX; FrameT::mnMinY = minY; FrameT::mnMaxY = maxY;
    std::vector<MapPointT> store(nq > 0 ? nq : 1);
    std::vector<MapPointT*> pts(nq);
    for (int i = 0; i < nq; ++i) {
        MapPointT& m = store[i];
        m.mbTrackInView = in_view[i] != 0; m.bad = bad[i] != 0; m.mnTrackScaleLevel = level[i]; m.nObs = obs[i];
        m.mTrackViewCos = view_cos[i]; m.mTrackProjX = proj[3 * i]; m.mTrackProjY = proj[3 * i + 1]; m.mTrackProjXR = proj[3 * i + 2];
        fill_desc(m.descriptor, qdesc + 32 * (size_t)i, 1);
        pts[i] = &m;
    }
    F.mvpMapPoints.assign(n, static_cast<MapPointT*>(0));
    if (init_assign) for (int k = 0; k < n; ++k) if (init_assign[k] >= 0) F.mvpMapPoints[k] = &store[init_assign[k]];
    const int nm = ORB_SLAM2::b200::SearchByProjection(F, pts, th, nnratio);
    for (int k = 0; k < n; ++k) assign_out[k] = F.mvpMapPoints[k] ? (int)(F.mvpMapPoints[k] - &store[0]) : -1;
    return nm;
}

// One Frame, `calls` searches: how many uploads the forwarders' device-side frame cache needed (1 when it works) and
// whether every call returned the same matches; seconds[0] / seconds[1] = host time of the first call / mean of the rest.
int fwd_resident_reuse(int calls, int n, const Kp* kps, const unsigned char* desc, const float* scale, int nlevels,
                       float minX, float maxX, float minY, float maxY, int nq, const float* proj, const int* level,
                       const float* view_cos, const unsigned char* qdesc, int* nmatches_out, double* seconds)
{
    FrameT F;
    F.N = n;
    fill_keys(F.mvKeysUn, kps, n); F.mvKeys = F.mvKeysUn;
    fill_desc(F.mDescriptors, desc, n);
    F.mvuRight.assign(n, -1.f);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    FrameT::mnMinX = minX; FrameT::mnMaxX = maxX; FrameT::mnMinY = minY; FrameT::mnMaxY = maxY;
    std::vector<MapPointT> store(nq > 0 ? nq : 1);
    std::vector<MapPointT*> pts(nq);
    for (int i = 0; i < nq; ++i) {
        MapPointT& m = store[i];
        m.mbTrackInView = true; m.bad = false; m.mnTrackScaleLevel = level[i]; m.nObs = 1;
        m.mTrackViewCos = view_cos[i]; m.mTrackProjX = proj[3 * i]; m.mTrackProjY = proj[3 * i + 1]; m.mTrackProjXR = proj[3 * i + 2];
        fill_desc(m.descriptor, qdesc + 32 * (size_t)i, 1);
        pts[i] = &m;
    }
    ORB_SLAM2::b200::ResidentFrames& cache = ORB_SLAM2::b200::ResidentFrames::Local();
    const unsigned long before = cache.Uploads();
    int first = -1, same = 1;
    seconds[0] = seconds[1] = 0.0;
    for (int c = 0; c < calls; ++c) {
        F.mvpMapPoints.assign(n, static_cast<MapPointT*>(0));
        const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        const int nm = ORB_SLAM2::b200::SearchByProjection(F, pts, 3.0f, 0.8f);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (c == 0) { first = nm; seconds[0] = dt; } else { seconds[1] += dt / (calls - 1); if (nm != first) same = 0; }
    }
    *nmatches_out = first;
    return same ? (int)(cache.Uploads() - before) : -1;
}

// b200::ResidentFrames bookkeeping: `frames` small Frames are made resident one after the other (capacity 8, least recently
// used entry dropped), then the first one is asked for again (evicted: one more upload) and the last one (still there: none).
// Returns uploads_total * 100 + entries held, or -1 if a view does not describe its frame.
int fwd_resident_cache_lru(int frames, int n, const Kp* kps, const unsigned char* desc)
{
    ORB_SLAM2::b200::ResidentFrames& cache = ORB_SLAM2::b200::ResidentFrames::Local();
    cache.Clear();
    const unsigned long before = cache.Uploads();
    std::vector<FrameT> F(frames);
    for (int f = 0; f < frames; ++f) {
        F[f].N = n - f;                                                     // different keypoint counts
        fill_keys(F[f].mvKeysUn, kps, n - f); F[f].mvKeys = F[f].mvKeysUn;
        fill_desc(F[f].mDescriptors, desc, n - f);
        F[f].mvuRight.assign(n - f, -1.f);
        const orbm_frame v = ORB_SLAM2::b200::Resident(F[f], f % 2 == 0);
        if (v.n != n - f || !v.kps || !v.desc || (f % 2 == 0) != (v.u_right != 0)) return -1;
    }
    const orbm_frame first = ORB_SLAM2::b200::Resident(F[0], true), last = ORB_SLAM2::b200::Resident(F[frames - 1], false);
    if (first.n != n || last.n != n - (frames - 1)) return -1;
    // a frame cached without right coordinates is uploaded again when they are asked for, and then serves both kinds of view
    const orbm_frame odd = ORB_SLAM2::b200::Resident(F[frames - 1], true), odd2 = ORB_SLAM2::b200::Resident(F[frames - 1], false);
    if (!odd.u_right || odd2.u_right) return -1;
    const int r = (int)(cache.Uploads() - before) * 100 + (int)cache.Size();
    cache.Clear();
    return r;
}

// both ORBmatcher::SearchByBoW overloads through b200::SearchByBoW; match12 [n1] = feature of side 2 or -1
int fwd_search_by_bow(int kf_kf, int n1, const Kp* k1, const unsigned char* d1, const unsigned char* has1, const unsigned char* bad1,
                      int nn1, const int* id1, const int* off1, const int* f1,
                      int n2, const Kp* k2, const unsigned char* d2, const unsigned char* has2, const unsigned char* bad2,
                      int nn2, const int* id2, const int* off2, const int* f2, float nnratio, int check_ori, int* match12)
{
    KeyFrameT A, B;
    std::vector<MapPointT> s1(n1 > 0 ? n1 : 1), s2(n2 > 0 ? n2 : 1);
    fill_keys(A.mvKeysUn, k1, n1); fill_desc(A.mDescriptors, d1, n1); fill_featvec(A.mFeatVec, nn1, id1, off1, f1);
    A.pts.assign(n1, static_cast<MapPointT*>(0));
    for (int i = 0; i < n1; ++i) if (has1[i]) { s1[i].bad = bad1[i] != 0; A.pts[i] = &s1[i]; }
    std::vector<MapPointT*> matches;
    int nm;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (kf_kf) {
        fill_keys(B.mvKeysUn, k2, n2); fill_desc(B.mDescriptors, d2, n2); fill_featvec(B.mFeatVec, nn2, id2, off2, f2);
        B.pts.assign(n2, static_cast<MapPointT*>(0));
        for (int i = 0; i < n2; ++i) if (has2[i]) { s2[i].bad = bad2[i] != 0; B.pts[i] = &s2[i]; }
        nm = ORB_SLAM2::b200::SearchByBoW(&A, &B, matches, nnratio, check_ori != 0);
        for (int i = 0; i < n1; ++i) if (matches[i]) match12[i] = (int)(matches[i] - &s2[0]);
    } else {
        FrameT F;
        F.N = n2;
        fill_keys(F.mvKeys, k2, n2); F.mvKeysUn = F.mvKeys; fill_desc(F.mDescriptors, d2, n2); fill_featvec(F.mFeatVec, nn2, id2, off2, f2);
        nm = ORB_SLAM2::b200::SearchByBoW(&A, F, matches, nnratio, check_ori != 0);
        for (int i = 0; i < n2; ++i) if (matches[i]) match12[(int)(matches[i] - &s1[0])] = i;   // vpMapPointMatches is per FRAME feature
    }
    return nm;
}

// the isInFrustum loop of Tracking::SearchLocalPoints through b200::IsInFrustum
int fwd_is_in_frustum(const float* Tcw, const float* K, float bf, float minX, float maxX, float minY, float maxY, float scale_factor,
                      int nlevels, float cos_limit, int n, const float* xyz, const float* normal, const float* max_d, const float* min_d,
                      unsigned char* in_view, float* proj, int* level, float* view_cos)
{
    struct MP : MapPointT {
        cv::Mat nrm; float mx, mn;
        cv::Mat GetNormal() { return nrm.clone(); }
        float GetMaxDistance() { return mx; }
        float GetMinDistance() { return mn; }
    };
    FrameT F;
    F.mTcw = cv::Mat(4, 4, CV_32F);
    for (int i = 0; i < 16; ++i) F.mTcw.at<float>(i / 4, i % 4) = Tcw[i];
    FrameT::fx = K[0]; FrameT::fy = K[1]; FrameT::cx = K[2]; FrameT::cy = K[3];
    FrameT::mnMinX = minX; FrameT::mnMaxX = maxX; FrameT::mnMinY = minY; FrameT::mnMaxY = maxY;
    F.mbf = bf; F.mfScaleFactor = scale_factor; F.mnScaleLevels = nlevels;
    std::vector<MP> store(n > 0 ? n : 1);
    std::vector<MP*> pts(n);
    for (int i = 0; i < n; ++i) {
        MP& m = store[i];
        m.pos = cv::Mat(3, 1, CV_32F); m.nrm = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) { m.pos.at<float>(k) = xyz[3 * i + k]; m.nrm.at<float>(k) = normal[3 * i + k]; }
        m.mx = max_d[i]; m.mn = min_d[i];
        m.mbTrackInView = true; m.mTrackProjX = m.mTrackProjY = m.mTrackProjXR = m.mTrackViewCos = 0.f; m.mnTrackScaleLevel = 0;
        pts[i] = &m;
    }
    std::vector<bool> iv;
    const int cnt = ORB_SLAM2::b200::IsInFrustum(F, pts, cos_limit, iv);
    for (int i = 0; i < n; ++i) {
        in_view[i] = (iv[i] && store[i].mbTrackInView) ? 1 : 0;
        if ((iv[i] ? 1 : 0) != (store[i].mbTrackInView ? 1 : 0)) return -1;
        proj[3 * i] = store[i].mTrackProjX; proj[3 * i + 1] = store[i].mTrackProjY; proj[3 * i + 2] = store[i].mTrackProjXR;
        level[i] = store[i].mnTrackScaleLevel; view_cos[i] = store[i].mTrackViewCos;
    }
    return cnt;
}

// the median search of MapPoint::ComputeDistinctiveDescriptors through b200::DistinctiveDescriptor
int fwd_distinctive(const unsigned char* desc, int n)
{
    std::vector<cv::Mat> v(n);
    for (int i = 0; i < n; ++i) fill_desc(v[i], desc + 32 * (size_t)i, 1);
    return ORB_SLAM2::b200::DistinctiveDescriptor(v);
}

// ORBmatcher::Fuse (both overloads) through b200::Fuse; the arrays and their meaning are those of oracle/ref_fuse_harness.cc
namespace {
cv::Mat mat_of(const float* v, int rows, int cols)
{
    cv::Mat m(rows, cols, CV_32F);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) m.at<float>(r, c) = v[r * cols + c];
    return m;
}
void fill_kf(KeyFrameT& KF, int n, const Kp* kps, const unsigned char* desc, const float* u_right, float minX, float maxX, float minY, float maxY,
             const float* scale, const float* inv_sigma2, int nlevels, const float* K, float bf)
{
    fill_keys(KF.mvKeysUn, kps, n); fill_desc(KF.mDescriptors, desc, n);
    KF.mvuRight.assign(n, -1.f);
    if (u_right) KF.mvuRight.assign(u_right, u_right + n);
    KF.mvScaleFactors.assign(scale, scale + nlevels);
    if (inv_sigma2) KF.mvInvLevelSigma2.assign(inv_sigma2, inv_sigma2 + nlevels);
    KF.fx = K[0]; KF.fy = K[1]; KF.cx = K[2]; KF.cy = K[3]; KF.mbf = bf;
    KF.mnMinX = (int)minX; KF.mnMaxX = (int)maxX; KF.mnMinY = (int)minY; KF.mnMaxY = (int)maxY;
    KF.pts.assign(n, static_cast<MapPointT*>(0));
}
void fill_points(std::vector<MapPointT>& mps, int npts, const unsigned char* bad, const float* xyz, const float* normal, const unsigned char* mp_desc,
                 const int* pred_level, const float* min_dist, const float* max_dist, const int* nobs)
{
    for (int i = 0; i < npts; ++i) {
        MapPointT& m = mps[i];
        m.bad = bad[i] != 0;
        m.pos = mat_of(xyz + 3 * i, 3, 1);
        if (normal) m.normalv = mat_of(normal + 3 * i, 3, 1);
        fill_desc(m.descriptor, mp_desc + 32 * (size_t)i, 1);
        m.mnTrackScaleLevel = pred_level[i];
        m.minD = min_dist[i]; m.maxD = max_dist[i]; m.haveRange = true;
        m.nObs = nobs ? nobs[i] : 0;
        m.replacedBy = 0;
    }
}
} // namespace

int fwd_fuse(int sim3, int n, const Kp* kps, const unsigned char* desc, const float* u_right,
             float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2, int nlevels,
             const float* K, float bf, const float* Rcw, const float* tcw, const float* Ow, const float* Scw,
             int npts, unsigned char* bad, const float* xyz, const float* normal, const unsigned char* mp_desc,
             const int* pred_level, const float* min_dist, const float* max_dist, int* nobs, int* kf_idx, int* replaced_by,
             int nlist, const int* list, int* kf_mp, int* replace_out, float th)
{
    KeyFrameT KF;
    fill_kf(KF, n, kps, desc, u_right, minX, maxX, minY, maxY, scale, inv_sigma2, nlevels, K, bf);
    if (!sim3) { KF.Rcw = mat_of(Rcw, 3, 3); KF.tcw = mat_of(tcw, 3, 1); KF.Ow = mat_of(Ow, 3, 1); }
    std::vector<MapPointT> mps(npts > 0 ? npts : 1);
    fill_points(mps, npts, bad, xyz, normal, mp_desc, pred_level, min_dist, max_dist, nobs);
    for (int i = 0; i < npts; ++i) if (kf_idx[i] >= 0) mps[i].obs[&KF] = (size_t)kf_idx[i];
    for (int k = 0; k < n; ++k) if (kf_mp[k] >= 0) KF.pts[k] = &mps[kf_mp[k]];
    std::vector<MapPointT*> pts(nlist);
    for (int i = 0; i < nlist; ++i) pts[i] = list[i] >= 0 ? &mps[list[i]] : static_cast<MapPointT*>(0);
    int nf;
    if (sim3) {
        std::vector<MapPointT*> rep(nlist, static_cast<MapPointT*>(0));
        nf = ORB_SLAM2::b200::Fuse(&KF, mat_of(Scw, 4, 4), pts, th, rep);
        for (int i = 0; i < nlist; ++i) replace_out[i] = rep[i] ? (int)(rep[i] - &mps[0]) : -1;
    } else {
        nf = ORB_SLAM2::b200::Fuse(&KF, pts, th);
    }
    for (int i = 0; i < npts; ++i) {
        bad[i] = mps[i].bad ? 1 : 0;
        nobs[i] = mps[i].nObs;
        kf_idx[i] = mps[i].GetIndexInKeyFrame(&KF);
        replaced_by[i] = mps[i].replacedBy ? (int)(mps[i].replacedBy - &mps[0]) : -1;
    }
    for (int k = 0; k < n; ++k) kf_mp[k] = KF.pts[k] ? (int)(KF.pts[k] - &mps[0]) : -1;
    ORB_SLAM2::b200::ResidentFrames::Local().Clear();
    return nf;
}

// ORBmatcher::SearchBySim3 through b200::SearchBySim3 (arrays as in oracle/ref_fuse_harness.cc)
int fwd_search_by_sim3(int n1, const Kp* kps1, const unsigned char* desc1, const int* mp1,
                       int n2, const Kp* kps2, const unsigned char* desc2, const int* mp2,
                       float minX, float maxX, float minY, float maxY, const float* scale, int nlevels, const float* K,
                       const float* R1w, const float* t1w, const float* R2w, const float* t2w, float s12, const float* R12, const float* t12,
                       int npts, const unsigned char* bad, const float* xyz, const unsigned char* mp_desc, const int* pred_level,
                       const float* min_dist, const float* max_dist, const int* idx_in_kf2, int* matches12, float th)
{
    KeyFrameT KF1, KF2;
    fill_kf(KF1, n1, kps1, desc1, 0, minX, maxX, minY, maxY, scale, 0, nlevels, K, 0.f);
    fill_kf(KF2, n2, kps2, desc2, 0, minX, maxX, minY, maxY, scale, 0, nlevels, K, 0.f);
    KF1.Rcw = mat_of(R1w, 3, 3); KF1.tcw = mat_of(t1w, 3, 1);
    KF2.Rcw = mat_of(R2w, 3, 3); KF2.tcw = mat_of(t2w, 3, 1);
    std::vector<MapPointT> mps(npts > 0 ? npts : 1);
    fill_points(mps, npts, bad, xyz, 0, mp_desc, pred_level, min_dist, max_dist, 0);
    for (int i = 0; i < npts; ++i) if (idx_in_kf2[i] >= 0) mps[i].obs[&KF2] = (size_t)idx_in_kf2[i];
    for (int k = 0; k < n1; ++k) if (mp1[k] >= 0) KF1.pts[k] = &mps[mp1[k]];
    for (int k = 0; k < n2; ++k) if (mp2[k] >= 0) KF2.pts[k] = &mps[mp2[k]];
    std::vector<MapPointT*> m12(n1, static_cast<MapPointT*>(0));
    for (int k = 0; k < n1; ++k) if (matches12[k] >= 0) m12[k] = &mps[matches12[k]];
    const int nf = ORB_SLAM2::b200::SearchBySim3(&KF1, &KF2, m12, s12, mat_of(R12, 3, 3), mat_of(t12, 3, 1), th);
    for (int k = 0; k < n1; ++k) matches12[k] = m12[k] ? (int)(m12[k] - &mps[0]) : -1;
    ORB_SLAM2::b200::ResidentFrames::Local().Clear();
    return nf;
}

// ORBmatcher::SearchForTriangulation through b200::SearchForTriangulation (arrays as in oracle/ref_bow_harness.cc)
int fwd_search_for_triangulation(int n1, const Kp* kps1, const unsigned char* desc1, const unsigned char* has_mp1, const float* u_right1,
                                 int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                                 int n2, const Kp* kps2, const unsigned char* desc2, const unsigned char* has_mp2, const float* u_right2,
                                 int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                                 const float* F12, const float* Cw, const float* pose2, const float* K2,
                                 const float* scale, const float* sigma2, int nlevels, int only_stereo, int check_ori, int* match12)
{
    KeyFrameT A, B;
    std::vector<MapPointT> s1(n1 > 0 ? n1 : 1), s2(n2 > 0 ? n2 : 1);
    fill_keys(A.mvKeysUn, kps1, n1); fill_desc(A.mDescriptors, desc1, n1); fill_featvec(A.mFeatVec, nn1, node_id1, node_off1, feat1);
    fill_keys(B.mvKeysUn, kps2, n2); fill_desc(B.mDescriptors, desc2, n2); fill_featvec(B.mFeatVec, nn2, node_id2, node_off2, feat2);
    A.pts.assign(n1, static_cast<MapPointT*>(0)); B.pts.assign(n2, static_cast<MapPointT*>(0));
    for (int i = 0; i < n1; ++i) if (has_mp1[i]) A.pts[i] = &s1[i];
    for (int i = 0; i < n2; ++i) if (has_mp2[i]) B.pts[i] = &s2[i];
    A.mvuRight.assign(n1, -1.f); B.mvuRight.assign(n2, -1.f);
    if (u_right1) A.mvuRight.assign(u_right1, u_right1 + n1);
    if (u_right2) B.mvuRight.assign(u_right2, u_right2 + n2);
    B.mvScaleFactors.assign(scale, scale + nlevels); B.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
    B.fx = K2[0]; B.fy = K2[1]; B.cx = K2[2]; B.cy = K2[3];
    A.Ow = mat_of(Cw, 3, 1); B.Rcw = mat_of(pose2, 3, 3); B.tcw = mat_of(pose2 + 9, 3, 1);
    std::vector<std::pair<size_t, size_t> > pairs;
    const int nm = ORB_SLAM2::b200::SearchForTriangulation(&A, &B, mat_of(F12, 3, 3), pairs, only_stereo != 0, check_ori != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (size_t k = 0; k < pairs.size(); ++k) match12[pairs[k].first] = (int)pairs[k].second;
    ORB_SLAM2::b200::ResidentFrames::Local().Clear();
    return nm;
}

} // extern "C"
#endif
