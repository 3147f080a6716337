// ORBmatcher_b200.hpp -- host forwarders for the Hamming searches of ORB_SLAM2::ORBmatcher.
//
// ORBmatcher's window searches take the SLAM system's own Frame / MapPoint objects, so they cannot be
// replaced by a standalone class.  Instead the bodies of the reference's methods become one-line calls
// into these templates (INTEGRATION.md shows the patch), which flatten the objects into the plain arrays
// of the C ABI (include/orb_b200.h), run the search on the GPU and write the results back:
//
//   int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
//   { return b200::SearchByProjection(F, vpMapPoints, th, mfNNratio); }                 // src/ORBmatcher.cc:73
//   int ORBmatcher::SearchByProjection(Frame &Cur, const Frame &Last, const float th, const bool bMono)
//   { return b200::SearchByProjection(Cur, Last, th, bMono, mbCheckOrientation); }      // src/ORBmatcher.cc:160
//   int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched,
//                                           vector<int> &vnMatches12, int windowSize)
//   { return b200::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize, mfNNratio,
//                                          mbCheckOrientation); }                       // src/ORBmatcher.cc:1055
//   int ORBmatcher::SearchByProjection(Frame &Cur, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist)
//   { return b200::SearchByProjection(Cur, pKF, sAlreadyFound, th, ORBdist, mbCheckOrientation); }   // src/ORBmatcher.cc:303
//   int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
//   { return b200::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th); }              // src/ORBmatcher.cc:434
//   int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
//   { return b200::SearchByBoW(pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation); }          // src/ORBmatcher.cc:552
//   int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)
//   { return b200::SearchByBoW(pKF1, pKF2, vpMatches12, mfNNratio, mbCheckOrientation); }            // src/ORBmatcher.cc:700
//   int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)
//   { return b200::Fuse(pKF, vpMapPoints, th); }                                         // src/ORBmatcher.cc:1364
//   int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)
//   { return b200::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint); }                       // src/ORBmatcher.cc:1516
//   int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12, const float &s12,
//                                const cv::Mat &R12, const cv::Mat &t12, const float th)
//   { return b200::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th); }           // src/ORBmatcher.cc:836
//   int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
//                                          vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)
//   { return b200::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo, mbCheckOrientation); }   // src/ORBmatcher.cc:1183
//   the isInFrustum loop of Tracking::SearchLocalPoints -> b200::IsInFrustum(F, vpPoints, 0.5f, inView)
//   the distance-matrix / median part of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:311-334)
//   -> BestIdx = b200::DistinctiveDescriptor(vDescriptors)
//   int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)   stays on the CPU for single
//   pairs (a 32-byte popcount); batches go through orbm_hamming_bf.
//
// The templates only use the members the reference's own code uses (listed per function), so they
// compile against the unmodified include/Frame.h and include/MapPoint.h.
#ifndef ORBMATCHER_B200_HPP
#define ORBMATCHER_B200_HPP

#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "orb_b200.h"

namespace ORB_SLAM2 { namespace b200 {

inline int& Device() { static int d = 0; return d; }

// &v[0] of an empty vector is undefined: a frame without keypoints (a black image) must reach the C ABI as n = 0, null
template <class T> inline T* Ptr(std::vector<T>& v) { return v.empty() ? (T*)0 : &v[0]; }
template <class T> inline const T* Ptr(const std::vector<T>& v) { return v.empty() ? (const T*)0 : &v[0]; }

template <class KeyPointT>
inline void FlattenKeys(const std::vector<KeyPointT>& keys, std::vector<orbx_kp>& out)
{
    out.resize(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
        orbx_kp k = { keys[i].pt.x, keys[i].pt.y, keys[i].size, keys[i].angle, keys[i].response, keys[i].octave, keys[i].class_id };
        out[i] = k;
    }
}

template <class MatT>
inline void FlattenDescriptors(const MatT& m, int n, std::vector<unsigned char>& out)
{
    out.resize((size_t)(n > 0 ? n : 1) * 32);
    for (int i = 0; i < n; ++i) std::memcpy(&out[(size_t)i * 32], m.ptr(i), 32);
}

inline void Check(int rc, const char* what)
{
    if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + " failed (B200 matcher, no CPU fallback)");
}

// ---- frames kept on the device --------------------------------------------------------------------------------
// Tracking runs two to four searches on the same Frame (TrackWithMotionModel / TrackReferenceKeyFrame, then
// SearchLocalPoints, Relocalization's retries), and a Frame's mvKeysUn, mDescriptors and mvuRight never change after its
// constructor (src/Frame.cc:58-174), nor do a KeyFrame's (they are copies of its Frame's, src/KeyFrame.cc:31-45).  So the
// forwarders upload them once (orbm_frame_upload) and every later search of that Frame or KeyFrame reuses the device copy:
// a small per-THREAD cache (Tracking, LocalMapping and LoopClosing run on their own threads) keyed by the frame id --
// Frame::mnId, which a KeyFrame carries as mnFrameId, so a KeyFrame finds the entry its Frame made -- with the keypoint
// count and the time stamp as a guard, least recently used entry dropped first.  Frame::nNextId restarts at 0 in
// Tracking::Reset (src/Tracking.cc:1636): call b200::ResidentFrames::Local().Clear() there.
class ResidentFrames {
public:
    enum { kCapacity = 8 };
    static ResidentFrames& Local() { static thread_local ResidentFrames c; return c; }
    ~ResidentFrames() { Clear(); }
    void Clear()
    {
        for (size_t i = 0; i < e_.size(); ++i) orbm_frame_release(e_[i].h);
        e_.clear();
    }
    size_t Size() const { return e_.size(); }
    unsigned long Uploads() const { return uploads_; }

    // The device-resident view of (keys, descriptors, uRight); uRight may be null.  Bounds are the caller's.
    template <class KeyPointT, class MatT>
    orbm_frame Get(unsigned long id, double stamp, const std::vector<KeyPointT>& keys, const MatT& descriptors, const std::vector<float>* uRight,
                   float minX, float maxX, float minY, float maxY)
    {
        const int n = (int)keys.size();
        const bool wantRight = uRight && !uRight->empty();
        Entry* hit = 0;
        for (size_t i = 0; i < e_.size(); ++i)
            if (e_[i].id == id && e_[i].n == n && e_[i].stamp == stamp && e_[i].device == Device() && (e_[i].right || !wantRight)) { hit = &e_[i]; break; }
        if (!hit) {
            std::vector<orbx_kp> kps; FlattenKeys(keys, kps);
            std::vector<unsigned char> desc; FlattenDescriptors(descriptors, n, desc);
            orbm_frame host;
            host.n = n; host.kps = Ptr(kps); host.desc = &desc[0]; host.u_right = wantRight ? &(*uRight)[0] : 0;
            host.min_x = minX; host.max_x = maxX; host.min_y = minY; host.max_y = maxY;
            Entry e;
            e.id = id; e.n = n; e.stamp = stamp; e.device = Device(); e.right = wantRight; e.use = 0; e.h = 0;
            Check(orbm_frame_upload(&host, Device(), &e.h), "orbm_frame_upload");
            ++uploads_;
            // an entry of the same frame without right coordinates is superseded; else the least recently used one goes
            size_t victim = e_.size();
            for (size_t i = 0; i < e_.size(); ++i) if (e_[i].id == id && e_[i].n == n && e_[i].stamp == stamp && e_[i].device == e.device) victim = i;
            if (victim == e_.size() && e_.size() >= (size_t)kCapacity) {
                victim = 0;
                for (size_t i = 1; i < e_.size(); ++i) if (e_[i].use < e_[victim].use) victim = i;
            }
            if (victim < e_.size()) { orbm_frame_release(e_[victim].h); e_[victim] = e; hit = &e_[victim]; }
            else { e_.push_back(e); hit = &e_.back(); }
        }
        hit->use = ++tick_;
        orbm_frame v;
        Check(orbm_frame_view(hit->h, &v), "orbm_frame_view");
        if (!wantRight) v.u_right = 0;
        v.min_x = minX; v.max_x = maxX; v.min_y = minY; v.max_y = maxY;
        return v;
    }

private:
    struct Entry { unsigned long id; int n; double stamp; int device; bool right; unsigned long use; orbm_frame_handle* h; };
    ResidentFrames() : tick_(0), uploads_(0) {}
    ResidentFrames(const ResidentFrames&);
    ResidentFrames& operator=(const ResidentFrames&);
    std::vector<Entry> e_;
    unsigned long tick_, uploads_;
};

// A Frame's mvKeysUn / mDescriptors (/ mvuRight when withRight) on the device.  Reads F.mnId, F.mTimeStamp.
template <class FrameT>
inline orbm_frame Resident(const FrameT& F, bool withRight)
{
    return ResidentFrames::Local().Get((unsigned long)F.mnId, (double)F.mTimeStamp, F.mvKeysUn, F.mDescriptors, withRight ? &F.mvuRight : 0,
                                       FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY);
}
// A KeyFrame's mvKeysUn / mDescriptors on the device (the entry of the Frame it was made from, if still cached).  Reads
// pKF->mnFrameId, mTimeStamp.  `grid`: with the key frame's own image bounds (the window searches), else the 0..1 dummy
// bounds of the searches that use no grid (SearchByBoW).
template <class KeyFrameT>
inline orbm_frame ResidentKF(const KeyFrameT* pKF, bool grid)
{
    if (grid)
        return ResidentFrames::Local().Get((unsigned long)pKF->mnFrameId, (double)pKF->mTimeStamp, pKF->mvKeysUn, pKF->mDescriptors, 0,
                                           (float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY);
    return ResidentFrames::Local().Get((unsigned long)pKF->mnFrameId, (double)pKF->mTimeStamp, pKF->mvKeysUn, pKF->mDescriptors, 0, 0.f, 1.f, 0.f, 1.f);
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:73-157.
// Reads: F.mvKeysUn, F.mDescriptors, F.mvuRight, F.mvScaleFactors, F.mvpMapPoints, Frame::mnMin/Max*;
// MapPoint::mbTrackInView, isBad(), mnTrackScaleLevel, mTrackViewCos, mTrackProjX/Y/XR, Observations(), GetDescriptor().
template <class FrameT, class MapPointT>
int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float th, const float nnratio)
{
    const orbm_frame view = Resident(F, true);
    const size_t nk = F.mvKeysUn.size();
    // map points already attached to the frame but not in the list still block their keypoint when
    // Observations()>0 (:115-117): they ride along as extra, never-queried entries
    std::vector<MapPointT*> pts(vpMapPoints.begin(), vpMapPoints.end());
    const int nq = (int)pts.size();
    std::vector<int> init(nk, -1);
    std::unordered_map<MapPointT*, int> index;                      // built only when some keypoint has a point attached
    for (size_t k = 0; k < nk; ++k) {
        MapPointT* p = F.mvpMapPoints[k];
        if (!p) continue;
        if (index.empty()) {
            index.reserve(2 * pts.size() + nk);
            for (size_t i = 0; i < pts.size(); ++i) index.insert(std::make_pair(pts[i], (int)i));   // first occurrence wins, as with std::map
        }
        typename std::unordered_map<MapPointT*, int>::iterator it = index.find(p);
        if (it == index.end()) { it = index.insert(std::make_pair(p, (int)pts.size())).first; pts.push_back(p); }
        init[k] = it->second;
    }
    const size_t n = pts.size();
    std::vector<float> proj(3 * n, 0.f), viewCos(n, 0.f);
    std::vector<int> level(n, 0), obs(n, 0);
    std::vector<unsigned char> inView(n, 0), bad(n, 0), qdesc(32 * (n ? n : 1), 0);
    for (size_t i = 0; i < n; ++i) {
        MapPointT* p = pts[i];
        obs[i] = p->Observations();
        if ((int)i >= nq) continue;                       // ride-along entry: only its Observations() matter
        inView[i] = p->mbTrackInView ? 1 : 0;
        bad[i] = p->isBad() ? 1 : 0;
        level[i] = p->mnTrackScaleLevel; viewCos[i] = p->mTrackViewCos;
        proj[3 * i] = p->mTrackProjX; proj[3 * i + 1] = p->mTrackProjY; proj[3 * i + 2] = p->mTrackProjXR;
        if (inView[i] && !bad[i]) std::memcpy(&qdesc[32 * i], p->GetDescriptor().ptr(0), 32);
    }
    std::vector<int> assign(nk, -1);
    int nmatches = 0;
    Check(orbm_search_by_projection_points(&view, &F.mvScaleFactors[0], (int)F.mvScaleFactors.size(), (int)n, Ptr(proj), Ptr(level), Ptr(viewCos),
                                           Ptr(inView), Ptr(bad), Ptr(obs), &qdesc[0], Ptr(init), Ptr(assign), th, nnratio,
                                           &nmatches, Device()), "orbm_search_by_projection_points");
    for (size_t k = 0; k < nk; ++k) if (assign[k] >= 0) F.mvpMapPoints[k] = pts[(size_t)assign[k]];
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:160-300.
// Reads: cur.mvKeysUn, mDescriptors, mvuRight, mvScaleFactors, mvpMapPoints, mTcw, mbf, Frame::fx/fy/cx/cy/mnMin/Max*;
// last.N, mvKeys, mvKeysUn, mvpMapPoints, mvbOutlier, mTcw; MapPoint::GetWorldPos(), GetDescriptor(), Observations().
template <class FrameT>
int SearchByProjection(FrameT& Cur, const FrameT& Last, const float th, const bool bMono, const bool checkOri)
{
    const orbm_frame view = Resident(Cur, true);
    const size_t nk = Cur.mvKeysUn.size();
    const int nl = Last.N;
    std::vector<orbx_kp> lk; FlattenKeys(Last.mvKeysUn, lk);
    for (int i = 0; i < nl; ++i) lk[(size_t)i].octave = Last.mvKeys[(size_t)i].octave;   // :211 uses mvKeys for the octave
    std::vector<unsigned char> hasMp((size_t)(nl ? nl : 1), 0), outlier((size_t)(nl ? nl : 1), 0), mdesc((size_t)(nl ? nl : 1) * 32, 0);
    std::vector<float> xyz((size_t)(nl ? nl : 1) * 3, 0.f);
    std::vector<int> mobs((size_t)(nl ? nl : 1), 0);
    for (int i = 0; i < nl; ++i) {
        if (!Last.mvpMapPoints[(size_t)i]) continue;
        hasMp[(size_t)i] = 1; outlier[(size_t)i] = Last.mvbOutlier[(size_t)i] ? 1 : 0;
        const cv::Mat x = Last.mvpMapPoints[(size_t)i]->GetWorldPos();
        for (int r = 0; r < 3; ++r) xyz[(size_t)(3 * i + r)] = x.template at<float>(r);
        std::memcpy(&mdesc[(size_t)i * 32], Last.mvpMapPoints[(size_t)i]->GetDescriptor().ptr(0), 32);
        mobs[(size_t)i] = Last.mvpMapPoints[(size_t)i]->Observations();
    }
    float Tc[16], Tl[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { Tc[4 * r + c] = Cur.mTcw.template at<float>(r, c); Tl[4 * r + c] = Last.mTcw.template at<float>(r, c); }
    const float K[4] = { FrameT::fx, FrameT::fy, FrameT::cx, FrameT::cy };
    std::vector<int> initObs(nk, -1), assign(nk, -1);
    for (size_t k = 0; k < nk; ++k) if (Cur.mvpMapPoints[k]) initObs[k] = Cur.mvpMapPoints[k]->Observations();
    int nmatches = 0;
    Check(orbm_search_by_projection_frame(&view, nl, lk.empty() ? 0 : &lk[0], &hasMp[0], &outlier[0], &xyz[0], &mdesc[0], &mobs[0], Tc, Tl, K, Cur.mbf,
                                          &Cur.mvScaleFactors[0], (int)Cur.mvScaleFactors.size(), Ptr(initObs),
                                          Ptr(assign), th, bMono ? 1 : 0, checkOri ? 1 : 0, &nmatches, Device()),
          "orbm_search_by_projection_frame");
    for (size_t k = 0; k < nk; ++k) {
        if (assign[k] >= 0) Cur.mvpMapPoints[k] = Last.mvpMapPoints[(size_t)assign[k]];
        else if (assign[k] == -1) Cur.mvpMapPoints[k] = 0;
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& cur, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
// src/ORBmatcher.cc:303-431 (relocalisation).  The projection, the distance gate and MapPoint::PredictScale
// stay on the host in the reference's arithmetic (cv::gemm float accumulation, cv::norm in double);
// windows, distances, ordered claims and the rotation histogram run on the GPU (orbm_window_search_best).
template <class FrameT, class KeyFrameT, class MapPointT>
int SearchByProjection(FrameT& Cur, KeyFrameT* pKF, const std::set<MapPointT*>& sAlreadyFound, const float th, const int ORBdist,
                       const bool checkOri)
{
    const orbm_frame view = Resident(Cur, false);
    const size_t nk = Cur.mvKeysUn.size();
    float T[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = Cur.mTcw.template at<float>(r, c);
    float Ow[3];
    for (int i = 0; i < 3; ++i) {
        float s = (-T[0 * 4 + i]) * T[3];
        s = s + (-T[1 * 4 + i]) * T[7];
        s = s + (-T[2 * 4 + i]) * T[11];
        Ow[i] = s;
    }
    const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();
    const size_t nq = vpMPs.size();
    std::vector<float> uvr(3 * (nq ? nq : 1), 0.f), qangle(nq ? nq : 1, 0.f);
    std::vector<int> minl(nq ? nq : 1, 0), maxl(nq ? nq : 1, 0);
    std::vector<unsigned char> valid(nq ? nq : 1, 0), qdesc(32 * (nq ? nq : 1), 0);
    for (size_t i = 0; i < nq; ++i) {
        MapPointT* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const float x[3] = { x3Dw.template at<float>(0), x3Dw.template at<float>(1), x3Dw.template at<float>(2) };
        float pc[3];
        for (int r = 0; r < 3; ++r) {
            float s = T[4 * r] * x[0];
            s = s + T[4 * r + 1] * x[1];
            s = s + T[4 * r + 2] * x[2];
            pc[r] = s + T[4 * r + 3];
        }
        const float invzc = 1.0 / pc[2];
        const float u = FrameT::fx * pc[0] * invzc + FrameT::cx, v = FrameT::fy * pc[1] * invzc + FrameT::cy;
        if (u < FrameT::mnMinX || u > FrameT::mnMaxX) continue;
        if (v < FrameT::mnMinY || v > FrameT::mnMaxY) continue;
        double acc = 0;
        for (int r = 0; r < 3; ++r) { const float d = x[r] - Ow[r]; acc += (double)d * (double)d; }
        float dist3D = (float)std::sqrt(acc);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        const int lvl = pMP->PredictScale(dist3D, &Cur);
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = th * Cur.mvScaleFactors[(size_t)lvl];
        minl[i] = lvl - 1; maxl[i] = lvl + 1; valid[i] = 1;
        qangle[i] = pKF->mvKeysUn[i].angle;
        std::memcpy(&qdesc[32 * i], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<int> initObs(nk, -1), assign(nk, -1);
    for (size_t k = 0; k < nk; ++k) if (Cur.mvpMapPoints[k]) initObs[k] = 1;   // any attached point blocks (:373-374)
    int nmatches = 0;
    Check(orbm_window_search_best(&view, (int)nq, &uvr[0], &minl[0], &maxl[0], 0, 0, &valid[0], &qdesc[0], &qangle[0], 0,
                                  Ptr(initObs), Ptr(assign), ORBdist, checkOri ? 1 : 0, &nmatches, Device()),
          "orbm_window_search_best");
    for (size_t k = 0; k < nk; ++k) {
        if (assign[k] >= 0) Cur.mvpMapPoints[k] = vpMPs[(size_t)assign[k]];
        else if (assign[k] == -1) Cur.mvpMapPoints[k] = 0;
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th),
// src/ORBmatcher.cc:434-549 (loop closing).  The Sim3 decomposition, projection, distance / viewing-angle gates and
// MapPoint::PredictScale stay on the host in the reference's arithmetic (cv::Mat::dot and cv::norm accumulate in double,
// Mat / scalar multiplies by the double reciprocal, matrix products accumulate in float); windows over the key frame's grid,
// distances and the ordered "keypoint already matched" rule run on the GPU (orbm_window_search_best, TH_LOW, no
// orientation check).  Reads pKF->fx, fy, cx, cy, mnMinX..mnMaxY, mvKeysUn, mDescriptors, mvScaleFactors.
// (The key frame's grid origin is the int-truncated bound, include/KeyFrame.h:186-189; results are exact whenever the
// image bounds are integers, i.e. for undistorted / rectified input.)
template <class KeyFrameT, class MatT, class MapPointT>
int SearchByProjection(KeyFrameT* pKF, const MatT& Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th)
{
    const int TH_LOW = 50;                                                               // src/ORBmatcher.cc:38
    const orbm_frame view = ResidentKF(pKF, true);
    const size_t nk = pKF->mvKeysUn.size();
    const float fx = pKF->fx, fy = pKF->fy, cx = pKF->cx, cy = pKF->cy;
    float S[12];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) S[4 * r + c] = Scw.template at<float>(r, c);
    double d = 0;
    for (int c = 0; c < 3; ++c) d += (double)S[c] * (double)S[c];
    const float scw = (float)std::sqrt(d);
    const double inv = 1.0 / (double)scw;
    float R[9], t[3], Ow[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (float)((double)S[4 * r + c] * inv);
        t[r] = (float)((double)S[4 * r + 3] * inv);
    }
    for (int i = 0; i < 3; ++i) {
        float s = (float)((double)R[i] * -1.0) * t[0];
        s = s + (float)((double)R[3 + i] * -1.0) * t[1];
        s = s + (float)((double)R[6 + i] * -1.0) * t[2];
        Ow[i] = s;
    }
    const std::set<MapPointT*> found(vpMatched.begin(), vpMatched.end());                // :449-450 (NULL is never a candidate)
    const size_t nq = vpPoints.size();
    std::vector<float> uvr(3 * (nq ? nq : 1), 0.f);
    std::vector<int> minl(nq ? nq : 1, 0), maxl(nq ? nq : 1, 0);
    std::vector<unsigned char> valid(nq ? nq : 1, 0), qdesc(32 * (nq ? nq : 1), 0);
    for (size_t i = 0; i < nq; ++i) {
        MapPointT* pMP = vpPoints[i];
        if (pMP->isBad() || found.count(pMP)) continue;
        const MatT p3Dw = pMP->GetWorldPos();
        const float x[3] = { p3Dw.template at<float>(0), p3Dw.template at<float>(1), p3Dw.template at<float>(2) };
        float pc[3];
        for (int r = 0; r < 3; ++r) {
            float s = R[3 * r] * x[0];
            s = s + R[3 * r + 1] * x[1];
            s = s + R[3 * r + 2] * x[2];
            pc[r] = s + t[r];
        }
        if (pc[2] < 0.0) continue;
        const float invz = 1 / pc[2];
        const float xn = pc[0] * invz, yn = pc[1] * invz;
        const float u = fx * xn + cx, v = fy * yn + cy;
        if (!(u >= pKF->mnMinX && u < pKF->mnMaxX && v >= pKF->mnMinY && v < pKF->mnMaxY)) continue;
        float PO[3];
        double acc = 0;
        for (int r = 0; r < 3; ++r) { PO[r] = x[r] - Ow[r]; acc += (double)PO[r] * (double)PO[r]; }
        const float dist = (float)std::sqrt(acc);
        if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
        const MatT Pn = pMP->GetNormal();
        double dn = 0;
        for (int r = 0; r < 3; ++r) dn += (double)PO[r] * (double)Pn.template at<float>(r);
        if (dn < 0.5 * dist) continue;
        const int lvl = pMP->PredictScale(dist, pKF);
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = th * pKF->mvScaleFactors[(size_t)lvl];
        minl[i] = lvl - 1; maxl[i] = lvl; valid[i] = 1;
        std::memcpy(&qdesc[32 * i], pMP->GetDescriptor().ptr(0), 32);
    }
    std::vector<int> initObs(nk, -1), assign(nk, -1);
    for (size_t k = 0; k < nk; ++k) if (vpMatched[k]) initObs[k] = 1;
    int nmatches = 0;
    Check(orbm_window_search_best(&view, (int)nq, &uvr[0], &minl[0], &maxl[0], 0, 0, &valid[0], &qdesc[0], 0, 0,
                                  Ptr(initObs), Ptr(assign), TH_LOW, 0, &nmatches, Device()),
          "orbm_window_search_best");
    for (size_t k = 0; k < nk; ++k) if (assign[k] >= 0) vpMatched[k] = vpPoints[(size_t)assign[k]];
    return nmatches;
}

// ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:1055-1180.
template <class FrameT, class Point2fT>
int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<Point2fT>& vbPrevMatched, std::vector<int>& vnMatches12,
                            int windowSize, const float nnratio, const bool checkOri)
{
    const orbm_frame v1 = Resident(F1, false), v2 = Resident(F2, false);
    const size_t n1 = F1.mvKeysUn.size();
    std::vector<float> prev(2 * (n1 ? n1 : 1), 0.f);
    for (size_t i = 0; i < n1; ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    vnMatches12.assign(n1, -1);
    int nmatches = 0;
    std::vector<int> m12(n1 ? n1 : 1, -1);
    Check(orbm_search_for_initialization(&v1, &v2, &prev[0], &m12[0], windowSize, nnratio, checkOri ? 1 : 0, &nmatches, Device()),
          "orbm_search_for_initialization");
    for (size_t i = 0; i < n1; ++i) { vnMatches12[i] = m12[i]; vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    return nmatches;
}

// A DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>) as the CSR arrays of orbm_featvec.
template <class FeatVecT>
inline void FlattenFeatureVector(const FeatVecT& fv, std::vector<int>& node_id, std::vector<int>& node_off, std::vector<int>& feat)
{
    node_id.clear(); node_off.assign(1, 0); feat.clear();
    for (typename FeatVecT::const_iterator it = fv.begin(); it != fv.end(); ++it) {
        node_id.push_back((int)it->first);
        for (size_t j = 0; j < it->second.size(); ++j) feat.push_back((int)it->second[j]);
        node_off.push_back((int)feat.size());
    }
}

template <class MapPointT>
inline void ValidPoints(const std::vector<MapPointT*>& pts, size_t n, std::vector<unsigned char>& valid)
{
    valid.assign(n ? n : 1, 0);
    for (size_t i = 0; i < n && i < pts.size(); ++i) valid[i] = (pts[i] && !pts[i]->isBad()) ? 1 : 0;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:552-697.
// Reads: pKF->GetMapPointMatches(), mFeatVec, mDescriptors, mvKeysUn; F.N, mFeatVec, mDescriptors, mvKeys; MapPoint::isBad().
template <class KeyFrameT, class FrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches, const float nnratio, const bool checkOri)
{
    const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPointT*>((size_t)F.N, static_cast<MapPointT*>(0));
    // (the reference reads F.mvKeys here, :655, but only its angle, which UndistortKeyPoints leaves as it is, src/Frame.cc:425-466)
    const orbm_frame va = ResidentKF(pKF, false);
    orbm_frame vb = Resident(F, false);
    vb.min_x = 0.f; vb.max_x = 1.f; vb.min_y = 0.f; vb.max_y = 1.f;
    const size_t na = pKF->mvKeysUn.size();
    std::vector<int> ia, oa, fa, ib, ob, fb;
    FlattenFeatureVector(pKF->mFeatVec, ia, oa, fa); FlattenFeatureVector(F.mFeatVec, ib, ob, fb);
    std::vector<unsigned char> valid; ValidPoints(vpMapPointsKF, na, valid);
    std::vector<int> m12(na ? na : 1, -1);
    int nmatches = 0;
    Check(orbm_search_by_bow(&va, &valid[0], (int)ia.size(), ia.empty() ? 0 : &ia[0], &oa[0], fa.empty() ? 0 : &fa[0],
                             &vb, 0, (int)ib.size(), ib.empty() ? 0 : &ib[0], &ob[0], fb.empty() ? 0 : &fb[0],
                             0, nnratio, checkOri ? 1 : 0, &m12[0], &nmatches, Device()), "orbm_search_by_bow");
    for (size_t i = 0; i < na; ++i) if (m12[i] >= 0) vpMapPointMatches[(size_t)m12[i]] = vpMapPointsKF[i];
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), src/ORBmatcher.cc:700-832.
template <class KeyFrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float nnratio, const bool checkOri)
{
    const std::vector<MapPointT*> p1 = pKF1->GetMapPointMatches(), p2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPointT*>(p1.size(), static_cast<MapPointT*>(0));
    const orbm_frame va = ResidentKF(pKF1, false), vb = ResidentKF(pKF2, false);
    const size_t na = pKF1->mvKeysUn.size(), nb = pKF2->mvKeysUn.size();
    std::vector<int> ia, oa, fa, ib, ob, fb;
    FlattenFeatureVector(pKF1->mFeatVec, ia, oa, fa); FlattenFeatureVector(pKF2->mFeatVec, ib, ob, fb);
    std::vector<unsigned char> v1, v2; ValidPoints(p1, na, v1); ValidPoints(p2, nb, v2);
    std::vector<int> m12(na ? na : 1, -1);
    int nmatches = 0;
    Check(orbm_search_by_bow(&va, &v1[0], (int)ia.size(), ia.empty() ? 0 : &ia[0], &oa[0], fa.empty() ? 0 : &fa[0],
                             &vb, &v2[0], (int)ib.size(), ib.empty() ? 0 : &ib[0], &ob[0], fb.empty() ? 0 : &fb[0],
                             1, nnratio, checkOri ? 1 : 0, &m12[0], &nmatches, Device()), "orbm_search_by_bow");
    for (size_t i = 0; i < na && i < p1.size(); ++i) if (m12[i] >= 0) vpMatches12[i] = p2[(size_t)m12[i]];
    return nmatches;
}

// The loop `if (mCurrentFrame.isInFrustum(pMP, 0.5)) ...` of Tracking::SearchLocalPoints: Frame::isInFrustum
// (src/Frame.cc:288-345) for all points at once.  Sets mbTrackInView and, for points in view, mTrackProjX/Y/XR,
// mnTrackScaleLevel, mTrackViewCos exactly as the reference does; inView[i] is isInFrustum's return value.
// MapPoint needs two one-line accessors beside Get{Min,Max}DistanceInvariance:
//   float GetMaxDistance() { unique_lock<mutex> lock(mMutexPos); return mfMaxDistance; }   and GetMinDistance() likewise,
// because PredictScale (src/MapPoint.cc:459-475) divides the RAW mfMaxDistance.
template <class FrameT, class MapPointT>
int IsInFrustum(FrameT& F, const std::vector<MapPointT*>& pts, const float viewingCosLimit, std::vector<bool>& inView)
{
    const size_t n = pts.size();
    inView.assign(n, false);
    if (!n) return 0;
    std::vector<float> xyz(3 * n), nrm(3 * n), mx(n), mn(n), proj(3 * n, 0.f), vc(n, 0.f);
    std::vector<int> lvl(n, 0);
    std::vector<unsigned char> iv(n, 0);
    for (size_t i = 0; i < n; ++i) {
        pts[i]->mbTrackInView = false;
        const typename std::remove_reference<decltype(pts[i]->GetWorldPos())>::type P = pts[i]->GetWorldPos(), N = pts[i]->GetNormal();
        for (int k = 0; k < 3; ++k) { xyz[3 * i + k] = P.template at<float>(k); nrm[3 * i + k] = N.template at<float>(k); }
        mx[i] = pts[i]->GetMaxDistance(); mn[i] = pts[i]->GetMinDistance();
    }
    float T[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = F.mTcw.template at<float>(r, c);
    const float K[4] = { FrameT::fx, FrameT::fy, FrameT::cx, FrameT::cy };
    Check(orbm_project_points(T, K, F.mbf, FrameT::mnMinX, FrameT::mnMaxX, FrameT::mnMinY, FrameT::mnMaxY, F.mfScaleFactor, F.mnScaleLevels,
                              viewingCosLimit, (int)n, &xyz[0], &nrm[0], &mx[0], &mn[0], &iv[0], &proj[0], &lvl[0], &vc[0], Device()),
          "orbm_project_points");
    int count = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!iv[i]) continue;
        inView[i] = true; ++count;
        pts[i]->mbTrackInView = true;
        pts[i]->mTrackProjX = proj[3 * i]; pts[i]->mTrackProjY = proj[3 * i + 1]; pts[i]->mTrackProjXR = proj[3 * i + 2];
        pts[i]->mnTrackScaleLevel = lvl[i]; pts[i]->mTrackViewCos = vc[i];
    }
    return count;
}

// The distance matrix and median search of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:311-334):
// returns BestIdx for the descriptors the reference collected in vDescriptors (-1 when empty).
template <class MatT>
int DistinctiveDescriptor(const std::vector<MatT>& vDescriptors)
{
    const size_t n = vDescriptors.size();
    if (!n) return -1;
    std::vector<unsigned char> d(32 * n);
    for (size_t i = 0; i < n; ++i) std::memcpy(&d[32 * i], vDescriptors[i].ptr(0), 32);
    int best = -1;
    Check(orbm_distinctive_descriptor(&d[0], (int)n, 0, &best, 0, Device()), "orbm_distinctive_descriptor");
    return best;
}

// ---- Fuse and SearchBySim3: searches whose queries do not depend on one another --------------------------------------
// What the reference does per map point is (1) project, (2) find the closest descriptor in the window at the predicted level
// or the one below, (3) update the map (Replace / AddObservation / AddMapPoint, or the mutual check of SearchBySim3).  Only
// (3) reads state that earlier points changed, and it never feeds back into (2), so the forwarders project every point on the
// host in the reference's arithmetic (matrix products accumulate in float, cv::norm and cv::Mat::dot in double), search all of
// them in one device call (orbm_window_best_free) and then replay (3) in the reference's order with its tests re-evaluated
// on the changing state.
namespace detail {
inline void MulAdd3(const float* R, const float* t, const float* x, float* out)      // R*x + t as cv::Mat computes it
{
    for (int r = 0; r < 3; ++r) {
        float s = R[3 * r] * x[0];
        s = s + R[3 * r + 1] * x[1];
        s = s + R[3 * r + 2] * x[2];
        out[r] = s + t[r];
    }
}
template <class MatT> inline void Read3x3(const MatT& m, float* R) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[3 * r + c] = m.template at<float>(r, c); }
template <class MatT> inline void Read3(const MatT& m, float* t) { for (int r = 0; r < 3; ++r) t[r] = m.template at<float>(r); }

struct FreeQueries {
    std::vector<float> uvr, ur;
    std::vector<int> level;
    std::vector<unsigned char> valid, desc;
    explicit FreeQueries(size_t n) : uvr(3 * (n ? n : 1), 0.f), ur(n ? n : 1, 0.f), level(n ? n : 1, 0), valid(n ? n : 1, 0), desc(32 * (n ? n : 1), 0) {}
};

// Projection + gates of one candidate point of Fuse (src/ORBmatcher.cc:1388-1425 / :1546-1583); doubleInvZ: the Sim3 overload
// divides 1.0 (double) by z, the pose overload 1 (int -> float).  Fills query i and returns true when the point reaches the search.
template <class KeyFrameT, class MapPointT>
inline bool FuseQuery(KeyFrameT* pKF, MapPointT* pMP, const float* R, const float* t, const float* Ow, float bf, float th, bool doubleInvZ,
                      FreeQueries& Q, size_t i)
{
    float x[3], pc[3];
    Read3(pMP->GetWorldPos(), x);
    MulAdd3(R, t, x, pc);
    if (pc[2] < 0.0f) return false;
    const float invz = doubleInvZ ? (float)(1.0 / (double)pc[2]) : 1 / pc[2];
    const float xn = pc[0] * invz, yn = pc[1] * invz;
    const float u = pKF->fx * xn + pKF->cx, v = pKF->fy * yn + pKF->cy;
    if (!(u >= pKF->mnMinX && u < pKF->mnMaxX && v >= pKF->mnMinY && v < pKF->mnMaxY)) return false;     // KeyFrame::IsInImage
    float PO[3];
    double acc = 0;
    for (int r = 0; r < 3; ++r) { PO[r] = x[r] - Ow[r]; acc += (double)PO[r] * (double)PO[r]; }
    const float dist3D = (float)std::sqrt(acc);
    if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) return false;
    float Pn[3];
    Read3(pMP->GetNormal(), Pn);
    double dn = 0;
    for (int r = 0; r < 3; ++r) dn += (double)PO[r] * (double)Pn[r];
    if (dn < 0.5 * dist3D) return false;
    const int lvl = pMP->PredictScale(dist3D, pKF);
    Q.uvr[3 * i] = u; Q.uvr[3 * i + 1] = v; Q.uvr[3 * i + 2] = th * pKF->mvScaleFactors[(size_t)lvl];
    Q.ur[i] = u - bf * invz;
    Q.level[i] = lvl; Q.valid[i] = 1;
    std::memcpy(&Q.desc[32 * i], pMP->GetDescriptor().ptr(0), 32);
    return true;
}
} // namespace detail

// ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th), src/ORBmatcher.cc:1364-1513.
// Reads pKF->GetRotation(), GetTranslation(), GetCameraCenter(), fx, fy, cx, cy, mbf, mnMinX..mnMaxY, mvKeysUn, mvuRight,
// mDescriptors, mvScaleFactors, mvInvLevelSigma2, GetMapPoint(); MapPoint::isBad(), IsInKeyFrame(), GetWorldPos(),
// GetNormal(), Get{Min,Max}DistanceInvariance(), PredictScale(), GetDescriptor(), Observations(); writes through
// MapPoint::Replace(), AddObservation() and KeyFrame::AddMapPoint() exactly where the reference does.
template <class KeyFrameT, class MapPointT>
int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, const float th)
{
    const int TH_LOW = 50;
    float R[9], t[3], Ow[3];
    detail::Read3x3(pKF->GetRotation(), R); detail::Read3(pKF->GetTranslation(), t); detail::Read3(pKF->GetCameraCenter(), Ow);
    const size_t nq = vpMapPoints.size();
    detail::FreeQueries Q(nq);
    for (size_t i = 0; i < nq; ++i) {
        MapPointT* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                   // :1383-1386 (re-tested below)
        detail::FuseQuery(pKF, pMP, R, t, Ow, pKF->mbf, th, false, Q, i);
    }
    const orbm_frame view = ResidentFrames::Local().Get((unsigned long)pKF->mnFrameId, (double)pKF->mTimeStamp, pKF->mvKeysUn, pKF->mDescriptors,
                                                        &pKF->mvuRight, (float)pKF->mnMinX, (float)pKF->mnMaxX, (float)pKF->mnMinY, (float)pKF->mnMaxY);
    std::vector<int> best(nq ? nq : 1, -1), dist(nq ? nq : 1, 256);
    int found = 0;
    Check(orbm_window_best_free(&view, (int)nq, &Q.uvr[0], &Q.level[0], &Q.ur[0], &Q.valid[0], &Q.desc[0], Ptr(pKF->mvInvLevelSigma2),
                                (int)pKF->mvInvLevelSigma2.size(), TH_LOW, &best[0], &dist[0], &found, Device()), "orbm_window_best_free");
    int nFused = 0;
    for (size_t i = 0; i < nq; ++i) {
        MapPointT* pMP = vpMapPoints[i];
        if (!pMP || !Q.valid[i] || best[i] < 0) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                           // state earlier points may have changed
        MapPointT* pMPinKF = pKF->GetMapPoint((size_t)best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {                                                    // :1489-1495
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, (size_t)best[i]);                                  // :1500-1501
            pKF->AddMapPoint(pMP, (size_t)best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint),
// src/ORBmatcher.cc:1516-1633 (loop closing): pose from the Sim3, no chi-square test, points the key frame holds at the start
// are skipped, a keypoint that holds a point already is reported in vpReplacePoint instead of replaced.
template <class KeyFrameT, class MatT, class MapPointT>
int Fuse(KeyFrameT* pKF, const MatT& Scw, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint)
{
    const int TH_LOW = 50;
    float S[12];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) S[4 * r + c] = Scw.template at<float>(r, c);
    double d = 0;
    for (int c = 0; c < 3; ++c) d += (double)S[c] * (double)S[c];
    const float scw = (float)std::sqrt(d);                                              // :1525
    const double inv = 1.0 / (double)scw;
    float R[9], t[3], Ow[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (float)((double)S[4 * r + c] * inv);
        t[r] = (float)((double)S[4 * r + 3] * inv);
    }
    for (int i = 0; i < 3; ++i) {                                                       // Ow = -Rcw.t()*tcw (:1528)
        float s = (float)((double)R[i] * -1.0) * t[0];
        s = s + (float)((double)R[3 + i] * -1.0) * t[1];
        s = s + (float)((double)R[6 + i] * -1.0) * t[2];
        Ow[i] = s;
    }
    const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();                    // :1531
    const size_t nq = vpPoints.size();
    detail::FreeQueries Q(nq);
    for (size_t i = 0; i < nq; ++i) {
        MapPointT* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        detail::FuseQuery(pKF, pMP, R, t, Ow, 0.f, th, true, Q, i);
    }
    const orbm_frame view = ResidentKF(pKF, true);
    std::vector<int> best(nq ? nq : 1, -1), dist(nq ? nq : 1, 256);
    int found = 0;
    Check(orbm_window_best_free(&view, (int)nq, &Q.uvr[0], &Q.level[0], 0, &Q.valid[0], &Q.desc[0], 0, 0, TH_LOW, &best[0], &dist[0], &found,
                                Device()), "orbm_window_best_free");
    int nFused = 0;
    for (size_t i = 0; i < nq; ++i) {
        if (!Q.valid[i] || best[i] < 0) continue;
        MapPointT* pMP = vpPoints[i];
        if (pMP->isBad()) continue;
        MapPointT* pMPinKF = pKF->GetMapPoint((size_t)best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;                         // :1612-1613
        } else {
            pMP->AddObservation(pKF, (size_t)best[i]);                                  // :1617-1618
            pKF->AddMapPoint(pMP, (size_t)best[i]);
        }
        nFused++;
    }
    return nFused;
}

// ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
// const cv::Mat& t12, const float th), src/ORBmatcher.cc:836-1052: the map points of each key frame moved through the similarity
// into the other one and searched there (TH_HIGH), matches kept when both directions agree.
template <class KeyFrameT, class MatT, class MapPointT>
int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12, const MatT& R12, const MatT& t12, const float th)
{
    const int TH_HIGH = 100;
    float R1w[9], t1w[3], R2w[9], t2w[3], r12[9], T12[3], sR12[9], sR21[9], nsR21[9], t21[3];
    detail::Read3x3(pKF1->GetRotation(), R1w); detail::Read3(pKF1->GetTranslation(), t1w);
    detail::Read3x3(pKF2->GetRotation(), R2w); detail::Read3(pKF2->GetTranslation(), t2w);
    detail::Read3x3(R12, r12); detail::Read3(t12, T12);
    const double inv_s = 1.0 / (double)s12;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            sR12[3 * r + c] = (float)((double)r12[3 * r + c] * (double)s12);            // :854
            sR21[3 * r + c] = (float)((double)r12[3 * c + r] * inv_s);                  // :855
            nsR21[3 * r + c] = (float)((double)sR21[3 * r + c] * -1.0);
        }
    for (int r = 0; r < 3; ++r) {                                                       // t21 = -sR21*t12 (:856)
        float s = nsR21[3 * r] * T12[0];
        s = s + nsR21[3 * r + 1] * T12[1];
        s = s + nsR21[3 * r + 2] * T12[2];
        t21[r] = s;
    }
    const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPointT*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1((size_t)N1, false), vbAlreadyMatched2((size_t)N2, false);
    for (int i = 0; i < N1; ++i) {                                                      // :864-877
        MapPointT* pMP = vpMatches12[(size_t)i];
        if (pMP) {
            vbAlreadyMatched1[(size_t)i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[(size_t)idx2] = true;
        }
    }
    // one direction: points of A through (Raw, taw) into A's camera, through (sR, tt) into B's, searched in B
    struct Dir {
        static void Run(const std::vector<MapPointT*>& ptsA, const std::vector<bool>& already, const float* Raw, const float* taw, const float* sR,
                        const float* tt, KeyFrameT* pKFB, const float* K, float th, int thAccept, std::vector<int>& match)
        {
            const size_t nA = ptsA.size();
            detail::FreeQueries Q(nA);
            for (size_t i = 0; i < nA; ++i) {
                MapPointT* pMP = ptsA[i];
                if (!pMP || already[i]) continue;
                if (pMP->isBad()) continue;
                float x[3], pa[3], pb[3];
                detail::Read3(pMP->GetWorldPos(), x);
                detail::MulAdd3(Raw, taw, x, pa);
                detail::MulAdd3(sR, tt, pa, pb);
                if (pb[2] < 0.0) continue;
                const float invz = (float)(1.0 / (double)pb[2]);
                const float xn = pb[0] * invz, yn = pb[1] * invz;
                const float u = K[0] * xn + K[2], v = K[1] * yn + K[3];
                if (!(u >= pKFB->mnMinX && u < pKFB->mnMaxX && v >= pKFB->mnMinY && v < pKFB->mnMaxY)) continue;
                double acc = 0;
                for (int r = 0; r < 3; ++r) acc += (double)pb[r] * (double)pb[r];
                const float dist3D = (float)std::sqrt(acc);
                if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
                const int lvl = pMP->PredictScale(dist3D, pKFB);
                Q.uvr[3 * i] = u; Q.uvr[3 * i + 1] = v; Q.uvr[3 * i + 2] = th * pKFB->mvScaleFactors[(size_t)lvl];
                Q.level[i] = lvl; Q.valid[i] = 1;
                std::memcpy(&Q.desc[32 * i], pMP->GetDescriptor().ptr(0), 32);
            }
            const orbm_frame view = ResidentKF(pKFB, true);
            match.assign(nA ? nA : 1, -1);
            std::vector<int> dist(nA ? nA : 1, 256);
            int found = 0;
            Check(orbm_window_best_free(&view, (int)nA, &Q.uvr[0], &Q.level[0], 0, &Q.valid[0], &Q.desc[0], 0, 0, thAccept, &match[0], &dist[0],
                                        &found, Device()), "orbm_window_best_free");
        }
    };
    const float K[4] = { pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy };                      // pKF1's intrinsics project in BOTH directions (:838-841)
    std::vector<int> vnMatch1, vnMatch2;
    Dir::Run(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, K, th, TH_HIGH, vnMatch1);
    Dir::Run(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, T12, pKF1, K, th, TH_HIGH, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; ++i1) {                                                   // :1029-1042
        const int idx2 = vnMatch1[(size_t)i1];
        if (idx2 >= 0 && vnMatch2[(size_t)idx2] == i1) { vpMatches12[(size_t)i1] = vpMapPoints2[(size_t)idx2]; nFound++; }
    }
    return nFound;
}

// ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t,size_t>>& vMatchedPairs,
// const bool bOnlyStereo), src/ORBmatcher.cc:1183-1361.  The epipole (:1190-1196) is computed here in the reference's
// arithmetic; the walk over shared vocabulary nodes, the distance / epipole / epipolar-line tests and the rotation histogram
// run on the GPU (orbm_search_for_triangulation).  Reads pKF->mFeatVec, GetMapPointMatches(), mvuRight, mvKeysUn, mDescriptors,
// GetCameraCenter() (pKF1), GetRotation(), GetTranslation(), fx, fy, cx, cy, mvScaleFactors, mvLevelSigma2 (pKF2).
template <class KeyFrameT, class MatT>
int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, const MatT& F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs,
                           const bool bOnlyStereo, const bool checkOri)
{
    float Cw[3], R2w[9], t2w[3], C2[3], F[9];
    detail::Read3(pKF1->GetCameraCenter(), Cw); detail::Read3x3(pKF2->GetRotation(), R2w); detail::Read3(pKF2->GetTranslation(), t2w);
    detail::MulAdd3(R2w, t2w, Cw, C2);                                                  // C2 = R2w*Cw + t2w
    const float invz = 1.0f / C2[2];
    float epipole[2];
    epipole[0] = pKF2->fx * C2[0] * invz + pKF2->cx;
    epipole[1] = pKF2->fy * C2[1] * invz + pKF2->cy;
    detail::Read3x3(F12, F);
    const size_t na = pKF1->mvKeysUn.size(), nb = pKF2->mvKeysUn.size();
    std::vector<unsigned char> v1(na ? na : 1, 0), v2(nb ? nb : 1, 0);
    {
        const auto p1 = pKF1->GetMapPointMatches();
        const auto p2 = pKF2->GetMapPointMatches();
        for (size_t i = 0; i < na; ++i) v1[i] = (!(i < p1.size() && p1[i]) && (!bOnlyStereo || pKF1->mvuRight[i] >= 0)) ? 1 : 0;   // :1222-1232
        for (size_t i = 0; i < nb; ++i) v2[i] = (!(i < p2.size() && p2[i]) && (!bOnlyStereo || pKF2->mvuRight[i] >= 0)) ? 1 : 0;   // :1254-1262
    }
    const orbm_frame va = ResidentFrames::Local().Get((unsigned long)pKF1->mnFrameId, (double)pKF1->mTimeStamp, pKF1->mvKeysUn, pKF1->mDescriptors,
                                                      &pKF1->mvuRight, 0.f, 1.f, 0.f, 1.f);
    const orbm_frame vb = ResidentFrames::Local().Get((unsigned long)pKF2->mnFrameId, (double)pKF2->mTimeStamp, pKF2->mvKeysUn, pKF2->mDescriptors,
                                                      &pKF2->mvuRight, 0.f, 1.f, 0.f, 1.f);
    std::vector<int> ia, oa, fa, ib, ob, fb;
    FlattenFeatureVector(pKF1->mFeatVec, ia, oa, fa); FlattenFeatureVector(pKF2->mFeatVec, ib, ob, fb);
    std::vector<int> m12(na ? na : 1, -1);
    int nmatches = 0;
    Check(orbm_search_for_triangulation(&va, &v1[0], (int)ia.size(), ia.empty() ? 0 : &ia[0], &oa[0], fa.empty() ? 0 : &fa[0],
                                        &vb, &v2[0], (int)ib.size(), ib.empty() ? 0 : &ib[0], &ob[0], fb.empty() ? 0 : &fb[0],
                                        F, epipole, Ptr(pKF2->mvScaleFactors), Ptr(pKF2->mvLevelSigma2), (int)pKF2->mvScaleFactors.size(),
                                        checkOri ? 1 : 0, &m12[0], &nmatches, Device()), "orbm_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve((size_t)(nmatches > 0 ? nmatches : 0));
    for (size_t i = 0; i < na; ++i) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));   // :1350-1356
    return nmatches;
}

}} // namespace ORB_SLAM2::b200

#endif
