// oracle/ref_bow_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points around the reference's UNMODIFIED vocabulary-guided matchers
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)        src/ORBmatcher.cc:552-697
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)     src/ORBmatcher.cc:700-832
// DBoW2 itself is not part of the reference checkout; the feature vectors (NodeId -> feature indices, the
// std::map DBoW2::FeatureVector is) are supplied by the caller in CSR form: node_id [nn] ascending, node_off [nn+1],
// feat [node_off[nn]].
#define private public
#define protected public
#include "Frame.h"
#include "ORBmatcher.h"
#undef private
#undef protected

#include <cstring>
#include <vector>

#include "ref_arena.hpp"

using namespace ORB_SLAM2;

namespace {
struct RefKp { float x, y, size, angle, response; int octave, class_id; };

void fill_featvec(DBoW2::FeatureVector& fv, int nn, const int* node_id, const int* node_off, const int* feat)
{
    for (int k = 0; k < nn; ++k) {
        std::vector<unsigned int>& v = fv[(DBoW2::NodeId)node_id[k]];
        for (int j = node_off[k]; j < node_off[k + 1]; ++j) v.push_back((unsigned int)feat[j]);
    }
}
void fill_kf(KeyFrame& KF, int n, const RefKp* kps, const unsigned char* desc, const unsigned char* has_mp, const unsigned char* mp_bad,
             std::vector<MapPoint>& store)
{
    KF.N = n;
    KF.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) KF.mvKeysUn[i] = cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
    KF.mvKeys = KF.mvKeysUn;
    KF.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(KF.mDescriptors.data, desc, (size_t)n * 32);
    store.resize(n > 0 ? n : 1);
    KF.mapPoints.assign(n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < n; ++i)
        if (has_mp[i]) { store[i].bad = mp_bad && mp_bad[i]; KF.mapPoints[i] = &store[i]; }
}
} // namespace

extern "C" {

// match_f [nf]: index of the key-frame feature whose map point ends up in vpMapPointMatches[i], -1 for NULL.
int orbref_search_by_bow_kf_frame(int nk, const RefKp* kps_k, const unsigned char* desc_k, const unsigned char* has_mp, const unsigned char* mp_bad,
                                  int nn_k, const int* node_id_k, const int* node_off_k, const int* feat_k,
                                  int nf, const RefKp* kps_f, const unsigned char* desc_f,
                                  int nn_f, const int* node_id_f, const int* node_off_f, const int* feat_f,
                                  float nnratio, int check_ori, int* match_f)
{
    ref_arena::Scope scope;
    int nm;
    {
        std::vector<MapPoint> store;
        KeyFrame KF;
        fill_kf(KF, nk, kps_k, desc_k, has_mp, mp_bad, store);
        fill_featvec(KF.mFeatVec, nn_k, node_id_k, node_off_k, feat_k);
        Frame F;
        F.N = nf;
        F.mvKeys.resize(nf);
        for (int i = 0; i < nf; ++i) F.mvKeys[i] = cv::KeyPoint(kps_f[i].x, kps_f[i].y, kps_f[i].size, kps_f[i].angle, kps_f[i].response, kps_f[i].octave, kps_f[i].class_id);
        F.mvKeysUn = F.mvKeys;
        F.mDescriptors.create(nf > 0 ? nf : 1, 32, CV_8U);
        if (nf > 0) std::memcpy(F.mDescriptors.data, desc_f, (size_t)nf * 32);
        fill_featvec(F.mFeatVec, nn_f, node_id_f, node_off_f, feat_f);
        std::vector<MapPoint*> matches;
        ORBmatcher matcher(nnratio, check_ori != 0);
        nm = matcher.SearchByBoW(&KF, F, matches);
        for (int i = 0; i < nf; ++i) match_f[i] = matches[i] ? (int)(matches[i] - &store[0]) : -1;
    }
    return nm;
}

// match12 [n1]: index of the key-frame-2 feature whose map point ends up in vpMatches12[i], -1 for NULL.
int orbref_search_by_bow_kf_kf(int n1, const RefKp* kps1, const unsigned char* desc1, const unsigned char* has_mp1, const unsigned char* mp_bad1,
                               int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                               int n2, const RefKp* kps2, const unsigned char* desc2, const unsigned char* has_mp2, const unsigned char* mp_bad2,
                               int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                               float nnratio, int check_ori, int* match12)
{
    ref_arena::Scope scope;
    int nm;
    {
        std::vector<MapPoint> store1, store2;
        KeyFrame K1, K2;
        fill_kf(K1, n1, kps1, desc1, has_mp1, mp_bad1, store1);
        fill_kf(K2, n2, kps2, desc2, has_mp2, mp_bad2, store2);
        fill_featvec(K1.mFeatVec, nn1, node_id1, node_off1, feat1);
        fill_featvec(K2.mFeatVec, nn2, node_id2, node_off2, feat2);
        std::vector<MapPoint*> matches;
        ORBmatcher matcher(nnratio, check_ori != 0);
        nm = matcher.SearchByBoW(&K1, &K2, matches);
        for (int i = 0; i < n1; ++i) match12[i] = matches[i] ? (int)(matches[i] - &store2[0]) : -1;
    }
    return nm;
}

} // extern "C"

// ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12, vector<pair<size_t,size_t>>&, bOnlyStereo),
// src/ORBmatcher.cc:1183-1361.  pose2 = R2w (9) | t2w (3) of key frame 2, Cw = key frame 1's camera centre, K2 = fx, fy, cx, cy
// of key frame 2 (the epipole is computed inside, :1190-1196); match12 [n1] = matched feature of key frame 2 or -1.
extern "C" int orbref_search_for_triangulation(int n1, const RefKp* kps1, const unsigned char* desc1, const unsigned char* has_mp1, const float* u_right1,
                                               int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                                               int n2, const RefKp* kps2, const unsigned char* desc2, const unsigned char* has_mp2, const float* u_right2,
                                               int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                                               const float* F12, const float* Cw, const float* pose2, const float* K2,
                                               const float* scale, const float* sigma2, int nlevels, int only_stereo, int check_ori, int* match12)
{
    ref_arena::Scope scope;
    int nm;
    {
        std::vector<MapPoint> store1, store2;
        KeyFrame K1, KB;
        fill_kf(K1, n1, kps1, desc1, has_mp1, NULL, store1);
        fill_kf(KB, n2, kps2, desc2, has_mp2, NULL, store2);
        fill_featvec(K1.mFeatVec, nn1, node_id1, node_off1, feat1);
        fill_featvec(KB.mFeatVec, nn2, node_id2, node_off2, feat2);
        K1.mvuRight.assign(n1, -1.0f); KB.mvuRight.assign(n2, -1.0f);
        if (u_right1) K1.mvuRight.assign(u_right1, u_right1 + n1);
        if (u_right2) KB.mvuRight.assign(u_right2, u_right2 + n2);
        KB.mvScaleFactors.assign(scale, scale + nlevels);
        KB.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
        KB.fx = K2[0]; KB.fy = K2[1]; KB.cx = K2[2]; KB.cy = K2[3];
        K1.Ow.create(3, 1, CV_32F); KB.R.create(3, 3, CV_32F); KB.t.create(3, 1, CV_32F);
        for (int r = 0; r < 3; ++r) {
            K1.Ow.at<float>(r) = Cw[r]; KB.t.at<float>(r) = pose2[9 + r];
            for (int c = 0; c < 3; ++c) KB.R.at<float>(r, c) = pose2[3 * r + c];
        }
        cv::Mat F(3, 3, CV_32F);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F.at<float>(r, c) = F12[3 * r + c];
        std::vector<std::pair<size_t, size_t> > pairs;
        ORBmatcher matcher(0.6f, check_ori != 0);
        nm = matcher.SearchForTriangulation(&K1, &KB, F, pairs, only_stereo != 0);
        for (int i = 0; i < n1; ++i) match12[i] = -1;
        for (size_t k = 0; k < pairs.size(); ++k) match12[pairs[k].first] = (int)pairs[k].second;
    }
    return nm;
}
