/*
 * oracle/orb_fuse_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * CPU restatement of the searches whose queries do not depend on one another:
 *   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th)                               src/ORBmatcher.cc:1364-1513
 *   ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, th, vpReplacePoint)  src/ORBmatcher.cc:1516-1633
 *   ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vpMatches12, s12, R12, t12, th)          src/ORBmatcher.cc:836-1052
 * and of the search they share (orbo_window_best_free = the contract of the device's orbm_window_best_free).
 * Pinned against the reference's unmodified code (oracle/_ref, ref_fuse_harness.cc) by tests/test_fuse_oracle.py.
 * Map points are rows of plain arrays ("the universe"); a point is named by its row, -1 = NULL.  The map bookkeeping the
 * reference does after a match follows the mock of oracle/mock/mock_slam.hpp, which keeps the parts of
 * MapPoint::Replace / AddObservation (src/MapPoint.cc:93-118, 204-258) that the loops of Fuse can observe.
 * Matrix arithmetic as oracle/cvshim does it (checked against cv2 in tests/test_cv_prims.py): products accumulate in float,
 * Mat * scalar and -Mat multiply in double and narrow, dot / norm accumulate in double.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* best candidate of one window: KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:637-676) + the candidate loop
 * (src/ORBmatcher.cc:1428-1481 with chi2, :1586-1604 / :900-918 without) */
static int best_in_window(const void* grid, int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                          float u, float v, float radius, int level, float ur, const uint8_t* qd,
                          const float* inv_sigma2, int* cand, int* best_dist_out)
{
    int nc = orbo_grid_query(grid, u, v, radius, -1, -1, cand, n);
    if (nc > n) nc = n;
    int bestDist = INT_MAX, bestIdx = -1;
    for (int c = 0; c < nc; ++c) {
        const int idx = cand[c];
        const int kpLevel = kps[idx].octave;
        if (kpLevel < level - 1 || kpLevel > level) continue;
        if (inv_sigma2) {
            const float ex = u - kps[idx].x, ey = v - kps[idx].y;
            if (u_right && u_right[idx] >= 0) {
                const float er = ur - u_right[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_sigma2[kpLevel] > 5.99) continue;
            }
        }
        const int dist = orbo_descriptor_distance(qd, desc + (size_t)idx * 32);
        if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    *best_dist_out = bestDist;
    return bestIdx;
}

/* The device contract: best_idx = -1 beyond th_accept, best_dist = 256 without a candidate.  Returns the entries >= 0. */
int orbo_window_best_free(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                          float minX, float maxX, float minY, float maxY,
                          int nq, const float* uvr, const int* level, const float* ur, const uint8_t* valid, const uint8_t* qdesc,
                          const float* inv_sigma2, int th_accept, int* best_idx, int* best_dist)
{
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    void* grid = orbo_grid_create(n, kps, minX, maxX, minY, maxY);
    int found = 0;
    for (int i = 0; i < nq; ++i) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (valid && !valid[i]) continue;
        int d;
        const int b = best_in_window(grid, n, kps, desc, u_right, uvr[3 * i], uvr[3 * i + 1], uvr[3 * i + 2], level[i],
                                     ur ? ur[i] : 0.f, qdesc + (size_t)i * 32, inv_sigma2, cand, &d);
        if (b >= 0) best_dist[i] = d;
        if (b >= 0 && d <= th_accept) { best_idx[i] = b; ++found; }
    }
    orbo_grid_destroy(grid);
    free(cand);
    return found;
}

/* mock MapPoint::AddObservation (src/MapPoint.cc:93-105) for the one key frame of the call */
static void add_observation(int p, int idx, const float* u_right, int* nobs, int* kf_idx)
{
    if (kf_idx[p] >= 0) return;
    kf_idx[p] = idx;
    nobs[p] += (u_right && u_right[idx] >= 0) ? 2 : 1;
}
/* mock MapPoint::Replace (src/MapPoint.cc:204-258): `from` goes bad and its slot in the key frame passes to `to` */
static void replace_point(int from, int to, const float* u_right, uint8_t* bad, int* replaced_by, int* nobs, int* kf_idx, int* kf_mp)
{
    if (from == to) return;
    const int idx = kf_idx[from];
    kf_idx[from] = -1;
    bad[from] = 1;
    replaced_by[from] = to;
    if (idx >= 0) {
        if (kf_idx[to] < 0) { kf_mp[idx] = to; add_observation(to, idx, u_right, nobs, kf_idx); }
        else kf_mp[idx] = -1;
    }
}

static void project3(const float* R, const float* t, const float* x, float* pc)   /* R*x + t, float accumulation */
{
    for (int r = 0; r < 3; ++r) {
        float s = R[3 * r] * x[0];
        s = s + R[3 * r + 1] * x[1];
        s = s + R[3 * r + 2] * x[2];
        pc[r] = s + t[r];
    }
}

/* Shared body of the two Fuse overloads.  sim3 = 0: src/ORBmatcher.cc:1364-1513 (R, t, Ow = the key frame's pose, chi-square
 * test, replacement by observation count); sim3 = 1: :1516-1633 (R, t, Ow from Scw, already-found set, vpReplacePoint).
 * list [nlist]: the candidate points (rows of the universe, -1 = NULL).  State arrays are updated in place:
 * bad, nobs (Observations()), kf_idx (GetIndexInKeyFrame(pKF), -1 = not in the key frame) [npts]; kf_mp [n] (GetMapPoint).
 * replaced_by [npts] (init -1) receives Replace's argument; replace_out [nlist] (sim3 only) = vpReplacePoint.
 * The queries the loop searched come back in uvr_out [nlist][3], level_out, ur_out, valid_out (valid = reached the search
 * at the time of the call's START: the device searches these, the host then replays the bookkeeping). */
static int fuse_body(int sim3, int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                     float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2,
                     const float* K, float bf, const float* R, const float* t, const float* Ow,
                     int npts, uint8_t* bad, const float* xyz, const float* normal, const uint8_t* mp_desc, const int* pred_level,
                     const float* min_dist, const float* max_dist, int* nobs, int* kf_idx, int* replaced_by,
                     int nlist, const int* list, int* kf_mp, int* replace_out, float th,
                     float* uvr_out, int* level_out, float* ur_out, uint8_t* valid_out)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const int iminX = (int)minX, imaxX = (int)maxX, iminY = (int)minY, imaxY = (int)maxY;   /* KeyFrame keeps int bounds */
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    void* grid = orbo_grid_create(n, kps, (float)iminX, maxX, (float)iminY, maxY);   /* the key frame's grid origin is its int bound */
    uint8_t* found0 = (uint8_t*)calloc((size_t)(npts > 0 ? npts : 1), 1);
    uint8_t* bad0 = (uint8_t*)malloc((size_t)(npts > 0 ? npts : 1));
    int* kfidx0 = (int*)malloc(sizeof(int) * (size_t)(npts > 0 ? npts : 1));
    memcpy(bad0, bad, (size_t)npts);
    memcpy(kfidx0, kf_idx, sizeof(int) * (size_t)npts);
    if (sim3) for (int k = 0; k < n; ++k) if (kf_mp[k] >= 0) found0[kf_mp[k]] = 1;          /* spAlreadyFound = pKF->GetMapPoints() (:1533) */
    int nFused = 0;
    for (int i = 0; i < nlist; ++i) {
        if (uvr_out) { uvr_out[3 * i] = uvr_out[3 * i + 1] = uvr_out[3 * i + 2] = 0.f; level_out[i] = 0; ur_out[i] = 0.f; valid_out[i] = 0; }
        const int p = list[i];
        if (p < 0) continue;                                                                 /* :1383-1384 (NULL) */
        const int skip_now = sim3 ? (bad[p] || found0[p]) : (bad[p] || kf_idx[p] >= 0);      /* :1385 / :1543 */
        const int skip_start = sim3 ? (bad0[p] || found0[p]) : (bad0[p] || kfidx0[p] >= 0);
        const float* x = xyz + 3 * p;
        float pc[3];
        project3(R, t, x, pc);
        if (pc[2] < 0.0f) continue;
        const float invz = sim3 ? (float)(1.0 / (double)pc[2]) : 1 / pc[2];                  /* :1552 / :1394 */
        const float xn = pc[0] * invz, yn = pc[1] * invz;
        const float u = fx * xn + cx, v = fy * yn + cy;
        if (!(u >= iminX && u < imaxX && v >= iminY && v < imaxY)) continue;                 /* KeyFrame::IsInImage */
        const float ur = u - bf * invz;
        float PO[3];
        double acc = 0;
        for (int r = 0; r < 3; ++r) { PO[r] = x[r] - Ow[r]; acc += (double)PO[r] * (double)PO[r]; }
        const float dist3D = (float)sqrt(acc);
        if (dist3D < min_dist[p] || dist3D > max_dist[p]) continue;
        double dn = 0;
        for (int r = 0; r < 3; ++r) dn += (double)PO[r] * (double)normal[3 * p + r];
        if (dn < 0.5 * dist3D) continue;
        const int lvl = pred_level[p];
        const float radius = th * scale[lvl];
        if (uvr_out && !skip_start) { uvr_out[3 * i] = u; uvr_out[3 * i + 1] = v; uvr_out[3 * i + 2] = radius; level_out[i] = lvl; ur_out[i] = ur; valid_out[i] = 1; }
        if (skip_now) continue;
        int bestDist;
        const int bestIdx = best_in_window(grid, n, kps, desc, u_right, u, v, radius, lvl, ur,
                                           mp_desc + (size_t)p * 32, sim3 ? NULL : inv_sigma2, cand, &bestDist);
        if (bestIdx >= 0 && bestDist <= 50) {                                                /* TH_LOW, :1483 / :1607 */
            const int inKF = kf_mp[bestIdx];
            if (inKF >= 0) {
                if (!bad[inKF]) {
                    if (sim3) replace_out[i] = inKF;                                         /* :1612-1613 */
                    else if (nobs[inKF] > nobs[p]) replace_point(p, inKF, u_right, bad, replaced_by, nobs, kf_idx, kf_mp);   /* :1491-1494 */
                    else replace_point(inKF, p, u_right, bad, replaced_by, nobs, kf_idx, kf_mp);
                }
            } else {
                add_observation(p, bestIdx, u_right, nobs, kf_idx);                          /* :1500-1501 / :1617-1618 */
                kf_mp[bestIdx] = p;
            }
            nFused++;
        }
    }
    orbo_grid_destroy(grid);
    free(kfidx0); free(bad0); free(found0); free(cand);
    return nFused;
}

int orbo_fuse(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
              float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2,
              const float* K, float bf, const float* Rcw, const float* tcw, const float* Ow,
              int npts, uint8_t* bad, const float* xyz, const float* normal, const uint8_t* mp_desc, const int* pred_level,
              const float* min_dist, const float* max_dist, int* nobs, int* kf_idx, int* replaced_by,
              int nlist, const int* list, int* kf_mp, float th,
              float* uvr_out, int* level_out, float* ur_out, uint8_t* valid_out)
{
    return fuse_body(0, n, kps, desc, u_right, minX, maxX, minY, maxY, scale, inv_sigma2, K, bf, Rcw, tcw, Ow, npts, bad, xyz, normal,
                     mp_desc, pred_level, min_dist, max_dist, nobs, kf_idx, replaced_by, nlist, list, kf_mp, NULL, th,
                     uvr_out, level_out, ur_out, valid_out);
}

/* Sim3 pieces the way the reference computes them (:1524-1530 / src/ORBmatcher.cc:443-449) */
static void decompose_sim3(const float* Scw, float* R, float* t, float* Ow)
{
    double d = 0;
    for (int c = 0; c < 3; ++c) d += (double)Scw[c] * (double)Scw[c];
    const float scw = (float)sqrt(d);
    const double inv = 1.0 / (double)scw;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (float)((double)Scw[4 * r + c] * inv);
        t[r] = (float)((double)Scw[4 * r + 3] * inv);
    }
    for (int i = 0; i < 3; ++i) {
        float s = (float)((double)R[0 * 3 + i] * -1.0) * t[0];
        s = s + (float)((double)R[1 * 3 + i] * -1.0) * t[1];
        s = s + (float)((double)R[2 * 3 + i] * -1.0) * t[2];
        Ow[i] = s;
    }
}

int orbo_fuse_sim3(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                   float minX, float maxX, float minY, float maxY, const float* scale,
                   const float* K, const float* Scw,
                   int npts, uint8_t* bad, const float* xyz, const float* normal, const uint8_t* mp_desc, const int* pred_level,
                   const float* min_dist, const float* max_dist, int* nobs, int* kf_idx,
                   int nlist, const int* list, int* kf_mp, int* replace_out, float th,
                   float* uvr_out, int* level_out, float* ur_out, uint8_t* valid_out)
{
    float R[9], t[3], Ow[3];
    decompose_sim3(Scw, R, t, Ow);
    int* replaced_by = (int*)malloc(sizeof(int) * (size_t)(npts > 0 ? npts : 1));
    const int nf = fuse_body(1, n, kps, desc, u_right, minX, maxX, minY, maxY, scale, NULL, K, 0.f, R, t, Ow, npts, bad, xyz, normal,
                             mp_desc, pred_level, min_dist, max_dist, nobs, kf_idx, replaced_by, nlist, list, kf_mp, replace_out, th,
                             uvr_out, level_out, ur_out, valid_out);
    free(replaced_by);
    return nf;
}

/* One direction of SearchBySim3 (:878-924 / :929-975): the points of key frame A (mpA [nA], rows of the universe) that are not
 * matched yet, moved into A's camera (Raw, taw), through the similarity (sR, tt) into B's camera, searched in B.
 * match [nA] = vnMatch (keypoint of B or -1). */
static void sim3_direction(int nA, const int* mpA, const uint8_t* already, const float* Raw, const float* taw, const float* sR, const float* tt,
                           int nB, const orbo_kp* kpsB, const uint8_t* descB, float minX, float maxX, float minY, float maxY,
                           const float* scale, const float* K, const uint8_t* bad, const float* xyz, const uint8_t* mp_desc, const int* pred_level,
                           const float* min_dist, const float* max_dist, float th, int* match,
                           float* uvr_out, int* level_out, uint8_t* valid_out)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const int iminX = (int)minX, imaxX = (int)maxX, iminY = (int)minY, imaxY = (int)maxY;
    int* cand = (int*)malloc(sizeof(int) * (size_t)(nB > 0 ? nB : 1));
    void* grid = orbo_grid_create(nB, kpsB, (float)iminX, maxX, (float)iminY, maxY);
    for (int i = 0; i < nA; ++i) {
        match[i] = -1;
        if (uvr_out) { uvr_out[3 * i] = uvr_out[3 * i + 1] = uvr_out[3 * i + 2] = 0.f; level_out[i] = 0; valid_out[i] = 0; }
        const int p = mpA[i];
        if (p < 0 || already[i]) continue;
        if (bad[p]) continue;
        float pa[3], pb[3];
        project3(Raw, taw, xyz + 3 * p, pa);
        project3(sR, tt, pa, pb);
        if (pb[2] < 0.0) continue;
        const float invz = (float)(1.0 / (double)pb[2]);
        const float xn = pb[0] * invz, yn = pb[1] * invz;
        const float u = fx * xn + cx, v = fy * yn + cy;
        if (!(u >= iminX && u < imaxX && v >= iminY && v < imaxY)) continue;
        double acc = 0;
        for (int r = 0; r < 3; ++r) acc += (double)pb[r] * (double)pb[r];
        const float dist3D = (float)sqrt(acc);
        if (dist3D < min_dist[p] || dist3D > max_dist[p]) continue;
        const int lvl = pred_level[p];
        const float radius = th * scale[lvl];
        if (uvr_out) { uvr_out[3 * i] = u; uvr_out[3 * i + 1] = v; uvr_out[3 * i + 2] = radius; level_out[i] = lvl; valid_out[i] = 1; }
        int bestDist;
        const int bestIdx = best_in_window(grid, nB, kpsB, descB, NULL, u, v, radius, lvl, 0.f,
                                           mp_desc + (size_t)p * 32, NULL, cand, &bestDist);
        if (bestIdx >= 0 && bestDist <= 100) match[i] = bestIdx;                             /* TH_HIGH, :920 / :971 */
    }
    orbo_grid_destroy(grid);
    free(cand);
}

/* src/ORBmatcher.cc:836-1052.  mp1 [n1] / mp2 [n2]: GetMapPointMatches() of the two key frames (rows of the universe, -1 =
 * NULL); matches12 [n1]: vpMatches12 in and out; idx_in_kf2 [npts]: GetIndexInKeyFrame(pKF2).  Both key frames share the
 * camera (K, bounds, scale factors).  Optional outputs: the queries of both directions. */
int orbo_search_by_sim3(int n1, const orbo_kp* kps1, const uint8_t* desc1, const int* mp1,
                        int n2, const orbo_kp* kps2, const uint8_t* desc2, const int* mp2,
                        float minX, float maxX, float minY, float maxY, const float* scale, const float* K,
                        const float* R1w, const float* t1w, const float* R2w, const float* t2w, float s12, const float* R12, const float* t12,
                        int npts, const uint8_t* bad, const float* xyz, const uint8_t* mp_desc, const int* pred_level,
                        const float* min_dist, const float* max_dist, const int* idx_in_kf2,
                        int* matches12, float th,
                        float* uvr1_out, int* level1_out, uint8_t* valid1_out, float* uvr2_out, int* level2_out, uint8_t* valid2_out)
{
    (void)npts;
    float sR12[9], sR21[9], nsR21[9], t21[3];
    const double inv_s = 1.0 / (double)s12;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            sR12[3 * r + c] = (float)((double)R12[3 * r + c] * (double)s12);                /* :854 */
            sR21[3 * r + c] = (float)((double)R12[3 * c + r] * inv_s);                       /* :855 */
            nsR21[3 * r + c] = (float)((double)sR21[3 * r + c] * -1.0);
        }
    for (int r = 0; r < 3; ++r) {                                                            /* t21 = -sR21*t12 (:856) */
        float s = nsR21[3 * r] * t12[0];
        s = s + nsR21[3 * r + 1] * t12[1];
        s = s + nsR21[3 * r + 2] * t12[2];
        t21[r] = s;
    }
    uint8_t* am1 = (uint8_t*)calloc((size_t)(n1 > 0 ? n1 : 1), 1);
    uint8_t* am2 = (uint8_t*)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
    for (int i = 0; i < n1; ++i) {                                                           /* :864-877 */
        const int p = matches12[i];
        if (p >= 0) {
            am1[i] = 1;
            const int idx2 = idx_in_kf2[p];
            if (idx2 >= 0 && idx2 < n2) am2[idx2] = 1;
        }
    }
    int* m1 = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    int* m2 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    sim3_direction(n1, mp1, am1, R1w, t1w, sR21, t21, n2, kps2, desc2, minX, maxX, minY, maxY, scale, K, bad, xyz, mp_desc, pred_level,
                   min_dist, max_dist, th, m1, uvr1_out, level1_out, valid1_out);
    sim3_direction(n2, mp2, am2, R2w, t2w, sR12, t12, n1, kps1, desc1, minX, maxX, minY, maxY, scale, K, bad, xyz, mp_desc, pred_level,
                   min_dist, max_dist, th, m2, uvr2_out, level2_out, valid2_out);
    int nFound = 0;
    for (int i1 = 0; i1 < n1; ++i1) {                                                        /* :1029-1042 */
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { matches12[i1] = mp2[idx2]; nFound++; }
    }
    free(m2); free(m1); free(am2); free(am1);
    return nFound;
}

/* The per-point prologue alone (projection, gates, MapPoint::PredictScale(dist, KeyFrame*) src/MapPoint.cc:442-457, radius), i.e. the
 * contract of the device's orbm_fuse_project_batch for one problem.  sim3 = 0: Fuse (src/ORBmatcher.cc:1388-1426, :1546-1584), pose =
 * R(9) t(3) Ow(3); sim3 = 1: one direction of SearchBySim3 (:886-909 / :937-960), pose = R_a(9) t_a(3) sR(9) tt(3).  max_d / min_d = the
 * RAW mfMaxDistance / mfMinDistance (the invariance factors 1.2 / 0.8 are applied here, src/MapPoint.cc:424-435).  Consistent with
 * fuse_body / sim3_direction above (tests/test_fuse_oracle.py feeds this function's levels to them as the preset PredictScale and
 * compares the queries), which are pinned against the reference. */
void orbo_fuse_project(int sim3, const float* pose, const float* K, float bf, float minX, float maxX, float minY, float maxY,
                       float scale_factor, const float* scale, int nlevels, float th,
                       int n, const float* xyz, const float* normal, const float* max_d, const float* min_d, const uint8_t* skip,
                       float* uvr, int* level, float* ur, uint8_t* valid)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const float logs = orbo_log_scale_factor(scale_factor);
    for (int i = 0; i < n; ++i) {
        uvr[3 * i] = uvr[3 * i + 1] = uvr[3 * i + 2] = 0.f; level[i] = 0; valid[i] = 0;
        if (ur) ur[i] = 0.f;
        if (skip && skip[i]) continue;
        const float* x = xyz + 3 * i;
        float pc[3];
        project3(pose, pose + 9, x, pc);
        if (sim3) { float pb[3]; project3(pose + 12, pose + 21, pc, pb); pc[0] = pb[0]; pc[1] = pb[1]; pc[2] = pb[2]; }
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float xn = pc[0] * invz, yn = pc[1] * invz;
        const float u = fx * xn + cx, v = fy * yn + cy;
        if (!(u >= minX && u < maxX && v >= minY && v < maxY)) continue;
        const float maxD = 1.2f * max_d[i], minD = 0.8f * min_d[i];
        float d[3];
        double acc = 0;
        for (int r = 0; r < 3; ++r) { d[r] = sim3 ? pc[r] : x[r] - pose[12 + r]; acc += (double)d[r] * (double)d[r]; }
        const float dist3D = (float)sqrt(acc);
        if (dist3D < minD || dist3D > maxD) continue;
        if (!sim3) {
            double dn = 0;
            for (int r = 0; r < 3; ++r) dn += (double)d[r] * (double)normal[3 * i + r];
            if (dn < 0.5 * dist3D) continue;
        }
        const int lvl = orbo_predict_scale(max_d[i], dist3D, logs, nlevels);
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = th * scale[lvl];
        level[i] = lvl; valid[i] = 1;
        if (ur) ur[i] = u - bf * invz;
    }
}
