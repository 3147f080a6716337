/*
 * oracle/orb_window_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * The "best candidate only" window search shared by the reference's projection overloads once the
 * points are projected, and the relocalisation overload built on it:
 *   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)  src/ORBmatcher.cc:303-431
 *   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, vector<MapPoint*>&, th)  :434-549
 * Pinned against the reference's unmodified code by tests/test_matcher_oracle.py (reloc and loop-closing overloads) and
 * against orbo_search_by_projection_frame (same loop, TH_HIGH).
 */
#include "orb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HISTO_LENGTH 30

static int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

static void three_maxima(const int* sizes, int L, int* ind1, int* ind2, int* ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < L; ++i) {
        const int s = sizes[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

/* Loop body of src/ORBmatcher.cc:213-272 / :355-398 / :500-536 with the projection done by the caller. */
int orbo_window_search_best(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                            float minX, float maxX, float minY, float maxY,
                            int nq, const float* uvr, const int* min_level, const int* max_level,
                            const float* ur, const float* er_max, const uint8_t* valid, const uint8_t* qdesc,
                            const float* q_angle, const int* q_obs, const int* init_obs, int* assign_out,
                            int th_accept, int check_ori)
{
    int nmatches = 0;
    int* obs = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* he = (int*)malloc(sizeof(int) * (size_t)(nq > 0 ? nq : 1));
    int* hb = (int*)malloc(sizeof(int) * (size_t)(nq > 0 ? nq : 1));
    int nh = 0, sizes[HISTO_LENGTH];
    memset(sizes, 0, sizeof(sizes));
    for (int k = 0; k < n; ++k) { obs[k] = init_obs ? init_obs[k] : -1; assign_out[k] = (init_obs && init_obs[k] >= 0) ? -2 : -1; }
    for (int i = 0; i < nq; ++i) {
        if (valid && !valid[i]) continue;
        const float u = uvr[3 * i], v = uvr[3 * i + 1], radius = uvr[3 * i + 2];
        int nc = orbo_features_in_area(n, kps, minX, maxX, minY, maxY, u, v, radius, min_level[i], max_level[i], cand, n);
        if (nc > n) nc = n;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            const int i2 = cand[c];
            if (obs[i2] > 0) continue;
            if (ur && u_right && u_right[i2] > 0) {
                const float er = fabsf(ur[i] - u_right[i2]);
                if (er > (er_max ? er_max[i] : 3.0e38f)) continue;
            }
            const int dist = orbo_descriptor_distance(qdesc + (size_t)i * 32, desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= th_accept) {
            assign_out[bestIdx2] = i;
            obs[bestIdx2] = q_obs ? q_obs[i] : 1;
            nmatches++;
            if (check_ori) { const int b = rot_bin(q_angle[i], kps[bestIdx2].angle); he[nh] = bestIdx2; hb[nh] = b; sizes[b]++; nh++; }
        }
    }
    if (check_ori) {
        int i1, i2, i3;
        three_maxima(sizes, HISTO_LENGTH, &i1, &i2, &i3);
        for (int e = 0; e < nh; ++e) if (hb[e] != i1 && hb[e] != i2 && hb[e] != i3) { assign_out[he[e]] = -1; nmatches--; }
    }
    free(hb); free(he); free(cand); free(obs);
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame& cur, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
 * src/ORBmatcher.cc:303-431.  Key-frame map points as arrays [nkf]: has_mp, bad, already_found, world position,
 * descriptor, predicted level (what MapPoint::PredictScale returns), min/max distance invariance, and the
 * key-frame keypoint angle.  cur_taken [n_cur]: 1 if the current keypoint already has a map point.
 * Writes the query arrays (for the device entry point) when the out pointers are given. */
int orbo_search_by_projection_reloc(int n_cur, const orbo_kp* kps_cur, const uint8_t* desc_cur,
                                    float minX, float maxX, float minY, float maxY, const float* scale,
                                    int nkf, const uint8_t* has_mp, const uint8_t* bad, const uint8_t* already_found,
                                    const float* xyz, const uint8_t* mp_desc, const int* pred_level,
                                    const float* min_dist, const float* max_dist, const float* kf_angle,
                                    const float* Tcw, const float* K, const uint8_t* cur_taken, int* assign_out,
                                    float th, int ORBdist, int check_ori,
                                    float* uvr_out, int* minl_out, int* maxl_out, uint8_t* valid_out)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    float Ow[3];                                           /* Ow = -Rcw^T * tcw (:311), float accumulation */
    for (int i = 0; i < 3; ++i) {
        float s = (-Tcw[0 * 4 + i]) * Tcw[3];
        s = s + (-Tcw[1 * 4 + i]) * Tcw[7];
        s = s + (-Tcw[2 * 4 + i]) * Tcw[11];
        Ow[i] = s;
    }
    float* uvr = (float*)calloc((size_t)(nkf > 0 ? nkf : 1) * 3, sizeof(float));
    int* minl = (int*)calloc((size_t)(nkf > 0 ? nkf : 1), sizeof(int));
    int* maxl = (int*)calloc((size_t)(nkf > 0 ? nkf : 1), sizeof(int));
    uint8_t* valid = (uint8_t*)calloc((size_t)(nkf > 0 ? nkf : 1), 1);
    for (int i = 0; i < nkf; ++i) {
        if (!has_mp[i] || bad[i] || already_found[i]) continue;
        const float* x = xyz + 3 * i;
        float pc[3];
        for (int r = 0; r < 3; ++r) {
            float s = Tcw[4 * r] * x[0];
            s = s + Tcw[4 * r + 1] * x[1];
            s = s + Tcw[4 * r + 2] * x[2];
            pc[r] = s + Tcw[4 * r + 3];
        }
        const float invzc = (float)(1.0 / (double)pc[2]);
        const float u = fx * pc[0] * invzc + cx, v = fy * pc[1] * invzc + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        /* dist3D = cv::norm(x3Dw - Ow): float differences, squares accumulated in double, sqrt, narrowed */
        double acc = 0;
        for (int r = 0; r < 3; ++r) { const float d = x[r] - Ow[r]; acc += (double)d * (double)d; }
        const float dist3D = (float)sqrt(acc);
        if (dist3D < min_dist[i] || dist3D > max_dist[i]) continue;
        const int lvl = pred_level[i];
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = th * scale[lvl];
        minl[i] = lvl - 1; maxl[i] = lvl + 1; valid[i] = 1;
    }
    /* any attached point blocks its keypoint (:373-374): Observations() plays no role here */
    int* init_obs = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
    for (int k = 0; k < n_cur; ++k) init_obs[k] = (cur_taken && cur_taken[k]) ? 1 : -1;
    const int nm = orbo_window_search_best(n_cur, kps_cur, desc_cur, NULL, minX, maxX, minY, maxY, nkf, uvr, minl, maxl, NULL, NULL, valid,
                                           mp_desc, kf_angle, NULL, init_obs, assign_out, ORBdist, check_ori);
    if (uvr_out) memcpy(uvr_out, uvr, sizeof(float) * 3 * (size_t)nkf);
    if (minl_out) memcpy(minl_out, minl, sizeof(int) * (size_t)nkf);
    if (maxl_out) memcpy(maxl_out, maxl, sizeof(int) * (size_t)nkf);
    if (valid_out) memcpy(valid_out, valid, (size_t)nkf);
    free(init_obs); free(valid); free(maxl); free(minl); free(uvr);
    return nm;
}

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th),
 * src/ORBmatcher.cc:434-549 (loop closing).  Candidate points as arrays [npts]; matched_in [n]: -1 free, k >= 0 candidate k already
 * matched to the keypoint (such candidates are not searched again, :449-450, 460), -2 some other point.  Accept at TH_LOW, no
 * orientation check.  The Sim3 is decomposed and the points projected in the reference's arithmetic (cv::Mat algebra of
 * oracle/cvshim: dot / norm accumulate in double, Mat/scalar multiplies by the double reciprocal, products accumulate in float). */
int orbo_search_by_projection_sim3(int n, const orbo_kp* kps, const uint8_t* desc,
                                   float minX, float maxX, float minY, float maxY, const float* scale,
                                   const float* K, const float* Scw, int npts, const uint8_t* bad, const float* xyz, const float* normal,
                                   const uint8_t* mp_desc, const int* pred_level, const float* min_dist, const float* max_dist,
                                   const int* matched_in, int* assign_out, int th,
                                   float* uvr_out, int* minl_out, int* maxl_out, uint8_t* valid_out)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    /* scw = sqrt(sRcw.row(0).dot(sRcw.row(0))) (:446); Rcw = sRcw/scw; tcw = Scw(0:3,3)/scw (:447-448) */
    double d = 0;
    for (int c = 0; c < 3; ++c) d += (double)Scw[c] * (double)Scw[c];
    const float scw = (float)sqrt(d);
    const double inv = 1.0 / (double)scw;
    float R[9], t[3], Ow[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[3 * r + c] = (float)((double)Scw[4 * r + c] * inv);
        t[r] = (float)((double)Scw[4 * r + 3] * inv);
    }
    for (int i = 0; i < 3; ++i) {                          /* Ow = -Rcw.t()*tcw (:449) */
        float s = (float)((double)R[0 * 3 + i] * -1.0) * t[0];
        s = s + (float)((double)R[1 * 3 + i] * -1.0) * t[1];
        s = s + (float)((double)R[2 * 3 + i] * -1.0) * t[2];
        Ow[i] = s;
    }
    const int iminX = (int)minX, imaxX = (int)maxX, iminY = (int)minY, imaxY = (int)maxY;   /* KeyFrame keeps int bounds */
    float* uvr = (float*)calloc((size_t)(npts > 0 ? npts : 1) * 3, sizeof(float));
    int* minl = (int*)calloc((size_t)(npts > 0 ? npts : 1), sizeof(int));
    int* maxl = (int*)calloc((size_t)(npts > 0 ? npts : 1), sizeof(int));
    uint8_t* valid = (uint8_t*)calloc((size_t)(npts > 0 ? npts : 1), 1);
    uint8_t* found = (uint8_t*)calloc((size_t)(npts > 0 ? npts : 1), 1);
    for (int k = 0; k < n; ++k) if (matched_in[k] >= 0) found[matched_in[k]] = 1;
    for (int i = 0; i < npts; ++i) {
        if (bad[i] || found[i]) continue;
        const float* x = xyz + 3 * i;
        float pc[3];
        for (int r = 0; r < 3; ++r) {                      /* p3Dc = Rcw*p3Dw + tcw (:465) */
            float s = R[3 * r] * x[0];
            s = s + R[3 * r + 1] * x[1];
            s = s + R[3 * r + 2] * x[2];
            pc[r] = s + t[r];
        }
        if (pc[2] < 0.0) continue;
        const float invz = 1 / pc[2];                      /* float division here (:472), unlike the other overloads */
        const float xn = pc[0] * invz, yn = pc[1] * invz;
        const float u = fx * xn + cx, v = fy * yn + cy;
        if (!(u >= iminX && u < imaxX && v >= iminY && v < imaxY)) continue;   /* KeyFrame::IsInImage */
        float PO[3];
        double acc = 0;
        for (int r = 0; r < 3; ++r) { PO[r] = x[r] - Ow[r]; acc += (double)PO[r] * (double)PO[r]; }
        const float dist = (float)sqrt(acc);
        if (dist < min_dist[i] || dist > max_dist[i]) continue;
        double dn = 0;                                     /* PO.dot(Pn) < 0.5*dist (:495) */
        for (int r = 0; r < 3; ++r) dn += (double)PO[r] * (double)normal[3 * i + r];
        if (dn < 0.5 * dist) continue;
        const int lvl = pred_level[i];
        uvr[3 * i] = u; uvr[3 * i + 1] = v; uvr[3 * i + 2] = th * scale[lvl];
        minl[i] = lvl - 1; maxl[i] = lvl; valid[i] = 1;    /* level test of the loop (:526) */
    }
    int* init_obs = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) init_obs[k] = matched_in[k] != -1 ? 1 : -1;
    /* the key frame's bounds are ints: its grid origin is (float)(int)minX */
    const int nm = orbo_window_search_best(n, kps, desc, NULL, (float)iminX, maxX, (float)iminY, maxY, npts, uvr, minl, maxl, NULL, NULL, valid,
                                           mp_desc, NULL, NULL, init_obs, assign_out, 50, 0);
    for (int k = 0; k < n; ++k) if (assign_out[k] == -2) assign_out[k] = matched_in[k];   /* pre-matched entries stay as they were */
    if (uvr_out) memcpy(uvr_out, uvr, sizeof(float) * 3 * (size_t)npts);
    if (minl_out) memcpy(minl_out, minl, sizeof(int) * (size_t)npts);
    if (maxl_out) memcpy(maxl_out, maxl, sizeof(int) * (size_t)npts);
    if (valid_out) memcpy(valid_out, valid, (size_t)npts);
    free(init_obs); free(found); free(valid); free(maxl); free(minl); free(uvr);
    return nm;
}
