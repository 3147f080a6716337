/* oracle/orb_mappoint_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the map-point side of the matching path:
 *   MapPoint::ComputeDistinctiveDescriptors   src/MapPoint.cc:275-340
 *   MapPoint::PredictScale                    src/MapPoint.cc:442-475
 *   Frame::SetPose / UpdatePoseMatrices       src/Frame.cc:271-285
 *   Frame::isInFrustum                        src/Frame.cc:288-345
 * Float semantics = what oracle/cvshim gives the reference (cv::gemm order for 3x3 * 3x1: products left to
 * right in float; cv::norm and Mat::dot accumulate in double), glibc logf (cv_prims.c).  Pinned against the
 * reference's own sources through oracle/_ref (tests/test_projection_oracle.py) and fixtures under tests/golden.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "cv_prims.h"
#include "orb_oracle.h"

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

/* desc [n][32], bad [n] or NULL (KeyFrame::isBad of the observing key frame, :299).  Returns the index (in the
 * caller's order) of the chosen descriptor, -1 when none is usable; *median_out = its median distance. */
int orbo_distinctive_descriptor(const uint8_t* desc, int n, const uint8_t* bad, int* median_out)
{
    int* idx = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int N = 0;
    for (int i = 0; i < n; ++i) if (!bad || !bad[i]) idx[N++] = i;
    if (N == 0) { free(idx); return -1; }
    int* row = (int*)malloc(sizeof(int) * (size_t)N);
    int best_median = INT_MAX, best = 0;
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) row[j] = i == j ? 0 : orbo_descriptor_distance(desc + 32 * (size_t)idx[i], desc + 32 * (size_t)idx[j]);
        qsort(row, (size_t)N, sizeof(int), cmp_int);
        const int median = row[(size_t)(0.5 * (N - 1))];           /* :327 */
        if (median < best_median) { best_median = median; best = i; }
    }
    if (median_out) *median_out = best_median;
    best = idx[best];
    free(row); free(idx);
    return best;
}

int orbo_predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)ceilf(cvp_logf(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}

float orbo_log_scale_factor(float scale_factor) { return cvp_logf(scale_factor); }

int orbo_is_in_frustum(const float* Tcw, const float* K, float bf, float minX, float maxX, float minY, float maxY,
                       float scale_factor, int nlevels, float viewing_cos_limit, int n, const float* xyz, const float* normal,
                       const float* max_distance, const float* min_distance, unsigned char* in_view, float* proj_xyxr,
                       int* level, float* view_cos)
{
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const float logs = cvp_logf(scale_factor);
    /* mOw = -mRcw.t() * mtcw (src/Frame.cc:284): negate, then products left to right */
    float Ow[3];
    for (int i = 0; i < 3; ++i) {
        float s = (-Tcw[0 * 4 + i]) * Tcw[3];
        s = s + (-Tcw[1 * 4 + i]) * Tcw[7];
        s = s + (-Tcw[2 * 4 + i]) * Tcw[11];
        Ow[i] = s;
    }
    int count = 0;
    for (int p = 0; p < n; ++p) {
        const float* P = xyz + 3 * p;
        in_view[p] = 0;
        float Pc[3];
        for (int r = 0; r < 3; ++r) {                                /* :295 */
            float s = Tcw[4 * r] * P[0];
            s = s + Tcw[4 * r + 1] * P[1];
            s = s + Tcw[4 * r + 2] * P[2];
            Pc[r] = s + Tcw[4 * r + 3];
        }
        if (Pc[2] < 0.0f) continue;                                  /* :301 */
        const float invz = 1.0f / Pc[2];
        const float u = fx * Pc[0] * invz + cx, v = fy * Pc[1] * invz + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const float maxD = 1.2f * max_distance[p], minD = 0.8f * min_distance[p];
        float PO[3];
        double s2 = 0, dot = 0;
        for (int k = 0; k < 3; ++k) { PO[k] = P[k] - Ow[k]; s2 += (double)PO[k] * (double)PO[k]; }
        const float dist = (float)sqrt(s2);                          /* :318 */
        if (dist < minD || dist > maxD) continue;
        for (int k = 0; k < 3; ++k) dot += (double)PO[k] * (double)normal[3 * p + k];
        const float vc = (float)(dot / (double)dist);                /* :324 */
        if (vc < viewing_cos_limit) continue;
        in_view[p] = 1;
        proj_xyxr[3 * p] = u; proj_xyxr[3 * p + 1] = v; proj_xyxr[3 * p + 2] = u - bf * invz;
        level[p] = orbo_predict_scale(max_distance[p], dist, logs, nlevels);
        view_cos[p] = vc;
        ++count;
    }
    return count;
}
