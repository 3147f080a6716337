/*
 * oracle/orb_match_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * Plain-C restatement of the reference's window matchers and stereo row matcher:
 *   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:243-259, 412-422, 348-409
 *   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) src/ORBmatcher.cc:73-157
 *   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, mono) src/ORBmatcher.cc:160-300
 *   ORBmatcher::SearchForInitialization                            src/ORBmatcher.cc:1055-1180
 *   ORBmatcher::ComputeThreeMaxima / RadiusByViewingCos            src/ORBmatcher.cc:1663-1707, 1653-1660
 *   Frame::ComputeStereoMatches                                    src/Frame.cc:513-699
 * Pinned against the reference's unmodified sources (oracle/_ref) by tests/test_matcher_oracle.py.
 * Build with -ffp-contract=off; every float operation is written in the reference's types and order.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GRID_COLS 64   /* FRAME_GRID_COLS, include/Frame.h:38 */
#define GRID_ROWS 48   /* FRAME_GRID_ROWS, include/Frame.h:37 */
#define TH_HIGH 100    /* src/ORBmatcher.cc:37 */
#define TH_LOW 50      /* :38 */
#define HISTO_LENGTH 30 /* :39 */

typedef struct {
    int n;
    const orbo_kp* kps;
    float minX, minY, invW, invH;
    int* cell_start;   /* [GRID_COLS*GRID_ROWS + 1], cell = ix*GRID_ROWS + iy */
    int* cell_items;   /* keypoint indices, insertion (index) order inside a cell */
} grid_t;

/* AssignFeaturesToGrid + PosInGrid, src/Frame.cc:243-259, 412-422 */
static void grid_build(grid_t* g, int n, const orbo_kp* kps, float minX, float maxX, float minY, float maxY)
{
    g->n = n; g->kps = kps; g->minX = minX; g->minY = minY;
    g->invW = (float)GRID_COLS / (maxX - minX);     /* src/Frame.cc:108 */
    g->invH = (float)GRID_ROWS / (maxY - minY);     /* :109 */
    const int nc = GRID_COLS * GRID_ROWS;
    g->cell_start = (int*)calloc((size_t)nc + 1, sizeof(int));
    g->cell_items = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        int px = (int)roundf((kps[i].x - minX) * g->invW);   /* round(), not floor (:414-415) */
        int py = (int)roundf((kps[i].y - minY) * g->invH);
        cell_of[i] = (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) ? -1 : px * GRID_ROWS + py;
        if (cell_of[i] >= 0) g->cell_start[cell_of[i] + 1]++;
    }
    for (int c = 0; c < nc; ++c) g->cell_start[c + 1] += g->cell_start[c];
    int* fill = (int*)malloc(sizeof(int) * (size_t)nc);
    memcpy(fill, g->cell_start, sizeof(int) * (size_t)nc);
    for (int i = 0; i < n; ++i) if (cell_of[i] >= 0) g->cell_items[fill[cell_of[i]]++] = i;
    free(fill); free(cell_of);
}
static void grid_free(grid_t* g) { free(g->cell_start); free(g->cell_items); }

/* GetFeaturesInArea, src/Frame.cc:348-409.  Result order: ix, then iy, then insertion. */
static int grid_query(const grid_t* g, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    int n = 0;
    int a = (int)floorf((x - g->minX - r) * g->invW);
    const int nMinCellX = a > 0 ? a : 0;
    if (nMinCellX >= GRID_COLS) return 0;
    a = (int)ceilf((x - g->minX + r) * g->invW);
    const int nMaxCellX = a < GRID_COLS - 1 ? a : GRID_COLS - 1;
    if (nMaxCellX < 0) return 0;
    a = (int)floorf((y - g->minY - r) * g->invH);
    const int nMinCellY = a > 0 ? a : 0;
    if (nMinCellY >= GRID_ROWS) return 0;
    a = (int)ceilf((y - g->minY + r) * g->invH);
    const int nMaxCellY = a < GRID_ROWS - 1 ? a : GRID_ROWS - 1;
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);       /* :375 */
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = g->cell_start[c]; j < g->cell_start[c + 1]; ++j) {
                const int idx = g->cell_items[j];
                const orbo_kp* kp = &g->kps[idx];
                if (bCheckLevels) {
                    if (kp->octave < minLevel) continue;
                    if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                }
                const float distx = kp->x - x, disty = kp->y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) { if (n < cap) out[n] = idx; ++n; }
            }
        }
    return n;
}

int orbo_features_in_area(int n, const orbo_kp* kps, float minX, float maxX, float minY, float maxY,
                          float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    grid_t g;
    grid_build(&g, n, kps, minX, maxX, minY, maxY);
    int cnt = grid_query(&g, x, y, r, minLevel, maxLevel, out, cap);
    grid_free(&g);
    return cnt;
}

/* The same grid kept across queries (a Frame / KeyFrame builds it once, src/Frame.cc:243-259): orb_fuse_oracle.c */
void* orbo_grid_create(int n, const orbo_kp* kps, float minX, float maxX, float minY, float maxY)
{
    grid_t* g = (grid_t*)malloc(sizeof(grid_t));
    grid_build(g, n, kps, minX, maxX, minY, maxY);
    return g;
}
int orbo_grid_query(const void* g, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    return grid_query((const grid_t*)g, x, y, r, minLevel, maxLevel, out, cap);
}
void orbo_grid_destroy(void* g) { if (g) { grid_free((grid_t*)g); free(g); } }

/* ComputeThreeMaxima, src/ORBmatcher.cc:1663-1707 */
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

typedef struct { int* v; int n, cap; } ivec;
static void ivec_push(ivec* a, int x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (int*)realloc(a->v, sizeof(int) * (size_t)a->cap); }
    a->v[a->n++] = x;
}

/* rotation-histogram bin, src/ORBmatcher.cc:263-268 (factor = 1.0f/HISTO_LENGTH: the famous bug is kept) */
static int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* src/ORBmatcher.cc:1055-1180 */
int orbo_search_for_initialization(int n1, const orbo_kp* kps1, const uint8_t* desc1,
                                   int n2, const orbo_kp* kps2, const uint8_t* desc2,
                                   float minX, float maxX, float minY, float maxY,
                                   float* prev_matched, int* matches12, int windowSize, float nnratio, int checkOri)
{
    int nmatches = 0;
    grid_t g;
    grid_build(&g, n2, kps2, minX, maxX, minY, maxY);
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    ivec hist[HISTO_LENGTH]; memset(hist, 0, sizeof(hist));
    int* vMatchedDistance = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    int* vnMatches21 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; ++i) { vMatchedDistance[i] = INT_MAX; vnMatches21[i] = -1; }
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i1 = 0; i1 < n1; ++i1) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;                                                   /* :1075 */
        const int nc = grid_query(&g, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)windowSize, level1, level1, cand, n2);
        if (nc == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            const int i2 = cand[c];
            const int dist = orbo_descriptor_distance(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;                             /* :1094 */
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if ((float)bestDist < (float)bestDist2 * nnratio) {                     /* :1112 */
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) ivec_push(&hist[rot_bin(kps1[i1].angle, kps2[bestIdx2].angle)], i1);
            }
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = hist[i].n;
        three_maxima(sizes, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j = 0; j < hist[i].n; ++j) {
                const int idx1 = hist[i].v[j];
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }     /* :1163-1167 */
            }
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)
        if (matches12[i1] >= 0) { prev_matched[2 * i1] = kps2[matches12[i1]].x; prev_matched[2 * i1 + 1] = kps2[matches12[i1]].y; }
    for (int i = 0; i < HISTO_LENGTH; ++i) free(hist[i].v);
    free(cand); free(vnMatches21); free(vMatchedDistance);
    grid_free(&g);
    return nmatches;
}

/* src/ORBmatcher.cc:73-157 */
int orbo_search_by_projection_points(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                                     const float* scale, float minX, float maxX, float minY, float maxY,
                                     int nq, const float* proj_xyxr, const int* level, const float* view_cos,
                                     const uint8_t* in_view, const uint8_t* bad, const int* observations,
                                     const uint8_t* qdesc, const int* init_assign, int* assign_out, float th, float nnratio)
{
    int nmatches = 0;
    grid_t g;
    grid_build(&g, n, kps, minX, maxX, minY, maxY);
    for (int k = 0; k < n; ++k) assign_out[k] = init_assign ? init_assign[k] : -1;
    const int bFactor = th != 1.0;
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int iMP = 0; iMP < nq; ++iMP) {
        if (!in_view[iMP]) continue;
        if (bad[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = ((double)view_cos[iMP] > 0.998) ? 2.5f : 4.0f;                    /* :1653-1660 */
        if (bFactor) r *= th;
        const int nc = grid_query(&g, proj_xyxr[3 * iMP], proj_xyxr[3 * iMP + 1], r * scale[nPredictedLevel],
                                  nPredictedLevel - 1, nPredictedLevel, cand, n);
        if (nc == 0) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; ++c) {
            const int idx = cand[c];
            if (assign_out[idx] >= 0 && observations[assign_out[idx]] > 0) continue; /* :115-117 */
            if (u_right && u_right[idx] > 0) {
                const float er = fabsf(proj_xyxr[3 * iMP + 2] - u_right[idx]);
                if (er > r * scale[nPredictedLevel]) continue;
            }
            const int dist = orbo_descriptor_distance(qdesc + (size_t)iMP * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = kps[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue; /* :146-149 */
            assign_out[bestIdx] = iMP;
            nmatches++;
        }
    }
    free(cand);
    grid_free(&g);
    return nmatches;
}

/* float 3x3 * 3x1 + 3x1 the way cv::gemm does it for CV_32F */
static void rt_apply(const float* T /* row-major 4x4 */, const float* x, float* out)
{
    for (int i = 0; i < 3; ++i) {
        float s = T[4 * i] * x[0];
        s = s + T[4 * i + 1] * x[1];
        s = s + T[4 * i + 2] * x[2];
        out[i] = s + T[4 * i + 3];
    }
}

/* src/ORBmatcher.cc:160-300 */
int orbo_search_by_projection_frame(int n_cur, const orbo_kp* kps_cur, const uint8_t* desc_cur, const float* u_right_cur,
                                    int n_last, const orbo_kp* kps_last, const uint8_t* last_mp, const uint8_t* last_outlier,
                                    const float* last_xyz, const uint8_t* last_mp_desc, const int* last_mp_obs,
                                    const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                    const float* scale, float minX, float maxX, float minY, float maxY,
                                    const int* cur_init_obs, int* assign_out, float th, int bMono, float nnratio, int checkOri)
{
    (void)nnratio;
    int nmatches = 0;
    grid_t g;
    grid_build(&g, n_cur, kps_cur, minX, maxX, minY, maxY);
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const float mb = bf / fx;
    /* twc = -Rcw^T * tcw; tlc = Rlw*twc + tlw (:172-181) */
    float twc[3], tlc[3];
    for (int i = 0; i < 3; ++i) {
        float s = (-Tcw_cur[0 * 4 + i]) * Tcw_cur[3];
        s = s + (-Tcw_cur[1 * 4 + i]) * Tcw_cur[7];
        s = s + (-Tcw_cur[2 * 4 + i]) * Tcw_cur[11];
        twc[i] = s;
    }
    rt_apply(Tcw_last, twc, tlc);
    const int bForward = tlc[2] > mb && !bMono;
    const int bBackward = -tlc[2] > mb && !bMono;
    /* per current keypoint: -1 free, else Observations() of the attached point; and who attached it */
    int* obs = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
    for (int k = 0; k < n_cur; ++k) { obs[k] = cur_init_obs ? cur_init_obs[k] : -1; assign_out[k] = (cur_init_obs && cur_init_obs[k] >= 0) ? -2 : -1; }
    ivec hist[HISTO_LENGTH]; memset(hist, 0, sizeof(hist));
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
    for (int i = 0; i < n_last; ++i) {
        if (!last_mp[i]) continue;
        if (last_outlier && last_outlier[i]) continue;
        float xc3[3];
        rt_apply(Tcw_cur, last_xyz + 3 * i, xc3);
        const float xc = xc3[0], yc = xc3[1];
        const float invzc = (float)(1.0 / (double)xc3[2]);                          /* :199 */
        if (invzc < 0) continue;
        const float u = fx * xc * invzc + cx, v = fy * yc * invzc + cy;
        if (u < minX || u > maxX) continue;
        if (v < minY || v > maxY) continue;
        const int nLastOctave = kps_last[i].octave;
        const float radius = th * scale[nLastOctave];
        int nc;
        if (bForward) nc = grid_query(&g, u, v, radius, nLastOctave, -1, cand, n_cur);
        else if (bBackward) nc = grid_query(&g, u, v, radius, 0, nLastOctave, cand, n_cur);
        else nc = grid_query(&g, u, v, radius, nLastOctave - 1, nLastOctave + 1, cand, n_cur);
        if (nc == 0) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; ++c) {
            const int i2 = cand[c];
            if (obs[i2] > 0) continue;                                              /* :234-236 */
            if (u_right_cur && u_right_cur[i2] > 0) {
                const float ur = u - bf * invzc;
                const float er = fabsf(ur - u_right_cur[i2]);
                if (er > radius) continue;
            }
            const int dist = orbo_descriptor_distance(last_mp_desc + (size_t)i * 32, desc_cur + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            assign_out[bestIdx2] = i;
            obs[bestIdx2] = last_mp_obs ? last_mp_obs[i] : 1;
            nmatches++;
            if (checkOri) ivec_push(&hist[rot_bin(kps_last[i].angle, kps_cur[bestIdx2].angle)], bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = hist[i].n;
        three_maxima(sizes, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != i1 && i != i2 && i != i3)
                for (int j = 0; j < hist[i].n; ++j) { assign_out[hist[i].v[j]] = -1; nmatches--; }   /* :286-296 */
    }
    for (int i = 0; i < HISTO_LENGTH; ++i) free(hist[i].v);
    free(cand); free(obs);
    grid_free(&g);
    return nmatches;
}

/* ------------------------------------------------------------------ stereo */
typedef struct { int dist, idx; } distidx;
static int cmp_distidx(const void* a, const void* b)
{
    const distidx* A = (const distidx*)a; const distidx* B = (const distidx*)b;
    if (A->dist != B->dist) return A->dist < B->dist ? -1 : 1;
    return A->idx < B->idx ? -1 : (A->idx > B->idx);
}

/* Frame::ComputeStereoMatches, src/Frame.cc:513-699.  eL / eR: extractors whose last orbo_extract call
 * produced the left / right keypoints (their bordered pyramids are read, like mvImagePyramid). */
int orbo_stereo_matches(const orbo_extractor* eL, const orbo_extractor* eR,
                        int nl, const orbo_kp* kps_l, const uint8_t* desc_l,
                        int nr, const orbo_kp* kps_r, const uint8_t* desc_r,
                        float bf, float fx, float* u_right, float* depth)
{
    const int nlevels = orbo_levels(eL);
    float scale[32], inv_scale[32];
    orbo_tables(eL, scale, inv_scale, NULL, NULL, NULL, NULL, NULL);
    (void)nlevels;
    for (int i = 0; i < nl; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    int w0, nRows;
    orbo_stage_level_size(eL, 0, &w0, &nRows);
    /* row table (:524-540) */
    ivec* rows = (ivec*)calloc((size_t)nRows, sizeof(ivec));
    for (int iR = 0; iR < nr; ++iR) {
        const float kpY = kps_r[iR].y;
        const float r = 2.0f * scale[kps_r[iR].octave];
        const int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi) if (yi >= 0 && yi < nRows) ivec_push(&rows[yi], iR);
    }
    const float mb = bf / fx;
    const float minZ = mb, minD = 0, maxD = bf / minZ;
    distidx* vDistIdx = (distidx*)malloc(sizeof(distidx) * (size_t)(nl > 0 ? nl : 1));
    int nd = 0;
    /* bordered pyramid levels of both images, fetched once */
    uint8_t* pl[32]; uint8_t* pr[32]; int lw[32], lh[32];
    for (int l = 0; l < nlevels; ++l) {
        orbo_stage_level_size(eL, l, &lw[l], &lh[l]);
        const size_t st = (size_t)(lw[l] + 38);
        pl[l] = (uint8_t*)malloc(st * (size_t)(lh[l] + 38));
        pr[l] = (uint8_t*)malloc(st * (size_t)(lh[l] + 38));
        orbo_stage_pyramid(eL, l, 1, pl[l], st);
        orbo_stage_pyramid(eR, l, 1, pr[l], st);
    }
    for (int iL = 0; iL < nl; ++iL) {
        const int levelL = kps_l[iL].octave;
        const float vL = kps_l[iL].y, uL = kps_l[iL].x;
        const ivec* cands = &rows[(int)vL];                                          /* :559 */
        if (cands->n == 0) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        for (int iC = 0; iC < cands->n; ++iC) {
            const int iR = cands->v[iC];
            if (kps_r[iR].octave < levelL - 1 || kps_r[iR].octave > levelL + 1) continue;
            const float uR = kps_r[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orbo_descriptor_distance(desc_l + (size_t)iL * 32, desc_r + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kps_r[bestIdxR].x;
            const float scaleFactor = inv_scale[levelL];
            const float scaleduL = roundf(uL * scaleFactor), scaledvL = roundf(vL * scaleFactor), scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const size_t st = (size_t)(lw[levelL] + 38);
            const uint8_t* IL = pl[levelL] + 19 * st + 19;     /* ROI origin inside the bordered buffer */
            const uint8_t* IR = pr[levelL] + 19 * st + 19;
            const int r0 = (int)(scaledvL - w), cL0 = (int)(scaleduL - w);
            const float cL = (float)IL[(size_t)(r0 + w) * st + cL0 + w];
            int bestDistS = INT_MAX, bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;      /* :624-625 */
            if (iniu < 0 || endu >= (float)lw[levelL]) continue;
            for (int incR = -L; incR <= L; ++incR) {
                const int cR0 = (int)(scaleduR0 + incR - w);
                const float cR = (float)IR[(size_t)(r0 + w) * st + cR0 + w];
                double sum = 0;                                                      /* cv::norm(IL, IR, NORM_L1) accumulates in double */
                for (int yy = 0; yy < 2 * w + 1; ++yy)
                    for (int xx = 0; xx < 2 * w + 1; ++xx) {
                        const float a = (float)IL[(size_t)(r0 + yy) * st + cL0 + xx] - cL;
                        const float b = (float)IR[(size_t)(r0 + yy) * st + cR0 + xx] - cR;
                        sum += fabs((double)a - (double)b);
                    }
                const float dist = (float)sum;
                if (dist < (float)bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = uL - bestuR;
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = (float)0.01; bestuR = (float)((double)uL - 0.01); }
                depth[iL] = bf / disparity;
                u_right[iL] = bestuR;
                vDistIdx[nd].dist = bestDistS; vDistIdx[nd].idx = iL; ++nd;
            }
        }
    }
    if (nd > 0) {                                                                    /* :685-698 (UB in the reference when nd == 0) */
        qsort(vDistIdx, (size_t)nd, sizeof(distidx), cmp_distidx);
        const float median = (float)vDistIdx[nd / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nd - 1; i >= 0; --i) {
            if ((float)vDistIdx[i].dist < thDist) break;
            u_right[vDistIdx[i].idx] = -1; depth[vDistIdx[i].idx] = -1;
        }
    }
    for (int l = 0; l < nlevels; ++l) { free(pl[l]); free(pr[l]); }
    for (int i = 0; i < nRows; ++i) free(rows[i].v);
    free(rows); free(vDistIdx);
    return nd;
}

/* ORBmatcher::SearchByBoW, src/ORBmatcher.cc:552-697 (KeyFrame -> Frame; valid2 == NULL, strict == 0) and
 * :700-832 (KeyFrame -> KeyFrame; valid2 = "has a map point that is not bad", strict == 1: bestDist1 < TH_LOW).
 * The DBoW2::FeatureVector of each side in CSR form: node_id [nn] ascending, node_off [nn + 1], feat [node_off[nn]].
 * valid1 [n1]: key-frame-1 feature has a map point that is not bad.  match12 [n1] = matched feature of side 2 or -1,
 * match21 [n2] (or NULL) the inverse.  Returns nmatches. */
int orbo_search_by_bow(int n1, const orbo_kp* kps1, const uint8_t* desc1, const uint8_t* valid1,
                       int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                       int n2, const orbo_kp* kps2, const uint8_t* desc2, const uint8_t* valid2,
                       int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                       float nnratio, int checkOri, int strict, int* match12, int* match21)
{
    int nmatches = 0;
    uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
    ivec hist[HISTO_LENGTH];
    memset(hist, 0, sizeof(hist));
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (match21) for (int i = 0; i < n2; ++i) match21[i] = -1;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (node_id1[a] == node_id2[b]) {
            for (int p1 = node_off1[a]; p1 < node_off1[a + 1]; ++p1) {
                const int idx1 = feat1[p1];
                if (!valid1[idx1]) continue;
                int bestDist1 = 256, bestDist2 = 256, bestIdx2 = -1;
                for (int p2 = node_off2[b]; p2 < node_off2[b + 1]; ++p2) {
                    const int idx2 = feat2[p2];
                    if (matched2[idx2] || (valid2 && !valid2[idx2])) continue;
                    const int dist = orbo_descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (strict ? bestDist1 < TH_LOW : bestDist1 <= TH_LOW) {                        /* :768 / :618 */
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        match12[idx1] = bestIdx2;
                        matched2[bestIdx2] = 1;
                        if (checkOri) ivec_push(&hist[rot_bin(kps1[idx1].angle, kps2[bestIdx2].angle)], idx1);
                        nmatches++;
                    }
                }
            }
            ++a; ++b;
        } else if (node_id1[a] < node_id2[b]) {
            while (a < nn1 && node_id1[a] < node_id2[b]) ++a;                                   /* lower_bound */
        } else {
            while (b < nn2 && node_id2[b] < node_id1[a]) ++b;
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = hist[i].n;
        three_maxima(sizes, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j = 0; j < hist[i].n; ++j) { match12[hist[i].v[j]] = -1; nmatches--; }
        }
    }
    if (match21) for (int i = 0; i < n1; ++i) if (match12[i] >= 0) match21[match12[i]] = i;
    for (int i = 0; i < HISTO_LENGTH; ++i) free(hist[i].v);
    free(matched2);
    return nmatches;
}

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:1183-1361, with CheckDistEpipolarLine (:1636-1650).
 * has_mp1 / has_mp2: the key frame's feature already has a map point (pKF->GetMapPoint(idx) != NULL); u_right1 / u_right2 =
 * mvuRight (NULL = monocular: every entry negative); F12 row-major 3x3; epipole = (ex, ey) of :1195-1196 (the caller projects
 * KF1's camera centre into KF2); scale = pKF2->mvScaleFactors, sigma2 = pKF2->mvLevelSigma2.  A candidate replaces the
 * running best when its distance is <= the best so far (:1248: `dist>bestDist` continues), i.e. the LAST of equally close
 * candidates that pass the epipolar tests wins.  The reference declares vbMatched2 (:1205) and tests it (:1254) but never sets
 * it, so no feature of key frame 1 depends on another one.  match12 [n1] = vMatches12 (vMatchedPairs lists its entries >= 0
 * in order). */
int orbo_search_for_triangulation(int n1, const orbo_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1,
                                  int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                                  int n2, const orbo_kp* kps2, const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2,
                                  int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                                  const float* F12, const float* epipole, const float* scale, const float* sigma2,
                                  int onlyStereo, int checkOri, int* match12)
{
    int nmatches = 0;
    uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
    ivec hist[HISTO_LENGTH];
    memset(hist, 0, sizeof(hist));
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    const float ex = epipole[0], ey = epipole[1];
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (node_id1[a] == node_id2[b]) {
            for (int p1 = node_off1[a]; p1 < node_off1[a + 1]; ++p1) {
                const int idx1 = feat1[p1];
                if (has_mp1[idx1]) continue;                                                    /* :1222-1225 */
                const int bStereo1 = u_right1 && u_right1[idx1] >= 0;
                if (onlyStereo && !bStereo1) continue;
                const orbo_kp* kp1 = &kps1[idx1];
                /* epipolar line of kp1 in image 2, :1640-1642 */
                const float la = kp1->x * F12[0] + kp1->y * F12[3] + F12[6];
                const float lb = kp1->x * F12[1] + kp1->y * F12[4] + F12[7];
                const float lc = kp1->x * F12[2] + kp1->y * F12[5] + F12[8];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int p2 = node_off2[b]; p2 < node_off2[b + 1]; ++p2) {
                    const int idx2 = feat2[p2];
                    if (matched2[idx2] || has_mp2[idx2]) continue;                              /* :1245-1246 */
                    const int bStereo2 = u_right2 && u_right2[idx2] >= 0;
                    if (onlyStereo && !bStereo2) continue;
                    const int dist = orbo_descriptor_distance(desc1 + (size_t)idx1 * 32, desc2 + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const orbo_kp* kp2 = &kps2[idx2];
                    if (!bStereo1 && !bStereo2) {                                               /* too close to the epipole, :1263-1269 */
                        const float distex = ex - kp2->x, distey = ey - kp2->y;
                        if (distex * distex + distey * distey < 100 * scale[kp2->octave]) continue;
                    }
                    const float num = la * kp2->x + lb * kp2->y + lc;                           /* CheckDistEpipolarLine */
                    const float den = la * la + lb * lb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * sigma2[kp2->octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;      /* vbMatched2 is NOT set here (:1289-1292): several idx1 may share an idx2 */
                    nmatches++;
                    if (checkOri) ivec_push(&hist[rot_bin(kp1->angle, kps2[bestIdx2].angle)], idx1);
                }
            }
            ++a; ++b;
        } else if (node_id1[a] < node_id2[b]) {
            while (a < nn1 && node_id1[a] < node_id2[b]) ++a;
        } else {
            while (b < nn2 && node_id2[b] < node_id1[a]) ++b;
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], i1, i2, i3;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = hist[i].n;
        three_maxima(sizes, HISTO_LENGTH, &i1, &i2, &i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j = 0; j < hist[i].n; ++j) { match12[hist[i].v[j]] = -1; nmatches--; }
        }
    }
    for (int i = 0; i < HISTO_LENGTH; ++i) free(hist[i].v);
    free(matched2);
    return nmatches;
}

/* Frame::UndistortKeyPoints, src/Frame.cc:436-468: cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) restated from
 * OpenCV 4.13 (third-party arithmetic, not under /root/reference; pinned against cv2 4.13.0 live and against
 * tests/golden/cv2_undistort.npz by tests/test_frame_steps.py): double arithmetic, normalise, five iterations of the inverse
 * distortion model (TermCriteria(MAX_ITER, 5, 0.01): the count only), project with the new camera matrix, narrow to float.
 * K = fx, fy, cx, cy; dist = k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]].  Only pt.x / pt.y change. */
void orbo_undistort_keypoints(int n, const orbo_kp* kps, const float* K, const float* dist, int n_dist, orbo_kp* out)
{
    if (n_dist == 0 || dist[0] == 0.0f) { for (int i = 0; i < n; ++i) out[i] = kps[i]; return; }   /* :438-442 */
    double k[12] = { 0 };
    for (int i = 0; i < n_dist && i < 12; ++i) k[i] = (double)dist[i];
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3], ifx = 1.0 / fx, ify = 1.0 / fy;
    for (int i = 0; i < n; ++i) {
        double x = ((double)kps[i].x - cx) * ifx, y = ((double)kps[i].y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        out[i] = kps[i];
        out[i].x = (float)(fx * x + cx);
        out[i].y = (float)(fy * y + cy);
    }
}

/* Frame::ComputeStereoFromRGBD, src/Frame.cc:702-727.  depth: h x w floats, row-major. */
void orbo_stereo_from_rgbd(int n, const orbo_kp* kps, const orbo_kp* kps_un, const float* depth, int w, int h, float bf,
                           float* u_right, float* depth_out)
{
    for (int i = 0; i < n; ++i) {
        u_right[i] = -1.0f; depth_out[i] = -1.0f;
        const int v = (int)kps[i].y, u = (int)kps[i].x;      /* imDepth.at<float>(v, u) with float arguments */
        if (u < 0 || u >= w || v < 0 || v >= h) continue;    /* (the reference would read out of bounds) */
        const float d = depth[(size_t)v * (size_t)w + (size_t)u];
        if (d > 0) { depth_out[i] = d; u_right[i] = kps_un[i].x - bf / d; }
    }
}
