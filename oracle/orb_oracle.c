/*
 * oracle/orb_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see orb_oracle.h).
 *
 * Plain-C restatement of the reference's ORB extractor.  Every function cites the
 * reference lines it follows (paths relative to /root/reference).  Float arithmetic is
 * written operation by operation in the reference's own types and order; build with
 * -ffp-contract=off.
 */
#define _POSIX_C_SOURCE 200809L
#include "orb_oracle.h"
#include "cv_prims.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PATCH_SIZE 31       /* src/ORBextractor.cc:72 */
#define HALF_PATCH_SIZE 15  /* :73 */
#define EDGE_THRESHOLD 19   /* :74 */
#define MAX_LEVELS 32

static const signed char k_pattern[1024] = {
#include "orb_pattern.inc"
};

typedef struct {
    int w, h;
    size_t step;      /* of the bordered buffer */
    uint8_t* buf;     /* (w+38) x (h+38), ROI at (19,19) */
    uint8_t* blur;    /* w x h, pitch w */
    orbo_cand* cand; int ncand, cand_cap;
    orbo_kp* kps; int nkps, kps_cap;
} level_t;

struct orbo_extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor; /* include/ORBextractor.h:99: the member is a double */
    float scale[MAX_LEVELS], inv_scale[MAX_LEVELS], sigma2[MAX_LEVELS], inv_sigma2[MAX_LEVELS];
    int per_level[MAX_LEVELS];
    int umax[HALF_PATCH_SIZE + 1];
    level_t lv[MAX_LEVELS];
};

static inline uint8_t* roi(const level_t* L) { return L->buf + EDGE_THRESHOLD * L->step + EDGE_THRESHOLD; }
static inline int ifloor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int iceil_f(float v) { int i = (int)v; return i + (i < v); }

/* src/ORBextractor.cc:498-559 */
orbo_extractor* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    if (nlevels < 1 || nlevels > MAX_LEVELS) return NULL;
    orbo_extractor* e = (orbo_extractor*)calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniThFAST; e->minTh = minThFAST;
    e->scaleFactor = (double)scaleFactor;
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        e->scale[i] = (float)((double)e->scale[i - 1] * e->scaleFactor);   /* :506 float*double */
        e->sigma2[i] = e->scale[i] * e->scale[i];                           /* :507 */
    }
    for (int i = 0; i < nlevels; ++i) {
        e->inv_scale[i] = 1.0f / e->scale[i];                               /* :513 */
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];                             /* :514 */
    }
    float factor = (float)(1.0 / e->scaleFactor);                           /* :520 1.0f/double */
    float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels)); /* :522-523 */
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e->per_level[l] = cvp_round((double)nDesired);                      /* :527 cvRound(float) */
        sum += e->per_level[l];
        nDesired *= factor;
    }
    e->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;  /* :531 */

    /* :544-558 umax */
    int v, v0;
    int vmax = ifloor_f((float)HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = iceil_f((float)HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cvp_round(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

static void free_levels(orbo_extractor* e)
{
    for (int l = 0; l < MAX_LEVELS; ++l) {
        free(e->lv[l].buf); free(e->lv[l].blur); free(e->lv[l].cand); free(e->lv[l].kps);
        memset(&e->lv[l], 0, sizeof(level_t));
    }
}

void orbo_destroy(orbo_extractor* e)
{
    if (!e) return;
    free_levels(e);
    free(e);
}

int orbo_levels(const orbo_extractor* e) { return e->nlevels; }

void orbo_tables(const orbo_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                 int* per_level, int* umax16, int* pattern1024)
{
    for (int i = 0; i < e->nlevels; ++i) {
        if (scale) scale[i] = e->scale[i];
        if (inv_scale) inv_scale[i] = e->inv_scale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = e->inv_sigma2[i];
        if (per_level) per_level[i] = e->per_level[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
    if (pattern1024) for (int i = 0; i < 1024; ++i) pattern1024[i] = k_pattern[i];
}

/* ------------------------------------------------------------------ pyramid */
/* src/ORBextractor.cc:1153-1180 */
static void compute_pyramid(orbo_extractor* e, const uint8_t* img, int w, int h, size_t step)
{
    for (int l = 0; l < e->nlevels; ++l) {
        level_t* L = &e->lv[l];
        float scale = e->inv_scale[l];
        int lw = cvp_round((double)((float)w * scale));   /* :1158 cvRound(float) */
        int lh = cvp_round((double)((float)h * scale));
        L->w = lw; L->h = lh;
        L->step = (size_t)(lw + 2 * EDGE_THRESHOLD);
        free(L->buf);
        L->buf = (uint8_t*)malloc(L->step * (size_t)(lh + 2 * EDGE_THRESHOLD));
        if (l != 0) {
            const level_t* P = &e->lv[l - 1];
            cvp_resize_linear_8u(roi(P), P->w, P->h, P->step, roi(L), lw, lh, L->step);          /* :1166 */
            cvp_border_reflect101(roi(L), lw, lh, L->step, L->buf, L->step, EDGE_THRESHOLD);       /* :1168 */
        } else {
            cvp_border_reflect101(img, w, h, step, L->buf, L->step, EDGE_THRESHOLD);               /* :1173 */
        }
    }
}

/* ------------------------------------------------------------------ quadtree */
/* src/ORBextractor.cc:436-495 (DivideNode) and :562-792 (DistributeOctTree).  The list is a
 * doubly linked list over an index pool; "seq" is the creation counter that stands in
 * for the node address in the sort of :711 (bump-arena order). */
typedef struct {
    int ulx, urx, uly, bry;
    int* keys; int nkeys;
    int no_more;
    int prev, next;
    int seq;
} qnode;

typedef struct {
    qnode* nodes; int nnodes, cap;
    int head, tail, size;
    int seq;
    int** chunks; int nchunks, chunks_cap; /* key-index storage, freed together */
    size_t chunk_used, chunk_cap;
} qlist;

static int* pool_alloc(qlist* q, int n)
{
    if (q->nchunks == 0 || q->chunk_used + (size_t)n > q->chunk_cap) {
        if (q->nchunks == q->chunks_cap) {
            q->chunks_cap = q->chunks_cap ? q->chunks_cap * 2 : 16;
            q->chunks = (int**)realloc(q->chunks, sizeof(int*) * (size_t)q->chunks_cap);
        }
        q->chunk_cap = (size_t)n * 8 > 65536 ? (size_t)n * 8 : 65536;
        q->chunks[q->nchunks++] = (int*)malloc(sizeof(int) * q->chunk_cap);
        q->chunk_used = 0;
    }
    int* p = q->chunks[q->nchunks - 1] + q->chunk_used;
    q->chunk_used += (size_t)n;
    return p;
}

static int node_new(qlist* q)
{
    if (q->nnodes == q->cap) {
        q->cap = q->cap ? q->cap * 2 : 1024;
        q->nodes = (qnode*)realloc(q->nodes, sizeof(qnode) * (size_t)q->cap);
    }
    qnode* n = &q->nodes[q->nnodes];
    memset(n, 0, sizeof(*n));
    n->prev = n->next = -1;
    n->seq = q->seq++;
    return q->nnodes++;
}
static void list_push_front(qlist* q, int i)
{
    q->nodes[i].prev = -1; q->nodes[i].next = q->head;
    if (q->head >= 0) q->nodes[q->head].prev = i; else q->tail = i;
    q->head = i; q->size++;
}
static void list_push_back(qlist* q, int i)
{
    q->nodes[i].next = -1; q->nodes[i].prev = q->tail;
    if (q->tail >= 0) q->nodes[q->tail].next = i; else q->head = i;
    q->tail = i; q->size++;
}
static int list_erase(qlist* q, int i) /* returns next */
{
    int p = q->nodes[i].prev, n = q->nodes[i].next;
    if (p >= 0) q->nodes[p].next = n; else q->head = n;
    if (n >= 0) q->nodes[n].prev = p; else q->tail = p;
    q->size--;
    return n;
}

/* DivideNode :436-495: children c[0..3] = n1..n4 */
static void divide(qlist* q, int parent, const orbo_cand* cand, int c[4])
{
    for (int k = 0; k < 4; ++k) c[k] = node_new(q); /* may realloc: take the parent by index afterwards */
    qnode* P = &q->nodes[parent];
    const int halfX = (int)ceilf((float)(P->urx - P->ulx) / 2);
    const int halfY = (int)ceilf((float)(P->bry - P->uly) / 2);
    qnode* n1 = &q->nodes[c[0]]; qnode* n2 = &q->nodes[c[1]]; qnode* n3 = &q->nodes[c[2]]; qnode* n4 = &q->nodes[c[3]];
    n1->ulx = P->ulx;         n1->urx = P->ulx + halfX; n1->uly = P->uly;         n1->bry = P->uly + halfY;
    n2->ulx = P->ulx + halfX; n2->urx = P->urx;         n2->uly = P->uly;         n2->bry = P->uly + halfY;
    n3->ulx = P->ulx;         n3->urx = P->ulx + halfX; n3->uly = P->uly + halfY; n3->bry = P->bry;
    n4->ulx = P->ulx + halfX; n4->urx = P->urx;         n4->uly = P->uly + halfY; n4->bry = P->bry;
    for (int k = 0; k < 4; ++k) { q->nodes[c[k]].keys = pool_alloc(q, P->nkeys); q->nodes[c[k]].nkeys = 0; }
    const float midx = (float)n1->urx, midy = (float)n1->bry;
    for (int i = 0; i < P->nkeys; ++i) {
        const int ki = P->keys[i];
        const float x = (float)cand[ki].x, y = (float)cand[ki].y;
        qnode* t;
        if (x < midx) t = (y < midy) ? n1 : n3;
        else t = (y < midy) ? n2 : n4;
        t->keys[t->nkeys++] = ki;
    }
    for (int k = 0; k < 4; ++k) if (q->nodes[c[k]].nkeys == 1) q->nodes[c[k]].no_more = 1;
}

typedef struct { int size, seq, node; } szptr;
static int cmp_szptr(const void* a, const void* b)
{
    const szptr* A = (const szptr*)a; const szptr* B = (const szptr*)b;
    if (A->size != B->size) return A->size < B->size ? -1 : 1;
    return A->seq < B->seq ? -1 : (A->seq > B->seq);
}

int orbo_distribute(const orbo_cand* cand, int n, int minX, int maxX, int minY, int maxY, int N,
                    int* out_idx, int cap)
{
    qlist q; memset(&q, 0, sizeof(q)); q.head = q.tail = -1;

    const int nIni = (int)roundf((float)(maxX - minX) / (float)(maxY - minY));  /* :567 */
    const float hX = (float)(maxX - minX) / (float)nIni;                        /* :568 */
    int* roots = (int*)malloc(sizeof(int) * (size_t)(nIni > 0 ? nIni : 1));
    for (int i = 0; i < nIni; ++i) {                                            /* :575-589 */
        int r = node_new(&q);
        qnode* R = &q.nodes[r];
        R->ulx = (int)(hX * (float)i); R->urx = (int)(hX * (float)(i + 1));
        R->uly = 0; R->bry = maxY - minY;
        R->keys = pool_alloc(&q, n); R->nkeys = 0;
        list_push_back(&q, r);
        roots[i] = r;
    }
    for (int i = 0; i < n; ++i) {                                               /* :593-597 */
        qnode* R = &q.nodes[roots[(size_t)((float)cand[i].x / hX)]];
        R->keys[R->nkeys++] = i;
    }
    for (int it = q.head; it >= 0;) {                                           /* :600-615 */
        qnode* nd = &q.nodes[it];
        if (nd->nkeys == 1) { nd->no_more = 1; it = nd->next; }
        else if (nd->nkeys == 0) it = list_erase(&q, it);
        else it = nd->next;
    }

    int finish = 0;
    szptr* vec = (szptr*)malloc(sizeof(szptr) * (size_t)(4 * (n + nIni) + 16));
    szptr* prev = (szptr*)malloc(sizeof(szptr) * (size_t)(4 * (n + nIni) + 16));
    int nvec = 0;
    while (!finish) {                                                           /* :624 */
        int prevSize = q.size;
        int nToExpand = 0;
        nvec = 0;
        for (int it = q.head; it >= 0;) {                                       /* :636-692 */
            if (q.nodes[it].no_more) { it = q.nodes[it].next; continue; }
            int c[4];
            divide(&q, it, cand, c);
            for (int k = 0; k < 4; ++k) {
                if (q.nodes[c[k]].nkeys > 0) {
                    list_push_front(&q, c[k]);
                    if (q.nodes[c[k]].nkeys > 1) {
                        nToExpand++;
                        vec[nvec].size = q.nodes[c[k]].nkeys; vec[nvec].seq = q.nodes[c[k]].seq; vec[nvec].node = c[k];
                        nvec++;
                    }
                }
            }
            it = list_erase(&q, it);
        }
        if (q.size >= N || q.size == prevSize) {                                /* :696 */
            finish = 1;
        } else if (q.size + nToExpand * 3 > N) {                                /* :702 */
            while (!finish) {
                prevSize = q.size;
                int nprev = nvec;
                memcpy(prev, vec, sizeof(szptr) * (size_t)nvec);
                nvec = 0;
                qsort(prev, (size_t)nprev, sizeof(szptr), cmp_szptr);           /* :711 (seq == address order) */
                for (int j = nprev - 1; j >= 0; --j) {                          /* :713 */
                    int c[4];
                    divide(&q, prev[j].node, cand, c);
                    for (int k = 0; k < 4; ++k) {
                        if (q.nodes[c[k]].nkeys > 0) {
                            list_push_front(&q, c[k]);
                            if (q.nodes[c[k]].nkeys > 1) {
                                vec[nvec].size = q.nodes[c[k]].nkeys; vec[nvec].seq = q.nodes[c[k]].seq; vec[nvec].node = c[k];
                                nvec++;
                            }
                        }
                    }
                    list_erase(&q, prev[j].node);
                    if (q.size >= N) break;                                     /* :759 */
                }
                if (q.size >= N || q.size == prevSize) finish = 1;              /* :763 */
            }
        }
    }

    int nout = 0;
    for (int it = q.head; it >= 0; it = q.nodes[it].next) {                     /* :773-789 */
        const qnode* nd = &q.nodes[it];
        int best = nd->keys[0];
        int maxResp = cand[best].score;
        for (int k = 1; k < nd->nkeys; ++k)
            if (cand[nd->keys[k]].score > maxResp) { best = nd->keys[k]; maxResp = cand[best].score; }
        if (nout < cap) out_idx[nout] = best;
        ++nout;
    }
    for (int i = 0; i < q.nchunks; ++i) free(q.chunks[i]);
    free(q.chunks); free(prev); free(vec); free(roots); free(q.nodes);
    return nout;
}

/* ------------------------------------------------------------------ keypoints */
/* IC_Angle, src/ORBextractor.cc:78-105 */
static float ic_angle(const uint8_t* center, int step, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return cvp_fast_atan2((float)m_01, (float)m_10);
}

static void cand_push(level_t* L, int x, int y, int score)
{
    if (L->ncand == L->cand_cap) {
        L->cand_cap = L->cand_cap ? L->cand_cap * 2 : 4096;
        L->cand = (orbo_cand*)realloc(L->cand, sizeof(orbo_cand) * (size_t)L->cand_cap);
    }
    L->cand[L->ncand].x = x; L->cand[L->ncand].y = y; L->cand[L->ncand].score = score;
    L->ncand++;
}

/* src/ORBextractor.cc:795-902 */
static void compute_keypoints_octtree(orbo_extractor* e)
{
    const float W = 30;
    for (int level = 0; level < e->nlevels; ++level) {
        level_t* L = &e->lv[level];
        L->ncand = 0; L->nkps = 0;
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = L->w - EDGE_THRESHOLD + 3, maxBorderY = L->h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);                 /* :816-817 */
        if (nCols > 0 && nRows > 0) {
            const int wCell = (int)ceilf(width / (float)nCols), hCell = (int)ceilf(height / (float)nRows); /* :818-819 */
            cvp_corner* cell = (cvp_corner*)malloc(sizeof(cvp_corner) * 4096);
            int cell_cap = 4096;
            for (int i = 0; i < nRows; ++i) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + (float)hCell + 6;
                if (iniY >= (float)(maxBorderY - 3)) continue;                         /* :830 */
                if (maxY > (float)maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; ++j) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + (float)wCell + 6;
                    if (iniX >= (float)(maxBorderX - 6)) continue;                     /* :845 */
                    if (maxX > (float)maxBorderX) maxX = (float)maxBorderX;
                    const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                    const uint8_t* p = roi(L) + (size_t)y0 * L->step + x0;
                    int need = ((x1 - x0 + 1) / 2) * ((y1 - y0 + 1) / 2) + 16;
                    if (need > cell_cap) { cell_cap = need; cell = (cvp_corner*)realloc(cell, sizeof(cvp_corner) * (size_t)cell_cap); }
                    int n = cvp_fast9_nms(p, x1 - x0, y1 - y0, L->step, e->iniTh, cell, cell_cap);      /* :853 */
                    if (n == 0) n = cvp_fast9_nms(p, x1 - x0, y1 - y0, L->step, e->minTh, cell, cell_cap); /* :857-861 */
                    for (int k = 0; k < n; ++k)
                        cand_push(L, cell[k].x + j * wCell, cell[k].y + i * hCell, cell[k].score);     /* :868-870 */
                }
            }
            free(cell);
        }
        /* :882 DistributeOctTree, :886-896 border offset, octave, size */
        int cap = L->ncand > 0 ? L->ncand : 1;
        int* idx = (int*)malloc(sizeof(int) * (size_t)cap);
        int nk = 0;
        nk = orbo_distribute(L->cand, L->ncand, minBorderX, maxBorderX, minBorderY, maxBorderY, e->per_level[level], idx, cap);
        if (nk > L->kps_cap) { L->kps_cap = nk; L->kps = (orbo_kp*)realloc(L->kps, sizeof(orbo_kp) * (size_t)nk); }
        const int scaledPatchSize = (int)((float)PATCH_SIZE * e->scale[level]);        /* :886 */
        for (int k = 0; k < nk; ++k) {
            orbo_kp* kp = &L->kps[k];
            kp->x = (float)L->cand[idx[k]].x + (float)minBorderX;
            kp->y = (float)L->cand[idx[k]].y + (float)minBorderY;
            kp->size = (float)scaledPatchSize;
            kp->angle = -1.f;
            kp->response = (float)L->cand[idx[k]].score;
            kp->octave = level;
            kp->class_id = -1;
        }
        L->nkps = nk;
        free(idx);
    }
    for (int level = 0; level < e->nlevels; ++level) {                                 /* :900-901 */
        level_t* L = &e->lv[level];
        for (int k = 0; k < L->nkps; ++k) {
            const uint8_t* c = roi(L) + (size_t)cvp_round((double)L->kps[k].y) * L->step + cvp_round((double)L->kps[k].x);
            L->kps[k].angle = ic_angle(c, (int)L->step, e->umax);
        }
    }
}

/* computeOrbDescriptor, src/ORBextractor.cc:117-161 */
static void orb_descriptor(const orbo_kp* kpt, const uint8_t* img, int step, uint8_t* desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);   /* :117 */
    float angle = (float)kpt->angle * factorPI;
    float a = cvp_cosf(angle), b = cvp_sinf(angle);                              /* :125 libm cosf/sinf */
    const uint8_t* center = img + (size_t)cvp_round((double)kpt->y) * (size_t)step + cvp_round((double)kpt->x);
    const signed char* pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; ++j) {
            const float px = (float)pat[4 * j], py = (float)pat[4 * j + 1];
            const float qx = (float)pat[4 * j + 2], qy = (float)pat[4 * j + 3];
            /* GET_VALUE :132-134: row = cvRound(x*b + y*a), col = cvRound(x*a - y*b), float ops */
            int t0 = center[cvp_round((double)(px * b + py * a)) * step + cvp_round((double)(px * a - py * b))];
            int t1 = center[cvp_round((double)(qx * b + qy * a)) * step + cvp_round((double)(qx * a - qy * b))];
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

static int shape_supported(const orbo_extractor* e, int w, int h)
{
    /* every level must leave a non-degenerate FAST area: the reference divides by
     * (maxBorderY-minBorderY) and by nIni = round(W/H) (src/ORBextractor.cc:567-568) */
    for (int l = 0; l < e->nlevels; ++l) {
        int lw = cvp_round((double)((float)w * e->inv_scale[l]));
        int lh = cvp_round((double)((float)h * e->inv_scale[l]));
        int W = lw - 32, H = lh - 32;
        if (W <= 0 || H <= 0) return 0;
        if ((int)roundf((float)W / (float)H) < 1) return 0;
    }
    return 1;
}

/* ORBextractor::operator(), src/ORBextractor.cc:1084-1150 */
int orbo_extract(orbo_extractor* e, const uint8_t* img, int w, int h, size_t step,
                 orbo_kp* kps, uint8_t* desc, int cap)
{
    if (!img || w <= 0 || h <= 0) return -1;                                     /* :1087 */
    if (!shape_supported(e, w, h)) return -2;
    compute_pyramid(e, img, w, h, step);                                         /* :1094 */
    compute_keypoints_octtree(e);                                                /* :1097 */
    int n = 0;
    for (int level = 0; level < e->nlevels; ++level) {
        level_t* L = &e->lv[level];
        free(L->blur); L->blur = NULL;
        if (L->nkps == 0) continue;                                              /* :1123 */
        L->blur = (uint8_t*)malloc((size_t)L->w * (size_t)L->h);
        cvp_gaussian7x7_s2(roi(L), L->w, L->h, L->step, L->blur, (size_t)L->w);  /* :1129-1130 on a clone of the ROI */
        const float scale = e->scale[level];
        for (int k = 0; k < L->nkps; ++k, ++n) {
            if (n >= cap) continue;
            if (desc) orb_descriptor(&L->kps[k], L->blur, L->w, desc + (size_t)n * 32); /* :1134 */
            if (kps) {
                kps[n] = L->kps[k];
                if (level != 0) { kps[n].x = kps[n].x * scale; kps[n].y = kps[n].y * scale; } /* :1140-1146 */
            }
        }
    }
    return n;
}

int orbo_stage_level_size(const orbo_extractor* e, int level, int* w, int* h)
{
    if (level < 0 || level >= e->nlevels || !e->lv[level].buf) return -1;
    if (w) *w = e->lv[level].w;
    if (h) *h = e->lv[level].h;
    return 0;
}
int orbo_stage_pyramid(const orbo_extractor* e, int level, int with_border, uint8_t* dst, size_t dst_step)
{
    if (level < 0 || level >= e->nlevels || !e->lv[level].buf) return -1;
    const level_t* L = &e->lv[level];
    const int b = with_border ? EDGE_THRESHOLD : 0;
    for (int r = -b; r < L->h + b; ++r)
        memcpy(dst + (size_t)(r + b) * dst_step, roi(L) + (ptrdiff_t)r * (ptrdiff_t)L->step - b, (size_t)(L->w + 2 * b));
    return 0;
}
int orbo_stage_blurred(const orbo_extractor* e, int level, uint8_t* dst, size_t dst_step)
{
    if (level < 0 || level >= e->nlevels || !e->lv[level].blur) return -1;
    const level_t* L = &e->lv[level];
    for (int r = 0; r < L->h; ++r) memcpy(dst + (size_t)r * dst_step, L->blur + (size_t)r * L->w, (size_t)L->w);
    return 0;
}
int orbo_stage_candidates(const orbo_extractor* e, int level, orbo_cand* out, int cap)
{
    if (level < 0 || level >= e->nlevels) return -1;
    const level_t* L = &e->lv[level];
    for (int i = 0; i < L->ncand && i < cap; ++i) out[i] = L->cand[i];
    return L->ncand;
}
int orbo_stage_level_keypoints(const orbo_extractor* e, int level, orbo_kp* out, int cap)
{
    if (level < 0 || level >= e->nlevels) return -1;
    const level_t* L = &e->lv[level];
    for (int i = 0; i < L->nkps && i < cap; ++i) out[i] = L->kps[i];
    return L->nkps;
}

/* ------------------------------------------------------------------ CPU timing helper */
typedef struct {
    int nfeatures, nlevels, ini, mn, nframes, w, h, tid, nthreads;
    float sf;
    const uint8_t* frames;
    long long kps;
} bench_arg;

static void* bench_worker(void* p)
{
    bench_arg* a = (bench_arg*)p;
    orbo_extractor* e = orbo_create(a->nfeatures, a->sf, a->nlevels, a->ini, a->mn);
    int cap = a->nfeatures * 2 + 64;
    orbo_kp* kps = (orbo_kp*)malloc(sizeof(orbo_kp) * (size_t)cap);
    uint8_t* desc = (uint8_t*)malloc((size_t)cap * 32);
    for (int f = a->tid; f < a->nframes; f += a->nthreads) {
        int n = orbo_extract(e, a->frames + (size_t)f * a->w * a->h, a->w, a->h, (size_t)a->w, kps, desc, cap);
        if (n > 0) a->kps += n;
    }
    free(desc); free(kps);
    orbo_destroy(e);
    return NULL;
}

double orbo_extract_bench(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                          const uint8_t* frames, int nframes, int w, int h, int nthreads,
                          long long* total_kps)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    bench_arg* args = (bench_arg*)calloc((size_t)nthreads, sizeof(bench_arg));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; ++t) {
        bench_arg a = { nfeatures, nlevels, iniThFAST, minThFAST, nframes, w, h, t, nthreads, scaleFactor, frames, 0 };
        args[t] = a;
        pthread_create(&th[t], NULL, bench_worker, &args[t]);
    }
    long long tot = 0;
    for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); tot += args[t].kps; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_kps) *total_kps = tot;
    free(args); free(th);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ matcher */
/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:46-63 (bit-twiddling popcount) */
int orbo_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4);
        memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

void orbo_hamming_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx, int* best_dist, int* second_dist)
{
    for (int i = 0; i < nq; ++i) {
        int bd = 256, bd2 = 256, bi = -1; /* same initial values as src/ORBmatcher.cc:101-105 */
        for (int j = 0; j < nt; ++j) {
            int d = orbo_descriptor_distance(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < bd) { bd2 = bd; bd = d; bi = j; }
            else if (d < bd2) bd2 = d;
        }
        best_idx[i] = bi; best_dist[i] = bd; second_dist[i] = bd2;
    }
}
