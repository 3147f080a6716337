// orb_plan.cpp -- host-side plan: extractor tables and per-shape geometry.
// Pure C++ (no CUDA): compiled with -ffp-contract=off; every float operation below is
// written in the reference's own types and order so the tables are bit-identical.
#include "orb_plan.h"

#include <cmath>
#include <cstring>


static inline int rne_f(float v) { return (int)nearbyintf(v); }   // cvRound(float): cvtss2si
static inline int rne_d(double v) { return (int)nearbyint(v); }   // cvRound(double)
static inline int floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int ceil_f(float v) { int i = (int)v; return i + (i < v); }

// ORBextractor::ORBextractor, src/ORBextractor.cc:498-559
int orb_params_init(OrbParams* p, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    if (!p || nlevels < 1 || nlevels > ORB_MAX_LEVELS || nfeatures < 0 || !(scaleFactor > 1.0f)) return 1;
    std::memset(p, 0, sizeof(*p));
    p->nfeatures = nfeatures; p->nlevels = nlevels; p->iniTh = iniTh; p->minTh = minTh;
    p->scaleFactor = (double)scaleFactor;
    p->scale[0] = 1.0f; p->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        p->scale[i] = (float)((double)p->scale[i - 1] * p->scaleFactor);   // :506 (float * double member)
        p->sigma2[i] = p->scale[i] * p->scale[i];                          // :507
    }
    for (int i = 0; i < nlevels; ++i) {
        p->inv_scale[i] = 1.0f / p->scale[i];                              // :513
        p->inv_sigma2[i] = 1.0f / p->sigma2[i];                            // :514
    }
    float factor = (float)(1.0 / p->scaleFactor);                          // :520
    float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)); // :522
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        p->per_level[l] = rne_f(nDesired);                                 // :527
        sum += p->per_level[l];
        nDesired *= factor;                                                // :529
    }
    p->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0; // :531
    // :544-558 umax: row half-widths of the 31-px circular patch
    int v, v0;
    int vmax = floor_f((float)ORB_HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    int vmin = ceil_f((float)ORB_HALF_PATCH * std::sqrt(2.f) / 2);
    const double hp2 = ORB_HALF_PATCH * ORB_HALF_PATCH;
    for (v = 0; v <= vmax; ++v) p->umax[v] = rne_d(std::sqrt(hp2 - v * v));
    for (v = ORB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (p->umax[v0] == p->umax[v0 + 1]) ++v0;
        p->umax[v] = v0;
        ++v0;
    }
    return 0;
}

// cv::resize INTER_LINEAR 8U coefficient table for one axis (OpenCV 4.13 imgproc/resize.cpp):
// fx = (float)((d+0.5)*scale-0.5) in double then one cast, weights saturate_cast<short>(w*2048).
static void axis_table(int n_src, int n_dst, std::vector<OrbTap>* out)
{
    double inv_scale = (double)n_dst / (double)n_src;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < n_dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
        int c0 = rne_f((1.f - f) * 2048.f);
        int c1 = rne_f(f * 2048.f);
        OrbTap t; t.ofs = s; t.c01 = (uint32_t)(c0 & 0xffff) | ((uint32_t)(c1 & 0xffff) << 16);
        out->push_back(t);
    }
}

static inline uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

int orb_plan_build(const OrbParams* p, int w, int h, OrbPlan* plan, std::vector<OrbTap>* taps)
{
    if (!p || !plan || w <= 0 || h <= 0) return 1;
    std::memset(plan, 0, sizeof(*plan));
    if (taps) taps->clear();
    plan->nlevels = p->nlevels; plan->w = w; plan->h = h; plan->iniTh = p->iniTh; plan->minTh = p->minTh;
    for (int i = 0; i <= ORB_HALF_PATCH; ++i) plan->umax[i] = p->umax[i];
    if ((long long)w * h >= (1 << 24)) return 1;
    uint32_t pyr = 0, blur = 0;
    int cells = 0, tiles = 0, cand = 0, kp = 0;
    const float Wcell = 30.f;                                              // src/ORBextractor.cc:799
    for (int l = 0; l < p->nlevels; ++l) {
        OrbLevel& L = plan->lv[l];
        L.w = rne_f((float)w * p->inv_scale[l]);                           // :1158
        L.h = rne_f((float)h * p->inv_scale[l]);
        if (L.w > ORB_MAX_DIM || L.h > ORB_MAX_DIM) return 1;
        L.W = L.w - 2 * ORB_BORDER0; L.H = L.h - 2 * ORB_BORDER0;          // :804-807
        if (L.W <= 0 || L.H <= 0) return 1;
        L.nIni = (int)std::round((float)L.W / (float)L.H);                 // :567 (round half away)
        if (L.nIni < 1) return 1;                                          // the reference divides by zero here
        L.hX = (float)L.W / (float)L.nIni;                                 // :568
        L.pitch = (int)align_up((uint32_t)L.w, 64);
        if (l > 0) { L.img_off = pyr; pyr += align_up((uint32_t)L.pitch * (uint32_t)L.h, 256); }
        L.blur_off = blur; blur += align_up((uint32_t)L.pitch * (uint32_t)L.h, 256);
        // :816-819 cell grid
        const float width = (float)L.W, height = (float)L.H;
        const int nCols = (int)(width / Wcell), nRows = (int)(height / Wcell);
        L.ncx = L.ncy = 0; L.wCell = L.hCell = 1;
        int cap = 0;
        if (nCols > 0 && nRows > 0) {
            L.wCell = (int)std::ceil(width / (float)nCols);
            L.hCell = (int)std::ceil(height / (float)nRows);
            const int maxBorderX = L.w - ORB_BORDER0, maxBorderY = L.h - ORB_BORDER0;
            // :828-848 skip rules; iniX/iniY grow with the index so processed cells are a prefix
            for (int i = 0; i < nRows; ++i) if ((float)(ORB_BORDER0 + i * L.hCell) < (float)(maxBorderY - 3)) L.ncy = i + 1;
            for (int j = 0; j < nCols; ++j) if ((float)(ORB_BORDER0 + j * L.wCell) < (float)(maxBorderX - 6)) L.ncx = j + 1;
            for (int i = 0; i < L.ncy; ++i) {
                int y0 = ORB_BORDER0 + i * L.hCell, y1 = y0 + L.hCell + 6; if (y1 > maxBorderY) y1 = maxBorderY;
                int eh = y1 - y0 - 6; if (eh < 0) eh = 0;
                for (int j = 0; j < L.ncx; ++j) {
                    int x0 = ORB_BORDER0 + j * L.wCell, x1 = x0 + L.wCell + 6; if (x1 > maxBorderX) x1 = maxBorderX;
                    int ew = x1 - x0 - 6; if (ew < 0) ew = 0;
                    cap += ((ew + 1) / 2) * ((eh + 1) / 2);                // strict 3x3 maxima cannot be denser
                }
            }
            // orb_fast.cu: a warp takes a band of one or two cells (at most ORB_FAST_BAND evaluated columns) and
            // keeps one word per row pair and lane; the slot mask of a lane is 32 bits wide (31 row pairs + a dummy)
            if (L.wCell > ORB_FAST_BAND || L.hCell > 62) return 1;
            if ((L.hCell + 1) / 2 > plan->fast_stash_slots) plan->fast_stash_slots = (L.hCell + 1) / 2;
            // order key (cell, y-in-cell, x-in-cell) must fit 24 bits
            if ((long long)L.ncx * L.ncy * L.wCell * L.hCell >= (1 << 24)) return 1;
        }
        L.cell_first = cells; cells += L.ncx * L.ncy;
        L.fcpb = 2 * L.wCell <= ORB_FAST_BAND ? 2 : 1;
        L.fbands = (L.ncx + L.fcpb - 1) / L.fcpb;
        L.quota = p->per_level[l];
        L.cand_off = cand; L.cand_cap = cap; cand += (cap + 31) / 32 * 32;
        // list size never exceeds max(N+2, 4*nIni) (see orb_octree.cuh); keep a little slack
        int nodes = L.quota + 3 > 4 * L.nIni ? L.quota + 3 : 4 * L.nIni;
        L.kp_off = kp; L.kp_cap = nodes; kp += (nodes + 31) / 32 * 32;
        if (nodes > plan->max_nodes) plan->max_nodes = nodes;
        if (nodes > 4095) return 1;                                        // node ids are 12-bit (orb_octree.cu)
        L.scale = p->scale[l];
        L.size = (float)(int)((float)ORB_PATCH * p->scale[l]);             // :886 int scaledPatchSize
        if (l > 0 && taps) {
            L.xtab = (int)taps->size(); axis_table(plan->lv[l - 1].w, L.w, taps);
            L.ytab = (int)taps->size(); axis_table(plan->lv[l - 1].h, L.h, taps);
            // the same rows again for the pyramid walker: 16-byte entries { source row, c0 << 12, c1 << 12, 0 }
            if (taps->size() & 1) taps->push_back(OrbTap{ 0, 0u });
            L.ytab4 = (int)taps->size();
            for (int d = 0; d < L.h; ++d) {
                const OrbTap t = (*taps)[(size_t)L.ytab + d];
                taps->push_back(OrbTap{ t.ofs, (t.c01 & 0xffffu) << 12 });
                taps->push_back(OrbTap{ (int)((t.c01 >> 16) << 12), 0u });
            }
        }
        L.blur_tiles_x = (L.w + ORB_BLUR_TW - 1) / ORB_BLUR_TW;
        L.blur_tiles_y = (L.h + ORB_BLUR_TH - 1) / ORB_BLUR_TH;
        L.blur_tile_first = tiles; tiles += L.blur_tiles_x * L.blur_tiles_y;
    }
    plan->total_cells = cells; plan->total_blur_tiles = tiles;
    plan->cand_per_frame = cand; plan->kp_per_frame = kp;
    plan->pyr_bytes = align_up(pyr + 64, 256); plan->blur_bytes = align_up(blur + 64, 256);
    return 0;
}
