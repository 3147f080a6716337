// orb_plan.h -- per-shape execution plan shared by host and device code.
//
// Everything the kernels need to know about one (extractor parameters, image shape) pair
// is computed once on the host (orb_plan.cpp) and handed to every kernel BY VALUE as a
// __grid_constant__ parameter, so contexts with different shapes can coexist.
//
// Reference arithmetic reproduced here (paths relative to the reference checkout):
//   scale tables, quotas        src/ORBextractor.cc:498-534
//   level sizes                 src/ORBextractor.cc:1157-1158
//   30-px cell grid             src/ORBextractor.cc:799-850
//   quadtree roots              src/ORBextractor.cc:567-586
#pragma once

#include <stdint.h>

#define ORB_MAX_LEVELS 16
#define ORB_SMEM_OPTIN (224 * 1024)   // dynamic shared memory a kernel opts in to (sm_100 allows 227 KB per block INCLUDING its static shared memory); set as a constant, never per launch
#define ORB_EDGE 19          // EDGE_THRESHOLD, src/ORBextractor.cc:74
#define ORB_HALF_PATCH 15    // HALF_PATCH_SIZE, :73
#define ORB_PATCH 31         // PATCH_SIZE, :72
#define ORB_BORDER0 16       // minBorderX = EDGE_THRESHOLD-3, :804
#define ORB_MAX_DIM 4128     // candidate coordinates are packed in 12 bits (border frame)
#define ORB_FAST_BAND 64     // evaluated columns a FAST warp covers (orb_fast.cu): 32 lanes x 2 pixels
#define ORB_BLUR_TW 120      // blur tile (orb_dense.cu): a warp makes 120 output pixels per row (30 lanes x 4 + 2 apron lanes)
#ifndef ORB_BLUR_TH
#define ORB_BLUR_TH 64       // ... and walks down this many rows
#endif

struct OrbLevel {
    int w, h;               // level image size
    int pitch;              // row pitch of this level in the pyramid / blur blocks (bytes, multiple of 64)
    uint32_t img_off;       // byte offset inside a frame's pyramid block (level 0: unused, image is external)
    uint32_t blur_off;      // byte offset inside a frame's blur block
    // FAST cell grid (reference grid restricted to the cells that are not skipped)
    int wCell, hCell;       // reference cell size
    int ncx, ncy;           // processed cell columns / rows
    int cell_first;         // index of this level's first cell in the all-level cell list
    int fcpb;               // FAST: cells per band (2 while two cells fit ORB_FAST_BAND columns, else 1)
    int fbands;             // FAST: bands per cell row
    // quadtree
    int W, H;               // maxBorder-minBorder extents = w-32, h-32
    int nIni;               // number of root nodes
    float hX;               // root width (float, as in the reference)
    int quota;              // mnFeaturesPerLevel[level]
    int cand_off, cand_cap; // slice of a frame's candidate block (entries)
    int kp_off, kp_cap;     // slice of a frame's level-keypoint block (entries)
    // output
    float scale;            // mvScaleFactor[level]
    float size;             // KeyPoint::size = (int)(31*scale)
    // resize tables (entries into the table buffer): for level >= 1, from level-1 to this level
    int xtab, ytab;
    int ytab4;              // the y table again as 16-byte entries { source row, c0 << 12, c1 << 12, 0 } (even index)
    // blur tiles
    int blur_tiles_x, blur_tiles_y, blur_tile_first;
};

struct OrbPlan {
    int nlevels;
    int w, h;
    int iniTh, minTh;
    int total_cells;        // sum of processed cells over levels
    int total_blur_tiles;
    int cand_per_frame;     // candidate entries per frame (sum of cand_cap)
    int kp_per_frame;       // level-keypoint entries per frame (sum of kp_cap)
    int max_nodes;          // largest kp_cap: quadtree shared-memory sizing
    int fast_stash_slots;   // FAST: row pairs of the tallest cell (survivor stash of a warp)
    int umax[ORB_HALF_PATCH + 1]; // row half-widths of the orientation patch
    uint32_t pyr_bytes;     // bytes of one frame's pyramid block (levels 1..)
    uint32_t blur_bytes;    // bytes of one frame's blur block (levels 0..)
    OrbLevel lv[ORB_MAX_LEVELS];
};

// packed candidate / level keypoint: x | y<<12 | score<<24, x,y in the border frame
static inline __host__ __device__ uint32_t orb_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
#define ORB_PX(p) ((int)((p) & 0xfffu))
#define ORB_PY(p) ((int)(((p) >> 12) & 0xfffu))
#define ORB_PS(p) ((int)((p) >> 24))

struct OrbParams {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;     // include/ORBextractor.h:99: stored as double
    float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS], sigma2[ORB_MAX_LEVELS], inv_sigma2[ORB_MAX_LEVELS];
    int per_level[ORB_MAX_LEVELS];
    int umax[ORB_HALF_PATCH + 1];
};

// resize table entry: source index and the two 11-bit weights (c0 | c1 << 16)
struct OrbTap { int ofs; uint32_t c01; };

#include <vector>
int orb_params_init(OrbParams* p, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh);
// returns 0 on success, 1 if the shape is unsupported; taps receives all resize tables
int orb_plan_build(const OrbParams* p, int w, int h, OrbPlan* plan, std::vector<OrbTap>* taps);
