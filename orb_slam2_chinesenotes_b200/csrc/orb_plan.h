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
#ifndef ORB_FAST_STRIP
#define ORB_FAST_STRIP 4     // cells per FAST block (orb_fast.cu); the block has 32 threads per cell
#endif
#define ORB_FAST_WC_STATIC 32    // widest cell handled with compile-time tile geometry
#define ORB_FAST_WPC_STATIC (((ORB_FAST_STRIP * ORB_FAST_WC_STATIC + 9) / 2 + 2) & ~1)   // orb_fast_wpc(ORB_FAST_STRIP, 32)
#define ORB_FAST_RW_STATIC (((ORB_FAST_STRIP * ORB_FAST_WC_STATIC + 15) >> 2) + 1)            // orb_fast_rw(ORB_FAST_STRIP, 32)
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
    int spr;                // FAST strips (runs of up to ORB_FAST_STRIP cells) per cell row
    int strip_first;        // index of this level's first strip in the all-level strip list
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
    int total_strips;       // sum of FAST strips over levels
    int fast_tile_words;    // largest FAST strip tile (both copies) in 32-bit words, multiple of 4
    int fast_eval_max;      // largest evaluated area of a cell (pixels)
    int fast_score_words;   // largest strip score map in 32-bit words, multiple of 4
    int fast_surv_max;      // largest possible number of NMS survivors in a strip
    int fast_raw_words;     // largest fetched strip image in 32-bit words
    int umax[ORB_HALF_PATCH + 1]; // row half-widths of the orientation patch
    uint32_t pyr_bytes;     // bytes of one frame's pyramid block (levels 1..)
    uint32_t blur_bytes;    // bytes of one frame's blur block (levels 0..)
    OrbLevel lv[ORB_MAX_LEVELS];
};

// FAST strip tile (orb_fast.cu): 32-bit words per row of ONE of the two 16-bit copies of a strip of ncs
// cells of width wc.  Strip columns 0 .. ncs*wc+7 are read; a word holds two columns; even count.
static inline __host__ __device__ int orb_fast_wpc(int ncs, int wc) { return (((ncs * wc + 8 + 1) >> 1) + 1 + 1) & ~1; }
// image words fetched per strip row (from the aligned address at or below the first pixel)
static inline __host__ __device__ int orb_fast_rw(int ncs, int wc) { return ((ncs * wc + 15) >> 2) + 1; }
// word offset of the second copy inside a tile row: the first value >= wpc that is 1 (mod 32) (bank spreading)
#ifndef ORB_FAST_OBX
#define ORB_FAST_OBX 1      // copy B starts this many words past a multiple of 32 (odd: rows stay 8-byte aligned)
#endif
static inline __host__ __device__ int orb_fast_ob(int wpc) { return ((wpc + 30) & ~31) + ORB_FAST_OBX; }

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
