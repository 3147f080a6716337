// orb_launch.h -- host-callable launchers of the extraction kernels.
#pragma once
#include <cuda_runtime.h>
#include "orb_plan.h"

struct OrbBatch;

size_t orb_fast_smem_bytes(const OrbPlan& plan);
size_t orb_octree_smem_bytes(const OrbPlan& plan);
cudaError_t orb_launch_pyramid(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_pyramid_level(const OrbPlan& plan, const OrbBatch& io, int batch, int l, cudaStream_t st);   // level l from level l-1
int orb_pyramid_launch_count(const OrbPlan& plan);   // kernels orb_launch_pyramid launches
cudaError_t orb_launch_blur(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
// FAST fetches its tiles with one tensor-map TMA copy per chunk wherever the level's base address, row pitch and
// frame stride are multiples of 16 bytes (all pyramid levels; level 0 when the caller's layout allows).  The
// descriptors of one batch layout are cached here (one per extractor slot) and only re-encoded when it changes.
struct alignas(64) OrbFastMaps {
    unsigned char map[ORB_MAX_LEVELS][128];   // CUtensorMap per level
    unsigned use = 0;                          // bit l: level l has a valid descriptor
    const void* key_img0 = nullptr; const void* key_pyr = nullptr;
    size_t key_stride = 0; int key_pitch = 0, key_batch = 0, key_w = 0, key_h = 0, key_levels = 0;
};
void fast_maps_prepare(const OrbPlan& plan, const OrbBatch& io, int batch, OrbFastMaps* maps);   // encode (or reuse) the descriptors of this layout
cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st, OrbFastMaps* maps = nullptr,
                            int level_lo = 0, int level_hi = ORB_MAX_LEVELS);   // levels [level_lo, level_hi) only
cudaError_t orb_launch_octree(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st, int level_lo = 0, int level_hi = ORB_MAX_LEVELS);
cudaError_t orb_launch_describe(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
// REFLECT_101-padded copies of ALL levels of `frames` frames in one launch (mvImagePyramid's parent buffers)
struct OrbBorderJob { uint32_t off[ORB_MAX_LEVELS]; int pitch[ORB_MAX_LEVELS]; size_t frame_bytes; };
cudaError_t orb_launch_border_levels(const OrbPlan& plan, const OrbBatch& io, const OrbBorderJob& job, uint8_t* dst, int frames, cudaStream_t st);
cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st);

// Internal: device pointers / pitches / sizes of all pyramid levels of one frame of the context's last
// batch (level 0 is the caller's image or its staged copy), plus the scale tables.  Any output may be
// null.  Returns non-zero when the frame is not resident.  Used by the stereo matcher (orb_match.cu).
struct orbx_ctx;
int orb_ctx_levels(orbx_ctx* ctx, int frame, const uint8_t** ptr, int* pitch, int* w, int* h,
                   float* scale, float* inv_scale, int* nlevels, int* device);
