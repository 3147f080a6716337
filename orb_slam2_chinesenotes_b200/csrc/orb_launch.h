// orb_launch.h -- host-callable launchers of the extraction kernels.
#pragma once
#include <cuda_runtime.h>
#include "orb_plan.h"

struct OrbBatch;

size_t orb_fast_smem_bytes(const OrbPlan& plan);
size_t orb_octree_smem_bytes(const OrbPlan& plan);
cudaError_t orb_launch_pyramid(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_blur(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_fast(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_octree(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_describe(const OrbPlan& plan, const OrbBatch& io, int batch, cudaStream_t st);
cudaError_t orb_launch_border(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b, cudaStream_t st);

// Internal: device pointers / pitches / sizes of all pyramid levels of one frame of the context's last
// batch (level 0 is the caller's image or its staged copy), plus the scale tables.  Any output may be
// null.  Returns non-zero when the frame is not resident.  Used by the stereo matcher (orb_match.cu).
struct orbx_ctx;
int orb_ctx_levels(orbx_ctx* ctx, int frame, const uint8_t** ptr, int* pitch, int* w, int* h,
                   float* scale, float* inv_scale, int* nlevels, int* device);
