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
