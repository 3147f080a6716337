// orb_stereo.h -- by-value view of a batch of stereo pairs for the kernels of orb_stereo.cu.
// Pair p reads its left/right level-l image at l[l] + p * lstride[l] / r[l] + p * rstride[l], its keypoints at
// kl / kr + p * kstride (descriptors likewise, 8 words each) and its counts at nl / nr [p * nstride], so the same
// kernels serve one extractor holding L0,R0,L1,R1,... (orbx_extract_stereo_batch) and two extractors holding
// one frame each (orbm_stereo_matches).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orb_b200.h"
#include "orb_plan.h"

struct OrbStereoView {
    int nlevels;
    const uint8_t* l[ORB_MAX_LEVELS]; const uint8_t* r[ORB_MAX_LEVELS];
    size_t lstride[ORB_MAX_LEVELS], rstride[ORB_MAX_LEVELS];
    int lpitch[ORB_MAX_LEVELS], rpitch[ORB_MAX_LEVELS], w[ORB_MAX_LEVELS], h[ORB_MAX_LEVELS];
    float scale[ORB_MAX_LEVELS], inv_scale[ORB_MAX_LEVELS];     // mvScaleFactors / mvInvScaleFactors (src/Frame.cc:73-79)
    const orbx_kp* kl; const orbx_kp* kr;                       // mvKeys / mvKeysRight
    const uint32_t* dl; const uint32_t* dr;                     // mDescriptors / mDescriptorsRight
    size_t kstride;                                             // keypoints between consecutive pairs
    const int* nl; const int* nr; int nstride;
    int cap;                                                    // keypoints per frame the buffers hold
    float bf, mb;                                               // mbf, mb = mbf / fx (src/Frame.cc:121)
    float* u_right; float* depth; size_t ostride;               // mvuRight / mvDepth, [pairs][ostride]
    int* n_stereo;                                              // [pairs] matches before the median cut
    int* sad;                                                   // scratch [pairs][cap]
    uint4* rec;                                                 // scratch [pairs][cap]: right keypoints by row, {x, minr|maxr<<16, index|octave<<16, 0}
    int* row_start;                                             // scratch [pairs][h[0] + 2]
};

// rows either side that certainly cover every band [floor(y - r), ceil(y + r)], r = 2 * scale <= 2 * max_scale
static inline __host__ __device__ int orb_stereo_band(float max_scale) { return (int)(2.0f * max_scale) + 3; }

cudaError_t orb_launch_stereo(const OrbStereoView& V, int pairs, int max_left, cudaStream_t st);
