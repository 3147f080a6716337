// orb_match_common.cuh -- pieces shared by the matcher kernels (orb_match.cu, orb_match_batch.cu, orb_match_bow.cu).
#pragma once
#include <cstdio>

// compute-sanitizer is closed on the GPU pool this was developed on, so the kernels added last carry their own index checks,
// compiled in by -DORB_BOUNDS_CHECK (tools/build_checked.sh builds liborb_b200_checked.so; the GPU tests run against it with
// ORB_B200_LIB): a failed check prints its location and traps, which fails the launch and the test.
#ifdef ORB_BOUNDS_CHECK
#define ORB_CHECK(cond) do { if (!(cond)) { printf("ORB_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define ORB_CHECK(cond) do { } while (0)
#endif

#ifndef HISTO_LENGTH
#define HISTO_LENGTH 30 // src/ORBmatcher.cc:39
#endif

// ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1663-1707: the three fullest bins of the rotation histogram; the
// second and third are dropped (-1) when they hold less than a tenth of the first.  `sizes` = entries per bin.
__device__ __forceinline__ void orb_three_maxima(const int* sizes, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; ++i) {
        const int s = sizes[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
}
