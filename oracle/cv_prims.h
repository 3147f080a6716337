/*
 * oracle/cv_prims.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the OpenCV-4.13 / glibc-2.39 primitives that the
 * reference's ORB front end calls.  OpenCV C++ is absent from /root/reference and
 * from this image, so these restate the published algorithms; they are pinned
 * bit-for-bit against Python cv2 4.13.0 and libm by tests/test_cv_prims.py and by
 * the committed fixtures in tests/golden/ (SURVEY.md App. A).
 *
 * Call sites in the reference (all in /root/reference/src/ORBextractor.cc):
 *   resize          :1166      copyMakeBorder :1168,1173     GaussianBlur :1130
 *   FAST            :853,859   fastAtan2      :104           cos/sin      :125
 *   cvRound         :82,128,133,134,530,1158
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call anything in this directory.
 */
#ifndef ORACLE_CV_PRIMS_H
#define ORACLE_CV_PRIMS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int x, y, score; } cvp_corner;

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 (11-bit fixed point). */
void cvp_resize_linear_8u(const uint8_t* src, int sw, int sh, size_t sstep,
                          uint8_t* dst, int dw, int dh, size_t dstep);

/* Per-axis tables of the same resize: ofs[d] = left/top source index,
 * c0[d], c1[d] = 11-bit weights of src[ofs] and src[min(ofs+1,n-1)]. */
void cvp_resize_axis_table(int n_src, int n_dst, int* ofs, short* c0, short* c1);

/* cv::copyMakeBorder(src, dst, b,b,b,b, BORDER_REFLECT_101): dst is (w+2b)x(h+2b).
 * src may alias the interior of dst. */
void cvp_border_reflect101(const uint8_t* src, int w, int h, size_t sstep,
                           uint8_t* dst, size_t dstep, int b);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1.
 * In-place (src == dst) is allowed. */
void cvp_gaussian7x7_s2(const uint8_t* src, int w, int h, size_t sstep,
                        uint8_t* dst, size_t dstep);

/* cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16.
 * Writes up to cap corners in row-major order; returns the total number found
 * (may exceed cap).  score is KeyPoint::response. */
int cvp_fast9_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                  cvp_corner* out, int cap);

/* Threshold-independent FAST-9 corner score of one pixel (ring must be inside). */
int cvp_fast9_score(const uint8_t* p, size_t step);

/* cv::fastAtan2 (scalar), degrees in [0,360). */
float cvp_fast_atan2(float y, float x);

/* glibc 2.39 sinf/cosf restated (double arithmetic, no FMA). */
float cvp_sinf(float x);
float cvp_cosf(float x);

/* glibc 2.39 logf restated (double arithmetic, no FMA). */
float cvp_logf(float x);

/* cvRound(double) on x86-64: round-half-to-even. */
int cvp_round(double v);

#ifdef __cplusplus
}
#endif
#endif
