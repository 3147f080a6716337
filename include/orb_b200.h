/*
 * orb_b200.h -- C ABI of the B200-native ORB front end (liborb_b200.so).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  Each entry
 * point names the reference interface it replaces (paths relative to the reference
 * checkout, Hello-Water/ORB-SLAM2-ChineseNotes).  Plain pointers and sizes only; no
 * exceptions cross the boundary; every function returns an orbx_status (0 = OK).
 * The library has NO CPU fallback: without a CUDA device every compute entry point
 * fails with ORBX_E_CUDA.
 *
 * The C++ classes a SLAM build links against (ORB_SLAM2::ORBextractor / ORBmatcher with the
 * reference's exact signatures) are thin forwarders over this ABI; see
 * orb_slam2_chinesenotes_b200/host/ and INTEGRATION.md.
 */
#ifndef ORB_B200_H
#define ORB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ORBX_OK = 0,
    ORBX_E_ARG = 1,        /* null pointer / bad parameter */
    ORBX_E_SHAPE = 2,      /* image shape the path cannot handle (see orbx_shape_supported) */
    ORBX_E_CAPACITY = 3,   /* caller's keypoint capacity smaller than the result; n_out still valid */
    ORBX_E_CUDA = 4,       /* CUDA runtime failure; orbx_last_error() has the text */
    ORBX_E_EMPTY = 5       /* empty image: outputs untouched (src/ORBextractor.cc:1087) */
} orbx_status;

/* Binary layout of cv::KeyPoint (28 bytes): what operator() appends to its vector. */
typedef struct { float x, y, size, angle, response; int octave, class_id; } orbx_kp;

typedef struct orbx_ctx orbx_ctx; /* one per ORBextractor instance; thread-compatible, not shared */

/* ---- extractor ------------------------------------------------------------------- */

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * include/ORBextractor.h:52, src/ORBextractor.cc:498-559.  device = CUDA ordinal. */
int orbx_create(orbx_ctx** out, int nfeatures, float scaleFactor, int nlevels,
                int iniThFAST, int minThFAST, int device);
void orbx_destroy(orbx_ctx* ctx);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares, include/ORBextractor.h:64-84; per_level = mnFeaturesPerLevel
 * (:103).  Any output pointer may be NULL.  Arrays hold orbx_levels() entries. */
int orbx_levels(const orbx_ctx* ctx);
int orbx_tables(const orbx_ctx* ctx, float* scale, float* inv_scale, float* sigma2,
                float* inv_sigma2, int* per_level);

/* 1 if (w,h) can be processed: every pyramid level leaves a non-degenerate FAST area
 * (the reference itself divides by zero otherwise, src/ORBextractor.cc:567-568), level
 * sizes <= 4128 and w*h < 2^24. */
int orbx_shape_supported(const orbx_ctx* ctx, int w, int h);

/* ORBextractor::operator()(image, mask, keypoints, descriptors), include/ORBextractor.h:60,
 * src/ORBextractor.cc:1084-1150, for one 8-bit single-channel image (mask is ignored by the
 * reference).  img/kps/desc/n_out may each be host or device memory.  kps[capacity],
 * desc[capacity*32]; *n_out is the number of keypoints found and can exceed nfeatures
 * (and capacity: then ORBX_E_CAPACITY and the first capacity entries are valid). */
int orbx_extract(orbx_ctx* ctx, const uint8_t* img, int w, int h, size_t pitch,
                 orbx_kp* kps, uint8_t* desc, int capacity, int* n_out);

/* What Frame::ExtractORB (src/Frame.cc:262-268) needs from ONE call, at the lowest latency: operator() for one host
 * image plus, if with_pyramid, mvImagePyramid (include/ORBextractor.h:86: every level as a view inside its
 * REFLECT_101-padded buffer, src/ORBextractor.cc:1157-1178, which Frame::ComputeStereoMatches reads, src/Frame.cc:520,
 * 611-633).  The image goes up through a pinned staging buffer, the whole pipeline replays as one CUDA graph, and keypoints, descriptors and all padded levels come back
 * with it into context-owned pinned host memory: everything in `out` points INTO the context and stays valid until
 * the next call on it, like the reference's mvImagePyramid.  orbx_extract() on a single host image takes the same
 * path and copies the n records out. */
#define ORBX_MAX_LEVELS 16
typedef struct {
    int n;                                   /* keypoints found */
    const orbx_kp* kps;                      /* [n] */
    const uint8_t* desc;                     /* [n][32] */
    int nlevels;                             /* 0 unless with_pyramid */
    const uint8_t* level[ORBX_MAX_LEVELS];   /* first pixel of level l INSIDE its padded buffer (19 px border all round) */
    int level_w[ORBX_MAX_LEVELS], level_h[ORBX_MAX_LEVELS];
    size_t level_pitch[ORBX_MAX_LEVELS];
} orbx_frame_out;
int orbx_extract_frame(orbx_ctx* ctx, const uint8_t* image, int w, int h, size_t pitch, int with_pyramid, orbx_frame_out* out);
/* Host-side microseconds of the last single-call invocation: {copy into the pinned staging buffer, cudaGraphLaunch,
 * wait until the results are back}.  Development aid (tools/latency_once.py). */
int orbx_debug_last_call_us(const orbx_ctx* ctx, double* us3);

/* The same over a batch of equally sized frames: frame f starts at imgs + f*frame_stride.
 * Outputs are [batch][cap_per_frame] / [batch][cap_per_frame][32] / [batch].  This is the
 * call Frame::ExtractORB (src/Frame.cc:262-268) would make once per camera per frame;
 * batching independent frames is how one GPU is kept busy. */
int orbx_extract_batch(orbx_ctx* ctx, const uint8_t* imgs, size_t frame_stride, int batch,
                       int w, int h, size_t pitch, orbx_kp* kps, uint8_t* desc,
                       int cap_per_frame, int* n_out);

/* As above, but the call only enqueues its work (on the context's stream and, with host buffers, on its copy streams)
 * and returns; orbx_sync() waits.  n_out overflow is not reported here.  With every pointer in DEVICE memory nothing is
 * staged.  Host buffers (PINNED, or the copies block) are allowed too and must stay valid until orbx_sync(): several calls
 * may be in flight, they complete in order, and the next call's uploads run beside the previous call's last kernels and
 * downloads -- a stream of batches keeps the copy engines and the SMs busy across call boundaries. */
int orbx_extract_batch_async(orbx_ctx* ctx, const uint8_t* d_imgs, size_t frame_stride, int batch,
                             int w, int h, size_t pitch, orbx_kp* d_kps, uint8_t* d_desc,
                             int cap_per_frame, int* d_n_out);
int orbx_sync(orbx_ctx* ctx);

/* The front end of the stereo Frame constructor (src/Frame.cc:61-115): ExtractORB on the left and the right
 * image (:82-85) followed by ComputeStereoMatches (:107, :513-699), for `pairs` rectified pairs stored as
 * L0,R0,L1,R1,... (frame 2p = left, 2p+1 = right; both extractors of the reference are built from the same
 * settings, so one context serves both).  kps/desc/n_out are as in orbx_extract_batch with batch = 2*pairs;
 * u_right/depth [pairs][cap_per_frame] receive mvuRight/mvDepth of the LEFT keypoints (-1 = no match; entries
 * past n_out[2p] are not written), n_stereo [pairs] the number of matches before the median cut.  bf = mbf,
 * fx = K(0,0) (mb = bf/fx, :121).  Keypoints, descriptors and pyramids stay on the device between the two
 * steps.  Pointers host or device; the _async form only enqueues (see orbx_extract_batch_async). */
int orbx_extract_stereo_batch(orbx_ctx* ctx, const uint8_t* imgs, size_t frame_stride, int pairs,
                              int w, int h, size_t pitch, orbx_kp* kps, uint8_t* desc,
                              int cap_per_frame, int* n_out, float bf, float fx,
                              float* u_right, float* depth, int* n_stereo);
int orbx_extract_stereo_batch_async(orbx_ctx* ctx, const uint8_t* d_imgs, size_t frame_stride, int pairs,
                                    int w, int h, size_t pitch, orbx_kp* d_kps, uint8_t* d_desc,
                                    int cap_per_frame, int* d_n_out, float bf, float fx,
                                    float* d_u_right, float* d_depth, int* d_n_stereo);

/* mvImagePyramid[level] of frame `frame` of the last call (include/ORBextractor.h:86; read by
 * Frame::ComputeStereoMatches, src/Frame.cc:520,611-633).  with_border: the (w+38)x(h+38)
 * REFLECT_101-padded buffer of src/ORBextractor.cc:1159-1174 instead of the ROI.  dst host or
 * device, may be NULL to query the size. */
int orbx_pyramid_level(orbx_ctx* ctx, int frame, int level, int with_border,
                       uint8_t* dst, size_t dst_pitch, int* w, int* h);

/* Use a caller-owned CUDA stream (cudaStream_t) for all work of this context; NULL restores
 * the context's own stream.  orbx_stream returns the stream in use. */
int orbx_set_stream(orbx_ctx* ctx, void* cuda_stream);
void* orbx_stream(const orbx_ctx* ctx);

/* Frames processed per internal pass.  Defaults: 64 when any buffer is host memory (the passes are what the copy /
 * compute pipeline overlaps), 512 when everything is device resident (bounds the work buffers only).  Sets both. */
int orbx_set_chunk(orbx_ctx* ctx, int frames_per_chunk);

const char* orbx_last_error(const orbx_ctx* ctx);

/* ---- stage taps (parity tests and profiling; not needed by a SLAM build) -------------- */

enum { ORBX_STAGE_PYRAMID = 0, ORBX_STAGE_FAST = 1, ORBX_STAGE_BLUR = 2, ORBX_STAGE_OCTREE = 3,
       ORBX_STAGE_DESCRIBE = 4, ORBX_STAGE_STEREO = 5, ORBX_STAGE_COUNT = 6 };

/* Blurred level (src/ORBextractor.cc:1129-1130) of a frame of the last call. */
int orbx_debug_blurred(orbx_ctx* ctx, int frame, int level, uint8_t* dst, size_t dst_pitch);
/* FAST candidates of a level before DistributeOctTree (src/ORBextractor.cc:853-870), as
 * (x, y, score) int triples in the "border frame"; UNORDERED.  Returns the count in *n. */
int orbx_debug_candidates(orbx_ctx* ctx, int frame, int level, int* xys, int cap, int* n);
/* Keypoints of a level after DistributeOctTree in list order, (x, y, score) triples. */
int orbx_debug_level_keypoints(orbx_ctx* ctx, int frame, int level, int* xys, int cap, int* n);
/* Accumulate per-stage GPU time (CUDA events on the context's stream) while enabled. */
int orbx_profile(orbx_ctx* ctx, int enable);
int orbx_stage_ms(orbx_ctx* ctx, float* ms /*[ORBX_STAGE_COUNT]*/, int* launches, int reset);

/* Host-only view of the per-shape plan (no GPU needed): level sizes, processed 30-px cells
 * (src/ORBextractor.cc:826-850), quotas, quadtree roots (:567), candidate capacities. */
int orbx_plan_describe(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int w, int h,
                       int* level_w, int* level_h, int* cells, int* quota, int* n_ini, int* cand_cap);

/* Host-only: the largest n_out operator() can produce for one w x h image with these parameters (every level's
 * DistributeOctTree returns at most max(N + 2, 4 * nIni) keypoints, src/ORBextractor.cc:617-766).  Negative
 * orbx_status on bad parameters / unsupported shape.  Lets a caller size capacities -- and tell the batched
 * matchers a tight bound (orbm_frames.max_n) -- without reading n_out back. */
int orbx_max_keypoints(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int w, int h);

/* Kernels of this library one chunk of frames launches for the context's current shape (pyramid levels, FAST, blur,
 * quadtree, describe, plus the three stereo kernels when `stereo`): what bench.py reports as gpu_launches.  Negative
 * ORBX_E_EMPTY when no shape has been seen yet. */
int orbx_launches_per_chunk(orbx_ctx* ctx, int stereo);

/* ---- matcher -------------------------------------------------------------------- */

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:46-63) over all pairs: for each of the nq
 * 32-byte query descriptors the index of the closest of nt train descriptors (strict <, first
 * wins: the reference's comparison, :129), its distance and the second-best distance
 * (256 when absent).  nprob independent problems are laid out back to back.
 * Pointers host or device. */
int orbm_hamming_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, int nprob,
                    int* best_idx, int* best_dist, int* second_dist, int device);

/* The same with every pointer in DEVICE memory: only enqueues on cuda_stream (cudaStream_t, NULL = default stream). */
int orbm_hamming_bf_async(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int nprob,
                          int* d_best_idx, int* d_best_dist, int* d_second_dist, void* cuda_stream);

/* The Frame fields the window searches read (include/Frame.h:120-175): undistorted keypoints,
 * descriptors, right-image coordinates (NULL = monocular) and the image bounds mnMinX..mnMaxY.
 * The 64x48 grid of Frame::AssignFeaturesToGrid (src/Frame.cc:243-259) is rebuilt on the device. */
typedef struct {
    int n;
    const orbx_kp* kps;      /* mvKeysUn */
    const uint8_t* desc;     /* mDescriptors, n x 32 */
    const float* u_right;    /* mvuRight or NULL */
    float min_x, max_x, min_y, max_y;
} orbm_frame;

/* A Frame kept on the device.  Tracking runs two to four searches on the same Frame (src/Tracking.cc: TrackWithMotionModel,
 * TrackReferenceKeyFrame, SearchLocalPoints, Relocalization), and its mvKeysUn / mDescriptors / mvuRight never change after
 * the constructor (src/Frame.cc:58-174), so they can be uploaded once: orbm_frame_upload copies them into device memory,
 * orbm_frame_view fills an orbm_frame whose kps / desc / u_right are DEVICE pointers (bounds as given at upload; the caller
 * may null u_right in its copy of the view to search monocularly), and EVERY single-problem search below
 * (orbm_search_by_projection_points / _frame, orbm_window_search_best, orbm_search_for_initialization, orbm_search_by_bow,
 * orbm_debug_features_in_area) accepts such a view in place of one with host pointers and skips the upload; it is also a
 * valid one-problem input of the *_batch entry points (kps / desc / u_right with kp_stride = n).  Frames with n = 0 are
 * allowed (the view then holds null pointers).  orbm_frame_release frees the device memory; the handle must outlive
 * every search that uses its view. */
typedef struct orbm_frame_handle orbm_frame_handle;
int orbm_frame_upload(const orbm_frame* F, int device, orbm_frame_handle** handle);
int orbm_frame_view(const orbm_frame_handle* handle, orbm_frame* view);
int orbm_frame_device(const orbm_frame_handle* handle);
void orbm_frame_release(orbm_frame_handle* handle);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:73-157.
 * Map points as arrays [nq]: mTrackProjX/Y/XR (3 floats each), mnTrackScaleLevel, mTrackViewCos,
 * mbTrackInView, isBad(), Observations(), GetDescriptor().  init_assign [n] (or NULL): index of the map
 * point already attached to a keypoint, -1 if none.  assign_out [n]: F.mvpMapPoints afterwards (may be NULL when the
 * frame has no keypoints, as may init_assign: the call then returns 0 matches like the reference on a black image).
 * scale = F.mvScaleFactors.  All pointers are HOST memory.  *nmatches = the function's return value. */
int orbm_search_by_projection_points(const orbm_frame* F, const float* scale, int nlevels,
                                     int nq, const float* proj_xyxr, const int* level, const float* view_cos,
                                     const uint8_t* in_view, const uint8_t* bad, const int* observations,
                                     const uint8_t* qdesc, const int* init_assign, int* assign_out,
                                     float th, float nnratio, int* nmatches, int device);

/* ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:160-300.
 * last_mp[i] != 0: last-frame keypoint i has a map point (world position last_xyz[3i..], descriptor
 * last_mp_desc[32i..], Observations() last_mp_obs[i]); last_outlier = mvbOutlier.  Tcw_*: row-major 4x4
 * poses, K = fx,fy,cx,cy, bf = mbf.  cur_init_obs [n_cur] (or NULL): Observations() of a map point
 * already attached to a current keypoint, -1 if free.  assign_out [n_cur]: index of the last-frame
 * keypoint whose map point is attached afterwards, -2 for a kept pre-attached point, -1 for none. */
int orbm_search_by_projection_frame(const orbm_frame* cur, int n_last, const orbx_kp* kps_last,
                                    const uint8_t* last_mp, const uint8_t* last_outlier, const float* last_xyz,
                                    const uint8_t* last_mp_desc, const int* last_mp_obs,
                                    const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                    const float* scale, int nlevels, const int* cur_init_obs, int* assign_out,
                                    float th, int bMono, int checkOri, int* nmatches, int device);

/* The "best candidate only" window search that the projection overloads share once the caller has
 * projected its points (src/ORBmatcher.cc:160-300 motion model, :303-431 relocalisation, :434-549 loop
 * closing): per query i in order, the candidates of Frame::GetFeaturesInArea(u, v, radius, min_level,
 * max_level) (uvr = 3 floats per query) that are still free -- a keypoint is taken while the point attached
 * to it has Observations() > 0 (init_obs [F->n]: -1 free, else Observations() of the attached point; q_obs
 * [nq]: Observations() of the queries' points, NULL = 1) -- and, if ur != NULL, pass |ur[i] - mvuRight[k]| <=
 * er_max[i] when mvuRight[k] > 0; smallest distance wins (first on ties), accepted when <= th_accept
 * (TH_HIGH, ORBdist or TH_LOW in the reference).  check_ori applies the rotation histogram with q_angle.
 * assign_out [F->n]: query index attached to each keypoint, -2 kept pre-attached point, -1 none. */
int orbm_window_search_best(const orbm_frame* F, int nq, const float* uvr, const int* min_level, const int* max_level,
                            const float* ur, const float* er_max, const uint8_t* valid, const uint8_t* qdesc,
                            const float* q_angle, const int* q_obs, const int* init_obs, int* assign_out,
                            int th_accept, int check_ori, int* nmatches, int device);

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:1055-1180.  prev_matched [F1.n][2] is
 * vbPrevMatched (updated in place), matches12 [F1.n] receives vnMatches12. */
int orbm_search_for_initialization(const orbm_frame* F1, const orbm_frame* F2, float* prev_matched, int* matches12,
                                   int windowSize, float nnratio, int checkOri, int* nmatches, int device);

/* Frame::ComputeStereoMatches, src/Frame.cc:513-699.  The two extractor contexts must still hold the
 * pyramids of the frames that produced the keypoints (mvImagePyramid of both extractors, :520,611-633):
 * frame_l / frame_r index into their last batch.  u_right / depth [nl] receive mvuRight / mvDepth. */
int orbm_stereo_matches(orbx_ctx* ex_left, int frame_l, orbx_ctx* ex_right, int frame_r,
                        int nl, const orbx_kp* kps_l, const uint8_t* desc_l,
                        int nr, const orbx_kp* kps_r, const uint8_t* desc_r,
                        float bf, float fx, float* u_right, float* depth, int* nmatched);

/* ---- matcher, batched and device resident ------------------------------------------------
 * Many independent (frame, query set) problems per launch, one thread block each; every pointer
 * (except `scale`) is DEVICE memory and the calls only enqueue work on `cuda_stream` (cudaStream_t,
 * NULL = default stream).  Frames are taken exactly as orbx_extract_batch leaves them, so
 * extraction -> matching never leaves the GPU.  Results equal those of the single-problem entry
 * points above (and of the reference) problem by problem; a problem whose n exceeds 8192 or
 * kp_stride, or whose nq exceeds nq_stride, gets nmatches = -1.  rounds (or NULL) [nprob] receives
 * the number of passes the order-resolution needed (orb_match_batch.cu). */
typedef struct {
    int nprob;
    const orbx_kp* kps;      /* mvKeysUn of frame p at kps + p*kp_stride */
    const uint8_t* desc;     /* mDescriptors of frame p at desc + p*kp_stride*32 (16-byte aligned) */
    const float* u_right;    /* mvuRight, [nprob][kp_stride], or NULL */
    const int* n;            /* [nprob] keypoints per frame (<= min(kp_stride, 8192)) */
    int kp_stride;
    float min_x, max_x, min_y, max_y;   /* mnMinX..mnMaxY, shared by all frames */
    int max_n;               /* 0, or a bound the caller guarantees for every n[p] (orbx_max_keypoints): the kernels
                                size their shared memory by it instead of kp_stride; a frame that exceeds it gets
                                nmatches = -1 */
} orbm_frames;

/* The map points of SearchByProjection(Frame&, const vector<MapPoint*>&, th): arrays [nprob][nq_stride]
 * with the meaning of orbm_search_by_projection_points. */
typedef struct {
    const int* nq; int nq_stride;
    const float* proj_xyxr; const int* level; const float* view_cos;
    const uint8_t* in_view; const uint8_t* bad; const int* observations; const uint8_t* qdesc;
} orbm_points;

/* src/ORBmatcher.cc:73-157 for every problem.  init_assign / assign_out are [nprob][kp_stride]. */
int orbm_search_by_projection_points_batch(const orbm_frames* F, const float* scale, int nlevels, const orbm_points* Q,
                                           const int* init_assign, int* assign_out, float th, float nnratio,
                                           int* nmatches, int* rounds, void* cuda_stream);

/* Already projected queries of the best-candidate-only overloads (orbm_window_search_best): arrays
 * [nprob][nq_stride]; ur, er_max, valid, q_angle (unless check_ori) and q_obs may be NULL. */
typedef struct {
    const int* nq; int nq_stride;
    const float* uvr; const int* min_level; const int* max_level; const float* ur; const float* er_max;
    const uint8_t* valid; const uint8_t* qdesc; const float* q_angle; const int* q_obs;
} orbm_windows;

/* src/ORBmatcher.cc:160-300 / :303-431 / :434-549 after the projection, for every problem. */
int orbm_window_search_best_batch(const orbm_frames* F, const orbm_windows* Q, const int* init_obs, int* assign_out,
                                  int th_accept, int check_ori, int* nmatches, int* rounds, void* cuda_stream);

/* ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:160-300, for nprob
 * (current frame, last frame) pairs: the projection of the last frame's map points (:172-244) runs on the device too.
 * Tcw_cur / Tcw_last [nprob][16] row-major poses; last-frame arrays [nprob][last_stride]: kps_last (octave, angle),
 * last_mp (has a map point), last_outlier (mvbOutlier, may be NULL), last_xyz [.][3], last_mp_desc [.][32],
 * last_mp_obs (Observations(), NULL = 1); n_last [nprob].  cur_init_obs / assign_out [nprob][cur->kp_stride] as in
 * orbm_search_by_projection_frame.  K, scale: host.  Device pointers otherwise; only enqueues. */
int orbm_search_by_projection_frame_batch(const orbm_frames* cur, const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                          const float* scale, int nlevels, const int* n_last, int last_stride,
                                          const orbx_kp* kps_last, const uint8_t* last_mp, const uint8_t* last_outlier,
                                          const float* last_xyz, const uint8_t* last_mp_desc, const int* last_mp_obs,
                                          const int* cur_init_obs, int* assign_out, float th, int bMono, int checkOri,
                                          int* nmatches, void* cuda_stream);

/* ---- order-free searches: ORBmatcher::Fuse and ORBmatcher::SearchBySim3 ------------------------------------------
 * The search the reference runs for every projected map point in Fuse(KeyFrame*, const vector<MapPoint*>&, th)
 * (src/ORBmatcher.cc:1364-1513), Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, th, vpReplacePoint) (:1516-1633)
 * and SearchBySim3 (:836-1052): among the keypoints KeyFrame::GetFeaturesInArea(u, v, radius) returns
 * (src/KeyFrame.cc:637-676) at level[i] - 1 or level[i] (nPredictedLevel, :1436-1437 / :907-908 / :1593-1594) the one with
 * the smallest descriptor distance, the first of them on ties.  With inv_sigma2 (= pKF->mvInvLevelSigma2, host, nlevels
 * entries) the candidates must also pass Fuse's reprojection-error test (:1440-1469): (ex^2 + ey^2 + er^2) * inv_sigma2 <=
 * 7.8 with er = ur[i] - mvuRight[k] when mvuRight[k] >= 0, (ex^2 + ey^2) * inv_sigma2 <= 5.99 otherwise.  No query depends
 * on another one; what the reference does with a match afterwards (Replace / AddObservation / AddMapPoint, the mutual
 * check of SearchBySim3) is the caller's (host/ORBmatcher_b200.hpp).
 * best_idx [nq]: the keypoint, -1 when there is none or its distance exceeds th_accept (TH_LOW = 50 in Fuse, TH_HIGH = 100
 * in SearchBySim3); best_dist [nq]: the smallest distance (256 without a candidate); *nfound: entries >= 0.
 * uvr = u, v, radius per query; valid (or NULL) = 0 skips a query.  F's bounds are the KEY FRAME's (ints, include/KeyFrame.h:186-189). */
int orbm_window_best_free(const orbm_frame* F, int nq, const float* uvr, const int* level, const float* ur, const uint8_t* valid,
                          const uint8_t* qdesc, const float* inv_sigma2, int nlevels, int th_accept,
                          int* best_idx, int* best_dist, int* nfound, int device);

/* The same for nprob (key frame, point set) problems, device resident: arrays [nprob][nq_stride] (ur, valid may be NULL),
 * best_idx / best_dist [nprob][nq_stride], nfound [nprob] (-1: a frame exceeds 8192 keypoints or a stride).  Only enqueues. */
typedef struct {
    const int* nq; int nq_stride;
    const float* uvr; const int* level; const float* ur; const uint8_t* valid; const uint8_t* qdesc;
} orbm_free_windows;
int orbm_window_best_free_batch(const orbm_frames* F, const orbm_free_windows* Q, const float* inv_sigma2, int nlevels,
                                int th_accept, int* best_idx, int* best_dist, int* nfound, void* cuda_stream);

/* The per-point prologue of Fuse / SearchBySim3 on the device, so that projection -> orbm_window_best_free_batch never leaves
 * the GPU: for every (key frame, map point) the projection and gates of ORBmatcher::Fuse (src/ORBmatcher.cc:1388-1426; the Sim3
 * overload :1546-1584 is the same once the caller has decomposed Scw into R, t, Ow as :1524-1529 do) or, with sim3 != 0, of one
 * direction of SearchBySim3 (:886-909 / :937-960), MapPoint::PredictScale(dist, KeyFrame*) (src/MapPoint.cc:442-457) and the
 * search radius th * scale[level].  pose [nprob][24] (device): Fuse = Rcw (9, row-major), tcw (3), Ow (3), 9 unused; SearchBySim3 =
 * R_a (9), t_a (3) of the key frame that owns the points, then sR (9), tt (3) of the similarity into the other one.  K = fx, fy,
 * cx, cy and scale = mvScaleFactors (host); min_x..max_y = the TARGET key frame's (int) image bounds; scale_factor =
 * mfScaleFactor.  Points as in orbm_project_points_batch (max_distance / min_distance = the RAW mfMaxDistance / mfMinDistance),
 * normal may be NULL with sim3; skip (or NULL) [.][nq_stride]: 1 = the reference's loop skips the point before projecting it
 * (NULL / isBad() / IsInKeyFrame / already found / already matched).  Outputs [nprob][nq_stride] in the layout of
 * orbm_free_windows: uvr, level, ur (or NULL; u - bf * invz, :1402), valid.  Only enqueues. */
int orbm_fuse_project_batch(int nprob, int sim3, const float* pose, const float* K, float bf, float min_x, float max_x, float min_y,
                            float max_y, float scale_factor, const float* scale, int nlevels, float th,
                            const int* nq, int nq_stride, int points_shared, const float* xyz, const float* normal,
                            const float* max_distance, const float* min_distance, const uint8_t* skip,
                            float* uvr, int* level, float* ur, uint8_t* valid, void* cuda_stream);

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:1055-1180, for nprob (F1, F2) pairs.  prev_matched
 * [nprob][F1.kp_stride][2] = vbPrevMatched (window centres; updated in place like the reference does), matches12
 * [nprob][F1.kp_stride] = vnMatches12, nmatches [nprob] the return value.  Frames with at most 8192 keypoints.
 * rounds (or NULL): passes of the order resolution, negative when a problem fell back to the in-order walk. */
int orbm_search_for_initialization_batch(const orbm_frames* F1, const orbm_frames* F2, float* prev_matched, int* matches12,
                                         int windowSize, float nnratio, int checkOri, int* nmatches, int* rounds, void* cuda_stream);

/* ---- map-point side of the matching path (device resident, only enqueue) ------------------
 * Frame::isInFrustum (src/Frame.cc:288-345, camera centre as Frame::UpdatePoseMatrices :280-285, level by
 * MapPoint::PredictScale src/MapPoint.cc:459-475) for every (frame, map point): the loop of
 * Tracking::SearchLocalPoints that prepares SearchByProjection(Frame&, vector<MapPoint*>&, th).
 * Tcw [nprob][16] row-major poses (device); K = fx, fy, cx, cy (host); scale_factor = mfScaleFactor.
 * Map points: xyz / normal [.][3] (GetWorldPos / GetNormal), max_distance / min_distance = mfMaxDistance /
 * mfMinDistance (the 1.2 / 0.8 invariance factors are applied here); [nprob][nq_stride] or, with
 * points_shared, one set [nq_stride] for all frames.  Outputs [nprob][nq_stride] in the layout of orbm_points:
 * in_view = mbTrackInView, proj_xyxr = mTrackProjX/Y/XR, level = mnTrackScaleLevel, view_cos = mTrackViewCos
 * (entries of points not in view keep their old contents, as the reference leaves those members alone);
 * n_in_view [nprob] or NULL. */
int orbm_project_points_batch(int nprob, const float* Tcw, const float* K, float bf, float min_x, float max_x, float min_y,
                              float max_y, float scale_factor, int nlevels, float viewing_cos_limit,
                              const int* nq, int nq_stride, int points_shared, const float* xyz, const float* normal,
                              const float* max_distance, const float* min_distance, uint8_t* in_view, float* proj_xyxr,
                              int* level, float* view_cos, int* n_in_view, void* cuda_stream);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:275-340) for npoints map points at once: the
 * observations of point p are the descriptors desc[offsets[p] .. offsets[p+1]) in the order the reference
 * iterates its observation map (key-frame pointer order); bad (or NULL) marks observations whose key frame
 * isBad().  best_idx [npoints] = index inside the point's own range of the descriptor with the least median
 * distance to the others (first on ties), -1 when there is none (mDescriptor stays as it is), -2 when the
 * point has more than 4096 observations; best_median [npoints] (or NULL) = that median. */
int orbm_distinctive_descriptors(const uint8_t* desc, const int* offsets, int npoints, const uint8_t* bad,
                                 int* best_idx, int* best_median, void* cuda_stream);

/* A DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>, Thirdparty/DBoW2 in upstream ORB-SLAM2; read at
 * src/ORBmatcher.cc:560-577) per frame in CSR form, device memory: node_id [nprob][node_stride] ascending NodeIds,
 * node_off [nprob][node_stride + 1] offsets into feat, n_nodes [nprob], feat [nprob][kp_stride of the frame]
 * feature indices node by node, in the order the vocabulary added them. */
typedef struct {
    const int* node_id; const int* node_off; const int* n_nodes; const int* feat; int node_stride;
} orbm_featvec;

/* ORBmatcher::SearchByBoW for nprob (side A, side B) problems: kf_kf = 0: (KeyFrame*, Frame&, matches),
 * src/ORBmatcher.cc:552-697 (A = key frame, B = frame, bestDist1 <= TH_LOW, b_valid = NULL); kf_kf = 1:
 * (KeyFrame*, KeyFrame*, matches12), :700-832 (bestDist1 < TH_LOW, b_valid marks side-B features whose map point
 * exists and is not bad).  a_valid [nprob][A.kp_stride] likewise for side A.  The frames' kps supply the angles of
 * the rotation histogram (check_ori).  match12 [nprob][A.kp_stride] = matched feature of side B or -1 (the
 * reference's vpMatches12 / vpMapPointMatches hold that feature's map point); match21 [nprob][B.kp_stride] (or
 * NULL) the inverse; nmatches [nprob] the return value (-1: sizes exceed 8192 or the strides).  Only enqueues. */
int orbm_search_by_bow_batch(const orbm_frames* A, const orbm_featvec* VA, const uint8_t* a_valid,
                             const orbm_frames* B, const orbm_featvec* VB, const uint8_t* b_valid,
                             int kf_kf, float nnratio, int check_ori, int* match12, int* match21,
                             int* nmatches, int* rounds, void* cuda_stream);

/* ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t,size_t>>&, bOnlyStereo),
 * src/ORBmatcher.cc:1183-1361 with CheckDistEpipolarLine (:1636-1650), for nprob key-frame pairs: features of shared
 * vocabulary nodes (orbm_featvec, as SearchByBoW), a_valid / b_valid [nprob][kp_stride] = the feature has NO map point yet
 * (and a right coordinate when bOnlyStereo); A->u_right / B->u_right = mvuRight of the key frames (NULL = monocular);
 * F12 [nprob][9] row-major, epipole [nprob][2] = (ex, ey) of :1190-1196 (device); scale = pKF2->mvScaleFactors, sigma2 =
 * pKF2->mvLevelSigma2 (host, nlevels).  A candidate must be within TH_LOW, off the epipole (:1272-1279) and on the epipolar
 * line; of equally close ones the last wins; like the reference (which tests vbMatched2 but never sets it) no feature blocks
 * another.  match12 [nprob][A.kp_stride] = vMatches12 (vMatchedPairs = its entries >= 0 in index order), nmatches [nprob]
 * the return value (-1: sizes exceed 8192 or the strides).  Only enqueues. */
int orbm_search_for_triangulation_batch(const orbm_frames* A, const orbm_featvec* VA, const uint8_t* a_valid,
                                        const orbm_frames* B, const orbm_featvec* VB, const uint8_t* b_valid,
                                        const float* F12, const float* epipole, const float* scale, const float* sigma2, int nlevels,
                                        int check_ori, int* match12, int* nmatches, void* cuda_stream);

/* ---- the same three for ONE problem with HOST arrays (what the C++ forwarders call) ---------
 * orbm_search_by_bow: SearchByBoW (src/ORBmatcher.cc:552-832), feature vectors in CSR form (see orbm_featvec).
 * orbm_project_points: Frame::isInFrustum for n map points of one frame (src/Frame.cc:288-345).
 * orbm_distinctive_descriptor: MapPoint::ComputeDistinctiveDescriptors for one map point (src/MapPoint.cc:275-340). */
int orbm_search_by_bow(const orbm_frame* A, const uint8_t* a_valid, int nn_a, const int* node_id_a, const int* node_off_a, const int* feat_a,
                       const orbm_frame* B, const uint8_t* b_valid, int nn_b, const int* node_id_b, const int* node_off_b, const int* feat_b,
                       int kf_kf, float nnratio, int check_ori, int* match12, int* nmatches, int device);
/* orbm_search_for_triangulation: SearchForTriangulation for one key-frame pair (host arrays; A->u_right / B->u_right = mvuRight
 * or NULL; frames may be resident views) */
int orbm_search_for_triangulation(const orbm_frame* A, const uint8_t* a_valid, int nn_a, const int* node_id_a, const int* node_off_a, const int* feat_a,
                                  const orbm_frame* B, const uint8_t* b_valid, int nn_b, const int* node_id_b, const int* node_off_b, const int* feat_b,
                                  const float* F12, const float* epipole, const float* scale, const float* sigma2, int nlevels,
                                  int check_ori, int* match12, int* nmatches, int device);
int orbm_project_points(const float* Tcw, const float* K, float bf, float min_x, float max_x, float min_y, float max_y,
                        float scale_factor, int nlevels, float viewing_cos_limit, int n, const float* xyz, const float* normal,
                        const float* max_distance, const float* min_distance, uint8_t* in_view, float* proj_xyxr, int* level,
                        float* view_cos, int device);
int orbm_distinctive_descriptor(const uint8_t* desc, int n, const uint8_t* bad, int* best_idx, int* best_median, int device);

/* ---- the per-keypoint steps of the monocular / RGB-D Frame constructors (src/Frame.cc:127-240), device resident --------
 * A stereo frame stays on the GPU from extraction to depth (orbx_extract_stereo_batch); these two do the same for the other
 * two constructors: every pointer except K / dist_coef is DEVICE memory laid out as orbx_extract_batch leaves it
 * (kps [batch][cap_per_frame], n [batch]); the calls only enqueue on cuda_stream.
 * orbx_undistort_keypoints_batch: Frame::UndistortKeyPoints (src/Frame.cc:436-468), i.e. cv::undistortPoints(mat, mat, mK,
 * mDistCoef, cv::Mat(), mK) on every keypoint: K = fx, fy, cx, cy (host); dist_coef = mDistCoef (host; 0, 4, 5, 8 or 12
 * entries k1 k2 p1 p2 [k3 [k4 k5 k6 [s1..s4]]]); with dist_coef[0] == 0 the keypoints are copied (:438-442).  In place
 * (d_kps_un == d_kps) is allowed.
 * orbx_stereo_from_rgbd_batch: Frame::ComputeStereoFromRGBD (src/Frame.cc:702-727): depth image of frame f (float32, already
 * scaled by mDepthMapFactor, src/Tracking.cc:257-258) at d_depth + f * depth_frame_stride bytes, rows depth_pitch bytes apart,
 * read at the DISTORTED keypoint; d_kps_un (NULL = d_kps) supplies the undistorted x of :720.  d_u_right / d_depth_out
 * [batch][cap_per_frame] = mvuRight / mvDepth (-1 without depth and behind n[f]); a keypoint outside the image reads as no depth. */
int orbx_undistort_keypoints_batch(const orbx_kp* d_kps, orbx_kp* d_kps_un, const int* d_n, int cap_per_frame, int batch,
                                   const float* K, const float* dist_coef, int n_dist, void* cuda_stream);
int orbx_stereo_from_rgbd_batch(const orbx_kp* d_kps, const orbx_kp* d_kps_un, const int* d_n, int cap_per_frame, int batch,
                                const float* d_depth, size_t depth_pitch, size_t depth_frame_stride, int w, int h, float bf,
                                float* d_u_right, float* d_depth_out, void* cuda_stream);

/* ---- test taps -------------------------------------------------------------------- */
/* ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1663-1707) as the matcher kernels run it: sizes [n][30] bin
 * counts (host) -> ind [n][3]. */
int orbm_debug_three_maxima(const int* sizes, int n, int* ind, int device);
/* Frame::GetFeaturesInArea (src/Frame.cc:348-409) on the device grid: for query q the keypoint indices inside the
 * window (xyr[3q], xyr[3q+1]) +- xyr[3q+2] at levels [min_level[q], max_level[q]] (-1 = open), IN THE REFERENCE'S ORDER
 * (grid column, grid row, insertion), idx_out [nq][cap], count_out [nq].  Host pointers. */
int orbm_debug_features_in_area(const orbm_frame* F, int nq, const float* xyr, const int* min_level, const int* max_level,
                                int cap, int* idx_out, int* count_out, int device);

#ifdef __cplusplus
}
#endif
#endif
