/*
 * oracle/orb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's ORB front-end hot path
 * (/root/reference/src/ORBextractor.cc, ORBmatcher.cc, Frame.cc).
 * It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Pinning: the OpenCV/libm primitives underneath (cv_prims.c) are checked bit-for-bit
 * against Python cv2 4.13.0 and glibc; the pipeline above them is checked against the
 * reference's own unmodified sources compiled into oracle/_ref/liborbref.so
 * (tests/test_oracle_vs_ref.py) and against fixtures generated from that library and
 * committed under tests/golden/.  The reference itself ships no tests or golden
 * vectors.
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, size, angle, response; int octave, class_id; } orbo_kp; /* == cv::KeyPoint */
typedef struct { int x, y, score; } orbo_cand; /* FAST candidate, "border frame" (origin at pixel 16,16) */

typedef struct orbo_extractor orbo_extractor;

/* ORBextractor::ORBextractor, ORBextractor.cc:498-559 */
orbo_extractor* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orbo_destroy(orbo_extractor* e);
int orbo_levels(const orbo_extractor* e);
void orbo_tables(const orbo_extractor* e, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                 int* per_level, int* umax16, int* pattern1024);

/* ORBextractor::operator(), ORBextractor.cc:1084-1150.  Returns n (may exceed cap; at most cap
 * entries are written), -1 for an empty image (outputs untouched), -2 for an unsupported shape. */
int orbo_extract(orbo_extractor* e, const uint8_t* img, int w, int h, size_t step,
                 orbo_kp* kps, uint8_t* desc, int cap);

/* Stage outputs of the last orbo_extract call on e. */
int orbo_stage_level_size(const orbo_extractor* e, int level, int* w, int* h);
int orbo_stage_pyramid(const orbo_extractor* e, int level, int with_border, uint8_t* dst, size_t dst_step);
int orbo_stage_blurred(const orbo_extractor* e, int level, uint8_t* dst, size_t dst_step);
int orbo_stage_candidates(const orbo_extractor* e, int level, orbo_cand* out, int cap);
int orbo_stage_level_keypoints(const orbo_extractor* e, int level, orbo_kp* out, int cap);

/* DistributeOctTree as a pure function (ORBextractor.cc:562-792 with the pointer
 * tie-break of :711 replaced by creation order).  Candidates in
 * reference order; returns the number of retained keypoints, indices into cand in
 * list order in out_idx. */
int orbo_distribute(const orbo_cand* cand, int n, int minX, int maxX, int minY, int maxY, int N,
                    int* out_idx, int cap);

/* CPU timing helper for the "port" baseline: frames contiguous with pitch w, round-robin
 * over nthreads (pthreads), one extractor per thread.  Returns wall seconds. */
double orbo_extract_bench(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                          const uint8_t* frames, int nframes, int w, int h, int nthreads,
                          long long* total_kps);

/* ---- matcher (ORBmatcher.cc, Frame.cc) ------------------------------------------- */

/* ORBmatcher::DescriptorDistance, ORBmatcher.cc:46-63 */
int orbo_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* All-pairs best / second-best (strict <, first wins), the brute-force kernel's checker. */
void orbo_hamming_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, int* best_idx, int* best_dist, int* second_dist);

/* ---- window matchers and stereo rows (orb_match_oracle.c) --------------------------- */

/* Frame::GetFeaturesInArea on a 64x48 grid built like Frame::AssignFeaturesToGrid (src/Frame.cc:243-259, 348-409) */
int orbo_features_in_area(int n, const orbo_kp* kps, float minX, float maxX, float minY, float maxY,
                          float x, float y, float r, int minLevel, int maxLevel, int* out, int cap);

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:1055-1180 */
int orbo_search_for_initialization(int n1, const orbo_kp* kps1, const uint8_t* desc1,
                                   int n2, const orbo_kp* kps2, const uint8_t* desc2,
                                   float minX, float maxX, float minY, float maxY,
                                   float* prev_matched, int* matches12, int windowSize, float nnratio, int checkOri);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:73-157 */
int orbo_search_by_projection_points(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                                     const float* scale, float minX, float maxX, float minY, float maxY,
                                     int nq, const float* proj_xyxr, const int* level, const float* view_cos,
                                     const uint8_t* in_view, const uint8_t* bad, const int* observations,
                                     const uint8_t* qdesc, const int* init_assign, int* assign_out, float th, float nnratio);

/* ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:160-300 */
int orbo_search_by_projection_frame(int n_cur, const orbo_kp* kps_cur, const uint8_t* desc_cur, const float* u_right_cur,
                                    int n_last, const orbo_kp* kps_last, const uint8_t* last_mp, const uint8_t* last_outlier,
                                    const float* last_xyz, const uint8_t* last_mp_desc, const int* last_mp_obs,
                                    const float* Tcw_cur, const float* Tcw_last, const float* K, float bf,
                                    const float* scale, float minX, float maxX, float minY, float maxY,
                                    const int* cur_init_obs, int* assign_out, float th, int bMono, float nnratio, int checkOri);

/* best-candidate-only window search with caller-side projection, and the relocalisation overload
 * (src/ORBmatcher.cc:303-431) built on it: orb_window_oracle.c */
int orbo_window_search_best(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                            float minX, float maxX, float minY, float maxY,
                            int nq, const float* uvr, const int* min_level, const int* max_level,
                            const float* ur, const float* er_max, const uint8_t* valid, const uint8_t* qdesc,
                            const float* q_angle, const int* q_obs, const int* init_obs, int* assign_out,
                            int th_accept, int check_ori);
int orbo_search_by_projection_reloc(int n_cur, const orbo_kp* kps_cur, const uint8_t* desc_cur,
                                    float minX, float maxX, float minY, float maxY, const float* scale,
                                    int nkf, const uint8_t* has_mp, const uint8_t* bad, const uint8_t* already_found,
                                    const float* xyz, const uint8_t* mp_desc, const int* pred_level,
                                    const float* min_dist, const float* max_dist, const float* kf_angle,
                                    const float* Tcw, const float* K, const uint8_t* cur_taken, int* assign_out,
                                    float th, int ORBdist, int check_ori,
                                    float* uvr_out, int* minl_out, int* maxl_out, uint8_t* valid_out);

/* ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:434-549 (orb_window_oracle.c) */
int orbo_search_by_projection_sim3(int n, const orbo_kp* kps, const uint8_t* desc,
                                   float minX, float maxX, float minY, float maxY, const float* scale,
                                   const float* K, const float* Scw, int npts, const uint8_t* bad, const float* xyz, const float* normal,
                                   const uint8_t* mp_desc, const int* pred_level, const float* min_dist, const float* max_dist,
                                   const int* matched_in, int* assign_out, int th,
                                   float* uvr_out, int* minl_out, int* maxl_out, uint8_t* valid_out);

/* Frame::AssignFeaturesToGrid once, Frame / KeyFrame::GetFeaturesInArea many times */
void* orbo_grid_create(int n, const orbo_kp* kps, float minX, float maxX, float minY, float maxY);
int orbo_grid_query(const void* grid, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap);
void orbo_grid_destroy(void* grid);

/* order-free searches, orb_fuse_oracle.c: ORBmatcher::Fuse (src/ORBmatcher.cc:1364-1513, :1516-1633) and
 * ORBmatcher::SearchBySim3 (:836-1052); map points are rows of plain arrays, -1 = NULL (see the file's header) */
int orbo_window_best_free(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                          float minX, float maxX, float minY, float maxY,
                          int nq, const float* uvr, const int* level, const float* ur, const uint8_t* valid, const uint8_t* qdesc,
                          const float* inv_sigma2, int th_accept, int* best_idx, int* best_dist);
int orbo_fuse(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
              float minX, float maxX, float minY, float maxY, const float* scale, const float* inv_sigma2,
              const float* K, float bf, const float* Rcw, const float* tcw, const float* Ow,
              int npts, uint8_t* bad, const float* xyz, const float* normal, const uint8_t* mp_desc, const int* pred_level,
              const float* min_dist, const float* max_dist, int* nobs, int* kf_idx, int* replaced_by,
              int nlist, const int* list, int* kf_mp, float th,
              float* uvr_out, int* level_out, float* ur_out, uint8_t* valid_out);
int orbo_fuse_sim3(int n, const orbo_kp* kps, const uint8_t* desc, const float* u_right,
                   float minX, float maxX, float minY, float maxY, const float* scale,
                   const float* K, const float* Scw,
                   int npts, uint8_t* bad, const float* xyz, const float* normal, const uint8_t* mp_desc, const int* pred_level,
                   const float* min_dist, const float* max_dist, int* nobs, int* kf_idx,
                   int nlist, const int* list, int* kf_mp, int* replace_out, float th,
                   float* uvr_out, int* level_out, float* ur_out, uint8_t* valid_out);
int orbo_search_by_sim3(int n1, const orbo_kp* kps1, const uint8_t* desc1, const int* mp1,
                        int n2, const orbo_kp* kps2, const uint8_t* desc2, const int* mp2,
                        float minX, float maxX, float minY, float maxY, const float* scale, const float* K,
                        const float* R1w, const float* t1w, const float* R2w, const float* t2w, float s12, const float* R12, const float* t12,
                        int npts, const uint8_t* bad, const float* xyz, const uint8_t* mp_desc, const int* pred_level,
                        const float* min_dist, const float* max_dist, const int* idx_in_kf2,
                        int* matches12, float th,
                        float* uvr1_out, int* level1_out, uint8_t* valid1_out, float* uvr2_out, int* level2_out, uint8_t* valid2_out);

void orbo_fuse_project(int sim3, const float* pose, const float* K, float bf, float minX, float maxX, float minY, float maxY,
                       float scale_factor, const float* scale, int nlevels, float th,
                       int n, const float* xyz, const float* normal, const float* max_d, const float* min_d, const uint8_t* skip,
                       float* uvr, int* level, float* ur, uint8_t* valid);

/* Frame::ComputeStereoMatches, src/Frame.cc:513-699 */
/* ORBmatcher::SearchByBoW, src/ORBmatcher.cc:552-697 (strict = 0, valid2 = NULL) and :700-832 (strict = 1). */
int orbo_search_by_bow(int n1, const orbo_kp* kps1, const uint8_t* desc1, const uint8_t* valid1,
                       int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                       int n2, const orbo_kp* kps2, const uint8_t* desc2, const uint8_t* valid2,
                       int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                       float nnratio, int checkOri, int strict, int* match12, int* match21);

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:1183-1361 (+ CheckDistEpipolarLine :1636-1650) */
int orbo_search_for_triangulation(int n1, const orbo_kp* kps1, const uint8_t* desc1, const uint8_t* has_mp1, const float* u_right1,
                                  int nn1, const int* node_id1, const int* node_off1, const int* feat1,
                                  int n2, const orbo_kp* kps2, const uint8_t* desc2, const uint8_t* has_mp2, const float* u_right2,
                                  int nn2, const int* node_id2, const int* node_off2, const int* feat2,
                                  const float* F12, const float* epipole, const float* scale, const float* sigma2,
                                  int onlyStereo, int checkOri, int* match12);

/* Frame::UndistortKeyPoints (src/Frame.cc:436-468, cv::undistortPoints of OpenCV 4.13) and Frame::ComputeStereoFromRGBD (:702-727) */
void orbo_undistort_keypoints(int n, const orbo_kp* kps, const float* K, const float* dist, int n_dist, orbo_kp* out);
void orbo_stereo_from_rgbd(int n, const orbo_kp* kps, const orbo_kp* kps_un, const float* depth, int w, int h, float bf,
                           float* u_right, float* depth_out);

/* map-point side, orb_mappoint_oracle.c */
int orbo_distinctive_descriptor(const uint8_t* desc, int n, const uint8_t* bad, int* median_out);
int orbo_predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels);
float orbo_log_scale_factor(float scale_factor);
int orbo_is_in_frustum(const float* Tcw, const float* K, float bf, float minX, float maxX, float minY, float maxY,
                       float scale_factor, int nlevels, float viewing_cos_limit, int n, const float* xyz, const float* normal,
                       const float* max_distance, const float* min_distance, unsigned char* in_view, float* proj_xyxr,
                       int* level, float* view_cos);

int orbo_stereo_matches(const orbo_extractor* eL, const orbo_extractor* eR,
                        int nl, const orbo_kp* kps_l, const uint8_t* desc_l,
                        int nr, const orbo_kp* kps_r, const uint8_t* desc_r,
                        float bf, float fx, float* u_right, float* depth);

#ifdef __cplusplus
}
#endif
#endif
