/*
 * erp_oracle.h -- CPU oracle for the ERP match + eight-point hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library, the
 * C++ class wrappers, the Python binding) may include, link or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker.
 *
 * It is a plain-C restatement of the reference's algorithm
 *   /root/reference/src/feature_matcher.cpp:42-59   (2-NN + ratio test)
 *   /root/reference/src/eight_point.cpp:16-192      (eight-point, initial_guess, find)
 *   /root/reference/src/eight_point.hpp:30-59       (random_array)
 *   /root/reference/src/erp_rotation.cpp:14-63      (eular2rot, rot2eular)
 *   /root/reference/src/epipolar_tool.cpp:100-107   (epipolar residual)
 * plus the third-party arithmetic those lines call, which is NOT under
 * /root/reference: OpenCV 3.4.x (unpinned: README says 3.4.2, CMake dir opencv342,
 * Windows libs 341, vcxproj 3.4.10):
 *   cv::SVDecomp            -> one-sided Jacobi (modules/core/src/lapack.cpp, JacobiSVDImpl_)
 *   cv::decomposeEssentialMat (modules/calib3d/src/five-point.cpp)
 *   cv::DescriptorMatcher FLANNBASED knnMatch -> restated as its exact limit, brute-force L2.
 *
 * PARITY PINNING: the reference holds no golden vectors or asserting tests
 *  and cannot be compiled here (no OpenCV C++), so parity with
 * the reference binary is UNPINNED.  The oracle is pinned instead against outputs of
 * the in-container Python cv2 4.13 (BFMatcher, SVDecomp, decomposeEssentialMat),
 * committed under tests/golden/ with the generating script.
 */
#ifndef ERP_ORACLE_H
#define ERP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::DMatch layout (16 bytes) */
typedef struct {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} orc_dmatch;

enum { ORC_METRIC_ALGEBRAIC = 0, ORC_METRIC_SAMPSON = 1, ORC_METRIC_ANGULAR = 2 };

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* torchrun exports OMP_NUM_THREADS=1: the timed CPU arm undoes it */

/* ---- matching (feature_matcher.cpp:42-59, exact-L2 limit of FLANN) ---- */
/* idx/dist: nq x 2, d2: nq x 2 (fp64 squared distances, may be NULL).  nt >= 2. */
void orc_knn2(const float* q, int nq, const float* t, int nt, int dim,
              int32_t* idx, float* dist, double* d2);
/* per train row: nearest query (exact, lowest index on ties). best_q: nt */
void orc_nn1_reverse(const float* q, int nq, const float* t, int nt, int dim,
                     int32_t* best_q, double* best_d2);
/* full match_two_image: ratio < 0 disables ratio test.  returns #matches, out has room for nq */
int orc_match(const float* q, int nq, const float* t, int nt, int dim,
              float ratio, int cross_check, orc_dmatch* out);

/* ---- geometry ---- */
/* eight_point.cpp:163-186.  xy: n points, stride in bytes between (x,y) float pairs
 * (28 = sizeof(cv::KeyPoint)).  out3: n x 3 fp64 */
void orc_bearings(const void* xy, int stride_bytes, int n, int W, int H, double* out3);
void orc_eular2rot(const double* theta3, double* R9);   /* erp_rotation.cpp:14-40 */
void orc_rot2eular(const double* R9, double* e3);       /* erp_rotation.cpp:43-63 */

/* ---- image / keypoint rotation (SURVEY section 8f "next" rows) ---- */
/* cv::Mat::inv() of a 3x3 CV_64F matrix: OpenCV's closed-form cofactor path (modules/core/src/lapack.cpp) */
int orc_inv3(const double* A9, double* out9);
/* erp_rotation::rotate_pixel, erp_rotation.cpp:66-92; in/out are (row, col) */
void orc_rotate_pixel(const int32_t* rc_in, const double* R9, int width, int height, int32_t* rc_out);
void orc_rotate_pixels(const int32_t* rc_in, int n, const double* R9, int width, int height, int32_t* rc_out);
/* erp_rotation::rotate_image, erp_rotation.cpp:94-122 (8-bit, 3 channels; unmapped pixels are left 0) */
void orc_rotate_image(const uint8_t* im, int width, int height, const double* R9, uint8_t* out);
/* spherical_surf::crop_rotated_image, spherical_surf.cpp:16-48: out is (height/4) x width x 3 */
void orc_crop_rotated_image(const uint8_t* im, int width, int height, float pitch_rot_deg, uint8_t* out);
/* epipolar_tool (ctor geometry + draw_epipole), epipolar_tool.cpp:7-128, sequential semantics: per
 * pixel the LAST key whose |l^T E^T p| < 0.002 colours it, then the 11x11 dots of the right keypoints
 * are painted in key order (clipped to the image: the reference writes out of bounds).
 * left_xy/right_xy: (x, y) float pairs of the n_key selected correspondences; out: out_h x out_w x 3 */
void orc_draw_epipole(const double* E9, const void* left_xy, const void* right_xy, int stride_bytes, int n_key,
                      int im_w, int im_h, int out_w, int out_h, uint8_t* out);
/* spherical_surf::rotate_keypoint, spherical_surf.cpp:50-63: xy (x, y) float pairs, in place */
void orc_rotate_keypoints(void* xy, int stride_bytes, int n, float pitch_rot_inv_deg, int width, int height);

/* OpenCV SVD::compute semantics (no FULL_UV).  A: m x n row-major.
 * k = min(m,n).  w: k, u: m x k, vt: k x n.  u/vt may be NULL. */
void orc_svd(const double* A, int m, int n, double* w, double* u, double* vt);
/* cv::decomposeEssentialMat */
void orc_decompose_essential(const double* E9, double* R1, double* R2, double* t3);

/* eight_point.cpp:16-85.  l3,r3: n x 3 fp64.
 * null_mode = 0: faithful (last row of vt as OpenCV returns it; for n == 8 this is the
 *                smallest NON-ZERO singular vector)
 * null_mode = 1: true null / least-squares vector (pads A to 9 rows with zeros when n < 9)
 * outputs (any may be NULL): e9 = raw vector, Ec9 = rank-2 corrected E,
 * R1e/R2e = XYZ euler (fp32 like Vec3f), T = t (fp32), valid flags. */
void orc_eight_point(const double* l3, const double* r3, int n, int null_mode,
                     double* e9, double* Ec9, float* R1e, float* R2e, float* T,
                     int* R1_valid, int* R2_valid);

/* eight_point.hpp:54-58 + libstdc++ random_shuffle driven by glibc rand().
 * reseed != 0 -> srand(reseed) first (srand(1) == never-seeded process state). */
void orc_random_array(int size, int32_t* out, unsigned reseed);
/* H rows of the first S entries of H successive random_array(M) (eight_point.cpp:99-111) */
void orc_ref_sample_table(int M, int H, int S, int32_t* table, unsigned reseed);

/* Philox4x32-10 minimal-sample generator shared by spec with the device:
 * S distinct indices in [0,M) for hypothesis hyp_id. */
void orc_philox_samples(uint64_t seed, uint64_t hyp_id, int M, int S, int32_t* out);

/* eight_point.cpp:87-150 with an explicit sample table (H x S).
 * cand_R/cand_T: up to 2H x 3 floats (may be NULL), n_cand out, chosen index out.
 * returns 0 ok, 1 if no valid candidate. */
int orc_initial_guess(const double* l3, const double* r3, int M,
                      const int32_t* table, int H, int S, int null_mode,
                      float* R_out, float* T_out,
                      float* cand_R, float* cand_T, int* n_cand, int* chosen);
/* the consensus pick alone (eight_point.cpp:131-149) */
int orc_consensus_pick(const float* cand_R, int C, double* trimmed_mean /* C or NULL */);

/* ---- hypothesis scoring (residual from epipolar_tool.cpp:100-107) ---- */
/* E: H x 9 fp64 (any scale).  l4/r4: m x 4 float bearings.  counts: H */
void orc_score(const double* E, int H, const float* l4, const float* r4, int m,
               int metric, float tau, int32_t* counts);
void orc_inlier_mask(const double* E9, const float* l4, const float* r4, int m,
                     int metric, float tau, uint8_t* mask);
/* minimal-sample RANSAC: hypotheses hyp0..hyp0+H-1, Philox samples, null_mode=1 solve,
 * fp32 scoring, best = max count then lowest id.  returns packed (count<<32 | ~id) */
uint64_t orc_ransac(const double* l3, const double* r3, int M, uint64_t seed,
                    uint64_t hyp0, int H, int S, int metric, float tau,
                    double* best_E9, int32_t* counts /* H or NULL */);

#ifdef __cplusplus
}
#endif
#endif
