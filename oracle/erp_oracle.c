/*
 * erp_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see erp_oracle.h).
 *
 * Every function cites the reference lines it restates.  Paths are relative to
 * /root/reference.  OpenCV itself is a third-party dependency that is absent from
 * the reference tree; its published algorithms (one-sided Jacobi SVD,
 * decomposeEssentialMat) are restated here and pinned against cv2 outputs in
 * tests/golden/.
 *
 * Build: make -C oracle        (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp)
 * -ffp-contract=off matters: fused operations appear only where fma()/fmaf() is
 * written, so the fp32 scoring chain and the fp64 distance chain are bit-identical
 * to the CUDA kernels, which spell the same chain with __fmaf_rn / __fma_rn.
 */
#include "erp_oracle.h"

#include <float.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* image / keypoint rotation: src/erp_rotation.cpp:66-122, src/spherical_surf.cpp:16-63 */
/* ------------------------------------------------------------------------- */
int orc_inv3(const double* S, double* t)
{
    /* cv::invert, n == 3, CV_64F: determinant by the first row, then the adjugate times 1/det */
    double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.0) { memset(t, 0, 9 * sizeof(double)); return 0; }
    d = 1.0 / d;
    t[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
    return 1;
}

void orc_rotate_pixel(const int32_t* in, const double* R, int width, int height, int32_t* out)
{
    /* :68  Vec2d(M_PI*in[0]/height, 2*M_PI*in[1]/width)  -- evaluated left to right in double */
    double lat = M_PI * in[0] / height, lon = 2 * M_PI * in[1] / width;
    double c0 = -sin(lat) * cos(lon), c1 = sin(lat) * sin(lon), c2 = cos(lat);       /* :71-73 */
    double r0 = R[0] * c0 + R[1] * c1 + R[2] * c2;                                      /* :77-79 */
    double r1 = R[3] * c0 + R[4] * c1 + R[5] * c2;
    double r2 = R[6] * c0 + R[7] * c1 + R[8] * c2;
    double a = acos(r2), b = atan2(r1, -r0);                                            /* :82-83 */
    if (b < 0) b += M_PI * 2;
    out[0] = (int32_t)(height * a / M_PI);                                              /* :88-89 */
    out[1] = (int32_t)(width * b / (2 * M_PI));
}

void orc_rotate_pixels(const int32_t* in, int n, const double* R, int width, int height, int32_t* out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) orc_rotate_pixel(in + 2 * (size_t)i, R, width, height, out + 2 * (size_t)i);
}

void orc_rotate_image(const uint8_t* im, int width, int height, const double* R, uint8_t* out)
{
    double Rinv[9];
    orc_inv3(R, Rinv);                                   /* :103 rot_mat.inv() */
    memset(out, 0, (size_t)width * height * 3);          /* the reference leaves unmapped pixels uninitialised */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width; j++) {
            int32_t in[2] = {i, j}, o[2];
            orc_rotate_pixel(in, Rinv, width, height, o);
            if (o[0] >= 0 && o[1] >= 0 && o[0] < height && o[1] < width)
                memcpy(out + ((size_t)i * width + j) * 3, im + ((size_t)o[0] * width + o[1]) * 3, 3);
        }
}

static void pitch_matrix(float pitch_deg, double* R)
{
    /* eular2rot(Vec3f(0, RAD(pitch), 0)): the angle passes through a float (spherical_surf.cpp:26,52) */
    float ang = (float)(M_PI * pitch_deg / 180.0);
    double th[3] = {0.0, (double)ang, 0.0};
    orc_eular2rot(th, R);
}

void orc_crop_rotated_image(const uint8_t* im, int width, int height, float pitch_rot_deg, uint8_t* out)
{
    double R[9];
    pitch_matrix(pitch_rot_deg, R);
    int oh = height / 4;
    memset(out, 0, (size_t)oh * width * 3);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < oh; i++)
        for (int j = 0; j < width; j++) {
            int32_t in[2] = {i + height * 3 / 8, j}, o[2];       /* :31 */
            orc_rotate_pixel(in, R, width, height, o);
            if (o[0] >= 0 && o[1] >= 0 && o[0] < height && o[1] < width)
                memcpy(out + ((size_t)i * width + j) * 3, im + ((size_t)o[0] * width + o[1]) * 3, 3);
        }
}

void orc_rotate_keypoints(void* xy, int stride_bytes, int n, float pitch_rot_inv_deg, int width, int height)
{
    double R[9];
    pitch_matrix(pitch_rot_inv_deg, R);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        float* p = (float*)((char*)xy + (size_t)i * stride_bytes);
        int32_t in[2], o[2];
        in[0] = (int32_t)(p[1] + height * 3 / 8);            /* :57  int offset_i = pt.y + height*3/8 (float sum) */
        in[1] = (int32_t)p[0];                               /* :58  Vec2i(offset_i, pt.x) */
        orc_rotate_pixel(in, R, width, height, o);
        p[0] = (float)o[1];                                  /* :61-62 */
        p[1] = (float)o[0];
    }
}

/* epipolar_tool.cpp:18-24 (BGR) */
static const uint8_t kEpiColor[7][3] = {{0, 0, 255}, {0, 127, 255}, {0, 255, 255}, {0, 255, 0}, {255, 0, 0}, {135, 0, 75}, {211, 0, 148}};

void orc_draw_epipole(const double* e, const void* left_xy, const void* right_xy, int stride_bytes, int n_key,
                      int im_w, int im_h, int out_w, int out_h, uint8_t* out)
{
    if (n_key > 7) n_key = 7;
    double l[7][3];
    int di[7], dj[7];
    double rw = (double)out_w / (double)im_w, rh = (double)out_h / (double)im_h;       /* :28-29 */
    for (int k = 0; k < n_key; k++) {
        const float* lp = (const float*)((const char*)left_xy + (size_t)k * stride_bytes);
        const float* rp = (const float*)((const char*)right_xy + (size_t)k * stride_bytes);
        double lon = 2 * M_PI * (lp[0] / im_w), lat = M_PI * (lp[1] / im_h);            /* :38-39 float quotient */
        l[k][0] = -sin(lat) * cos(lon); l[k][1] = sin(lat) * sin(lon); l[k][2] = cos(lat);
        di[k] = (int)(rp[1] * rh);                                                      /* :113-114 float * double */
        dj[k] = (int)(rp[0] * rw);
    }
    memset(out, 0, (size_t)out_w * out_h * 3);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < out_h; i++)
        for (int j = 0; j < out_w; j++) {
            double lon = 2 * M_PI * ((double)j / out_w), lat = M_PI * ((double)i / out_h);   /* :70-71 */
            double p[3] = {-sin(lat) * cos(lon), sin(lat) * sin(lon), cos(lat)};
            uint8_t* o = out + ((size_t)i * out_w + j) * 3;
            for (int k = 0; k < n_key; k++) {
                double result = l[k][0] * (p[0] * e[0] + p[1] * e[3] + p[2] * e[6])          /* :103-105 */
                              + l[k][1] * (p[0] * e[1] + p[1] * e[4] + p[2] * e[7])
                              + l[k][2] * (p[0] * e[2] + p[1] * e[5] + p[2] * e[8]);
                if (fabs(result) < 0.002) memcpy(o, kEpiColor[k], 3);
            }
            for (int k = 0; k < n_key; k++)                                                  /* :113-123, clipped */
                if (i >= di[k] - 5 && i <= di[k] + 5 && j >= dj[k] - 5 && j <= dj[k] + 5) memcpy(o, kEpiColor[k], 3);
        }
}

/* ------------------------------------------------------------------------- */
/* matching: src/feature_matcher.cpp:42-59                                   */
/* ------------------------------------------------------------------------- */
/*
 * knnMatch(desc1, desc2, out, 2) at feature_matcher.cpp:45 is FLANN (approximate,
 * rand()-seeded KD-trees).  The oracle is its exact limit: brute force over all
 * train rows.  Distance definition shared with the device:
 *   d2 = sum_k fma(diff_k, diff_k, acc), diff_k = (double)q_k - (double)t_k, k ascending
 *   order = (d2, trainIdx) ascending  (lowest trainIdx wins exact ties, as cv::BFMatcher)
 *   DMatch.distance = (float)sqrt(d2)
 */
#define TB 8 /* train rows per transposed block */

static float* transpose_blocks(const float* t, int nt, int dim, int* nblk_out)
{
    int nblk = (nt + TB - 1) / TB;
    float* tt = (float*)aligned_alloc(64, (size_t)nblk * dim * TB * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nblk; b++)
        for (int k = 0; k < dim; k++)
            for (int r = 0; r < TB; r++) {
                int row = b * TB + r;
                tt[((size_t)b * dim + k) * TB + r] = row < nt ? t[(size_t)row * dim + k] : 0.0f;
            }
    *nblk_out = nblk;
    return tt;
}

static inline void block_d2(const double* qd, const float* ttb, int dim, double acc[TB])
{
#if defined(__AVX2__) && defined(__FMA__)
    /* 8 train rows per step as two 4-wide fp64 lanes; per lane the chain is still the scalar
     * one (k ascending, one correctly rounded fma per step), so results equal the #else branch */
    __m256d a0 = _mm256_setzero_pd(), a1 = _mm256_setzero_pd();
    for (int k = 0; k < dim; k++) {
        __m256 row = _mm256_load_ps(ttb + (size_t)k * TB);
        __m256d qk = _mm256_set1_pd(qd[k]);
        __m256d d0 = _mm256_sub_pd(qk, _mm256_cvtps_pd(_mm256_castps256_ps128(row)));
        __m256d d1 = _mm256_sub_pd(qk, _mm256_cvtps_pd(_mm256_extractf128_ps(row, 1)));
        a0 = _mm256_fmadd_pd(d0, d0, a0);
        a1 = _mm256_fmadd_pd(d1, d1, a1);
    }
    _mm256_storeu_pd(acc, a0);
    _mm256_storeu_pd(acc + 4, a1);
#else
    for (int r = 0; r < TB; r++) acc[r] = 0.0;
    for (int k = 0; k < dim; k++) {
        const float* row = ttb + (size_t)k * TB;
        double qk = qd[k];
        for (int r = 0; r < TB; r++) {
            double diff = qk - (double)row[r];
            acc[r] = fma(diff, diff, acc[r]);
        }
    }
#endif
}

void orc_knn2(const float* q, int nq, const float* t, int nt, int dim,
              int32_t* idx, float* dist, double* d2out)
{
    int nblk;
    float* tt = transpose_blocks(t, nt, dim, &nblk);
#pragma omp parallel
    {
        double* qd = (double*)malloc(sizeof(double) * dim);
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < nq; i++) {
            for (int k = 0; k < dim; k++) qd[k] = (double)q[(size_t)i * dim + k];
            double b0 = INFINITY, b1 = INFINITY;
            int i0 = -1, i1 = -1;
            for (int b = 0; b < nblk; b++) {
                double acc[TB];
                block_d2(qd, tt + (size_t)b * dim * TB, dim, acc);
                int lim = nt - b * TB < TB ? nt - b * TB : TB;
                for (int r = 0; r < lim; r++) {
                    double d = acc[r];
                    if (d < b1) { /* ascending scan: strict < keeps the lowest index on ties */
                        if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = b * TB + r; }
                        else { b1 = d; i1 = b * TB + r; }
                    }
                }
            }
            idx[2 * i] = i0; idx[2 * i + 1] = i1;
            dist[2 * i] = (float)sqrt(b0); dist[2 * i + 1] = (float)sqrt(b1);
            if (d2out) { d2out[2 * i] = b0; d2out[2 * i + 1] = b1; }
        }
        free(qd);
    }
    free(tt);
}

void orc_nn1_reverse(const float* q, int nq, const float* t, int nt, int dim,
                     int32_t* best_q, double* best_d2)
{
    /* nearest QUERY for each train row: cv::BFMatcher(crossCheck=true) semantics */
    int nblk;
    float* qq = transpose_blocks(q, nq, dim, &nblk);
#pragma omp parallel
    {
        double* td = (double*)malloc(sizeof(double) * dim);
#pragma omp for schedule(dynamic, 16)
        for (int j = 0; j < nt; j++) {
            for (int k = 0; k < dim; k++) td[k] = (double)t[(size_t)j * dim + k];
            double b0 = INFINITY; int i0 = -1;
            for (int b = 0; b < nblk; b++) {
                double acc[TB];
                /* note (q-t)^2 == (t-q)^2 exactly, so operand order does not matter */
                block_d2(td, qq + (size_t)b * dim * TB, dim, acc);
                int lim = nq - b * TB < TB ? nq - b * TB : TB;
                for (int r = 0; r < lim; r++)
                    if (acc[r] < b0) { b0 = acc[r]; i0 = b * TB + r; }
            }
            best_q[j] = i0;
            if (best_d2) best_d2[j] = b0;
        }
        free(td);
    }
    free(qq);
}

int orc_match(const float* q, int nq, const float* t, int nt, int dim,
              float ratio, int cross_check, orc_dmatch* out)
{
    if (nq <= 0) return 0;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)nq);
    float* dist = (float*)malloc(sizeof(float) * 2 * (size_t)nq);
    int32_t* rev = NULL;
    orc_knn2(q, nq, t, nt, dim, idx, dist, NULL);
    if (cross_check) {
        rev = (int32_t*)malloc(sizeof(int32_t) * (size_t)nt);
        orc_nn1_reverse(q, nq, t, nt, dim, rev, NULL);
    }
    int n = 0;
    for (int i = 0; i < nq; i++) {
        /* feature_matcher.cpp:47,52: if (m[0].distance < 0.3f * m[1].distance) */
        float rhs = ratio * dist[2 * i + 1];
        int keep = ratio < 0.0f ? 1 : (dist[2 * i] < rhs);
        if (keep && cross_check) keep = rev[idx[2 * i]] == i;
        if (keep) {
            out[n].queryIdx = i; out[n].trainIdx = idx[2 * i];
            out[n].imgIdx = 0;   out[n].distance = dist[2 * i];
            n++;
        }
    }
    free(idx); free(dist); free(rev);
    return n;
}

/* ------------------------------------------------------------------------- */
/* geometry                                                                  */
/* ------------------------------------------------------------------------- */
/* src/eight_point.cpp:163-186: float quotient pt.x / im_width, promoted to double */
void orc_bearings(const void* xy, int stride_bytes, int n, int W, int H, double* out3)
{
    const char* base = (const char*)xy;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const float* p = (const float*)(base + (size_t)i * stride_bytes);
        float fx = p[0] / (float)W;      /* KeyPoint.pt.x (float) / int -> float */
        float fy = p[1] / (float)H;
        double lon = 2 * M_PI * fx;
        double lat = M_PI * fy;
        out3[3 * i + 0] = -sin(lat) * cos(lon);   /* MPEG OMAF axes */
        out3[3 * i + 1] = sin(lat) * sin(lon);
        out3[3 * i + 2] = cos(lat);
    }
}

static void mat3_mul(const double* A, const double* B, double* C)
{
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
            T[3 * i + j] = s;
        }
    memcpy(C, T, sizeof T);
}

/* src/erp_rotation.cpp:14-40: R = Rx * Ry * Rz */
void orc_eular2rot(const double* th, double* R)
{
    double Rx[9] = {1, 0, 0, 0, cos(th[0]), -sin(th[0]), 0, sin(th[0]), cos(th[0])};
    double Ry[9] = {cos(th[1]), 0, sin(th[1]), 0, 1, 0, -sin(th[1]), 0, cos(th[1])};
    double Rz[9] = {cos(th[2]), -sin(th[2]), 0, sin(th[2]), cos(th[2]), 0, 0, 0, 1};
    double T[9];
    mat3_mul(Rx, Ry, T);
    mat3_mul(T, Rz, R);
}

/* src/erp_rotation.cpp:43-63 */
void orc_rot2eular(const double* R, double* e)
{
    double sy = sqrt(R[8] * R[8] + R[5] * R[5]);
    int singular = sy < 1e-6;
    e[0] = singular ? 0.0 : atan2(-R[5], R[8]);
    e[1] = atan2(R[2], sy);
    e[2] = atan2(-R[1], R[0]);
}

/* ------------------------------------------------------------------------- */
/* OpenCV SVD (third party, restated): modules/core/src/lapack.cpp            */
/* ------------------------------------------------------------------------- */
static uint64_t cvrng_state;
static unsigned cvrng_next(void)
{
    cvrng_state = (uint64_t)(unsigned)cvrng_state * 4164903690U + (unsigned)(cvrng_state >> 32);
    return (unsigned)cvrng_state;
}

/* JacobiSVDImpl_<double>: At is n rows of length m (row stride m); W[n]; Vt n x n or NULL.
 * n1 = number of rows of At to normalise into left singular vectors. */
static void jacobi_svd(double* At, double* Wout, double* Vt, int m, int n, int n1)
{
    const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
    double* W = (double*)malloc(sizeof(double) * n);
    int i, j, k, iter, max_iter = m > 30 ? m : 30;
    double c, s, sd;

    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * m + k]; sd += t * t; }
        W[i] = sd;
        if (Vt) { for (k = 0; k < n; k++) Vt[i * n + k] = 0; Vt[i * n + i] = 1; }
    }
    for (iter = 0; iter < max_iter; iter++) {
        int changed = 0;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * m, *Aj = At + j * m;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = 1;
                if (Vt) {
                    double *Vi = Vt + i * n, *Vj = Vt + j * n;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0; Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) { double t = At[i * m + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            double tmp = W[i]; W[i] = W[j]; W[j] = tmp;
            if (Vt) {
                for (k = 0; k < m; k++) { tmp = At[i * m + k]; At[i * m + k] = At[j * m + k]; At[j * m + k] = tmp; }
                for (k = 0; k < n; k++) { tmp = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = tmp; }
            }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    if (Vt) {
        cvrng_state = 0x12345678;
        for (i = 0; i < n1; i++) {
            sd = i < n ? W[i] : 0;
            for (int ii = 0; ii < 100 && sd <= minval; ii++) {
                /* zero singular value: pseudo-random vector, Gram-Schmidt against previous rows */
                const double val0 = 1. / m;
                for (k = 0; k < m; k++) At[i * m + k] = (cvrng_next() & 256) != 0 ? val0 : -val0;
                for (iter = 0; iter < 2; iter++)
                    for (j = 0; j < i; j++) {
                        sd = 0;
                        for (k = 0; k < m; k++) sd += At[i * m + k] * At[j * m + k];
                        double asum = 0;
                        for (k = 0; k < m; k++) {
                            double t = At[i * m + k] - sd * At[j * m + k];
                            At[i * m + k] = t; asum += fabs(t);
                        }
                        asum = asum > eps * 100 ? 1 / asum : 0;
                        for (k = 0; k < m; k++) At[i * m + k] *= asum;
                    }
                sd = 0;
                for (k = 0; k < m; k++) { double t = At[i * m + k]; sd += t * t; }
                sd = sqrt(sd);
            }
            s = sd > minval ? 1 / sd : 0.;
            for (k = 0; k < m; k++) At[i * m + k] *= s;
        }
    }
    free(W);
}

/* cv::SVD::compute without FULL_UV (what cv::SVDecomp(src,w,u,vt) calls). */
void orc_svd(const double* A, int m0, int n0, double* w, double* u, double* vt)
{
    int m = m0, n = n0, at = 0;
    if (m < n) { int t = m; m = n; n = t; at = 1; }
    double* ta = (double*)malloc(sizeof(double) * (size_t)n * m);  /* n x m */
    double* tv = (double*)malloc(sizeof(double) * (size_t)n * n);
    if (!at) { for (int i = 0; i < m0; i++) for (int j = 0; j < n0; j++) ta[(size_t)j * m + i] = A[(size_t)i * n0 + j]; }
    else memcpy(ta, A, sizeof(double) * (size_t)m0 * n0);
    jacobi_svd(ta, w, tv, m, n, n);
    if (!at) {
        if (u) for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) u[(size_t)i * n + j] = ta[(size_t)j * m + i];
        if (vt) memcpy(vt, tv, sizeof(double) * n * n);
    } else {
        /* u = temp_v^T (m0 x k), vt = temp_u (k x n0) */
        if (u) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) u[i * n + j] = tv[j * n + i];
        if (vt) memcpy(vt, ta, sizeof(double) * (size_t)n * m);
    }
    free(ta); free(tv);
}

static double det3(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* cv::decomposeEssentialMat (modules/calib3d/src/five-point.cpp) */
void orc_decompose_essential(const double* E, double* R1, double* R2, double* t)
{
    double D[3], U[9], Vt[9];
    orc_svd(E, 3, 3, D, U, Vt);
    if (det3(U) < 0) for (int i = 0; i < 9; i++) U[i] = -U[i];
    if (det3(Vt) < 0) for (int i = 0; i < 9; i++) Vt[i] = -Vt[i];
    const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    const double Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double T[9];
    mat3_mul(U, Wm, T); mat3_mul(T, Vt, R1);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, R2);
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

/* ------------------------------------------------------------------------- */
/* eight-point: src/eight_point.cpp:16-85                                    */
/* ------------------------------------------------------------------------- */
static double max_vec3(const float* v) /* eight_point.cpp:6-14 */
{
    if ((v[0] > v[1]) && (v[0] > v[2])) return v[0];
    else if (v[1] > v[2]) return v[1];
    else return v[2];
}

void orc_eight_point(const double* l3, const double* r3, int n, int null_mode,
                     double* e9, double* Ec9, float* R1e, float* R2e, float* T,
                     int* R1_valid, int* R2_valid)
{
    int rows = n;
    if (null_mode && rows < 9) rows = 9;           /* zero-padded rows keep the null space */
    double* A = (double*)calloc((size_t)rows * 9, sizeof(double));
    for (int i = 0; i < n; i++)                    /* :26-37  A[i] = kron(l_i, r_i) */
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) A[(size_t)i * 9 + 3 * a + b] = l3[3 * i + a] * r3[3 * i + b];
    int k = rows < 9 ? rows : 9;
    double w[9];
    double* vt = (double*)malloc(sizeof(double) * (size_t)k * 9);
    orc_svd(A, rows, 9, w, NULL, vt);              /* :39 */
    const double* e = vt + (size_t)(k - 1) * 9;    /* :42 vt.row(vt.rows-1) */
    if (e9) memcpy(e9, e, sizeof(double) * 9);

    double wf[3], uf[9], vtf[9];                   /* :45-50 */
    orc_svd(e, 3, 3, wf, uf, vtf);
    wf[2] = 0.0;
    double D[9] = {wf[0], 0, 0, 0, wf[1], 0, 0, 0, wf[2]}, Tm[9], Ec[9];
    mat3_mul(uf, D, Tm); mat3_mul(Tm, vtf, Ec);
    if (Ec9) memcpy(Ec9, Ec, sizeof Ec);

    double R1[9], R2[9], t[3], e1[3], e2[3];       /* :53-61 */
    orc_decompose_essential(Ec, R1, R2, t);
    orc_rot2eular(R1, e1); orc_rot2eular(R2, e2);
    float f1[3], f2[3];
    for (int i = 0; i < 3; i++) { f1[i] = (float)e1[i]; f2[i] = (float)e2[i]; }
    if (R1e) memcpy(R1e, f1, sizeof f1);
    if (R2e) memcpy(R2e, f2, sizeof f2);
    if (T) for (int i = 0; i < 3; i++) T[i] = (float)t[i];
    float a1[3] = {fabsf(f1[0]), fabsf(f1[1]), fabsf(f1[2])};   /* :72-84 */
    float a2[3] = {fabsf(f2[0]), fabsf(f2[1]), fabsf(f2[2])};
    if (R1_valid) *R1_valid = max_vec3(a1) < 1.57;
    if (R2_valid) *R2_valid = max_vec3(a2) < 1.57;
    free(A); free(vt);
}

/* ------------------------------------------------------------------------- */
/* sampling                                                                  */
/* ------------------------------------------------------------------------- */
/* src/eight_point.hpp:54-58: iota + std::random_shuffle.  libstdc++'s two-iterator
 * random_shuffle is  for i in 1..n-1: j = rand() % (i+1); if (i != j) swap(a[i], a[j])
 * (bits/stl_algo.h), driven by glibc rand(), which the reference never seeds. */
void orc_random_array(int size, int32_t* out, unsigned reseed)
{
    if (reseed) srand(reseed);
    for (int i = 0; i < size; i++) out[i] = i;
    for (int i = 1; i < size; i++) {
        int j = rand() % (i + 1);
        if (i != j) { int32_t t = out[i]; out[i] = out[j]; out[j] = t; }
    }
}

/* src/eight_point.cpp:99-111: a fresh random_array(M) per iteration, first S entries */
void orc_ref_sample_table(int M, int H, int S, int32_t* table, unsigned reseed)
{
    int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    for (int h = 0; h < H; h++) {
        orc_random_array(M, perm, h == 0 ? reseed : 0);
        for (int s = 0; s < S; s++) table[(size_t)h * S + s] = perm[s % (M > 0 ? M : 1)];
    }
    free(perm);
}

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* Replaces random_array for minimal samples.  Spec shared with the device:
 * counter = (hyp_lo, hyp_hi, block, 'ERP8'), key = (seed_lo, seed_hi); each 32-bit word r
 * proposes idx = (r * M) >> 32; duplicates of earlier picks are skipped. */
void orc_philox_samples(uint64_t seed, uint64_t hyp_id, int M, int S, int32_t* out)
{
    int count = 0;
    uint32_t block = 0;
    if (M < S) { for (int i = 0; i < S; i++) out[i] = -1; return; }
    while (count < S) {
        uint32_t c[4] = {(uint32_t)hyp_id, (uint32_t)(hyp_id >> 32), block++, 0x45525038u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        for (int w = 0; w < 4 && count < S; w++) {
            int32_t idx = (int32_t)(((uint64_t)c[w] * (uint32_t)M) >> 32);
            int dup = 0;
            for (int j = 0; j < count; j++) dup |= out[j] == idx;
            if (!dup) out[count++] = idx;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* initial_guess: src/eight_point.cpp:87-150                                 */
/* ------------------------------------------------------------------------- */
static int cmp_double(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* :131-149.  returns arg-min index (first minimum, like std::min_element) */
int orc_consensus_pick(const float* R, int C, double* trimmed_mean)
{
    if (C <= 0) return -1;
    double* dist = (double*)malloc(sizeof(double) * C);
    double* dn = (double*)malloc(sizeof(double) * C);
    for (int i = 0; i < C; i++) {
        for (int j = 0; j < C; j++) {
            float d0 = R[3 * i] - R[3 * j], d1 = R[3 * i + 1] - R[3 * j + 1], d2 = R[3 * i + 2] - R[3 * j + 2];
            /* Vec3f arithmetic: float products and sums, sqrt(float) -> float -> double */
            float ss = d0 * d0 + d1 * d1 + d2 * d2;
            dn[j] = (double)sqrtf(ss);
        }
        qsort(dn, C, sizeof(double), cmp_double);
        int lo = (int)(C * 0.2), hi = (int)(C * 0.8);
        double acc = 0.0;
        for (int j = lo; j < hi; j++) acc += dn[j];
        dist[i] = acc / ((hi - lo) * 1.0);
    }
    int best = 0;
    for (int i = 1; i < C; i++) if (dist[i] < dist[best]) best = i;
    if (trimmed_mean) memcpy(trimmed_mean, dist, sizeof(double) * C);
    free(dist); free(dn);
    return best;
}

int orc_initial_guess(const double* l3, const double* r3, int M,
                      const int32_t* table, int H, int S, int null_mode,
                      float* R_out, float* T_out,
                      float* cand_R, float* cand_T, int* n_cand, int* chosen)
{
    float* cr = (float*)malloc(sizeof(float) * 6 * (size_t)H);
    float* ct = (float*)malloc(sizeof(float) * 6 * (size_t)H);
    float* res = (float*)malloc(sizeof(float) * 9 * (size_t)H);
    int* val = (int*)malloc(sizeof(int) * 2 * (size_t)H);
    (void)M;
#pragma omp parallel for schedule(dynamic, 1)
    for (int h = 0; h < H; h++) {
        double* ls = (double*)malloc(sizeof(double) * 3 * (size_t)S);
        double* rs = (double*)malloc(sizeof(double) * 3 * (size_t)S);
        for (int s = 0; s < S; s++) {                       /* :105-111 */
            int idx = table[(size_t)h * S + s];
            memcpy(ls + 3 * s, l3 + 3 * (size_t)idx, 3 * sizeof(double));
            memcpy(rs + 3 * s, r3 + 3 * (size_t)idx, 3 * sizeof(double));
        }
        orc_eight_point(ls, rs, S, null_mode, NULL, NULL, res + 9 * h, res + 9 * h + 3,
                        res + 9 * h + 6, val + 2 * h, val + 2 * h + 1);
        free(ls); free(rs);
    }
    int C = 0;
    for (int h = 0; h < H; h++) {                           /* :117-126 */
        if (val[2 * h]) { memcpy(cr + 3 * C, res + 9 * h, 12); memcpy(ct + 3 * C, res + 9 * h + 6, 12); C++; }
        if (val[2 * h + 1]) { memcpy(cr + 3 * C, res + 9 * h + 3, 12); memcpy(ct + 3 * C, res + 9 * h + 6, 12); C++; }
    }
    int rc = 1, best = -1;
    if (C > 0) {
        best = orc_consensus_pick(cr, C, NULL);
        memcpy(R_out, cr + 3 * best, 12); memcpy(T_out, ct + 3 * best, 12);
        rc = 0;
    }
    if (cand_R) memcpy(cand_R, cr, sizeof(float) * 3 * C);
    if (cand_T) memcpy(cand_T, ct, sizeof(float) * 3 * C);
    if (n_cand) *n_cand = C;
    if (chosen) *chosen = best;
    free(cr); free(ct); free(res); free(val);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* scoring: residual of src/epipolar_tool.cpp:100-107                        */
/* ------------------------------------------------------------------------- */
/*
 * epipolar_tool evaluates result = l . (E_tool^T p) with |result| < 0.002 and
 * E_tool = R^-1 [t]x (||E||_F = sqrt 2).  Under eight_point's convention
 * (l^T E r = 0, eight_point.cpp:28-36) the same scalar is l^T E r.
 * Spec shared bit-for-bit with the device (fp32, explicit fma chain):
 *   Eh   = (float)(E * sqrt(2)/||E||_F)
 *   k_ab = l_a * r_b                                  (fp32 products, a,b in 0..2)
 *   res  = fma(Eh8,k8, ... fma(Eh1,k1, Eh0*k0))       (ascending index)
 *   ALGEBRAIC: |res| < tau
 *   SAMPSON  : res^2 < (tau*tau) * (|Eh r|^2 + |Eh^T l|^2)
 *   ANGULAR  : res^2 < sin(tau)^2 * |Eh r|^2          (angle of l to the epipolar plane)
 */
static void scale_E(const double* E, float* Eh)
{
    double n = 0;
    for (int i = 0; i < 9; i++) n += E[i] * E[i];
    double s = n > 0 ? sqrt(2.0) / sqrt(n) : 0.0;
    for (int i = 0; i < 9; i++) Eh[i] = (float)(E[i] * s);
}

static inline int is_inlier(const float* Eh, const float* l, const float* r, int metric,
                            float tau, float tau2, float sin2)
{
    float k[9];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) k[3 * a + b] = l[a] * r[b];
    float res = Eh[0] * k[0];
    for (int i = 1; i < 9; i++) res = fmaf(Eh[i], k[i], res);
    if (metric == ORC_METRIC_ALGEBRAIC) return fabsf(res) < tau;
    float n0 = fmaf(Eh[2], r[2], fmaf(Eh[1], r[1], Eh[0] * r[0]));
    float n1 = fmaf(Eh[5], r[2], fmaf(Eh[4], r[1], Eh[3] * r[0]));
    float n2 = fmaf(Eh[8], r[2], fmaf(Eh[7], r[1], Eh[6] * r[0]));
    float nn = fmaf(n2, n2, fmaf(n1, n1, n0 * n0));
    float rr = res * res;
    if (metric == ORC_METRIC_ANGULAR) return rr < sin2 * nn;
    float m0 = fmaf(Eh[6], l[2], fmaf(Eh[3], l[1], Eh[0] * l[0]));
    float m1 = fmaf(Eh[7], l[2], fmaf(Eh[4], l[1], Eh[1] * l[0]));
    float m2 = fmaf(Eh[8], l[2], fmaf(Eh[5], l[1], Eh[2] * l[0]));
    float mm = fmaf(m2, m2, fmaf(m1, m1, m0 * m0));
    return rr < tau2 * (nn + mm);
}

void orc_score(const double* E, int H, const float* l4, const float* r4, int m,
               int metric, float tau, int32_t* counts)
{
    float tau2 = tau * tau;
    double sd = sin((double)tau);
    float sin2 = (float)(sd * sd);
#pragma omp parallel for schedule(static)
    for (int h = 0; h < H; h++) {
        float Eh[9];
        scale_E(E + 9 * (size_t)h, Eh);
        int c = 0;
        for (int i = 0; i < m; i++) c += is_inlier(Eh, l4 + 4 * (size_t)i, r4 + 4 * (size_t)i, metric, tau, tau2, sin2);
        counts[h] = c;
    }
}

void orc_inlier_mask(const double* E9, const float* l4, const float* r4, int m,
                     int metric, float tau, uint8_t* mask)
{
    float Eh[9];
    float tau2 = tau * tau;
    double sd = sin((double)tau);
    float sin2 = (float)(sd * sd);
    scale_E(E9, Eh);
    for (int i = 0; i < m; i++) mask[i] = (uint8_t)is_inlier(Eh, l4 + 4 * (size_t)i, r4 + 4 * (size_t)i, metric, tau, tau2, sin2);
}

uint64_t orc_ransac(const double* l3, const double* r3, int M, uint64_t seed,
                    uint64_t hyp0, int H, int S, int metric, float tau,
                    double* best_E9, int32_t* counts_out)
{
    float* l4 = (float*)malloc(sizeof(float) * 4 * (size_t)M);
    float* r4 = (float*)malloc(sizeof(float) * 4 * (size_t)M);
    for (int i = 0; i < M; i++) {
        for (int a = 0; a < 3; a++) { l4[4 * i + a] = (float)l3[3 * i + a]; r4[4 * i + a] = (float)r3[3 * i + a]; }
        l4[4 * i + 3] = r4[4 * i + 3] = 0.0f;
    }
    double* E = (double*)malloc(sizeof(double) * 9 * (size_t)H);
    int32_t* counts = (int32_t*)malloc(sizeof(int32_t) * (size_t)H);
#pragma omp parallel for schedule(dynamic, 64)
    for (int h = 0; h < H; h++) {
        int32_t smp[64];
        double ls[3 * 64], rs[3 * 64];
        int s_eff = S > 64 ? 64 : S;
        orc_philox_samples(seed, hyp0 + (uint64_t)h, M, s_eff, smp);
        for (int s = 0; s < s_eff; s++) {
            memcpy(ls + 3 * s, l3 + 3 * (size_t)smp[s], 24);
            memcpy(rs + 3 * s, r3 + 3 * (size_t)smp[s], 24);
        }
        orc_eight_point(ls, rs, s_eff, 1, NULL, E + 9 * (size_t)h, NULL, NULL, NULL, NULL, NULL);
    }
    orc_score(E, H, l4, r4, M, metric, tau, counts);
    uint64_t best = 0;
    int bh = 0;
    for (int h = 0; h < H; h++) {
        uint64_t id = hyp0 + (uint64_t)h;
        uint64_t packed = ((uint64_t)(uint32_t)counts[h] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)id);
        if (h == 0 || packed > best) { best = packed; bh = h; }
    }
    if (best_E9) memcpy(best_E9, E + 9 * (size_t)bh, 72);
    if (counts_out) memcpy(counts_out, counts, sizeof(int32_t) * (size_t)H);
    free(l4); free(r4); free(E); free(counts);
    return best;
}
