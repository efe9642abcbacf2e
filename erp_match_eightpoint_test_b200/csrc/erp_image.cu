// erp_image.cu -- the rows next to the hot path: ERP pixel rotation, image warp,
// strip cropping and keypoint back-rotation.
//
//   erp_rotation::rotate_pixel          /root/reference/src/erp_rotation.cpp:66-92
//   erp_rotation::rotate_image          /root/reference/src/erp_rotation.cpp:94-122  (nearest-neighbour inverse warp)
//   spherical_surf::crop_rotated_image  /root/reference/src/spherical_surf.cpp:16-48
//   spherical_surf::rotate_keypoint     /root/reference/src/spherical_surf.cpp:50-63
//
// The map is the reference's fp64 chain (sin/cos -> 3x3 -> acos/atan2 -> truncation), with separate
// multiplies and adds as the CPU code is compiled (the library builds with --fmad=false).  sin/cos
// of the latitude depend on the row only and of the longitude on the column only: a block computes
// them once for its rows and columns.  The warp is therefore bound by the fp64 acos/atan2 per pixel
// (about 150 fp64 instructions against 6 bytes of traffic), not by HBM.
// CUDA's fp64 sin/cos/acos/atan2 are not bit-identical to glibc's: a coordinate that lands within an
// ulp of an integer can truncate differently (measure zero for generic rotations; the identity and
// axis-aligned rotations hit integers exactly and are libm-dependent in the reference itself).
#include "common.cuh"

namespace erp {

struct Rot9 { double m[9]; };

__device__ __forceinline__ void finish_pixel(double c0, double c1, double c2, const Rot9& R, int W, int H, int& orow, int& ocol)
{
    const double PI = 3.14159265358979323846;
    double r0 = R.m[0] * c0 + R.m[1] * c1 + R.m[2] * c2;
    double r1 = R.m[3] * c0 + R.m[4] * c1 + R.m[5] * c2;
    double r2 = R.m[6] * c0 + R.m[7] * c1 + R.m[8] * c2;
    double a = acos(r2), b = atan2(r1, -r0);
    if (b < 0) b += PI * 2;
    orow = (int)(H * a / PI);
    ocol = (int)(W * b / (2 * PI));
}

__device__ __forceinline__ void rotate_pixel_dev(int row, int col, const Rot9& R, int W, int H, int& orow, int& ocol)
{
    const double PI = 3.14159265358979323846;
    double lat = PI * row / H, lon = 2 * PI * col / W;
    double sl, cl, so, co;
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
    finish_pixel(-sl * co, sl * so, cl, R, W, H, orow, ocol);
}

__global__ void rotate_pixels_kernel(const int32_t* __restrict__ rc, int n, Rot9 R, int W, int H, int32_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r, c;
    rotate_pixel_dev(rc[2 * i], rc[2 * i + 1], R, W, H, r, c);
    out[2 * i] = r; out[2 * i + 1] = c;
}

__global__ void rotate_keypoints_kernel(char* __restrict__ xy, size_t stride, int n, Rot9 R, int W, int H)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* p = reinterpret_cast<float*>(xy + (size_t)i * stride);
    int row = (int)__fadd_rn(p[1], (float)(H * 3 / 8));      // int offset_i = pt.y + height*3/8
    int col = (int)p[0];
    int r, c;
    rotate_pixel_dev(row, col, R, W, H, r, c);
    p[0] = (float)c; p[1] = (float)r;
}

// out[i][j] = im[map(i + row_offset, j)] for i < out_rows; unmapped pixels are written 0
constexpr int WT_X = 128, WT_Y = 8;
__global__ void __launch_bounds__(WT_X* WT_Y)
warp_image_kernel(const uint8_t* __restrict__ im, int W, int H, size_t stride, Rot9 R, int row_offset, int out_rows,
                  uint8_t* __restrict__ out, size_t ostride)
{
    __shared__ double s_lat[WT_Y][2], s_lon[WT_X][2];
    const double PI = 3.14159265358979323846;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = blockIdx.x * WT_X + tx, i = blockIdx.y * WT_Y + ty;
    if (ty == 0) {
        double s, c;
        sincos(2 * PI * j / W, &s, &c);
        s_lon[tx][0] = s; s_lon[tx][1] = c;
    } else if (ty == 1 && tx < WT_Y) {
        double s, c;
        sincos(PI * (blockIdx.y * WT_Y + tx + row_offset) / H, &s, &c);
        s_lat[tx][0] = s; s_lat[tx][1] = c;
    }
    __syncthreads();
    if (i >= out_rows || j >= W) return;
    const double sl = s_lat[ty][0], cl = s_lat[ty][1], so = s_lon[tx][0], co = s_lon[tx][1];
    int r, c;
    finish_pixel(-sl * co, sl * so, cl, R, W, H, r, c);
    uint8_t* o = out + (size_t)i * ostride + (size_t)j * 3;
    if (r >= 0 && c >= 0 && r < H && c < W) {
        const uint8_t* s = im + (size_t)r * stride + (size_t)c * 3;
        o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
    } else { o[0] = 0; o[1] = 0; o[2] = 0; }
}

// ---- epipolar_tool::draw_epipole (/root/reference/src/epipolar_tool.cpp:84-128), one thread per pixel.
// Sequential semantics of the reference loop (its collapse(3) pragma races): the last key whose
// residual passes colours the pixel, then the 11 x 11 dots are painted in key order, clipped.
struct EpiParams {
    double e[9];
    double l[7][3];
    int di[7], dj[7];
    int n_key;
};
__constant__ uint8_t kEpiColor[7][3] = {{0, 0, 255}, {0, 127, 255}, {0, 255, 255}, {0, 255, 0}, {255, 0, 0}, {135, 0, 75}, {211, 0, 148}};

__global__ void __launch_bounds__(WT_X* WT_Y)
draw_epipole_kernel(EpiParams P, int out_w, int out_h, uint8_t* __restrict__ out, size_t ostride)
{
    __shared__ double s_lat[WT_Y][2], s_lon[WT_X][2];
    const double PI = 3.14159265358979323846;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int j = blockIdx.x * WT_X + tx, i = blockIdx.y * WT_Y + ty;
    if (ty == 0) {
        double s, c;
        sincos(2 * PI * ((double)j / out_w), &s, &c);
        s_lon[tx][0] = s; s_lon[tx][1] = c;
    } else if (ty == 1 && tx < WT_Y) {
        double s, c;
        sincos(PI * ((double)(blockIdx.y * WT_Y + tx) / out_h), &s, &c);
        s_lat[tx][0] = s; s_lat[tx][1] = c;
    }
    __syncthreads();
    if (i >= out_h || j >= out_w) return;
    const double sl = s_lat[ty][0], cl = s_lat[ty][1], so = s_lon[tx][0], co = s_lon[tx][1];
    const double p0 = -sl * co, p1 = sl * so, p2 = cl;
    const double* e = P.e;
    const double q0 = p0 * e[0] + p1 * e[3] + p2 * e[6];
    const double q1 = p0 * e[1] + p1 * e[4] + p2 * e[7];
    const double q2 = p0 * e[2] + p1 * e[5] + p2 * e[8];
    int col = -1;
    for (int k = 0; k < P.n_key; k++) {
        double result = P.l[k][0] * q0 + P.l[k][1] * q1 + P.l[k][2] * q2;
        if (fabs(result) < 0.002) col = k;
    }
    for (int k = 0; k < P.n_key; k++)
        if (i >= P.di[k] - 5 && i <= P.di[k] + 5 && j >= P.dj[k] - 5 && j <= P.dj[k] + 5) col = k;
    uint8_t* o = out + (size_t)i * ostride + (size_t)j * 3;
    if (col >= 0) { o[0] = kEpiColor[col][0]; o[1] = kEpiColor[col][1]; o[2] = kEpiColor[col][2]; }
    else { o[0] = 0; o[1] = 0; o[2] = 0; }
}

// cv::Mat::inv() for 3x3 CV_64F (closed form of OpenCV's lapack.cpp), evaluated on the host
static bool inv3_host(const double* S, double* t)
{
    volatile double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.0) return false;
    double q = 1.0 / d;
    t[0] = (S[4] * S[8] - S[5] * S[7]) * q; t[1] = (S[2] * S[7] - S[1] * S[8]) * q; t[2] = (S[1] * S[5] - S[2] * S[4]) * q;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * q; t[4] = (S[0] * S[8] - S[2] * S[6]) * q; t[5] = (S[2] * S[3] - S[0] * S[5]) * q;
    t[6] = (S[3] * S[7] - S[4] * S[6]) * q; t[7] = (S[1] * S[6] - S[0] * S[7]) * q; t[8] = (S[0] * S[4] - S[1] * S[3]) * q;
    return true;
}

// eular2rot(Vec3f(0, RAD(pitch), 0)): R = Ry(angle), the angle rounded through a float as the reference does
static void pitch_matrix(float pitch_deg, Rot9& R)
{
    float ang = (float)(3.14159265358979323846 * pitch_deg / 180.0);
    double c = cos((double)ang), s = sin((double)ang);
    // Rx(0) * Ry * Rz(0) evaluated like erp_rotation.cpp:14-40 (products with exact 0 and 1)
    double m[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
    for (int i = 0; i < 9; i++) R.m[i] = m[i] + 0.0;       // normalise -0.0
}

static int launch_warp(erp_ctx* ctx, const uint8_t* d_im, int W, int H, size_t stride, const Rot9& R, int row_offset, int out_rows,
                       uint8_t* d_out, size_t ostride)
{
    if (out_rows <= 0 || W <= 0) return ERP_OK;
    dim3 grid(cdiv(W, WT_X), cdiv(out_rows, WT_Y)), block(WT_X, WT_Y);
    warp_image_kernel<<<grid, block, 0, ctx->stream>>>(d_im, W, H, stride, R, row_offset, out_rows, d_out, ostride);
    ERP_LAUNCH(ctx, "warp_image_kernel");
    return ERP_OK;
}

} // namespace erp

using namespace erp;

ERP_API int erp_rotate_image_dev(erp_ctx* ctx, const uint8_t* d_im, int width, int height, size_t stride_bytes,
                                 const double* R9, uint8_t* d_out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && d_im && d_out && R9 && width > 0 && height > 0, ERP_E_ARG, "erp_rotate_image_dev: bad argument");
    ERP_ARG(stride_bytes >= (size_t)width * 3 && out_stride_bytes >= (size_t)width * 3, ERP_E_ARG, "erp_rotate_image_dev: stride smaller than a row");
    Rot9 Rinv;
    ERP_ARG(inv3_host(R9, Rinv.m), ERP_E_ARG, "erp_rotate_image_dev: singular rotation matrix");     // rot_mat.inv(): inverse mapping
    DeviceGuard g(ctx->device);
    return launch_warp(ctx, d_im, width, height, stride_bytes, Rinv, 0, height, d_out, out_stride_bytes);
}

ERP_API int erp_crop_rotated_image_dev(erp_ctx* ctx, const uint8_t* d_im, int width, int height, size_t stride_bytes,
                                       float pitch_rot_deg, uint8_t* d_out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && d_im && d_out && width > 0 && height >= 4, ERP_E_ARG, "erp_crop_rotated_image_dev: bad argument");
    ERP_ARG(stride_bytes >= (size_t)width * 3 && out_stride_bytes >= (size_t)width * 3, ERP_E_ARG, "erp_crop_rotated_image_dev: stride smaller than a row");
    Rot9 R;
    pitch_matrix(pitch_rot_deg, R);
    DeviceGuard g(ctx->device);
    return launch_warp(ctx, d_im, width, height, stride_bytes, R, height * 3 / 8, height / 4, d_out, out_stride_bytes);
}

static int image_host(erp_ctx* ctx, const uint8_t* im, int width, int height, size_t stride, uint8_t* out, size_t ostride, int out_rows,
                      uint8_t** d_im, uint8_t** d_out)
{
    int st = ERP_OK;
    *d_im = ctx->scratch<uint8_t>(S_IMG_IN, (size_t)height * width * 3, &st);
    *d_out = ctx->scratch<uint8_t>(S_IMG_OUT, (size_t)out_rows * width * 3, &st);
    ERP_TRY(st);
    ERP_CUDA(cudaMemcpy2DAsync(*d_im, (size_t)width * 3, im, stride, (size_t)width * 3, height, cudaMemcpyHostToDevice, ctx->stream));
    (void)out; (void)ostride;
    return ERP_OK;
}

ERP_API int erp_rotate_image(erp_ctx* ctx, const uint8_t* im, int width, int height, size_t stride_bytes,
                             const double* R9, uint8_t* out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && im && out && R9 && width > 0 && height > 0, ERP_E_ARG, "erp_rotate_image: bad argument");
    ERP_ARG(stride_bytes >= (size_t)width * 3 && out_stride_bytes >= (size_t)width * 3, ERP_E_ARG, "erp_rotate_image: stride smaller than a row");
    DeviceGuard g(ctx->device);
    uint8_t *d_im, *d_out;
    ERP_TRY(image_host(ctx, im, width, height, stride_bytes, out, out_stride_bytes, height, &d_im, &d_out));
    ERP_TRY(erp_rotate_image_dev(ctx, d_im, width, height, (size_t)width * 3, R9, d_out, (size_t)width * 3));
    ERP_CUDA(cudaMemcpy2DAsync(out, out_stride_bytes, d_out, (size_t)width * 3, (size_t)width * 3, height, cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}

ERP_API int erp_crop_rotated_image(erp_ctx* ctx, const uint8_t* im, int width, int height, size_t stride_bytes,
                                   float pitch_rot_deg, uint8_t* out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && im && out && width > 0 && height >= 4, ERP_E_ARG, "erp_crop_rotated_image: bad argument");
    ERP_ARG(stride_bytes >= (size_t)width * 3 && out_stride_bytes >= (size_t)width * 3, ERP_E_ARG, "erp_crop_rotated_image: stride smaller than a row");
    DeviceGuard g(ctx->device);
    uint8_t *d_im, *d_out;
    ERP_TRY(image_host(ctx, im, width, height, stride_bytes, out, out_stride_bytes, height / 4, &d_im, &d_out));
    ERP_TRY(erp_crop_rotated_image_dev(ctx, d_im, width, height, (size_t)width * 3, pitch_rot_deg, d_out, (size_t)width * 3));
    ERP_CUDA(cudaMemcpy2DAsync(out, out_stride_bytes, d_out, (size_t)width * 3, (size_t)width * 3, height / 4, cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}

ERP_API int erp_rotate_pixels(erp_ctx* ctx, const int32_t* rc, int n, const double* R9, int width, int height, int32_t* out)
{
    ERP_ARG(ctx && n >= 0 && R9 && width > 0 && height > 0, ERP_E_ARG, "erp_rotate_pixels: bad argument");
    if (n == 0) return ERP_OK;
    ERP_ARG(rc && out, ERP_E_ARG, "erp_rotate_pixels: null buffer");
    DeviceGuard g(ctx->device);
    int st = ERP_OK;
    int32_t* d_in = ctx->scratch<int32_t>(S_IMG_IN, (size_t)n * 2, &st);
    int32_t* d_o = ctx->scratch<int32_t>(S_IMG_OUT, (size_t)n * 2, &st);
    ERP_TRY(st);
    Rot9 R;
    memcpy(R.m, R9, sizeof R.m);
    ERP_CUDA(cudaMemcpyAsync(d_in, rc, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rotate_pixels_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_in, n, R, width, height, d_o);
    ERP_LAUNCH(ctx, "rotate_pixels_kernel");
    ERP_CUDA(cudaMemcpyAsync(out, d_o, sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}

ERP_API int erp_rotate_keypoints_dev(erp_ctx* ctx, void* d_xy, size_t stride_bytes, int n, float pitch_rot_inv_deg, int width, int height)
{
    ERP_ARG(ctx && n >= 0 && width > 0 && height > 0 && stride_bytes >= 8 && stride_bytes % 4 == 0, ERP_E_ARG, "erp_rotate_keypoints_dev: bad argument");
    if (n == 0) return ERP_OK;
    ERP_ARG(d_xy, ERP_E_ARG, "erp_rotate_keypoints_dev: null buffer");
    DeviceGuard g(ctx->device);
    Rot9 R;
    pitch_matrix(pitch_rot_inv_deg, R);
    rotate_keypoints_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>((char*)d_xy, stride_bytes, n, R, width, height);
    ERP_LAUNCH(ctx, "rotate_keypoints_kernel");
    return ERP_OK;
}

ERP_API int erp_rotate_keypoints(erp_ctx* ctx, void* xy, size_t stride_bytes, int n, float pitch_rot_inv_deg, int width, int height)
{
    ERP_ARG(ctx && n >= 0 && stride_bytes >= 8 && stride_bytes % 4 == 0, ERP_E_ARG, "erp_rotate_keypoints: bad argument");
    if (n == 0) return ERP_OK;
    ERP_ARG(xy, ERP_E_ARG, "erp_rotate_keypoints: null buffer");
    DeviceGuard g(ctx->device);
    int st = ERP_OK;
    float* d = ctx->scratch<float>(S_XY, (size_t)n * 2, &st);
    ERP_TRY(st);
    ERP_CUDA(cudaMemcpy2DAsync(d, 8, xy, stride_bytes, 8, n, cudaMemcpyHostToDevice, ctx->stream));
    ERP_TRY(erp_rotate_keypoints_dev(ctx, d, 8, n, pitch_rot_inv_deg, width, height));
    ERP_CUDA(cudaMemcpy2DAsync(xy, stride_bytes, d, 8, 8, n, cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}

// left_xy / right_xy: the n_key (<= 7) selected correspondences, (x, y) float pairs on the HOST
ERP_API int erp_draw_epipole_dev(erp_ctx* ctx, const double* E9, const void* left_xy, const void* right_xy, size_t stride_bytes,
                                 int n_key, int im_width, int im_height, int out_width, int out_height,
                                 uint8_t* d_out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && E9 && d_out && n_key >= 0 && n_key <= 7 && im_width > 0 && im_height > 0 && out_width > 0 && out_height > 0,
            ERP_E_ARG, "erp_draw_epipole_dev: bad argument (at most 7 keys: the reference has 7 colours)");
    ERP_ARG(n_key == 0 || (left_xy && right_xy && stride_bytes >= 8), ERP_E_ARG, "erp_draw_epipole_dev: null keypoints");
    ERP_ARG(out_stride_bytes >= (size_t)out_width * 3, ERP_E_ARG, "erp_draw_epipole_dev: stride smaller than a row");
    EpiParams P;
    memcpy(P.e, E9, sizeof P.e);
    P.n_key = n_key;
    const double PI = 3.14159265358979323846;
    const double rw = (double)out_width / (double)im_width, rh = (double)out_height / (double)im_height;
    for (int k = 0; k < 7; k++) { P.di[k] = P.dj[k] = -1000; P.l[k][0] = P.l[k][1] = P.l[k][2] = 0.0; }
    for (int k = 0; k < n_key; k++) {
        const float* lp = (const float*)((const char*)left_xy + (size_t)k * stride_bytes);
        const float* rp = (const float*)((const char*)right_xy + (size_t)k * stride_bytes);
        // epipolar_tool.cpp:38-45: float quotient, fp64 trig (host libm, as the reference's constructor)
        volatile double lon = 2 * PI * (lp[0] / im_width), lat = PI * (lp[1] / im_height);
        P.l[k][0] = -sin(lat) * cos(lon); P.l[k][1] = sin(lat) * sin(lon); P.l[k][2] = cos(lat);
        P.di[k] = (int)(rp[1] * rh);
        P.dj[k] = (int)(rp[0] * rw);
    }
    DeviceGuard g(ctx->device);
    dim3 grid(cdiv(out_width, WT_X), cdiv(out_height, WT_Y)), block(WT_X, WT_Y);
    draw_epipole_kernel<<<grid, block, 0, ctx->stream>>>(P, out_width, out_height, d_out, out_stride_bytes);
    ERP_LAUNCH(ctx, "draw_epipole_kernel");
    return ERP_OK;
}

ERP_API int erp_draw_epipole(erp_ctx* ctx, const double* E9, const void* left_xy, const void* right_xy, size_t stride_bytes,
                             int n_key, int im_width, int im_height, int out_width, int out_height,
                             uint8_t* out, size_t out_stride_bytes)
{
    ERP_ARG(ctx && out && out_width > 0 && out_height > 0, ERP_E_ARG, "erp_draw_epipole: bad argument");
    DeviceGuard g(ctx->device);
    int st = ERP_OK;
    uint8_t* d_out = ctx->scratch<uint8_t>(S_IMG_OUT, (size_t)out_width * out_height * 3, &st);
    ERP_TRY(st);
    ERP_TRY(erp_draw_epipole_dev(ctx, E9, left_xy, right_xy, stride_bytes, n_key, im_width, im_height, out_width, out_height,
                                 d_out, (size_t)out_width * 3));
    ERP_CUDA(cudaMemcpy2DAsync(out, out_stride_bytes, d_out, (size_t)out_width * 3, (size_t)out_width * 3, out_height,
                               cudaMemcpyDeviceToHost, ctx->stream));
    ERP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ERP_OK;
}
