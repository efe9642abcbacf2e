// strip_main.cpp -- drives erp_rotation::rotate_pixel the way the reference's callers do: once per pixel of a strip,
// from an OpenMP loop (the pattern of spherical_surf::crop_rotated_image / rotate_keypoint, src/spherical_surf.cpp:16-63).
// With the drop-in this must be plain host arithmetic: no CUDA context, no device round trip per pixel.
//
//   strip_main <width> <height> <pitch_deg> <out.bin> [image.bin out_strip.bin]
// out.bin: int32 (row, col) per pixel of the strip rows [3H/8, 3H/8 + H/4); prints the seconds the loop took.
// With image.bin (H x W x 3 bytes) it also fills the strip from the mapped source pixels and, on a GPU box, compares
// the result with erp_crop_rotated_image (the batched device entry point) -- the two must pick the same pixels.
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "erp_b200.h"
#include "erp_rotation.hpp"

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s width height pitch_deg out.bin [image.bin strip.bin]\n", argv[0]); return 2; }
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    const double pitch = atof(argv[3]);
    erp_rotation rot;
    // the callers build the pitch matrix from a Vec3f (src/spherical_surf.cpp:26,52): the angle is rounded to float first
    cv::Mat R = rot.eular2rot(cv::Vec3f(0, RAD(pitch), 0));
    const int rows = H / 4, first = H * 3 / 8;
    std::vector<int> map((size_t)rows * W * 2);
    const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < W; j++) {
            const cv::Vec2i px = rot.rotate_pixel(cv::Vec2i(first + i, j), R, W, H);
            map[((size_t)i * W + j) * 2] = px[0];
            map[((size_t)i * W + j) * 2 + 1] = px[1];
        }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    FILE* o = fopen(argv[4], "wb");
    if (!o) { perror(argv[4]); return 2; }
    fwrite(map.data(), sizeof(int), map.size(), o);
    fclose(o);
    printf("%.6f seconds for %d x %d pixels on %d threads\n", sec, rows, W, omp_get_max_threads());
    if (argc < 7) return 0;

    std::vector<unsigned char> im((size_t)H * W * 3), strip((size_t)rows * W * 3, 0), dev((size_t)rows * W * 3, 0);
    FILE* f = fopen(argv[5], "rb");
    if (!f || fread(im.data(), 1, im.size(), f) != im.size()) { fprintf(stderr, "cannot read %s\n", argv[5]); return 2; }
    fclose(f);
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < W; j++) {
            const int r = map[((size_t)i * W + j) * 2], c = map[((size_t)i * W + j) * 2 + 1];
            if (r >= 0 && c >= 0 && r < H && c < W)
                for (int k = 0; k < 3; k++) strip[((size_t)i * W + j) * 3 + k] = im[((size_t)r * W + c) * 3 + k];
        }
    o = fopen(argv[6], "wb");
    fwrite(strip.data(), 1, strip.size(), o);
    fclose(o);
    if (erp_device_count() > 0) {
        erp_ctx* ctx = nullptr;
        if (erp_ctx_create(0, &ctx) != ERP_OK) { fprintf(stderr, "%s\n", erp_last_error()); return 1; }
        if (erp_crop_rotated_image(ctx, im.data(), W, H, (size_t)W * 3, (float)pitch, dev.data(), (size_t)W * 3) != ERP_OK) {
            fprintf(stderr, "%s\n", erp_last_error());
            return 1;
        }
        erp_ctx_destroy(ctx);
        size_t differ = 0;
        for (size_t p = 0; p < (size_t)rows * W; p++)
            differ += strip[3 * p] != dev[3 * p] || strip[3 * p + 1] != dev[3 * p + 1] || strip[3 * p + 2] != dev[3 * p + 2];
        printf("device strip differs in %zu of %zu pixels\n", differ, (size_t)rows * W);
    }
    return 0;
}
