// dropin_main.cpp -- exercises the drop-in classes the way the reference's mains do
// (src/automatic.cpp:117-136: match, then eight_point::find), on inputs read from a binary file so
// that the Python parity test can compare the outputs with the CPU oracle.
//
//   dropin_main <in.bin> <out.bin> [estimated_extrinsic.txt]
// in : int32 nq, nt, dim, W, H | float q[nq*dim] | float t[nt*dim] | float lxy[nq*2] | float rxy[nt*2]
// out: int32 n_matches | DMatch[n] | float R[3] | float T[3] | int32 v1,v2 | float R1[3],R2[3],Tn[3]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "eight_point.hpp"
#include "extrinsic_log.hpp"
#include "feature_matcher.hpp"

static void rd(FILE* f, void* p, size_t n) { if (fread(p, 1, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int hdr[5];
    rd(f, hdr, sizeof hdr);
    const int nq = hdr[0], nt = hdr[1], dim = hdr[2], W = hdr[3], H = hdr[4];
    cv::Mat d1(nq, dim, CV_32FC1), d2(nt, dim, CV_32FC1);
    rd(f, d1.data, (size_t)nq * dim * 4);
    rd(f, d2.data, (size_t)nt * dim * 4);
    std::vector<float> lxy((size_t)nq * 2), rxy((size_t)nt * 2);
    rd(f, lxy.data(), lxy.size() * 4);
    rd(f, rxy.data(), rxy.size() * 4);
    fclose(f);

    try {
        feature_matcher fm;                                           // src/spherical_surf.cpp:96
        std::vector<cv::DMatch> m = fm.match_two_image(d1, d2);       // src/spherical_surf.cpp:153
        // gather the matched keypoints (src/spherical_surf.cpp:155-162)
        std::vector<cv::KeyPoint> kl(m.size()), kr(m.size());
        for (size_t i = 0; i < m.size(); i++) {
            kl[i].pt = cv::Point2f(lxy[2 * m[i].queryIdx], lxy[2 * m[i].queryIdx + 1]);
            kr[i].pt = cv::Point2f(rxy[2 * m[i].trainIdx], rxy[2 * m[i].trainIdx + 1]);
        }
        eight_point estimater;                                        // src/automatic.cpp:124
        cv::Vec3f R, T;
        estimater.find(W, H, kl, kr, R, T, (int)m.size());            // src/automatic.cpp:126

        // the N-point call of src/manual.cpp:152 on the first 64 correspondences
        int n8 = m.size() < 64 ? (int)m.size() : 64;
        std::vector<cv::Point3d> pl(n8), pr(n8);
        for (int i = 0; i < n8; i++) {
            double lon = 2 * M_PI * (kl[i].pt.x / W), lat = M_PI * (kl[i].pt.y / H);
            pl[i] = cv::Point3d(-sin(lat) * cos(lon), sin(lat) * sin(lon), cos(lat));
            lon = 2 * M_PI * (kr[i].pt.x / W); lat = M_PI * (kr[i].pt.y / H);
            pr[i] = cv::Point3d(-sin(lat) * cos(lon), sin(lat) * sin(lon), cos(lat));
        }
        cv::Vec3f R1, R2, Tn;
        bool v1 = false, v2 = false;
        estimater.eight_point_estimation(W, H, pl, pr, R1, R2, Tn, v1, v2, n8);

        FILE* o = fopen(argv[2], "wb");
        if (!o) { perror(argv[2]); return 2; }
        int n = (int)m.size(), vv[2] = {v1, v2};
        fwrite(&n, 4, 1, o);
        fwrite(m.data(), sizeof(cv::DMatch), m.size(), o);
        fwrite(R.val, 4, 3, o); fwrite(T.val, 4, 3, o);
        fwrite(vv, 4, 2, o);
        fwrite(R1.val, 4, 3, o); fwrite(R2.val, 4, 3, o); fwrite(Tn.val, 4, 3, o);
        fclose(o);
        if (argc > 3) {
            // estimated_extrinsic.txt as src/automatic.cpp:128-136 writes it, and back through the parser
            std::ofstream log(argv[3]);
            const cv::Vec3d rv(R[0], R[1], R[2]), tv(T[0], T[1], T[2]);
            erp_host::write_initial_pose(log, rv, tv);
            log.close();
            std::ifstream in(argv[3]);
            cv::Vec3d r2, t2;
            if (!erp_host::read_initial_pose(in, r2, t2)) { fprintf(stderr, "dropin_main: log does not parse\n"); return 1; }
            for (int i = 0; i < 3; i++)
                if (std::fabs(r2[i] - rv[i]) > 1e-5 * (1 + std::fabs(rv[i])) || std::fabs(t2[i] - tv[i]) > 1e-5) { fprintf(stderr, "dropin_main: log round trip\n"); return 1; }
        }
        DEBUG_PRINT_OUT("matches " << n << "  R " << R[0] << " " << R[1] << " " << R[2] << "  T " << T[0] << " " << T[1] << " " << T[2]);
    } catch (const std::exception& e) {
        fprintf(stderr, "dropin_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
