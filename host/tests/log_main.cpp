// log_main.cpp -- writes estimated_extrinsic.txt for the pose given on the command line (radians, unit t), reads it
// back and prints the parsed values: the CPU-only check of extrinsic_log.{hpp,cpp} (format of src/automatic.cpp:135-136).
//   log_main out.txt rx ry rz tx ty tz
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "extrinsic_log.hpp"

int main(int argc, char** argv)
{
    if (argc != 8) { fprintf(stderr, "usage: %s out.txt rx ry rz tx ty tz\n", argv[0]); return 2; }
    const cv::Vec3d r(atof(argv[2]), atof(argv[3]), atof(argv[4])), t(atof(argv[5]), atof(argv[6]), atof(argv[7]));
    {
        std::ofstream log(argv[1]);
        log << "match result" << std::endl;                 // other lines of the reference's log are skipped by the reader
        erp_host::write_initial_pose(log, r, t);
    }
    std::ifstream in(argv[1]);
    cv::Vec3d r2, t2;
    if (!erp_host::read_initial_pose(in, r2, t2)) return 1;
    printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", r2[0], r2[1], r2[2], t2[0], t2[1], t2[2]);
    return 0;
}
