// extrinsic_log.cpp -- see extrinsic_log.hpp (format of src/automatic.cpp:135-136).
#include "extrinsic_log.hpp"

#include <cmath>
#include <cstdio>
#include <istream>
#include <ostream>
#include <string>

namespace erp_host {

static void put_vec(std::ostream& o, double a, double b, double c) { o << "[" << a << ", " << b << ", " << c << "]"; }

void write_initial_pose(std::ostream& log, const cv::Vec3d& r, const cv::Vec3d& t)
{
    // DEGREE(x) = 180.0*(x)/M_PI, evaluated in that order (src/erp_rotation.hpp:7)
    log << "initial_R_vector: ";
    put_vec(log, 180.0 * r[0] / M_PI, 180.0 * r[1] / M_PI, 180.0 * r[2] / M_PI);
    log << std::endl;
    log << "initial_T_vector: ";
    put_vec(log, t[0], t[1], t[2]);
    log << std::endl;
}

static bool parse_line(const std::string& line, const char* key, double out[3])
{
    const size_t k = line.find(key);
    if (k == std::string::npos) return false;
    return std::sscanf(line.c_str() + k + std::string(key).size(), " [%lf, %lf, %lf]", &out[0], &out[1], &out[2]) == 3;
}

bool read_initial_pose(std::istream& log, cv::Vec3d& r, cv::Vec3d& t)
{
    bool have_r = false, have_t = false;
    std::string line;
    double v[3];
    while (std::getline(log, line)) {
        if (parse_line(line, "initial_R_vector:", v)) { r = cv::Vec3d(v[0] * M_PI / 180.0, v[1] * M_PI / 180.0, v[2] * M_PI / 180.0); have_r = true; }
        else if (parse_line(line, "initial_T_vector:", v)) { t = cv::Vec3d(v[0], v[1], v[2]); have_t = true; }
    }
    return have_r && have_t;
}

} // namespace erp_host
