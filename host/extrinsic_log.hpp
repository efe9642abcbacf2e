// extrinsic_log.hpp -- the reference's on-disk result format (estimated_extrinsic.txt).
// src/automatic.cpp:135-136 writes, after eight_point::find,
//     log << "initial_R_vector: " << DEGREE(initial_rot_vec) << endl;      // XYZ Euler, degrees
//     log << "initial_T_vector: " << initial_t_vec << endl;                // unit translation
// where cv::Vec3d streams as "[a, b, c]" with the stream's own number formatting (6 significant digits
// unless the caller changed the precision).  write_initial_pose reproduces those two lines byte for byte so
// that logs of the reference and of this build can be diffed; read_initial_pose parses them back.
#pragma once
#include <iosfwd>

#include "opencv2/core.hpp"

namespace erp_host {

void write_initial_pose(std::ostream& log, const cv::Vec3d& rot_vec_rad, const cv::Vec3d& t_vec);
// returns false when the two lines are not found; rot_vec comes back in radians
bool read_initial_pose(std::istream& log, cv::Vec3d& rot_vec_rad, cv::Vec3d& t_vec);

} // namespace erp_host
