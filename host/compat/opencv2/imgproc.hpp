// OpenCV compat shim: see opencv2/core.hpp
#pragma once
#include "opencv2/core.hpp"
