// compat_io.cpp -- the image I/O entry points of the OpenCV type shim (host/compat/opencv2/core.hpp).  The shim has no
// codecs: the functions exist so that callers written against OpenCV (src/automatic.cpp:91-92,150-153) link, and a call
// says what is missing.  Not compiled when the build uses a real OpenCV.
#include <opencv2/core.hpp>

#ifdef ERP_OPENCV_COMPAT
namespace cv {

Mat imread(const std::string& path, int)
{
    CV_Error(Error::StsNotImplemented, "compat imread(" + path + "): build against a real OpenCV for image I/O");
}

bool imwrite(const std::string& path, const Mat&)
{
    CV_Error(Error::StsNotImplemented, "compat imwrite(" + path + "): build against a real OpenCV for image I/O");
}

} // namespace cv
#endif
