// host_impl.hpp -- private state and small helpers shared by the host-side classes (multicalib.cpp, mymulticalib.cpp).
#pragma once
#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

#include "mccba_host.hpp"

namespace mccba {

struct CameraIntrinsics {
    int model = 0, ndist = 0;
    double K5[5] = {0, 0, 0, 0, 0}, dist8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xi = 0;
};
struct ImageRecord {   // one (camera, timestamp) image: what loadOneSerial keeps per file (src/mymulticalib.cpp:268-301)
    int camera = 0, timestamp = 0, n_points = 0;
    Mat44f transform{};
    size_t first = 0;  // offset of its corners in the concatenated point arrays
    std::string path;  // source file (directory ingest only)
    int side = 0;      // 0 front pattern, 1 back pattern (double-sided boards)
};

struct MultiCameraCalibration::Impl {
    mccba_handle h = nullptr;
    std::vector<CameraIntrinsics> cams;
    std::vector<ImageRecord> images;                       // load order: cameras outer loop, files in glob order
    std::vector<std::vector<int>> imagesOfCamera;          // per camera: indices into images (photoIndex order)
    std::vector<float> obj, img;                           // concatenated CV_32F points of all images
    std::vector<int> edgeImage;                            // edge -> image record
    std::unordered_map<int, int> tsToVertex;               // timestamp -> photo vertex (first-seen order)
    bool loaded = false, initialised = false;
};

inline Mat44f eye4()
{
    Mat44f m{};
    m[0] = m[5] = m[10] = m[15] = 1.f;
    return m;
}
// cv::Rodrigues vector -> matrix (paras2vertex src/multicalib.cpp:449)
inline void exp_so3(const double* om, double* R)
{
    const double x = om[0], y = om[1], z = om[2], th2 = x * x + y * y + z * z, th = std::sqrt(th2);
    double a, b;
    if (th < 1e-8) { a = 1 - th2 / 6; b = 0.5 - th2 / 24; }
    else { a = std::sin(th) / th; b = (1 - std::cos(th)) / th2; }
    R[0] = 1 - b * (y * y + z * z); R[1] = b * x * y - a * z; R[2] = b * x * z + a * y;
    R[3] = b * x * y + a * z; R[4] = 1 - b * (x * x + z * z); R[5] = b * y * z - a * x;
    R[6] = b * x * z - a * y; R[7] = b * y * z + a * x; R[8] = 1 - b * (x * x + y * y);
}
// cv::Rodrigues matrix -> vector for a row-major 3x3 (theta < pi)
inline void log_so3_3x3(const double* R, double* om)
{
    const double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = std::max(-1.0, std::min(1.0, c));
    const double th = std::atan2(s, c);
    if (s < 1e-9) { om[0] = 0.5 * rx; om[1] = 0.5 * ry; om[2] = 0.5 * rz; return; }
    const double k = th / (2 * s);
    om[0] = rx * k; om[1] = ry * k; om[2] = rz * k;
}

// ---- OpenCV FileStorage subset (cvstorage.cpp): matrices and real scalars, YAML 1.0 and XML flavours -------------
struct CvMatrix {
    int rows = 0, cols = 0, channels = 1;
    char depth = 'd';                 // u c w s i f d
    std::vector<double> data;         // rows * cols * channels
    bool empty() const { return data.empty(); }
};
class CvFileReader {
public:
    explicit CvFileReader(const std::string& path);   // throws std::runtime_error if the file cannot be read
    bool matrix(const std::string& name, CvMatrix& out) const;
    bool scalar(const std::string& name, double& out) const;
private:
    std::string text_;
    bool xml_ = false;
};
std::string cv_format_real(double v, int digits);
void cv_write_mat_xml(std::ostream& os, const std::string& name, int rows, int cols, char dt, const double* data);

// ---- cv::solvePnP (SOLVEPNP_ITERATIVE) restated (pnp.cpp): pinhole + radtan / rational distortion --------------------
// obj (n x 3), img (n x 2); K5 = fx fy cx cy skew(ignored); returns false if the pose cannot be initialised.
bool solve_pnp(int n, const double* obj, const double* img, const double* K5, const double* dist8, int ndist, double* rvec,
               double* tvec);

}  // namespace mccba

// the plain-C handle of include/mccba_host.hpp (one definition for multicalib.cpp and mymulticalib.cpp)
struct mccbah_s {
    mccba::MultiCameraCalibration* obj = nullptr;
    std::string err;
};
