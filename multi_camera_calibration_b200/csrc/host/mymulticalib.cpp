// mymulticalib.cpp -- MyMultiCameraCalibration: the reference's directory ingest in front of the GPU path
// (src/mymulticalib.cpp:72-98 constructor, :118-131 intrinsics, :182-233 corner files + solvePnP, :268-301 loadOneSerial,
// :348-405 loadImages, :406-423 removeOutlier, :425-460 writeParameters2config; call sequence of
// samples/multi_cameras_calibration.cpp:71-83).
#include <algorithm>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "host_impl.hpp"

namespace fs = std::filesystem;

namespace mccba {

MyMultiCameraCalibration::MyMultiCameraCalibration(const std::vector<std::string>& cameraSerials, int cameraType, int nCameras,
                                                   const std::string& dataFolder, const std::string& cameraConfigFolder,
                                                   const std::string doubleSideConfig, Size frontPatternSize, Size backPatternSize,
                                                   float patternWidth, float patternHeight, int verbose, int showExtration,
                                                   int nMiniMatches, int flags, TermCriteria criteria, SolverOptions solver)
    : MultiCameraCalibration(cameraType, nCameras, dataFolder, patternWidth, patternHeight, verbose, showExtration, nMiniMatches,
                             flags, criteria, solver),
      _serials(cameraSerials), _dataFolder(dataFolder), _configFolder(cameraConfigFolder), _doubleSideConfig(doubleSideConfig),
      _front(frontPatternSize), _back(backPatternSize)
{
    if ((int)_serials.size() != nCameras) throw std::invalid_argument("one serial per camera expected");
    if (cameraType != PINHOLE) throw std::invalid_argument("the directory ingest initialises poses with solvePnP: PINHOLE cameras only");
    readCameraIntrinsics();
}

void MyMultiCameraCalibration::readCameraIntrinsics()   // src/mymulticalib.cpp:118-131
{
    Impl& I = *_impl;
    I.cams.assign(_serials.size(), CameraIntrinsics());
    for (size_t c = 0; c < _serials.size(); ++c) {
        const std::string filename = _configFolder + "/" + _serials[c] + ".xml";
        CvFileReader f(filename);
        CvMatrix K, D;
        if (!f.matrix("Intrinsics", K) || K.data.size() != 9) throw std::runtime_error(filename + ": no 3x3 \"Intrinsics\" matrix");
        if (!f.matrix("Distortion", D)) D = CvMatrix();
        CameraIntrinsics& ci = I.cams[c];
        ci.model = PINHOLE;
        // convertTo(CV_32F): the reference keeps intrinsics in float32
        ci.K5[0] = (float)K.data[0]; ci.K5[1] = (float)K.data[4]; ci.K5[2] = (float)K.data[2]; ci.K5[3] = (float)K.data[5];
        ci.K5[4] = (float)K.data[1];
        const int nd = (int)D.data.size();
        if (nd != 0 && nd != 4 && nd != 5 && nd != 8)
            throw std::runtime_error(filename + ": " + std::to_string(nd) + " distortion coefficients (0, 4, 5 or 8 supported)");
        ci.ndist = nd;
        for (int i = 0; i < nd; ++i) ci.dist8[i] = (float)D.data[i];
    }
}

void MyMultiCameraCalibration::loadImages(const std::set<std::string>& outliers)
{
    Impl& I = *_impl;
    if (!outliers.empty()) _fileOutliers = outliers;          // :350-353
    I.images.clear();
    I.obj.clear();
    I.img.clear();
    size_t total = 0;
    for (int c = 0; c < _nCamera; ++c) {                      // loadOneSerial, :268-301
        const std::string folder = _dataFolder + "/" + _serials[c];
        std::vector<std::string> files;
        if (fs::is_directory(folder))
            for (const auto& de : fs::directory_iterator(folder))
                if (de.is_regular_file() && de.path().extension() == ".yaml") files.push_back(folder + "/" + de.path().filename().string());
        std::sort(files.begin(), files.end());                // cv::glob returns the paths sorted as strings
        for (const std::string& file : files) {
            if (_fileOutliers.count(file)) {
                if (_verbose) std::fprintf(stderr, "outlier:%s skipped\n", file.c_str());
                continue;
            }
            const int timestamp = std::stoi(fs::path(file).stem().string());     // readTimestamps, :213-219
            CvFileReader f(file);
            CvMatrix corners, objects;
            if (!f.matrix("corners", corners) || !f.matrix("objects", objects)) throw std::runtime_error(file + ": \"corners\" / \"objects\" missing");
            const int n = (int)(corners.data.size() / 2);
            if ((int)(objects.data.size() / 3) != n || n < 4) throw std::runtime_error(file + ": corners and objects do not match");
            // objects are converted to CV_32F before solvePnP (:190), corners keep the file's depth
            std::vector<double> o(3 * (size_t)n), m(2 * (size_t)n);
            for (int i = 0; i < 3 * n; ++i) o[i] = (double)(float)objects.data[i];
            for (int i = 0; i < 2 * n; ++i) m[i] = corners.data[i];
            double rvec[3], tvec[3];
            const CameraIntrinsics& ci = I.cams[c];
            if (!solve_pnp(n, o.data(), m.data(), ci.K5, ci.dist8, ci.ndist, rvec, tvec)) throw std::runtime_error(file + ": solvePnP failed");
            const float tf[3] = {(float)tvec[0], (float)tvec[1], (float)tvec[2]};      // tvec.convertTo(CV_32F), :209
            const double r = std::sqrt((double)tf[0] * tf[0] + (double)tf[1] * tf[1] + (double)tf[2] * tf[2]);
            if (!(r < 3000 && r > 300))                                                   // isValidPose, src/multicalib.cpp:107-113
                throw std::runtime_error(file + ": pattern pose outside (300, 3000) units (the reference asserts here, src/mymulticalib.cpp:294-298)");
            int side = 0;
            if (n != _front.width * _front.height) {                                     // storeReaded, :236-241: front pattern only ...
                if (!(_keepBackPattern && n == _back.width * _back.height)) continue;    // ... unless the subclass keeps both sides
                side = 1;
            }
            ImageRecord im;
            im.side = side;
            im.camera = c; im.timestamp = timestamp; im.n_points = n; im.first = total; im.path = file;
            const double omf[3] = {(double)(float)rvec[0], (double)(float)rvec[1], (double)(float)rvec[2]};   // :387-390 CV_32F
            double R[9];
            exp_so3(omf, R);
            im.transform = eye4();
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) im.transform[i * 4 + j] = (float)R[i * 3 + j];
                im.transform[i * 4 + 3] = tf[i];
            }
            I.images.push_back(im);
            for (int i = 0; i < 3 * n; ++i) I.obj.push_back((float)o[i]);
            for (int i = 0; i < 2 * n; ++i) I.img.push_back((float)m[i]);              // storeReadedImp converts to CV_32F (:228-229)
            total += (size_t)n;
        }
    }
    buildEdges();
}

std::set<std::string> MyMultiCameraCalibration::removeOutlier()   // :406-423
{
    Impl& I = *_impl;
    std::set<std::string> names;
    std::vector<edge> keep;
    std::vector<int> keepImg;
    for (size_t e = 0; e < _edgeList.size(); ++e) {
        if (_edgeList[e].reprojecterror > 0.5f) names.insert(I.images[I.edgeImage[e]].path);
        else { keep.push_back(_edgeList[e]); keepImg.push_back(I.edgeImage[e]); }
    }
    _edgeList.swap(keep);
    I.edgeImage.swap(keepImg);
    return names;
}

double MyMultiCameraCalibration::run()   // samples/multi_cameras_calibration.cpp:71-80
{
    loadImages();
    initialize();
    optimizeExtrinsics();
    const std::set<std::string> out = removeOutlier();
    reset();
    loadImages(out);
    initialize();
    return optimizeExtrinsics();
}

void MyMultiCameraCalibration::writeParameters2config()   // :425-454
{
    for (size_t c = 0; c < _serials.size(); ++c) {
        const std::string filename = _configFolder + "/" + _serials[c] + ".xml";
        double depth_scale = 0, height = 0;
        CvMatrix K, D;
        {
            CvFileReader f(filename);
            f.scalar("depth_scale", depth_scale);     // a missing key reads as 0, like FileNode >> float
            f.scalar("height", height);
            f.matrix("Intrinsics", K);
            f.matrix("Distortion", D);
        }
        std::ofstream os(filename);
        if (!os) throw std::runtime_error("cannot write " + filename);
        os << "<?xml version=\"1.0\"?>\n<opencv_storage>\n";
        os << "<depth_scale>" << cv_format_real((double)(float)depth_scale, 9) << "</depth_scale>\n";
        os << "<height>" << cv_format_real((double)(float)height, 9) << "</height>\n";
        double pose[16];
        for (int i = 0; i < 16; ++i) pose[i] = _vertexList[c].pose[i];
        cv_write_mat_xml(os, "CameraMatrix", 4, 4, 'f', pose);            // mats[0] = _vertexList[camIdx].pose
        if (!K.empty()) cv_write_mat_xml(os, "Intrinsics", K.rows, K.cols, K.depth == 'f' ? 'f' : 'd', K.data.data());
        if (!D.empty()) cv_write_mat_xml(os, "Distortion", D.rows, D.cols, D.depth == 'f' ? 'f' : 'd', D.data.data());
        os << "</opencv_storage>\n";
    }
}

void MyMultiCameraCalibration::writeParameters(const std::string& filename)   // :457-460
{
    MultiCameraCalibration::writeParameters(filename);
    writeParameters2config();
}

}  // namespace mccba

// ---- plain-C access ------------------------------------------------------------------------------------------
namespace {
std::vector<std::string> split(const char* s, char sep)
{
    std::vector<std::string> out;
    if (!s) return out;
    std::string cur;
    for (const char* p = s; *p; ++p) {
        if (*p == sep) { if (!cur.empty()) out.push_back(cur); cur.clear(); }
        else cur.push_back(*p);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}
template <typename F>
int guarded_my(mccbah h, F&& f)
{
    if (!h || !h->obj) return 1;
    auto* my = dynamic_cast<mccba::MyMultiCameraCalibration*>(h->obj);
    if (!my) { h->err = "handle is not a MyMultiCameraCalibration"; return 1; }
    try {
        f(*my);
        return 0;
    } catch (const std::exception& e) {
        h->err = e.what();
        return 2;
    }
}
}  // namespace

extern "C" {

int mccbah_create_my(const char* serials, int cameraType, int nCameras, const char* dataFolder, const char* cameraConfigFolder,
                     const char* doubleSideConfig, int frontW, int frontH, int backW, int backH, float patternWidth,
                     float patternHeight, int verbose, int critType, int critMaxCount, double critEps, int mode, int device,
                     mccbah* out)
{
    if (!out || !serials || !dataFolder || !cameraConfigFolder) return 1;
    mccbah h = new mccbah_s();
    *out = h;
    try {
        mccba::SolverOptions so;
        so.mode = mode;
        so.device = device;
        h->obj = new mccba::MyMultiCameraCalibration(split(serials, ','), cameraType, nCameras, dataFolder, cameraConfigFolder,
                                                     doubleSideConfig ? doubleSideConfig : "", mccba::Size(frontW, frontH),
                                                     mccba::Size(backW, backH), patternWidth, patternHeight, verbose, 0, 20, 0,
                                                     mccba::TermCriteria(critType, critMaxCount, critEps), so);
    } catch (const std::exception& e) {
        h->err = e.what();
        return 2;
    }
    return 0;
}
int mccbah_load_images_my(mccbah h, const char* outliers)
{
    return guarded_my(h, [&](mccba::MyMultiCameraCalibration& my) {
        std::set<std::string> o;
        for (const auto& s : split(outliers, '\n')) o.insert(s);
        my.loadImages(o);
    });
}
int mccbah_remove_outlier_my(mccbah h, char* out, int cap, int* n_removed)
{
    return guarded_my(h, [&](mccba::MyMultiCameraCalibration& my) {
        const std::set<std::string> names = my.removeOutlier();
        std::string all;
        for (const auto& s : names) { all += s; all += '\n'; }
        if (n_removed) *n_removed = (int)names.size();
        if (out && cap > 0) {
            const size_t k = std::min<size_t>(all.size(), (size_t)cap - 1);
            std::copy(all.begin(), all.begin() + (long)k, out);
            out[k] = 0;
        }
    });
}
int mccbah_run_my(mccbah h, double* error)
{
    return guarded_my(h, [&](mccba::MyMultiCameraCalibration& my) { const double e = my.run(); if (error) *error = e; });
}
int mccbah_solve_pnp(int n, const double* obj, const double* img, const double* K5, const double* dist8, int ndist, double* rvec,
                     double* tvec)
{
    if (n < 4 || !obj || !img || !K5 || !rvec || !tvec) return 1;
    return mccba::solve_pnp(n, obj, img, K5, dist8, ndist, rvec, tvec) ? 0 : 2;
}
}
