// mccba_host.hpp -- C++17 host-side mirror of cv::multicalib::MultiCameraCalibration for the bundle-adjustment path.
//
// Same class name, constructor parameter list, enum values and public methods as the reference
// (include/opencv2/ccalib/multicalib.hpp:72-165), so the reference's own call sequence
// (samples/multi_cameras_calibration.cpp:71-83, tutorials/multi_camera_tutorial.markdown:39-43)
//     MultiCameraCalibration calib(cameraType, nCameras, fileName, patternWidth, patternHeight, ...);
//     calib.run();                      // loadImages(); initialize(); optimizeExtrinsics();
//     calib.writeParameters("out.xml");
// compiles against it.  OpenCV headers are not available in this build, so cv::TermCriteria / cv::Mat are replaced by
// the small structs below (TermCriteria keeps cv's type bits COUNT=1, EPS=2, decoded exactly as at
// src/multicalib.cpp:475-477).  The feature detector / descriptor / matcher arguments of the reference constructor
// belong to the image front-end, which is out of scope: `fileName` names an observation file (corner
// correspondences + intrinsics + per-edge initial transforms, the same information MyMultiCameraCalibration reads
// from <data>/<serial>/<ts>.yaml and <cfg>/<serial>.xml, src/mymulticalib.cpp:118-131, 182-233) instead of an image
// list.  Numerics run on the GPU through the C ABI in mccba.h; there is no CPU fallback.
#ifndef MCCBA_HOST_HPP_
#define MCCBA_HOST_HPP_

#include <array>
#include <cstdint>
#include <set>
#include <string>
#include <vector>

#include "mccba.h"

namespace mccba {

struct TermCriteria {
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type;
    int maxCount;
    double epsilon;
    TermCriteria(int t = COUNT, int n = 20, double e = 1e-7) : type(t), maxCount(n), epsilon(e) {}
};

using Mat44f = std::array<float, 16>;  // row-major 4x4, CV_32F like vertex::pose / edge::transform

struct SolverOptions {             // extensions that have no counterpart in the reference
    int mode = MCCBA_MODE_REFERENCE_GN;   // MCCBA_MODE_LM for the north-star Levenberg-Marquardt
    double lambda0 = 1e-3, lambdaUp = 10.0, lambdaDown = 1.0 / 3.0;
    int device = 0;
};

class MultiCameraCalibration {
public:
    enum aa1 { PINHOLE, OMNIDIRECTIONAL };      // multicalib.hpp:76-80
    enum PatternSide { FRONT_PATTERN, BACK_PATTERN };   // multicalib.hpp:81-84

    struct edge {                               // multicalib.hpp:86-103
        int cameraVertex, photoVertex, photoIndex;
        Mat44f transform;                       // pattern -> camera
        float reprojecterror = 0.f;
        int patternSide = FRONT_PATTERN;        // multicalib.hpp:91
    };
    struct vertex {                             // multicalib.hpp:105-122
        Mat44f pose;                            // relative pose to the first camera
        int timestamp = -1;
        int timestampCnt = 1;
    };

    MultiCameraCalibration(int cameraType, int nCameras, const std::string& fileName, float patternWidth,
                           float patternHeight, int verbose = 0, int showExtration = 0, int nMiniMatches = 20,
                           int flags = 0, TermCriteria criteria = TermCriteria(TermCriteria::COUNT, 20, 1e-7),
                           SolverOptions solver = SolverOptions());
    virtual ~MultiCameraCalibration();
    MultiCameraCalibration(const MultiCameraCalibration&) = delete;
    MultiCameraCalibration& operator=(const MultiCameraCalibration&) = delete;

    void loadImages();            // reads the observation file (multicalib.hpp:147; src/mymulticalib.cpp:348-405)
    void initialize();            // spanning-tree pose chaining (src/multicalib.cpp:380-420), sparse graph
    double optimizeExtrinsics();  // src/multicalib.cpp:462-514, on the GPU
    double run();                 // src/multicalib.cpp:127-133
    void reset();                 // src/multicalib.cpp:134-152
    virtual void writeParameters(const std::string& filename);   // src/multicalib.cpp:1092-1127, OpenCV XML

    // outlier loop of the shipped workflow (samples/multi_cameras_calibration.cpp:71-83, src/mymulticalib.cpp:406-423)
    std::set<int> removeOutlier(float threshold = 0.5f);   // returns the indices (load order) of the dropped edges

    // inspection (tests, tools)
    const std::vector<edge>& edges() const { return _edgeList; }
    const std::vector<vertex>& vertices() const { return _vertexList; }
    double meanReprojectError() const { return _error; }
    double rms() const { return _rms; }
    const mccba_report& report() const { return _report; }
    std::vector<double> parameters() const { return _params; }   // 6*(nV-1), [rvec|tvec] per vertex (buildParas layout)
    std::vector<double> initialParameters() const;               // buildParas() of the current vertex poses, widened

protected:
    struct Impl;
    Impl* _impl;
    int _camType, _nCamera, _nMiniMatches, _flags, _verbose, _showExtraction;
    float _patternWidth, _patternHeight;
    TermCriteria _criteria;
    SolverOptions _solver;
    std::string _filename;
    double _error = 0, _rms = 0;
    mccba_report _report{};
    std::vector<edge> _edgeList;
    std::vector<vertex> _vertexList;
    std::vector<double> _params;
    std::set<int> _outliers;

    void buildEdges();                                    // src/mymulticalib.cpp:314-403 on the loaded image records
    int getPhotoVertex(int timestamp);                    // src/multicalib.cpp:323-346
    std::vector<float> buildParas() const;                // src/multicalib.cpp:422-440
    void paras2vertex(const std::vector<float>& p);       // src/multicalib.cpp:442-459
};

// The subclass the reference's sample actually runs (include/opencv2/ccalib/mymulticalib.hpp:95-100,
// samples/multi_cameras_calibration.cpp:53-83): corners come from <dataFolder>/<serial>/<timestamp>.yaml ("corners",
// "objects"; files in cv::glob order), intrinsics from <cameraConfigFolder>/<serial>.xml ("Intrinsics", "Distortion"), the
// initial pattern->camera transform of every image from solvePnP (src/mymulticalib.cpp:118-131, 182-233, 268-301).
// Only images of the front pattern (frontPatternSize corners) are kept, as in storeReaded (:236-241).
struct Size {
    int width = 0, height = 0;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};

class MyMultiCameraCalibration : public MultiCameraCalibration {
public:
    MyMultiCameraCalibration(const std::vector<std::string>& cameraSerials, int cameraType, int nCameras,
                             const std::string& dataFolder, const std::string& cameraConfigFolder,
                             const std::string doubleSideConfig, Size frontPatternSize, Size backPatternSize, float patternWidth,
                             float patternHeight, int verbose = 0, int showExtration = 0, int nMiniMatches = 20, int flags = 0,
                             TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 200, 1e-7),
                             SolverOptions solver = SolverOptions());
    void loadImages(const std::set<std::string>& outliers = std::set<std::string>());   // src/mymulticalib.cpp:348-405
    std::set<std::string> removeOutlier();               // files of edges with mean error > 0.5 px (:406-423)
    void writeParameters(const std::string& filename) override;   // base XML + writeParameters2config (:425-460)
    void writeParameters2config();
    double run();                                        // load / initialise / optimise / drop outliers / again (sample :71-80)

protected:
    std::vector<std::string> _serials;
    std::string _dataFolder, _configFolder, _doubleSideConfig;
    Size _front, _back;
    std::set<std::string> _fileOutliers;
    bool _keepBackPattern = false;                       // DoubleSideCalibration::storeReaded keeps both sides (src/doubleSide.cpp:114-118)
    void readCameraIntrinsics();
};

// Double-sided board calibration (include/opencv2/ccalib/doubleSide.hpp:82-178, src/doubleSide.cpp): the cameras are
// FIXED at the poses their config files hold ("CameraMatrix", what writeParameters2config wrote), a board carries a
// front pattern (frontPatternSize corners) and a back pattern (backPatternSize corners, a different count); unknown
// are the front<->back transform and one board pose per timestamp.  loadImages keeps the images of both sides;
// initialize() takes the transform from the first timestamp that one camera sees from the front and another from the
// back (findTransformOfTwoEdge, :119-148) and the board poses from the first edge of each timestamp; the optimisation
// runs on the GPU (mccba_ds_*).  writeParameters writes doublesideTransform.yaml (key "transform", 4 x 4: back-pattern
// coordinates -> front-pattern coordinates), the file MyMultiCameraCalibration::readDoubleSide reads (:98-104).
class DoubleSideCalibration : public MyMultiCameraCalibration {
public:
    DoubleSideCalibration(const std::vector<std::string>& cameraSerials, int cameraType, int nCameras, const std::string& dataFolder,
                          const std::string& cameraConfigFolder, Size frontPatternSize, Size backPatternSize, float patternWidth,
                          float patternHeight, int verbose = 0, int showExtration = 0, int nMiniMatches = 20, int flags = 0,
                          TermCriteria criteria = TermCriteria(TermCriteria::COUNT + TermCriteria::EPS, 200, 1e-8),
                          SolverOptions solver = SolverOptions());
    void initialize();                                   // src/doubleSide.cpp:150-231
    double optimizeExtrinsics();                         // the inherited loop on [transform | board poses], on the GPU; returns the RMS
    double run();                                        // loadImages(); initialize(); optimizeExtrinsics()
    void writeParameters(const std::string& filename) override;   // :582-590; filename = the yaml to write
    const std::array<double, 16>& doubleSideTransform() const { return _dst; }
    const std::vector<std::array<double, 16>>& camerasPose() const { return _camPose; }

private:
    std::array<double, 16> _dst{};                       // row-major 4 x 4
    std::vector<std::array<double, 16>> _camPose;        // world -> camera, per camera
    void loadCameraPose();                               // :276-287
};

}  // namespace mccba

// Plain-C access to the class for bindings and tests (ctypes).  Every function returns 0 on success; the last error
// text of a handle is available through mccbah_last_error.
extern "C" {
typedef struct mccbah_s* mccbah;
int mccbah_create(int cameraType, int nCameras, const char* fileName, float patternWidth, float patternHeight,
                  int verbose, int showExtraction, int nMiniMatches, int flags, int critType, int critMaxCount,
                  double critEps, int mode, int device, mccbah* out);
int mccbah_destroy(mccbah h);
const char* mccbah_last_error(mccbah h);
int mccbah_load_images(mccbah h);
int mccbah_reset(mccbah h);
int mccbah_initialize(mccbah h);
int mccbah_optimize_extrinsics(mccbah h, double* error);
int mccbah_run(mccbah h, double* error);
int mccbah_remove_outlier(mccbah h, float threshold, int* n_removed);
int mccbah_write_parameters(mccbah h, const char* filename);
int mccbah_sizes(mccbah h, int* n_vertex, int* n_edge);
int mccbah_get_indexing(mccbah h, int* edge_cam, int* edge_pv, int* edge_photo_index, int* vertex_timestamp);
int mccbah_get_parameters(mccbah h, double* params /* 6*(nV-1) */);
int mccbah_get_initial_parameters(mccbah h, double* params /* 6*(nV-1), buildParas() of the current poses */);
int mccbah_get_stats(mccbah h, double* mean_error, double* rms, int* iterations, double* device_ms);
/* MyMultiCameraCalibration: serials separated by ','; outliers / returned file lists separated by '\n' */
int mccbah_create_my(const char* serials, int cameraType, int nCameras, const char* dataFolder, const char* cameraConfigFolder,
                     const char* doubleSideConfig, int frontW, int frontH, int backW, int backH, float patternWidth,
                     float patternHeight, int verbose, int critType, int critMaxCount, double critEps, int mode, int device,
                     mccbah* out);
int mccbah_load_images_my(mccbah h, const char* outliers);
int mccbah_remove_outlier_my(mccbah h, char* out, int cap, int* n_removed);
int mccbah_run_my(mccbah h, double* error);
/* DoubleSideCalibration */
int mccbah_create_ds(const char* serials, int cameraType, int nCameras, const char* dataFolder, const char* cameraConfigFolder,
                     int frontW, int frontH, int backW, int backH, float patternWidth, float patternHeight, int verbose,
                     int critType, int critMaxCount, double critEps, int device, mccbah* out);
int mccbah_initialize_ds(mccbah h);
int mccbah_optimize_ds(mccbah h, double* rms);
int mccbah_get_double_side_transform(mccbah h, double* T16);
/* cv::solvePnP (iterative) restated; obj n x 3, img n x 2 (doubles) */
int mccbah_solve_pnp(int n, const double* obj, const double* img, const double* K5, const double* dist8, int ndist, double* rvec,
                     double* tvec);
}

#endif  // MCCBA_HOST_HPP_
