// multicalib.cpp -- host side of the drop-in: the reference's MultiCameraCalibration surface (include/mccba_host.hpp)
// on top of the C ABI (include/mccba.h).  Host work only: file ingest, the bit-exact vertex/edge indexing of the
// reference, spanning-tree initialisation, parameter (de)serialisation and the OpenCV-format XML output.  All
// optimisation arithmetic happens in libmccba.so on the GPU.
#include "mccba_host.hpp"
#include "host_impl.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <queue>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace mccba {

namespace {

// products / inverses of rigid 4x4 transforms, evaluated in double and stored CV_32F like the reference's poses
Mat44f mul44(const Mat44f& a, const Mat44f& b)
{
    Mat44f c{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)a[i * 4 + k] * (double)b[k * 4 + j];
            c[i * 4 + j] = (float)s;
        }
    return c;
}
Mat44f inv_rigid(const Mat44f& a)
{
    Mat44f c = eye4();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i * 4 + j] = a[j * 4 + i];
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)a[k * 4 + i] * (double)a[k * 4 + 3];
        c[i * 4 + 3] = (float)(-s);
    }
    return c;
}
// cv::Rodrigues matrix -> vector (theta < pi; buildParas src/multicalib.cpp:432)
void log_so3(const Mat44f& m, double* om)
{
    const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
    const double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = std::max(-1.0, std::min(1.0, c));
    const double th = std::atan2(s, c);
    if (s < 1e-9) { om[0] = 0.5 * rx; om[1] = 0.5 * ry; om[2] = 0.5 * rz; return; }
    const double k = th / (2 * s);
    om[0] = rx * k; om[1] = ry * k; om[2] = rz * k;
}
}  // namespace

MultiCameraCalibration::MultiCameraCalibration(int cameraType, int nCameras, const std::string& fileName,
                                               float patternWidth, float patternHeight, int verbose, int showExtration,
                                               int nMiniMatches, int flags, TermCriteria criteria, SolverOptions solver)
    : _impl(new Impl()), _camType(cameraType), _nCamera(nCameras), _nMiniMatches(nMiniMatches), _flags(flags),
      _verbose(verbose), _showExtraction(showExtration), _patternWidth(patternWidth), _patternHeight(patternHeight),
      _criteria(criteria), _solver(solver), _filename(fileName)
{
    if (nCameras < 1) throw std::invalid_argument("nCameras must be >= 1");
    for (int i = 0; i < _nCamera; ++i) {   // src/multicalib.cpp:100-103: camera vertices first, identity pose
        vertex v;
        v.pose = eye4();
        _vertexList.push_back(v);
    }
}

MultiCameraCalibration::~MultiCameraCalibration()
{
    if (_impl->h) mccba_destroy(_impl->h);
    delete _impl;
}

void MultiCameraCalibration::reset()   // src/multicalib.cpp:134-152
{
    _edgeList.clear();
    _vertexList.clear();
    for (int i = 0; i < _nCamera; ++i) {
        vertex v;
        v.pose = eye4();
        _vertexList.push_back(v);
    }
    _impl->edgeImage.clear();
    _impl->tsToVertex.clear();
    _impl->images.clear();
    _impl->imagesOfCamera.clear();
    _impl->obj.clear();
    _impl->img.clear();
    _impl->loaded = _impl->initialised = false;
}

int MultiCameraCalibration::getPhotoVertex(int timestamp)
{
    // src/multicalib.cpp:323-346 scans the vertex list linearly (O(nV) per image); the scan also visits the camera
    // vertices, whose timestamp is -1, so a timestamp of -1 would alias camera 0.  A hash map gives the same
    // first-seen numbering in O(1).
    auto& m = _impl->tsToVertex;
    if (m.empty()) m[-1] = 0;
    auto it = m.find(timestamp);
    if (it != m.end()) { _vertexList[it->second].timestampCnt++; return it->second; }
    vertex v;
    v.pose = eye4();
    v.timestamp = timestamp;
    _vertexList.push_back(v);
    m[timestamp] = (int)_vertexList.size() - 1;
    return (int)_vertexList.size() - 1;
}

void MultiCameraCalibration::loadImages()
{
    std::ifstream f(_filename, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open observation file " + _filename);
    char magic[8];
    int32_t hdr[4];
    f.read(magic, 8);
    f.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
    if (!f || std::memcmp(magic, "MCCBOBS1", 8) != 0) throw std::runtime_error("not an MCCBOBS1 observation file: " + _filename);
    const int nC = hdr[0], nImg = hdr[1];
    if (nC != _nCamera) throw std::runtime_error("observation file holds " + std::to_string(nC) + " cameras, expected " + std::to_string(_nCamera));
    Impl& I = *_impl;
    I.cams.resize(nC);
    for (int c = 0; c < nC; ++c) {
        int32_t mi[2];
        f.read(reinterpret_cast<char*>(mi), sizeof(mi));
        I.cams[c].model = mi[0];
        I.cams[c].ndist = mi[1];
        f.read(reinterpret_cast<char*>(I.cams[c].K5), sizeof(double) * 5);
        f.read(reinterpret_cast<char*>(I.cams[c].dist8), sizeof(double) * 8);
        f.read(reinterpret_cast<char*>(&I.cams[c].xi), sizeof(double));
        if (I.cams[c].model != _camType)   // the reference has ONE camera type for the whole rig (multicalib.hpp:193)
            if (_verbose) std::fprintf(stderr, "camera %d: model %d differs from cameraType %d (per-camera model is an extension)\n", c, I.cams[c].model, _camType);
    }
    I.images.resize(nImg);
    size_t total = 0;
    for (int i = 0; i < nImg; ++i) {
        int32_t r[4];
        f.read(reinterpret_cast<char*>(r), sizeof(r));
        ImageRecord& im = I.images[i];
        im.camera = r[0]; im.timestamp = r[1]; im.n_points = r[2];
        f.read(reinterpret_cast<char*>(im.transform.data()), sizeof(float) * 16);
        if (im.camera < 0 || im.camera >= nC || im.n_points < 0) throw std::runtime_error("corrupt image record");
        im.first = total;
        total += (size_t)im.n_points;
    }
    I.obj.resize(3 * total);
    I.img.resize(2 * total);
    for (int i = 0; i < nImg; ++i) {
        const ImageRecord& im = I.images[i];
        f.read(reinterpret_cast<char*>(I.obj.data() + 3 * im.first), sizeof(float) * 3 * (size_t)im.n_points);
        f.read(reinterpret_cast<char*>(I.img.data() + 2 * im.first), sizeof(float) * 2 * (size_t)im.n_points);
    }
    if (!f) throw std::runtime_error("observation file truncated: " + _filename);
    buildEdges();
}

// vertices and edges from the loaded image records: the indexing contract of the reference (bit-exact)
void MultiCameraCalibration::buildEdges()
{
    Impl& I = *_impl;
    const int nC = _nCamera, nImg = (int)I.images.size();
    // photoIndex = index into the camera's own image list (multicalib.hpp:90)
    I.imagesOfCamera.assign(nC, {});
    for (int i = 0; i < nImg; ++i) I.imagesOfCamera[I.images[i].camera].push_back(i);
    // identifyMultiCameraTimestamps (src/mymulticalib.cpp:314-347): keep timestamps some OTHER camera also has
    std::unordered_map<int, int> firstCam, multi;
    for (int i = 0; i < nImg; ++i) {
        if (_outliers.count(i)) continue;                        // outlier files are never loaded (:277-281)
        const ImageRecord& im = I.images[i];
        auto it = firstCam.find(im.timestamp);
        if (it == firstCam.end()) firstCam[im.timestamp] = im.camera;
        else if (it->second != im.camera) multi[im.timestamp] = 1;
    }
    // edges: cameras outer loop, the camera's images in load order (src/mymulticalib.cpp:360-403)
    for (int c = 0; c < nC; ++c) {
        int pi = -1;
        for (size_t q = 0; q < I.imagesOfCamera[c].size(); ++q) {
            const int ii = I.imagesOfCamera[c][q];
            const ImageRecord& im = I.images[ii];
            if (_outliers.count(ii)) continue;                   // m_outliers, src/mymulticalib.cpp:277-281
            ++pi;                                                // index among the images actually loaded for this camera
            if (!multi.count(im.timestamp)) continue;            // :374-376
            if (im.n_points <= 0) continue;
            edge e;
            e.cameraVertex = c;
            e.photoVertex = getPhotoVertex(im.timestamp);
            e.photoIndex = pi;
            e.transform = im.transform;
            e.patternSide = im.side;
            _edgeList.push_back(e);
            I.edgeImage.push_back(ii);
        }
    }
    I.loaded = true;
    if (_verbose) std::fprintf(stderr, "loaded %zu edges, %zu vertices\n", _edgeList.size(), _vertexList.size());
}

void MultiCameraCalibration::initialize()
{
    // src/multicalib.cpp:380-420.  The reference builds a dense nV x nV int matrix (40 GB at 100k frames) and scans
    // rows for neighbours; adjacency lists sorted by neighbour index give the same BFS order and the same
    // "last edge wins" rule for G(cameraVertex, photoVertex) = edgeIdx + 1.
    if (!_impl->loaded) throw std::logic_error("initialize() before loadImages()");
    const int nV = (int)_vertexList.size();
    std::vector<std::vector<std::pair<int, int>>> adj(nV);   // (neighbour, edge)
    for (int e = 0; e < (int)_edgeList.size(); ++e) {
        adj[_edgeList[e].cameraVertex].push_back({_edgeList[e].photoVertex, e});
        adj[_edgeList[e].photoVertex].push_back({_edgeList[e].cameraVertex, e});
    }
    for (auto& a : adj) {
        std::stable_sort(a.begin(), a.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; });
        std::vector<std::pair<int, int>> u;   // keep the LAST edge of a duplicated (camera, photo) pair
        for (auto& p : a) {
            if (!u.empty() && u.back().first == p.first) u.back() = p;
            else u.push_back(p);
        }
        a.swap(u);
    }
    std::vector<int> pre(nV, -2), order, preEdge(nV, -1);
    std::queue<int> q;
    pre[0] = -1;
    q.push(0);
    order.push_back(0);
    while (!q.empty()) {   // graphTraverse, src/multicalib.cpp:825-857
        const int v = q.front();
        q.pop();
        for (auto& p : adj[v])
            if (pre[p.first] == -2) {
                pre[p.first] = v;
                preEdge[p.first] = p.second;
                q.push(p.first);
                order.push_back(p.first);
            }
    }
    for (int i = 0; i < _nCamera; ++i)
        if (pre[i] == -2 && _verbose) std::fprintf(stderr, "camera %d is not connected\n", i);
    for (size_t i = 1; i < order.size(); ++i) {
        const int v = order[i];
        const Mat44f& prePose = _vertexList[pre[v]].pose;
        const Mat44f& T = _edgeList[preEdge[v]].transform;
        if (v < _nCamera) _vertexList[v].pose = mul44(T, inv_rigid(prePose));   // :405
        else _vertexList[v].pose = mul44(inv_rigid(prePose), T);                // :416  cameraPose * photoPose = transform
    }
    _impl->initialised = true;
}

std::vector<float> MultiCameraCalibration::buildParas() const
{
    const int nV = (int)_vertexList.size();
    std::vector<float> p((size_t)(nV - 1) * 6);
    for (int i = 1; i < nV; ++i) {
        double om[3];
        log_so3(_vertexList[i].pose, om);
        for (int k = 0; k < 3; ++k) {
            p[(size_t)(i - 1) * 6 + k] = (float)om[k];
            p[(size_t)(i - 1) * 6 + 3 + k] = _vertexList[i].pose[k * 4 + 3];
        }
    }
    return p;
}

std::vector<double> MultiCameraCalibration::initialParameters() const
{
    const std::vector<float> p = buildParas();
    return std::vector<double>(p.begin(), p.end());
}

void MultiCameraCalibration::paras2vertex(const std::vector<float>& p)
{
    for (int v = 1; v < (int)_vertexList.size(); ++v) {
        const double om[3] = {p[(size_t)(v - 1) * 6], p[(size_t)(v - 1) * 6 + 1], p[(size_t)(v - 1) * 6 + 2]};
        double R[9];
        exp_so3(om, R);
        Mat44f m = eye4();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) m[i * 4 + j] = (float)R[i * 3 + j];
            m[i * 4 + 3] = p[(size_t)(v - 1) * 6 + 3 + i];
        }
        _vertexList[v].pose = m;
    }
}

double MultiCameraCalibration::optimizeExtrinsics()
{
    Impl& I = *_impl;
    if (!I.loaded) throw std::logic_error("optimizeExtrinsics() before loadImages()");
    const int nV = (int)_vertexList.size(), nE = (int)_edgeList.size(), nC = _nCamera, nF = nV - nC;
    if (nE == 0 || nF <= 0) throw std::runtime_error("no multi-camera observations to optimise");
    auto check = [&](int rc, const char* what) {
        if (rc != MCCBA_OK) throw std::runtime_error(std::string(what) + ": " + (I.h ? mccba_last_error(I.h) : "no CUDA device"));
    };
    if (!I.h) {
        mccba_options o;
        mccba_default_options(&o);
        o.device = _solver.device;
        o.verbose = _verbose;
        check(mccba_create(&o, &I.h), "mccba_create");
    }
    std::vector<int> model(nC), ndist(nC);
    std::vector<double> K5(5 * (size_t)nC), d8(8 * (size_t)nC), xi(nC);
    for (int c = 0; c < nC; ++c) {
        model[c] = I.cams[c].model; ndist[c] = I.cams[c].ndist; xi[c] = I.cams[c].xi;
        std::copy(I.cams[c].K5, I.cams[c].K5 + 5, K5.begin() + 5 * c);
        std::copy(I.cams[c].dist8, I.cams[c].dist8 + 8, d8.begin() + 8 * c);
    }
    check(mccba_set_cameras(I.h, nC, model.data(), K5.data(), d8.data(), ndist.data(), xi.data()), "mccba_set_cameras");
    std::vector<int> ecam(nE), epv(nE);
    std::vector<int64_t> eoff(nE + 1, 0);
    for (int e = 0; e < nE; ++e) {
        ecam[e] = _edgeList[e].cameraVertex;
        epv[e] = _edgeList[e].photoVertex;
        eoff[e + 1] = eoff[e] + I.images[I.edgeImage[e]].n_points;
    }
    std::vector<float> obj(3 * (size_t)eoff[nE]), img(2 * (size_t)eoff[nE]);
    for (int e = 0; e < nE; ++e) {
        const ImageRecord& im = I.images[I.edgeImage[e]];
        std::copy(I.obj.begin() + 3 * im.first, I.obj.begin() + 3 * (im.first + im.n_points), obj.begin() + 3 * eoff[e]);
        std::copy(I.img.begin() + 2 * im.first, I.img.begin() + 2 * (im.first + im.n_points), img.begin() + 2 * eoff[e]);
    }
    check(mccba_set_observations(I.h, nF, nE, ecam.data(), epv.data(), eoff.data(), obj.data(), img.data()), "mccba_set_observations");
    const std::vector<float> p32 = buildParas();          // CV_32F parameter vector, src/multicalib.cpp:426
    std::vector<double> p(p32.begin(), p32.end());
    check(mccba_set_parameters(I.h, (int64_t)p.size(), p.data()), "mccba_set_parameters");
    mccba_solve_opts so;
    mccba_default_solve_opts(&so);
    so.mode = _solver.mode;
    so.crit_type = _criteria.type;
    so.max_count = _criteria.maxCount;
    so.epsilon = _criteria.epsilon;
    so.lambda0 = _solver.lambda0; so.lambda_up = _solver.lambdaUp; so.lambda_down = _solver.lambdaDown;
    check(mccba_solve(I.h, &so, &_report), "mccba_solve");
    check(mccba_get_parameters(I.h, (int64_t)p.size(), p.data()), "mccba_get_parameters");
    mccba_error_stats st;
    std::vector<double> perEdge(nE);
    check(mccba_reproj_error(I.h, &st, perEdge.data()), "mccba_reproj_error");   // computeProjectError, :509
    for (int e = 0; e < nE; ++e) _edgeList[e].reprojecterror = (float)perEdge[e];   // :979-980
    _error = st.mean_reproj_error;
    _rms = st.rms;
    _params = p;
    std::vector<float> pf(p.begin(), p.end());
    paras2vertex(pf);                                      // :512
    return _error;
}

double MultiCameraCalibration::run()
{
    loadImages();
    initialize();
    return optimizeExtrinsics();
}

std::set<int> MultiCameraCalibration::removeOutlier(float threshold)
{
    // src/mymulticalib.cpp:406-423: edges whose mean reprojection error exceeds 0.5 px are dropped and remembered so
    // that the next loadImages() skips them.
    std::set<int> dropped;
    std::vector<edge> keep;
    std::vector<int> keepImg;
    for (size_t e = 0; e < _edgeList.size(); ++e) {
        if (_edgeList[e].reprojecterror > threshold) {
            dropped.insert(_impl->edgeImage[e]);
            _outliers.insert(_impl->edgeImage[e]);
        } else {
            keep.push_back(_edgeList[e]);
            keepImg.push_back(_impl->edgeImage[e]);
        }
    }
    _edgeList.swap(keep);
    _impl->edgeImage.swap(keepImg);
    return dropped;
}

// ---- OpenCV FileStorage XML ----------------------------------------------------------------------------------
namespace {
std::string fmt_real(double v, int digits)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.*g", digits, v);
    std::string s(buf);
    if (s.find_first_of(".eEni") == std::string::npos) s += ".";   // OpenCV writes "1." for integral reals
    return s;
}
void write_mat_f(std::ostream& os, const std::string& name, int rows, int cols, const float* data)
{
    os << "<" << name << " type_id=\"opencv-matrix\">\n  <rows>" << rows << "</rows>\n  <cols>" << cols
       << "</cols>\n  <dt>f</dt>\n  <data>\n    ";
    size_t col = 4;
    for (int i = 0; i < rows * cols; ++i) {
        const std::string s = fmt_real(data[i], 9);
        if (i > 0) {
            if (col + 1 + s.size() > 72) { os << "\n    "; col = 4; }
            else { os << " "; ++col; }
        }
        os << s;
        col += s.size();
    }
    os << "</data></" << name << ">\n";
}
}  // namespace

void MultiCameraCalibration::writeParameters(const std::string& filename)
{
    // keys, order and dtypes of src/multicalib.cpp:1092-1127
    std::ofstream os(filename);
    if (!os) throw std::runtime_error("cannot write " + filename);
    os << "<?xml version=\"1.0\"?>\n<opencv_storage>\n";
    os << "<nCameras>" << _nCamera << "</nCameras>\n";
    for (int c = 0; c < _nCamera; ++c) {
        const std::string id = std::to_string(c);
        float K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 1};
        std::vector<float> D;
        float xi = 0;
        if (c < (int)_impl->cams.size()) {
            const CameraIntrinsics& ci = _impl->cams[c];
            K[0] = (float)ci.K5[0]; K[1] = (float)ci.K5[4]; K[2] = (float)ci.K5[2];
            K[4] = (float)ci.K5[1]; K[5] = (float)ci.K5[3];
            for (int i = 0; i < ci.ndist; ++i) D.push_back((float)ci.dist8[i]);
            xi = (float)ci.xi;
        }
        write_mat_f(os, "camera_matrix_" + id, 3, 3, K);
        write_mat_f(os, "camera_distortion_" + id, 1, (int)D.size(), D.data());
        if (_camType == OMNIDIRECTIONAL) os << "<xi_" << id << ">" << fmt_real(xi, 9) << "</xi_" << id << ">\n";
        write_mat_f(os, "camera_pose_" + id, 4, 4, _vertexList[c].pose.data());
    }
    os << "<meanReprojectError>" << fmt_real(_error, 17) << "</meanReprojectError>\n";
    for (size_t v = _nCamera; v < _vertexList.size(); ++v)
        write_mat_f(os, "pose_timestamp_" + std::to_string(_vertexList[v].timestamp), 4, 4, _vertexList[v].pose.data());
    os << "</opencv_storage>\n";
}

}  // namespace mccba

// ---- plain-C access ------------------------------------------------------------------------------------------

namespace {
template <typename F>
int guarded(mccbah h, F&& f)
{
    if (!h || !h->obj) return 1;
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        h->err = e.what();
        return 2;
    }
}
}  // namespace

extern "C" {

int mccbah_create(int cameraType, int nCameras, const char* fileName, float patternWidth, float patternHeight, int verbose,
                  int showExtraction, int nMiniMatches, int flags, int critType, int critMaxCount, double critEps,
                  int mode, int device, mccbah* out)
{
    if (!out || !fileName) return 1;
    mccbah h = new mccbah_s();
    *out = h;
    try {
        mccba::SolverOptions so;
        so.mode = mode;
        so.device = device;
        h->obj = new mccba::MultiCameraCalibration(cameraType, nCameras, fileName, patternWidth, patternHeight, verbose,
                                                   showExtraction, nMiniMatches, flags,
                                                   mccba::TermCriteria(critType, critMaxCount, critEps), so);
    } catch (const std::exception& e) {
        h->err = e.what();
        return 2;
    }
    return 0;
}
int mccbah_destroy(mccbah h)
{
    if (!h) return 0;
    delete h->obj;
    delete h;
    return 0;
}
const char* mccbah_last_error(mccbah h) { return h ? h->err.c_str() : "null handle"; }
int mccbah_load_images(mccbah h) { return guarded(h, [&] { h->obj->loadImages(); }); }
int mccbah_reset(mccbah h) { return guarded(h, [&] { h->obj->reset(); }); }
int mccbah_initialize(mccbah h) { return guarded(h, [&] { h->obj->initialize(); }); }
int mccbah_optimize_extrinsics(mccbah h, double* error)
{
    return guarded(h, [&] { const double e = h->obj->optimizeExtrinsics(); if (error) *error = e; });
}
int mccbah_run(mccbah h, double* error)
{
    return guarded(h, [&] { const double e = h->obj->run(); if (error) *error = e; });
}
int mccbah_remove_outlier(mccbah h, float threshold, int* n_removed)
{
    return guarded(h, [&] { const auto s = h->obj->removeOutlier(threshold); if (n_removed) *n_removed = (int)s.size(); });
}
int mccbah_write_parameters(mccbah h, const char* filename) { return guarded(h, [&] { h->obj->writeParameters(filename); }); }
int mccbah_sizes(mccbah h, int* n_vertex, int* n_edge)
{
    return guarded(h, [&] {
        if (n_vertex) *n_vertex = (int)h->obj->vertices().size();
        if (n_edge) *n_edge = (int)h->obj->edges().size();
    });
}
int mccbah_get_indexing(mccbah h, int* edge_cam, int* edge_pv, int* edge_photo_index, int* vertex_timestamp)
{
    return guarded(h, [&] {
        const auto& E = h->obj->edges();
        for (size_t e = 0; e < E.size(); ++e) {
            if (edge_cam) edge_cam[e] = E[e].cameraVertex;
            if (edge_pv) edge_pv[e] = E[e].photoVertex;
            if (edge_photo_index) edge_photo_index[e] = E[e].photoIndex;
        }
        const auto& V = h->obj->vertices();
        if (vertex_timestamp)
            for (size_t v = 0; v < V.size(); ++v) vertex_timestamp[v] = V[v].timestamp;
    });
}
int mccbah_get_parameters(mccbah h, double* params)
{
    return guarded(h, [&] {
        const auto p = h->obj->parameters();
        std::copy(p.begin(), p.end(), params);
    });
}
int mccbah_get_initial_parameters(mccbah h, double* params)
{
    return guarded(h, [&] {
        const auto p = h->obj->initialParameters();
        std::copy(p.begin(), p.end(), params);
    });
}
int mccbah_get_stats(mccbah h, double* mean_error, double* rms, int* iterations, double* device_ms)
{
    return guarded(h, [&] {
        if (mean_error) *mean_error = h->obj->meanReprojectError();
        if (rms) *rms = h->obj->rms();
        if (iterations) *iterations = h->obj->report().iterations;
        if (device_ms) *device_ms = h->obj->report().device_ms;
    });
}
}
