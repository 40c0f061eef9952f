// doubleside.cpp -- DoubleSideCalibration: double-sided board in front of FIXED cameras (include/opencv2/ccalib/
// doubleSide.hpp:82-178; src/doubleSide.cpp:114-118 storeReaded, :119-165 initial transform, :167-231 initialize,
// :233-261 buildParas, :276-287 loadCameraPose, :582-590 writeParameters, :612-638 paras2vertex).  The optimisation is
// mccba_ds_* (include/mccba.h).  The reference marks the class "not working" (README.md:22); this file follows its
// model -- cameraPose * photoPose * doubleSideTransform = back pose (:474-475) -- not its debugging state.
#include <cmath>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "host_impl.hpp"

namespace mccba {
namespace {
using M4 = std::array<double, 16>;
M4 eye()
{
    M4 m{};
    m[0] = m[5] = m[10] = m[15] = 1.0;
    return m;
}
M4 mul(const M4& a, const M4& b)
{
    M4 c{};
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
    return c;
}
M4 inv_rigid(const M4& a)
{
    M4 r = eye();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i * 4 + j] = a[j * 4 + i];
    for (int i = 0; i < 3; ++i) r[i * 4 + 3] = -(r[i * 4] * a[3] + r[i * 4 + 1] * a[7] + r[i * 4 + 2] * a[11]);
    return r;
}
M4 widen(const Mat44f& a)
{
    M4 r;
    for (int i = 0; i < 16; ++i) r[i] = a[i];
    return r;
}
void to_rt(const M4& m, double* p6)   // [rvec | tvec]
{
    const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
    log_so3_3x3(R, p6);
    p6[3] = m[3]; p6[4] = m[7]; p6[5] = m[11];
}
M4 from_rt(const double* p6)
{
    double R[9];
    exp_so3(p6, R);
    M4 m = eye();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) m[i * 4 + j] = R[i * 3 + j];
        m[i * 4 + 3] = p6[3 + i];
    }
    return m;
}
}  // namespace

DoubleSideCalibration::DoubleSideCalibration(const std::vector<std::string>& cameraSerials, int cameraType, int nCameras,
                                             const std::string& dataFolder, const std::string& cameraConfigFolder, Size frontPatternSize,
                                             Size backPatternSize, float patternWidth, float patternHeight, int verbose, int showExtration,
                                             int nMiniMatches, int flags, TermCriteria criteria, SolverOptions solver)
    : MyMultiCameraCalibration(cameraSerials, cameraType, nCameras, dataFolder, cameraConfigFolder, "", frontPatternSize, backPatternSize,
                               patternWidth, patternHeight, verbose, showExtration, nMiniMatches, flags, criteria, solver)
{
    if (frontPatternSize.width * frontPatternSize.height == backPatternSize.width * backPatternSize.height)
        throw std::invalid_argument("front and back pattern must differ in their corner count (isBackPattern tells them apart by it)");
    _keepBackPattern = true;
    _dst = eye();
    loadCameraPose();
}

void DoubleSideCalibration::loadCameraPose()   // src/doubleSide.cpp:276-287
{
    _camPose.assign(_serials.size(), eye());
    for (size_t c = 0; c < _serials.size(); ++c) {
        const std::string filename = _configFolder + "/" + _serials[c] + ".xml";
        CvFileReader f(filename);
        CvMatrix P;
        if (!f.matrix("CameraMatrix", P) || P.data.size() != 16) throw std::runtime_error(filename + ": no 4x4 \"CameraMatrix\" (camera pose)");
        for (int i = 0; i < 16; ++i) _camPose[c][i] = P.data[i];
    }
}

void DoubleSideCalibration::initialize()
{
    Impl& I = *_impl;
    if (!I.loaded) throw std::logic_error("initialize() before loadImages()");
    const int nC = _nCamera, nV = (int)_vertexList.size();
    for (int c = 0; c < nC; ++c)
        for (int i = 0; i < 16; ++i) _vertexList[c].pose[i] = (float)_camPose[c][i];
    // first edge of each side per photo vertex (edges are in load order: cameras outer loop)
    std::vector<int> front(nV, -1), back(nV, -1);
    for (int e = 0; e < (int)_edgeList.size(); ++e) {
        std::vector<int>& slot = _edgeList[e].patternSide == BACK_PATTERN ? back : front;
        if (slot[_edgeList[e].photoVertex] < 0) slot[_edgeList[e].photoVertex] = e;
    }
    // initializeDoublesideTransform (:150-165): a timestamp seen from both sides;  front pose = camPose^-1 * transform,
    // doubleSideTransform = frontpose^-1 * backpose (findTransformOfTwoEdge :139-143)
    int both = -1;
    for (int v = nC; v < nV && both < 0; ++v)
        if (front[v] >= 0 && back[v] >= 0) both = v;
    if (both < 0) throw std::runtime_error("no timestamp is seen from the front by one camera and from the back by another: the front<->back transform cannot be initialised");
    auto worldPose = [&](int e) { return mul(inv_rigid(_camPose[_edgeList[e].cameraVertex]), widen(_edgeList[e].transform)); };
    _dst = mul(inv_rigid(worldPose(front[both])), worldPose(back[both]));
    const M4 dstInv = inv_rigid(_dst);
    for (int v = nC; v < nV; ++v) {   // :196-229: board pose = camPose^-1 * transform (* doubleSideTransform^-1 for a back edge)
        M4 P;
        if (front[v] >= 0) P = worldPose(front[v]);
        else if (back[v] >= 0) P = mul(worldPose(back[v]), dstInv);
        else continue;
        for (int i = 0; i < 16; ++i) _vertexList[v].pose[i] = (float)P[i];
    }
    I.initialised = true;
}

double DoubleSideCalibration::optimizeExtrinsics()
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
    std::vector<unsigned char> back(nE);
    for (int e = 0; e < nE; ++e) {
        ecam[e] = _edgeList[e].cameraVertex;
        epv[e] = _edgeList[e].photoVertex;
        back[e] = _edgeList[e].patternSide == BACK_PATTERN ? 1 : 0;
        eoff[e + 1] = eoff[e] + I.images[I.edgeImage[e]].n_points;
    }
    std::vector<float> obj(3 * (size_t)eoff[nE]), img(2 * (size_t)eoff[nE]);
    for (int e = 0; e < nE; ++e) {
        const ImageRecord& im = I.images[I.edgeImage[e]];
        std::copy(I.obj.begin() + 3 * im.first, I.obj.begin() + 3 * (im.first + im.n_points), obj.begin() + 3 * eoff[e]);
        std::copy(I.img.begin() + 2 * im.first, I.img.begin() + 2 * (im.first + im.n_points), img.begin() + 2 * eoff[e]);
    }
    check(mccba_set_observations(I.h, nF, nE, ecam.data(), epv.data(), eoff.data(), obj.data(), img.data()), "mccba_set_observations");
    std::vector<double> camPose(6 * (size_t)nC);
    for (int c = 0; c < nC; ++c) to_rt(_camPose[c], camPose.data() + 6 * c);       // cameraPose2vec, :262-274
    check(mccba_ds_set_problem(I.h, back.data(), camPose.data()), "mccba_ds_set_problem");
    // buildParas (:233-261): [doubleSideTransform | photo vertices], stored CV_32F
    std::vector<double> p(6 + 6 * (size_t)nF);
    to_rt(_dst, p.data());
    for (int v = nC; v < nV; ++v) to_rt(widen(_vertexList[v].pose), p.data() + 6 * (size_t)(v - nC + 1));
    for (double& x : p) x = (double)(float)x;
    check(mccba_ds_set_parameters(I.h, (int64_t)p.size(), p.data()), "mccba_ds_set_parameters");
    check(mccba_ds_solve(I.h, _criteria.type, _criteria.maxCount, _criteria.epsilon, &_report), "mccba_ds_solve");
    check(mccba_ds_get_parameters(I.h, (int64_t)p.size(), p.data()), "mccba_ds_get_parameters");
    _params = p;
    _dst = from_rt(p.data());                                                       // paras2vertex, :612-638
    for (int v = nC; v < nV; ++v) {
        const M4 P = from_rt(p.data() + 6 * (size_t)(v - nC + 1));
        for (int i = 0; i < 16; ++i) _vertexList[v].pose[i] = (float)P[i];
    }
    _rms = std::sqrt(_report.cost / (double)eoff[nE]);
    _error = _rms;
    return _rms;
}

double DoubleSideCalibration::run()
{
    loadImages();
    initialize();
    return optimizeExtrinsics();
}

void DoubleSideCalibration::writeParameters(const std::string& filename)   // writeDoubleSideTransform, :582-590
{
    std::ofstream os(filename);
    if (!os) throw std::runtime_error("cannot write " + filename);
    os << "%YAML:1.0\n---\ntransform: !!opencv-matrix\n   rows: 4\n   cols: 4\n   dt: d\n   data: [ ";
    for (int i = 0; i < 16; ++i) {
        os << cv_format_real(_dst[i], 17);
        if (i != 15) os << (i % 4 == 3 ? ",\n       " : ", ");
    }
    os << " ]\n";
}

}  // namespace mccba

// ---- plain-C access ------------------------------------------------------------------------------------------
namespace {
template <typename F>
int guarded_ds(mccbah h, F&& f)
{
    if (!h) return 1;
    mccba::DoubleSideCalibration* o = dynamic_cast<mccba::DoubleSideCalibration*>(h->obj);
    if (!o) { h->err = "the handle is not a DoubleSideCalibration"; return 1; }
    try {
        f(*o);
        return 0;
    } catch (const std::exception& e) {
        h->err = e.what();
        return 1;
    }
}
}  // namespace

extern "C" {

int mccbah_create_ds(const char* serials, int cameraType, int nCameras, const char* dataFolder, const char* cameraConfigFolder, int frontW,
                     int frontH, int backW, int backH, float patternWidth, float patternHeight, int verbose, int critType, int critMaxCount,
                     double critEps, int device, mccbah* out)
{
    if (!out || !serials || !dataFolder || !cameraConfigFolder) return 1;
    mccbah h = new mccbah_s();
    *out = h;
    try {
        std::vector<std::string> names;
        std::string cur;
        for (const char* c = serials; *c; ++c) {
            if (*c == ',') { names.push_back(cur); cur.clear(); }
            else cur.push_back(*c);
        }
        if (!cur.empty()) names.push_back(cur);
        mccba::SolverOptions so;
        so.device = device;
        h->obj = new mccba::DoubleSideCalibration(names, cameraType, nCameras, dataFolder, cameraConfigFolder, mccba::Size(frontW, frontH),
                                                  mccba::Size(backW, backH), patternWidth, patternHeight, verbose, 0, 20, 0,
                                                  mccba::TermCriteria(critType, critMaxCount, critEps), so);
        return 0;
    } catch (const std::exception& e) {
        h->err = e.what();
        return 1;
    }
}
int mccbah_initialize_ds(mccbah h) { return guarded_ds(h, [&](mccba::DoubleSideCalibration& o) { o.initialize(); }); }
int mccbah_optimize_ds(mccbah h, double* rms)
{
    return guarded_ds(h, [&](mccba::DoubleSideCalibration& o) {
        const double r = o.optimizeExtrinsics();
        if (rms) *rms = r;
    });
}
int mccbah_get_double_side_transform(mccbah h, double* T16)
{
    return guarded_ds(h, [&](mccba::DoubleSideCalibration& o) { std::memcpy(T16, o.doubleSideTransform().data(), 16 * sizeof(double)); });
}
}
