// Counterpart of the reference's samples/multi_cameras_calibration.cpp (call sequence at :55-83) on the B200 drop-in.
//   multi_cameras_calibration <observations.mccb> <nCameras> <output.xml> [cameraType=0] [mode: 0 reference GN | 1 LM]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mccba_host.hpp"

int main(int argc, char** argv)
{
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <observations.mccb> <nCameras> <output.xml> [cameraType] [mode]\n", argv[0]);
        return 2;
    }
    const std::string file = argv[1], out = argv[3];
    const int nCameras = std::atoi(argv[2]);
    const int cameraType = argc > 4 ? std::atoi(argv[4]) : mccba::MultiCameraCalibration::PINHOLE;
    mccba::SolverOptions so;
    so.mode = argc > 5 ? std::atoi(argv[5]) : MCCBA_MODE_REFERENCE_GN;
    try {
        // TermCriteria(COUNT + EPS, 200, 1e-7): the MyMultiCameraCalibration default (mymulticalib.hpp:95)
        mccba::MultiCameraCalibration calib(cameraType, nCameras, file, 360.f, 200.f, 1, 0, 20, 0,
                                            mccba::TermCriteria(mccba::TermCriteria::COUNT + mccba::TermCriteria::EPS, 200, 1e-7), so);
        calib.loadImages();
        calib.initialize();
        double err = calib.optimizeExtrinsics();
        std::printf("pass 1: mean reprojection error %.6f px (rms %.6f), %d iterations, %.3f ms on device\n", err, calib.rms(),
                    calib.report().iterations, calib.report().device_ms);
        const auto outliers = calib.removeOutlier(0.5f);          // reference: edges above 0.5 px are dropped
        if (!outliers.empty()) {
            std::printf("%zu outlier images removed, second pass\n", outliers.size());
            calib.reset();
            calib.loadImages();
            calib.initialize();
            err = calib.optimizeExtrinsics();
            std::printf("pass 2: mean reprojection error %.6f px (rms %.6f)\n", err, calib.rms());
        }
        calib.writeParameters(out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
