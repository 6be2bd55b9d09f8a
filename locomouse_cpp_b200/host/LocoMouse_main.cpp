// LocoMouse_main.cpp — the reference's driver (main.cpp:36-106) against the B200 class mirror: same
// argument list, same call sequence, same exception handling and exit codes.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <stdexcept>

#include "LocoMouse_class.hpp"

int main(int argc, char *argv[]) {
    const auto t0 = std::chrono::steady_clock::now();
    int return_val = EXIT_SUCCESS;
    try {
        LocoMouse_ParseInputs inputs = LocoMouse_ParseInputs(argc, argv);
        std::unique_ptr<LocoMouse> L = LocoMouse_Initialize(inputs);
        L->getBoundingBox();
        L->initializeFeatureLoop();
        for (unsigned int i_frames = 0; i_frames < L->N_frames(); ++i_frames) {
            L->readFrame();
            L->cropBoundingBox();
            L->detectTail();
            L->detectBottomCandidates();
            L->computeUnaryCostsBottom();
            L->computePairwiseCostsBottom();
            L->detectSideCandidates();
            L->matchBottomSideCandidates();
            L->storePreviousImage();
        }
        L->computeBottomTracks();
        L->computeSideTracks();
        L->exportResults();
    } catch (const std::invalid_argument &e) {
        std::cout << "Invalid inputs: " << e.what() << std::endl;
        return_val = EXIT_FAILURE;
    } catch (const std::runtime_error &e) {
        std::cout << "Runtime Error: " << e.what() << std::endl;
        return_val = EXIT_FAILURE;
    }
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Total Elapsed time: " << t << "s" << std::endl;
    return return_val;
}
