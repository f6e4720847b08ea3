// examples/drop_in_driver.cpp -- what the reference's driver does (main.cpp:26-75: build an n x n all-ones image,
// construct the pyramid once with S = 2, call GenerateDoG repeatedly until 100 ms of wall clock have accumulated,
// print the mean milliseconds per call), written against the B200 build.  Switching the reference's own main.cpp
// takes two edits: `#include "GaussDePyramid-CUDA.h"` and the type name `GaussPyramid_cuda`.
//
//   g++ -O2 -std=gnu++14 -Iinclude examples/drop_in_driver.cpp -Lsift-parallel-optimization_b200 -lsspyr \
//       -Wl,-rpath,$PWD/sift-parallel-optimization_b200 -o build/main_cuda
//   build/main_cuda [n]            (n <= 4096, default 512)
//
// Built with -I/root/reference -DWITH_REFERENCE it also runs the serial header on the same image and exits 0 only
// if every value of every level agrees exactly.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifdef WITH_REFERENCE
#include "GuassDePyramid.h"
#endif
#include "GaussDePyramid-CUDA.h"

int main(int argc, char* argv[]) {
    const int side = argc > 1 ? std::atoi(argv[1]) : 512;
    const int scales = 2;                                        // the S the reference driver passes (main.cpp:61)
    if (side < 1 || side > 4096) { std::fprintf(stderr, "side must be in 1..4096\n"); return 2; }

    std::vector<int> pixels((size_t)side * side, 1);             // all ones, like main.cpp:31-35
    std::vector<int*> rows(side);
    for (int r = 0; r < side; ++r) rows[r] = pixels.data() + (size_t)r * side;

    GaussPyramid_cuda pyramid(rows.data(), side, scales);

    using clock = std::chrono::steady_clock;
    double wall_ms = 0.0, device_ms = 0.0;
    int calls = 0;
    do {                                                         // accumulate >= 100 ms, as main.cpp:66-73
        const auto t0 = clock::now();
        pyramid.GenerateDoG();
        wall_ms += std::chrono::duration<double, std::milli>(clock::now() - t0).count();
        device_ms += pyramid.last_device_ms();
        ++calls;
    } while (wall_ms < 100.0);
    std::printf("%g\n", wall_ms / calls);                        // mean ms per call (main.cpp:74)
    std::printf("device ms per call: %g  calls: %d  octaves: %d\n", device_ms / calls, calls, pyramid.octaves());

#ifdef WITH_REFERENCE
    GaussPyramid serial(rows.data(), side, scales);
    serial.GenerateDoG();
    double worst = 0.0;
    for (int o = 0; o < pyramid.octaves(); ++o)
        for (int s = 0; s < scales + 3; ++s)
            for (int r = 0; r < pyramid.side(o); ++r)
                for (int c = 0; c < pyramid.side(o); ++c)
                    worst = std::fmax(worst, std::fabs((double)pyramid.GaussPy[o][s][r][c] - (double)serial.GaussPy[o][s][r][c]));
    std::printf("max |cuda - serial header| = %g\n", worst);
    return worst == 0.0 ? 0 : 1;
#else
    return 0;
#endif
}
