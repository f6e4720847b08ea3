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
    // Self-check against the unmodified serial header on a non-trivial image (the all-ones frame of main.cpp is ~0
    // away from the window centre): the result of GenerateDoG(), of EVERY variant entry point the reference's other
    // headers add (they all forward to the same fused build), of the `for o: GaussFilter(o)` idiom, and of the superset
    // constructor fed the same pixels as floats.  Exit 0 only if every value of every level agrees exactly.
    for (int r = 0; r < side; ++r)
        for (int c = 0; c < side; ++c) pixels[(size_t)r * side + c] = (131 * r + 71 * c + (r * c) % 13) % 256;
    GaussPyramid serial(rows.data(), side, scales);
    serial.GenerateDoG();
    GaussPyramid_cuda cuda(rows.data(), side, scales);
    auto worst_vs = [&](GaussPyramid_cuda& g, GaussPyramid& ref) {
        double worst = 0.0;
        for (int o = 0; o < g.octaves(); ++o)
            for (int s = 0; s < scales + 3; ++s)
                for (int r = 0; r < g.side(o); ++r)
                    for (int c = 0; c < g.side(o); ++c)
                        worst = std::fmax(worst, std::fabs((double)g.GaussPy[o][s][r][c] - (double)ref.GaussPy[o][s][r][c]));
        return worst;
    };
    double worst = 0.0;
    struct Shim { const char* name; void (*call)(GaussPyramid_cuda&, int, char**); };
    const Shim shims[] = {
        {"GenerateDoG", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG(); }},
        {"GenerateDoG_omp", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_omp(); }},
        {"GenerateDoG_omp_dynamic", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_omp_dynamic(); }},
        {"GenerateDoG_omp_guided", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_omp_guided(); }},
        {"GenerateDoG_nomp_dynamic", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_nomp_dynamic(); }},
        {"GenerateDoG_nomp_static", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_nomp_static(); }},
        {"GenerateDoG_i", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_i(); }},
        {"GenerateDoG_n_new", [](GaussPyramid_cuda& g, int, char**) { g.GenerateDoG_n_new(); }},
        {"GenerateDoG_mpi", [](GaussPyramid_cuda& g, int ac, char** av) { g.GenerateDoG_mpi(ac, av); }},
        {"GenerateDoG_mpi_normal", [](GaussPyramid_cuda& g, int ac, char** av) { g.GenerateDoG_mpi_normal(ac, av); }},
    };
    for (const Shim& sh : shims) {
        cuda.GaussPyInit();                                      // back to the K0 state, as pThread.h:315-317 does between runs
        sh.call(cuda, argc, argv);
        const double w = worst_vs(cuda, serial);
        std::printf("%-26s max |cuda - serial header| = %g\n", sh.name, w);
        worst = std::fmax(worst, w);
    }
    {   // GaussFilter on every octave (GuassDePyramid.h:106-134), no DoG
        GaussPyramid filtered(rows.data(), side, scales);
        cuda.GaussPyInit();
        for (int o = 0; o < cuda.octaves(); ++o) { filtered.GaussFilter(o); cuda.GaussFilter(o); }
        const double w = worst_vs(cuda, filtered);
        std::printf("%-26s max |cuda - serial header| = %g\n", "GaussFilter(o) for all o", w);
        worst = std::fmax(worst, w);
    }
    {   // superset constructor: same pixels as floats, all octaves, sigma0 = the reference's 2.0
        std::vector<float> fpix(pixels.begin(), pixels.end());
        GaussPyramid_cuda sup(fpix.data(), side, side, 0, scales, 2.0f, SSPYR_MODE_REF);
        sup.GenerateDoG();
        const double w = worst_vs(sup, serial);
        std::printf("%-26s max |cuda - serial header| = %g\n", "superset ctor (float)", w);
        worst = std::fmax(worst, w);
    }
    std::printf("max |cuda - serial header| = %g\n", worst);
    return worst == 0.0 ? 0 : 1;
#else
    return 0;
#endif
}
