// examples/main.cpp -- the reference's driver (main.cpp:26-75) switched to the B200 build: same image,
// same constructor call, same loop-until-100-ms timing, one #include and one type name changed.
//
//   g++ -O2 -std=gnu++14 -Iinclude examples/main.cpp -Lsift-parallel-optimization_b200 -lsspyr \
//       -Wl,-rpath,$PWD/sift-parallel-optimization_b200 -o build/main_cuda
//
// With /root/reference on the include path and -DWITH_REFERENCE it also runs the serial header on the same
// image and reports the largest difference (expected: 0).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#ifdef WITH_REFERENCE
#include "GuassDePyramid.h"
#endif
#include "GaussDePyramid-CUDA.h"

using namespace std;
const int MAX = 1024;
int n = 512;

int main(int argc, char* argv[]) {
    if (argc > 1) n = atoi(argv[1]);
    if (n < 1 || n > MAX) { cerr << "n must be in 1.." << MAX << endl; return 2; }
    int** p = new int*[MAX];
    for (int i = 0; i < MAX; ++i) {
        p[i] = new int[MAX];
    }
    for (int i = 0; i < MAX; ++i) {
        for (int j = 0; j < MAX; ++j) {
            p[i][j] = 1;
        }
    }
    int times = 0;
    GaussPyramid_cuda g(p, n, 2);                       // main.cpp:61
    std::chrono::duration<double, std::milli> elapsed{};
    double device_ms = 0;
    while (elapsed.count() < 100) {                     // main.cpp:66-73
        auto start = std::chrono::high_resolution_clock::now();
        g.GenerateDoG();
        auto end = std::chrono::high_resolution_clock::now();
        elapsed += end - start;
        device_ms += g.last_device_ms();
        times += 1;
    }
    cout << float(elapsed.count()) / float(times) << endl;   // main.cpp:74: mean ms per call (host wall clock)
    cout << "device ms per call: " << device_ms / times << "  calls: " << times << "  octaves: " << g.octaves() << endl;
#ifdef WITH_REFERENCE
    GaussPyramid r(p, n, 2);
    r.GenerateDoG();
    double worst = 0;
    for (int o = 0; o < g.octaves(); ++o)
        for (int s = 0; s < 2 + 3; ++s)
            for (int i = 0; i < g.side(o); ++i)
                for (int j = 0; j < g.side(o); ++j)
                    worst = fmax(worst, fabs((double)g.GaussPy[o][s][i][j] - (double)r.GaussPy[o][s][i][j]));
    cout << "max |cuda - serial header| = " << worst << endl;
    return worst == 0 ? 0 : 1;
#else
    return 0;
#endif
}
