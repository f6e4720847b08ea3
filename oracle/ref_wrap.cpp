// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shell around the UNMODIFIED reference headers, compiled where they lie
// (-I/root/reference) into oracle/_ref/libsiftref.so by oracle/Makefile.  No reference source is
// copied into this repository; this file only instantiates the reference classes and flattens their
// jagged float**** result into one dense array:
//
//     out[ o ][ s ][ r ][ c ]   o < octaves, s < S+3, r,c < len>>o      (planes back to back, no padding)
//
// Used by (and only by) tests/, oracle/make_golden.py, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py.
//
// Reference entry points driven here:
//   GaussPyramid::GaussPyramid(int**,int,int)   GuassDePyramid.h:36-58
//   GaussPyramid::GaussPyInit()                 GuassDePyramid.h:60-87
//   GaussPyramid::GaussFilter(int)              GuassDePyramid.h:106-134
//   GaussPyramid::GenerateDoG()                 GuassDePyramid.h:136-149
//   GaussPyramid_p::GenerateDoG_i()             GaussDePyramid-pThread.h:328-342   (bit-exact vs serial)
//   GaussPyramid_omp::GenerateDoG()             GaussDePyramid-OpenMP.h:180-197    (bit-exact vs serial)
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "GuassDePyramid.h"
#include "GaussDePyramid-pThread.h"
#include "GaussDePyramid-OpenMP.h"

namespace {

// `layer` / `length` are protected in every reference class: reach them through a subclass.
template <class Base>
struct Peek : Base {
    using Base::Base;
    int octaves() const { return this->layer; }
    int side() const { return this->length; }
};

struct IntImage {
    std::vector<int*> rows;
    IntImage(const int32_t* img, int len) : rows(len) {
        for (int i = 0; i < len; ++i) rows[i] = const_cast<int*>(img) + (size_t)i * len;
    }
    int** ptr() { return rows.data(); }
};

template <class Pyr>
long long flatten(Pyr& g, int len, int S, float* out) {
    long long n = 0;
    int lo = len;
    for (int o = 0; o < g.octaves(); ++o) {
        for (int s = 0; s < S + 3; ++s)
            for (int r = 0; r < lo; ++r) {
                std::memcpy(out + n, g.GaussPy[o][s][r], sizeof(float) * (size_t)lo);
                n += lo;
            }
        lo /= 2;
    }
    return n;
}

double now_ms() {
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

// Time `reps` calls of fn(g) after `warm` untimed ones; GaussPyInit() (untimed unless include_init)
// resets the levels from `data` before every call, as GaussDePyramid-pThread.h:315-317 does.
template <class Pyr, class Fn>
int time_loop(Pyr& g, Fn fn, int warm, int reps, int include_init, double* ms_each) {
    for (int i = 0; i < warm; ++i) { g.GaussPyInit(); fn(g); }
    for (int i = 0; i < reps; ++i) {
        double t0;
        if (include_init) { t0 = now_ms(); g.GaussPyInit(); }
        else              { g.GaussPyInit(); t0 = now_ms(); }
        fn(g);
        ms_each[i] = now_ms() - t0;
    }
    return reps;
}

}  // namespace

extern "C" {

// floor(log2(len))+1, by the reference's own loop (GuassDePyramid.h:48-53).
int sref_octaves(int len) {
    int x = 0;
    while (len) { x++; len /= 2; }
    return x;
}

long long sref_total_floats(int len, int S) {
    long long n = 0;
    for (int lo = len; lo; lo /= 2) n += (long long)(S + 3) * lo * lo;
    return n;
}

// Reference result layout after GenerateDoG(): slots 0..S+1 = DoG_s, slot S+2 = G_{S+2}.
long long sref_serial_dog(const int32_t* img, int len, int S, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid> g(im.ptr(), len, S);
    g.GenerateDoG();
    return flatten(g, len, S, out);
}

// Gaussian ("window-multiplied") levels G_0..G_{S+2}: public GaussFilter(o) on every octave, no DoG.
long long sref_serial_gauss(const int32_t* img, int len, int S, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid> g(im.ptr(), len, S);
    for (int o = 0; o < g.octaves(); ++o) g.GaussFilter(o);
    return flatten(g, len, S, out);
}

// K0 only: every level = decimated original cast to float (GuassDePyramid.h:76-86).
long long sref_serial_init(const int32_t* img, int len, int S, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid> g(im.ptr(), len, S);
    return flatten(g, len, S, out);
}

long long sref_pthread_i_dog(const int32_t* img, int len, int S, int threads, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid_p> g(im.ptr(), len, S);
    GaussPyramid_p::THREAD_COUNT = threads > 0 ? threads : 7;
    g.GenerateDoG_i();
    return flatten(g, len, S, out);
}

long long sref_omp_dog(const int32_t* img, int len, int S, int threads, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid_omp> g(im.ptr(), len, S);
    if (threads > 0) g.thread_count = threads;
    g.GenerateDoG();
    return flatten(g, len, S, out);
}

// ---- timing legs (wall clock, like main.cpp:62-74 but with warm-up and a reset before each rep) ----
int sref_time_serial(const int32_t* img, int len, int S, int warm, int reps, int include_init,
                     double* ms_each) {
    IntImage im(img, len);
    Peek<GaussPyramid> g(im.ptr(), len, S);
    return time_loop(g, [](GaussPyramid& p) { p.GenerateDoG(); }, warm, reps, include_init, ms_each);
}

int sref_time_pthread_i(const int32_t* img, int len, int S, int threads, int warm, int reps,
                        int include_init, double* ms_each) {
    IntImage im(img, len);
    Peek<GaussPyramid_p> g(im.ptr(), len, S);
    GaussPyramid_p::THREAD_COUNT = threads > 0 ? threads : 7;
    return time_loop(g, [](GaussPyramid_p& p) { p.GenerateDoG_i(); }, warm, reps, include_init, ms_each);
}

int sref_time_omp(const int32_t* img, int len, int S, int threads, int warm, int reps,
                  int include_init, double* ms_each) {
    IntImage im(img, len);
    Peek<GaussPyramid_omp> g(im.ptr(), len, S);
    if (threads > 0) g.thread_count = threads;
    return time_loop(g, [](GaussPyramid_omp& p) { p.GenerateDoG(); }, warm, reps, include_init, ms_each);
}

}  // extern "C"
