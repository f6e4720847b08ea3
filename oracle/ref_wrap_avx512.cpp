// oracle/ref_wrap_avx512.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" shell around the reference's AVX-512 variants, compiled where they lie
// (-I/root/reference) into oracle/_ref/libsiftref_avx512.so.  Kept in its own library because the
// host of a GPU box may lack AVX-512: callers check the CPU flags before loading it.
//
// Both headers use ALIGNED _mm512_store_ps on rows that come from plain `new float[]`
// (GaussDePyramid-AVX512xPTHREAD.h:231,239,249; GaussDePyramid-AVX512xOpenMP.h:295,306,353) and
// fault on glibc.  The no-source-change shim below (redirect the store to the unaligned form
// before the headers are read) is the only deviation; the reference files are compiled untouched.
//
//   GaussPyramid_a512xp::GenerateDoG()                GaussDePyramid-AVX512xPTHREAD.h:143-155,178-261
//        7 threads, octave-cyclic; bit-exact vs the serial header on power-of-two sides >= 16.
//   GaussPyramid_a512omp::GenerateDoG_nomp_dynamic()  GaussDePyramid-AVX512xOpenMP.h:240-364
//        the "fastest CPU variant" BASELINE.json names.  Filters S of the S+3 levels (:242) and
//        subtracts S-1 pairs with a race across `omp for` iterations (:337,:346): TIMING ONLY,
//        never a parity oracle.
#include <immintrin.h>
#define _mm512_store_ps _mm512_storeu_ps

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "GuassDePyramid.h"
#include "GaussDePyramid-AVX512xPTHREAD.h"
#include "GaussDePyramid-AVX512xOpenMP.h"

namespace {

template <class Base>
struct Peek : Base {
    using Base::Base;
    int octaves() const { return this->layer; }
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

// K0 by hand: the a512 classes' GaussPyInit() re-allocates every row on every call (no
// `initialized` guard, AVX512xPTHREAD.h:69-80), so the timing loop resets the levels itself with
// the same assignment the reference uses (GuassDePyramid.h:76-86).
template <class Pyr>
void reset_levels(Pyr& g, int len, int S) {
    int lo = len, step = 1;
    for (int o = 0; o < g.octaves(); ++o) {
        for (int s = 0; s < S + 3; ++s)
            for (int r = 0; r < lo; ++r)
                for (int c = 0; c < lo; ++c) g.GaussPy[o][s][r][c] = g.data[r * step][c * step];
        lo /= 2;
        step *= 2;
    }
}

}  // namespace

extern "C" {

long long sref_a512xp_dog(const int32_t* img, int len, int S, float* out) {
    IntImage im(img, len);
    Peek<GaussPyramid_a512xp> g(im.ptr(), len, S);
    g.GenerateDoG();
    return flatten(g, len, S, out);
}

int sref_time_a512xp(const int32_t* img, int len, int S, int warm, int reps, double* ms_each) {
    IntImage im(img, len);
    Peek<GaussPyramid_a512xp> g(im.ptr(), len, S);
    for (int i = 0; i < warm; ++i) { reset_levels(g, len, S); g.GenerateDoG(); }
    for (int i = 0; i < reps; ++i) {
        reset_levels(g, len, S);
        double t0 = now_ms();
        g.GenerateDoG();
        ms_each[i] = now_ms() - t0;
    }
    return reps;
}

// threads -> the reference's global `counnt` (AVX512xOpenMP.h:18, default 2).
int sref_time_a512omp(const int32_t* img, int len, int S, int threads, int warm, int reps,
                      double* ms_each) {
    IntImage im(img, len);
    Peek<GaussPyramid_a512omp> g(im.ptr(), len, S);
    if (threads > 0) counnt = threads;
    for (int i = 0; i < warm; ++i) { reset_levels(g, len, S); g.GenerateDoG_nomp_dynamic(); }
    for (int i = 0; i < reps; ++i) {
        reset_levels(g, len, S);
        double t0 = now_ms();
        g.GenerateDoG_nomp_dynamic();
        ms_each[i] = now_ms() - t0;
    }
    return reps;
}

}  // extern "C"
