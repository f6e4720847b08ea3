// include/GaussDePyramid-CUDA.h -- the reference-side binding: a header-only C++14 class with the same
// public surface as the reference's `GaussPyramid` (GuassDePyramid.h:11-29), backed by libsspyr.so.
//
// The reference selects an implementation at compile time by `#include "<variant>.h"` and a class name
// (main.cpp:2-13,61).  To switch a driver to the B200 build:
//
//     #include "GuassDePyramid.h"            // optional: this header does not need it and does not
//     #include "GaussDePyramid-CUDA.h"       //           redefine its `sigma` / `PI` constants
//     ...
//     GaussPyramid_cuda g(p, n, 2);          // was: GaussPyramid_mpi g(p,n,2);       main.cpp:61
//     g.GenerateDoG();                       // was: g.GenerateDoG_mpi(argc,argv);    main.cpp:68
//     float v = g.GaussPy[o][s][r][c];       // unchanged: public float**** GaussPy   GuassDePyramid.h:16
//
// and link with  -lsspyr  (no CUDA headers or nvcc needed on the caller's side).
//
// Semantics kept: constructor deep-copies the image (:38-46) and runs GaussPyInit (:57); octave count
// floor(log2 len)+1 (:48-53); S+3 levels per octave (:64); GaussPyInit() resets every level to the
// decimated original (:76-86); GaussFilter(o) leaves the window-multiplied levels of octave o in
// GaussPy[o] (:106-134); GenerateDoG() leaves DoG_s = G_s - G_{s+1} in slots 0..S+1 and G_{S+2} in slot
// S+2 (:136-149), bit-identical to the serial header.  One difference, on purpose: every call recomputes
// from the image uploaded by the last GaussPyInit(), so calling GenerateDoG() twice gives the same result
// twice (the reference multiplies its stored levels again -- an artefact of main.cpp:66-73's loop).
// Errors (the reference has none) throw std::runtime_error carrying sspyr_last_error().
#ifndef SIFT_GUASS_NORMAL_GAUSSDEPYRAMID_CUDA_H
#define SIFT_GUASS_NORMAL_GAUSSDEPYRAMID_CUDA_H

#include <cstddef>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sspyr.h"

class GaussPyramid_cuda {
public:
    int** data;  // image grey values, as in the reference (public, caller may edit then call GaussPyInit())
    GaussPyramid_cuda() : data(nullptr), GaussPy(nullptr), initialized(false), length(0), S(0), layer(0),
                          filter(nullptr), h_(nullptr), mirror_(nullptr), staging_(nullptr), download_(true),
                          rows_(0), cols_(0), float_pixels_(false), mode_(SSPYR_MODE_REF), state_(NOTHING) {}

    // The reference's constructor (GuassDePyramid.h:36-58): square int image, all octaves, sigma 2.0.
    GaussPyramid_cuda(int** img, int len, int S_) : GaussPyramid_cuda() {
        length = rows_ = cols_ = len;
        S = S_;
        data = new int*[len]();                                 // GuassDePyramid.h:38-46
        for (int i = 0; i < len; ++i) {
            data[i] = new int[len];
            std::memcpy(data[i], img[i], sizeof(int) * (size_t)len);
        }
        setup(0, 0.0f);
        GaussPyInit();                                          // :57
    }

    // Superset constructor: a dense row-major H x W float image, a chosen octave count
    // (0 = all), S, sigma0 (<= 0: the mode's default) and the mode (SSPYR_MODE_REF: the reference's pointwise window;
    // SSPYR_MODE_CONV: the separable blur chain).  `data` stays null: the float pixels are kept in a private copy,
    // reachable through pixels() -- edit them there, then call GaussPyInit() as with the reference's `data`.
    GaussPyramid_cuda(const float* img, int H, int W, int octaves, int S_, float sigma0 = 0.0f, int mode = SSPYR_MODE_REF)
        : GaussPyramid_cuda() {
        rows_ = H;
        cols_ = W;
        length = H < W ? H : W;
        S = S_;
        float_pixels_ = true;
        mode_ = mode;
        setup(octaves, sigma0);
        std::memcpy(staging_, img, sizeof(float) * (size_t)H * W);
        GaussPyInit();
    }

    // K0 (GuassDePyramid.h:60-87): re-read `data`, upload, every level := decimated original.
    void GaussPyInit() {
        check(sspyr_sync(h_), "sspyr_sync");                    // an earlier asynchronous upload may still be reading staging_
        if (!float_pixels_)
            for (int i = 0; i < rows_; ++i)
                std::memcpy(static_cast<int*>(staging_) + (size_t)i * cols_, data[i], sizeof(int) * (size_t)cols_);
        check(sspyr_upload(h_, 0, staging_, 0), "sspyr_upload");
        state_ = NOTHING;
        if (mode_ == SSPYR_MODE_REF) {                          // (the blur chain has no "levels = decimated original" state)
            check(sspyr_build_stage(h_, 0, SSPYR_STAGE_INIT), "sspyr_build_stage");
            state_ = INIT;
            if (download_) { check(sspyr_download_gauss(h_, 0, mirror_), "sspyr_download_gauss"); check(sspyr_sync(h_), "sspyr_sync"); }
        }
        initialized = true;
    }

    // GuassDePyramid.h:106-134: the S+3 filtered levels of octave `theLayer`.  The device builds all octaves in one
    // fused pass; that pass runs once and is reused until the next GaussPyInit()/GenerateDoG(), so the reference's
    // `for (o...) GaussFilter(o)` idiom costs ONE build plus one download per octave.
    void GaussFilter(int theLayer) {
        if (theLayer < 0 || theLayer >= layer) throw std::out_of_range("GaussFilter: theLayer");
        if (state_ != FILTERED) {                               // (a full build also leaves every Gaussian level in place)
            if (mode_ == SSPYR_MODE_REF) check(sspyr_build_stage(h_, 0, SSPYR_STAGE_FILTER), "sspyr_build_stage");
            else check(sspyr_build(h_, 0), "sspyr_build");      // CONV: the Gaussian levels are outputs of the full build
            state_ = FILTERED;
        }
        for (int s = 0; s < S + 3; ++s)
            check(sspyr_download(h_, 0, theLayer, s, SSPYR_KIND_GAUSS, GaussPy[theLayer][s][0], 0), "sspyr_download");
    }

    // GuassDePyramid.h:136-149: the whole pipeline, fused on the GPU.
    void GenerateDoG() {
        check(sspyr_build(h_, 0), "sspyr_build");
        state_ = FILTERED;
        if (download_) check(sspyr_download_inplace(h_, 0, mirror_), "sspyr_download_inplace");
        check(sspyr_sync(h_), "sspyr_sync");
    }

    // The entry points the reference's variants add; all forward to the same fused GPU build, so any of
    // the reference's drivers compiles against this class unchanged
    // (OpenMP.h:44-48, AVX512xOpenMP.h:37-39, pThread.h:50, MPI.h:35, NEON.h).
    void GenerateDoG_omp() { GenerateDoG(); }
    void GenerateDoG_omp_dynamic() { GenerateDoG(); }
    void GenerateDoG_omp_guided() { GenerateDoG(); }
    void GenerateDoG_nomp_dynamic() { GenerateDoG(); }
    void GenerateDoG_nomp_static() { GenerateDoG(); }
    void GenerateDoG_i() { GenerateDoG(); }
    void GenerateDoG_n_new() { GenerateDoG(); }
    void GenerateDoG_mpi(int, char**) { GenerateDoG(); }
    void GenerateDoG_mpi_normal(int, char**) { GenerateDoG(); }

    // GuassDePyramid.h:89-104: level 0 of every octave, one text row per image row, a ruler of "==" per octave.
    void output() {
        for (int o = 0; o < layer; ++o) {
            for (int r = 0; r < rows(o); ++r) {
                const float* row = GaussPy[o][0][r];
                for (int c = 0; c < cols(o); ++c) std::cout << row[c] << " ";
                std::cout << std::endl;
            }
            std::cout << std::string(2 * (size_t)cols(o), '=') << std::endl;
        }
    }

    ~GaussPyramid_cuda() {                                      // GuassDePyramid.h:151-170 (and frees `data`)
        if (GaussPy) {
            for (int o = 0; o < layer; ++o) {
                if (!GaussPy[o]) continue;                      // (tables are value-initialised: a constructor that threw half-way
                for (int s = 0; s < S + 3; ++s) delete[] GaussPy[o][s];   //  leaves null entries, never garbage)
                delete[] GaussPy[o];
            }
            delete[] GaussPy;
        }
        if (data) {
            for (int i = 0; i < rows_; ++i) delete[] data[i];
            delete[] data;
        }
        if (mirror_) sspyr_host_free(mirror_);
        if (staging_) sspyr_host_free(staging_);
        if (h_) sspyr_destroy(h_);
    }

    float**** GaussPy;
    bool initialized;

    // ---- additions (not in the reference) ----
    int octaves() const { return layer; }
    int side(int o) const { return length >> o; }
    int rows(int o) const { return rows_ >> o; }
    int cols(int o) const { return cols_ >> o; }
    float* pixels() { return float_pixels_ ? static_cast<float*>(staging_) : nullptr; }   // superset ctor: the image copy
    float last_device_ms() { float ms = 0; check(sspyr_elapsed_ms(h_, &ms), "sspyr_elapsed_ms"); return ms; }
    void set_download(bool on) { download_ = on; }   // false: results stay on the device (kernel-only timing)
    sspyr_handle handle() const { return h_; }

    GaussPyramid_cuda(const GaussPyramid_cuda&) = delete;
    GaussPyramid_cuda& operator=(const GaussPyramid_cuda&) = delete;

protected:
    int length;
    int S;
    int layer;
    float* filter;  // kept for layout familiarity; the window tables live on the device

private:
    enum State { NOTHING, INIT, FILTERED };   // what the Gaussian-level planes of slot 0 hold on the device
    void check(int rc, const char* what) {
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + sspyr_last_error(h_));
    }
    // handle + pinned host mirror in the reference's dense in-place order + its float**** row tables (:55, :62-72)
    void setup(int octaves, float sigma0) {
        sspyr_config cfg;
        sspyr_default_config(&cfg);
        cfg.height = rows_;
        cfg.width = cols_;
        cfg.S = S;
        cfg.octaves = octaves;                                  // 0 = all: floor(log2 len)+1, :48-53
        cfg.outputs = SSPYR_OUT_ALL;
        cfg.mode = mode_;
        cfg.sigma0 = sigma0;
        cfg.pixel_type = float_pixels_ ? SSPYR_PIXEL_F32 : SSPYR_PIXEL_I32;
        check(sspyr_create(&cfg, &h_), "sspyr_create");
        check(sspyr_set_tuning(h_, "timing", 1), "sspyr_set_tuning");   // last_device_ms() for the driver's report
        layer = sspyr_num_octaves(h_);
        size_t floats = 0;
        for (int o = 0; o < layer; ++o) floats += (size_t)(S + 3) * rows(o) * cols(o);
        check(sspyr_host_alloc(sizeof(float) * (floats ? floats : 1), (void**)&mirror_), "sspyr_host_alloc");
        check(sspyr_host_alloc(4 * (size_t)rows_ * cols_, &staging_), "sspyr_host_alloc");
        GaussPy = new float***[layer]();                        // :55 (value-initialised)
        float* p = mirror_;
        for (int o = 0; o < layer; ++o) {
            GaussPy[o] = new float**[S + 3]();
            for (int s = 0; s < S + 3; ++s) {
                GaussPy[o][s] = new float*[rows(o)];
                for (int r = 0; r < rows(o); ++r, p += cols(o)) GaussPy[o][s][r] = p;
            }
        }
    }
    sspyr_handle h_;
    float* mirror_;
    void* staging_;   // pinned: int pixels (reference constructor) or float pixels (superset constructor)
    bool download_;
    int rows_, cols_;
    bool float_pixels_;
    int mode_;
    State state_;
};

#endif  // SIFT_GUASS_NORMAL_GAUSSDEPYRAMID_CUDA_H
