// examples/pgm_pyramid.cpp -- real image ingest through the C ABI: read an 8-bit binary PGM (P5), build the
// scale space with 1-byte-per-pixel input (SSPYR_PIXEL_U8: 4x less input traffic than the reference's int**),
// write chosen levels back as PGM.  The reference only ever feeds an all-ones int** (main.cpp:31-35).
//
//   g++ -O2 -std=gnu++14 -Iinclude examples/pgm_pyramid.cpp -Lsift-parallel-optimization_b200 -lsspyr \
//       -Wl,-rpath,$PWD/sift-parallel-optimization_b200 -o build/pgm_pyramid
//   build/pgm_pyramid in.pgm out_prefix [ref|conv] [octaves]
//       -> out_prefix_o<octave>_g<level>.pgm (Gaussian levels, clamped to 0..255)
//          out_prefix_o<octave>_d<level>.pgm (DoG levels, 128 + 4*value, clamped)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sspyr.h"

static bool read_pgm(const char* path, std::vector<unsigned char>& px, int& w, int& h) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    char magic[3] = {0};
    int maxv = 0;
    auto skip = [&]() {                       // whitespace and # comments between header tokens
        int c;
        while ((c = std::fgetc(f)) != EOF) {
            if (c == '#') { while ((c = std::fgetc(f)) != EOF && c != '\n') {} }
            else if (c != ' ' && c != '\t' && c != '\n' && c != '\r') { std::ungetc(c, f); break; }
        }
    };
    bool ok = std::fscanf(f, "%2s", magic) == 1 && !std::strcmp(magic, "P5");
    if (ok) { skip(); ok = std::fscanf(f, "%d", &w) == 1; }
    if (ok) { skip(); ok = std::fscanf(f, "%d", &h) == 1; }
    if (ok) { skip(); ok = std::fscanf(f, "%d", &maxv) == 1 && maxv > 0 && maxv < 256; }
    if (ok) {
        std::fgetc(f);                        // the single whitespace after maxval
        px.resize((size_t)w * h);
        ok = w > 0 && h > 0 && std::fread(px.data(), 1, px.size(), f) == px.size();
    }
    std::fclose(f);
    return ok;
}

static bool write_pgm(const std::string& path, const std::vector<unsigned char>& px, int w, int h) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P5\n%d %d\n255\n", w, h);
    const bool ok = std::fwrite(px.data(), 1, px.size(), f) == px.size();
    std::fclose(f);
    return ok;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s in.pgm out_prefix [ref|conv] [octaves]\n", argv[0]); return 2; }
    std::vector<unsigned char> img;
    int w = 0, h = 0;
    if (!read_pgm(argv[1], img, w, h)) { std::fprintf(stderr, "cannot read %s as a binary PGM\n", argv[1]); return 2; }
    sspyr_config cfg;
    sspyr_default_config(&cfg);
    cfg.height = h;
    cfg.width = w;
    cfg.pixel_type = SSPYR_PIXEL_U8;
    cfg.mode = (argc > 3 && !std::strcmp(argv[3], "conv")) ? SSPYR_MODE_CONV : SSPYR_MODE_REF;
    cfg.octaves = argc > 4 ? std::atoi(argv[4]) : 0;
    sspyr_handle hd = nullptr;
    if (sspyr_create(&cfg, &hd) != SSPYR_OK) { std::fprintf(stderr, "sspyr_create: %s\n", sspyr_last_error(nullptr)); return 1; }
    auto check = [&](int rc, const char* what) {
        if (rc < 0) { std::fprintf(stderr, "%s: %s\n", what, sspyr_last_error(hd)); std::exit(1); }
    };
    check(sspyr_set_tuning(hd, "timing", 1), "sspyr_set_tuning");
    check(sspyr_upload(hd, 0, img.data(), 0), "sspyr_upload");
    check(sspyr_build(hd, 0), "sspyr_build");
    float ms = 0;
    check(sspyr_elapsed_ms(hd, &ms), "sspyr_elapsed_ms");
    const int octs = sspyr_num_octaves(hd), nl = sspyr_num_levels(hd);
    std::printf("%dx%d, %d octaves x %d levels, %s mode: %.3f ms on the GPU\n", w, h, octs, nl,
                cfg.mode == SSPYR_MODE_CONV ? "CONV" : "REF", ms);
    std::vector<float> plane;
    std::vector<unsigned char> out;
    for (int o = 0; o < octs; ++o) {
        int r = 0, c = 0;
        check(sspyr_level_dims(hd, o, &r, &c, nullptr), "sspyr_level_dims");
        plane.resize((size_t)r * c);
        out.resize(plane.size());
        for (int s : {0, nl - 1}) {
            check(sspyr_download(hd, 0, o, s, SSPYR_KIND_GAUSS, plane.data(), 0), "sspyr_download");
            for (size_t i = 0; i < plane.size(); ++i) out[i] = (unsigned char)std::min(255.0f, std::max(0.0f, plane[i] + 0.5f));
            write_pgm(std::string(argv[2]) + "_o" + std::to_string(o) + "_g" + std::to_string(s) + ".pgm", out, c, r);
        }
        check(sspyr_download(hd, 0, o, 1, SSPYR_KIND_DOG, plane.data(), 0), "sspyr_download");
        for (size_t i = 0; i < plane.size(); ++i) out[i] = (unsigned char)std::min(255.0f, std::max(0.0f, 128.0f + 4.0f * plane[i]));
        write_pgm(std::string(argv[2]) + "_o" + std::to_string(o) + "_d1.pgm", out, c, r);
    }
    sspyr_destroy(hd);
    return 0;
}
