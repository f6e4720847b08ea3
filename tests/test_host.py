"""CPU suite, part 2: host logic and the C-ABI surface (no compute calls without a GPU)."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_symbols() -> list[str]:
    text = open(os.path.join(ROOT, "include", "sspyr.h")).read()
    return re.findall(r"^SSPYR_API\s+[\w\s\*]+?\b(sspyr_\w+)\s*\(", text, flags=re.M)


def test_library_exports_every_declared_symbol(pkg):
    declared = _declared_symbols()
    assert len(declared) >= 25 and len(set(declared)) == len(declared)
    assert set(declared) == set(pkg._lib.SYMBOLS)                    # the ctypes table tracks the header
    lib = C.CDLL(pkg._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libsspyr.so does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {e for e in exported if e.startswith("sspyr_")} == set(declared)
    assert not [e for e in exported if not e.startswith("sspyr_") and not e.startswith("_")], exported


def test_config_struct_matches_the_header(pkg):
    """sizeof/offsetof as the C compiler sees them == the ctypes mirror."""
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "sspyr.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(sspyr_config), offsetof(sspyr_config, sigma0),
        offsetof(sspyr_config, band_row0), offsetof(sspyr_config, extrema_thresh), offsetof(sspyr_config, reserved)); return 0; }
    '''
    exe = "/tmp/sspyr_cfg_probe"
    subprocess.run(["/usr/bin/gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe],
                   input=src, text=True, check=True)
    got = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    Cfg = pkg._lib.Config
    assert got == [C.sizeof(Cfg), Cfg.sigma0.offset, Cfg.band_row0.offset, Cfg.extrema_thresh.offset,
                   Cfg.reserved.offset]


def test_default_config_and_version(pkg):
    lib = pkg._lib.load()
    assert lib.sspyr_version() == 100
    cfg = pkg._lib.Config()
    assert lib.sspyr_default_config(C.byref(cfg)) == 0
    assert (cfg.S, cfg.mode, cfg.outputs, cfg.pixel_type, cfg.frames, cfg.device) == (3, 0, 3, 0, 1, -1)
    assert lib.sspyr_default_config(None) == pkg._lib.ERR_ARG


def test_create_rejects_bad_configs_before_touching_cuda(pkg):
    lib = pkg._lib.load()
    h = C.c_void_p()

    def rc(**kw):
        cfg = pkg._lib.Config()
        lib.sspyr_default_config(C.byref(cfg))
        cfg.height, cfg.width = 64, 64
        for k, v in kw.items():
            setattr(cfg, k, v)
        return lib.sspyr_create(C.byref(cfg), C.byref(h))

    assert rc(height=0) == pkg._lib.ERR_ARG
    assert rc(S=-1) == pkg._lib.ERR_ARG
    assert rc(S=14) == pkg._lib.ERR_ARG                   # S + 3 levels <= SSPYR_MAX_LEVELS = 16
    assert rc(mode=7) == pkg._lib.ERR_ARG
    assert rc(octaves=8) == pkg._lib.ERR_ARG              # 64 -> at most 7 octaves (GuassDePyramid.h:48-53)
    assert rc(full_height=128, band_row0=8, octaves=5) == pkg._lib.ERR_ARG   # band not aligned to 2^(O-1)
    assert rc(outputs=8) == pkg._lib.ERR_ARG              # EXTREMA without DOG
    assert b"aligned" in lib.sspyr_last_error(None) or lib.sspyr_last_error(None)


def test_no_cpu_fallback(pkg):
    """Without a GPU the product refuses to run (it must never route through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SspyrError) as e:
        pkg.ScaleSpace(64, 64)
    assert e.value.code == pkg._lib.ERR_CUDA and "no CPU fallback" in str(e.value)
    src = "".join(open(os.path.join(ROOT, "sift-parallel-optimization_b200", f)).read()
                  for f in ("__init__.py", "_lib.py", "pyramid.py", "partition.py", "synth.py"))
    assert "oracle" not in src.replace("the oracle", "")    # the product package never imports oracle/


def test_shard_frames(pkg):
    for n, world in ((256, 8), (10, 4), (3, 8), (0, 2)):
        spans = [pkg.shard_frames(n, world, r) for r in range(world)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for first, count in spans:
            assert first == pos and count in (n // world, n // world + 1)
            pos += count
    with pytest.raises(ValueError):
        pkg.shard_frames(4, 2, 2)


@pytest.mark.parametrize("h,octs,world", [(4320, 5, 8), (16384, 8, 8), (1080, 5, 4), (2160, 5, 2), (100, 3, 8),
                                          (48, 5, 8), (4320, 5, 1)])
def test_band_rows(pkg, h, octs, world):
    align = 1 << (octs - 1)
    spans = [pkg.band_rows(h, octs, world, r) for r in range(world)]
    pos = 0
    for row0, rows in spans:
        if rows == 0:
            assert row0 == h
            continue
        assert row0 == pos and row0 % align == 0                # decimation r<<o stays band-local
        pos += rows
        if pos != h:
            assert rows % align == 0
        for o in range(octs):                                    # band octave rows tile the full octave rows
            assert (row0 >> o) + (rows >> o) == (pos >> o)
    assert pos == h
    live = [r for _, r in spans if r]
    assert max(live) - min(live) <= align + h % align


def test_reference_side_binding_compiles_as_cxx14(pkg, tmp_path):
    """include/GaussDePyramid-CUDA.h is what a maintainer drops next to the reference's variant headers: it must
    compile as C++14 with plain g++ (no CUDA headers), alone and next to the reference's own header, and link
    against libsspyr.so only."""
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "probe.cpp"
    ref = "/root/reference/GuassDePyramid.h"
    src.write_text(('#include "GuassDePyramid.h"\n' if os.path.exists(ref) else "") + """
#include "GaussDePyramid-CUDA.h"
int main() {
    int row[4] = {1, 2, 3, 4};
    int* img[4] = {row, row, row, row};
    try { GaussPyramid_cuda g(img, 4, 2); g.GenerateDoG(); g.GenerateDoG_nomp_dynamic(); return g.GaussPy[0][0][0][0] != 0; }
    catch (const std::exception&) { return 42; }     // no GPU here: sspyr_create reports it, the class throws
}
""")
    cmd = ["/usr/bin/g++", "-std=gnu++14", "-Wall", "-I", inc, str(src), "-L", os.path.dirname(pkg._lib.LIB_PATH), "-lsspyr",
           "-Wl,-rpath," + os.path.dirname(pkg._lib.LIB_PATH), "-o", str(tmp_path / "probe")]
    if os.path.exists(ref):
        cmd[1:1] = ["-I/root/reference"]
    subprocess.run(cmd, check=True)
    import torch
    rc = subprocess.run([str(tmp_path / "probe")]).returncode
    assert rc == (0 if torch.cuda.is_available() else 42)


def test_conv_scheduling_arithmetic(tmp_path):
    """csrc/conv_sched.h (host-only C++): how CONV levels are cut into (strip, segment) CTAs, which levels are
    chained, how many builds are in flight.  Compiled with plain g++ and checked on BASELINE.json's shapes and on
    random geometries: segments are whole 32-row steps that cover the level, identical for every level of an
    octave (the chaining counters are indexed by them), long where the level is large, and a level of less than a
    wave is never chained automatically."""
    src = tmp_path / "sched.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include "conv_sched.h"
using namespace sspyr;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)
int main() {
    const int sms = 148;
    // BASELINE shapes, octave 0, automatic segmentation (waves = 0), minimum 32 rows
    CHECK(march_seg_rows(4320, 7680, 1, sms, 0, 32) == 256);            // 8K: 8-step segments, 60 x 17 CTAs = 1.7 waves
    CHECK(march_ctas(4320, 7680, 1, 256) == 1020);
    CHECK(march_seg_rows(16384, 16384, 1, sms, 0, 32) == 1280);         // 16K: 3 waves already give 40-step segments
    CHECK(march_seg_rows(2160, 3840, 8, sms, 0, 32) == 320);            // 4K x 8 frames per launch
    CHECK(march_seg_rows(1080, 1920, 1, sms, 0, 32) == 32);             // 1080p: one step per CTA (less than a wave)
    CHECK(march_seg_rows(4320, 7680, 1, sms, 3, 32) == 160);            // explicit wave counts are honoured
    CHECK(march_seg_rows(4320, 7680, 1, sms, 3, 512) == 512);
    // chaining: only multi-wave grids unless forced
    CHECK(level_chained(1, march_ctas(4320, 7680, 1, 256), sms));
    CHECK(level_chained(1, march_ctas(2160, 3840, 1, march_seg_rows(2160, 3840, 1, sms, 0, 32)), sms));   // 8K octave 1
    CHECK(!level_chained(1, march_ctas(1080, 1920, 1, 32), sms));
    CHECK(!level_chained(0, 1 << 20, sms) && level_chained(2, 1, sms));
    // lanes
    CHECK(frame_lanes(8, 8, false) == 8 && frame_lanes(8, 5, false) == 5 && frame_lanes(8, 1, false) == 1);
    CHECK(frame_lanes(8, 8, true) == 3 && frame_lanes(2, 8, true) == 2 && frame_lanes(0, 8, false) == 1);
    CHECK(frame_lanes(64, 64, false) == 16);
    // random geometries
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    auto rnd = [&](int lo, int hi) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return lo + (int)(x % (unsigned long long)(hi - lo + 1)); };
    for (int it = 0; it < 20000; ++it) {
        const int H = rnd(1, 20000), W = rnd(1, 20000), frames = rnd(1, 9), waves = rnd(0, 6), seg_min = 32 * rnd(1, 4);
        const int r = march_seg_rows(H, W, frames, sms, waves, seg_min);
        CHECK(r % STRIP_TH == 0 && r >= seg_min);
        const long long ctas = march_ctas(H, W, frames, r);
        CHECK(ctas >= 1 && (long long)((H + r - 1) / r) * r >= H);
        if (waves == 0 && r == 8 * STRIP_TH && r > seg_min)              // long segments were chosen: still >= 1.5 waves
            CHECK(2 * ctas >= 3LL * sms * STRIP_CTAS_PER_SM || march_seg_rows(H, W, frames, sms, 3, seg_min) >= r);
        if (waves > 0 && r > seg_min)                                    // never more CTAs than the waves asked for (rounding up rows)
            CHECK(ctas <= (long long)sms * STRIP_CTAS_PER_SM * waves || (H + r - 1) / r == 1);
    }
    std::printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "sched"
    subprocess.run(["/usr/bin/g++", "-std=gnu++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "sift-parallel-optimization_b200", "csrc"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout


def test_cascade_block_order_is_dependency_safe(tmp_path):
    """csrc/conv_sched.h cascade_item_table: the block order of the one-launch CONV build.  Deadlock freedom rests on
    every work item coming after everything it reads (CTAs are dispatched in block order and spin on their producers'
    counters), so the order is checked here, on the CPU, for BASELINE.json's shapes and for random geometries: the
    row-keyed order must pass the check on the BASELINE shapes (it is the fast one), and whatever the function returns
    must pass it always, hold every item exactly once and keep the strips of a group together."""
    src = tmp_path / "casc.cpp"
    src.write_text(r'''
#include <cstdio>
#include <set>
#include "conv_sched.h"
using namespace sspyr;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)
static int run(int H, int W, int octaves, int S, const int* rad, int seg_tuned, bool must_be_keyed) {
    const int nl = S + 3;
    CascItemGeom g[16];
    int radius[16];
    for (int s = 0; s < nl; ++s) radius[s] = cascade_radius_class(rad[s]);
    size_t want = 0;
    for (int o = 0; o < octaves; ++o) {
        const int h = H >> o, w = W >> o;
        const int sr = cascade_seg_rows(w, seg_tuned);
        g[o] = CascItemGeom{sr, (h + sr - 1) / sr, (w + CONV_TW - 1) / CONV_TW, o == 0 ? 0 : 1, h, w};
        want += (size_t)g[o].nsegs * g[o].nstrips * (nl - g[o].first_level);
    }
    bool keyed = false;
    const std::vector<unsigned> tab = cascade_item_table(g, octaves, nl, S, radius, &keyed);
    CHECK(tab.size() == want);
    CHECK(std::set<unsigned>(tab.begin(), tab.end()).size() == want);         // every item exactly once
    CHECK(cascade_order_is_safe(tab, g, octaves, nl, S, radius));
    if (must_be_keyed) CHECK(keyed);
    std::vector<unsigned> bad = tab;                                          // the checker does catch a broken order
    if (bad.size() > 1 && (bad.front() >> 24) != (bad.back() >> 24)) { std::swap(bad.front(), bad.back()); CHECK(!cascade_order_is_safe(bad, g, octaves, nl, S, radius)); }
    return 0;
}
int main() {
    const int def[6] = {5, 4, 5, 6, 8, 10};                                   // sigma0 1.6, S = 3, radius 3 sigma
    CHECK(cascade_seg_rows(7680, 0) == 128 && cascade_seg_rows(16384, 0) == 64 && cascade_seg_rows(3840, 0) == 256);
    CHECK(cascade_seg_rows(1920, 0) == 256 && cascade_seg_rows(100, 0) == 256 && cascade_seg_rows(7680, 70) == 96);
    CHECK(cascade_radius_class(1) == 6 && cascade_radius_class(6) == 6 && cascade_radius_class(7) == 10 && cascade_radius_class(12) == 12);
    if (run(1080, 1920, 5, 3, def, 0, true)) return 1;                        // C2
    if (run(2160, 3840, 5, 3, def, 0, true)) return 1;                        // C3
    if (run(4320, 7680, 5, 3, def, 0, true)) return 1;                        // C4
    if (run(16384, 16384, 8, 3, def, 0, true)) return 1;                      // C5
    if (run(512, 512, 4, 3, def, 0, true)) return 1;                          // C1
    for (int seg : {32, 64, 96, 160, 256}) if (run(4320, 7680, 5, 3, def, seg, true)) return 1;
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    auto rnd = [&](int lo, int hi) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return lo + (int)(x % (unsigned long long)(hi - lo + 1)); };
    for (int it = 0; it < 300; ++it) {
        const int H = rnd(1, 3000), W = rnd(1, 3000), S = rnd(1, 5);
        int octs = 1;
        while ((std::min(H, W) >> octs) >= 1 && octs < 8) ++octs;
        octs = rnd(1, octs);
        int rad[16];
        for (int s = 0; s < S + 3; ++s) rad[s] = rnd(1, 12);
        if (run(H, W, octs, S, rad, 32 * rnd(0, 8), false)) { std::printf("geometry %d x %d, %d octaves, S=%d\n", H, W, octs, S); return 1; }
    }
    std::printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "casc"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=gnu++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "sift-parallel-optimization_b200", "csrc"),
                    str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout
