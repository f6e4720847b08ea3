#!/bin/bash
# scripts/gpu_bands8.sh -- one 8-GPU call: band check with chained seams, then A/B of seam protocol and builds in flight.
set -u
N=8 TEST=1 BCS="0" bash scripts/gpu_bands.sh || exit 1
mv gpurun_out/bands8/bench_bc0.json gpurun_out/bands8/bench_lanes3.json
N=8 TEST=0 BCS="1" bash scripts/gpu_bands.sh
N=8 TEST=0 BCS="0" XT=",conv_band_lanes=6" XA="--extras-slots 6" bash scripts/gpu_bands.sh; mv gpurun_out/bands8/bench_bc0.json gpurun_out/bands8/bench_lanes6.json
N=8 TEST=0 BCS="0" XT=",conv_band_lanes=8" XA="--extras-slots 8" bash scripts/gpu_bands.sh; mv gpurun_out/bands8/bench_bc0.json gpurun_out/bands8/bench_lanes8.json
