#!/bin/bash
# Phase timing of the posterior kernel.  Build a copy of the library with the phase switches compiled in
# (the shipped library has none):
#   cd barefoot_framework_b200/csrc && make            # the other objects
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC \
#        -DBF_PHASE_PROFILE -c bf_posterior.cu -o /tmp/bf_posterior_prof.o
#   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/prof_lib/libbarefoot_sm100.so \
#        /tmp/bf_posterior_prof.o $(ls build/*.o | grep -v bf_posterior.o)
# then, on the GPU box, put it in place of the shipped library and run this script:
#   cp tools/prof_lib/libbarefoot_sm100.so barefoot_framework_b200/lib/libbarefoot_sm100.so; bash tools/dbg_skip.sh
# BF_DEBUG_SKIP bits: 1 = no phase B, 2 = no phase A, 4 = no phase C (results are wrong in those runs);
# 64 = per-warp clock64 accounting, printed by one CTA per launch (lines starting with "prof": cycles spent
# waiting at the loop-top barrier, in phase C, in phase A, waiting at the barrier before phase B, in phase B).
for m in 0 1 2 4 3 6; do BF_DEBUG_SKIP=$m timeout 100 python bench.py --configs 2 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$m kernel_ms', d['roofline']['kernel_ms'])"; done
BF_DEBUG_SKIP=64 timeout 100 python bench.py --configs 2 --steps 5 --warmup 3 2>/dev/null | grep '^prof' | tail -16
