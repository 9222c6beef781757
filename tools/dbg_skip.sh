#!/bin/bash
# Phase timing of the posterior kernel: build the library with phase switches compiled in
#   make -C barefoot_framework_b200/csrc clean && make -C barefoot_framework_b200/csrc NVCCFLAGS_EXTRA=-DBF_PHASE_PROFILE
# then time the kernel with phases removed (results are wrong in these runs; rebuild without the flag afterwards).
for m in 0 1 2 4 3 6; do BF_DEBUG_SKIP=$m timeout 100 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$m kernel_ms', d['roofline']['kernel_ms'])"; done
