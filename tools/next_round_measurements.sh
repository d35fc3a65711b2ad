#!/bin/bash
# First GPU call of the next round: one `gpurun -- bash tools/next_round_measurements.sh`, ~2 min.
#  - EVOPF projection with / without a flag buffer, and the RPO_FLAG_MASK=1 build of the EVOPF kernels
#  - cart / pendulum K-step kernels after the flag rewrite
#  - the default bench line (refreshes other_configs)
set -x
mkdir -p build/variants gpurun_out
F='-O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DRPO_BUILD_ARCH="sm_100a"'
( cd rpo_b200/csrc && nvcc $F -shared -o ../../build/variants/a_base.so rpo_kernels.cu -lcudart \
  && nvcc $F -DRPO_FLAG_MASK=1 -shared -o ../../build/variants/b_flagmask.so rpo_kernels.cu -lcudart )
timeout 300 python tools/variant_sweep.py | tee gpurun_out/next_variants.txt
timeout 120 python tools/quick_gpu.py | tee gpurun_out/next_quick.txt
timeout 120 python tools/classic_flags_gpu.py | tee gpurun_out/next_classic.txt
timeout 300 python bench.py > gpurun_out/next_bench.json 2> gpurun_out/next_bench.err
cut -c1-400 gpurun_out/next_bench.json
