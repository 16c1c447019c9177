#!/bin/bash
# Developer tool: build libwvd variants that differ only in attention compile-time knobs.
#   tools/build_attn_variants.sh name1 "-DFOO=1 -DBAR=2" name2 "..." ...
# Writes video_styler_b200/variants/libwvd_<name>.so (git-ignored; they travel to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/.."
python -m video_styler_b200.build > /dev/null
mkdir -p video_styler_b200/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
      -c video_styler_b200/csrc/attention_sm100.cu -o video_styler_b200/variants/attention_$name.o &&
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
      -c video_styler_b200/csrc/attention_pair_sm100.cu -o video_styler_b200/variants/attention_pair_$name.o &&
    nvcc -shared -o video_styler_b200/variants/libwvd_$name.so video_styler_b200/build/api.o video_styler_b200/build/elementwise.o \
      video_styler_b200/build/gemm_sm100.o video_styler_b200/variants/attention_$name.o video_styler_b200/variants/attention_pair_$name.o video_styler_b200/build/fallthrough_f32.o \
      -gencode arch=compute_100a,code=sm_100a -cudart static && echo built $name ) &
done
wait
