#!/bin/bash
# Developer tool (GPU box): time every variant built by tools/build_attn_variants.sh on the c3 self-attention shape.
cd "$(dirname "$0")/.."
cp video_styler_b200/libwvd.so /tmp/libwvd_orig.so
for f in video_styler_b200/variants/libwvd_*.so; do
  n=$(basename $f .so); n=${n#libwvd_}
  cp $f video_styler_b200/libwvd.so
  for e in ${EMUS:-0}; do
    r=$(WVD_ATTN_EMU=$e timeout 120 python tools/kernel_check.py --only attn --big 2>&1 | grep -E "time attn 29640x29640|29640x29640 h40 vs" | tr '\n' ' ')
    echo "$n emu=$e: $r"
  done
done
cp /tmp/libwvd_orig.so video_styler_b200/libwvd.so
