#!/bin/bash
# Developer A/B probe: per-stage times of several builds of the library on one GPU box.
#   tools/ab_probe.sh out.jsonl tag1 tag2 ...     (tag "main" = the in-tree libsiftb200.so, else sift-gpu_b200/libsiftb200_<tag>.so)
out=$1; shift
: > "$out"
for t in "$@"; do
  if [ "$t" = main ]; then unset SIFT_B200_LIB; else export SIFT_B200_LIB=$PWD/sift-gpu_b200/libsiftb200_$t.so; fi
  python tools/stage_probe.py "$t" >> "$out" 2>> "${out%.jsonl}.err"
done
cat "$out"
