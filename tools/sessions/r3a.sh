#!/bin/bash
# Round 2, session 3a: compute-sanitizer (racecheck, memcheck, synccheck) on the warp-specialised kernels.
set -u
O=gpurun_out
mkdir -p $O
for tool in racecheck synccheck memcheck; do
  echo "== $tool stream"
  timeout -s KILL 600 compute-sanitizer --tool $tool python tools/sanitize_stream.py stream 2>&1 | tail -8
done | tee $O/r3a_sanitizer_stream.log
for tool in racecheck synccheck; do
  echo "== $tool fused"
  timeout -s KILL 900 compute-sanitizer --tool $tool python tools/sanitize_stream.py fused 2>&1 | tail -8
done | tee $O/r3a_sanitizer_fused.log
