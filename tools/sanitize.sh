#!/usr/bin/env bash
# Race / memory checking of the CUDA kernels (reference analogue: `make DEBUG=1` ASan build, Makefile:8-12; Vulkan validation
# layers, nn-vulkan.cpp:23-30). Run on a GPU box:   gpurun --timeout 900 -- 'bash tools/sanitize.sh memcheck'
#   memcheck  : out-of-bounds / misaligned global + shared accesses
#   racecheck : shared-memory hazards between warps of a CTA (ring stages, activation planes, partial sums)
#   synccheck : illegal bar.sync / mbarrier usage
#   initcheck : reads of uninitialised global memory
# Inter-CTA protocols (grid barrier, LL all-reduce words) are outside racecheck's model; they are covered by the bit-exactness
# tests (tests/test_gpu_engine.py::test_graph_decode_equals_eager, tools/tp_check.py: TP=N == TP=1 token for token).
set -euo pipefail
tool="${1:-memcheck}"
shift || true
out="gpurun_out/sanitize_${tool}.log"
mkdir -p gpurun_out
compute-sanitizer --tool "$tool" --error-exitcode 3 --print-limit 20 python tools/sanitize_target.py "$@" > "$out" 2>&1 && rc=0 || rc=$?
tail -n 15 "$out"
echo "compute-sanitizer $tool exit code $rc (full log: $out)"
exit $rc
