#!/bin/bash
# tools/sanitize.sh [outdir] -- the GPU parity tests of the tracker under compute-sanitizer (memcheck, then initcheck),
# with the product library as the only instrumented code of interest.  The persistent tracker kernel spins on relaxed
# global loads; under the sanitizer it runs 10-50x slower but must still terminate, read no uninitialised device memory
# and touch nothing outside its arena.  Logs: <outdir>/sanitize_memcheck.log, <outdir>/sanitize_initcheck.log.
# Small images keep the run within minutes.  Run on the GPU box:  bash tools/sanitize.sh gpurun_out
out=${1:-gpurun_out}
mkdir -p "$out"
cd "$(dirname "$0")/.."
SEL='tests/test_tracker_edge_gpu.py::test_ragged_sizes_match_reference tests/test_tracker_edge_gpu.py::test_empty_depth_and_black_image_do_not_hang_and_keep_the_pose tests/test_tracker_edge_gpu.py::test_degenerate_geometry_icp_only tests/test_array_entry_gpu.py::test_same_entry_twice_keeps_both_calls_apart'
rc=0
for tool in memcheck initcheck; do
  timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 600 \
      --kernel-name kns=2ef --log-file "$out/sanitize_$tool.log" \
      python -m pytest $SEL -x -q -m gpu -k "${SANITIZE_K:-168 or 96 or 321 or empty or plane or twice}" > "$out/sanitize_$tool.pytest.log" 2>&1
  code=$?
  echo "[$tool] exit code $code" | tee -a "$out/sanitize_$tool.log"
  tail -3 "$out/sanitize_$tool.pytest.log" | tee -a "$out/sanitize_$tool.log"
  [ $code -ne 0 ] && rc=$code
done
exit $rc
