#!/bin/bash
# tools/sanitize.sh [outdir] -- GPU tests of the tracker under compute-sanitizer, memcheck then initcheck.
# The persistent tracker kernel spins on relaxed global loads; under the sanitizer it runs 10-50x slower but must still
# terminate, touch nothing outside its buffers and read no uninitialised device memory.
#   memcheck : product AND reference kernels (parity tests at small sizes, degenerate geometry, cudaArray entries)
#   initcheck: product-only tests -- the reference library itself trips initcheck (DeviceMemory::download copies its
#              result struct's never-written tail, 228 reports in round 2), which would bury anything of ours
# Logs: <outdir>/sanitize_memcheck.log, <outdir>/sanitize_initcheck.log (+ .pytest.log).  Run on the GPU box.
out=${1:-gpurun_out}
mkdir -p "$out"
cd "$(dirname "$0")/.."
E=tests/test_tracker_edge_gpu.py
MEM="$E::test_ragged_sizes_match_reference $E::test_empty_depth_and_black_image_do_not_hang_and_keep_the_pose $E::test_degenerate_geometry_icp_only tests/test_array_entry_gpu.py::test_same_entry_twice_keeps_both_calls_apart tests/test_batch_gpu.py $E::test_fused_host_iteration_matches_the_operator_path"
MEMK="168 or 96 or 321 or empty or plane or twice or k3 or 322x242 or different_iterations or size1"
INIT="$E::test_empty_depth_and_black_image_do_not_hang_and_keep_the_pose $E::test_sparse_depth_matches_host_mode $E::test_sm_subsets_give_the_same_pose $E::test_unaligned_device_inputs_take_the_chained_builders $E::test_deferred_build_through_the_reference_calls tests/test_array_entry_gpu.py::test_same_entry_twice_keeps_both_calls_apart tests/test_batch_gpu.py $E::test_fused_host_iteration_matches_the_operator_path tests/test_predict.py"
INITK="not size0 and not size2 and not 640x480"
rc=0
run() { # tool, tests, -k expression
  timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $1 --error-exitcode 86 --launch-timeout 600 --print-limit 400 \
      --log-file "$out/sanitize_$1.log" python -m pytest $2 -q -m gpu -k "$3" > "$out/sanitize_$1.pytest.log" 2>&1
  code=$?
  echo "[$1] exit code $code" | tee -a "$out/sanitize_$1.log"
  tail -3 "$out/sanitize_$1.pytest.log" | tee -a "$out/sanitize_$1.log"
  [ $code -ne 0 ] && rc=$code
}
run memcheck "$MEM" "$MEMK"
run initcheck "$INIT" "$INITK"
# shared-memory hazards and barrier misuse of the persistent kernel (named barriers per thread group in the batched build)
SMALL="$E::test_empty_depth_and_black_image_do_not_hang_and_keep_the_pose $E::test_sparse_depth_matches_host_mode $E::test_fused_host_iteration_matches_the_operator_path tests/test_batch_gpu.py"
run racecheck "$SMALL" "not size0 and not size2 and not 640x480 and not argument"
run synccheck "$SMALL" "not size0 and not size2 and not 640x480 and not argument"
exit $rc
