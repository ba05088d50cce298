#!/bin/bash
# compute-sanitizer over the product path: memcheck and racecheck on smoke() and on a small slice of the GPU parity tests.
mkdir -p gpurun_out
echo "== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_memcheck_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_memcheck_smoke.log
echo "== racecheck smoke"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_racecheck_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_racecheck_smoke.log
echo "== memcheck tests (batch, fallbacks, cloud2, multi-stream)"; timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -m gpu -k "batch_equals_single or fallback_paths or pointcloud2 or multi_stream or edge_cases or radius_ties or speculation" > gpurun_out/san_memcheck_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/san_memcheck_tests.log
echo "== initcheck smoke"; timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_initcheck_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/san_initcheck_smoke.log
