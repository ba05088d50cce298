mkdir -p gpurun_out
echo "== memcheck variants"; timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/san_variants.py > gpurun_out/san_memcheck_variants.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/san_memcheck_variants.log
echo "== racecheck variants"; timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/san_variants.py > gpurun_out/san_racecheck_variants.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/san_racecheck_variants.log
echo "== memcheck smoke"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_memcheck_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_memcheck_smoke.log
echo "== memcheck C2 packed"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/san_c2_packed.py > gpurun_out/san_memcheck_c2_packed.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_memcheck_c2_packed.log
echo "== racecheck C2 packed"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/san_c2_packed.py > gpurun_out/san_racecheck_c2_packed.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_racecheck_c2_packed.log
echo "== initcheck smoke"; timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_initcheck_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_initcheck_smoke.log
