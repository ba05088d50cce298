mkdir -p gpurun_out
for tool in initcheck memcheck racecheck; do
echo "== $tool variants"; timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python scripts/san_variants.py > gpurun_out/san_${tool}_variants.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/san_${tool}_variants.log
done
echo "== initcheck smoke"; timeout 600 compute-sanitizer --tool initcheck --print-limit 20 --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_initcheck_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/san_initcheck_smoke.log
