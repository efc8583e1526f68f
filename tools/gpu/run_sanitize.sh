# plain run first (so that a failure of the workload itself is told apart from a sanitizer finding), then the tools
python tools/sanitize_workload.py > gpurun_out/r02_sanitize_plain.log 2>&1; echo "plain rc=$?"
tail -3 gpurun_out/r02_sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_sanitize_memcheck.txt python tools/sanitize_workload.py > gpurun_out/r02_sanitize_memcheck.out 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r02_sanitize_racecheck.txt python tools/sanitize_workload.py > gpurun_out/r02_sanitize_racecheck.out 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r02_sanitize_racecheck.txt
for tool in initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r02_sanitize_$tool.txt python tools/sanitize_workload.py > gpurun_out/r02_sanitize_$tool.out 2>&1; echo "$tool rc=$?"
  tail -3 gpurun_out/r02_sanitize_$tool.txt
done
