set -x
python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 && tail -2 gpurun_out/sanitize_plain.log &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|sanitize case done|error" gpurun_out/sanitize_memcheck.log | head -20
