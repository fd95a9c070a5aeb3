python tools/sanitize_small.py > gpurun_out/san_plain.log 2>&1 && tail -3 gpurun_out/san_plain.log && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|=========" gpurun_out/san_memcheck.log | head -20; tail -3 gpurun_out/san_memcheck.log
