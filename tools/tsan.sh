#!/bin/bash
# Builds nothing: runs the prebuilt TSAN host test (tools/tsan_host.c against pixelbox_b200/lib/exp/lib_tsan.so, the library
# compiled with -Xcompiler -fsanitize=thread).  Reports in gpurun_out/sanitizer/tsan_host.log.
mkdir -p gpurun_out/sanitizer
TSAN_OPTIONS="report_signal_unsafe=0 history_size=4 second_deadlock_stack=1" timeout 600 tools/bin/tsan_host > gpurun_out/sanitizer/tsan_host.log 2>&1
echo "tsan_host rc=$? warnings=$(grep -c 'WARNING: ThreadSanitizer' gpurun_out/sanitizer/tsan_host.log)"
tail -4 gpurun_out/sanitizer/tsan_host.log
grep -A12 "WARNING: ThreadSanitizer" gpurun_out/sanitizer/tsan_host.log | grep "#0\|#1\|#2\|WARNING" | head -30
