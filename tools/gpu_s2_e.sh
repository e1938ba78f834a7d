#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_workload.py --edm > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_workload.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; grep -E "ok|SUMMARY" gpurun_out/r02_sanitizer_racecheck.txt | tail -4; grep -E "Race reported|access at" gpurun_out/r02_sanitizer_racecheck.txt | sed 's/.*access at //' | sort | uniq -c | sort -rn | head -8
