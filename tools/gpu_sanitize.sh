#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_workload.py --edm > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.txt
timeout -s KILL 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_workload.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_racecheck.txt
timeout -s KILL 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_workload.py > gpurun_out/r02_sanitizer_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_synccheck.txt
