mkdir -p gpurun_out
timeout 300 python tools/sanitize_case.py 32 > gpurun_out/r2j_sanitize_plain.log 2>&1; echo rc=$? >> gpurun_out/r2j_sanitize_plain.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_case.py 32 > gpurun_out/r2j_racecheck.log 2>&1; echo rc=$? >> gpurun_out/r2j_racecheck.log
tail -5 gpurun_out/r2j_sanitize_plain.log; tail -15 gpurun_out/r2j_racecheck.log
