mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_render.py > gpurun_out/r02_memcheck_render.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py > gpurun_out/r02_memcheck_raster.log 2>&1
timeout 800 compute-sanitizer --tool racecheck python tools/sanitize_case.py > gpurun_out/r02_racecheck_raster.log 2>&1
timeout 800 compute-sanitizer --tool racecheck python tools/sanitize_render.py > gpurun_out/r02_racecheck_render.log 2>&1
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY\|sanitize_render ok\|Error" gpurun_out/r02_memcheck_*.log gpurun_out/r02_racecheck_*.log | head -20
