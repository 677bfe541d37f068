#!/bin/bash
# Round-2 starter for the rows past the step (light map, ray marches, volume files), ONE gpurun call (~3 min):
#   gpurun --timeout 600 -- 'bash tools/profile_next_rows.sh'
# 1. their GPU tests (first execution on a B200) and the 512^3 cross-path test; 2. per-launch durations and one
# `ncu --set full` capture of the light-map and ray-march kernels on a developed 256^3 plume.  Output: gpurun_out/next_*.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzx_gpu_full_size.py tests/test_zzy_gpu_volume.py tests/test_zzz_gpu_lightmap.py \
    tests/test_zzz_gpu_raymarch.py -m gpu -q > gpurun_out/next_tests.log 2>&1; tail -5 gpurun_out/next_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/next_launches.csv \
    -k regex:"extract_density|light_map_kernel|ray_march" python tools/profile_next_rows.py 256 > gpurun_out/next_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"light_map_kernel|ray_march" -c 3 -o gpurun_out/next_rows -f \
    python tools/profile_next_rows.py 256 > /dev/null 2>&1
tail -3 gpurun_out/next_run.log
