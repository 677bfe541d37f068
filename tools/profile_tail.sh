#!/bin/bash
# Round-2 starter: everything needed to see where a tail launch spends its time, in ONE gpurun call (~3 min).
#   gpurun --timeout 600 -- 'bash tools/profile_tail.sh'
# Writes into gpurun_out/: tail_shot.jsonl (parity + timing, tools/gpu_shot.py), tail_launches.csv (per-launch
# durations of two steps), tail_first.ncu-rep / tail_mid.ncu-rep (ncu --set full of the first and the sixth tail
# launch of a developed 256^3 step), tail_probe.txt (in-kernel phase cycles; needs the timing build, done here).
set -x
mkdir -p gpurun_out
# every opt-in variant against the oracle first (the gated tests: cp.async / TMA staging, second dense path, pass-0 kernel,
# second advection kernel), then parity + timing of the variants side by side
FXB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_gpu_tail.py tests/test_zy_gpu_golden.py -m gpu -q \
    > gpurun_out/tail_tests.log 2>&1; tail -3 gpurun_out/tail_tests.log
python tools/gpu_shot.py > gpurun_out/tail_shot.log 2>&1; cp gpurun_out/shot.jsonl gpurun_out/tail_shot.jsonl
export FXB_TAIL=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tail_launches.csv \
    -k regex:"jacobi_tail|jacobi_pass" -s 2400 -c 48 python tools/profile_step.py --grid 256 256 256 --steps 103 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_tail -s 1600 -c 1 -o gpurun_out/tail_first -f \
    python tools/profile_step.py --grid 256 256 256 --steps 101 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_tail -s 1605 -c 1 -o gpurun_out/tail_mid -f \
    python tools/profile_step.py --grid 256 256 256 --steps 101 > /dev/null 2>&1
make -C fluidx12_b200/csrc -B -j8 EXTRA=-DFXB_TAIL_TIMING > /dev/null 2>&1 && python tools/tail_probe.py 256 120 > gpurun_out/tail_probe.txt 2>&1
make -C fluidx12_b200/csrc -B -j8 > /dev/null 2>&1
