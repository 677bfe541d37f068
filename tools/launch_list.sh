#!/bin/bash
# Per-launch device times of two steps (after 100 spin-up steps) at the given grids: gpurun -- 'bash tools/launch_list.sh tag 256 512'
tag=$1; shift
out=gpurun_out; mkdir -p $out
for g in "$@"; do
  K=$(python tools/profile_step.py --grid $g $g $g --steps 1 | sed -n 's/^kernels\/step \([0-9]*\).*/\1/p')
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((100*K)) -c $((2*K)) --csv --log-file $out/${tag}_launches_$g.csv \
      python tools/profile_step.py --grid $g $g $g --steps 102 > $out/${tag}_launches_$g.log 2>&1
  python tools/summarize_profiles.py launches $out/${tag}_launches_$g.csv $out/${tag}_launches_$g.txt > /dev/null
  sed -n 1,$((K+14))p $out/${tag}_launches_$g.txt
done
