#!/bin/bash
# One gpurun call of the build -> measure loop: GPU tests, smoke, the bench line, launch lists and ncu captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_shot.sh [tests|notests] [tag]'
# Everything lands in gpurun_out/ (scratch); tools/summarize_profiles.py turns it into profiles/.
what=${1:-tests}; tag=${2:-r2}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
if [ "$what" = tests ]; then
  timeout 1100 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
  timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
  tail -3 $out/${tag}_smoke.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python tools/benchsum.py $out/${tag}_bench.json 2>/dev/null | head -40
for g in 256 512; do
  K=$(python tools/profile_step.py --grid $g $g $g --steps 1 | sed -n 's/^kernels\/step \([0-9]*\).*/\1/p')
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((100*K)) -c $((2*K)) --csv --log-file $out/${tag}_launches_$g.csv \
      python tools/profile_step.py --grid $g $g $g --steps 102 > $out/${tag}_launches_$g.log 2>&1
done
# DRAM bytes of every launch of two steps -> profiles/traffic.json (bench.py's roofline.traffic)
for g in 256 512; do
  K=$(python tools/profile_step.py --grid $g $g $g --steps 1 | sed -n 's/^kernels\/step \([0-9]*\).*/\1/p')
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $((100*K)) -c $((2*K)) --csv --log-file $out/${tag}_traffic_$g.csv \
      python tools/profile_step.py --grid $g $g $g --steps 102 > $out/${tag}_traffic_$g.log 2>&1
done
# the first pass of step 101 (one jacobi_pass_kernel launch per step) and three later passes of a developed step
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jacobi_pass_kernel -s 100 -c 1 -o $out/${tag}_jacobi_pass0_256 -f \
    python tools/profile_step.py --grid 256 256 256 --steps 101 > $out/${tag}_ncu_j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jacobi_resident_kernel -s 1500 -c 3 -o $out/${tag}_jacobi_256 -f \
    python tools/profile_step.py --grid 256 256 256 --steps 101 >> $out/${tag}_ncu_j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"advect_kernel|divergence_quad|gradient_quad" -s 300 -c 3 -o $out/${tag}_adg_256 -f \
    python tools/profile_step.py --grid 256 256 256 --steps 101 > $out/${tag}_ncu_a.log 2>&1
ls -la $out | tail -15
