#!/bin/bash
# Re-captures the ncu evidence under profiles/<round>/ on a GPU box (run through gpurun from the repo root):
#   bash tools/capture_profiles.sh        -> gpurun_out/cap_*   (copy what is to be judged into profiles/)
# Bench lines are taken WITHOUT a profiler (python bench.py); the ncu passes below are separate runs of the same command.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-extras"
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/cap_launches.csv $B > $O/cap_launches.log 2>&1
# full capture of the roofline kernel at the default workload (cfg3; 3 launches, with source) and at cfg2 / cfg2x2 (2 launches)
ncu --set full --clock-control none --import-source on -k regex:k_march_lean --launch-skip 30 --launch-count 3 \
    -f -o $O/cap_march_cfg3 $B --workload cfg3 > $O/cap_ncu_march_cfg3.log 2>&1
for w in cfg2 cfg2x2; do
  ncu --set full --clock-control none -k regex:k_march_lean --launch-skip 30 --launch-count 2 \
      -f -o $O/cap_march_$w $B --workload $w > $O/cap_ncu_march_$w.log 2>&1
done
# DRAM traffic only (dram__bytes) for the beam sweep: CSV text, no report file (gpurun_out/ is capped at 64 MiB)
for b in 270 540 1080 2160; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_march_lean \
      --launch-skip 30 --launch-count 3 --csv --log-file $O/cap_march_cfg5_$b.csv $B --workload cfg5_$b > $O/cap_ncu_march_cfg5_$b.log 2>&1
done
# the other two kernels of the tick at cfg3
ncu --set full --clock-control none --import-source on -k regex:"k_dynamics|k_tail" --launch-skip 60 --launch-count 2 \
    -f -o $O/cap_dyn_tail_cfg3 $B > $O/cap_ncu_dyn_tail.log 2>&1
du -sh $O
ls -la $O/cap_*
