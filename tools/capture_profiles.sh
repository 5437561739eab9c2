#!/bin/bash
# Re-captures everything under profiles/<round>/ on a GPU box (run through gpurun from the repo root):
#   bash tools/capture_profiles.sh        -> gpurun_out/cap_*   (copy what is to be judged into profiles/)
# Bench lines are taken WITHOUT a profiler; the ncu passes are separate runs of the same command.
set -u
O=gpurun_out
mkdir -p $O
python bench.py                                   > $O/cap_bench_cfg2.json     2> $O/cap_bench_cfg2.err
python bench.py --workload cfg2x2 --no-cpu        > $O/cap_bench_cfg2x2.json   2> $O/cap_bench_cfg2x2.err
python bench.py --workload cfg3 --no-cpu --steps 200 --warmup 20 > $O/cap_bench_cfg3.json 2> $O/cap_bench_cfg3.err
for b in 270 540 1080 2160; do
  python bench.py --workload cfg5_$b --no-cpu --steps 100 --warmup 10 > $O/cap_bench_cfg5_$b.json 2> $O/cap_bench_cfg5_$b.err
done
python bench.py --impl reference --steps 50 --warmup 5 > $O/cap_bench_reference.json 2> $O/cap_bench_reference.err
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/cap_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu > $O/cap_launches.log 2>&1
# full capture of the roofline kernel (3 launches) and of the other two kernels of the tick (1 launch each)
ncu --set full --clock-control none --import-source on -k regex:k_march_persistent --launch-skip 30 --launch-count 3 \
    -f -o $O/cap_march python bench.py --steps 20 --warmup 5 --no-cpu > $O/cap_ncu_march.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dynamics|k_tail" --launch-skip 60 --launch-count 2 \
    -f -o $O/cap_dyn_tail python bench.py --steps 20 --warmup 5 --no-cpu > $O/cap_ncu_dyn_tail.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tail --launch-skip 30 --launch-count 1 \
    -f -o $O/cap_tail_a2 python bench.py --workload cfg2x2 --steps 20 --warmup 5 --no-cpu > $O/cap_ncu_tail_a2.log 2>&1
python tools/march_timeline.py > $O/cap_timeline.log 2>&1
ls -la $O/cap_*
