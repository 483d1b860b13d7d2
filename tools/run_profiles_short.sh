#!/bin/bash
# The cheap part of tools/run_profiles.sh (GPU tests, bench lines, launch lists) + one ncu capture of the kernels named on the
# command line (default: the weight-gradient kernel) — for refreshing the evidence after a change that touched few kernels.
P=fp16x2
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --precision $P > $O/bench_final.json 2> $O/bench_final.err || tail -5 $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_ref.json 2> $O/bench_final_ref.err
NCU="ncu --clock-control none"
timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_$P.csv python tools/profile_step.py $P > $O/ncu_step.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_bench_$P.csv -c 2000 python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-extra > $O/ncu_bench.log 2>&1
export N=4
timeout 120 python tools/bench_kernel.py resblock_wgrad $P > /dev/null 2>&1 && timeout 300 $NCU --set full --import-source on -k regex:wgrad_tc_kernel -s 4 -c 1 -f -o $O/prof_wgrad_$P python tools/bench_kernel.py resblock_wgrad $P > $O/ncu_wgrad.log 2>&1
python tools/trace_wgrad.py > $O/trace_wgrad.txt 2>&1
cut -c1-200 $O/bench_final.json
