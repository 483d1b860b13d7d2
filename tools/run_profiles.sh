#!/bin/bash
# Evidence pass on the GPU box (run through gpurun from the repo root): GPU tests, the bench line, the per-launch time
# list of one eager train_step and one `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/;
# tools/make_profiles.py turns it into the tracked summaries under profiles/.
# Usage: gpurun --timeout 1500 -- 'bash tools/run_profiles.sh [precision]'
P=${1:-fp16x2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --precision $P > $O/bench_final.json 2> $O/bench_final.err || tail -5 $O/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_final_ref.json 2> $O/bench_final_ref.err
NCU="ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_$P.csv python tools/profile_step.py $P > $O/ncu_step.log 2>&1
# the same pass over the bench command itself (eager first step, capture, graph replays: kernel nodes are listed one by one)
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_bench_$P.csv -c 2500 python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > $O/ncu_bench.log 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  "$@" > /dev/null 2>&1 || { echo "plain run of $name failed"; return; }
  $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
}
full rb_fwd_$P rb_tc_kernel 4 python tools/bench_kernel.py resblock_fwd $P
full rb_fwd_masks_$P rb_tc_kernel 5 python tools/bench_kernel.py resblock_fwd_masks $P
full rb_bwd_masks_$P rb_tc_kernel 5 python tools/bench_kernel.py resblock_bwd_masks $P
DIL=27 full rb_bwd_masks_d27_$P rb_tc_kernel 5 python tools/bench_kernel.py resblock_bwd_masks $P
full wgrad_$P wgrad_tc_kernel 4 python tools/bench_kernel.py wgrad $P
full vq_search vq2_kernel 2 python tools/profile_vq.py
full vq_finish vq_finish_smem_kernel 2 python tools/profile_vq.py
ls -la $O/prof_*_$P.ncu-rep $O/prof_vq_*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
cut -c1-400 $O/bench_final.json
