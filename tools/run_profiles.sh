#!/bin/bash
# Evidence pass on the GPU box (run through gpurun from the repo root): GPU tests, the bench line (+ the reference arm), the
# per-launch time list of one eager train_step and of the bench command, one `ncu --set full` capture per hot kernel (each only
# after the same command has exited 0 without ncu).  Everything lands in gpurun_out/; tools/make_profiles.py turns it into the
# tracked summaries under profiles/.
# Usage: gpurun --timeout 1500 -- 'bash tools/run_profiles.sh [precision]'
P=${1:-fp16x2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --precision $P > $O/bench_final.json 2> $O/bench_final.err || tail -5 $O/bench_final.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_ref.json 2> $O/bench_final_ref.err
timeout 200 python bench.py --workload infer > $O/bench_final_infer.json 2> $O/bench_final_infer.err
timeout 300 python tools/precision_report.py fp32 fp16x2 bf16x3 > /dev/null 2> $O/precision.err
NCU="ncu --clock-control none"
timeout 300 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_$P.csv python tools/profile_step.py $P > $O/ncu_step.log 2>&1
# the same pass over the bench command itself (eager first step, capture, graph replays: kernel nodes are listed one by one)
timeout 400 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_bench_$P.csv -c 2000 python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-extra > $O/ncu_bench.log 2>&1
full() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 120 "$@" > /dev/null 2>&1 || { echo "plain run of $name failed"; return; }
  timeout 300 $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/prof_$name "$@" > $O/ncu_$name.log 2>&1
}
export N=4
full rs_infer rs_kernel 3 python tools/bench_kernel.py stack_infer $P
full rs_train rs_kernel 4 python tools/bench_kernel.py stack_train $P
full rs_bwd rs_kernel 3 python tools/bench_kernel.py stack_bwd $P
DILS=27,9,3,1 full rs_train_rev rs_kernel 4 python tools/bench_kernel.py stack_train $P
full wgrad_$P wgrad_tc_kernel 4 python tools/bench_kernel.py resblock_wgrad $P
full conv_tc conv_tc_kernel 2 python tools/bench_kernel.py conv_down $P
L=3520 full conv3_tc conv3_tc_kernel 2 python tools/bench_kernel.py conv3_up $P
full wgrad4_tc wgrad4_tc_kernel 2 python tools/bench_kernel.py conv_down_wgrad $P
full rb_fwd_masks_$P rb_tc_kernel 5 python tools/bench_kernel.py resblock_fwd_masks $P
full vq_search vq2_kernel 2 python tools/profile_vq.py
full vq_finish vq_finish_smem_kernel 2 python tools/profile_vq.py
python tools/trace_vq.py > $O/trace_vq.txt 2>&1
python tools/trace_wgrad.py > $O/trace_wgrad.txt 2>&1
[ -x tools/mma_probe ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I"vae-based-music--deep-generative-models_b200/csrc" -Iinclude -o tools/mma_probe tools/mma_probe.cu
./tools/mma_probe > $O/mma_probe.txt 2>&1
ls -la $O/prof_*.ncu-rep 2>/dev/null | awk '{print $5, $9}'
cut -c1-300 $O/bench_final.json
