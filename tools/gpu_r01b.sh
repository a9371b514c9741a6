#!/bin/bash
# Round-1 pass B: barrier-skeleton timing of the expert-stream kernel, A/B of the two expert schedules, ncu launch
# list of one call and full-set captures of the three dominant kernels. Usage: bash tools/gpu_r01b.sh <tag>
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 300 python tools/expert_phases.py --flags 0,14,6 > $OUT/expert_phases.txt 2>&1; tail -30 $OUT/expert_phases.txt
VLASER_EXPERT_MEGA=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_mega0.json 2> $OUT/bench_mega0.err; echo "bench mega0 rc=$?"; cat $OUT/bench_mega0.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/launches.csv python tools/profile_call.py --graph 0 > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; cat $OUT/launch_summary.txt
for K in "pair:gemm_tcgen05_pair_kernel:40:2" "expert:expert_pass_kernel:1:1" "attn:^attention_kernel:30:2" "vitattn:vit_attention:5:1"; do
  IFS=: read NAME PAT SKIP CNT <<< "$K"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$PAT -s $SKIP -c $CNT \
      -f -o $OUT/full_$NAME python tools/profile_call.py --graph 0 > $OUT/ncu_full_$NAME.log 2>&1; echo "ncu full $NAME rc=$?"
done
ls -la $OUT
