#!/bin/bash
# Round-2 profile / bench collection on the GPU box (run through gpurun).  ncu reports stay in /tmp on the box; only their
# CSV exports and the bench lines are written to gpurun_out/ (which is copied back, <= 64 MiB).
#   usage: bash profiles/micro/r2_collect.sh [profiles|benches|all]
what=${1:-all}
mkdir -p gpurun_out
if [ "$what" != benches ]; then
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python bench.py --profile 2 > gpurun_out/launches_r2.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/step_r2 -f python bench.py --profile 1 > gpurun_out/step_r2.log 2>&1
ncu -i /tmp/step_r2.ncu-rep --page raw --csv > gpurun_out/step_r2_raw.csv 2>/dev/null
# the producers: skip the query-side and first encoder launches, capture one encoder layer and the first 5H block at the
# passage size (LayerNorm, QKV GEMM, attention, out-projection GEMM, fused feed-forward, Interaction)
timeout -k 10 900 ncu --set full --clock-control none -k regex:"gemm_rows|ffn_rows|interaction_kernel|enc_attention|ln_rows" -s 44 -c 36 -o /tmp/producers_r2 -f python profiles/micro/producers_timing.py c2 tc > gpurun_out/producers_r2.log 2>&1
ncu -i /tmp/producers_r2.ncu-rep --page raw --csv > gpurun_out/producers_r2_raw.csv 2>/dev/null
fi
if [ "$what" != profiles ]; then
for c in c1 c3 c4 c5; do timeout -k 10 600 python bench.py --config $c --no-extras > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; done
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err
for c in c1 c2 c5; do timeout -k 10 300 python profiles/micro/producers_timing.py $c tc 2>&1 | grep -v -i warn; done > gpurun_out/producers_timing_r2.txt
(timeout -k 10 200 python profiles/micro/gemm_rows_bench.py 2>&1 | grep -v -i warn; timeout -k 10 200 python profiles/micro/gemm_rows_bench.py ffn 2>&1 | grep -v -i warn) > gpurun_out/gemm_rows_bench_r2.txt
fi
ls -la gpurun_out | tail -20; du -sh gpurun_out
