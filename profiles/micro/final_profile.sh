#!/bin/bash
# Round-end evidence run (one B200): default bench line, launch lists (official cold-cache recipe pass and a
# warm-cache pass), one `ncu --set full` capture of every kernel of one decode step.  Outputs in gpurun_out/.
set -x
python bench.py > gpurun_out/r1_final_bench.json 2> gpurun_out/r1_final_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1_final_ref.json 2> gpurun_out/r1_final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --profile 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r1_final_launches_warm.csv python bench.py --profile 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r1_final_top python bench.py --profile 1 > /dev/null 2>&1
ls -la gpurun_out/r1_final*
