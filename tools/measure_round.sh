#!/bin/bash
# The single-GPU measurement pass behind profiles/r2_*: driver-style bench lines (config 4 + reference arm, configs 2 / 3 / 5), the ncu launch
# list of a bench run and one ncu --set full capture of the second LM iteration.  Run on the GPU box: gpurun -- "bash tools/measure_round.sh"
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_err.log | grep "^{" > gpurun_out/r2_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | grep "^{" > gpurun_out/r2_bench_reference.json
for c in 2 3 5; do timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_err.log | grep "^{" > gpurun_out/r2_bench_cfg$c.json; done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 27 -c 23 -f -o gpurun_out/r2_ncu_full python tools/profile_lm_step.py 4 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python -c "
import json
for f in ['r2_bench','r2_bench_reference','r2_bench_cfg2','r2_bench_cfg3','r2_bench_cfg5']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('wall_clock_to_convergence_s'), d.get('clocks'))
"
tail -3 gpurun_out/bench_err.log
