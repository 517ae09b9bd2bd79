cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_err.log | grep "^{" > gpurun_out/r2_bench.json
for c in 2 3 5; do timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_err.log | grep "^{" > gpurun_out/r2_bench_cfg$c.json; done
python -c "
import json
for f in ['r2_bench','r2_bench_cfg2','r2_bench_cfg3','r2_bench_cfg5']:
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('wall_clock_to_convergence_s'), d.get('step_ms', [])[:4])
"
python __graft_entry__.py smoke 2>&1 | tail -2
