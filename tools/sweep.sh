# bench lines for a list of configurations: each argument is "workload|extra bench flags"
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  wl=${spec%%|*}; flags=${spec#*|}
  timeout 400 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline $flags > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - "$spec" gpurun_out/sweep_$i.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    r=d['roofline']
    print('SWEEP', sys.argv[1], '| value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v,3) for k,v in r.get('step_ms_by_kernel',{}).items()}, r.get('work_items'))
except Exception as e:
    print('SWEEP', sys.argv[1], 'FAILED', e)
PY
done
