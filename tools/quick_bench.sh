# quick device-time check of the bench workload; extra bench.py flags as arguments
python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/quick.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('QB', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || { tail -5 gpurun_out/quick.err; exit 1; }
