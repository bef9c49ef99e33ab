mkdir -p gpurun_out
timeout 60 python tools/debug1.py "AND body t0000001 t0000005 t0000020" || exit 1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v8.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/pytest_v8.log
[ $rc -ne 0 ] && exit 1
for w in C3-AND2 C3-AND4; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_$w.err | tee gpurun_out/wl_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WL $w', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('doc_range_splits'), d['gpu_launches'])" || { echo "WL $w FAILED"; tail -3 gpurun_out/wl_$w.err; }
done
