mkdir -p gpurun_out
for w in C3-AND2 C3-AND4 C4 C1; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_$w.err | tee gpurun_out/wl_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WL $w', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline'].get('window_docs'), d['roofline'].get('doc_range_splits'), d['roofline']['postings_per_launch'])" || { echo "WL $w FAILED"; tail -3 gpurun_out/wl_$w.err; }
done
timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_C5s.err | tee gpurun_out/wl_C5s.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WL C5 x0.1', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline']['postings_per_launch'])" || echo "WL C5 FAILED"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
