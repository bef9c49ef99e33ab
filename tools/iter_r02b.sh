# parity of the filter / union paths, then C4, C5 x0.1 and C2 (with the host-phase trace of one end-to-end call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log
for w in C4; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_$w.err | tee gpurun_out/wl_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WL $w', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || { echo "WL $w FAILED"; tail -3 gpurun_out/wl_$w.err; }
done
timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/wl_C5s.err | tee gpurun_out/wl_C5s.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('WL C5 x0.1', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline']['work_items'])" || echo "WL C5 FAILED"
DGPU_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/c2_trace.err | tee gpurun_out/c2_iter.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_step'], d['roofline']['step_ms_by_kernel'])"
grep "dgpu trace" gpurun_out/c2_trace.err | tail -8
