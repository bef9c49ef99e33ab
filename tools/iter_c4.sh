mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for fs in 1 0; do
timeout 300 python bench.py --workload C4 --filter-stream $fs --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 fs=$fs', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || tail -5 gpurun_out/it.err
done
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || tail -5 gpurun_out/it.err
