mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_image_cache.py tests/test_segment_reader.py -m gpu -x -q 2>&1 | tail -3
for w in C2 C4; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])" || tail -5 gpurun_out/it.err
done
