mkdir -p gpurun_out
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | one "C2 "
timeout 300 python bench.py --workload C4 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | one "C4 "
