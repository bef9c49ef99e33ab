# same-box A/B of the C5-shaped workload: round-1 tree (r1tree/) against the working tree; then a C4 capture
mkdir -p gpurun_out
one() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'])"; }
(cd r1tree && timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>../gpurun_out/ab_r1.err | one "C5 r1  ")
timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_now.err | one "C5 now "
(cd r1tree && timeout 300 python bench.py --workload C5 --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>../gpurun_out/ab_r1.err | one "C5 r1  ")
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_c2.err | one "C2 now "
timeout 600 ncu --set full --clock-control none --import-source on -k regex:union_topk -c 1 -o gpurun_out/prof_c4_r02b python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_r02b.log 2>&1; echo "ncu c4 rc=$?"
