# parity of union_topk_kernel (bounds-checked build), then one --set full capture of it on a scaled C2
mkdir -p gpurun_out
DGPU_LIB=$PWD/diagon_b200/libdiagon_b200_chk.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "union or lane_merge3" > gpurun_out/pytest_union_chk.log 2>&1; rc=$?
echo "pytest(chk) rc=$rc"; tail -15 gpurun_out/pytest_union_chk.log
timeout 120 python bench.py --scale 0.1 --queries 1000 --steps 2 --warmup 1 --no-cpu-baseline --lane-merge 3 > gpurun_out/union_s01.json 2>gpurun_out/union_s01.err || { echo "plain run failed"; tail -3 gpurun_out/union_s01.err; exit 1; }
cat gpurun_out/union_s01.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S01', round(d['value']), d['roofline']['step_ms_by_kernel'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:union_topk -c 1 -o gpurun_out/prof_union python bench.py --scale 0.1 --queries 1000 --steps 1 --warmup 1 --no-cpu-baseline --lane-merge 3 > gpurun_out/ncu_union.log 2>&1; echo "ncu rc=$?"
