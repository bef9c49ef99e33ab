# one --set full capture of lane_merge_topk_kernel on a scaled C2 (the full-size launch does not finish under ncu replay)
mkdir -p gpurun_out
timeout 120 python bench.py --scale 0.1 --queries 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/lane_s01.json 2>gpurun_out/lane_s01.err || { echo "plain run failed"; tail -3 gpurun_out/lane_s01.err; exit 1; }
cat gpurun_out/lane_s01.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S01', round(d['value']), d['roofline']['step_ms_by_kernel'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:staged_merge -c 1 -o gpurun_out/prof_lane python bench.py --scale 0.1 --queries 1000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lane.log 2>&1; echo "ncu rc=$?"
