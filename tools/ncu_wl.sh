# one --set full capture of a scoring kernel on a scaled workload: bash tools/ncu_wl.sh <workload> <kernel regex> <out name>
mkdir -p gpurun_out
W=$1; K=$2; O=$3
timeout 120 python bench.py --workload $W --scale 0.1 --queries 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/$O.json 2>gpurun_out/$O.err || { echo "plain run failed"; tail -3 gpurun_out/$O.err; exit 1; }
cat gpurun_out/$O.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S01', round(d['value']), d['roofline']['step_ms_by_kernel'], d['roofline']['postings_per_launch'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o gpurun_out/$O python bench.py --workload $W --scale 0.1 --queries 1000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$O.log 2>&1; echo "ncu rc=$?"
