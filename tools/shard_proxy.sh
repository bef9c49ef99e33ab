# a 1/8-size C2 corpus on one GPU as a stand-in for one shard of the 8-GPU run: how finely should items be cut?
mkdir -p gpurun_out
for pf in 0 2 3 4; do
timeout 300 python bench.py --scale 0.125 --part-factor $pf --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/it.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PF $pf', round(d['value']), round(d['e2e']['value']), d['roofline']['step_ms_by_kernel'], d['roofline']['work_items'])" || tail -5 gpurun_out/it.err
done
