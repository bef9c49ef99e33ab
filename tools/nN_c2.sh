# C2 bench line on N GPUs: bash tools/nN_c2.sh N
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n${N}_final.json 2> gpurun_out/bench_c2_n${N}_final.err; echo "rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/bench_c2_n${N}_final.json').read().strip().split('\n')[-1]); print('N$N', round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['one_call_at_a_time']['value'], d['roofline']['step_ms_by_kernel'])"
