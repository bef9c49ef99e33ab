# C2 on 8 GPUs: bench line + host-phase trace of rank 0's end-to-end calls
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n8_final.json 2> gpurun_out/bench_c2_n8_final.err; echo "rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/bench_c2_n8_final.json').read().strip().split('\n')[-1]); print('N8', round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['one_call_at_a_time']['value'], d['roofline']['step_ms_by_kernel'])"
grep "dgpu trace" gpurun_out/bench_c2_n8_final.err | tail -40
