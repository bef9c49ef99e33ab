# 2 GPUs: parity of the sharded calls (one call, chunked, stream of batches) against an unsharded reader, then the C2 line
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/sharded_check.py 2>gpurun_out/n2_check.err | tail -14; echo "check rc=${PIPESTATUS[0]}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n2_final.json 2> gpurun_out/bench_c2_n2_final.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/bench_c2_n2_final.json').read().strip().split('\n')[-1]); print('N2', round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['one_call_at_a_time']['value'], d['roofline']['step_ms_by_kernel'])" || tail -5 gpurun_out/bench_c2_n2_final.err
