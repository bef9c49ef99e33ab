# end-to-end call: chunk geometry and part-cut hints (experiment)
mkdir -p gpurun_out
run() { env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $EXTRA 2>gpurun_out/pipe.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PIPE $* $EXTRA', round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2))" || tail -3 gpurun_out/pipe.err; }
EXTRA=""
run DGPU_PIPE_HINT=0
run DGPU_PIPE_HINT=3
run DGPU_PIPE_HINT=2
run DGPU_PIPE_HINT=4
EXTRA="--pipeline-chunks 4"
run DGPU_PIPE_HINT=0
run DGPU_PIPE_HINT=3
run DGPU_PIPE_HINT=0 DGPU_PIPE_RATIO=2
run DGPU_PIPE_HINT=3 DGPU_PIPE_RATIO=2
run DGPU_PIPE_HINT=4 DGPU_PIPE_RATIO=2
EXTRA="--pipeline-chunks 3"
run DGPU_PIPE_HINT=0 DGPU_PIPE_RATIO=3
run DGPU_PIPE_HINT=4 DGPU_PIPE_RATIO=3
run DGPU_PIPE_HINT=0 DGPU_PIPE_RATIO=2
EXTRA="--pipeline-chunks 2"
run DGPU_PIPE_HINT=0 DGPU_PIPE_RATIO=4
run DGPU_PIPE_HINT=0 DGPU_PIPE_RATIO=2
EXTRA="--pipeline-chunks 5"
run DGPU_PIPE_HINT=3 DGPU_PIPE_RATIO=1.6
