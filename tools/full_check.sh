# whole GPU suite (bounds-checked build first on the parity file, then the shipped build on everything), other workloads
mkdir -p gpurun_out
DGPU_LIB=$PWD/diagon_b200/libdiagon_b200_chk.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_chk.log 2>&1; rc=$?
echo "pytest(chk) rc=$rc"; tail -4 gpurun_out/pytest_chk.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?
echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_gpu.log
bash tools/workloads.sh
