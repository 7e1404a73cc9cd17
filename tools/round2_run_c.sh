set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_keygen.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02c_pytest_gpu.log
