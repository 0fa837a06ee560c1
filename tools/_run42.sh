mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "overlaps or fuzz or errors" 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_overlaps.log
