set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_splat.py -x -q 2>&1 | grep -v "^chunkset" | tail -30
