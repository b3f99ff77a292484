mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_materials.py -q -m gpu -k "plane or catcher" 2>&1 | tail -25
