# whole GPU suite on the current default build (leaf cost 0.7, 8 walk CTAs / SM, new entry points)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r02zb_pytest_gpu.txt 2>&1; echo rc=$? >> gpurun_out/r02zb_pytest_gpu.txt
grep -E "passed|failed|rc=" gpurun_out/r02zb_pytest_gpu.txt | tail -3
grep -E "^(FAILED|ERROR)" gpurun_out/r02zb_pytest_gpu.txt | head -20
