mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt 2>&1
nproc >> gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api_edges.py -m gpu -q --maxfail=12 > gpurun_out/r02a_pytest_parity.txt 2>&1
echo "parity exit $?" >> gpurun_out/r02a_pytest_parity.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?" >> gpurun_out/r02a_bench.err
timeout 400 bash tools/time_dropin.sh > gpurun_out/r02a_dropin.txt 2>&1
echo "dropin exit $?" >> gpurun_out/r02a_dropin.txt
timeout 1500 python -m pytest tests/test_jm_dropin.py -m gpu -q --maxfail=12 > gpurun_out/r02a_pytest_dropin.txt 2>&1
echo "dropin tests exit $?" >> gpurun_out/r02a_pytest_dropin.txt
tail -5 gpurun_out/r02a_pytest_parity.txt; tail -3 gpurun_out/r02a_dropin.txt; tail -5 gpurun_out/r02a_pytest_dropin.txt; head -c 600 gpurun_out/r02a_bench.json
