set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r01o_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01o_smoke.txt 2>&1; echo smoke rc=$? >> gpurun_out/r01o_smoke.txt
python bench.py > gpurun_out/r01o_bench.json 2> gpurun_out/r01o_bench.err
python bench.py --impl reference > gpurun_out/r01o_bench_reference_arm.json 2> gpurun_out/r01o_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01o_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01o_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:me_ffs_tuned -s 5 -c 1 -o gpurun_out/r01o_ffs python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01o_ncu_ffs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:me_subpel_tuned -s 5 -c 1 -o gpurun_out/r01o_subpel python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01o_ncu_subpel.log 2>&1
ncu --set full --clock-control none -k regex:me_planes -s 5 -c 1 -o gpurun_out/r01o_planes python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01o_ncu_planes.log 2>&1
tail -2 gpurun_out/r01o_pytest_gpu.txt; tail -1 gpurun_out/r01o_smoke.txt; cut -c1-300 gpurun_out/r01o_bench.json
