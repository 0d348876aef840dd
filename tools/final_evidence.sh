# usage: bash tools/final_evidence.sh TAG   (one GPU box) -- the evidence set of a round: GPU test suite, smoke, both bench arms, the
# launch list and the ncu --set full captures of the step's kernels (each after its own command has run without ncu), the
# integer-pipe probes, the drop-in timing, side measurements.  Everything lands in gpurun_out/TAG_*.
T=${1:-r02z}
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1; nproc >> gpurun_out/${T}_gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu_full.txt 2>&1; tail -4 gpurun_out/${T}_pytest_gpu_full.txt > gpurun_out/${T}_pytest_gpu.txt
grep -E "^(FAILED|ERROR)" gpurun_out/${T}_pytest_gpu_full.txt | head -20
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo smoke rc=$? >> gpurun_out/${T}_smoke.txt
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
ME_BENCH_WORKLOAD=4k python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_4k_sr64_side.json 2> gpurun_out/${T}_4k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:me_ffs_tuned -s 5 -c 1 -o gpurun_out/${T}_ffs python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_ffs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:me_subpel_tuned -s 5 -c 1 -o gpurun_out/${T}_subpel python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_subpel.log 2>&1
ncu --set full --clock-control none -k regex:me_planes -s 5 -c 1 -o gpurun_out/${T}_planes python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_planes.log 2>&1
if [ -z "$SKIP_PROBES" ]; then
tools/probe/pipe_probe > gpurun_out/${T}_pipe_probe.txt 2>&1
tools/probe/pipe_probe2 > gpurun_out/${T}_pipe_probe2.txt 2>&1
fi
bash tools/time_dropin.sh > gpurun_out/${T}_dropin.txt 2>&1
tail -3 gpurun_out/${T}_pytest_gpu.txt; tail -1 gpurun_out/${T}_smoke.txt; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_dropin.txt
