# usage: bash tools/gpu_round2_check.sh TAG   -- full GPU test suite, smoke, both bench arms, the integer-peak kinds, the drop-in timing
T=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_gpu.txt 2>&1
nproc >> gpurun_out/${T}_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/${T}_pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
echo "ref exit $?" >> gpurun_out/${T}_ref.err
python - > gpurun_out/${T}_int_peaks.txt 2>&1 <<'PY'
import h264_b200_pkg
from h264_b200 import MeCuda
names = {0:"VABSDIFF4",1:"IADD3",2:"IMAD",3:"VIADDMNMX",4:"LOP3",5:"SHF",6:"VIMNMX",7:"PRMT",8:"VABSDIFF4+IADD3",9:"VABSDIFF4+IMAD",10:"VABSDIFF4+VIADDMNMX",
         11:"VABSDIFF4+LOP3",12:"VABSDIFF4+SHF",13:"VABSDIFF4+VIMNMX",14:"VABSDIFF4+PRMT",15:"VABSDIFF4+LDS",16:"VABSDIFF4+IMAD(sweep operands)",17:"VABSDIFF4+VIADDMNMX(sweep operands)"}
with MeCuda(176, 144) as me:
    for k in range(18):
        for rep in range(2):
            v = me.measure_int_peak(k)
        print("%2d %-40s %8.1f Glaneop/s" % (k, names[k], v / 1e9))
PY
timeout 400 bash tools/time_dropin.sh > gpurun_out/${T}_dropin.txt 2>&1
echo "dropin exit $?" >> gpurun_out/${T}_dropin.txt
tail -5 gpurun_out/${T}_pytest_gpu.txt; tail -2 gpurun_out/${T}_smoke.txt; tail -3 gpurun_out/${T}_dropin.txt; cat gpurun_out/${T}_int_peaks.txt; head -c 400 gpurun_out/${T}_bench.json
