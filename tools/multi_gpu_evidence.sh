# usage: bash tools/multi_gpu_evidence.sh TAG N   (a box with N GPUs): torchrun bench lines and the GOP-sharded real encoder
T=${1:-r02o}; N=${2:-8}
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_${N}gpu.err
if [ "$N" = "8" ]; then
  ME_BENCH_WORKLOAD=4k python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_4k_sr64_${N}gpu_side.json 2>> gpurun_out/${T}_${N}gpu.err
  # the real encoder, one lencod_cuda process per GOP and GPU, every GOP compared with the plain reference encoder
  python h.264_b200/gop_shard.py --gpus $N --gops 8 --gop-frames 3 --width 1920 --height 1088 --search-range 32 --refs 5 --compare > gpurun_out/${T}_gop_shard_1080p_${N}gpu.json 2> gpurun_out/${T}_gop_shard.err
  python h.264_b200/gop_shard.py --gpus 1 --gops 8 --gop-frames 3 --width 1920 --height 1088 --search-range 32 --refs 5 > gpurun_out/${T}_gop_shard_1080p_1gpu.json 2>> gpurun_out/${T}_gop_shard.err
  # BASELINE config 5 shape: 4K, SR 64, 8 GOPs (I + P each) on the 8 GPUs
  python h.264_b200/gop_shard.py --gpus $N --gops 8 --gop-frames 2 --width 3840 --height 2160 --search-range 64 --refs 5 > gpurun_out/${T}_gop_shard_4k_sr64_${N}gpu.json 2>> gpurun_out/${T}_gop_shard.err
  python -m pytest tests/test_jm_dropin.py -m gpu -q -k gop_sharded 2>&1 | tail -3 > gpurun_out/${T}_pytest_gop.txt
fi
cut -c1-300 gpurun_out/${T}_bench_${N}gpu.json; cat gpurun_out/${T}_gop_shard_*.json 2>/dev/null | cut -c1-600
