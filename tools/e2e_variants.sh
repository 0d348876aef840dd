# usage: bash tools/e2e_variants.sh TAG N [modes...] -- end-to-end figure of bench.py under torchrun on N GPUs for the host-buffer
# variants of me_cuda_frame_search (ME_CUDA_ZEROCOPY: 1 pinned caller arrays read and written in place, 0 chunked copy pipeline)
T=$1; N=$2; shift 2
for m in "${@:-1 0}"; do
  ME_CUDA_ZEROCOPY=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_mode${m}_${N}gpu.json 2> gpurun_out/${T}_mode${m}.err
  echo "mode $m, $N GPUs: e2e $(grep -o '"e2e": {[^}]*}' gpurun_out/${T}_mode${m}_${N}gpu.json | grep -o '"ms_per_step": [0-9.]*') device $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/${T}_mode${m}_${N}gpu.json | head -1)"
done
