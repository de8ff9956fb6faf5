#!/bin/bash
# BASELINE.json configs 2-5 at N GPUs of one box: PIDNet-S 1024x2048 bs32 (+ the training record, config 5), PIDNet-M CamVid
# 720x960 bs32 (config 3), PIDNet-L 1024x2048 bs16 (config 4).  usage: tools/measure_scaling.sh N  (under `gpurun --gpus N`)
set -u
N=${1:-1}
O=gpurun_out/scale_n$N; mkdir -p $O
run() {  # name, extra args...
  local name=$1; shift
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err
  fi
  echo "$name rc=$? $(cut -c1-160 $O/$name.json)"
}
if [ "$N" = 1 ]; then run bench_pidnet_s_1024x2048 --steps 20 --warmup 5; else run bench_pidnet_s_1024x2048 --steps 20 --warmup 5 --skip-ref-gpu --skip-cpu-baseline; fi
run bench_pidnet_m_720x960 --model pidnet_m --classes 11 --batch 32 --height 720 --width 960 --steps 20 --warmup 5 --skip-train --skip-ref-gpu --skip-cpu-baseline
run bench_pidnet_l_1024x2048 --model pidnet_l --classes 19 --batch 16 --height 1024 --width 2048 --steps 20 --warmup 5 --skip-train --skip-ref-gpu --skip-cpu-baseline
if [ "$N" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_grad_worker.py > $O/nccl_grad_worker.log 2>&1
  echo "nccl worker rc=$? $(grep -c NCCL_GRAD_OK $O/nccl_grad_worker.log)"
fi
nvidia-smi topo -m > $O/topo.txt 2>&1
