#!/bin/bash
# C4 (overlap-add separation of a 10-minute mix, strong scaling) and C5 (training step, weak scaling; with and without the
# MR-STFT term) at 1 / 2 / 4 / 8 GPUs on ONE 8-GPU box:  gpurun --gpus 8 -- 'bash tools/run_multigpu.sh r2'
# N = 8 runs alone; N = 4, 2, 1 then run side by side on disjoint GPUs (0-3 | 4-5 | 6) to keep the lease short.
# Output: gpurun_out/<tag>_{c4,c5,c5mr}_<N>gpu.json (+ .err)
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
run() {  # run <N> <devices> <port> <script...>
  local n=$1 dev=$2 port=$3; shift 3
  if [ "$n" = 1 ]; then CUDA_VISIBLE_DEVICES=$dev python "$@"
  else CUDA_VISIBLE_DEVICES=$dev python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; fi
}
suite() {  # suite <N> <devices> <port>
  local n=$1 dev=$2 port=$3
  run $n $dev $port tools/ola_bench.py --reps 2 > $out/${tag}_c4_${n}gpu.json 2> $out/${tag}_c4_${n}gpu.err
  run $n $dev $((port+1)) tools/train_bench.py --batch 8 --steps 5 --warmup 3 --measure-exposed > $out/${tag}_c5_${n}gpu.json 2> $out/${tag}_c5_${n}gpu.err
  run $n $dev $((port+2)) tools/train_bench.py --batch 8 --steps 5 --warmup 3 --mrstft > $out/${tag}_c5mr_${n}gpu.json 2> $out/${tag}_c5mr_${n}gpu.err
}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_multigpu_smi.txt
suite 8 0,1,2,3,4,5,6,7 29500
suite 4 0,1,2,3 29510 &
suite 2 4,5 29520 &
suite 1 6 29530 &
wait
for f in $out/${tag}_c4_*gpu.json $out/${tag}_c5_*gpu.json $out/${tag}_c5mr_*gpu.json; do echo "== $f"; tail -c 700 $f; echo; done
