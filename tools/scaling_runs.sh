#!/bin/bash
# Config 3 scaling runs: strong scaling of a 16M-clip DB and weak scaling at 12.5M clips per GPU.
# usage (on a box with G GPUs): bash tools/scaling_runs.sh G   -> one JSON line per run in gpurun_out/scale_*.json
G=${1:-4}
mkdir -p gpurun_out
run() {  # n_gpus clips_per_gpu tag
  local n=$1 c=$2 tag=$3
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 --clips-per-gpu $c --no-cold --no-cpu > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 30 --warmup 5 --clips-per-gpu $c --no-cold --no-cpu > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  fi
  echo "$tag rc=$? $(cut -c1-160 gpurun_out/scale_${tag}.json)"
}
for n in 1 2 4 8; do
  [ $n -le $G ] && run $n $((16000000 / n)) strong16M_n$n
done
for n in 2 4 8; do
  [ $n -le $G ] && [ $n -ge 2 ] && run $n 12500000 weak12p5M_n$n
done
