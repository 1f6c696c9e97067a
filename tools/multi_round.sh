#!/bin/bash
# One multi-GPU gpurun call: NCCL strip check, timeline, strong-scaling bench (and the sharded sweep at N = 8).
# Usage (via gpurun --gpus N):  bash tools/multi_round.sh N tag
N=${1:-2}; TAG=${2:-r02}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
FULL=""; [ "$N" = "8" ] && FULL="--full"
timeout 900 $TR tools/check_strips_nccl.py $FULL > gpurun_out/${TAG}_check_n$N.txt 2>&1; echo "check rc=$?"; tail -6 gpurun_out/${TAG}_check_n$N.txt
timeout 600 $TR tools/strip_timeline.py > gpurun_out/${TAG}_timeline_n$N.txt 2>&1; echo "timeline rc=$?"; tail -$N gpurun_out/${TAG}_timeline_n$N.txt | cut -c1-260
timeout 900 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_scale_n$N.json
if [ "$N" = "8" ]; then
  timeout 900 $TR bench.py --gpus $N --workload datagen256 --steps 1000 --warmup 20 > gpurun_out/${TAG}_datagen_n$N.json 2>> gpurun_out/${TAG}_scale_n$N.err; echo "datagen rc=$?"; cat gpurun_out/${TAG}_datagen_n$N.json
fi
