#!/bin/bash
# multi-GPU round: NCCL strip validation + strong-scaling bench at N = 8, 4, 2 (run with gpurun --gpus 8)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 4 --master-port 29501 tools/check_strips_nccl.py 2>&1 | grep -E "strips|sweep|STRIPS_OK|Error|error" | tail -12
for N in 8 4 2; do
  $TR --nproc-per-node $N --master-port $((29510+N)) bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  echo "N=$N rc=$?"; cat gpurun_out/scale_n$N.json; grep -iE "error|Traceback" gpurun_out/scale_n$N.err | head -3
done
$TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 --no-overlap > gpurun_out/scale_n8_nooverlap.json 2>/dev/null; cat gpurun_out/scale_n8_nooverlap.json
python bench.py --workload cavity32768 --steps 10 --warmup 3 > gpurun_out/scale_n1_32768.json 2> gpurun_out/scale_n1_32768.err; echo "N=1 32768 rc=$?"; cat gpurun_out/scale_n1_32768.json; tail -2 gpurun_out/scale_n1_32768.err
nvidia-smi topo -m | head -12
