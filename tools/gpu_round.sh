#!/bin/bash
# One gpurun call: parity tests, smoke, bench (N=1) + reference arm, launch list and full ncu captures of the step kernel.
# Usage (from the repo root, via gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
python bench.py --workload datagen256 --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_datagen256.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_datagen256.json
# launch list of the bench command (short K), only after the same command exited 0 without ncu
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# full capture of the dominant kernel (fp64 MRT then fp32 MRT), 2 launches each
bash tools/ncu_one.sh ${TAG}_f64 lbm_step_slide2 "" 4096 4096 float64
bash tools/ncu_one.sh ${TAG}_f32 lbm_step_slide2 "" 4096 4096 float32
ls -la gpurun_out/ | tail -20
