#!/bin/bash
# One gpurun call: parity tests, bench (N=1), launch list and one full ncu capture of the step kernel.
# Usage (from the repo root, via gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
# launch list of the bench command (short K), only after the same command exited 0 without ncu
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# full capture of the dominant kernel (fp64 MRT then fp32 MRT), 2 launches each
python tools/quick_perf.py 4096 4096 float64 > gpurun_out/${TAG}_plain64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lbm_step -s 14 -c 2 -o gpurun_out/${TAG}_f64 -f \
    python tools/quick_perf.py 4096 4096 float64 > gpurun_out/${TAG}_ncu64.log 2>&1
echo "ncu f64 rc=$?"
python tools/quick_perf.py 4096 4096 float32 > gpurun_out/${TAG}_plain32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lbm_step -s 14 -c 2 -o gpurun_out/${TAG}_f32 -f \
    python tools/quick_perf.py 4096 4096 float32 > gpurun_out/${TAG}_ncu32.log 2>&1
echo "ncu f32 rc=$?"
cat gpurun_out/${TAG}_plain64.log gpurun_out/${TAG}_plain32.log
ls -la gpurun_out/
