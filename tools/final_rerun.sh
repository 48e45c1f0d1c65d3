#!/bin/bash
# Runs on the GPU box: bench lines, tests and the launch list of the final build (the full capture of round 4 and the reference
# arm come from tools/final_captures.sh on the same kernels).
set -u
O=gpurun_out/final2
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_mt.json 2> $O/bench_mt.err; echo "bench rc=$?"
python bench.py --rng philox --no-cpu > $O/bench_philox.json 2> $O/bench_philox.err
python bench.py --samples-per-gpu 16 --steps 3 --warmup 3 --no-cpu > $O/bench_mt_16samples.json 2> $O/bench_16.err
if CBS_GPU_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_pre_ncu.json 2>/dev/null; then
  CBS_GPU_GRAPH=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
python tools/round_timeline.py > $O/timeline.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
