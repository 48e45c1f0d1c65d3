#!/bin/bash
# Runs on the GPU box (gpurun): the measurements and captures committed under profiles/ for this round.
# Everything is written below gpurun_out/final/.  ncu passes run only after the same command has exited 0 without ncu.
set -u
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_mt.json 2> $O/bench_mt.err; echo "bench rc=$?"
python bench.py --rng philox --no-cpu > $O/bench_philox.json 2> $O/bench_philox.err
python bench.py --samples-per-gpu 16 --steps 3 --warmup 3 --no-cpu > $O/bench_mt_16samples.json 2> $O/bench_16.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err
# launch list of the bench command; rounds enqueued by the host (CBS_GPU_GRAPH=0): ncu does not profile the kernel nodes
# of a graph that has conditional nodes
if CBS_GPU_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_pre_ncu.json 2>/dev/null; then
  CBS_GPU_GRAPH=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
# one call, planned work per round, then the full capture of round 4 (rounds enqueued by the host: launch k of a kernel = round k)
CBS_GPU_GRAPH=0 CBS_GPU_DEBUG_ROUNDS=1 python tools/one_step.py > $O/one_step.txt 2>&1 && \
CBS_GPU_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_chain|k_shuffle|k_scan" -s 40 -c 10 \
    -o $O/round4 -f python tools/one_step.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/round_timeline.py > $O/timeline.txt 2>&1
# the command-line driver end to end on a 32-sample matrix
tools/make_cn /tmp/c32.cn 32 && for i in 1 2 3; do genomic_b200/cna_segment_gpu --timing --chain 0 --nperm 10000 /tmp/c32.cn /tmp/c32.seg 2>> $O/driver_e2e.txt; done
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
ls -la $O
