#!/bin/bash
# Probe the GPU box: host cores/RAM, GPU clocks, then the FP64 peak microbenchmark with clock sampling.
mkdir -p gpurun_out
{ echo "nproc=$(nproc)"; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv; python -c "import os;print('cpu_count',os.cpu_count())"; lscpu | head -20; } > gpurun_out/box_info.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/fp64_peak_clocks.csv &
SMI=$!
tools/bin/fp64_peak 0.5 > gpurun_out/fp64_peak.jsonl 2> gpurun_out/fp64_peak.err
echo "exit=$?" >> gpurun_out/fp64_peak.err
kill $SMI
cat gpurun_out/box_info.txt; cat gpurun_out/fp64_peak.jsonl; tail -3 gpurun_out/fp64_peak.err
