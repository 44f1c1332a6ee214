#!/bin/bash
# round 2, first GPU visit: the whole GPU suite, then whole-program throughput of the drop-in
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests.log
tail -5 gpurun_out/r2_tests.log
timeout 600 python scripts/gmap_throughput.py --queries 2000 --threads 16,64,256,1024 > gpurun_out/r2_throughput.json 2> gpurun_out/r2_throughput.err
tail -12 gpurun_out/r2_throughput.err; cat gpurun_out/r2_throughput.json
