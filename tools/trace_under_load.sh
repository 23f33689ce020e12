#!/bin/bash
# per-round timing of one image's LSD passes alone and while a 19-rig bench runs on the same GPU (MPS so that the two processes run concurrently)
python tools/lsd_trace_batch.py > gpurun_out/trace_alone.txt 2>&1
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps CUDA_MPS_LOG_DIRECTORY=/tmp/mps_log; mkdir -p /tmp/mps /tmp/mps_log
nvidia-cuda-mps-control -d && sleep 2
python tools/lsd_trace_batch.py > gpurun_out/trace_mps_alone.txt 2>&1
export TRACE_T0=$(date +%s)
python bench.py --steps 300 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines 19 > gpurun_out/trace_bench.log 2>/dev/null &
python tools/lsd_trace_batch.py 38 > gpurun_out/trace_loaded.txt 2>&1          # keeps calling until 38 s after the start: inside the bench's first timed region
wait
echo quit | nvidia-cuda-mps-control; sleep 1
grep -o '"value": [0-9.]*' gpurun_out/trace_bench.log | head -1
head -1 gpurun_out/trace_alone.txt; head -1 gpurun_out/trace_mps_alone.txt; head -1 gpurun_out/trace_loaded.txt
