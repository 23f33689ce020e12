#!/bin/bash
# diagnostic: multi-rank bench with the CUDA-graph chain / TMA staging switched off one at a time (gpurun --gpus N -- bash tools/diag_multi.sh N)
N=${1:-2}
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --prewarm-steps 40 > gpurun_out/diag_$tag.log 2> gpurun_out/diag_$tag.err; rc=$?; echo "$tag rc=$rc $(grep -o '"value": [0-9.]*' gpurun_out/diag_$tag.log | head -2 | tr '\n' ' ') $(grep -m1 -o 'CUDA error.*' gpurun_out/diag_$tag.err)"; return $rc; }
run default A=1 && run default2 A=1 && exit 0
run nograph OLF_LSD_GRAPH=0
run notma OLF_NO_TMA=1
run neither OLF_LSD_GRAPH=0 OLF_NO_TMA=1
