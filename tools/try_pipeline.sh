#!/bin/bash
# one-shot comparison of the two instances of the grow kernel (OLF_LSD_PIPELINE=1: software-pipelined walk, =0: plain):
# parity first, then throughput at 1 and 20 rigs on the same box.   gpurun -- bash tools/try_pipeline.sh
# (first run, round 2: the pipelined kernel was a separate -DOLF_GROW_PIPELINE=1 build of the library; it is an instance of the template now)
bench() { tag=$1; P=$2; timeout 120 python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines $P > gpurun_out/pf_$tag.log 2>gpurun_out/pf_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/pf_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
}
export OLF_LSD_PIPELINE=1
timeout 150 python -m pytest tests/test_gpu_line.py tests/test_frontend_golden.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -4
timeout 90 python tools/lsd_gpu_stress.py 30 2>&1 | tail -3
bench pf_p1 1; bench pf_p20 20
export OLF_LSD_PIPELINE=0
bench base_p1 1; bench base_p20 20
