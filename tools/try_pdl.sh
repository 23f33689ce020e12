#!/bin/bash
# EXPERIMENT: programmatic dependent launch between the LSD pass kernels (a -DOLF_PDL=1 build of the library, plain-launch chain).
# Build it first:  (cd orb_line_slam_b200/csrc && nvcc <flags of build.py> -DOLF_PDL=1 -o ../../tools/libolf_pdl.so *.cu -lcudart)
# then:            gpurun -- bash tools/try_pdl.sh        (parity of the line half, then frames/s at 20 rigs and at 1 rig)
bench() { tag=$1; P=$2; timeout 45 python bench.py --steps 8 --warmup 3 --prewarm-steps 12 --no-cpu-baseline --pipelines $P > gpurun_out/pdl_$tag.log 2>gpurun_out/pdl_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/pdl_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
}
export OLF_LIB=$PWD/tools/libolf_pdl.so
timeout 40 python -m pytest tests/test_gpu_line.py -x -q -m gpu 2>&1 | tail -2
bench p20 20
bench p1 1
