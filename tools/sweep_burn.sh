#!/bin/bash
# needs tools/libburn.so: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/libburn.so tools/burn.cu -lcudart
# what does the front end's throughput depend on?  bench.py at 20 rigs while tools/burn.cu takes (a) issue slots: 1..4 dependent integer chains on one
# warp per scheduler, (b) memory requests: 100 + w = w warps per SM chasing random sectors
run() { tag=$1; shift; env $ENVV python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines 20 > gpurun_out/sweep_$tag.log 2>gpurun_out/sweep_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'call_p50', c['steadiness']['resident']['call_ms']['p50'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
grep -h "burn:" gpurun_out/sweep_$tag.err
}
for m in "$@"; do ENVV="OLF_BENCH_BURN=$m,25"; run burn$m; done
