#!/bin/bash
# throughput / latency of library variants (diagnostic; libraries in orb_line_slam_b200/_variants): usage sweep_variants.sh name...
run() { tag=$1; P=$2; shift; shift; env "$@" python bench.py --steps 10 --warmup 3 --prewarm-steps 20 --no-cpu-baseline --pipelines $P > gpurun_out/sweep_$tag.log 2>/dev/null
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
l=[x for x in open(f'gpurun_out/sweep_{tag}.log') if x.startswith('{')]
if not l: print(tag,'FAILED'); sys.exit()
d=json.loads(l[-1]); c=d['config']
print(tag, 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'rig_call_ms', c['rig_call_ms']['total'], 'chain_ms', c['line_call_ms']['enqueue_and_chain'], 'kernel_ms/img', round(d['roofline']['kernel_ms'],2))
PY
}
python -m pytest tests/test_gpu_line.py tests/test_frontend_golden.py -x -q -m gpu 2>&1 | tail -2
V=orb_line_slam_b200/_variants
run base_p1 1 A=1; run base_p20 20 A=1
for v in "$@"; do run ${v}_p1 1 OLF_LIB=$V/libolf_$v.so; run ${v}_p20 20 OLF_LIB=$V/libolf_$v.so; done
run base_p20b 20 A=1
