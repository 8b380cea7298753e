#!/bin/bash
# usage: profiles/sweep_variants.sh "<variant names>" "<threads list>" "<occ list>"
mkdir -p gpurun_out; rm -f gpurun_out/sweepv.log
for v in $1; do for t in $2; do for o in $3; do
MTFB_LIB=$PWD/mtf_b200/csrc/_variants/lib$v.so timeout 300 python bench.py --steps 10 --warmup 3 --threads $t --occ $o --no-cpu 2>&1 | V=$v python -c "
import sys,json,os
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(os.environ['V'], 'T', d['config']['threads_per_patch'], 'occ', d['config']['occupancy'], 'ms/step %.3f'%d['ms_per_step'], 'Miter/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), d['valid'])
" >> gpurun_out/sweepv.log; done; done; done; cat gpurun_out/sweepv.log
