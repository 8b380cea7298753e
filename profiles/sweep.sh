#!/bin/bash
# usage: profiles/sweep.sh "<threads list>" "<occ list>"  -> gpurun_out/sweep.log
mkdir -p gpurun_out; rm -f gpurun_out/sweep.log
for t in $1; do for o in $2; do
timeout 300 python bench.py --steps 10 --warmup 3 --threads $t --occ $o --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('T', d['config']['threads_per_patch'], 'occ', d['config']['occupancy'], 'ms/step %.3f'%d['ms_per_step'], 'Miter/s %.2f'%(d['value']/1e6), 'e2e %.2f'%(d['e2e']['value']/1e6), 'frac %.4f'%d['roofline']['frac'], d['valid'])
" >> gpurun_out/sweep.log; done; done; cat gpurun_out/sweep.log
