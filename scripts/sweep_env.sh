#!/bin/bash
# usage: scripts/sweep_env.sh OUT.txt "VAR=a VAR2=b" "VAR=c" ...   -- one short bench run per environment setting, one summary line each
out=$1; shift
: > $out
for setting in "$@"; do
  env $setting python bench.py --no-rows --cpu-frames 4 --steps 6 --warmup 3 2>/dev/null | grep '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$setting', 'value=%.0f'%d['value'], 'ms=%.2f'%d['ms_per_step'], 'e2e=%.0f'%d['e2e']['value'], 'serial=%.2f'%d['roofline'].get('serial_step_ms',0), ' '.join('%s=%.2f'%(n,k[n]['avg_ms']) for n in ('kht_link','canny_front','kht_peaks_sort','canny_finalize') if n in k))
" >> $out
done
cat $out
