#!/bin/bash
# A/B of experiment knobs on the default workload (tuning library): prints ms/step, images/s, conv / HBM family ms
#   scripts/ab_bench.sh "VAR1=a VAR2=b" "VAR1=c" ...
export TORTTO_B200_LIB=${TORTTO_B200_LIB:-tuning}
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --cpu-baseline 0 --extra-bf16 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('ms/step %.4f  img/s %.0f  e2e %.0f  conv_ms %.3f  hbm_ms %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['hbm']['family_ms_per_step']))
fam = d['family_ms_per_step']['by_entry_point']
print('   ' + '  '.join('%s %.3f' % (k.replace('ttb_', ''), v) for k, v in list(fam.items())[:9]))"
done
