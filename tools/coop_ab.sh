#!/bin/bash
# A/B of the cooperative multi-round kernel on the few-slot figures of bench.py (C1 plug-in, C4, C2)
for v in coop nocoop; do
  if [ $v = nocoop ]; then export QB_NO_COOP=1; else unset QB_NO_COOP; fi
  python bench.py --no-cpu --steps 1 --warmup 3 --ntraj 256 --slots 256 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
m = d['mesolve']
print('$v', 'c2 rhs/s %.0f' % m['mesolve_rhs_evals_per_s'], 'c2mf rhs/s %.0f' % m['matrix_free']['mesolve_rhs_evals_per_s'],
      'c4 ms %.2f' % d['td_mesolve_c4']['gpu_ms'], 'c1 plugin ms %.2f / per-call %.2f' % (d['plugin_c1']['b200_vern7_wall_ms'], d['plugin_c1']['b200_vern7_one_call_per_time_wall_ms']),
      'c2 plugin wall %.3f' % d['plugin_matrix_form_c2']['b200_vern7']['wall_s'])
"
done
