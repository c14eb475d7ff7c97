#!/bin/bash
# A/B of the density kernel configurations on the bench workload: prints the density time and the posterior checksum
for cfg in ${CFGS:-0 1 2 3 4 5 6}; do
 for g in ${GS:-0}; do
  NEMB200_DENS_G=$g NEMB200_DENS_CFG=$cfg python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.readlines()[-1])
r=j['roofline']
print('cfg $cfg G $g', 'value', round(j['value'],1), 'dens_ms', round(r['stage_ms']['density (k_density_sk)'],4), 'onehot_ms', round(r['kernels']['k_mstep_onehot']['ms'],4), 'checksum', j['check']['posterior_checksum'], 'clk', j['clocks']['sm_mhz'])
"
 done
done
