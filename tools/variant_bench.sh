#!/bin/bash
# Runs the device-resident bench leg for the default library and every experimental build under kalibr_b200/_exp/.
for lib in "" kalibr_b200/_exp/*.so; do
  if [ -n "$lib" ]; then export KB_LIB_PATH=$PWD/$lib; else unset KB_LIB_PATH; fi
  echo "== ${lib:-default}"
  python bench.py --steps 10 --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step %.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['stage_ms'].items()}, 'mat %.3f' % d['linearise_materialised']['ms'])"
done
