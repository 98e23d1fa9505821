#!/bin/bash
# sweeps the sub-batch stagger stage and the number of parts (bench, device-resident + e2e)
for cfg in "2 -1" "2 0" "2 1" "2 3" "2 4" "2 5" "3 1" "3 4" "4 1" "4 4"; do
  set -- $cfg
  echo "== parts $1 stagger $2"
  JN_ELAS_SPLIT=$1 JN_ELAS_STAGGER=$2 python bench.py --steps 15 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('  value %.0f  e2e %.0f  ms/step %.2f  ok %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['frames_ok']))"
done
