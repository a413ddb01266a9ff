#!/bin/bash
# e2e variants: warm-up length and DataLoader workers (host-side sensitivity of the b200mm.train() e2e number)
for w in 2 12; do for k in 0 2 4 8; do
  echo "warm=$w workers=$k"
  B200MM_E2E_WARM=$w B200MM_E2E_WORKERS=$k python bench.py --no-cpu-baseline --steps 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['e2e']['ms_per_step'])
"
done; done
