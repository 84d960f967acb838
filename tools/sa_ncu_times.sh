#!/usr/bin/env bash
# durations of the four fused SA launches of one steady-state pass under ncu (cold, serialised): quick A/B of csrc/sa_fused.cu
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"${1:-sa_mlp_fused}" --csv \
    python tools/prof_two_phase.py 1 2>/dev/null | python -c "
import csv, sys
rows = [r for r in csv.reader(l for l in sys.stdin if l.startswith('\"'))]
h = rows[0]
print(' '.join(f'{float(r[h.index(\"Metric Value\")].replace(\",\", \"\")) / 1e3:.1f}' for r in rows[1:]), 'us')
"
