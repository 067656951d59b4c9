#!/bin/bash
echo "C2 i=5k T=4736 : $(python tools/profile_run.py --passes 3 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
echo "C2 i=20k T=2960: $(python tools/profile_run.py --passes 3 --threads 2960 --scheme greedy:i=20k,a=1 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
for cfg in C4 C5 C2; do echo "$cfg: $(python tools/prefilter_run.py --config $cfg | tail -1)"; done
