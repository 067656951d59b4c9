#!/bin/bash
for lib in liblctp.so liblctp_eager.so; do
echo "$lib C2 i=5k T=4736 : $(LCTP_LIB=$PWD/locityper_b200/_lib/$lib python tools/profile_run.py --passes 3 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
echo "$lib C2 i=20k T=2960: $(LCTP_LIB=$PWD/locityper_b200/_lib/$lib python tools/profile_run.py --passes 3 --threads 2960 --scheme greedy:i=20k,a=1 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
done
