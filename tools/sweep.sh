#!/bin/bash
# tuning sweep: launch-bounds variants of the solver kernel (one C2 locus, 3 passes each)
for lib in liblctp.so liblctp_mb5.so liblctp_mb6.so liblctp_mb7.so; do
  export LCTP_LIB=$PWD/locityper_b200/_lib/$lib
  echo "$lib: $(python tools/profile_run.py --passes 3 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
done
