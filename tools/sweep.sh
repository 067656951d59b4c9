#!/bin/bash
# tuning sweep of the solver kernel's residency knobs (one C2 locus, 3 passes each)
for cfg in "0 0" "100 4" "75 3" "60 3" "50 2" "40 2"; do
  set -- $cfg
  if [ "$1" != "0" ]; then export LCTP_CARVEOUT=$1 LCTP_MAX_CTAS_PER_SM=$2; fi
  echo "carveout=$1 ctas=$2: $(python tools/profile_run.py --passes 3 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
done
