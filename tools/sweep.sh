#!/bin/bash
# prefilter tile variants (LCTP_PREFILTER_VARIANT) at the BASELINE shapes
for cfg in C4 C5 C2; do for v in 1 11 12 13 14 15; do
  echo "$cfg variant=$v: $(LCTP_PREFILTER_VARIANT=$v timeout 120 python tools/prefilter_run.py --config $cfg | tail -1)"
done; done
