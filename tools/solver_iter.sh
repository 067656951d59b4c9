#!/bin/bash
# one GPU round trip of the solver tuning loop: timing (solo, 3 loci in flight), then an ncu capture named $1
tag=${1:-x}
echo "solo: $(python tools/profile_run.py --passes 3 | sed -e 's/.*stage_ms.: \([0-9.]*\).*/stage_ms(3 passes)=\1/')"
python tools/concurrency_probe.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_solve_stage -s 1 -c 1 -f -o gpurun_out/solve_$tag python tools/profile_run.py > gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_ncu.log | cut -c1-80
