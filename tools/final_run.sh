set -x
# Round-end measurement pass (one GPU).  The two ncu --set full captures the bench lines quote (C2 greedy launch, C3
# annealing launch; ~10 min of box time, the annealing launch is 0.75 s per replay) are taken separately:
#   ncu --set full --clock-control none --import-source on -k regex:k_solve_stage -s 1 -c 1 -f -o gpurun_out/solve_s3 python tools/profile_run.py
#   ncu ... -o gpurun_out/anneal_s3 python tools/profile_run.py --config C3 --scheme anneal:i=5k,a=20
# and summarised here with tools/ncu_summary.py into profiles/r02_ncu_traffic.json BEFORE this script runs.
O=gpurun_out/final; mkdir -p $O
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.json
python bench.py --config C3 --scheme anneal:i=5k,a=20 --steps 2 --warmup 1 --no-shard-kir --no-kir-prefilter --no-t-sweep > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --config C5 --steps 3 --warmup 2 --no-shard-kir --no-kir-prefilter --no-t-sweep > $O/bench_c5.json 2> $O/bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-shard-kir --no-kir-prefilter --no-t-sweep > $O/bench_under_ncu.log 2>&1
python tools/pairs_run.py > $O/pairs_run.log 2>&1; tail -2 $O/pairs_run.log
python tools/prefilter_run.py --config C4 > $O/prefilter_c4.log 2>&1; tail -1 $O/prefilter_c4.log
