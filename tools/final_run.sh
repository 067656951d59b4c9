set -x
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; tail -4 $O/gpu_tests.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench.json 2> $O/bench.err
tail -c 300 $O/bench.json
python bench.py --config C3 --scheme anneal:i=5k,a=20 --steps 2 --warmup 1 --no-shard-kir --no-kir-prefilter --no-t-sweep > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --config C5 --steps 3 --warmup 2 --no-shard-kir --no-kir-prefilter --no-t-sweep > $O/bench_c5.json 2> $O/bench_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-shard-kir --no-kir-prefilter > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_solve_stage -s 1 -c 1 -f -o $O/solve_final python tools/profile_run.py > $O/solve_final_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"k_recruit_short|k_collect_read_ends|k_pair_groups" -c 6 -f -o $O/widen_final python tools/recruit_run.py --pairs 200000 > $O/widen_ncu.log 2>&1
python tools/pairs_run.py > $O/pairs_run.log 2>&1; tail -2 $O/pairs_run.log
python tools/rescore_run.py > $O/rescore_run.log 2>&1; tail -2 $O/rescore_run.log
python tools/recruit_run.py > $O/recruit_run.log 2>&1; tail -2 $O/recruit_run.log
python tools/prefilter_run.py --config C4 > $O/prefilter_c4.log 2>&1; tail -1 $O/prefilter_c4.log
