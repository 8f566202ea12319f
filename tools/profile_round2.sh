# round-2 evidence run (one B200): bench lines of every stated config, ncu launch list + full captures, per-C' table
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python bench.py > $O/r2_bench_n1_jumpcp.json 2> $O/r2_bench_n1_jumpcp.err || exit 1
timeout 200 python tools/per_cprime_graphs.py > $O/r2_per_cprime.txt 2> $O/r2_per_cprime.err
for wl in chammi so2sat vitb; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-eager > $O/r2_bench_n1_$wl.json 2> $O/r2_bench_n1_$wl.err
done
timeout 200 python bench.py --workload vitb --mode eval > $O/r2_bench_n1_vitb_eval.json 2> $O/r2_bench_n1_vitb_eval.err
DCV_GRAPHS=0 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > $O/r2_plain.log 2>&1 || exit 1
DCV_GRAPHS=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > $O/r2_ncu_launch.log 2>&1
timeout 100 python tools/one_attn.py 32 1569 6 > $O/r2_one_attn.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -f -o $O/r2_attn_bwd python tools/one_attn.py 32 1569 6 > $O/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -f -o $O/r2_attn_fwd python tools/one_attn.py 32 1569 6 > $O/r2_ncu_b.log 2>&1
ls -la $O/*.ncu-rep
