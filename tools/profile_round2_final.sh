# round-2 final evidence run at HEAD (one B200): bench lines of every stated config, ncu launch list of the bench
# command, full ncu capture of the fused patch-embedding kernel, per-C' table
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python bench.py > $O/r2f_bench_n1_jumpcp.json 2> $O/r2f_bench_n1_jumpcp.err || exit 1
tail -c 600 $O/r2f_bench_n1_jumpcp.json
DCV_GRAPHS=0 timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > $O/r2f_plain.log 2>&1 || exit 1
DCV_GRAPHS=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > $O/r2f_ncu_launch.log 2>&1
timeout 100 python tools/one_embed.py > $O/r2f_one_embed.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_fused_v2_kernel -s 1 -c 1 -f -o $O/r2f_embed_fused_v2 python tools/one_embed.py > $O/r2f_ncu_e.log 2>&1
for wl in chammi so2sat vitb; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-eager > $O/r2f_bench_n1_$wl.json 2> $O/r2f_bench_n1_$wl.err
done
timeout 200 python bench.py --workload vitb --mode eval > $O/r2f_bench_n1_vitb_eval.json 2> $O/r2f_bench_n1_vitb_eval.err
timeout 200 python tools/per_cprime_graphs.py > $O/r2f_per_cprime.txt 2> $O/r2f_per_cprime.err
ls -la $O/r2f_*
