set -x
python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > gpurun_out/plain_r1c.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-eager > gpurun_out/ncu_r1c.log 2>&1
python tools/one_attn.py 32 1569 6 > gpurun_out/one_attn2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -f -o gpurun_out/attn_bwd_r1c python tools/one_attn.py 32 1569 6 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 1 -c 1 -f -o gpurun_out/attn_fwd_r1c python tools/one_attn.py 32 1569 6 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 4 -c 3 -f -o gpurun_out/gemm_r1c python tools/probe_block_gemms.py > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
