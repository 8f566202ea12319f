# two GPUs: NCCL data-parallel tests (gradients vs one-GPU full batch, graphed DP vs eager DP) + the N = 2 bench line
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q > $O/r2f_dp_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_dp_pytest.log
tail -n 4 $O/r2f_dp_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2f_bench_n2_jumpcp.json 2> $O/r2f_bench_n2_jumpcp.err; echo "bench rc=$?"
tail -c 400 $O/r2f_bench_n2_jumpcp.json
