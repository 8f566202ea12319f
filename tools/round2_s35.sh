# one GPU call: full GPU suite at HEAD (fused embed v2 + backward side branch on by default), A/B of the side branch,
# timelines of both fused patch-embedding kernels
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/s35_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s35_pytest.log
tail -n 5 $O/s35_pytest.log
timeout 300 python tools/overlap_ab.py > $O/s35_overlap_ab.txt 2> $O/s35_overlap_ab.err; echo "rc=$?"
cat $O/s35_overlap_ab.txt; tail -n 5 $O/s35_overlap_ab.err
EF_MODE=1 timeout 100 python tools/embed_timeline.py > $O/s35_embed_timeline_m1.txt 2>&1
EF_MODE=2 timeout 100 python tools/embed_timeline.py > $O/s35_embed_timeline_m2.txt 2>&1
