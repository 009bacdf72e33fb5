#!/bin/bash
# The round's measurement pass on ONE B200 (run through gpurun): GPU test suite, smoke, both bench arms, the size and kernel
# tables, then - each only after its command has exited 0 without a profiler - the ncu launch list of the bench command and
# one full capture of its dominant kernel.  Everything lands in gpurun_out/ under the given tag.
# usage: tools/measure_all.sh TAG
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $O/${TAG}_smoke.log
python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || exit 1
python tools/perf_playout.py --games 65536 1048576 16777216 > $O/${TAG}_sizes.jsonl 2>> $O/${TAG}_sizes.err
python tools/perf_playout.py --games 1048576 --players 3 >> $O/${TAG}_sizes.jsonl 2>> $O/${TAG}_sizes.err
python tools/perf_playout.py --games 65536 1048576 --players 4 >> $O/${TAG}_sizes.jsonl 2>> $O/${TAG}_sizes.err
python tools/perf_kernels.py --envs 262144 1048576 > $O/${TAG}_kernels.jsonl 2> $O/${TAG}_kernels.err
python tools/perf_e2e.py > $O/${TAG}_e2e.json 2> $O/${TAG}_e2e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > $O/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:playout_pool -s 3 -c 1 -o $O/prof_${TAG}_bench -f python tools/perf_playout.py --games 65536 > $O/${TAG}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:playout_pool -s 3 -c 1 -o $O/prof_${TAG}_big -f python tools/perf_playout.py --games 1048576 > $O/${TAG}_ncu_b.log 2>&1
for k in legal_mask observe step_kernel env_move; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $O/prof_${TAG}_$k -f python tools/perf_kernels.py --only legal_mask observe step env_step > $O/${TAG}_ncu_$k.log 2>&1
done
echo done > $O/${TAG}_done
