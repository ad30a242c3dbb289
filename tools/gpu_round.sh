#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list + one full capture of the shading kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
tag=${1:-r01}
mkdir -p gpurun_out
if [ -z "$2" ]; then
	timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log
	tail -3 gpurun_out/${tag}_tests.log
fi
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -1 gpurun_out/${tag}_bench.json | cut -c1-400
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -1 gpurun_out/${tag}_bench_ref.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shading_kernel -c 1 -o gpurun_out/${tag}_full -f python tools/quick_time.py 64 8 1 3 > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out | tail -8
