#!/bin/bash
# One GPU-box visit of round 2 (under gpurun): bash tools/gpu_round2.sh <tag> [stage ...]
#   stages: variants tests bench ref c2 c4 matrix builder launches fullc2 full   (default: all but c4 and fullc2)
tag=${1:-r02}; shift
stages=${@:-variants tests bench ref c2 matrix builder launches full}
mkdir -p gpurun_out
has() { [[ " $stages " == *" $1 "* ]]; }
if has variants; then bash tools/gpu_explore.sh ${tag} 2>&1 | grep -v "^Triangle"; fi
if has tests; then timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log; fi
if has bench; then timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -1 gpurun_out/${tag}_bench.json | cut -c1-600; tail -3 gpurun_out/${tag}_bench.err; fi
if has ref; then timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; tail -1 gpurun_out/${tag}_bench_ref.json | cut -c1-400; fi
if has c2; then timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/${tag}_c2_bench.json 2> gpurun_out/${tag}_c2_bench.err; tail -1 gpurun_out/${tag}_c2_bench.json | cut -c1-600; tail -2 gpurun_out/${tag}_c2_bench.err; fi
if has c4; then timeout 1200 python bench.py --workload C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_c4_bench_n1.json 2> gpurun_out/${tag}_c4_bench_n1.err; tail -1 gpurun_out/${tag}_c4_bench_n1.json | cut -c1-600; tail -2 gpurun_out/${tag}_c4_bench_n1.err; fi
if has matrix; then timeout 1500 python tools/run_experiments.py --no-figs --no-screenshots --width 1920 --height 1080 --out gpurun_out/${tag}_experiments > gpurun_out/${tag}_experiments.log 2>&1; tail -3 gpurun_out/${tag}_experiments.log; fi
if has builder; then
	for builder in sah lbvh_gpu; do echo "== VKR_BVH_BUILDER=$builder"; VKR_COUNTERS=1 VKR_BVH_BUILDER=$builder timeout 600 python tools/quick_time.py 64 8 1 3 2>&1 | tail -7 | tee gpurun_out/${tag}_builder_$builder.log; done
fi
if has launches; then timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-counters > gpurun_out/${tag}_launches_bench.log 2>&1; tail -2 gpurun_out/${tag}_launches_bench.log | cut -c1-200; fi
if has fullc2; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:shading_kernel -c 1 -o gpurun_out/${tag}_c2_full -f python tools/quick_time.py 4 1 1 0 > gpurun_out/${tag}_c2_full.log 2>&1; ls -la gpurun_out/${tag}_c2_full.ncu-rep; fi
if has full; then timeout 900 ncu --set full --clock-control none --import-source on -k regex:shading_kernel -c 1 -o gpurun_out/${tag}_full -f python tools/quick_time.py 64 8 1 3 > gpurun_out/${tag}_full.log 2>&1; ls -la gpurun_out/${tag}_full.ncu-rep; fi
ls gpurun_out | wc -l
