#!/bin/bash
# Developer tool, one GPU-box visit of tuning: the in-tree library and every variant under vulkan_renderer_b200/variants on the bench scene
# (frame hash first: variants must reproduce the frame bit for bit), trace counters, optionally an ncu capture of the in-tree kernel.
#   usage (under gpurun): bash tools/gpu_explore.sh <tag> [ncu]
tag=${1:-explore}
mkdir -p gpurun_out
args="64 8 1 3"
echo "== in-tree"; VKR_COUNTERS=1 timeout 600 python tools/quick_time.py $args 2>&1 | tail -6 | tee gpurun_out/${tag}_intree.log
for lib in vulkan_renderer_b200/variants/libvkr_*.so; do
	[ -f "$lib" ] || continue
	name=$(basename $lib .so); name=${name#libvkr_}
	width=2; case $name in bvh4*) width=4;; esac
	echo "== $name"; VKR_COUNTERS=1 VKR_BVH_WIDTH=$width VKR_B200_LIB=$PWD/$lib timeout 600 python tools/quick_time.py $args 2>&1 | tail -6 | tee gpurun_out/${tag}_$name.log
done
if [ -n "$2" ]; then
	timeout 900 ncu --set full --clock-control none --import-source on -k regex:shading_kernel -c 1 -o gpurun_out/${tag}_full -f python tools/quick_time.py $args > gpurun_out/${tag}_full.log 2>&1
	python tools/summarize_ncu.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_summary.md 2>gpurun_out/${tag}_summary.err
	ls -la gpurun_out/${tag}_full.ncu-rep
fi
