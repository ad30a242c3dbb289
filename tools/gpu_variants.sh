#!/bin/bash
# Developer tool: times the in-tree library and every variant under vulkan_renderer_b200/variants on the bench scene.
mkdir -p gpurun_out
args=${@:-64 8 1 3}
echo "== in-tree"; timeout 300 python tools/quick_time.py $args 2>&1 | tail -3
for lib in vulkan_renderer_b200/variants/libvkr_*.so; do
	echo "== $lib"; VKR_B200_LIB=$PWD/$lib timeout 300 python tools/quick_time.py $args 2>&1 | tail -3
done
