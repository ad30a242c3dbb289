#!/bin/bash
# Developer tool for tuning experiments: builds the library with extra -D flags for the shading kernel only.
#   tools/build_variant.sh <name> "<nvcc flags>"   ->  vulkan_renderer_b200/variants/libvkr_<name>.so
# Select it with VKR_B200_LIB=<path> (tools/quick_time.py only; tests and bench always load the in-tree libvkr_b200.so).
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=vulkan_renderer_b200/variants; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
	-ccbin /usr/bin/g++ -Xcompiler -fPIC -I include -DVKR_MAXP_TU=5 $flags -c vulkan_renderer_b200/csrc/vkr_shading_kernel.cu -o $out/$name.o
b=vulkan_renderer_b200/build   # only the quad-light kernels (vertex bound 5) are rebuilt; everything else comes from the in-tree objects
others=$(ls $b/*.o | grep -v vkr_shading_kernel_maxp5)
nvcc -shared -o $out/libvkr_$name.so $out/$name.o $others -ccbin /usr/bin/g++ -Xcompiler -fopenmp -lgomp -cudart static
rm $out/$name.o
echo built $out/libvkr_$name.so
