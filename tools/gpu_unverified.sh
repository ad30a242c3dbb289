#!/bin/bash
# First GPU-box visit of the next round: everything that was written after the GPU budget of round 1 was spent, one log per item, no -x
# (a failure in one file must not hide the others).  usage (under gpurun): bash tools/gpu_unverified.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
for t in test_gpu_zzw_fuzz test_gpu_zzx_textured_lights test_gpu_zz_error_display test_gpu_zzy_textured_gbuffer test_gpu_zzz_lbvh_gpu; do
	timeout 600 python -m pytest tests/$t.py -m gpu -q > gpurun_out/${tag}_$t.log 2>&1; echo "exit $?" >> gpurun_out/${tag}_$t.log
	echo "== $t: $(tail -2 gpurun_out/${tag}_$t.log | tr '\n' ' ')"
done
# the GPU BVH builder on the benchmark scene: build time and the cost of its trees for the shading pass (expected: tens of ms; 20-40 % more node visits than SAH)
for builder in sah lbvh_gpu; do
	echo "== VKR_BVH_BUILDER=$builder"; VKR_BVH_BUILDER=$builder timeout 600 python tools/quick_time.py 64 8 1 3 2>&1 | tail -3 | tee gpurun_out/${tag}_builder_$builder.log
done
# the 4-wide traversal variant, if it was built here beforehand (tools/build_variant.sh bvh4 "-DVKR_BVH_WIDTH=4"): same frame first, then time
if [ -f vulkan_renderer_b200/variants/libvkr_bvh4.so ]; then
	echo "== bvh4 variant"; VKR_BVH_WIDTH=4 VKR_B200_LIB=$PWD/vulkan_renderer_b200/variants/libvkr_bvh4.so timeout 600 python tools/quick_time.py 64 8 1 3 2>&1 | tail -3 | tee gpurun_out/${tag}_bvh4.log
fi
ls -la gpurun_out | tail -12
