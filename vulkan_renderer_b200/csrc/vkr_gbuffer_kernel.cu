// vkr_gbuffer_kernel.cu -- producer of the shading pass's inputs.
//
// (1) visibility kernel: stand-in for the reference's rasterised visibility buffer
//     (src/shaders/visibility_pass.vert.glsl:27-33, src/main.c:1422-1427): closest hit of the
//     primary ray through each pixel centre, against the shader-decoded (fma) vertices.
// (2) gbuffer kernel: the reference's get_shading_data() (src/shaders/shading_pass.frag.glsl:721-822)
//     writing the 64 B/pixel G-buffer the shading megakernel reads. Materials whose textures are constant take one
//     texel (TEXTURED = false); otherwise the three material textures are filtered with screen-space derivatives as the
//     shader does (textureGrad at :779-783, filter definition in vkr_texture.cuh).
// A visibility buffer produced elsewhere (e.g. by a rasteriser) can be fed to (2) directly.
// Compile with -fmad=false.
#include "vkr_gbuffer.cuh"

namespace vkr {

constexpr int kGTileW = 16, kGTileH = 8, kGThreads = kGTileW * kGTileH;

__global__ void __launch_bounds__(kGThreads) visibility_kernel(const gbuffer_kernel_params p) {
	__shared__ int stack[kMaxStackDepth * kGThreads];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
	const int tiles_x = (p.width + kGTileW - 1) / kGTileW;
	const int x = (blockIdx.x % tiles_x) * kGTileW + lx;
	const int y = (blockIdx.x / tiles_x) * kGTileH + ly;
	if (x >= p.width || y >= p.height) return;
	const f3 camera = make3(gldf(p.constants, G_OFF_CAMERA), gldf(p.constants, G_OFF_CAMERA + 4), gldf(p.constants, G_OFF_CAMERA + 8));
	const f3 d = pixel_ray(p.constants, x, y);
	bvh_view bvh; bvh.nodes = p.bvh_nodes; bvh.tris = p.bvh_tris; bvh.tri_ids = p.bvh_tri_ids; bvh.tri_count = p.tri_count;
	int hit = -1;
	if (p.tri_count) hit = closest_hit(bvh, camera, d, 0.0f, __int_as_float(0x7f800000), stack + threadIdx.x, kGThreads);
	p.visibility[(size_t) y * p.width + x] = (hit < 0) ? 0xFFFFFFFFu : (uint32_t) hit;
}

template <bool TEXTURED>
__global__ void __launch_bounds__(128) gbuffer_kernel(const gbuffer_kernel_params p) {
	const size_t pixel = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (pixel < (size_t) p.width * p.height) shade_gbuffer_pixel<TEXTURED>(p, pixel);
}

} // namespace vkr

cudaError_t vkr_launch_visibility_kernel(const vkr::gbuffer_kernel_params& p, cudaStream_t stream) {
	const int tiles = ((p.width + vkr::kGTileW - 1) / vkr::kGTileW) * ((p.height + vkr::kGTileH - 1) / vkr::kGTileH);
	if (tiles <= 0) return cudaSuccess;
	vkr::visibility_kernel<<<tiles, vkr::kGThreads, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t vkr_launch_gbuffer_kernel(const vkr::gbuffer_kernel_params& p, cudaStream_t stream) {
	const size_t pixels = (size_t) p.width * p.height;
	if (!pixels) return cudaSuccess;
	if (p.texture_data) vkr::gbuffer_kernel<true><<<(unsigned) ((pixels + 127) / 128), 128, 0, stream>>>(p);
	else vkr::gbuffer_kernel<false><<<(unsigned) ((pixels + 127) / 128), 128, 0, stream>>>(p);
	return cudaGetLastError();
}
