/* shim/vkr_shim.c -- host-memory implementation of the Vulkan / vulkan_basics entry points that the reference's
 * loader files call (boundary B1). Compiled against the reference's UNCHANGED src/vulkan_basics.h.
 *
 * Observable semantics kept: all buffers of one create_*buffers call live in ONE allocation addressed by
 * buffers[i].offset after vkMapMemory(memory, 0, size) (src/vulkan_basics.c:712-722; scene.c:462-474 reads the *.vks
 * payload straight into staging + offset); copy_buffers / copy_buffers_to_images are the upload hooks;
 * vkCmdBuildAccelerationStructuresKHR is the BVH hook (it receives the dequantised float[3] soup, scene.c:197-209).
 * Everything is plain malloc'ed memory; the CUDA library picks the bytes up through the vkr_shim_* accessors.
 */
#include "vulkan_basics.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct VkDeviceMemory_T { char* data; VkDeviceSize size; };
struct VkBuffer_T { struct VkDeviceMemory_T* memory; VkDeviceSize offset, size; };
struct VkBufferView_T { VkBuffer buffer; VkFormat format; };
struct VkSampler_T { VkSamplerCreateInfo info; };
struct VkCommandBuffer_T { int recording; };
struct VkImage_T { VkImageCreateInfo info; VkDeviceSize* level_offsets; VkDeviceSize layer_size; char* data; };
struct VkImageView_T { VkImage image; };
struct VkAccelerationStructureKHR_T { VkAccelerationStructureTypeKHR type; float* vertices; uint64_t triangle_count; };
struct VkDevice_T { int dummy; };
struct VkInstance_T { int dummy; };

static struct VkDevice_T g_device;
static struct VkInstance_T g_instance;
VkDevice vkr_shim_device(void) { return &g_device; }
VkInstance vkr_shim_instance(void) { return &g_instance; }

/* bytes per 4x4 block (compressed) or per texel */
static VkDeviceSize format_block_bytes(VkFormat format, uint32_t* block_dim) {
	*block_dim = 1;
	switch (format) {
	case VK_FORMAT_R8_UINT: return 1;
	case VK_FORMAT_R8G8_SINT: return 2;
	case VK_FORMAT_R8G8B8A8_UNORM: case VK_FORMAT_R8G8B8A8_SRGB: case VK_FORMAT_B8G8R8A8_SRGB: case VK_FORMAT_R16G16_UNORM: case VK_FORMAT_R32_UINT: return 4;
	case VK_FORMAT_R16G16B16_SFLOAT: return 6;
	case VK_FORMAT_R16G16B16A16_UNORM: case VK_FORMAT_R16G16B16A16_SFLOAT: case VK_FORMAT_R32G32_UINT: return 8;
	case VK_FORMAT_R32G32B32_SFLOAT: return 12;
	case VK_FORMAT_R32G32B32A32_SFLOAT: return 16;
	case VK_FORMAT_BC1_RGB_UNORM_BLOCK: case VK_FORMAT_BC1_RGB_SRGB_BLOCK: *block_dim = 4; return 8;
	case VK_FORMAT_BC5_UNORM_BLOCK: *block_dim = 4; return 16;
	default: return 0;
	}
}
static VkDeviceSize level_bytes(const VkImageCreateInfo* info, uint32_t level) {
	uint32_t bd; VkDeviceSize bb = format_block_bytes(info->format, &bd);
	uint32_t w = info->extent.width >> level, h = info->extent.height >> level, d = info->extent.depth >> level;
	if (!w) w = 1; if (!h) h = 1; if (!d) d = 1;
	return bb * ((w + bd - 1) / bd) * ((h + bd - 1) / bd) * d;
}

/* ---- vulkan_basics.h functions (src/vulkan_basics.h:375-404, 417-424) ---- */
int create_aligned_buffers(buffers_t* buffers, const device_t* device, const VkBufferCreateInfo* buffer_infos, uint32_t buffer_count, VkMemoryPropertyFlags memory_properties, VkDeviceSize alignment) {
	(void) device; (void) memory_properties;
	memset(buffers, 0, sizeof(*buffers));
	if (alignment < 16) alignment = 16;
	buffers->buffer_count = buffer_count;
	buffers->buffers = (buffer_t*) calloc(buffer_count ? buffer_count : 1, sizeof(buffer_t));
	VkDeviceSize offset = 0;
	for (uint32_t i = 0; i != buffer_count; ++i) {
		offset = (offset + alignment - 1) / alignment * alignment;
		buffers->buffers[i].offset = offset; buffers->buffers[i].size = buffer_infos[i].size;
		offset += buffer_infos[i].size;
	}
	buffers->size = offset ? offset : 16;
	struct VkDeviceMemory_T* memory = (struct VkDeviceMemory_T*) calloc(1, sizeof(*memory));
	memory->data = (char*) calloc(1, buffers->size); memory->size = buffers->size;
	if (!memory->data) { free(memory); free(buffers->buffers); memset(buffers, 0, sizeof(*buffers)); return 1; }
	buffers->memory = memory;
	for (uint32_t i = 0; i != buffer_count; ++i) {
		struct VkBuffer_T* b = (struct VkBuffer_T*) calloc(1, sizeof(*b));
		b->memory = memory; b->offset = buffers->buffers[i].offset; b->size = buffers->buffers[i].size;
		buffers->buffers[i].buffer = b;
	}
	return 0;
}

void destroy_buffers(buffers_t* buffers, const device_t* device) {
	(void) device;
	for (uint32_t i = 0; i != buffers->buffer_count; ++i) if (buffers->buffers && buffers->buffers[i].buffer) free(buffers->buffers[i].buffer);
	if (buffers->memory) { free(buffers->memory->data); free(buffers->memory); }
	free(buffers->buffers);
	memset(buffers, 0, sizeof(*buffers));
}

int create_images(images_t* images, const device_t* device, const image_request_t* requests, uint32_t image_count, VkMemoryPropertyFlags memory_properties) {
	(void) device;
	memset(images, 0, sizeof(*images));
	images->image_count = image_count; images->memory_properties = memory_properties;
	images->images = (image_t*) calloc(image_count ? image_count : 1, sizeof(image_t));
	for (uint32_t i = 0; i != image_count; ++i) {
		image_t* im = &images->images[i];
		im->image_info = requests[i].image_info; im->view_info = requests[i].view_info;
		if (im->image_info.mipLevels == 0) im->image_info.mipLevels = get_mipmap_count_3d(im->image_info.extent);
		uint32_t bd; if (!format_block_bytes(im->image_info.format, &bd)) { printf("The Vulkan shim does not know VkFormat %d.\n", (int) im->image_info.format); destroy_images(images, device); return 1; }
		struct VkImage_T* image = (struct VkImage_T*) calloc(1, sizeof(*image));
		image->info = im->image_info;
		image->level_offsets = (VkDeviceSize*) calloc(image->info.mipLevels + 1, sizeof(VkDeviceSize));
		for (uint32_t l = 0; l != image->info.mipLevels; ++l) image->level_offsets[l + 1] = image->level_offsets[l] + level_bytes(&image->info, l);
		image->layer_size = image->level_offsets[image->info.mipLevels];
		image->data = (char*) calloc(1, image->layer_size * (image->info.arrayLayers ? image->info.arrayLayers : 1));
		im->image = image; im->memory_size = image->layer_size * image->info.arrayLayers;
		if (im->view_info.sType == VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO) {
			struct VkImageView_T* view = (struct VkImageView_T*) calloc(1, sizeof(*view)); view->image = image; im->view = view;
			im->view_info.image = image; im->view_info.format = image->info.format;
			if (!im->view_info.subresourceRange.layerCount) im->view_info.subresourceRange.layerCount = image->info.arrayLayers;
			if (!im->view_info.subresourceRange.levelCount) im->view_info.subresourceRange.levelCount = image->info.mipLevels;
		}
	}
	return 0;
}

void destroy_images(images_t* images, const device_t* device) {
	(void) device;
	for (uint32_t i = 0; i != images->image_count; ++i) {
		if (!images->images) break;
		struct VkImage_T* image = images->images[i].image;
		if (image) { free(image->level_offsets); free(image->data); free(image); }
		free(images->images[i].view);
	}
	free(images->images); free(images->memories);
	memset(images, 0, sizeof(*images));
}

int copy_buffers_and_images(const device_t* device,
	uint32_t buffer_count, const VkBuffer* source_buffers, const VkBuffer* destination_buffers, VkBufferCopy* buffer_regions,
	uint32_t image_count, const VkImage* source_images, const VkImage* destination_images, VkImageLayout source_layout,
	VkImageLayout destination_layout_before, VkImageLayout destination_layout_after, VkImageCopy* image_regions,
	uint32_t buffer_to_image_count, const VkBuffer* image_source_buffers, const VkImage* buffer_destination_images,
	VkImageLayout buffer_destination_layout_before, VkImageLayout buffer_destination_layout_after, VkBufferImageCopy* buffer_to_image_regions)
{
	(void) device; (void) source_layout; (void) destination_layout_before; (void) destination_layout_after; (void) buffer_destination_layout_before; (void) buffer_destination_layout_after;
	(void) source_images; (void) destination_images; (void) image_regions;
	if (image_count) { printf("The Vulkan shim does not implement image to image copies.\n"); return 1; }
	for (uint32_t i = 0; i != buffer_count; ++i) {
		const struct VkBuffer_T* s = source_buffers[i]; struct VkBuffer_T* d = destination_buffers[i];
		if (buffer_regions[i].srcOffset + buffer_regions[i].size > s->size || buffer_regions[i].dstOffset + buffer_regions[i].size > d->size) return 1;
		memcpy(d->memory->data + d->offset + buffer_regions[i].dstOffset, s->memory->data + s->offset + buffer_regions[i].srcOffset, buffer_regions[i].size);
	}
	for (uint32_t i = 0; i != buffer_to_image_count; ++i) {
		const struct VkBuffer_T* s = image_source_buffers[i]; struct VkImage_T* d = buffer_destination_images[i];
		const VkBufferImageCopy* r = &buffer_to_image_regions[i];
		uint32_t level = r->imageSubresource.mipLevel;
		VkDeviceSize bytes = level_bytes(&d->info, level);
		for (uint32_t l = 0; l != r->imageSubresource.layerCount; ++l) {
			VkDeviceSize src = r->bufferOffset + bytes * l;
			if (src + bytes > s->size) return 1;
			memcpy(d->data + d->layer_size * (r->imageSubresource.baseArrayLayer + l) + d->level_offsets[level], s->memory->data + s->offset + src, bytes);
		}
	}
	return 0;
}

/* ---- core Vulkan subset ---- */
VkResult vkMapMemory(VkDevice device, VkDeviceMemory memory, VkDeviceSize offset, VkDeviceSize size, VkMemoryMapFlags flags, void** ppData) {
	(void) device; (void) size; (void) flags;
	if (!memory || offset > memory->size) return VK_ERROR_INITIALIZATION_FAILED;
	*ppData = memory->data + offset; return VK_SUCCESS;
}
void vkUnmapMemory(VkDevice device, VkDeviceMemory memory) { (void) device; (void) memory; }
void vkFreeMemory(VkDevice device, VkDeviceMemory memory, const VkAllocationCallbacks* a) { (void) device; (void) a; if (memory) { free(memory->data); free(memory); } }
void vkDestroyBuffer(VkDevice device, VkBuffer buffer, const VkAllocationCallbacks* a) { (void) device; (void) a; free(buffer); }
VkResult vkCreateBufferView(VkDevice device, const VkBufferViewCreateInfo* info, const VkAllocationCallbacks* a, VkBufferView* view) {
	(void) device; (void) a;
	struct VkBufferView_T* v = (struct VkBufferView_T*) calloc(1, sizeof(*v)); v->buffer = info->buffer; v->format = info->format; *view = v; return VK_SUCCESS;
}
void vkDestroyBufferView(VkDevice device, VkBufferView view, const VkAllocationCallbacks* a) { (void) device; (void) a; free(view); }
VkResult vkCreateSampler(VkDevice device, const VkSamplerCreateInfo* info, const VkAllocationCallbacks* a, VkSampler* sampler) {
	(void) device; (void) a;
	struct VkSampler_T* s = (struct VkSampler_T*) calloc(1, sizeof(*s)); s->info = *info; *sampler = s; return VK_SUCCESS;
}
void vkDestroySampler(VkDevice device, VkSampler sampler, const VkAllocationCallbacks* a) { (void) device; (void) a; free(sampler); }
VkDeviceAddress vkGetBufferDeviceAddress(VkDevice device, const VkBufferDeviceAddressInfo* info) { (void) device; return (VkDeviceAddress) (uintptr_t) (info->buffer->memory->data + info->buffer->offset); }
VkResult vkAllocateCommandBuffers(VkDevice device, const VkCommandBufferAllocateInfo* info, VkCommandBuffer* out) {
	(void) device;
	for (uint32_t i = 0; i != info->commandBufferCount; ++i) out[i] = (VkCommandBuffer) calloc(1, sizeof(struct VkCommandBuffer_T));
	return VK_SUCCESS;
}
void vkFreeCommandBuffers(VkDevice device, VkCommandPool pool, uint32_t count, const VkCommandBuffer* buffers) { (void) device; (void) pool; for (uint32_t i = 0; i != count; ++i) free(buffers[i]); }
VkResult vkBeginCommandBuffer(VkCommandBuffer cmd, const VkCommandBufferBeginInfo* info) { (void) info; cmd->recording = 1; return VK_SUCCESS; }
VkResult vkEndCommandBuffer(VkCommandBuffer cmd) { cmd->recording = 0; return VK_SUCCESS; }
void vkCmdPipelineBarrier(VkCommandBuffer c, VkPipelineStageFlags s, VkPipelineStageFlags d, VkDependencyFlags f, uint32_t mc, const VkMemoryBarrier* m, uint32_t bc, const VkBufferMemoryBarrier* b, uint32_t ic, const VkImageMemoryBarrier* i) {
	(void) c; (void) s; (void) d; (void) f; (void) mc; (void) m; (void) bc; (void) b; (void) ic; (void) i;
}
VkResult vkQueueSubmit(VkQueue queue, uint32_t count, const VkSubmitInfo* submits, VkFence fence) { (void) queue; (void) count; (void) submits; (void) fence; return VK_SUCCESS; } /* work ran when it was recorded */
VkResult vkQueueWaitIdle(VkQueue queue) { (void) queue; return VK_SUCCESS; }
VkResult vkEnumeratePhysicalDevices(VkInstance instance, uint32_t* count, VkPhysicalDevice* devices) { (void) instance; (void) devices; *count = 0; return VK_SUCCESS; }

/* ---- VK_KHR_acceleration_structure through VK_LOAD ---- */
static void shim_GetAccelerationStructureBuildSizesKHR(VkDevice device, VkAccelerationStructureBuildTypeKHR type, const VkAccelerationStructureBuildGeometryInfoKHR* info, const uint32_t* counts, VkAccelerationStructureBuildSizesInfoKHR* sizes) {
	(void) device; (void) type; (void) info;
	sizes->accelerationStructureSize = 64 + 112 * (VkDeviceSize) counts[0]; /* node pairs + padded triangles of the software BVH2 */
	sizes->updateScratchSize = 16; sizes->buildScratchSize = 16;
}
static VkResult shim_CreateAccelerationStructureKHR(VkDevice device, const VkAccelerationStructureCreateInfoKHR* info, const VkAllocationCallbacks* a, VkAccelerationStructureKHR* out) {
	(void) device; (void) a;
	struct VkAccelerationStructureKHR_T* s = (struct VkAccelerationStructureKHR_T*) calloc(1, sizeof(*s)); s->type = info->type; *out = s; return VK_SUCCESS;
}
static void shim_DestroyAccelerationStructureKHR(VkDevice device, VkAccelerationStructureKHR s, const VkAllocationCallbacks* a) { (void) device; (void) a; if (s) { free(s->vertices); free(s); } }
static VkDeviceAddress shim_GetAccelerationStructureDeviceAddressKHR(VkDevice device, const VkAccelerationStructureDeviceAddressInfoKHR* info) { (void) device; return (VkDeviceAddress) (uintptr_t) info->accelerationStructure; }
/* The BVH build hook: keep a copy of the triangle soup (the staging buffer it lives in is destroyed right after) */
static void shim_CmdBuildAccelerationStructuresKHR(VkCommandBuffer cmd, uint32_t info_count, const VkAccelerationStructureBuildGeometryInfoKHR* infos, const VkAccelerationStructureBuildRangeInfoKHR* const* ranges) {
	(void) cmd;
	for (uint32_t i = 0; i != info_count; ++i) {
		struct VkAccelerationStructureKHR_T* dst = infos[i].dstAccelerationStructure;
		if (!dst || infos[i].type != VK_ACCELERATION_STRUCTURE_TYPE_BOTTOM_LEVEL_KHR || !infos[i].geometryCount) continue;
		const VkAccelerationStructureGeometryKHR* g = infos[i].pGeometries ? &infos[i].pGeometries[0] : infos[i].ppGeometries[0];
		if (g->geometryType != VK_GEOMETRY_TYPE_TRIANGLES_KHR || g->geometry.triangles.vertexFormat != VK_FORMAT_R32G32B32_SFLOAT || g->geometry.triangles.indexType != VK_INDEX_TYPE_NONE_KHR) continue;
		uint64_t n = ranges[i][0].primitiveCount;
		const char* src = (const char*) (uintptr_t) g->geometry.triangles.vertexData.deviceAddress;
		free(dst->vertices);
		dst->vertices = (float*) malloc(sizeof(float) * 9 * (n ? n : 1)); dst->triangle_count = n;
		for (uint64_t v = 0; v != 3 * n; ++v) memcpy(dst->vertices + 3 * v, src + g->geometry.triangles.vertexStride * v, sizeof(float) * 3);
	}
}

GLFWvkproc glfwGetInstanceProcAddress(VkInstance instance, const char* name) {
	(void) instance;
	if (!strcmp(name, "vkGetAccelerationStructureBuildSizesKHR")) return (GLFWvkproc) shim_GetAccelerationStructureBuildSizesKHR;
	if (!strcmp(name, "vkCreateAccelerationStructureKHR")) return (GLFWvkproc) shim_CreateAccelerationStructureKHR;
	if (!strcmp(name, "vkDestroyAccelerationStructureKHR")) return (GLFWvkproc) shim_DestroyAccelerationStructureKHR;
	if (!strcmp(name, "vkGetAccelerationStructureDeviceAddressKHR")) return (GLFWvkproc) shim_GetAccelerationStructureDeviceAddressKHR;
	if (!strcmp(name, "vkCmdBuildAccelerationStructuresKHR")) return (GLFWvkproc) shim_CmdBuildAccelerationStructuresKHR;
	return NULL;
}
int glfwGetKey(GLFWwindow* w, int key) { (void) w; (void) key; return GLFW_RELEASE; }
int glfwGetMouseButton(GLFWwindow* w, int button) { (void) w; (void) button; return GLFW_RELEASE; }
void glfwGetCursorPos(GLFWwindow* w, double* x, double* y) { (void) w; *x = 0.0; *y = 0.0; }
double glfwGetTime(void) { return 0.0; }

/* ---- accessors for the CUDA side ---- */
void* vkr_shim_buffer_data(VkBuffer buffer, VkDeviceSize* out_size) { if (out_size) *out_size = buffer->size; return buffer->memory->data + buffer->offset; }
void* vkr_shim_image_data(VkImage image, uint32_t mip_level, uint32_t array_layer, VkDeviceSize* out_size) {
	if (mip_level >= image->info.mipLevels || array_layer >= image->info.arrayLayers) return NULL;
	if (out_size) *out_size = image->level_offsets[mip_level + 1] - image->level_offsets[mip_level];
	return image->data + image->layer_size * array_layer + image->level_offsets[mip_level];
}
const float* vkr_shim_acceleration_structure_vertices(VkAccelerationStructureKHR s, uint64_t* out_triangle_count) { if (out_triangle_count) *out_triangle_count = s ? s->triangle_count : 0; return s ? s->vertices : NULL; }
