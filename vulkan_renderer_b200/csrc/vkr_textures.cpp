// vkr_textures.cpp -- material textures as the G-buffer producer consumes them.
//
// Replaces load_2d_textures() for material textures (src/textures.c:111-169, called at src/scene.c:529-540) and the texture units
// that decode them: every mip level of a *.vkt file is decoded to RGBA32F on the host once, at load time -- R16G16B16(A16)_SFLOAT,
// R32G32B32(A32)_SFLOAT, R8G8B8A8_UNORM / SRGB, BC1_RGB_UNORM / SRGB and BC5_UNORM, the formats the reference's converter writes
// (tools/texture_conversion/main.c:27-35) -- and filtered on the device by csrc/vkr_texture.cuh. Decoding BC blocks up front costs
// memory (16 bytes per texel instead of 0.5 or 1), not correctness; a block-compressed device path is a later optimisation.
// sRGB formats are converted to linear like the sampler hardware does before filtering.
#include "vkr_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

float half_to_float(uint16_t h) {
	const uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 31u, man = h & 1023u, bits;
	if (exp == 0) {
		if (man == 0) bits = sign;
		else { int e = -1; do { ++e; man <<= 1; } while (!(man & 1024u)); bits = sign | ((uint32_t) (127 - 15 - e) << 23) | ((man & 1023u) << 13); }
	}
	else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
	else bits = sign | ((exp + 112u) << 23) | (man << 13);
	float f; memcpy(&f, &bits, 4); return f;
}
float srgb_to_linear(float c) { return (c <= 0.04045f) ? (c / 12.92f) : powf((c + 0.055f) / 1.055f, 2.4f); }

// BC1 block (8 bytes) -> 16 RGB texels, row-major inside the block (the RGB variant has no punch-through alpha: index 3 of the 3-colour mode is black)
void decode_bc1_block(const uint8_t* block, float rgb[16][3]) {
	uint16_t c0, c1; memcpy(&c0, block, 2); memcpy(&c1, block + 2, 2);
	float palette[4][3];
	const float a[3] = { (float) ((c0 >> 11) & 31) / 31.0f, (float) ((c0 >> 5) & 63) / 63.0f, (float) (c0 & 31) / 31.0f };
	const float b[3] = { (float) ((c1 >> 11) & 31) / 31.0f, (float) ((c1 >> 5) & 63) / 63.0f, (float) (c1 & 31) / 31.0f };
	for (int i = 0; i != 3; ++i) {
		palette[0][i] = a[i]; palette[1][i] = b[i];
		if (c0 > c1) { palette[2][i] = (2.0f * a[i] + b[i]) / 3.0f; palette[3][i] = (a[i] + 2.0f * b[i]) / 3.0f; }
		else { palette[2][i] = 0.5f * (a[i] + b[i]); palette[3][i] = 0.0f; }
	}
	uint32_t indices; memcpy(&indices, block + 4, 4);
	for (int t = 0; t != 16; ++t) for (int i = 0; i != 3; ++i) rgb[t][i] = palette[(indices >> (2 * t)) & 3u][i];
}

// BC4 block (8 bytes) -> 16 values
void decode_bc4_block(const uint8_t* block, float value[16]) {
	const float r0 = (float) block[0] / 255.0f, r1 = (float) block[1] / 255.0f;
	float palette[8];
	palette[0] = r0; palette[1] = r1;
	if (block[0] > block[1]) for (int i = 2; i != 8; ++i) palette[i] = ((float) (8 - i) * r0 + (float) (i - 1) * r1) / 7.0f;
	else { for (int i = 2; i != 6; ++i) palette[i] = ((float) (6 - i) * r0 + (float) (i - 1) * r1) / 5.0f; palette[6] = 0.0f; palette[7] = 1.0f; }
	uint64_t bits = 0;
	for (int i = 0; i != 6; ++i) bits |= (uint64_t) block[2 + i] << (8 * i);
	for (int t = 0; t != 16; ++t) value[t] = palette[(bits >> (3 * t)) & 7u];
}

// One mip level of `format` -> RGBA32F. Returns false if the level is too small for its payload or the format is unknown.
bool decode_level(float* out, uint32_t width, uint32_t height, uint32_t format, const uint8_t* data, uint64_t size) {
	const size_t pixel_count = (size_t) width * height;
	const uint32_t bw = (width + 3) / 4, bh = (height + 3) / 4;
	{ // the payload must hold the level before a single texel is written (a corrupt header must not make this loop write gigabytes)
		uint64_t need;
		switch (format) {
		case 97: need = (uint64_t) pixel_count * 8; break; case 90: need = (uint64_t) pixel_count * 6; break;
		case 109: need = (uint64_t) pixel_count * 16; break; case 106: need = (uint64_t) pixel_count * 12; break;
		case 37: case 43: need = (uint64_t) pixel_count * 4; break;
		case 131: case 132: case 133: case 134: need = (uint64_t) bw * bh * 8; break;
		case 141: case 142: need = (uint64_t) bw * bh * 16; break;
		default: return false;
		}
		if (size < need) return false;
	}
	for (size_t i = 0; i != pixel_count; ++i) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = 0.0f; out[4 * i + 3] = 1.0f; }
	switch (format) {
	case 97: case 90: { // R16G16B16A16_SFLOAT, R16G16B16_SFLOAT
		const uint32_t channels = (format == 97) ? 4 : 3;
		if (size < pixel_count * channels * 2) return false;
		for (size_t i = 0; i != pixel_count; ++i) for (uint32_t c = 0; c != channels; ++c) { uint16_t h; memcpy(&h, data + 2 * (i * channels + c), 2); out[4 * i + c] = half_to_float(h); }
		return true; }
	case 109: case 106: { // R32G32B32A32_SFLOAT, R32G32B32_SFLOAT
		const uint32_t channels = (format == 109) ? 4 : 3;
		if (size < pixel_count * channels * 4) return false;
		for (size_t i = 0; i != pixel_count; ++i) memcpy(out + 4 * i, data + 4 * i * channels, 4 * channels);
		return true; }
	case 37: case 43: // R8G8B8A8_UNORM, R8G8B8A8_SRGB
		if (size < pixel_count * 4) return false;
		for (size_t i = 0; i != pixel_count; ++i) for (int c = 0; c != 4; ++c) {
			const float v = (float) data[4 * i + c] / 255.0f;
			out[4 * i + c] = (format == 43 && c != 3) ? srgb_to_linear(v) : v;
		}
		return true;
	case 131: case 132: // BC1_RGB_UNORM_BLOCK, BC1_RGB_SRGB_BLOCK
		if (size < (uint64_t) bw * bh * 8) return false;
		for (uint32_t by = 0; by != bh; ++by) for (uint32_t bx = 0; bx != bw; ++bx) {
			float rgb[16][3];
			decode_bc1_block(data + 8 * ((size_t) by * bw + bx), rgb);
			for (uint32_t t = 0; t != 16; ++t) {
				const uint32_t x = 4 * bx + (t & 3), y = 4 * by + (t >> 2);
				if (x < width && y < height) for (int c = 0; c != 3; ++c) out[4 * ((size_t) y * width + x) + c] = (format == 132) ? srgb_to_linear(rgb[t][c]) : rgb[t][c];
			}
		}
		return true;
	case 141: // BC5_UNORM_BLOCK: two BC4 blocks, red then green
		if (size < (uint64_t) bw * bh * 16) return false;
		for (uint32_t by = 0; by != bh; ++by) for (uint32_t bx = 0; bx != bw; ++bx) {
			float red[16], green[16];
			decode_bc4_block(data + 16 * ((size_t) by * bw + bx), red);
			decode_bc4_block(data + 16 * ((size_t) by * bw + bx) + 8, green);
			for (uint32_t t = 0; t != 16; ++t) {
				const uint32_t x = 4 * bx + (t & 3), y = 4 * by + (t >> 2);
				if (x < width && y < height) { out[4 * ((size_t) y * width + x)] = red[t]; out[4 * ((size_t) y * width + x) + 1] = green[t]; }
			}
		}
		return true;
	default:
		return false;
	}
}

} // namespace

extern "C" void vkr_destroy_texture(vkr_texture_t* texture) {
	free(texture->h_texels);
	memset(texture, 0, sizeof(*texture));
}

// Boundary B1: the mip levels of an image as the reference's load_2d_textures() holds them after reading a *.vkt file (textures.c:111-169): raw bytes of `vk_format`
extern "C" int vkr_texture_from_levels(vkr_texture_t* texture, uint32_t width, uint32_t height, uint32_t mip_count, uint32_t vk_format, const void* const* level_data, const uint64_t* level_sizes) {
	memset(texture, 0, sizeof(*texture));
	if (!mip_count || mip_count > 32 || !width || !height || width > 32768 || height > 32768) { printf("Cannot take over a texture of %ux%u texels with %u mipmaps.\n", width, height, mip_count); return 1; }
	uint64_t float_count = 0;
	for (uint32_t k = 0; k != mip_count; ++k) float_count += 4 * (uint64_t) ((width >> k) ? (width >> k) : 1) * ((height >> k) ? (height >> k) : 1);
	float* texels = (float*) malloc(sizeof(float) * (size_t) float_count);
	if (!texels) return 1;
	uint64_t at = 0;
	for (uint32_t k = 0; k != mip_count; ++k) {
		const uint32_t w = (width >> k) ? (width >> k) : 1, h = (height >> k) ? (height >> k) : 1;
		if (!level_data[k] || !decode_level(texels + at, w, h, vk_format, (const uint8_t*) level_data[k], level_sizes[k])) {
			printf("A texture handed over has VkFormat %u or a mipmap size that this library cannot read.\n", vk_format);
			free(texels); return 1;
		}
		at += 4 * (uint64_t) w * h;
	}
	int constant = 1;
	for (uint64_t i = 4; i < float_count && constant; ++i) constant = texels[i] == texels[i & 3];
	texture->width = width; texture->height = height; texture->mip_count = mip_count; texture->vk_format = vk_format;
	texture->h_texels = texels; texture->texel_float_count = float_count; texture->is_constant = constant;
	return 0;
}

extern "C" int vkr_load_texture(vkr_texture_t* texture, const char* file_path) {
	memset(texture, 0, sizeof(*texture));
	FILE* file = fopen(file_path, "rb");
	if (!file) { printf("Failed to open the texture file at path %s.\n", file_path); return 1; }
	uint32_t header[6] = {0}; uint64_t payload_size = 0;
	if (fread(header, 4, 6, file) != 6 || fread(&payload_size, 8, 1, file) != 1 || header[0] != 0xbc1bc1 || header[1] != 1) {
		printf("The texture at path %s does not have the *.vkt format. Aborting.\n", file_path); fclose(file); return 1; // textures.c:117-121
	}
	const uint32_t mip_count = header[2], width = header[3], height = header[4], format = header[5];
	if (mip_count == 0 || mip_count > 32 || width == 0 || height == 0 || width > 32768 || height > 32768 || payload_size > (1ull << 34)) {
		printf("The texture at path %s has an invalid header (%u mipmaps, %ux%u).\n", file_path, mip_count, width, height); fclose(file); return 1;
	}
	struct mip_header { uint32_t width, height; uint64_t size, offset; };
	std::vector<mip_header> mips(mip_count);
	uint64_t float_count = 0;
	for (uint32_t k = 0; k != mip_count; ++k) {
		if (fread(&mips[k].width, 4, 1, file) != 1 || fread(&mips[k].height, 4, 1, file) != 1 || fread(&mips[k].size, 8, 1, file) != 1 || fread(&mips[k].offset, 8, 1, file) != 1) { fclose(file); return 1; }
		const uint32_t ew = (width >> k) ? (width >> k) : 1, eh = (height >> k) ? (height >> k) : 1;
		if (mips[k].width != ew || mips[k].height != eh || mips[k].offset > payload_size || mips[k].size > payload_size - mips[k].offset) {
			printf("The texture at path %s has an unexpected mipmap %u (%ux%u).\n", file_path, k, mips[k].width, mips[k].height); fclose(file); return 1;
		}
		float_count += 4 * (uint64_t) ew * eh;
	}
	std::vector<uint8_t> payload((size_t) payload_size);
	uint32_t eof_marker = 0;
	if ((payload_size && fread(payload.data(), 1, (size_t) payload_size, file) != payload_size) || fread(&eof_marker, 4, 1, file) != 1 || eof_marker != 0xE0FE0F) {
		printf("The texture file at path %s seems to be invalid. The texture data is not followed by the expected end of file marker.\n", file_path); // textures.c:163-167
		fclose(file); return 1;
	}
	fclose(file);
	float* texels = (float*) malloc(sizeof(float) * (size_t) float_count);
	if (!texels) { printf("Failed to allocate %llu floats for the texture at path %s.\n", (unsigned long long) float_count, file_path); return 1; }
	uint64_t at = 0;
	for (uint32_t k = 0; k != mip_count; ++k) {
		if (!decode_level(texels + at, mips[k].width, mips[k].height, format, payload.data() + mips[k].offset, mips[k].size)) {
			printf("The texture at path %s has VkFormat %u or a mipmap size that this library cannot read.\n", file_path, format);
			free(texels); return 1;
		}
		at += 4 * (uint64_t) mips[k].width * mips[k].height;
	}
	// constant = every texel of every level equals the first one: such a texture needs no filtering (and the G-buffer producer then keeps its one-texel path)
	int constant = 1;
	for (uint64_t i = 4; i < float_count && constant; ++i) constant = texels[i] == texels[i & 3];
	texture->width = width; texture->height = height; texture->mip_count = mip_count; texture->vk_format = format;
	texture->h_texels = texels; texture->texel_float_count = float_count; texture->is_constant = constant;
	return 0;
}
