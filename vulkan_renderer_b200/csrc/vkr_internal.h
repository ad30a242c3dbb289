// vkr_internal.h -- helpers shared by the host translation units of libvkr_b200.so
#pragma once
#include "../../include/vkr_b200.h"
#include <stddef.h>
namespace vkr {
// cudaMalloc + synchronous copy; returns non-zero on failure and leaves *d_ptr NULL
int upload(void** d_ptr, const void* src, size_t bytes, const vkr_device_t* device);
}
