"""Arithmetic of the kernels that is written by hand instead of left to the compiler, checked on the device against its definition."""
import ctypes as C

import pytest

from tests import harness as H

pytestmark = pytest.mark.gpu


def test_reciprocal_square_root_equals_its_definition_for_every_float():
	"""rsqrt_ieee() (csrc/vkr_device_math.cuh) runs the compiler's fast paths of sqrtf and of the division behind one range check instead of two. Its definition
	is 1.0f / sqrtf(x) with both operations correctly rounded (arithmetic contract); the probe compares the two for all 2^32 bit patterns."""
	frame = H.open_frame(H.dataset("cornell"), cuda_device=0)
	try:
		mismatches = C.c_uint64(1); first = C.c_uint32(0)
		assert frame.lib.vkr_probe_rsqrt_exhaustive(C.byref(frame.device), C.byref(mismatches), C.byref(first)) == 0
		assert mismatches.value == 0, "first input that differs: 0x%08x" % first.value
	finally:
		frame.close()
