"""Multi-GPU partition of a frame: host-side mirror of the split the C-ABI implements, and the exchange of a frame between processes.

The frame is cut into 16x8 screen tiles; GPU `rank` of `world` shades tile column tx of every tile row if tx % world == rank
(vkr_shading_pass_desc_t.stripe_index / stripe_count). The product path exchanges pixels inside the shading kernel (vkr_frame_exchange_t:
peer stores over NVLink, csrc/vkr_exchange.cu); `connect_exchange` only carries the 64-byte IPC handles between the processes with
torch.distributed. `ShareGather` is the same exchange written with one torch.distributed all_gather: the CPU tests (gloo, world size 2)
use it to check the partition, and bench.py falls back to it if the GPUs of a box cannot map each other's memory.
"""
import ctypes as C

import torch
import torch.distributed as dist

TILE_WIDTH = 16
TILE_ROW_HEIGHT = 8


def share_columns(width, rank, world):
	"""Pixel columns owned by `rank`, in ascending order."""
	tiles_x = (width + TILE_WIDTH - 1) // TILE_WIDTH
	return [x for t in range(rank, tiles_x, world) for x in range(t * TILE_WIDTH, min((t + 1) * TILE_WIDTH, width))]


def share_tiles(width, height, rank, world):
	"""Tile indices (ty * tiles_x + tx) owned by `rank`, row-major: the launch order of a pass before it has measured tile costs."""
	tiles_x = (width + TILE_WIDTH - 1) // TILE_WIDTH
	tiles_y = (height + TILE_ROW_HEIGHT - 1) // TILE_ROW_HEIGHT
	return [ty * tiles_x + tx for ty in range(tiles_y) for tx in range(rank, tiles_x, world)]


class ShareGather:
	"""The exchange as ONE all_gather of packed tile columns (fallback and CPU test edition of vkr_frame_exchange_t)."""

	def __init__(self, height, width, rank, world, device):
		self.rank, self.world, self.height, self.width = rank, world, height, width
		cols = [share_columns(width, r, world) for r in range(world)]
		self.counts = [len(c) for c in cols]
		self.max_cols = max(self.counts)
		pad = lambda c: (c + [c[-1] if c else 0] * self.max_cols)[:self.max_cols]
		self.my_cols = torch.tensor(pad(cols[rank]), dtype=torch.long, device=device)
		# ranks without a column (more GPUs than tile columns) and the padding of short shares take no part in the scatter
		self.real = torch.tensor([r * self.max_cols + i for r in range(world) for i in range(self.counts[r])], dtype=torch.long, device=device)
		self.all_cols = torch.tensor([x for c in cols for x in c], dtype=torch.long, device=device)
		self.gathered = torch.empty((world, height, self.max_cols, 4), dtype=torch.float32, device=device)

	def gather_frame(self, frame):
		"""frame: [H, W, 4] with this rank's tile columns valid -> all columns valid on every rank (in place)."""
		if self.world == 1:
			return frame
		share = frame.index_select(1, self.my_cols).contiguous()
		dist.all_gather_into_tensor(self.gathered, share) if frame.is_cuda else dist.all_gather(list(self.gathered.unbind(0)), share)
		packed = self.gathered.permute(1, 0, 2, 3).reshape(self.height, self.world * self.max_cols, 4)
		frame.index_copy_(1, self.all_cols, packed.index_select(1, self.real))
		return frame


def connect_exchange(lib, exchange, device):
	"""Sends this process's IPC handle to all ranks and maps theirs (vkr_frame_exchange_get_handle / _connect). Collective: every rank calls it."""
	handle = (C.c_ubyte * 64)()
	if lib.vkr_frame_exchange_get_handle(C.byref(exchange), C.byref(device), handle) != 0:
		raise RuntimeError("vkr_frame_exchange_get_handle failed")
	handles = [None] * exchange.world
	dist.all_gather_object(handles, bytes(handle))
	blob = (C.c_ubyte * (64 * exchange.world)).from_buffer_copy(b"".join(handles))
	if lib.vkr_frame_exchange_connect(C.byref(exchange), C.byref(device), blob) != 0:
		raise RuntimeError("vkr_frame_exchange_connect failed: the GPUs of this box cannot map each other's memory")
