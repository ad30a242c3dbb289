"""Multi-GPU partition of a frame: host-side mirror of the split the C-ABI implements, and the exchange of a frame between processes.

The frame is cut into 16x8 screen tiles; GPU `rank` of `world` shades tile (tx, ty) if (tx + ty // 8) % world == rank: tile columns dealt out in turn,
the deal moving on by one column every 8 tile rows (vkr_shading_pass_desc_t.stripe_index / stripe_count, VKR_TILE_BAND_ROWS). The product path exchanges
pixels inside the shading kernel (vkr_frame_exchange_t: peer stores over NVLink, csrc/vkr_exchange.cu); `connect_exchange` only carries the 64-byte IPC
handles between the processes with torch.distributed. `ShareGather` is the same exchange written with one torch.distributed all_gather: the CPU tests
(gloo, world size 2) use it to check the partition, and bench.py falls back to it if the GPUs of a box cannot map each other's memory.
"""
import ctypes as C

import torch
import torch.distributed as dist

TILE_WIDTH = 16
TILE_ROW_HEIGHT = 8
TILE_BAND_ROWS = 8


def share_tiles(width, height, rank, world):
	"""Tile indices (ty * tiles_x + tx) owned by `rank`, row-major: the launch order of a pass before it has measured tile costs."""
	tiles_x = (width + TILE_WIDTH - 1) // TILE_WIDTH
	tiles_y = (height + TILE_ROW_HEIGHT - 1) // TILE_ROW_HEIGHT
	return [ty * tiles_x + tx for ty in range(tiles_y) for tx in range((rank - ty // TILE_BAND_ROWS) % world, tiles_x, world)]


def share_pixels(width, height, rank, world):
	"""Flat pixel indices (y * width + x) owned by `rank`, ascending."""
	tiles_x = (width + TILE_WIDTH - 1) // TILE_WIDTH
	out = []
	for t in share_tiles(width, height, rank, world):
		ty, tx = divmod(t, tiles_x)
		for y in range(ty * TILE_ROW_HEIGHT, min((ty + 1) * TILE_ROW_HEIGHT, height)):
			out.extend(range(y * width + tx * TILE_WIDTH, y * width + min((tx + 1) * TILE_WIDTH, width)))
	return sorted(out)


class ShareGather:
	"""The exchange as ONE all_gather of packed pixels (fallback and CPU test edition of vkr_frame_exchange_t)."""

	def __init__(self, height, width, rank, world, device):
		self.rank, self.world, self.height, self.width = rank, world, height, width
		pixels = [share_pixels(width, height, r, world) for r in range(world)]
		self.counts = [len(p) for p in pixels]
		self.max_count = max(max(self.counts), 1)
		pad = lambda p: (p + [p[-1] if p else 0] * self.max_count)[:self.max_count]
		self.mine = torch.tensor(pad(pixels[rank]), dtype=torch.long, device=device)
		# ranks without a tile (more GPUs than tile columns) and the padding of short shares take no part in the scatter
		self.real = torch.tensor([r * self.max_count + i for r in range(world) for i in range(self.counts[r])], dtype=torch.long, device=device)
		self.all_pixels = torch.tensor([x for p in pixels for x in p], dtype=torch.long, device=device)
		self.gathered = torch.empty((world, self.max_count, 4), dtype=torch.float32, device=device)

	def gather_frame(self, frame):
		"""frame: [H, W, 4] with this rank's tiles valid -> all tiles valid on every rank (in place)."""
		if self.world == 1:
			return frame
		flat = frame.view(-1, 4)
		share = flat.index_select(0, self.mine).contiguous()
		dist.all_gather_into_tensor(self.gathered, share) if frame.is_cuda else dist.all_gather(list(self.gathered.unbind(0)), share)
		flat.index_copy_(0, self.all_pixels, self.gathered.view(-1, 4).index_select(0, self.real))
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
