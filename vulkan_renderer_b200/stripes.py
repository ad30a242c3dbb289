"""Multi-GPU partition of a frame: interleaved 8-pixel screen-tile rows, one all-gather of HDR stripes.

Tile row t (pixel rows 8t .. 8t+7) belongs to rank t % world (vkr_shading_pass_desc_t.stripe_index / stripe_count).
Every rank shades its tile rows into a full-size frame buffer; `gather_frame` packs the rank's rows, exchanges them
with ONE torch.distributed all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests) and scatters the rows of
all ranks into every rank's frame. Read-only inputs (BVH, tables, constants) are replicated per GPU.
"""
import torch
import torch.distributed as dist

TILE_ROW_HEIGHT = 8


def stripe_rows(height, rank, world):
	"""Pixel rows owned by `rank`, in ascending order."""
	tile_rows = (height + TILE_ROW_HEIGHT - 1) // TILE_ROW_HEIGHT
	return [y for t in range(rank, tile_rows, world) for y in range(t * TILE_ROW_HEIGHT, min((t + 1) * TILE_ROW_HEIGHT, height))]


class StripeGather:
	"""Precomputed index tensors for one frame size."""

	def __init__(self, height, width, rank, world, device):
		self.rank, self.world, self.height, self.width = rank, world, height, width
		rows = [stripe_rows(height, r, world) for r in range(world)]
		self.max_rows = max(len(r) for r in rows)
		pad = lambda r: (r + [r[-1] if r else 0] * self.max_rows)[:self.max_rows]
		self.my_rows = torch.tensor(pad(rows[rank]), dtype=torch.long, device=device)
		self.all_rows = torch.tensor([y for r in rows for y in pad(r)], dtype=torch.long, device=device)
		self.row_counts = [len(r) for r in rows]
		self.gathered = torch.empty((world, self.max_rows, width, 4), dtype=torch.float32, device=device)

	def gather_frame(self, frame):
		"""frame: [H, W, 4] with this rank's rows valid -> all rows valid on every rank (in place)."""
		if self.world == 1:
			return frame
		stripe = frame.index_select(0, self.my_rows).contiguous()
		dist.all_gather_into_tensor(self.gathered, stripe) if frame.is_cuda else dist.all_gather(list(self.gathered.unbind(0)), stripe)
		# padded duplicates carry the same row content as the original, so the scatter is well defined
		frame.index_copy_(0, self.all_rows, self.gathered.reshape(-1, self.width, 4))
		return frame
