"""world_size-2 gloo test (CPU) of the multi-GPU host logic: interleaved tile-row stripes + one all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vulkan_renderer_b200.stripes import StripeGather, stripe_rows


def _pattern(height, width):
	y = torch.arange(height, dtype=torch.float32)[:, None, None]; x = torch.arange(width, dtype=torch.float32)[None, :, None]
	c = torch.arange(4, dtype=torch.float32)[None, None, :]
	return y * 1000.0 + x + c * 0.25


def _worker(rank, world, port, height, width, result_path):
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		sg = StripeGather(height, width, rank, world, torch.device("cpu"))
		frame = torch.full((height, width, 4), -1.0)
		rows = stripe_rows(height, rank, world)
		frame[rows] = _pattern(height, width)[rows]      # "shade" this rank's stripe
		sg.gather_frame(frame)
		ok = torch.equal(frame, _pattern(height, width))
		flags = [torch.zeros(1) for _ in range(world)]
		dist.all_gather(flags, torch.tensor([float(ok)]))
		if rank == 0:
			with open(result_path, "w") as f:
				f.write("ok" if all(bool(v.item()) for v in flags) else "mismatch")
	finally:
		dist.destroy_process_group()


def _free_port():
	s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close(); return port


def test_stripe_partition_covers_every_row_once():
	for height in (1080, 2160, 100, 8, 7):
		for world in (1, 2, 4, 8):
			rows = sorted(y for r in range(world) for y in stripe_rows(height, r, world))
			assert rows == list(range(height))
			counts = [len(stripe_rows(height, r, world)) for r in range(world)]
			assert max(counts) - min(counts) <= 8      # interleaving balances the ranks to one tile row


def test_gather_reassembles_the_frame_world_size_2(tmp_path):
	for height, width in ((100, 24), (1080 // 8, 16)):
		result = tmp_path / ("result_%d.txt" % height)
		mp.spawn(_worker, args=(2, _free_port(), height, width, str(result)), nprocs=2, join=True)
		assert result.read_text() == "ok"
