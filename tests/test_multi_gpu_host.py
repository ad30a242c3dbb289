"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the split of a frame into tile columns per GPU and the exchange written with
one all_gather (the product path exchanges pixels inside the shading kernel, csrc/vkr_exchange.cu; GPU edition: tests/test_gpu_zzzy_multi.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vulkan_renderer_b200.stripes import ShareGather, share_columns, share_tiles


def _pattern(height, width):
	y = torch.arange(height, dtype=torch.float32)[:, None, None]; x = torch.arange(width, dtype=torch.float32)[None, :, None]
	c = torch.arange(4, dtype=torch.float32)[None, None, :]
	return y * 1000.0 + x + c * 0.25


def _worker(rank, world, port, height, width, result_path):
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		sg = ShareGather(height, width, rank, world, torch.device("cpu"))
		frame = torch.full((height, width, 4), -1.0)
		cols = share_columns(width, rank, world)
		frame[:, cols] = _pattern(height, width)[:, cols]      # "shade" this rank's tile columns
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


def test_share_partition_covers_every_column_and_tile_once():
	for width in (1920, 3840, 100, 16, 7):
		for world in (1, 2, 3, 4, 8):
			cols = sorted(x for r in range(world) for x in share_columns(width, r, world))
			assert cols == list(range(width))
			counts = [len(share_columns(width, r, world)) for r in range(world)]
			assert max(counts) - min(counts) <= 16     # interleaving balances the ranks to one tile column
			height = 40
			tiles = sorted(t for r in range(world) for t in share_tiles(width, height, r, world))
			assert tiles == list(range(((width + 15) // 16) * 5))
	# the shapes of the benchmark configurations divide evenly: every GPU gets the same number of tiles
	for width, height in ((1920, 1080), (3840, 2160)):
		for world in (2, 4, 8):
			assert len({len(share_tiles(width, height, r, world)) for r in range(world)}) == 1


def test_gather_reassembles_the_frame_world_size_2(tmp_path):
	for height, width in ((24, 100), (16, 1080 // 8), (8, 16)):   # ragged last tile; odd tile count; one tile for two ranks (a rank without a column)
		result = tmp_path / ("result_%d_%d.txt" % (height, width))
		mp.spawn(_worker, args=(2, _free_port(), height, width, str(result)), nprocs=2, join=True)
		assert result.read_text() == "ok"
