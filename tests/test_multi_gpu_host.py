"""world_size-2 gloo test (CPU) of the multi-GPU host logic: the split of a frame into tile columns per GPU and the exchange written with
one all_gather (the product path exchanges pixels inside the shading kernel, csrc/vkr_exchange.cu; GPU edition: tests/test_gpu_zzzy_multi.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vulkan_renderer_b200.stripes import ShareGather, share_pixels, share_tiles


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
		mine = torch.tensor(share_pixels(width, height, rank, world), dtype=torch.long)
		frame.view(-1, 4)[mine] = _pattern(height, width).view(-1, 4)[mine]      # "shade" this rank's tiles
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


def test_share_partition_covers_every_tile_and_pixel_once():
	for width, height in ((1920, 1080), (3840, 2160), (100, 75), (16, 8), (7, 200)):
		for world in (1, 2, 3, 4, 8):
			tiles_x = (width + 15) // 16; tiles_y = (height + 7) // 8
			shares = [share_tiles(width, height, r, world) for r in range(world)]
			assert sorted(t for s in shares for t in s) == list(range(tiles_x * tiles_y))
			if width * height <= 20000:
				assert sorted(p for r in range(world) for p in share_pixels(width, height, r, world)) == list(range(width * height))
			# the deal moves on by one column from band to band (8 tile rows): in band b the first column of rank r is (r - b) % world
			for r in range(world):
				for band in range((tiles_y + 7) // 8):
					first = [t % tiles_x for t in shares[r] if t // tiles_x == 8 * band]
					assert first == list(range((r - band) % world, tiles_x, world))
	# the shapes of the benchmark configurations divide evenly: every GPU gets the same number of tiles
	for width, height in ((1920, 1080), (3840, 2160)):
		for world in (2, 4, 8):
			assert len({len(share_tiles(width, height, r, world)) for r in range(world)}) == 1


def test_gather_reassembles_the_frame_world_size_2(tmp_path):
	for height, width in ((24, 100), (136, 1080 // 8), (8, 16)):   # ragged last tile; more than two bands of tile rows; one tile for two ranks (a rank without a tile)
		result = tmp_path / ("result_%d_%d.txt" % (height, width))
		mp.spawn(_worker, args=(2, _free_port(), height, width, str(result)), nprocs=2, join=True)
		assert result.read_text() == "ok"
