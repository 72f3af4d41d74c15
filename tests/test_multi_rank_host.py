"""CPU test (gloo, world_size 2) of the N>1 path's host logic: interleaved row-block partition + one
reduce(sum) of the partial films per pass gives the single-rank film.  The renderer behind it here is the oracle (no GPU
in this container); bench.py runs the same partition/reduce with the CUDA renderer and NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition_covers_frame_once():
    sys.path.insert(0, ROOT)
    import bench
    for h in (90, 360, 1080, 2160):
        for world in (1, 2, 4, 8):
            cover = np.zeros(h, np.int32)
            for rank in range(world):
                for y0, y1 in bench.my_rows(h, rank, world):
                    assert 0 <= y0 < y1 <= h
                    cover[y0:y1] += 1
            assert np.all(cover == 1)
            if world > 1 and h >= bench.BLOCK_ROWS * world:
                sizes = [sum(y1 - y0 for y0, y1 in bench.my_rows(h, r, world)) for r in range(world)]
                assert max(sizes) - min(sizes) <= bench.BLOCK_ROWS + h % bench.BLOCK_ROWS


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bench
    from buas_pathtracer_b200 import scenes
    from oracle import ref_oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, spp = 96, 160, 2
    s = ref_oracle.RefScene()
    scenes.c1_week3(s, w, h)
    film = np.zeros((h, w, 4), np.float32)
    for y0, y1 in bench.my_rows(h, rank, world):
        s.render_parity(w, h, spp, rect=(0, y0, w, y1), film=film)
    t = torch.from_numpy(film)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)          # the one collective per progressive pass
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_two_rank_row_sharded_render_equals_single_rank(oracle, tmp_path):
    import torch.multiprocessing as mp
    from buas_pathtracer_b200 import scenes
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sharded = np.load(out)
    s = oracle.RefScene()
    scenes.c1_week3(s, 96, 160)
    full, _ = s.render_parity(96, 160, 2)
    # per-pixel seeding makes the samples identical; only the order of float adds across band borders differs
    assert np.allclose(sharded, full, rtol=1e-5, atol=1e-6)
    assert np.array_equal(sharded[..., 3] > 0, full[..., 3] > 0)
