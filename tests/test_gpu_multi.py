"""Multi-GPU on the hardware (-m gpu, needs >= 2 GPUs: `gpurun --gpus 2`; skipped on a one-GPU box): the row-block
partition + ONE NCCL reduce per progressive pass inside the library (bpt_reduce_film, include/bpt.h section 3) against one
GPU rendering the whole frame, same seeds.  The two differ only in the order of float additions (film atomics + the
reduce): rtol 1e-4 / atol 1e-5.  Also: progressive passes must not double count (the partial films are not modified by
the reduce), and the headless C++ driver's --gpus N writes the same bitmap as its one-GPU run up to 1 LSB."""
import os
import subprocess
import threading

import numpy as np
import pytest

from buas_pathtracer_b200 import scenes, lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _bands(h, rank, world, block=8):
    return [(y0, min(h, y0 + block)) for b, y0 in enumerate(range(0, h, block)) if b % world == rank]


@pytest.mark.parametrize("recipe,kw", [(scenes.c1_week3, {}), (scenes.c2_icosphere, dict(level=5))])
def test_two_rank_reduced_film_equals_single_gpu_film(bpt, recipe, kw):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    w, h, spp, passes, world = 320, 180, 4, 2, 2
    s = bpt.Scene()
    recipe(s, w, h, **kw)
    rs = [bpt.Renderer(d) for d in range(world)]
    comms = lib.nccl_comm_init_all(list(range(world)))
    errors = []
    per_pass = [np.zeros((h, w, 4), np.float32) for _ in range(passes)]

    def rank_main(rank):
        try:
            r = rs[rank]
            r.upload_scene(s)
            r.film_resize(w, h)
            for p in range(passes):                 # progressive: the partial films keep accumulating, every pass ends in a reduce
                r.render_pass_bands(spp, _bands(h, rank, world), frame_count=p * spp)
                r.reduce_film(comms[rank], 0)
                if rank == 0:
                    r.download_film_async(per_pass[p], reduced=True)    # the next pass is enqueued right behind: it must not leak into this image
            if rank == 0:
                r.wait_download()
            r.sync()
        except Exception as e:                      # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=rank_main, args=(k,)) for k in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not errors, errors
    reduced = rs[0].download_reduced_film()
    partial0 = rs[0].download_film()
    assert not np.allclose(reduced, partial0), "rank 0's own film must still be its partial film"
    # one GPU, whole frame, the same two passes
    r = rs[1]
    r.film_clear()
    for p in range(passes):
        r.render_pass(spp, frame_count=p * spp)
        alone = r.download_film()
        assert np.allclose(per_pass[p], alone, rtol=1e-4, atol=1e-5), (p, float(np.max(np.abs(per_pass[p] - alone))))
    assert np.all(alone[..., 3] > 0)
    assert np.allclose(reduced, alone, rtol=1e-4, atol=1e-5), float(np.max(np.abs(reduced - alone)))
    for c, rr in zip(comms, rs):
        rr.nccl_comm_destroy(c)
        rr.close()


def test_headless_driver_two_gpus(tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "driver", "bpt_headless")
    tables = os.path.join(ROOT, "buas_pathtracer_b200", "data", "sampler_tables.bin")
    outs = []
    for gpus in (1, 2):
        out = str(tmp_path / f"g{gpus}.bmp")
        p = subprocess.run([exe, "--tables", tables, "--scene", "icosphere", "--level", "5", "--w", "320", "--h", "180", "--spp", "8",
                            "--passes", "2", "--gpus", str(gpus), "--out", out], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        outs.append(np.frombuffer(open(out, "rb").read(), np.uint8))
    a, b = outs
    assert a.shape == b.shape and np.array_equal(a[:54], b[:54])
    d = np.abs(a[54:].astype(np.int32) - b[54:].astype(np.int32))
    assert d.max() <= 1 and np.count_nonzero(d) <= 0.01 * d.size
