"""N>1 host logic on CPU: two gloo ranks partition the image into row tiles exactly as bench.py
does, take their slice of the make_trace_state rng table, and exchange padded tiles with an
all-gather laid out like ygl_gather_image (equal chunks of ceil(H/N) rows) — the reassembled
buffers must equal the single-rank ones. (The NCCL collective itself runs in the GPU scaling run.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "yocto-gl_b200"))
    from ygl_b200 import abi, lib, scenes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = scenes.cornellbox()
        params = abi.trace_params(resolution=37, samples=1)  # 37 rows: ragged last tile
        w, h, rngs = lib.make_state_rngs(scene, params)
        rb, re = lib.tile_rows(h, rank, world)
        per = (h + world - 1) // world
        # this rank's "tile": its rng slice, padded to `per` rows like the NCCL send buffer
        send = np.zeros((per * w, 2), np.int64)
        mine = rngs[rb * w:re * w].view(np.int64)
        send[:len(mine)] = mine
        out = [torch.zeros(per * w, 2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(send))
        full = torch.cat(out).numpy()[:h * w].view(np.uint64)
        ok = np.array_equal(full, rngs)
        # every rank agrees on the partition
        rows = torch.tensor([rb, re])
        allrows = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allrows, rows)
        cover = allrows[0][0].item() == 0 and allrows[-1][1].item() == h and all(
            allrows[i][1] == allrows[i + 1][0] for i in range(world - 1))
        t = torch.tensor([int(ok and cover)])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(int(t.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tiles_and_gather_layout_gloo(world):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret.get(timeout=10) == 1
