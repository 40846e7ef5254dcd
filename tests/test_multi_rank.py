"""CPU, world_size 2 over gloo: the N>1 path of bench.py / SURVEY 8(e).  Every rank renders its interleaved
32x32-tile shard through the C ABI (the CPU emulation build of the product sources stands in for the GPU library),
the float3 accumulation framebuffers are summed onto rank 0 with ONE reduce — the only collective of the path —
and the result must be bit-identical to the unsharded image (disjoint tiles: exactly one non-zero contributor
per pixel)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "libb200pt_emu.so")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene(pt, name):
    if name == "smoke":                                  # heterogeneous medium, sequential kernel, one lane
        return pt.scenes.cornell_smoke(128, 64, 6, 1)
    return pt.scenes.cornell_pt(128, 64, 6)


def _worker(rank, world, port, spp, out_path, scene_name):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gpu_pathtracer_b200 as pt
    from gpu_pathtracer_b200 import _lib
    _lib.load(EMU)
    s = _scene(pt, scene_name)
    with pt.PathTracer(s, shard=(rank, world, 32, 32)) as r:
        acc = None
        for batch in range(2):                                   # two spp batches, one reduce each (north_star)
            r.render(1 + batch * spp, reset=(batch == 0), spp=spp)
            acc = torch.from_numpy(r.accum())
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("scene_name", ["cornell", "smoke"])
def test_two_ranks_reduce_to_the_single_rank_image(tmp_path, scene_name):
    sys.path.insert(0, ROOT)
    import gpu_pathtracer_b200 as pt
    from gpu_pathtracer_b200 import _lib
    spp = 2
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), spp, out, scene_name), nprocs=2, join=True)
    got = np.load(out)
    saved = _lib._lib
    _lib.load(EMU)
    try:
        s = _scene(pt, scene_name)
        with pt.PathTracer(s) as r:
            r.render(1, reset=True, spp=2 * spp)
            full = r.accum()
    finally:
        _lib._lib = saved
    assert np.array_equal(got.view(np.uint32), full.view(np.uint32))
    assert full.mean() > 0.01
