"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): every rank renders its interleaved tile shard on its own
GPU, ONE NCCL reduce of the float3 accumulation framebuffer per spp batch onto rank 0, tonemap there — the image must
be bit-identical to the single-GPU render."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import gpu_pathtracer_b200 as pt
    s = pt.scenes.cornell_pt(512, 512, 8)
    with pt.PathTracer(s, device=rank, shard=(rank, world, 32, 32)) as r:
        # the collective lives in the library: torch.distributed only ships the 128-byte NCCL id
        ids = [pt.PathTracer.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.comm_init(world, rank, ids[0])
        tone = np.zeros((512, 512, 3), np.float32)
        for batch in range(2):
            r.render_reduce(1 + batch * spp, reset=(batch == 0), spp=spp, root=0, output=tone if rank == 0 else None)
        if rank == 0:
            np.savez(out_path, acc=r.reduced_accum(), tone=tone)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpus_one_nccl_reduce_per_batch_is_bit_identical(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import gpu_pathtracer_b200 as pt
    spp = 4
    out = str(tmp_path / "img.npz")
    mp.spawn(_worker, args=(2, _free_port(), spp, out), nprocs=2, join=True)
    got = np.load(out)
    s = pt.scenes.cornell_pt(512, 512, 8)
    with pt.PathTracer(s) as r:
        tone = r.render(1, reset=True, spp=2 * spp)
        acc = r.accum()
    assert np.array_equal(got["acc"].view(np.uint32), acc.view(np.uint32))
    assert np.array_equal(got["tone"].view(np.uint32), tone.view(np.uint32))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_single_process_multi_gpu_context_is_bit_identical():
    """b200pt_create_multi (what the BeginRender / Render adapter uses with B200PT_GPUS=N): one process, every GPU of the box,
    ncclCommInitAll + one grouped ncclReduce per batch — same bits as one GPU, for a surface and a heterogeneous-media scene."""
    sys.path.insert(0, ROOT)
    import gpu_pathtracer_b200 as pt
    n = min(torch.cuda.device_count(), 8)
    for s in (pt.scenes.cornell_pt(512, 512, 8), pt.scenes.cornell_smoke(256, 256, 8, 1)):
        with pt.PathTracer(s) as r:
            r.render(1, reset=True, spp=3)
            tone = r.render(4, reset=False, spp=3)
            acc = r.accum()
        with pt.MultiPathTracer(s, n) as m:
            m.render(1, reset=True, spp=3)
            tone_m = m.render(4, reset=False, spp=3)
            acc_m = m.accum()
            assert m.stats()["samples"] == 3 * s.width * s.height
        assert np.array_equal(acc_m.view(np.uint32), acc.view(np.uint32))
        assert np.array_equal(tone_m.view(np.uint32), tone.view(np.uint32))


def test_multi_context_api_on_one_gpu():
    """b200pt_create_multi with n_gpus = 1 is the plain context behind the multi entry points (no communicator)."""
    sys.path.insert(0, ROOT)
    import gpu_pathtracer_b200 as pt
    s = pt.scenes.cornell_pt(256, 256, 6)
    with pt.PathTracer(s) as r:
        tone = r.render(1, reset=True, spp=4); acc = r.accum()
    with pt.MultiPathTracer(s, 1) as m:
        tone_m = m.render(1, reset=True, spp=4); acc_m = m.accum()
    assert np.array_equal(acc_m.view(np.uint32), acc.view(np.uint32)) and np.array_equal(tone_m.view(np.uint32), tone.view(np.uint32))
