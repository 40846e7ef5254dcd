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
    npix = 512 * 512
    with pt.PathTracer(s, device=rank, shard=(rank, world, 32, 32)) as r:
        ptr = r.accum_device_ptr()

        class _Holder:
            __cuda_array_interface__ = {"shape": (npix * 3,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        acc_t = torch.as_tensor(_Holder(), device=f"cuda:{rank}")
        full = torch.empty_like(acc_t)
        out = torch.empty_like(acc_t)
        for batch in range(2):
            r.render(1 + batch * spp, reset=(batch == 0), spp=spp)
            full.copy_(acc_t)
            dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            torch.cuda.current_stream().synchronize()
            r.tonemap_device(full.data_ptr(), 2 * spp, out.data_ptr())
            np.savez(out_path, acc=full.cpu().numpy().reshape(512, 512, 3), tone=out.cpu().numpy().reshape(512, 512, 3))
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
