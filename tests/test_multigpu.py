"""Multi-GPU path proper (BASELINE configs[4]): one process per GPU over NCCL.

ONE ETI stream is cut into transmission-frame ranges; every rank positions its own coder + modulator with
dabmod_b200_seek_eti on its own GPU, produces its range, and the ranges are gathered in stream order to rank 0
over NCCL (sharding.gather_stream).  The gathered stream must be bit-identical to the same stream run unsharded on
one GPU.  Needs two GPUs: skipped below that (the driver's 1-GPU box runs the single-GPU shard tests in
test_sharding.py / test_coder.py; `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu` runs this one).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dabmod_loader  # noqa: E402

KW = dict(mode=1, output_rate=10000000, tii=(1, 11, 0), fir_taps="default", normalise=1.0 / 46000.0,
          poly=[1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0])
N_TF = 13


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stream():
    dabmod_loader.load()
    import importlib
    eti = importlib.import_module("odr_dabmod_b200.eti")
    return eti.synth_eti_range(1, eti.default_multiplex(), 0, 4 * N_TF, seed=77)


def _worker(rank, world, port, out_path):
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dm = dabmod_loader.load()
        sh = importlib.import_module("odr_dabmod_b200.sharding")
        frames = _stream()                      # every rank derives the same synthetic stream
        _, streams = dm.eti_describe(frames[0])
        plan = sh.plan_shards(N_TF, world)
        mod = dm.Modulator(max_batch=4, device=rank, **KW)
        cod = dm.Coder(1, streams, max_frames=16, device=rank)
        local = sh.run_eti_shard(mod, cod, plan[rank], frames)
        t = None if local is None else torch.from_numpy(local).to("cuda:%d" % rank)
        full = sh.gather_stream(t, plan, dist, dst=0, device=torch.device("cuda", rank))
        if rank == 0:
            assert full.is_cuda                 # the gather ran over NCCL on device tensors
            np.save(out_path, full.cpu().numpy())
        else:
            assert full is None
        dist.barrier()
        mod.close()
        cod.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_eti_stream_is_bit_identical_to_one_gpu(tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "stream.npy")
    mp.spawn(_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    got = np.load(out_path)
    dm = dabmod_loader.load()
    frames = _stream()
    _, streams = dm.eti_describe(frames[0])
    mod = dm.Modulator(max_batch=N_TF, device=0, **KW)
    cod = dm.Coder(1, streams, max_frames=4 * N_TF, device=0)
    want = cod.modulate(mod, frames)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
