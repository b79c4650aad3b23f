"""Multi-GPU path: frame-parallel sharding of one stream.

CPU part (world_size 2 and 3 over gloo): the host logic in
odr-dabmod_b200/sharding.py -- shard plan, halo TF for the resampler history, TII
parity, ordered result gather -- with the oracle standing in for the CUDA
modulator (the oracle is the checker here; the product object is the CUDA
Modulator, covered by the gpu-marked test at the bottom).
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
from conftest import rel_rms  # noqa: E402
from oracle import oracle  # noqa: E402


def sharding():
    dabmod_loader.load()
    import importlib
    return importlib.import_module("odr_dabmod_b200.sharding")


class OracleModulator:
    """seek()/process_batch() on top of the (sequential, stateful) oracle chain."""

    def __init__(self, **kw):
        self.kw = kw
        self.chain = oracle.OracleChain(**kw)
        self.max_batch = 2

    def seek(self, tf_index, prev_bits=None):
        self.chain = oracle.OracleChain(**self.kw)
        if tf_index == 0:
            return
        # The resampler history after a TF depends on that TF alone, the TII toggle on
        # the number of TFs seen: feed the halo TF once (odd index) or twice (even).
        assert prev_bits is not None
        for _ in range(1 if tf_index % 2 else 2):
            self.chain.process(prev_bits)

    def process_batch(self, bits):
        return np.stack(self.chain.run(bits))


CASES = {
    "tm2_res_tii": dict(mode=2, output_rate=4096000, tii=(4, 17, 0), fir_taps=oracle.fir_default_taps()),
    "tm2_native_s16": dict(mode=2, gain_mode="max", digital_gain=0.8, fmt="s16"),
}


def test_plan_shards():
    sh = sharding()
    for n_tf, world in [(16384, 8), (10, 3), (2, 4), (0, 2), (7, 1)]:
        plan = sh.plan_shards(n_tf, world)
        assert len(plan) == world
        assert [s.rank for s in plan] == list(range(world))
        assert sum(s.n_tf for s in plan) == n_tf
        first = 0
        for s in plan:
            assert s.first_tf == first
            first += s.n_tf
            assert s.halo_tf == (s.first_tf - 1 if s.first_tf > 0 and s.n_tf > 0 else None)
        assert max(s.n_tf for s in plan) - min(s.n_tf for s in plan) <= 1
    # C5 = 16384 TFs over 8 GPUs -> 2048 per GPU
    assert all(s.n_tf == 2048 for s in sh.plan_shards(16384, 8))
    with pytest.raises(ValueError):
        sh.plan_shards(4, 0)


@pytest.mark.parametrize("case", sorted(CASES))
def test_shards_reproduce_the_single_stream(case):
    """No process group needed: rank-by-rank, the shard outputs concatenate to the stream."""
    sh = sharding()
    kw = CASES[case]
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 256, (5, oracle.mode_params(kw["mode"]).tf_bytes), dtype=np.uint8)
    want = np.stack(oracle.OracleChain(**kw).run(bits))
    for world in (2, 3):
        parts = [sh.run_shard(OracleModulator(**kw), s, bits) for s in sh.plan_shards(5, world)]
        got = np.concatenate([p for p in parts if p is not None], axis=0)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, n_tf, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = sharding()
        kw = CASES[case]
        rng = np.random.default_rng(7)       # every rank sees the same synthetic stream
        bits = rng.integers(0, 256, (n_tf, oracle.mode_params(kw["mode"]).tf_bytes), dtype=np.uint8)
        plan = sh.plan_shards(n_tf, world)
        local = sh.run_shard(OracleModulator(**kw), plan[rank], bits)
        full = sh.gather_stream(local, plan, dist, dst=0)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_tf,case", [(2, 5, "tm2_res_tii"), (3, 2, "tm2_native_s16"), (2, 4, "tm2_native_s16")])
def test_gloo_sharded_run_and_gather(tmp_path, world, n_tf, case):
    """world_size > 1 over gloo: shard, run, gather to rank 0; equals the one-process stream.
    (3 ranks, 2 TFs: one rank has nothing to do and still takes part in the gather.)"""
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "stream.npy")
    mp.spawn(_worker, args=(world, _free_port(), case, n_tf, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    kw = CASES[case]
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 256, (n_tf, oracle.mode_params(kw["mode"]).tf_bytes), dtype=np.uint8)
    want = np.stack(oracle.OracleChain(**kw).run(bits))
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


class OracleEtiCoder:
    """coder.seek()/modulate() of the ETI-fronted shard (sharding.run_eti_shard) on top of the oracle coder."""

    def __init__(self, mode, streams, max_frames):
        self.mode, self.streams, self.max_frames = mode, streams, max_frames
        self.frames_per_tf = {1: 4, 2: 1, 3: 1, 4: 2}[mode]
        self.coder = oracle.OracleCoder(mode, streams)

    def seek(self, modulator, tf_index, frames_before):
        cif = self.frames_per_tf
        fb = np.asarray(frames_before, np.uint8).reshape(-1, 6144)
        want = min(tf_index * cif, 15 + cif)
        assert fb.shape[0] >= want, "history too short"
        self.coder = oracle.OracleCoder(self.mode, self.streams)
        if tf_index == 0:
            modulator.seek(0)
            return
        fb = fb[fb.shape[0] - want:]
        # A fresh oracle coder emits one block per `cif` frames fed; the history may not be a multiple of cif,
        # so pad in front with frames whose content cannot reach the last block (older than 15 frames).
        pad = (-fb.shape[0]) % cif
        blocks = self.coder.run(np.concatenate([np.repeat(fb[:1], pad, axis=0), fb]) if pad else fb)
        modulator.seek(tf_index, blocks[-1])

    def modulate(self, modulator, frames):
        return modulator.process_batch(np.stack(self.coder.run(frames)))


def _eti_case():
    dabmod_loader.load()
    import importlib
    eti = importlib.import_module("odr_dabmod_b200.eti")
    subch = [(0, 12, eti.eep_tpl(0, 1)), (60, 24, eti.eep_tpl(0, 3)), (200, 36, eti.eep_tpl(1, 2))]
    frames = eti.synth_eti_range(2, subch, 0, 23, seed=5)
    _, streams = oracle.describe_eti(frames[0])
    return frames, [s.as_tuple() for s in streams]


def _eti_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = sharding()
        frames, streams = _eti_case()
        plan = sh.plan_shards(frames.shape[0], world)
        local = sh.run_eti_shard(OracleModulator(**CASES["tm2_res_tii"]), OracleEtiCoder(2, streams, 4), plan[rank], frames)
        full = sh.gather_stream(local, plan, dist, dst=0)
        if rank == 0:
            np.save(out_path, full.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gloo_sharded_eti_stream(tmp_path):
    """The ETI-fronted shard (BASELINE configs[4]) under world_size 2: time-interleaver history, resampler halo and
    TII parity are re-established per rank from the frames before its range; gather equals the one-process stream."""
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "stream.npy")
    mp.spawn(_eti_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    got = np.load(out_path)
    frames, streams = _eti_case()
    blocks = np.stack(oracle.OracleCoder(2, streams).run(frames))
    want = np.stack(oracle.OracleChain(**CASES["tm2_res_tii"]).run(blocks))
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_cuda_shards_reproduce_the_single_stream(world):
    """The product object: `world` CUDA handles on one GPU play the ranks."""
    dm = dabmod_loader.load()
    sh = sharding()
    taps = oracle.fir_default_taps()
    kw = dict(mode=1, output_rate=8192000, tii=(1, 11, 0), fir_taps=taps)
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 256, (5, oracle.mode_params(1).tf_bytes), dtype=np.uint8)
    want = oracle.OracleChain(**kw).run(bits)
    single = dm.Modulator(max_batch=5, **kw).process_batch(bits)
    parts = []
    for s in sh.plan_shards(5, world):
        mod = dm.Modulator(max_batch=2, **kw)
        parts.append(sh.run_shard(mod, s, bits))
        mod.close()
    got = np.concatenate(parts, axis=0)
    assert got.shape == single.shape
    for i in range(5):
        assert rel_rms(got[i], want[i]) < 2e-6, i
    # same kernels, same per-TF arithmetic: bit-identical to the unsharded CUDA run
    assert np.array_equal(got.view(np.uint32), single.view(np.uint32))
