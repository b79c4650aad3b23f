"""Frame-parallel sharding of one modulator stream over several GPUs (host logic).

The hot chain is stateless across transmission frames (TFs) except for two
things:

  * the TII block inserts its symbol on every second TF (reference
    src/TII.cpp:225-242), a function of the TF index alone;
  * the Resampler keeps half a block of input and of output (reference
    src/Resampler.cpp:143-145,185-191), a pure function of the last Ni samples
    of the previous TF's FIR-stage output.

So a stream of n_tf TFs is cut into contiguous TF ranges, one per rank; rank r
positions its modulator with `seek(first_tf, bits_of_tf[first_tf - 1])`, which
re-runs that one halo TF up to the resampler input, and then produces exactly
the bytes a single modulator would have produced for its range.  There is no
collective on the data path.  `gather_stream` is the optional result gather
("NCCL only for result gather"): every rank contributes its slice, the
destination rank receives the stream in order.

The compute object is whatever implements `seek(tf_index, prev_bits)` and
`process_batch(bits) -> (n, samples) array`; in the product that is
`odr_dabmod_b200.Modulator` (CUDA).  This module has no compute of its own.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    first_tf: int     # index of the first TF of this rank within the stream
    n_tf: int         # number of TFs (may be 0 when world > n_tf)

    @property
    def halo_tf(self):
        """Index of the TF whose bits prime the resampler history, or None at stream start."""
        return self.first_tf - 1 if self.first_tf > 0 and self.n_tf > 0 else None


def plan_shards(n_tf, world):
    """Contiguous, balanced TF ranges: the first n_tf % world ranks get one TF more."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if n_tf < 0:
        raise ValueError("n_tf must be >= 0")
    base, extra = divmod(n_tf, world)
    shards, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        shards.append(Shard(r, first, n))
        first += n
    return shards


def run_shard(modulator, shard, stream_bits):
    """Processes this rank's range of `stream_bits` ((n_tf_total, tf_in_bytes) uint8, or any
    object whose [a:b] slicing yields such rows, e.g. a memory-mapped file).
    Returns the (shard.n_tf, samples) output array."""
    if shard.n_tf == 0:
        return None
    halo = shard.halo_tf
    modulator.seek(shard.first_tf, None if halo is None else np.ascontiguousarray(stream_bits[halo]))
    rows = np.ascontiguousarray(stream_bits[shard.first_tf:shard.first_tf + shard.n_tf])
    step = getattr(modulator, "max_batch", None) or shard.n_tf
    outs = [modulator.process_batch(rows[i:i + step]) for i in range(0, shard.n_tf, step)]
    return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)


def run_eti_shard(modulator, coder, shard, stream_frames, n_before=None):
    """The same with the channel coding in front (BASELINE configs[4]: one ETI stream sharded by frame range):
    `stream_frames` = (n_eti_frames, 6144) uint8 (or anything sliceable like it).  The shard is positioned with
    `coder.seek(modulator, first_tf, frames before it)`: 15 ETI frames of time-interleaver history
    (src/TimeInterleaver.cpp:39-41) plus the transmission frame before the shard, which is coded and re-run up to
    the resampler input (src/Resampler.cpp:143-145,185-191); first_tf sets the TII toggle (src/TII.cpp:225-242)."""
    if shard.n_tf == 0:
        return None
    cif = coder.frames_per_tf
    f0 = shard.first_tf * cif
    hist = min(f0, 15 + cif) if n_before is None else n_before
    coder.seek(modulator, shard.first_tf, np.ascontiguousarray(stream_frames[f0 - hist:f0]))
    step = min(getattr(modulator, "max_batch", shard.n_tf), coder.max_frames // cif)
    outs = []
    for t in range(0, shard.n_tf, step):
        n = min(step, shard.n_tf - t)
        outs.append(coder.modulate(modulator, np.ascontiguousarray(stream_frames[f0 + t * cif:f0 + (t + n) * cif])))
    return outs[0] if len(outs) == 1 else np.concatenate(outs, axis=0)


def gather_stream(local_out, shards, dist, dst=0, device=None):
    """Gathers the per-rank outputs to rank `dst` in stream order over `dist`
    (an initialised torch.distributed; NCCL on GPUs, gloo in the CPU tests).
    local_out: numpy array or torch tensor of shape (shards[rank].n_tf, samples) (None if n_tf == 0).
    Returns the (n_tf_total, samples) torch tensor on rank dst, None elsewhere."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    assert len(shards) == world
    # agree on row length and dtype (a rank with no TFs has no array to look at)
    meta = [None] * world
    mine = None
    if local_out is not None:
        t = local_out if isinstance(local_out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_out))
        mine = (int(t.shape[1]), str(t.dtype))
    dist.all_gather_object(meta, mine)
    known = [m for m in meta if m is not None]
    if not known:
        return None
    if any(m != known[0] for m in known):
        raise RuntimeError("ranks disagree on the output row layout: %r" % (meta,))
    cols, dtype = known[0][0], getattr(torch, known[0][1].split(".")[-1])
    if local_out is None:
        t = torch.empty((0, cols), dtype=dtype)
    # rows travel as raw bytes: the collective backends do not take every dtype
    # (gloo: no complex64, no int16), and a gather does no arithmetic
    t = t.contiguous().view(torch.uint8).reshape(t.shape[0], cols * t.element_size())
    if device is not None:
        t = t.to(device)
    # ranks may hold different TF counts: pad to the largest, trim after the gather
    n_max = max(s.n_tf for s in shards)
    pad = torch.zeros((n_max, t.shape[1]), dtype=torch.uint8, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.cat([b[:s.n_tf] for b, s in zip(bufs, shards)], dim=0)
    out = out.contiguous().view(dtype).reshape(out.shape[0], cols)
    return out
