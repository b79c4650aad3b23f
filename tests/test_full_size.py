"""BASELINE.json's full batch sizes on the GPU, checked through size-independent properties (the oracle needs
~10 ms per TF, so only sampled frames are compared with it):

  * a frame of a big batch equals the same frame processed alone, bit for bit (frames are independent except for
    the resampler history and the TII toggle, both functions of the stream position);
  * the host-buffer entry point (sliced three-stream pipeline) and the device-buffer entry point agree bit for bit;
  * the FIR is linear in its taps (a power-of-two scale is exact in float32);
  * a batch equals the same stream cut into several calls (resampler history across calls).
"""
import numpy as np
import pytest

import dabmod_loader
from conftest import rel_rms
from oracle import oracle

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def dm():
    return dabmod_loader.load()


def test_config2_full_batch(dm):
    """configs[1]: TM I, 1024 frames in one call, FIR default taps (k_symbols_w compact + k_fir_sym)."""
    import torch
    n = 1024
    rng = np.random.default_rng(2024)
    m = oracle.mode_params(1)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    taps = oracle.fir_default_taps()
    mod = dm.Modulator(mode=1, fir_taps=taps, max_batch=n)
    out = mod.process_batch(bits)
    assert out.shape == (n, m.tf_samples)
    # sampled frames: alone == in the batch, and against the oracle
    one = dm.Modulator(mode=1, fir_taps=taps, max_batch=1)
    for i in (0, 1, 31, 32, 511, 1023):                      # 31 | 32: a slice boundary of the host pipeline
        assert np.array_equal(one.process(bits[i]).view(np.uint32), out[i].view(np.uint32)), i
    ora = oracle.OracleChain(mode=1, fir_taps=taps)
    for i in (0, 1023):
        assert rel_rms(out[i], ora.process(bits[i])) < TOL, i
    # device entry point: same bits as the host pipeline (whole batch, word for word)
    d_in = torch.from_numpy(bits).cuda()
    d_out = torch.empty(n * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    mod.process_batch_device(d_in.data_ptr(), n, d_out.data_ptr())
    mod.synchronize()
    dev = d_out.view(torch.int32)
    host = torch.from_numpy(out.view(np.int32).reshape(-1)).cuda()
    assert bool(torch.equal(dev, host))
    # gain var: every frame's symbols are scaled to the same spread (GainControl.cpp:251-340), null symbol silent
    x = d_out.view(torch.float32).view(n, -1, 2)
    assert float(x[:, :2600].abs().max()) == 0.0
    p = (x[:, 4000:] ** 2).sum(dim=(1, 2)) / (x.shape[1] - 4000)
    assert float(p.max() / p.min()) < 1.05       # the gain follows max(sigma_re, sigma_im), not the total power
    # linearity in the taps: 2 * taps -> exactly 2 * output
    mod2 = dm.Modulator(mode=1, fir_taps=[2.0 * t for t in taps], max_batch=n)
    mod2.process_batch_device(d_in.data_ptr(), n, d_out.data_ptr())
    mod2.synchronize()
    assert bool(torch.equal(d_out.view(torch.float32), 2.0 * host.view(torch.float32)))


def test_fixed_point_full_batch(dm):
    """Row N4 at 1024 frames with TII: bit-exact, the TII symbol on every second frame of the stream."""
    n = 1024
    rng = np.random.default_rng(2025)
    m = oracle.mode_params(1)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    mod = dm.Modulator(mode=1, fixed_point=True, tii=(7, 21, 0), max_batch=n)
    out = mod.process_batch(bits)
    ora = oracle.OracleChain(mode=1, fixed_point=True, tii=(7, 21, 0))
    for i in range(4):                                       # the oracle walks the stream from its start
        assert np.array_equal(out[i], ora.process(bits[i])), i
    one = dm.Modulator(mode=1, fixed_point=True, tii=(7, 21, 0), max_batch=1)
    for i in (510, 511, 1022, 1023):
        one.seek(i)
        assert np.array_equal(one.process(bits[i]), out[i]), i
    null = out[:, : 2 * 2656].astype(np.int32)
    energy = (null ** 2).sum(axis=1)
    assert (energy[0::2] > 0).all() and (energy[1::2] == 0).all()


@pytest.mark.parametrize("mode,n", [(2, 4096), (3, 4096), (4, 2048)])
def test_config4_full_batches(dm, mode, n):
    """configs[3] at the benchmarked batch sizes (4096 / 4096 / 2048 frames, ~45 k CTAs or one persistent grid of
    warps): sampled frames alone == in the batch, bit for bit; host pipeline == device entry point over the whole
    batch; sampled frames against the oracle; every frame's null symbol silent and its data symbols at the same spread."""
    import torch
    rng = np.random.default_rng(3000 + mode)
    m = oracle.mode_params(mode)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    mod = dm.Modulator(mode=mode, max_batch=n)
    mod.set_param("profile", 1)
    d_in = torch.from_numpy(bits).cuda()
    d_out = torch.empty(n * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    mod.process_batch_device(d_in.data_ptr(), n, d_out.data_ptr())
    mod.synchronize()
    names = [k for k, _ in mod.kernel_times()]
    assert names == (["k_symbols"] if mode == 3 else ["k_symbols_wg"]), names
    out = d_out.cpu().numpy().view(np.complex64).reshape(n, -1)
    assert out.shape == (n, m.tf_samples)
    # the same kernel on a short call (a different split of the symbols over the warps): same bits
    few = dm.Modulator(mode=mode, max_batch=40)
    for i0 in (0, n // 2 - 20, n - 40):
        part = few.process_batch(bits[i0:i0 + 40])
        assert np.array_equal(part.view(np.uint32), out[i0:i0 + 40].view(np.uint32)), i0
    ora = oracle.OracleChain(mode=mode)
    for i in (0, n - 1):
        assert rel_rms(out[i], ora.process(bits[i])) < TOL, i
    # host entry point (sliced three-stream pipeline) == device entry point, word for word
    host = mod.process_batch(bits)
    assert np.array_equal(host.view(np.uint32), out.view(np.uint32))
    x = d_out.view(torch.float32).view(n, -1, 2)
    assert float(x[:, :m.null_size - 8].abs().max()) == 0.0
    p = (x[:, m.null_size + m.sym_size:] ** 2).sum(dim=(1, 2)) / (x.shape[1] - m.null_size - m.sym_size)
    assert float(p.max() / p.min()) < 1.10


@pytest.mark.parametrize("rate,n", [(8192000, 256), (10000000, 128)])
def test_config3_5_benchmarked_batches(dm, rate, n):
    """configs[2] / [4] at the benchmarked batch sizes (256 / 128 frames per call): the batch == the same stream in
    calls of 100 + 1 + the rest == a second handle that seeks, bit for bit; first and last frame against the oracle."""
    rng = np.random.default_rng(2028)
    m = oracle.mode_params(1)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    kw = dict(mode=1, output_rate=rate, normalise=1.0 / 46000.0, fir_taps=oracle.fir_default_taps(),
              poly=[1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0])
    mod = dm.Modulator(max_batch=n, **kw)
    whole = mod.process_batch(bits).copy()
    mod.reset()
    parts = np.concatenate([mod.process_batch(bits[:100]), mod.process_batch(bits[100:101]), mod.process_batch(bits[101:])])
    assert np.array_equal(parts.view(np.uint32), whole.view(np.uint32))
    del parts
    shard = dm.Modulator(max_batch=n, **kw)
    shard.seek(n - 7, bits[n - 8])
    assert np.array_equal(shard.process_batch(bits[n - 7:]).view(np.uint32), whole[n - 7:].view(np.uint32))
    ora = oracle.OracleChain(**kw)
    assert rel_rms(whole[0], ora.process(bits[0])) < TOL


@pytest.mark.parametrize("rate", [10000000, 8192000])
def test_config5_stream_in_pieces(dm, rate):
    """configs[4] / [2] geometry: FIR + resampler + MemlessPoly; 96 frames in one call == the same stream in calls of
    40 + 1 + 55 frames == a second handle that seeks to frame 41 (what a shard of the 8-GPU run does)."""
    n = 96
    rng = np.random.default_rng(2026)
    m = oracle.mode_params(1)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    kw = dict(mode=1, output_rate=rate, normalise=1.0 / 46000.0, fir_taps=oracle.fir_default_taps(),
              poly=[1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0])
    mod = dm.Modulator(max_batch=n, **kw)
    whole = mod.process_batch(bits).copy()
    mod.reset()
    parts = np.concatenate([mod.process_batch(bits[:40]), mod.process_batch(bits[40:41]), mod.process_batch(bits[41:])])
    assert np.array_equal(parts.view(np.uint32), whole.view(np.uint32))
    shard = dm.Modulator(max_batch=n, **kw)
    shard.seek(41, bits[40])
    tail = shard.process_batch(bits[41:])
    assert np.array_equal(tail.view(np.uint32), whole[41:].view(np.uint32))
    ora = oracle.OracleChain(**kw)
    for i in range(2):
        assert rel_rms(whole[i], ora.process(bits[i])) < TOL, i


def test_file_sink(dm, tmp_path):
    """Row N2: dabmod_b200_process_batch_to_fd == OutputFile behind the chain (OutputFile.cpp:56-67): the bytes in the
    file are the bytes process_batch returns, for a regular file and for a pipe (short writes), in several calls."""
    import os
    import threading
    n = 300                                                  # 14 slices of the sink's pipeline at s16
    rng = np.random.default_rng(2027)
    m = oracle.mode_params(1)
    bits = rng.integers(0, 256, (n, m.tf_bytes), dtype=np.uint8)
    kw = dict(mode=1, fir_taps=oracle.fir_default_taps(), fmt="s16", digital_gain=0.5, max_batch=n)
    want = dm.Modulator(**kw).process_batch(bits).tobytes()
    mod = dm.Modulator(**kw)
    path = str(tmp_path / "out.iq")
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        assert mod.process_batch_to_fd(bits[:100], fd) == 100 * mod.tf_out_bytes
        assert mod.process_batch_to_fd(bits[100:], fd) == 200 * mod.tf_out_bytes
    finally:
        os.close(fd)
    with open(path, "rb") as f:
        assert f.read() == want
    # a pipe: 64 KiB kernel buffer, so every slice goes out in many short writes
    mod.reset()
    r, w = os.pipe()
    got = bytearray()

    def reader():
        while True:
            b = os.read(r, 1 << 20)
            if not b:
                break
            got.extend(b)

    t = threading.Thread(target=reader)
    t.start()
    try:
        assert mod.process_batch_to_fd(bits, w) == len(want)
    finally:
        os.close(w)
        t.join()
        os.close(r)
    assert bytes(got) == want
    # a closed descriptor: the reference's OutputFile throws, here DABMOD_B200_EIO
    with pytest.raises(dm.DabModError) as e:
        mod.process_batch_to_fd(bits[:2], w)
    assert e.value.code == -6
