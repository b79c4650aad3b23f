"""Golden-vector cases shared by tests/golden/make_golden.py (generator, runs the
reference) and tests/test_golden.py (checks the oracle and the CUDA path)."""
import numpy as np

HEAD = 3000      # samples kept from each end of a TF
STRIDE = 37      # and every STRIDE-th sample in between

POLY = [1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0]

# cfg keys follow oracle.OracleChain / Modulator keyword names; "fir": True = default taps
CASES = {
    "c1_tm1_native": dict(seed=101, n_tf=2, cfg=dict(mode=1)),
    "c4_tm2_native": dict(seed=102, n_tf=2, cfg=dict(mode=2)),
    "c4_tm3_native": dict(seed=103, n_tf=2, cfg=dict(mode=3)),
    "c4_tm4_native": dict(seed=104, n_tf=2, cfg=dict(mode=4)),
    "c2_tm1_fir": dict(seed=105, n_tf=2, cfg=dict(mode=1, fir=True)),
    "c3_tm1_fir_res8192_poly": dict(seed=106, n_tf=2, cfg=dict(mode=1, fir=True, output_rate=8192000,
                                                               normalise=1.0 / 46000.0, poly=POLY)),
    "c5_tm1_fir_res10000_poly": dict(seed=107, n_tf=3, cfg=dict(mode=1, fir=True, output_rate=10000000,
                                                                normalise=1.0 / 46000.0, poly=POLY)),
    "tm2_gain_max_s16": dict(seed=108, n_tf=2, cfg=dict(mode=2, gain_mode="max", digital_gain=0.8, fmt="s16")),
    "tm2_tii": dict(seed=109, n_tf=3, cfg=dict(mode=2, tii=(4, 17, 0))),
    "tm1_tii_old": dict(seed=110, n_tf=2, cfg=dict(mode=1, tii=(23, 69, 1), gain_mode="fix")),
    "tm4_window": dict(seed=111, n_tf=2, cfg=dict(mode=4, window_overlap=32)),
    "tm1_cfr": dict(seed=112, n_tf=2, cfg=dict(mode=1, cfr=(50.0, 0.1))),
    "tm1_ciceq": dict(seed=113, n_tf=1, cfg=dict(mode=1, clock_rate=32768000)),
    "tm2_res_down_1536": dict(seed=114, n_tf=3, cfg=dict(mode=2, fir=True, output_rate=1536000)),
    # row N4: the fixed-point engine (FFTEngine::KISS), int16 I/Q, bit-exact
    "fix_tm1": dict(seed=301, n_tf=2, cfg=dict(mode=1, fixed_point=True)),
    "fix_tm2_tii": dict(seed=302, n_tf=3, cfg=dict(mode=2, fixed_point=True, tii=(4, 17, 0))),
    "fix_tm3_window": dict(seed=303, n_tf=2, cfg=dict(mode=3, fixed_point=True, window_overlap=12)),
    "fix_tm4": dict(seed=304, n_tf=2, cfg=dict(mode=4, fixed_point=True)),
    "fix_tm1_window_tii": dict(seed=305, n_tf=2, cfg=dict(mode=1, fixed_point=True, window_overlap=40, tii=(23, 69, 1))),
}


def slices(o):
    """(head, tail, strided middle, checksums[sum_re, sum_im, sum |x|^2]) of one TF's output."""
    o = np.asarray(o)
    if np.iscomplexobj(o):
        v = o.astype(np.complex128)
        chk = np.array([v.real.sum(), v.imag.sum(), (np.abs(v) ** 2).sum()])
    else:
        v = o.astype(np.float64)
        chk = np.array([v[0::2].sum(), v[1::2].sum(), (v ** 2).sum()])
    return o[:HEAD].copy(), o[-HEAD:].copy(), o[HEAD:-HEAD:STRIDE].copy(), chk


# CFR read-outs ("clip_stats", "papr"): strings of the reference after `after` frames of the seeded stream
# (tests/golden/make_cfr_readouts.py -> cfr_readouts.json).  The PAPR windows fill after 50 frames.
READOUT_CASES = {
    "tm2_cfr_errclip": dict(seed=201, n_tf=56, after=[1, 9, 50, 51, 56], cfg=dict(mode=2, cfr=(40.0, 0.02))),
    "tm1_cfr_tii": dict(seed=202, n_tf=52, after=[3, 52], cfg=dict(mode=1, cfr=(80.0, 0.01), tii=(3, 5, 0))),
    "tm3_cfr_noerr": dict(seed=203, n_tf=12, after=[12], cfg=dict(mode=3, cfr=(30.0, 0.5), gain_mode="max")),
}
