#!/usr/bin/env python
"""Benchmark of the COFDM I/Q hot path (BASELINE.json metric: ETI frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): TM I, 1024 transmission frames (= 4096 ETI
frames) per step and per GPU, native 2.048 Msps, FIRFilter with the default
taps, complexf output.  One step = one pass of the whole kernel family over
that batch.  Synthetic input: uniform random bytes of the BlockPartitioner
block size (every bit pattern is a valid QPSK input).

  value  device-resident ETI frames/s (inputs and outputs in HBM), CUDA events
  e2e    the same through dabmod_b200_process_batch with pinned HOST buffers
         (H2D of the bits and D2H of the I/Q inside the timed region)

N > 1: one process per GPU (torchrun), frame-sharded weak scaling, no data-path
collective; time = max over ranks of the device time.

--impl reference: the reference's own CPU code (oracle/_ref/libdabmod_ref.so =
unmodified ODR-DabMod blocks, FFTW replaced by its vendored float KISS FFT),
one independent modulator per host core.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE = 1
TFS_PER_STEP = 1024
ETI_PER_TF = 4
TF_IN_BYTES = 28800
TF_SAMPLES = 196608
# algorithmic bytes per TM I TF
SYM_BYTES_PER_TF = TF_IN_BYTES + TF_SAMPLES * 8          # 1 601 664
FIR_BYTES_PER_TF = 2 * TF_SAMPLES * 8                    # 3 145 728
# k_fir_sym reads the symbol kernel's compact layout (76 symbols x 2048 samples)
COMPACT_SAMPLES = 76 * 2048                              # 155 648
FIR_SYM_BYTES_PER_TF = (COMPACT_SAMPLES + TF_SAMPLES) * 8   # 2 818 048
ETI_PER_TF_MODE = {1: 4, 2: 1, 3: 1, 4: 2}
FP32_PEAK_TFLOPS = 72.0                                   # measured FFMA rate of this B200 pool (tools/ubench/fp32_pipes.cu)
METRIC = "ETI frames/sec (TM I, 2.048 Msps I/Q) at 1/2/4/8 B200 vs reference CPU"
WORKLOAD = "TM I, batched 1024 frames on 1xB200, native rate, FIRFilter enabled (default taps)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]      # the idle->busy ramp lands in the lower half
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference / CPU baseline
# ---------------------------------------------------------------------------
def ref_kwargs(kw, tmpdir="/tmp"):
    """Modulator(**kw) keyword set -> RefChain(**...) keyword set (the same configuration for the reference)."""
    r = dict(mode=kw.get("mode", 1))
    for k in ("output_rate", "normalise", "digital_gain", "gain_mode", "fixed_point", "fmt"):
        if k in kw:
            r[k] = kw[k]
    if kw.get("fir_taps") is not None:
        r["fir_taps_file"] = "default"
    if kw.get("poly") is not None:
        path = os.path.join(tmpdir, "bench_poly_%d.coef" % os.getpid())
        with open(path, "w") as f:
            f.write("1\n5\n" + "\n".join("%.9g" % c for c in kw["poly"]) + "\n")
        r["poly_coef_file"] = path
        r["poly_threads"] = 1
    return r


def cpu_chain_factory(kw=None, variant=None):
    """Returns (kind, make_chain) for the CPU implementation of the path in configuration `kw`
    (default: the headline workload)."""
    from oracle import refwrap
    kw = kw if kw is not None else dict(mode=MODE, fir_taps="default")
    if refwrap.available(variant):
        rk = ref_kwargs(kw)

        def mk():
            return refwrap.RefChain(variant=variant, **rk)
        return "reference", mk
    if variant is not None:
        raise RuntimeError("reference library variant %r not built" % variant)
    from oracle import oracle

    def mk2():
        return oracle.OracleChain(mode=kw.get("mode", 1), fir_taps=oracle.fir_default_taps() if kw.get("fir_taps") else None)
    return "port", mk2


def run_cpu(n_threads, tfs_per_thread, seed=99, kw=None, variant=None):
    """n_threads independent modulators (one per host core), each fed tfs_per_thread TFs.
    Returns (eti_frames_per_s, seconds, kind)."""
    kind, mk = cpu_chain_factory(kw, variant)
    mode = (kw or {}).get("mode", MODE)
    rng = np.random.default_rng(seed)
    tf_bytes = {1: 28800, 2: 7200, 3: 7296, 4: 14400}[mode]
    bits = rng.integers(0, 256, (8, tf_bytes), dtype=np.uint8)
    chains = [mk() for _ in range(n_threads)]
    feed = (lambda c, b: c.feed_raw(b)) if kind == "reference" else (lambda c, b: c.process(b))
    # prime the pipelined stages (reference: up to 3 calls of latency) outside the timed region
    for c in chains:
        for i in range(4):
            feed(c, bits[i % 8])
    barrier = threading.Barrier(n_threads + 1)

    def work(c):
        barrier.wait()
        for i in range(tfs_per_thread):
            feed(c, bits[i % 8])
        barrier.wait()

    ths = [threading.Thread(target=work, args=(c,)) for c in chains]
    for t in ths:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in ths:
        t.join()
    for c in chains:
        c.close()
    return n_threads * tfs_per_thread * ETI_PER_TF_MODE[mode] / dt, dt, kind


def cpu_reference_for(kw, tfs_per_thread):
    """The reference's own CPU code on this box in configuration `kw`: one modulator on one core (the reference's
    native shape: main thread + its pipelined stages), one independent modulator per core, and the -ffast-math build
    (BASELINE.md section 3)."""
    from oracle import refwrap
    if not refwrap.available():
        return None
    cores = host_cores()
    out = {"cores": cores, "kind": "reference",
           "sample": "%d TFs per modulator; FFTW replaced by the reference's vendored float KISS FFT" % tfs_per_thread}
    try:
        v, dt, _ = run_cpu(1, tfs_per_thread, kw=kw)
        out["one_modulator_eti_frames_per_s"] = v
        v, dt, _ = run_cpu(cores, tfs_per_thread, kw=kw)
        out["all_cores_eti_frames_per_s"] = v
        out["all_cores_seconds"] = dt
        if refwrap.available("fast"):
            v, dt, _ = run_cpu(cores, tfs_per_thread, kw=kw, variant="fast")
            out["all_cores_ffast_math_eti_frames_per_s"] = v
    except Exception as e:
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    tfs = 96                                   # per thread per step: ~0.4 s of CPU work
    for _ in range(args.warmup):
        run_cpu(cores, tfs)
    t_total, frames = 0.0, 0
    kind = "reference"
    for _ in range(args.steps):
        v, dt, kind = run_cpu(cores, tfs)
        t_total += dt
        frames += cores * tfs * ETI_PER_TF
    value = frames / t_total
    sample = "%d TFs per thread per step x %d threads (one independent modulator per host core)" % (tfs, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ETI frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mode": MODE, "tfs_per_step": cores * tfs,
                   "note": "CPU arm: bounded sample of the same workload"},
        "cpu_baseline": {"value": value, "unit": "ETI frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "ETI frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
POLY = [1.0, 0.05, -0.02, 0.003, 0.0, 0.0, 0.1, -0.05, 0.01, 0.0]   # 5 + 5 odd-polynomial terms (config 3)
# algorithmic bytes / flops per TM I TF of the resampling stages
RES_INFO = {
    8192000: {"out_samples": 786432, "flop": 133.7e6},
    10000000: {"out_samples": 960000, "flop": 160.6e6},
}


def other_configs():
    """The BASELINE.json configs that are not the headline workload (device-resident only)."""
    return [
        ("c1 TM I native, no FIR", dict(mode=1), 1024),
        ("c3 TM I FIR + resample 8.192 Msps + poly(5)", dict(mode=1, fir_taps="default", output_rate=8192000,
                                                             normalise=1.0 / 46000.0, poly=POLY), 256),
        ("c5 TM I FIR + resample 10 Msps + poly(5), one GPU's share", dict(mode=1, fir_taps="default",
                                                                           output_rate=10000000,
                                                                           normalise=1.0 / 46000.0, poly=POLY), 128),
        ("c3s TM I FIR + resample 8.192 Msps + poly(5), s16 out", dict(mode=1, fir_taps="default", output_rate=8192000,
                                                                       normalise=1.0 / 46000.0, poly=POLY, fmt="s16"), 256),
        ("c4 TM II native", dict(mode=2), 4096),
        ("c4 TM III native", dict(mode=3), 4096),
        ("c4 TM IV native", dict(mode=4), 2048),
        ("n4 TM I fixed-point engine (FFTEngine::KISS), int16 I/Q", dict(mode=1, fixed_point=True), 1024),
    ]


CPU_TFS = {"c1": 64, "c3 ": 10, "c5": 8, "c3s": 10, "c4 TM II ": 256, "c4 TM III": 256, "c4 TM IV": 128, "n4": 64}


def measure_config(dm, torch, name, kw, n_tf, stream, peak, steps=5, warmup=3, with_cpu=False, with_e2e=True):
    """Device-resident throughput, per-kernel times, host-delivered throughput and the reference's CPU throughput
    of one configuration."""
    mod = dm.Modulator(max_batch=n_tf, **kw)
    g = torch.Generator(device="cpu").manual_seed(4321)
    bits = torch.randint(0, 256, (n_tf, mod.tf_in_bytes), dtype=torch.uint8, generator=g).to("cuda")
    out = torch.empty(n_tf * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(warmup):
        mod.process_batch_device(bits.data_ptr(), n_tf, out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        mod.process_batch_device(bits.data_ptr(), n_tf, out.data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mod.set_param("profile", 1)
    kt = {}
    for _ in range(3):
        mod.process_batch_device(bits.data_ptr(), n_tf, out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        for k, t in mod.kernel_times():
            kt.setdefault(k, []).append(t)
    mode = kw.get("mode", 1)
    res = {"workload": name, "tfs_per_step": n_tf, "ms_per_step": ms,
           "eti_frames_per_s": n_tf * ETI_PER_TF_MODE[mode] / (ms * 1e-3),
           "out_bytes_per_step": n_tf * mod.tf_out_bytes,
           "kernels_ms": {k: float(np.mean(v)) for k, v in kt.items()}}
    mod.set_param("profile", 0)
    if with_e2e:
        # host-delivered: the same batch through dabmod_b200_process_batch, pinned host buffers both ways
        try:
            ob = n_tf * mod.tf_out_bytes
            h_in = bits.cpu().pin_memory()
            h_out = torch.empty(ob, dtype=torch.uint8).pin_memory()
            mod.process_batch_ptr(h_in.data_ptr(), n_tf, h_out.data_ptr(), ob)
            torch.cuda.synchronize()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                mod.process_batch_ptr(h_in.data_ptr(), n_tf, h_out.data_ptr(), ob)
            dt = (time.perf_counter() - t0) / reps
            res["e2e"] = {"eti_frames_per_s": n_tf * ETI_PER_TF_MODE[mode] / dt, "ms_per_step": dt * 1e3,
                          "d2h_GB/s": ob / dt / 1e9, "h2d_bytes_per_step": n_tf * mod.tf_in_bytes,
                          "d2h_bytes_per_step": ob, "checksum": int(h_out[::65537].to(torch.int64).sum().item())}
            del h_in, h_out
        except Exception as e:
            res["e2e"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if with_cpu:
        key = [k for k in CPU_TFS if name.startswith(k)]
        cpu = cpu_reference_for(kw, CPU_TFS[key[0]] if key else 16)
        if cpu is not None:
            res["cpu_reference"] = cpu
            if "all_cores_eti_frames_per_s" in cpu:
                res["speedup_vs_all_cores"] = {"device_resident": res["eti_frames_per_s"] / cpu["all_cores_eti_frames_per_s"],
                                               "host_delivered": (res.get("e2e", {}).get("eti_frames_per_s", 0) /
                                                                  cpu["all_cores_eti_frames_per_s"])}
    rate = kw.get("output_rate", 2048000)
    if rate in RES_INFO and mode == 1:
        k = [x for x in kt if x.startswith("k_resample")][0]
        t = float(np.mean(kt[k])) * 1e-3
        info = RES_INFO[rate]
        byt = n_tf * (TF_SAMPLES * 8 + info["out_samples"] * (mod.tf_out_bytes // mod.tf_out_samples))
        fir_t = sum(float(np.mean(v)) for kk, v in kt.items() if kk.startswith("k_fir")) * 1e-3
        stage_bytes = n_tf * (TF_SAMPLES * 8 + info["out_samples"] * (mod.tf_out_bytes // mod.tf_out_samples))
        stage_flop = n_tf * (info["flop"] + 35.4e6 + 40.0 * info["out_samples"])     # resampler + FIR + poly, SURVEY 8(d)
        res["resampler"] = {"kernel": k, "ms": t * 1e3, "algorithmic_GB/s": byt / t / 1e9,
                            "frac_of_hbm_peak": byt / t / 1e9 / peak,
                            "algorithmic_TFLOP/s_fp32": n_tf * info["flop"] / t / 1e12,
                            "frac_of_fp32_peak": n_tf * info["flop"] / t / 1e12 / FP32_PEAK_TFLOPS,
                            "note": "FP32-bound stage (17-19 flop/B > the 11 flop/B ridge): see DESIGN.md"}
        res["fir_resample_stage"] = {"ms": (fir_t + t) * 1e3, "algorithmic_bytes": stage_bytes,
                                     "frac_of_hbm_peak": stage_bytes / (fir_t + t) / 1e9 / peak,
                                     "algorithmic_TFLOP/s_fp32": stage_flop / (fir_t + t) / 1e12,
                                     "frac_of_fp32_peak": stage_flop / (fir_t + t) / 1e12 / FP32_PEAK_TFLOPS,
                                     "fp32_peak_TFLOP/s": FP32_PEAK_TFLOPS,
                                     "fp32_peak_source": "measured FFMA throughput, profiles/r01_ubench_fp32_pipes.txt",
                                     "note": "north_star stage: FIRFilter + Resampler + MemlessPoly, against both roofs"}
    if kw.get("fixed_point"):
        t = float(np.mean(kt["k_symbols_fix"])) * 1e-3
        byt = n_tf * (mod.tf_in_bytes + mod.tf_out_bytes)        # bits in, int16 I/Q out
        res["symbols_fix"] = {"ms": t * 1e3, "algorithmic_GB/s": byt / t / 1e9, "frac_of_hbm_peak": byt / t / 1e9 / peak,
                              "note": "integer kernel (KISS FIXED_POINT=16 arithmetic, bit-exact), ALU/LSU bound"}
    mod.close()
    del bits, out
    torch.cuda.empty_cache()
    return res


def measure_binary(n_frames=4000):
    """BASELINE configs[0]: the real program, ETI file in -> I/Q file out (TM I, native rate, complexf), through
    InputFileReader -> EtiReader -> the reference's channel coding graph -> [chain] -> OutputFile(/dev/null):
    oracle/_ref/odr-dabmod-ref (every source unmodified, its own threads) against oracle/_ref/odr-dabmod-b200
    (fft_engine = b200: the same program with the chain behind BlockPartitioner on the GPU), one TF per call and with
    the adapter's 64-TF pipeline.  In both the channel coding runs on ONE host thread (src/DabMod.cpp:593-738)."""
    import importlib
    import tempfile
    eti = importlib.import_module("odr_dabmod_b200.eti")
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-ref")
    b200_bin = os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-b200")
    if not (os.path.exists(ref_bin) and os.path.exists(b200_bin)):
        return {"workload": "c0 binary", "error": "oracle/_ref binaries not built"}
    tmp = tempfile.mkdtemp(prefix="dabmod_bench_")
    path = os.path.join(tmp, "in.eti")
    eti.synth_eti_range(1, eti.default_multiplex(), 0, n_frames, seed=8).tofile(path)
    ini = ("[remotecontrol]\nzmqctrl=0\ntelnet=0\n[log]\nsyslog=0\n[input]\ntransport=file\nsource=%s\nloop=0\n"
           "[modulator]\nfft_engine=%s\ngainmode=var\nmode=1\nrate=2048000\n[firfilter]\nenabled=0\n"
           "[output]\noutput=file\n[fileoutput]\nformat=complexf\nfilename=/dev/null\n")
    res = {"workload": "c0 TM I, ETI file input, native 2.048 Msps, no FIR/resample, file output: the real binary",
           "eti_frames": n_frames}
    # the long input: the same file thirty times over (4000 frames = 16 FCT periods, so the frame counter stays continuous)
    long_path, long_n = os.path.join(tmp, "in_long.eti"), 30 * n_frames
    with open(path, "rb") as f:
        blob = f.read()
    with open(long_path, "wb") as f:
        for _ in range(30):
            f.write(blob)
    del blob
    out_file = os.path.join(tmp, "out.iq")
    runs = (
        # key, binary, engine, depth, input, frames, format, output
        ("reference_binary", ref_bin, "fftw", 0, path, n_frames, "complexf", "/dev/null"),
        ("b200_binary", b200_bin, "b200", 0, path, n_frames, "complexf", "/dev/null"),
        ("b200_binary_depth64", b200_bin, "b200", 64, path, n_frames, "complexf", "/dev/null"),
        # boundary shape B1: DabModulator.cpp / ConfigParser.cpp / DabMod.cpp unmodified, the hot path's translation
        # units substituted (adapter/B200Blocks.cpp), the reference's own configuration file
        ("b1_binary", os.path.join(ROOT, "oracle", "_ref", "odr-dabmod-b1"), "fftw", 0, path, n_frames, "complexf", "/dev/null"),
        ("b200_eti_binary", b200_bin, "b200_eti", 64, path, n_frames, "complexf", "/dev/null"),
        ("b200_eti_binary_long", b200_bin, "b200_eti", 64, long_path, long_n, "complexf", "/dev/null"),
        ("b200_eti_binary_long_s16", b200_bin, "b200_eti", 64, long_path, long_n, "s16", "/dev/null"),
        ("b200_eti_binary_long_u8", b200_bin, "b200_eti", 64, long_path, long_n, "u8", "/dev/null"),
        ("reference_binary_to_file", ref_bin, "fftw", 0, path, n_frames, "complexf", out_file),
        ("b200_eti_binary_to_file", b200_bin, "b200_eti", 64, path, n_frames, "complexf", out_file),
    )
    for key, binary, engine, depth, src, nfr, fmt, dst in runs:
        cfg = os.path.join(tmp, key + ".ini")
        with open(cfg, "w") as f:
            f.write((ini % (src, engine)).replace("format=complexf", "format=" + fmt).replace("filename=/dev/null",
                                                                                             "filename=" + dst))
        env = dict(os.environ, ODR_DABMOD_B200_DEPTH=str(depth), ODR_DABMOD_B200_TRACE="1")
        if not os.path.exists(binary):
            res[key] = {"error": "not built"}
            continue
        try:
            t0 = time.perf_counter()
            r = subprocess.run([binary, cfg], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                               timeout=300)
            dt = time.perf_counter() - t0
            res[key] = {"eti_frames_per_s": nfr / dt, "seconds": dt, "eti_frames": nfr, "returncode": r.returncode}
            # the ETI engines report the time from the first frame to the end of the stream (ODR_DABMOD_B200_TRACE):
            # the rate without process start-up and CUDA context creation (0.7 - 2.8 s, as large as the run itself)
            import re
            m = re.search(r"B200EtiChain: (\d+) frames in ([0-9.]+) s since the first frame", r.stdout)
            if m and float(m.group(2)) > 0:
                res[key]["eti_frames_per_s_streaming"] = int(m.group(1)) / float(m.group(2))
        except Exception as e:
            res[key] = {"error": "%s: %s" % (type(e).__name__, e)}
        if dst != "/dev/null" and os.path.exists(dst):
            res[key]["output_bytes"] = os.path.getsize(dst)
            os.remove(dst)
    res["note"] = ("wall clock of the whole process incl. start-up (CUDA context creation ~0.3 s for the b200 arms); "
                   "reference_binary and b200_binary are bound by the reference's single-threaded EtiReader + channel "
                   "coding (~0.25 ms per ETI frame); b200_eti = fft_engine b200_eti: coding and chain on the GPU, "
                   "the host keeps InputFileReader, EtiReader and OutputFile (the *_to_file runs write a real file in "
                   "the temporary directory)")
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return res


def measure_traffic(kernel):
    """dram__bytes_read + dram__bytes_write of one launch of `kernel` in the headline step, from an ncu capture taken
    by this run (a child process under ncu; nothing under the profiler is timed)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:^%s" % kernel, "-s", "3", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__),
           "--traffic-child"]
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
    except Exception as e:
        return None, "%s: %s" % (type(e).__name__, e)
    total, seen = 0.0, 0
    import csv
    import io
    rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 4]
    hdr = next((row for row in rows if "Metric Name" in row), None)
    if hdr is None:
        return None, "no ncu csv (%s)" % r.stdout[-200:].replace("\n", " ")
    iname, iunit, ival = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for row in rows:
        if row is hdr or len(row) <= ival or not row[iname].startswith("dram__bytes"):
            continue
        total += float(row[ival].replace(",", "")) * scale.get(row[iunit], 1.0)
        seen += 1
    return (total, "ncu capture of this run") if seen == 2 else (None, "ncu metrics missing")


def traffic_child():
    """The headline step a few times, device-resident (run under ncu by measure_traffic)."""
    import torch
    import dabmod_loader
    dm = dabmod_loader.load()
    n_tf = TFS_PER_STEP
    mod = dm.Modulator(mode=MODE, fir_taps="default", max_batch=n_tf)
    bits = torch.randint(0, 256, (n_tf, mod.tf_in_bytes), dtype=torch.uint8).cuda()
    out = torch.empty(n_tf * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    for _ in range(6):
        mod.process_batch_device(bits.data_ptr(), n_tf, out.data_ptr(), 0)
    torch.cuda.synchronize()
    mod.close()
    return 0


def measure_coder(dm, torch, stream, n_tf=1024, steps=5, warmup=3, with_cpu=True):
    """Row N1: ETI frames -> blocks (coder alone) and ETI frames -> I/Q (coder + TM I FIR chain),
    device-resident, plus the reference's coding graph on one host core."""
    import importlib
    eti = importlib.import_module("odr_dabmod_b200.eti")
    n_frames = n_tf * ETI_PER_TF
    frames = eti.synth_eti(1, eti.default_multiplex(), n_frames, seed=11)
    mode, streams = dm.eti_describe(frames[0])
    cod = dm.Coder(mode, streams, max_frames=n_frames)
    mod = dm.Modulator(mode=1, fir_taps="default", max_batch=n_tf)
    d_eti = torch.from_numpy(frames).to("cuda")
    d_bits = torch.empty(n_tf * cod.tf_bytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n_tf * mod.tf_out_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def run(chain):
        cod.process_device(d_eti.data_ptr(), n_frames, d_bits.data_ptr(), stream.cuda_stream)
        if chain:
            mod.process_batch_device(d_bits.data_ptr(), n_tf, d_out.data_ptr(), stream.cuda_stream)

    res = {"workload": "n1 TM I, 6 x 128 kbit/s EEP 3-A + FIC: ETI(NI) frames in", "eti_frames_per_step": n_frames}
    for key, chain in (("coder_only", False), ("eti_to_iq", True)):
        for _ in range(warmup):
            run(chain)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            run(chain)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[key] = {"ms_per_step": ms, "eti_frames_per_s": n_frames / (ms * 1e-3)}
    # algorithmic traffic of the coder: ETI frame in, block out
    res["coder_only"]["algorithmic_GB/s"] = n_frames * (6144 + 7200) / (res["coder_only"]["ms_per_step"] * 1e-3) / 1e9
    if with_cpu:
        try:
            from oracle import refwrap
            if refwrap.available():
                ref = refwrap.RefCoder()
                n_cpu = 2000
                t0 = time.perf_counter()
                ref.run(frames[:n_cpu])
                dt = time.perf_counter() - t0
                res["cpu_reference_coder"] = {"eti_frames_per_s": n_cpu / dt, "cores": 1, "kind": "reference",
                                              "sample": "%d ETI frames through the reference's EtiReader + coding graph" % n_cpu}
        except Exception as e:
            res["cpu_reference_coder"] = {"error": str(e)}
    cod.close()
    mod.close()
    return res


def measure_single_tf_latency(dm, torch, host_bits, calls=200):
    """The real-time use of the drop-in: one transmission frame (96 ms of signal) per call through
    dabmod_b200_process, pinned host buffers in and out -- what the ModCodec adapter does once per TF."""
    mod = dm.Modulator(mode=MODE, fir_taps="default", max_batch=1)
    out = torch.empty(mod.tf_out_bytes, dtype=torch.uint8).pin_memory()
    lat = []
    for i in range(calls + 20):
        t0 = time.perf_counter()
        mod.process_batch_ptr(host_bits[i % host_bits.shape[0]].data_ptr(), 1, out.data_ptr(), mod.tf_out_bytes)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[20:]) * 1e6
    mod.close()
    return {"workload": "latency: one TF per call (dabmod_b200_process, host buffers), TM I + FIR default taps",
            "us_median": float(np.median(lat)), "us_p99": float(np.percentile(lat, 99)), "calls": calls,
            "realtime_budget_us": 96000}


def measure_e2e_s16(dm, torch, host_bits, n_tf, steps=5):
    """The headline workload end to end with FormatConverter s16 fused into k_fir: `e2e` is bound by the
    PCIe link (1.57 MB of complexf per TF), so halving the output bytes is what moves it."""
    mod = dm.Modulator(mode=MODE, fir_taps="default", max_batch=n_tf, fmt="s16")
    out_bytes = n_tf * mod.tf_out_bytes
    host_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        mod.process_batch_ptr(host_bits.data_ptr(), n_tf, host_out.data_ptr(), out_bytes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        mod.process_batch_ptr(host_bits.data_ptr(), n_tf, host_out.data_ptr(), out_bytes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mod.close()
    return {"workload": "e2e s16: configs[1] with FormatConverter s16 output, host buffers in and out",
            "eti_frames_per_s": n_tf * ETI_PER_TF * steps / dt, "d2h_bytes_per_step": out_bytes,
            "d2h_GB/s": out_bytes * steps / dt / 1e9, "steps": steps}


# ---------------------------------------------------------------------------
# BASELINE configs[4]: ONE ETI stream sharded across the ranks by transmission-frame range
# ---------------------------------------------------------------------------
STREAM_WORKLOAD = "TM I full chain, 65536-frame synthetic ETI stream sharded across 8xB200, resample to 10 Msps"
STREAM_KW = dict(mode=1, fir_taps="default", output_rate=10000000, normalise=1.0 / 46000.0, poly=POLY)
SEAM_TFS = 2


def measure_sharded_stream(dm, torch, dist, rank, world, local_rank, n_eti_frames, batch_tfs=256):
    """One stream of n_eti_frames ETI(NI) frames -> channel coding -> TM I symbols -> FIR -> 10 Msps -> MemlessPoly,
    cut into contiguous transmission-frame ranges, one per rank (sharding.plan_shards).  Rank r positions coder and
    modulator with dabmod_b200_seek_eti (15 ETI frames of time-interleaver history + the TF before the shard re-run
    for the resampler overlap + TII parity) and streams its range in calls of `batch_tfs` TFs.  No data-path
    collective.  Parity gate: rank r runs SEAM_TFS frames past its range with its continuous state; their CRC must
    equal the CRC of the first SEAM_TFS frames of rank r+1, which started from seek_eti."""
    import importlib
    import zlib
    eti = importlib.import_module("odr_dabmod_b200.eti")
    sharding = importlib.import_module("odr_dabmod_b200.sharding")
    cif = 4
    n_tf = n_eti_frames // cif
    plan = sharding.plan_shards(n_tf, world)
    sh = plan[rank]
    mux = eti.default_multiplex()
    f0 = sh.first_tf * cif
    hist = min(f0, 15 + cif)
    tail = SEAM_TFS * cif if sh.first_tf + sh.n_tf + SEAM_TFS <= n_tf else 0
    n_local = hist + sh.n_tf * cif + tail
    host_eti = torch.empty((max(n_local, 1), 6144), dtype=torch.uint8).pin_memory()
    eti.synth_eti_range(1, mux, f0 - hist, n_local, seed=1234, out=host_eti.numpy())
    frames = host_eti.numpy()
    mode, streams = dm.eti_describe(frames[0])
    B = min(batch_tfs, max(sh.n_tf, SEAM_TFS))
    stream = torch.cuda.Stream()
    res = {"workload": STREAM_WORKLOAD, "eti_frames": n_eti_frames, "tfs": n_tf, "ranks": world,
           "tfs_per_rank": [s.n_tf for s in plan], "tfs_per_call": B, "output": "complexf, 960000 samples per TF",
           "collective": "none on the data path (control-plane all_gather of CRCs and timings only)"}

    def sync_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_formats(fmt):
        kw = dict(STREAM_KW)
        if fmt != "complexf":
            # integer output: no 1/46000 normalisation and no predistortion (MemlessPoly works on |x| <~ 1,
            # FormatConverter on file-scale amplitudes: the reference's "-F s16" file output of the resampled chain)
            kw = dict(mode=1, fir_taps="default", output_rate=10000000, fmt=fmt)
        mod = dm.Modulator(max_batch=B, device=local_rank, **kw)
        cod = dm.Coder(mode, streams, max_frames=B * cif, device=local_rank)
        return mod, cod

    mod, cod = run_formats("complexf")
    out_tf = mod.tf_out_bytes
    d_eti = host_eti.to("cuda")
    d_bits = torch.empty(B * cod.tf_bytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(B * out_tf, dtype=torch.uint8, device="cuda")
    shard_eti = hist * 6144                      # byte offset of the shard's first frame in the local buffers

    def device_pass(collect):
        """seek + the shard's calls, enqueued on `stream`; returns (seek seconds, [crc of the first SEAM_TFS], crc of the tail)"""
        t0 = time.perf_counter()
        cod.seek(mod, sh.first_tf, frames[:hist])
        seek_s = time.perf_counter() - t0
        head = tailcrc = None
        for t in range(0, sh.n_tf, B):
            n = min(B, sh.n_tf - t)
            cod.process_device(d_eti.data_ptr() + shard_eti + t * cif * 6144, n * cif, d_bits.data_ptr(), stream.cuda_stream)
            mod.process_batch_device(d_bits.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
            if collect and t == 0:
                stream.synchronize()
                head = zlib.crc32(d_out[:min(SEAM_TFS, n) * out_tf].cpu().numpy().tobytes())
        if collect and tail:
            t = sh.n_tf
            cod.process_device(d_eti.data_ptr() + shard_eti + t * cif * 6144, tail, d_bits.data_ptr(), stream.cuda_stream)
            mod.process_batch_device(d_bits.data_ptr(), SEAM_TFS, d_out.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            tailcrc = zlib.crc32(d_out[:SEAM_TFS * out_tf].cpu().numpy().tobytes())
        return seek_s, head, tailcrc

    # ---- pass 1 (untimed, also the warm-up): seam CRCs ----
    head = tailcrc = None
    if sh.n_tf:
        _, head, tailcrc = device_pass(True)
    torch.cuda.synchronize()
    crcs = [(head, tailcrc)]
    if dist is not None:
        crcs = [None] * world
        dist.all_gather_object(crcs, (head, tailcrc))
    seams = [(r, crcs[r][1], crcs[r + 1][0]) for r in range(world - 1)
             if crcs[r][1] is not None and crcs[r + 1][0] is not None and plan[r + 1].n_tf >= SEAM_TFS]
    res["seam_parity"] = {"seams_checked": len(seams), "tfs_per_seam": SEAM_TFS,
                          "bit_identical": all(a == b for _, a, b in seams),
                          "how": "CRC32 of the I/Q of the %d TFs after each seam: continued by rank r (stream state) "
                                 "vs started by rank r+1 (seek_eti)" % SEAM_TFS,
                          "crc": [{"seam_after_rank": r, "continued": a, "seeked": b} for r, a, b in seams]}
    if world == 1 and n_tf >= 4 * SEAM_TFS:
        # one rank: the same gate against a second handle that seeks to the middle of the stream
        mid = n_tf // 2
        mod2, cod2 = run_formats("complexf")
        cod2.seek(mod2, mid, frames[hist + mid * cif - (15 + cif): hist + mid * cif])
        a = cod2.modulate(mod2, frames[hist + mid * cif: hist + (mid + SEAM_TFS) * cif])
        cod.seek(mod, mid - SEAM_TFS, frames[hist + (mid - SEAM_TFS) * cif - min((mid - SEAM_TFS) * cif, 15 + cif):
                                             hist + (mid - SEAM_TFS) * cif])
        b = cod.modulate(mod, frames[hist + (mid - SEAM_TFS) * cif: hist + (mid + SEAM_TFS) * cif])[SEAM_TFS:]
        res["seam_parity"].update({"seams_checked": 1, "bit_identical": bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))),
                                   "how": "one rank: %d TFs at mid-stream, a handle that ran into them vs a second handle "
                                          "started there with seek_eti, compared bit for bit" % SEAM_TFS})
        mod2.close()
        cod2.close()
    if not res["seam_parity"]["bit_identical"]:
        raise SystemExit("bench.py: sharded stream: seam parity FAILED: %r" % (res["seam_parity"],))

    # ---- pass 2: device-resident (ETI frames and I/Q in HBM; the I/Q of a call is overwritten by the next) ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record(stream)
    seek_s = 0.0
    if sh.n_tf:
        seek_s, _, _ = device_pass(False)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = sync_max(time.perf_counter() - t_wall)
    dev_ms = sync_max(e0.elapsed_time(e1))         # includes the (synchronous) seek of this rank
    seek_ms = sync_max(seek_s * 1e3)
    res["device_resident"] = {"eti_frames_per_s": n_eti_frames / (dev_ms * 1e-3), "ms": dev_ms, "seek_ms": seek_ms,
                              "wall_ms": wall * 1e3,
                              "timing": "CUDA events on the launching stream around seek_eti + every call of the shard, "
                                        "max over ranks"}
    launches = mod.last_launch_count
    res["kernels_per_call"] = launches + 2
    mod.close()
    cod.close()
    del d_out, d_bits

    # ---- pass 3: host-delivered (ETI frames from pinned host memory, I/Q into pinned host memory) ----
    for fmt in ("complexf", "s16"):
        mod, cod = run_formats(fmt)
        out_tf = mod.tf_out_bytes
        host_out = torch.empty(B * out_tf, dtype=torch.uint8).pin_memory()
        base = host_eti.data_ptr() + shard_eti
        check = 0
        if sh.n_tf:                                  # warm-up call (buffers touched, pipeline events created)
            cod.seek(mod, sh.first_tf, frames[:hist])
            cod.modulate_ptr(mod, base, min(B, sh.n_tf) * cif, host_out.data_ptr(), host_out.numel())
        barrier()
        t0 = time.perf_counter()
        if sh.n_tf:
            cod.seek(mod, sh.first_tf, frames[:hist])
            for t in range(0, sh.n_tf, B):
                n = min(B, sh.n_tf - t)
                cod.modulate_ptr(mod, base + t * cif * 6144, n * cif, host_out.data_ptr(), host_out.numel())
                check ^= int(host_out[::1048573].to(torch.int64).sum().item())    # the delivered bytes are read
        dt = sync_max(time.perf_counter() - t0)
        res["host_delivered_" + fmt] = {"chain": "FIR + 10 Msps + MemlessPoly, complexf_normalised" if fmt == "complexf"
                                        else "FIR + 10 Msps, FormatConverter " + fmt,
                                        "eti_frames_per_s": n_eti_frames / dt, "s": dt,
                                        "d2h_bytes": n_tf * out_tf, "d2h_GB/s": n_tf * out_tf / dt / 1e9,
                                        "h2d_bytes": n_eti_frames * 6144, "checksum": check,
                                        "timing": "wall clock around seek_eti + dabmod_b200_process_eti_batch calls "
                                                  "(pinned host buffers both ways), max over ranks"}
        if fmt == "complexf" and dist is not None:
            # ---- result gather over NCCL ("NCCL only for result gather"), a bounded sample, timed separately ----
            g_tfs = min(64, min(s.n_tf for s in plan))
            if g_tfs > 0:
                dev = torch.device("cuda", local_rank)
                mine = host_out[:g_tfs * out_tf].to(dev).view(g_tfs, out_tf)
                my_crc = zlib.crc32(mine.cpu().numpy().tobytes())
                sub = [sharding.Shard(r, r * g_tfs, g_tfs) for r in range(world)]
                sharding.gather_stream(mine, sub, dist, dst=0, device=dev)        # warm-up (NCCL communicator setup)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                full = sharding.gather_stream(mine, sub, dist, dst=0, device=dev)
                g1.record()
                torch.cuda.synchronize()
                g_ms = sync_max(g0.elapsed_time(g1))
                all_crc = [None] * world
                dist.all_gather_object(all_crc, my_crc)
                ordered = None
                if rank == 0:
                    ordered = all(zlib.crc32(full[r * g_tfs:(r + 1) * g_tfs].cpu().numpy().tobytes()) == all_crc[r]
                                  for r in range(world))
                res["nccl_gather"] = {"tfs_per_rank": g_tfs, "bytes_to_rank0": world * g_tfs * out_tf, "ms": g_ms,
                                      "GB/s": world * g_tfs * out_tf / (g_ms * 1e-3) / 1e9, "in_stream_order": ordered,
                                      "note": "sharding.gather_stream (torch.distributed gather over NCCL/NVLink) of a "
                                              "bounded sample; the full 126 GB stream is drained per rank, not gathered"}
                del mine, full
        mod.close()
        cod.close()
        del host_out
    torch.cuda.empty_cache()
    return res



def measure_pcie_ceiling(torch, dist, world, nbytes=1 << 30, reps=6):
    """The denominator of every host-delivered number: raw cudaMemcpyAsync between pinned host memory and HBM, no
    kernels, all ranks at once (the host's memory system and PCIe root complexes are shared by the GPUs of a box).
    Returns aggregate GB/s in both directions and with both running together."""
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host2 = torch.empty(nbytes // 8, dtype=torch.uint8).pin_memory()
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes // 8, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def d2h():
        with torch.cuda.stream(s1):
            host.copy_(dev, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            dev.copy_(host, non_blocking=True)

    def both():                                  # the shape of the pipeline: a big D2H stream, a small H2D stream
        with torch.cuda.stream(s1):
            host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s2):
            dev2.copy_(host2, non_blocking=True)

    t_d2h, t_h2d, t_both = timed(d2h), timed(h2d), timed(both)
    gb = world * nbytes * reps / 1e9
    return {"ranks": world, "bytes_per_copy": nbytes, "copies": reps,
            "d2h_GB/s": gb / t_d2h, "h2d_GB/s": gb / t_h2d, "d2h_with_small_h2d_GB/s": gb / t_both,
            "how": "cudaMemcpyAsync pinned<->HBM on every rank at once, wall clock, max over ranks, no kernels"}


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank (and allocate its pinned host buffers) on the CPUs next to its GPU: the end-to-end leg
    moves 1.6 GB per step over PCIe, and a remote NUMA node costs a third of that bandwidth."""
    if os.environ.get("DABMOD_BENCH_NO_BIND"):
        return
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def gpu_arm(args):
    import torch
    import dabmod_loader

    dm = dabmod_loader.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.only:
        # development aid: one of the other_configs by prefix, device-resident, no headline line
        stream = torch.cuda.Stream()
        peak, _ = measured_peaks()
        for name, kw, ntf in other_configs():
            if name.startswith(args.only):
                print(json.dumps(measure_config(dm, torch, name, kw, ntf, stream, peak, steps=args.steps)))
        if args.only == "n1":
            print(json.dumps(measure_coder(dm, torch, stream, steps=args.steps)))
        return 0

    n_tf = TFS_PER_STEP
    mod = dm.Modulator(mode=MODE, fir_taps="default", max_batch=n_tf, device=local_rank)
    out_bytes = n_tf * mod.tf_out_bytes
    in_bytes = n_tf * mod.tf_in_bytes

    # synthetic input, distinct per rank; 4 rotating device copies
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_bits = torch.randint(0, 256, (n_tf, mod.tf_in_bytes), dtype=torch.uint8, generator=g).pin_memory()
    d_bits = [host_bits.to("cuda", non_blocking=True).clone() for _ in range(4)]
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    # a dedicated non-default stream: the C ABI takes it as a cudaStream_t and the
    # timing events below are recorded on the same stream the kernels run on
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    torch.cuda.synchronize()

    def step_device(i):
        mod.process_batch_device(d_bits[i % 4].data_ptr(), n_tf, d_out.data_ptr(), stream.cuda_stream)

    def step_e2e():
        return mod.process_batch_ptr(host_bits.data_ptr(), n_tf, host_out.data_ptr(), out_bytes)

    # ---- device-resident: `value` ----
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(i)
    e1.record(stream)
    barrier()
    launches = mod.last_launch_count * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())

    # ---- per-kernel timing for the roofline (same workload, events around each kernel) ----
    mod.set_param("profile", 1)
    ktimes = {}
    reps = max(3, min(args.steps, 10))
    for i in range(reps):
        step_device(i)
        torch.cuda.synchronize()
        for name, t in mod.kernel_times():
            ktimes.setdefault(name, []).append(t)
    mod.set_param("profile", 0)
    kavg = {k: float(np.mean(v)) for k, v in ktimes.items()}

    # ---- the same step as two kernels (fir_kernel = 2: k_symbols_w in its compact layout + k_fir_tma), for comparison:
    #      the fused kernel (fir_kernel = 3, the default) does not write and re-read the 1.245 MB/TF intermediate ----
    two_kernel = None
    if rank == 0:
        try:
            mod.set_param("fir_kernel", 2)
            mod.set_param("profile", 1)
            kt2 = {}
            for i in range(reps + 2):
                step_device(i)
                torch.cuda.synchronize()
                if i >= 2:
                    for name, t in mod.kernel_times():
                        kt2.setdefault(name, []).append(t)
            two_kernel = {k: float(np.mean(v)) for k, v in kt2.items()}
        finally:
            mod.set_param("profile", 0)
            mod.set_param("fir_kernel", 3)

    # ---- end to end through the host-buffer C ABI: `e2e` ----
    for _ in range(max(1, min(args.warmup, 3))):
        step_e2e()
    barrier()
    e2e_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s = float(dt.item())
    clocks = sampler.stop() if rank == 0 else None
    checksum = int(host_out[::65537].to(torch.int64).sum().item())  # the D2H result is read (strided over the whole buffer)

    # ---- the host's PCIe ceiling with every rank copying at once (denominator of the host-delivered numbers) ----
    pcie = None
    try:
        pcie = measure_pcie_ceiling(torch, dist, world)
    except Exception as e:
        pcie = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- BASELINE configs[4]: one ETI stream sharded across the ranks (every rank takes part) ----
    sharded = None
    if args.stream_frames > 0:
        mod.close()
        d_bits = d_out = host_out = None             # 3.2 GB of HBM and 1.6 GB of pinned memory back before the stream leg
        torch.cuda.empty_cache()
        try:
            sharded = measure_sharded_stream(dm, torch, dist, rank, world, local_rank, args.stream_frames)
        except SystemExit:
            raise
        except Exception as e:                      # never lose the headline line
            sharded = {"workload": STREAM_WORKLOAD, "error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    eti_per_step = world * n_tf * ETI_PER_TF
    value = eti_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = eti_per_step * e2e_steps / e2e_s

    peak, peak_src = measured_peaks()
    bytes_per_launch = {"k_symbols": SYM_BYTES_PER_TF * n_tf, "k_symbols_w": SYM_BYTES_PER_TF * n_tf,
                        "k_symbols_w_fir": SYM_BYTES_PER_TF * n_tf,      # bits in, filtered I/Q out: nothing in between
                        "k_fir": FIR_BYTES_PER_TF * n_tf, "k_fir_sym": FIR_SYM_BYTES_PER_TF * n_tf,
                        "k_fir_tma": FIR_SYM_BYTES_PER_TF * n_tf}
    if "k_fir_sym" in kavg or "k_fir_tma" in kavg:
        # compact intermediate: the symbol kernel writes 76 x 2048 samples per TF (no null symbol, no cyclic prefix)
        bytes_per_launch["k_symbols_w"] = (TF_IN_BYTES + COMPACT_SAMPLES * 8) * n_tf
    dom = max(kavg, key=lambda k: kavg[k])
    achieved = bytes_per_launch[dom] / (kavg[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": bytes_per_launch[dom], "kernel_ms": kavg[dom],
        "all_kernels": {k: {"ms": kavg[k], "GB/s": bytes_per_launch[k] / (kavg[k] * 1e-3) / 1e9,
                            "frac": bytes_per_launch[k] / (kavg[k] * 1e-3) / 1e9 / peak} for k in kavg},
    }
    if dom == "k_symbols_w_fir":
        # The fused kernel is bound by the FP32 pipe, not by HBM: per TF 35.4 MFLOP of FIR + 8.6 MFLOP of IFFT +
        # 0.95 MFLOP of gain statistics against 1.6 MB of compulsory traffic = 28 flop/B, the ridge of
        # this GPU being 72 TFLOP/s / 6.55 TB/s = 11 flop/B.  Both roofs are reported; `frac` above is the HBM one.
        flops = (35.4e6 + 8.6e6 + 0.95e6) * n_tf
        tf = flops / (kavg[dom] * 1e-3) / 1e12
        roofline["fp32"] = {"achieved": tf, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP32_PEAK_TFLOPS,
                            "peak_source": "measured FFMA rate, profiles/r01_ubench_fp32_pipes.txt",
                            "algorithmic_flops_per_launch": flops}
        roofline["binding_roof"] = "fp32"
    if two_kernel:
        b2 = {"k_symbols_w": (TF_IN_BYTES + COMPACT_SAMPLES * 8) * n_tf, "k_fir_tma": FIR_SYM_BYTES_PER_TF * n_tf,
              "k_fir_sym": FIR_SYM_BYTES_PER_TF * n_tf}
        roofline["two_kernel_step"] = {
            "how": "the same step with fir_kernel = 2 (symbol kernel -> compact intermediate in HBM -> k_fir_tma)",
            "ms": sum(two_kernel.values()),
            "kernels": {k: {"ms": v, "GB/s": b2.get(k, 0) / (v * 1e-3) / 1e9, "frac": b2.get(k, 0) / (v * 1e-3) / 1e9 / peak}
                        for k, v in two_kernel.items()},
            "bytes_moved_per_step": sum(b2.get(k, 0) for k in two_kernel),
        }
    if not args.no_extras:
        # ("k_symbols_w_fir" is the timing label of the FUSE instance of the function k_symbols_w)
        roofline["traffic"], roofline["traffic_source"] = measure_traffic("k_symbols_w" if dom == "k_symbols_w_fir" else dom)
    if roofline["traffic"] is None:
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as f:
                roofline["traffic"] = json.load(f).get(dom)
            roofline["traffic_source"] = "profiles/traffic.json (ncu capture of an earlier run: %s)" % roofline.get("traffic_source")

    os.sched_setaffinity(0, all_cpus)      # the CPU legs below use every host core again
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        tfs = 192
        runs = sorted(run_cpu(cores, tfs) for _ in range(3))       # median of three ~1 s samples (48 threads on the
        v, dt_cpu, kind = runs[1]                                   # host's cores: single samples scatter by 10-20 %)
        cpu = {"value": v, "unit": "ETI frames/s", "cores": cores, "kind": kind,
               "sample": "median of 3 runs of %d TFs per thread x %d threads, %.1f s each (same TM I + FIR default "
                         "taps chain); all three: %s" % (tfs, cores, dt_cpu, ", ".join("%.0f" % r[0] for r in runs))}

    others = None
    if world == 1 and not args.no_extras:
        others = []
        for name, kw, ntf in other_configs():
            try:
                others.append(measure_config(dm, torch, name, kw, ntf, stream, peak, with_cpu=not args.no_cpu))
            except Exception as e:                       # an extra must never cost the headline line
                others.append({"workload": name, "error": str(e)})
        try:
            others.append(measure_single_tf_latency(dm, torch, host_bits))
        except Exception as e:
            others.append({"workload": "latency", "error": str(e)})
        try:
            others.append(measure_e2e_s16(dm, torch, host_bits, n_tf))
        except Exception as e:
            others.append({"workload": "e2e s16", "error": str(e)})
        try:
            others.append(measure_coder(dm, torch, stream, with_cpu=not args.no_cpu))
        except Exception as e:
            others.append({"workload": "n1 coder", "error": str(e)})
        if not args.no_cpu:
            try:
                others.append(measure_binary())
            except Exception as e:
                others.append({"workload": "c0 binary", "error": str(e)})

    line = {
        "metric": METRIC, "value": value, "unit": "ETI frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "mode": MODE, "tfs_per_step_per_gpu": n_tf,
                   "eti_frames_per_step": eti_per_step, "output": "complexf", "gain": "var",
                   "l2": "inputs larger than L2: per step 1.6 GB of output written and 29 MB of input read (one fused kernel, "
                         "no intermediate) >> 126 MB L2; input bits rotate over 4 buffers",
                   "parallelism": "frame-sharded x%d, no collective" % world},
        "e2e": {"value": e2e_value, "unit": "ETI frames/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "checksum": checksum,
                "d2h_GB/s": world * out_bytes * e2e_steps / e2e_s / 1e9,
                "pcie_ceiling": pcie,
                "frac_of_d2h_ceiling": (world * out_bytes * e2e_steps / e2e_s / 1e9 / pcie["d2h_GB/s"]
                                        if pcie and "d2h_GB/s" in pcie else None),
                "bound": "PCIe: 1.57 MB of complexf per TF through one x16 link; with several GPUs the host side "
                         "(memory system, root complexes) all ranks share"},
        "gpu_launches": launches,
        "roofline": roofline,
        "clocks": clocks,
    }
    if sharded is not None:
        if pcie and "d2h_GB/s" in pcie:
            for k in ("host_delivered_complexf", "host_delivered_s16"):
                if k in sharded:
                    sharded[k]["frac_of_d2h_ceiling"] = sharded[k]["d2h_GB/s"] / pcie["d2h_GB/s"]
        line["sharded_stream"] = sharded
        line["config"]["sharded_stream_workload"] = STREAM_WORKLOAD
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if others is not None:
        line["other_configs"] = others
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only", default="", help="development: measure only the other_configs entry with this prefix")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (other_configs key)")
    ap.add_argument("--stream-frames", type=int, default=65536,
                    help="ETI frames of the sharded-stream leg (BASELINE configs[4]); 0 skips it")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.traffic_child:
        return traffic_child()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
