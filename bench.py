#!/usr/bin/env python
"""Benchmark of the enhancement forward path (BASELINE.json metric: R-CED V2 audio-seconds
enhanced per second; % of FP32 FFMA peak for the fused network kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1); utterances are independent, so every rank enhances
its own batch with no data-path collective (weak scaling) and the only communication is the
timing barrier / max-over-ranks.  A "step" is one pass of STFT -> fused network ->
reconstruction over one batch of synthetic utterances.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET_WORK = "FullyCNNV2"
N_UTT = 1024                 # BASELINE.json configs[1]
UTT_SAMPLES = 32000          # 4 s @ 8 kHz
E2E_CHUNKS = [int(c) for c in os.environ.get("RCED_E2E_CHUNKS", "32,96,128,128,128,128,128,128,96,32").split(",")]   # utterances per chunk of the host pipeline
SAMPLE_RATE = 8000
POOL = 64                    # distinct synthetic utterances, tiled to N_UTT
METRIC = "R-CED V2 audio-seconds enhanced per second"
UNIT = "audio-s/s"
DEFAULT_VARIANT = "tc"       # network kernel timed by default: tcgen05 tensor cores, FP16 x3 split


def synth_pool():
    from fullycnnspeechenhancement_b200.synth import noisy_utterance
    return [noisy_utterance(1000 + i, UTT_SAMPLES) for i in range(POOL)]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_path(pool, weights, n_utt, batch, timings=None):
    """The oracle port of the reference path on `n_utt` utterances; returns seconds."""
    from oracle.cpu_path import enhance_batch_cpu
    t0 = time.perf_counter()
    for s in range(0, n_utt, batch):
        waves = [pool[i % len(pool)] for i in range(s, min(n_utt, s + batch))]
        enhance_batch_cpu(waves, NET_WORK, weights, timings=timings)
    return time.perf_counter() - t0


def cpu_baseline(pool, weights, budget_s=15.0, batch=32):
    import torch
    cores = host_cores()
    torch.set_num_threads(cores)
    run_cpu_path(pool, weights, 4, 4)                       # warm-up (thread pools, FFT plans)
    t_probe = run_cpu_path(pool, weights, batch, batch)
    n = int(max(batch, min(1024, budget_s / max(t_probe, 1e-3) * batch)) // batch * batch)
    timings = {}
    t = run_cpu_path(pool, weights, n, batch, timings)
    return {
        "value": n * UTT_SAMPLES / SAMPLE_RATE / t, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d x 4 s utterances, batch %d, oracle port of the reference path (numpy STFT/rebuild with the "
                  "reference's per-sample de-emphasis loop + torch-CPU float32 conv2d on %d threads); %s; "
                  "stage seconds stft=%.2f network=%.2f rebuild=%.2f"
                  % (n, batch, cores, cpu_model(), timings["stft"], timings["network"], timings["rebuild"]),
    }


class ClockSampler(object):
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: TensorFlow
    1.14 cannot be installed) on the box's host cores, rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import network
    cores = host_cores()
    torch.set_num_threads(cores)
    pool = synth_pool()[:32]
    weights = network.random_weights(NET_WORK, seed=0, randomize_bn=False)
    per_step, batch = 32, 32
    for _ in range(args.warmup):
        run_cpu_path(pool, weights, per_step, batch)
    timings = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu_path(pool, weights, per_step, batch, timings)
    t = time.perf_counter() - t0
    value = args.steps * per_step * UTT_SAMPLES / SAMPLE_RATE / t
    sample = ("each step = %d x 4 s utterances (bounded sample of the 1024-utterance batch), batch %d, %d threads, %s; "
              "stage seconds stft=%.2f network=%.2f rebuild=%.2f" % (per_step, batch, cores, cpu_model(),
                                                                       timings["stft"], timings["network"], timings["rebuild"]))
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "R-CED V2 (FullyCNNV2) enhancement of 4 s 8 kHz synthetic noisy utterances, CPU",
                   "utterances_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written) or the fallback B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(path))
        return {"hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "source": "measured (MEASURED_PEAKS.json)"}
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


def roofline(variant, achieved, ffma_peak, flops_valid, k2_ms, share, traffic, rows, tc_status):
    """Roofline object of the dominant kernel (the fused network).  `achieved` is algorithmic:
    valid-tap MACs x 2 per launch / mean launch time measured with CUDA events in the timed steps."""
    common = {"achieved": achieved, "unit": "TFLOP/s", "flops_per_launch": flops_valid,
              "flop_basis": "valid-tap MACs x2 (3,959,092 MAC/frame), FP32-equivalent", "kernel_ms": k2_ms,
              "kernel_share_of_step": share}
    if variant != "tc":
        common.update({"bound": "fp32_ffma", "kernel": "rced_net_kernel<2,TMEM> (fused 16-layer network)", "peak": ffma_peak,
                       "frac": achieved / ffma_peak, "traffic": traffic,
                       "peak_source": "measured in this run by rced_ffma_peak (independent FFMA chains, 64 warps/SM); "
                                      "MEASURED_PEAKS.json holds no FP32 figure; nominal 2*128*148*1.965 GHz = 74.4"})
        return common
    pk = measured_peaks()
    # tensor-core work actually issued: per 128-row tile and unit (two (tap, 8-channel) chunks,
    # K = 16) A_hi x [Whi|Wlo] (N = 2 NP) and A_lo x Whi (N = NP); 8 tiles per 7-frame batch; output layer 5 row-shifted blocks
    import ctypes as ct
    from fullycnnspeechenhancement_b200 import _lib
    out = (ct.c_int64 * 4096)()
    _lib.check(_lib.lib().rced_tc_layout(2, out, 4096))
    ns = out[0]
    mac_tile, cyc_tile = 0, 0.0
    for st in range(ns):
        units, npad, final = out[16 + 6 * st], out[18 + 6 * st], out[21 + 6 * st]
        if final:   # row-shifted blocks x 3 products, for the even- and the odd-frame copy (14 of 16 tile-parity pairs)
            mac_tile += 1.75 * units * 3 * 128 * npad * 16
            cyc_tile += 1.75 * units * 3 * (32 + npad / 4.0)
        else:
            mac_tile += units * 128 * 16 * 3 * npad
            cyc_tile += units * ((32 + 2 * npad / 4.0) + (32 + npad / 4.0))
    batches = (rows + 6) // 7
    issued = 2.0 * mac_tile * 8 * batches
    common.update({
        "bound": "tensor", "kernel": "rced_net_tc_kernel<2> (fused 16-layer network, tcgen05 kind::f16)",
        "peak": pk["bf16_tflops"], "frac": achieved / pk["bf16_tflops"], "peak_source": pk["source"] + ", dense bf16/fp16",
        "traffic": traffic,
        "traffic_note": "dram__bytes_read + write of one launch (ncu --set full, profiles/k2tc_dram_traffic.json): 0.26 GB algorithmic, the "
                        "rest is write-back of the L2-resident skip scratch",
        "l1tex_throughput_pct": 91.0,
        "l1tex_note": "ncu (profiles/r01_k2_tc_v7_ncu_digest.txt): l1tex__throughput 91 % of peak -- the shared-memory operand fetch "
                      "of the small-N MMAs is the binding unit; sm__pipe_tensor_cycles_active 40 %",
        "issued_tflops": issued / (k2_ms * 1e-3) / 1e12,
        "issued_note": "tensor-core FLOP actually issued: 3 FP16 products per multiply, channels padded to 8 / 16 / 32, "
                       "136-row frame stride, 7 frames per 8 row tiles",
        "fp32_ffma_peak": ffma_peak, "achieved_vs_fp32_ffma_peak": achieved / ffma_peak,
        "operand_fetch_model": {
            "note": "a small-N tcgen05.mma is bound by its shared-memory operand fetch (32 + N/4 cycles at M = 128, "
                    "profiles/r01_umma_probe_rates.log), not by the tensor pipe: the kernel's own ceiling is the sum of "
                    "those cycles",
            "mma_cycles_per_batch": cyc_tile * 8,
            "pipe_busy_frac": cyc_tile * 8 * batches / 148.0 / (k2_ms * 1e-3 * 1.965e9)},
        "guard": tc_status,
    })
    return common


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print on fd 1 (NCCL's version banner, torchrun notices) goes to stderr;
    the one JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-in-global", action="store_true", help="park skips in global scratch instead of TMEM")
    ap.add_argument("--variant", default=DEFAULT_VARIANT, choices=["ffma", "tc"],
                    help="network kernel: FP32 FFMA, or tcgen05 tensor cores with the FP16 x3 split")
    ap.add_argument("--sweep-utterances", type=int, default=0,
                    help="also time one job of this many 4 s utterances partitioned over the ranks (BASELINE.json configs[4]: "
                         "100000), device-resident, and report it under 'sweep'")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from fullycnnspeechenhancement_b200 import _lib
    from fullycnnspeechenhancement_b200.engine import Enhancer, num_frames
    from fullycnnspeechenhancement_b200.model_utils import fold

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib = _lib.lib()

    # random-init weights of the reference architecture (no checkpoint ships with the reference)
    weights = fold.glorot_weights(NET_WORK, seed=0)
    eng = Enhancer(NET_WORK, weights, device=local_rank, variant=args.variant)
    if args.skip_in_global:
        eng.set_skip_in_tmem(False)
    eng.set_variant(args.variant)

    pool = synth_pool()
    lengths = np.full(N_UTT, UTT_SAMPLES, dtype=np.int64)
    total = int(lengths.sum())
    h_wav = torch.empty(total, dtype=torch.float32).pin_memory()
    hv = h_wav.numpy()
    for i in range(N_UTT):
        hv[i * UTT_SAMPLES:(i + 1) * UTT_SAMPLES] = pool[(i + rank) % POOL]
    h_out = torch.empty(total, dtype=torch.float32).pin_memory()
    d_wav = h_wav.to(dev)
    d_out = torch.empty_like(d_wav)

    T = int(num_frames(UTT_SAMPLES))
    rows = N_UTT * T
    plan = eng.plan(lengths)                       # one chunk: the whole batch per launch
    # pipelined over streams for the host path; a small first and last chunk shorten the first upload and the last
    # download, which nothing overlaps
    plan_e2e = eng.plan(lengths, chunk_utts=E2E_CHUNKS)
    row_off = plan["row_off_all"]
    mag = torch.empty((rows, 129), dtype=torch.float32, device=dev)
    phase = torch.empty((rows, 129, 2), dtype=torch.float32, device=dev)
    pred = torch.empty((rows, 129), dtype=torch.float32, device=dev)

    def step_device(ev=None):
        eng.stft_device(d_wav, plan["wav_off"], plan["wav_len"], row_off, rows, mag, phase)
        if ev:
            ev[0].record()
        eng.forward_device(mag, row_off, pred)
        if ev:
            ev[1].record()
        eng.istft_device(pred, phase, row_off, T, d_out, plan["wav_off"], plan["wav_len"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # measured FP32 FFMA peak of this GPU (the roofline denominator; MEASURED_PEAKS.json has none)
    tf = ctypes.c_double()
    _lib.check(lib.rced_ffma_peak(local_rank, 4096, ctypes.byref(tf)))
    ffma_peak = tf.value

    # ---------------- device-resident throughput (`value`) ------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.rced_launch_count()
    k2_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_device(k2_events[i])
    e1.record()
    barrier()
    launches = lib.rced_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_events]))
    audio_s_per_step = N_UTT * UTT_SAMPLES / SAMPLE_RATE
    value = world * audio_s_per_step * args.steps / (ms_total * 1e-3)

    # ---------------- end to end through the host API (`e2e`) ---------------------------------
    for _ in range(max(1, min(args.warmup, 3))):
        eng.run_plan_host(plan_e2e, h_wav, h_out, d_wav, d_out)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    for _ in range(args.steps):
        eng.run_plan_host(plan_e2e, h_wav, h_out, d_wav, d_out)
    f1.record()
    barrier()
    e2e_ms = max_over_ranks(f0.elapsed_time(f1))
    e2e_value = world * audio_s_per_step * args.steps / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    sweep = None
    if args.sweep_utterances > 0:   # every rank takes part (before the non-zero ranks leave)
        # BASELINE.json configs[4]: a fixed job partitioned across the ranks (strong scaling).  The job is cut into
        # batches of N_UTT utterances (the device-resident pool tiled; H2D is not in the timed region), batch i goes to
        # rank i % world, no collective.
        n_batches = (args.sweep_utterances + N_UTT - 1) // N_UTT
        mine = len(range(rank, n_batches, world))
        step_device()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(mine):
            step_device()
        s1.record()
        barrier()
        job_ms = max_over_ranks(s0.elapsed_time(s1))
        job_utts = n_batches * N_UTT
        sweep = {"workload": "R-CED V2, one job of %d synthetic 4 s utterances (%d batches of %d) partitioned over %d "
                                       "B200 (BASELINE.json configs[4]); inputs resident in HBM (pool tiled)" %
                                       (job_utts, n_batches, N_UTT, world),
                           "utterances": job_utts, "audio_seconds": job_utts * UTT_SAMPLES / SAMPLE_RATE,
                           "job_seconds": job_ms * 1e-3, "scaling": "strong",
                           "value": job_utts * UTT_SAMPLES / SAMPLE_RATE / (job_ms * 1e-3), "unit": UNIT}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_valid = 2.0 * lib.rced_mac_per_frame(eng.arch, 1) * rows
    achieved = flops_valid / (k2_ms * 1e-3) / 1e12
    tc_status = None
    if args.variant == "tc":
        amax, perr = eng.tc_status()
        tc_status = {"max_abs_activation": amax, "protocol_error": perr, "ffma_fallback_ran": bool(amax > 65504.0 or perr != 0)}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k2tc_dram_traffic.json" if args.variant == "tc" else "k2_dram_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except (ValueError, OSError):
            traffic = None
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16x3 (FP16 hi/lo split, 3 products per multiply, FP32 accumulate; 1e-6 of float64)" if args.variant == "tc" else "f32",
        "data": "synthetic (64 seeded tone/chirp + white/babble-noise utterances tiled to 1024 per GPU)",
        "config": {"workload": "R-CED V2 (FullyCNNV2, 16 layers, random-init Glorot weights) batched enhancement of "
                               "1024 synthetic 4 s 8 kHz utterances per B200 (BASELINE.json configs[1])",
                   "utterances_per_gpu": N_UTT, "samples_per_utterance": UTT_SAMPLES, "frames_per_gpu": rows,
                   "partition": "independent utterances per rank, no collective",
                   "l2": "per-step working set (131 MB wav in, 131 MB mag, 263 MB phase, 131 MB pred, 131 MB wav out) "
                         "exceeds the 126 MB L2, no explicit flush",
                   "network_kernel": "tcgen05 tensor cores, FP16 x3 error-compensated split (rced_net_tc_kernel); FP32 FFMA "
                                     "kernel queued behind it as range-guard fall-back" if args.variant == "tc"
                                     else "FP32 FFMA (rced_net_kernel)",
                   "skip_storage": ("global scratch (L2)" if args.variant == "tc" else
                                    "global" if args.skip_in_global else "tmem")},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": total * 4, "d2h_bytes_per_step": total * 4,
                "api": "Enhancer.run_plan_host: pinned host waveforms -> H2D -> rced_enhance (K1,K2,K3) -> D2H, "
                       "chunks of %s utterances over 3 streams" % "/".join(str(c) for c in E2E_CHUNKS)},
        "gpu_launches": int(launches),
        "roofline": roofline(args.variant, achieved, ffma_peak, flops_valid, k2_ms, k2_ms * args.steps / ms_total, traffic,
                             rows, tc_status),
        "clocks": clocks,
    }
    if sweep is not None:
        result["sweep"] = sweep
    if not args.no_cpu_baseline and world == 1:
        from oracle import network as onet
        result["cpu_baseline"] = cpu_baseline(pool, onet.random_weights(NET_WORK, seed=0, randomize_bn=False))
    emit(result)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
