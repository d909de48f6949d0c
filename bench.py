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


REF_UTTS_PER_STEP = 64     # bounded sample of the 1024-utterance batch the CPU arm enhances per step


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port: TensorFlow
    1.14 cannot be installed) on the box's host cores, rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle import network
    cores = host_cores()
    torch.set_num_threads(cores)
    pool = synth_pool()
    weights = network.random_weights(NET_WORK, seed=0, randomize_bn=False)
    per_step, batch = REF_UTTS_PER_STEP, 32
    for _ in range(args.warmup):
        run_cpu_path(pool, weights, per_step, batch)
    timings = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu_path(pool, weights, per_step, batch, timings)
    t = time.perf_counter() - t0
    value = args.steps * per_step * UTT_SAMPLES / SAMPLE_RATE / t
    sample = ("each step = %d x 4 s utterances in batches of %d -- a bounded sample of the 1024-utterance batch of the GPU arm "
              "(throughput-normalised: audio-seconds per second), %d threads, %s; stage seconds stft=%.2f network=%.2f "
              "rebuild=%.2f" % (per_step, batch, cores, cpu_model(), timings["stft"], timings["network"], timings["rebuild"]))
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "R-CED V2 (FullyCNNV2) enhancement of 4 s 8 kHz synthetic noisy utterances, CPU",
                   "utterances_per_step": per_step,
                   "note": "the GPU arm enhances 1024 utterances per step; the CPU arm a bounded sample of %d of the same "
                           "utterances per step, both reported as audio-seconds enhanced per second" % per_step},
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


def source_sha(*names):
    """sha256 (16 hex digits) of the kernel sources a profile refers to: tells whether a committed ncu figure
    still describes the kernel that was just timed."""
    import hashlib
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "fullycnnspeechenhancement_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_figures(name, sources):
    """profiles/<name>: what one `ncu --set full` capture of the kernel measured (tools/ncu_digest.py --json writes it),
    or None.  `kernel_source_matches` says whether the capture was taken from the sources timed now."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        return None
    d["file"] = "profiles/" + name
    d["kernel_source_matches"] = d.get("kernel_source_sha16") == source_sha(*sources)
    return d


def tc_issue_model(rows):
    """Tensor-core work the K2-TC kernel issues from the library's own step table: per 128-row tile
    and unit (two (tap, 8-channel) chunks, K = 16) A_hi x [Whi|Wlo'] (N = 2 NP) and A_lo' x Whi (N = NP rounded up to 16); 8 tiles per
    7-frame batch; output layer: 5 row-shifted blocks x 3 products for 14 of the 16 (tile, parity) pairs."""
    import ctypes as ct
    from fullycnnspeechenhancement_b200 import _lib
    out = (ct.c_int64 * 4096)()
    _lib.check(_lib.lib().rced_tc_layout(2, out, 4096))
    ns = out[0]
    mac_tile, cyc_tile = 0, 0.0
    for st in range(ns):
        units, npad, final = out[16 + 6 * st], out[18 + 6 * st], out[21 + 6 * st]
        if final:
            mac_tile += 1.75 * units * 3 * 128 * npad * 16
            cyc_tile += 1.75 * units * 3 * (32 + npad / 4.0)
        else:
            n2 = (npad + 15) // 16 * 16          # N of the second instruction (A_lo' x Whi)
            mac_tile += units * 128 * 16 * (2 * npad + n2)
            cyc_tile += units * ((32 + 2 * npad / 4.0) + (32 + n2 / 4.0))
    batches = (rows + 6) // 7
    return 2.0 * mac_tile * 8 * batches, cyc_tile * 8, batches


def roofline_tc(achieved, ffma_peak, flops_valid, k2_ms, share, rows, tc_status, sm_mhz):
    """Roofline object of the dominant kernel (the fused network, tensor-core variant).  `achieved` is algorithmic:
    valid-tap MACs x 2 per launch / mean launch time measured with CUDA events in the timed steps."""
    pk = measured_peaks()
    issued, cyc_batch, batches = tc_issue_model(rows)
    ncu = ncu_figures("r02_k2tc_ncu.json", ["rced_net_tc.cu", "rced_tc.cuh"])
    clk = (sm_mhz or 1965.0) * 1e6
    r = {
        "bound": "tensor", "kernel": "rced_net_tc_kernel<2> (fused 16-layer network, tcgen05 kind::f16)",
        "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops"],
        "peak_source": pk["source"] + ", dense bf16/fp16 (burst: the kernel is timed by its own events)",
        "flops_per_launch": flops_valid, "flop_basis": "valid-tap MACs x2 (3,959,092 MAC/frame), FP32-equivalent",
        "kernel_ms": k2_ms, "kernel_share_of_step": share,
        "traffic": ncu.get("dram_bytes_per_launch") if ncu else None,
        "ncu": ncu,
        "issued_tflops": issued / (k2_ms * 1e-3) / 1e12,
        "issued_per_useful_flop": issued / flops_valid,
        "issued_note": "tensor-core FLOP actually issued: 3 FP16 products per multiply, channels padded to 8 / 16 / 32, "
                       "136-row frame stride, 7 frames per 8 row tiles",
        "fp32_ffma_peak": ffma_peak, "achieved_vs_fp32_ffma_peak": achieved / ffma_peak,
        "operand_fetch_model": {
            "note": "a small-N tcgen05.mma is bound by its shared-memory operand fetch (32 + N/4 cycles at M = 128, "
                    "profiles/r01_umma_probe_rates.log), not by the tensor pipe: the kernel's own ceiling is the sum of "
                    "those cycles",
            "mma_cycles_per_batch": cyc_batch,
            "pipe_busy_frac": cyc_batch * batches / 148.0 / (k2_ms * 1e-3 * clk)},
        "guard": tc_status,
    }
    return r


def roofline_ffma(achieved, ffma_peak, flops_valid, k2_ms):
    ncu = ncu_figures("r02_k2ffma_ncu.json", ["rced_net.cu", "rced_arch.cuh"])
    return {"bound": "fp32_ffma", "kernel": "rced_net_kernel<2,TMEM> (fused 16-layer network, FP32 FFMA2)", "achieved": achieved,
            "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved / ffma_peak, "kernel_ms": k2_ms,
            "flops_per_launch": flops_valid, "flop_basis": "valid-tap MACs x2 (3,959,092 MAC/frame)",
            "traffic": ncu.get("dram_bytes_per_launch") if ncu else None, "ncu": ncu,
            "peak_source": "measured in this run by rced_ffma_peak (independent FFMA chains, 64 warps/SM); "
                           "MEASURED_PEAKS.json holds no FP32 figure; nominal 2*128*148*1.965 GHz = 74.4"}


def roofline_hbm(kernel, ms, rows, bytes_per_row, what, ncu_name, sources):
    pk = measured_peaks()
    algo = float(rows) * bytes_per_row
    gbs = algo / (ms * 1e-3) / 1e9
    ncu = ncu_figures(ncu_name, sources)
    return {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
            "kernel_ms": ms, "bytes_per_launch": algo, "byte_basis": what, "peak_source": pk["source"],
            "traffic": ncu.get("dram_bytes_per_launch") if ncu else None, "ncu": ncu}


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


def latency_block(eng, pool, torch):
    """Small-batch latencies through the host API (BASELINE.json configs[0] shape): one 4 s
    utterance per call, and the online enhancer's push for 512- and 4096-sample blocks (look-back 13 hops + block +
    look-ahead 6 hops per call).  Host wall clock around the synchronous call, median of the repetitions."""
    from fullycnnspeechenhancement_b200.streaming import LOOK_AHEAD, LOOK_BACK, StreamingEnhancer
    out = {}
    wv = pool[0]
    for _ in range(5):
        eng.enhance([wv])
    ts = []
    for _ in range(30):
        t0 = time.perf_counter()
        eng.enhance([wv])
        ts.append(time.perf_counter() - t0)
    one = float(np.median(ts))
    out["one_utterance_4s"] = {"ms": 1e3 * one, "real_time_factor": one / (UTT_SAMPLES / SAMPLE_RATE),
                               "what": "Enhancer.enhance([4 s waveform]): numpy in, numpy out, one rced_enhance_host call"}
    long_wv = np.concatenate([pool[i] for i in range(4)])
    for block in (512, 4096):
        st = StreamingEnhancer(eng, block=block)
        ts = []
        pos = 0
        while pos + block <= len(long_wv) and len(ts) < 60:
            t0 = time.perf_counter()
            st.push(long_wv[pos:pos + block])
            ts.append(time.perf_counter() - t0)
            pos += block
        st.flush()
        push = float(np.median(ts[5:]))
        out["stream_block_%d" % block] = {
            "push_ms": 1e3 * push, "real_time_factor": push / (block / SAMPLE_RATE),
            "algorithmic_latency_ms": 1e3 * (LOOK_AHEAD + block) / SAMPLE_RATE,
            "samples_per_call": LOOK_BACK + block + LOOK_AHEAD,
            "what": "StreamingEnhancer.push of one block (host wall clock, median); final samples lag the input by the "
                    "look-ahead plus the block"}
    return out


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default=DEFAULT_VARIANT, choices=["ffma", "tc"],
                    help="network kernel behind `value` / `e2e`: tcgen05 tensor cores with the FP16 x3 split (default), or FP32 "
                         "FFMA.  The other kernel is always timed for a few steps as well (`fp32_ffma` / `tensor_core`)")
    ap.add_argument("--sweep-utterances", type=int, default=100000,
                    help="one job of this many 4 s utterances partitioned over the ranks (BASELINE.json configs[4]), device-"
                         "resident, reported under 'sweep' (0: skip)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--relay", default="auto", choices=["auto", "off"],
                    help="multi-GPU: measure every GPU's host link with all ranks copying at once and let ranks on a slow / shared "
                         "link move their waveforms through a fast peer GPU (rced_host_set_relay); off: every rank copies directly")
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
    eng.set_variant(args.variant)
    other = "ffma" if args.variant == "tc" else "tc"

    pool = synth_pool()
    lengths = np.full(N_UTT, UTT_SAMPLES, dtype=np.int64)
    total = int(lengths.sum())
    # two sets of page-locked host buffers: step i + 1 is queued while step i still runs (a consumer would read the
    # other set meanwhile)
    # (library-allocated page-locked memory; RCED_BENCH_WC=1: write-combined input buffers, an experiment)
    from fullycnnspeechenhancement_b200.engine import PinnedArray
    wc = os.environ.get("RCED_BENCH_WC") == "1"
    pins = [PinnedArray(total, write_combined=wc) for _ in range(2)] + [PinnedArray(total) for _ in range(2)]
    h_wav = [pins[0].array, pins[1].array]
    h_out = [pins[2].array, pins[3].array]
    staged = np.empty(total, np.float32)
    for i in range(N_UTT):
        staged[i * UTT_SAMPLES:(i + 1) * UTT_SAMPLES] = pool[(i + rank) % POOL]
    h_wav[0][:] = staged
    h_wav[1][:] = staged
    d_wav = torch.from_numpy(staged).to(dev)
    d_out = torch.empty_like(d_wav)

    T = int(num_frames(UTT_SAMPLES))
    rows = N_UTT * T
    plan = eng.plan(lengths)                       # one chunk: the whole batch per launch
    tables = eng.host_tables(lengths)
    assert tables["total"] == total
    row_off = plan["row_off_all"]
    mag = torch.empty((rows, 129), dtype=torch.float32, device=dev)
    phase = torch.empty((rows, 129, 2), dtype=torch.float32, device=dev)
    pred = torch.empty((rows, 129), dtype=torch.float32, device=dev)

    def step_device(ev=None):
        if ev:
            ev[0].record()
        eng.stft_device(d_wav, plan["wav_off"], plan["wav_len"], row_off, rows, mag, phase)
        if ev:
            ev[1].record()
        eng.forward_device(mag, row_off, pred)
        if ev:
            ev[2].record()
        eng.istft_device(pred, phase, row_off, T, d_out, plan["wav_off"], plan["wav_len"])
        if ev:
            ev[3].record()

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

    def per_rank(x):
        """[x of rank 0, x of rank 1, ...] on every rank (diagnostics of the host side: which ranks are slow)."""
        if world == 1:
            return [float(x)]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def gpu_numa_node():
        try:
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            devid = torch.cuda.get_device_properties(local_rank).pci_device_id
            with open("/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devid)) as f:
                return int(f.read().strip())
        except Exception:
            return -99

    def timed_device_steps(n_steps, n_warm):
        """n_steps passes of K1 -> K2 -> K3 with CUDA events around every kernel; returns (ms total, [K1, K2, K3] mean ms)."""
        for _ in range(n_warm):
            step_device()
        barrier()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_steps)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for i in range(n_steps):
            step_device(evs[i])
        b.record()
        barrier()
        mine = a.elapsed_time(b)
        ms = max_over_ranks(mine)
        per = [float(np.mean([e[j].elapsed_time(e[j + 1]) for e in evs])) for j in range(3)]
        per.append(per_rank(mine / n_steps))
        return ms, per

    # ---------------- host links (multi-GPU): who copies through whom ---------------------------
    from fullycnnspeechenhancement_b200.engine import plan_relays, relay_candidates
    link, link_fast_only, relay_of = None, None, None
    if world > 1:
        # Every rank copies its own page-locked buffers both ways at the same time (the buffers exist already: the timed
        # copies start right behind the barrier on all ranks and run for ~100 ms).  If that shows ranks whose link is too slow
        # for their stream, a second pass with only the other ranks copying shows what THEIR links give once the slow ranks'
        # traffic is gone -- i.e. whether they can carry a second stream as relays.
        t_in, t_out = torch.from_numpy(h_wav[0]), torch.from_numpy(h_out[0])
        assert t_in.is_pinned() and t_out.is_pinned()
        sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        reps = 6

        def copy_pass(active):
            got = (0.0, 0.0)
            for timed in (False, True):
                barrier()
                if active:
                    with torch.cuda.stream(sa):
                        ev[0].record()
                        for _ in range(reps):
                            d_wav.copy_(t_in, non_blocking=True)
                        ev[1].record()
                    with torch.cuda.stream(sb):
                        ev[2].record()
                        for _ in range(reps):
                            t_out.copy_(d_out, non_blocking=True)
                        ev[3].record()
                    torch.cuda.synchronize()
                    got = (reps * total * 4 / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9, reps * total * 4 / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9)
            barrier()
            return [list(x) for x in zip(per_rank(got[0]), per_rank(got[1]))]

        link = copy_pass(True)
        needed = total * 4 / 14.0e-3 / 1e9              # one direction's bytes per ~14 ms step
        worst = [min(a, b) for a, b in link]
        slow, fast = relay_candidates(worst, needed)
        relay_of = [-1] * world
        if args.relay == "auto" and slow:
            link_fast_only = copy_pass(rank in fast)
            relay_of = plan_relays(worst, needed, relay_gbs=[min(a, b) for a, b in link_fast_only])
        if os.environ.get("RCED_BENCH_FORCE_RELAY"):      # experiment: "4,5,6,7,-1,-1,-1,-1"
            relay_of = [int(x) for x in os.environ["RCED_BENCH_FORCE_RELAY"].split(",")]
        mine = relay_of[rank]
        if mine >= 0:
            try:
                eng.host_set_relay(mine)                 # ranks are local GPU indices on one node
            except _lib.RcedError as exc:                # e.g. the two GPUs are not peers: copy directly
                sys.stderr.write("rank %d: no relay through GPU %d (%s)\n" % (rank, mine, exc))
                mine = -1
        relay_of = [int(x) for x in per_rank(mine)]      # what every rank really does
        barrier()

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
    ms_total, (k1_ms, k2_ms, k3_ms, dev_rank_ms) = timed_device_steps(args.steps, 0)
    launches = lib.rced_launch_count() - launches0
    audio_s_per_step = N_UTT * UTT_SAMPLES / SAMPLE_RATE
    value = world * audio_s_per_step * args.steps / (ms_total * 1e-3)
    tc_status = None
    if args.variant == "tc":
        amax, perr = eng.tc_status()
        tc_status = {"max_abs_activation_scaled_domain": amax, "protocol_error": perr,
                     "ffma_fallback_ran": bool(not amax <= 65504.0 or perr != 0)}

    # ---------------- end to end through the host-buffer C entry point (`e2e`) -----------------
    # rced_enhance_host_async: page-locked host waveforms -> H2D -> K1, K2, K3 -> D2H -> page-locked host output, chunked and
    # pipelined over the library's streams; consecutive steps are queued behind each other (two buffer sets) and the
    # timed region ends with rced_host_sync, so every step's copies are inside it.
    def e2e_steps(n):
        for i in range(n):
            eng.enhance_host(h_wav[i & 1], h_out[i & 1], tables, sync=False)
        eng.host_sync()

    e2e_steps(max(2, min(args.warmup, 3)))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    t0 = time.perf_counter()
    e2e_steps(args.steps)          # returns when the last step's output has arrived in host memory
    wall = time.perf_counter() - t0
    f1.record()
    barrier()
    e2e_rank_ms = per_rank(wall * 1e3 / args.steps)
    numa_nodes = per_rank(gpu_numa_node())
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), wall * 1e3))
    e2e_value = world * audio_s_per_step * args.steps / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- the other network kernel, a few steps ------------------------------------
    eng.set_variant(other)
    o_steps = max(3, min(args.steps, 5))
    o_ms_total, (_, o_k2_ms, _, _) = timed_device_steps(o_steps, 2)
    o_value = world * audio_s_per_step * o_steps / (o_ms_total * 1e-3)
    eng.set_variant(args.variant)

    # ---------------- BASELINE.json configs[4]: one job partitioned across the ranks -----------
    sweep = None
    if args.sweep_utterances > 0:   # every rank takes part (before the non-zero ranks leave)
        # the job is cut into batches of N_UTT utterances (the device-resident pool tiled; H2D is not in the timed region),
        # batch i goes to rank i % world, no collective
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
                             "B200 (BASELINE.json configs[4]); inputs resident in HBM (pool tiled)" % (job_utts, n_batches, N_UTT, world),
                 "utterances": job_utts, "audio_seconds": job_utts * UTT_SAMPLES / SAMPLE_RATE, "n_gpus": world,
                 "job_seconds": job_ms * 1e-3, "scaling": "strong", "network_kernel": args.variant,
                 "value": job_utts * UTT_SAMPLES / SAMPLE_RATE / (job_ms * 1e-3), "unit": UNIT}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_valid = 2.0 * lib.rced_mac_per_frame(eng.arch, 1) * rows
    achieved = flops_valid / (k2_ms * 1e-3) / 1e12
    o_achieved = flops_valid / (o_k2_ms * 1e-3) / 1e12
    share = k2_ms * args.steps / ms_total
    sm_mhz = clocks.get("sm_mhz") if clocks else None
    if args.variant == "tc":
        main_roof = roofline_tc(achieved, ffma_peak, flops_valid, k2_ms, share, rows, tc_status, sm_mhz)
        ffma_roof = roofline_ffma(o_achieved, ffma_peak, flops_valid, o_k2_ms)
        ffma_value, ffma_steps = o_value, o_steps
    else:
        main_roof = roofline_ffma(achieved, ffma_peak, flops_valid, k2_ms)
        main_roof["kernel_share_of_step"] = share
        ffma_roof, ffma_value, ffma_steps = main_roof, value, args.steps
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "ms_per_step_by_rank": dev_rank_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16x3 (FP16 hi/lo split, 3 products per multiply, FP32 accumulate; scale-invariant, 1e-6 of float64)"
                 if args.variant == "tc" else "f32",
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
                   "skip_storage": "global scratch (L2), one region per resident CTA" if args.variant == "tc" else "tensor memory"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": total * 4, "d2h_bytes_per_step": total * 4,
                "ms_per_step_by_rank": e2e_rank_ms, "gpu_numa_node_by_rank": [int(x) for x in numa_nodes],
                "host_cores": host_cores(),
                "host_link_gbs_by_rank": link, "host_link_gbs_fast_ranks_only": link_fast_only, "relay_gpu_by_rank": relay_of,
                "relay_note": None if not relay_of or max(relay_of) < 0 else
                              "ranks with a relay move their waveforms host -> relay GPU -> NVLink -> own GPU and back "
                              "(rced_host_set_relay): their own path to host memory is shared and too slow for the stream",
                "api": "rced_enhance_host_async (Enhancer.enhance_host): page-locked host waveforms -> H2D -> K1, K2, K3 -> D2H -> "
                       "page-locked host output, chunks of ~131072 spectrogram rows through the library's copy-in / compute / "
                       "copy-out streams; the steps alternate between two host buffer sets and are queued behind each other, "
                       "rced_host_sync closes the timed region (host wall clock)"},
        "gpu_launches": int(launches),
        "roofline": main_roof,
        # the metric BASELINE.json names for the network kernel: fraction of the FP32 FFMA peak reached by the FP32 kernel
        "fp32_ffma": {"value": ffma_value, "unit": UNIT, "steps": ffma_steps, "roofline": ffma_roof,
                      "what": "the same step with the FP32 FFMA network kernel (rced_net_kernel), device-resident"},
        "roofline_k1": roofline_hbm("rced_stft_kernel", k1_ms, rows, 512 + 516 + 1032,
                                    "512 B waveform in + 516 B magnitude + 1032 B phase out per frame (SURVEY.md 8d)",
                                    "r02_k1_ncu.json", ["rced_stft.cu", "rced_fft.cuh"]),
        "roofline_k3": roofline_hbm("rced_istft_kernel<512>", k3_ms, rows, 516 + 1032 + 512,
                                    "516 B prediction + 1032 B phase in + 512 B waveform out per frame (SURVEY.md 8d)",
                                    "r02_k3_ncu.json", ["rced_istft.cu", "rced_fft.cuh"]),
        "clocks": clocks,
    }
    if args.variant != "tc":
        result["tensor_core"] = {"value": o_value, "unit": UNIT, "steps": o_steps, "kernel_ms": o_k2_ms}
    if sweep is not None:
        result["sweep"] = sweep
    if not args.no_latency and world == 1:
        result["latency"] = latency_block(eng, pool, torch)
    if not args.no_cpu_baseline and world == 1:
        from oracle import network as onet
        result["cpu_baseline"] = cpu_baseline(pool, onet.random_weights(NET_WORK, seed=0, randomize_bn=False))
    emit(result)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
