"""Digest of an ncu report: key raw metrics of the first kernel plus the stall-sample breakdown
by reason, opcode and code region (source page).  Usage: python tools/ncu_digest.py X.ncu-rep
       python tools/ncu_digest.py X.ncu-rep json <kernel name substring> <out.json> <csrc file> [<csrc file> ...]
The json mode writes what bench.py reports next to a kernel's roofline (DRAM bytes per launch, the throughput of the
units that bind it) together with the sha256 of the kernel sources the capture was taken from."""
import collections
import csv
import io
import subprocess
import sys


KERNEL = None   # regex selecting one kernel of a report that holds several (argument `kernel=<regex>`)


def page(rep, name):
    cmd = ["ncu", "-i", rep, "--page", name, "--csv"]
    if KERNEL:
        cmd += ["--kernel-name", "regex:" + KERNEL]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, bucket=256):
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.per_cycle_active",
            "smsp__average_warp_latency_per_inst_issued.ratio",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sass__inst_executed_shared_loads", "sm__cycles_elapsed.avg.per_second",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
            # the units that bind the tensor-core kernel and K1 / K3
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum"]
    for h, u, v in zip(hdr, units, vals):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.004):
            print("%s [%s] = %s" % (h, u, v))
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot = sum(f(r, "# Samples") for r in data)
    print("\nstall samples: total %d" % tot)
    reasons = [h for h in hdr if h.startswith("stall_") and "(" not in h]
    for k in sorted(reasons, key=lambda k: -sum(f(r, k) for r in data)):
        s = sum(f(r, k) for r in data)
        if s / tot > 0.002:
            print("  %-24s %.3f" % (k, s / tot))
    d = collections.defaultdict(lambda: [0.0, 0.0])
    for r in data:
        t = r[ix["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else t[0]) if t else "?"
        d[op][0] += f(r, "# Samples")
        d[op][1] += f(r, "Instructions Executed")
    print("\nopcode: share of samples, warp instructions executed")
    ti = sum(v[1] for v in d.values())
    for op, e in sorted(d.items(), key=lambda kv: -kv[1][0])[:14]:
        print("  %-12s %.3f  %.4g (%.3f of instr)" % (op, e[0] / tot, e[1], e[1] / ti))
    print("\ncode regions of %d instructions: samples, share, and per-reason fractions" % bucket)
    for b in range(0, len(data), bucket):
        chunk = data[b:b + bucket]
        s = sum(f(r, "# Samples") for r in chunk)
        if s / tot < 0.004:
            continue
        parts = sorted(((sum(f(r, k) for r in chunk) / s, k[6:]) for k in reasons), reverse=True)[:4]
        print("  %5d %8d %.3f  %s" % (b, s, s / tot, "  ".join("%s %.2f" % (k, v) for v, k in parts)))


def segments(rep, frames_per_warp_total):
    """Consecutive instructions with the same execution count form a segment (a loop body or a
    once-per-layer stretch); prints each segment's share of the stall samples."""
    rows = page(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot = sum(f(r, "# Samples") for r in data)
    segs, cur = [], None
    for i, r in enumerate(data):
        e = round(f(r, "Instructions Executed") / frames_per_warp_total, 1)
        if cur is None or abs(cur[2] - e) > 0.15:
            cur = [i, i, e, 0.0, 0, 0]
            segs.append(cur)
        cur[1] = i
        cur[3] += f(r, "# Samples")
        cur[4] += 1 if "FFMA2" in r[ix["Source"]] else 0
        cur[5] += 1 if "LDS" in r[ix["Source"]] else 0
    loop = once = 0.0
    print("\nsegments: instr range, executions per warp-frame, instructions, FFMA2, LDS, share of samples, pipe-ideal share")
    for s in segs:
        share = s[3] / tot
        if share > 0.003:
            print("  %5d-%5d x%5.1f n=%4d ffma2=%4d lds=%3d share %.4f" % (s[0], s[1], s[2], s[1] - s[0] + 1, s[4], s[5], share))
        if s[2] > 1.5:
            loop += share
        else:
            once += share
    print("  loop bodies %.3f, once-per-layer code %.3f" % (loop, once))


def to_json(rep, kernel_substr, out_path, sources):
    import hashlib
    import json
    import os
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if kernel_substr in r[ix["Kernel Name"]]]
    if not sel:
        raise SystemExit("no kernel matching %r in %s" % (kernel_substr, rep))
    r = sel[-1]

    def val(name, scale_units=True):
        if name not in ix or r[ix[name]] in ("", "n/a"):
            return None
        v = float(r[ix[name]].replace(",", ""))
        if scale_units:
            u = units[ix[name]].lower()
            exact = {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}      # times are reported in ms
            if u in exact:
                return v * exact[u]
            for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("usecond", 1e-3), ("msecond", 1.0),
                           ("nsecond", 1e-6), ("second", 1e3)):
                if u.startswith(pre):
                    return v * m
        return v
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for n in sources:
        with open(os.path.join(root, "fullycnnspeechenhancement_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    d = {
        "kernel": r[ix["Kernel Name"]], "report": os.path.basename(rep), "kernel_source_sha16": h.hexdigest()[:16],
        "kernel_sources": list(sources),
        "gpu_time_ms_under_ncu": val("gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd or 0.0) + (wr or 0.0),
        "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "l1tex_throughput_pct": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "lts_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "tensor_pipe_active_pct": val("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", False),
        "fma_pipe_active_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "registers_per_thread": val("launch__registers_per_thread", False),
        "warp_instructions": val("smsp__inst_executed.sum", False),
    }
    with open(out_path, "w") as f:
        json.dump(d, f, indent=1)
        f.write("\n")
    print(json.dumps(d))


if __name__ == "__main__":
    for a in list(sys.argv):
        if a.startswith("kernel="):
            KERNEL = a[7:]
            sys.argv.remove(a)
    if len(sys.argv) > 5 and sys.argv[2] == "json":
        to_json(sys.argv[1], sys.argv[3], sys.argv[4], sys.argv[5:])
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[2] == "segments":
        segments(sys.argv[1], float(sys.argv[3]))
        sys.exit(0)
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 256)
