"""Report of a K2-TC development trace (RCED_TC_TRACE=<file>, see rced_api.cu / rced_net_tc.cu):
clock64 stamps of CTA 0's second batch, one line of 8 per (step, tile).
events: 0 MMA issue begins, 1 committed, 2 epilogue past its waits, 3 accumulator in registers,
4 epilogue done, 5 issuer starts polling for the tile, 6 first unit issued, 7 all units issued.
Usage: python tools/tc_trace_report.py gpurun_out/tc_trace.txt [tiles=8]"""
import sys

import numpy as np


def main(path, tiles=8):
    a = np.loadtxt(path, dtype=np.int64).reshape(-1, tiles, 8)
    ns = a.shape[0]
    base = a[a > 0].min()
    r = np.where(a > 0, a - base, -1)
    print("step | per tile: poll>issue0 issue>commit commit>epi_go epi_go>regs regs>done | step span")
    prev_end = 0
    for s in range(ns):
        row = []
        for t in range(tiles):
            e = r[s, t]
            row.append("%5d %4d %4d %4d %4d %4d" % (e[5], e[0] - e[5], e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3]))
        first = r[s, :, 5][r[s, :, 5] >= 0].min()
        last = r[s, :, 4].max()
        mma = int((r[s, :, 1] - r[s, :, 0]).sum())
        print("%2d first_poll=%6d last_epi_done=%6d (+%5d vs prev) mma_issue_sum=%5d" % (s, first, last, last - prev_end, mma))
        for t in range(tiles):
            print("     t%d: %s" % (t, row[t]))
        prev_end = last
    print("total span %d cycles" % (r[:, :, 4].max() - r[:, :, 5][r[:, :, 5] >= 0].min()))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
