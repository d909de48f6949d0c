"""numpy emulator of the fused network kernel's shared-memory data flow (development aid).

It executes, for one warp, exactly the address arithmetic of csrc/rced_net.cu -- the packed
weight image produced by rced_pack_weights, the per-warp slot with its row stride 136 / zero
halos, in-place layer updates, the "wide" layout of the (1,129) layer, the staging rows and the
halo restoration between frames, the channel split of every layer over the warps of a frame and
the tap pairing of the (1,129) layer -- vectorised over the 32 lanes.  It needs no GPU, only the
host functions of librced_b200.so, and is compared with the oracle in tests/test_host_cpu.py.
"""
import ctypes

import numpy as np

RS, BIN0, WS, WBIN0, KP = 136, 8, 196, 64, 132


def layout(lib, arch):
    nl = lib.rced_num_layers(arch)
    out = (ctypes.c_int64 * (8 + 4 * nl))()
    assert lib.rced_debug_layout(arch, out, len(out)) == 0
    shapes = []
    for i in range(nl):
        v = [ctypes.c_int() for _ in range(4)]
        assert lib.rced_layer_shape(arch, i, *v) == 0
        shapes.append(tuple(x.value for x in v))
    return dict(nl=nl, stage_row=out[0], slot_floats=out[1], wide_floats=out[2], split=out[5], combine_off=out[6],
                w_off=[out[8 + 4 * i] for i in range(nl)], b_off=[out[9 + 4 * i] for i in range(nl)],
                save=[out[10 + 4 * i] for i in range(nl)], add=[out[11 + 4 * i] for i in range(nl)],
                shapes=shapes)


def pack(lib, arch, folded):
    folded = np.ascontiguousarray(folded, np.float32)
    n = lib.rced_packed_weight_count(arch)
    packed = np.zeros(n, np.float32)
    rc = lib.rced_pack_weights(arch, folded.ctypes.data_as(ctypes.c_void_p), folded.size,
                               packed.ctypes.data_as(ctypes.c_void_p), packed.size)
    assert rc == 0, lib.rced_last_error()
    return packed


def run(lib, arch, folded, mag, relu_flags, after_flags, dtype=np.float64):
    """mag [T,129] of ONE utterance -> pred [T,129], all frames through one emulated warp slot.
    relu_flags / after_flags: per-layer flags (the emulator takes them from the caller so that it
    does not share the C++ table for those)."""
    lay = layout(lib, arch)
    packed = pack(lib, arch, folded).astype(dtype)
    nl, SR = lay["nl"], lay["stage_row"]
    T = mag.shape[0]
    slot = np.zeros(lay["slot_floats"] + 64, dtype)          # +64: reads of discarded lanes past the end
    tmem = np.zeros((32, 512), dtype)                        # [lane][column]
    lanes = np.arange(32)
    pred = np.zeros((T, 129), dtype)

    def prefetch(g):
        for i in range(9 * 7):
            slot[(SR + i // 7) * RS + 1 + (i % 7)] = 0
        for dt in range(8):
            r = g + dt - 3
            dst = (SR + dt) * RS + BIN0
            slot[dst:dst + 129] = mag[r] if 0 <= r < T else 0

    SPLIT = lay["split"]
    comb = lay["combine_off"]
    prefetch(0)
    for g in range(T):
        for li in range(nl - 1):
            kh, kw, cin, cout = lay["shapes"][li]
            CIN = kh if li == 0 else cin
            CH = ((cout + SPLIT - 1) // SPLIT + 1) & ~1        # ch_part
            CIB = (kw * CH + 3) & ~3                           # ci_block
            BP = (CH + 3) & ~3
            PADL = (kw - 1) // 2
            wide = PADL > 4
            XB = (8 if wide else 4) - PADL
            NX = 20 if wide else 12
            in0 = SR * RS if li == 0 else 0
            pre_add = lay["add"][li] >= 0 and not after_flags[li]
            out_wide = li == nl - 2
            results = []
            for part in range(SPLIT):                          # the warps of the frame: all read first ...
                W = packed[lay["w_off"][li] + part * CIN * CIB:]
                B = packed[lay["b_off"][li] + part * BP:]
                acc = np.zeros((32, 4, CH), dtype)
                cl = np.minimum(lanes, CH - 1)
                if pre_add:
                    col = lay["add"][li] + part * (4 * CH + 1)
                    for c in range(CH):
                        acc[:, :, c] = tmem[:, col + 4 * c: col + 4 * c + 4] + B[c]
                    tacc = tmem[:, col + 4 * CH] + B[cl]
                else:
                    acc[:] = B[:CH][None, None, :]
                    tacc = B[cl].copy()
                for ci in range(CIN):
                    base = in0 + (0 if wide else 4) + 4 * lanes + ci * RS
                    x = slot[base[:, None] + np.arange(NX)[None, :]]                  # [lane][NX]
                    for k in range(kw):
                        w = W[ci * CIB + k * CH: ci * CIB + k * CH + CH]
                        for f in range(4):
                            acc[:, f, :] += x[:, XB + f + k][:, None] * w[None, :]
                for ci in range(CIN):
                    for k in range(PADL + 1):
                        xt = slot[in0 + BIN0 + 128 - PADL + ci * RS + k]
                        tacc = tacc + xt * W[ci * CIB + k * CH + cl]
                results.append((acc, tacc))
            if out_wide:
                slot[:lay["wide_floats"]] = 0
            for part in range(SPLIT):                          # ... then write
                v, t = results[part][0].copy(), results[part][1].copy()
                if relu_flags[li]:
                    v = np.maximum(v, 0)
                    t = np.maximum(t, 0)
                if lay["add"][li] >= 0 and after_flags[li]:
                    col = lay["add"][li] + part * (4 * CH + 1)
                    for c in range(CH):
                        v[:, :, c] += tmem[:, col + 4 * c: col + 4 * c + 4]
                    t = t + tmem[:, col + 4 * CH]
                if lay["save"][li] >= 0:
                    col = lay["save"][li] + part * (4 * CH + 1)
                    for c in range(CH):
                        tmem[:, col + 4 * c: col + 4 * c + 4] = v[:, :, c]
                    tmem[:, col + 4 * CH] = t
                for c in range(CH):
                    cg = part * CH + c
                    if cg >= cout:
                        assert not v[:, :, c].any()            # zero padding channels stay zero
                        continue
                    o = (cg * WS + WBIN0 if out_wide else cg * RS + BIN0) + 4 * lanes
                    for f in range(4):
                        slot[o + f] = v[:, f, c]
                for lane in range(CH):
                    cg = part * CH + lane
                    if cg >= cout:
                        continue
                    o = cg * WS + WBIN0 + 128 if out_wide else cg * RS + BIN0 + 128
                    slot[o] = t[lane]
        # next frame's input lands while the final layer runs
        if g + 1 < T:
            prefetch(g + 1)
        li = nl - 1
        cin = lay["shapes"][li][2]
        Wf = packed[lay["w_off"][li]:]
        Sf = Wf[cin * KP:]
        bias = packed[lay["b_off"][li]]
        CP = (cin + SPLIT - 1) // SPLIT
        tot = np.zeros((32, 4), dtype)
        ttot = dtype(0) if not isinstance(dtype, type) else dtype(0)
        for part in range(SPLIT):
            a = np.zeros((32, 4, 2), dtype)       # (even-tap, odd-tap) partial sums
            lone = np.zeros((32, 4), dtype)
            for ci in range(part * CP, min(cin, part * CP + CP)):
                row = ci * WS + 4 * lanes
                X = slot[row[:, None] + np.arange(132)[None, :]]                      # x[4l .. 4l+131]
                w, sh = Wf[ci * KP: ci * KP + KP], Sf[ci * KP: ci * KP + KP]
                lone[:, 1] += X[:, 1] * w[0]
                lone[:, 3] += X[:, 3] * w[0]
                for q in range(32):
                    xa, xb = X[:, 4 * q: 4 * q + 4], X[:, 4 * q + 4: 4 * q + 8]
                    wq, sq = w[4 * q: 4 * q + 4], sh[4 * q: 4 * q + 4]
                    a[:, 0] += xa[:, 0:2] * wq[0:2] + xa[:, 2:4] * wq[2:4]
                    a[:, 2] += xa[:, 2:4] * wq[0:2] + xb[:, 0:2] * wq[2:4]
                    a[:, 1] += xa[:, 2:4] * sq[0:2] + xb[:, 0:2] * sq[2:4]
                    a[:, 3] += xb[:, 0:2] * sq[0:2] + xb[:, 2:4] * sq[2:4]
                lone[:, 0] += X[:, 128] * w[128]
                lone[:, 2] += X[:, 130] * w[128]
            tl = np.zeros(32, dtype)
            for ci in range(part * CP, min(cin, part * CP + CP)):
                r = ci * WS + 128
                tl += slot[r + lanes] * Wf[ci * KP + lanes] + slot[r + lanes + 32] * Wf[ci * KP + lanes + 32]
                tl[0] += slot[r + 64] * Wf[ci * KP + 64]
            r4 = a.sum(axis=2) + lone
            if part != 0:
                slot[comb + 4 * lanes[:, None] + np.arange(4)[None, :]] = r4
                slot[comb + 128] = tl.sum()
                tot += slot[comb + 4 * lanes[:, None] + np.arange(4)[None, :]]
                ttot = ttot + slot[comb + 128]
            else:
                tot += r4
                ttot = ttot + tl.sum()
        for lane in range(32):
            pred[g, 4 * lane: 4 * lane + 4] = tot[lane] + bias
        pred[g, 128] = ttot + bias
        for i in range((SR + 1) * 7):
            slot[(i // 7) * RS + 1 + (i % 7)] = 0
    return pred
