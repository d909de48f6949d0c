"""numpy emulator of the tensor-core network kernel's data flow (development / test aid).

It executes exactly the address arithmetic of csrc/rced_net_tc.cu on the CPU: the FP16 hi/lo
weight image and bias table produced by rced_tc_pack_weights, the per-unit A-descriptor table of
rced_tc_layout (start offset and leading-dimension offset in 16-byte units, rows linear at 16
bytes), the flattened (frame, bin) row space with its zero halo rows, in-place plane updates,
the two instructions per K step (A_hi x [Whi | Wlo'] into columns [0, 2 NP), A_lo' x Whi into columns
[NP, 2 NP), residuals stored times 2^11) with FP32 accumulation, the power-of-two scalings (per step weight
scale, per frame activation scale chosen from the frame's input rows and the largest bias), the
even / odd frame copies written by the last conv layer, the row-shifted blocks of the (1,129) layer
(reads before plane 0 land in the zero front rows) with its 32-column diagonal sums, and the FP32
skip scratch.
It needs no GPU, only the host functions of librced_b200.so; tests/test_tc_cpu.py compares it
with the float64 oracle, which pins both the layout and the accuracy of the FP16 x3 split.
"""
import ctypes

import numpy as np

BINS = 129


def layout(lib, arch):
    out = (ctypes.c_int64 * 4096)()
    assert lib.rced_tc_layout(arch, out, len(out)) == 0, lib.rced_last_error()
    ns, nu = out[0], out[1]
    lay = dict(ns=ns, nu=nu, image_bytes=out[2], smem=out[3], plane16=out[4], lead=out[5], fs=out[6], fb=out[7],
               tiles=out[8], lo16=out[9], final_n=out[10], skip_floats=out[11], final_shifts=out[12], front_rows=out[13],
               steps=[], units=[])
    for s in range(ns):
        o = out[16 + 6 * s: 22 + 6 * s]
        lay["steps"].append(dict(units=o[0], unit_base=o[1], np=o[2], tile_bytes=o[3], w_off=o[4], final=bool(o[5])))
    u0 = 16 + 6 * ns
    lay["units"] = [(out[u0 + 2 * i], out[u0 + 2 * i + 1]) for i in range(nu)]
    return lay


def pack(lib, arch, folded):
    folded = np.ascontiguousarray(folded, np.float32)
    img = np.zeros(lib.rced_tc_image_bytes(arch), np.uint8)
    bias = np.zeros(lib.rced_tc_bias_count(arch), np.float32)
    rc = lib.rced_tc_pack_weights(arch, folded.ctypes.data_as(ctypes.c_void_p), folded.size,
                                  img.ctypes.data_as(ctypes.c_void_p), img.size,
                                  bias.ctypes.data_as(ctypes.c_void_p), bias.size)
    assert rc == 0, lib.rced_last_error()
    return img, bias


LO_SHIFT = 11        # residuals are stored times 2^11 (csrc/rced_tc.cuh: kLoShift)
FRAME_TOP = 4        # a frame's reference magnitude lands in [2^3, 2^4) (kFrameTop)
BIAS_REF_SLOT = 16   # kBiasRefSlot


def split(v):
    v = np.asarray(v, np.float32)
    hi = v.astype(np.float16)
    lo = ((v - hi.astype(np.float32)) * np.float32(2.0 ** LO_SHIFT)).astype(np.float16)
    return hi, lo


def frame_scale(ref):
    """Power of two that puts `ref` (float32, finite, >= 0) into [2^(FRAME_TOP-1), 2^FRAME_TOP): exponent arithmetic on the
    bit pattern, like the prefetch warp."""
    bits = int(np.float32(ref).view(np.uint32))
    if bits == 0:
        return np.float32(1.0)
    k = 127 + FRAME_TOP - 1 - (bits >> 23)
    k = max(-100, min(100, k))
    return np.float32(2.0 ** k)


def run(lib, arch, folded, mag, row_off, table):
    """mag [rows,129] float32 (packed ragged batch), row_off [n_utt+1] -> pred [rows,129] float32.
    ``table``: the model's layer table (scope / skip / act / skip_after_act per layer) -- the emulator
    takes the graph wiring from the caller so that it does not share the C++ table for it."""
    lay = layout(lib, arch)
    img, bias = pack(lib, arch, folded)
    halfs = img.view(np.float16)
    P16, LEAD, FS, FB, TILES, LO16 = lay["plane16"], lay["lead"], lay["fs"], lay["fb"], lay["tiles"], lay["lo16"]
    ROWS = TILES * 128
    mag = np.asarray(mag, np.float32)
    row_off = np.asarray(row_off, np.int64)
    total = mag.shape[0]
    pred = np.zeros((total, BINS), np.float32)
    nl = len(table)
    ns = lay["ns"]
    scopes = [L["scope"] for L in table]
    amax = 0.0
    # per conv step: power of two between the domain of the skip tensor it adds and its own; output layer: inverse of
    # the total weight scale (the biases of the table are already in their step's domain)
    aux = bias[ns * 32: ns * 32 + ns].astype(np.float32)
    bias_ref = np.float32(bias[ns * 32 + BIAS_REF_SLOT])
    lo_inv = np.float32(2.0 ** -LO_SHIFT)

    r_idx = np.arange(ROWS)
    fi_of, b_of = r_idx // FS, r_idx % FS

    for g0 in range(0, total, FB):
        nf = min(FB, total - g0)
        valid = (fi_of < nf) & (b_of < BINS)
        # flat[16-byte unit][8 halfs]: zero front rows, hi planes, lo planes, like the shared-memory array (FR: index
        # of plane 0's first unit); what follows the lo planes in shared memory (the weight buffer) is finite garbage
        FR = lay["front_rows"]
        flat = np.zeros((FR + 2 * LO16 + 128, 8), np.float16)
        flat[FR + 2 * LO16:] = 3.0
        # ---- staging of the first layer's input ("channel" = time tap)
        v = np.zeros((ROWS, 8), np.float32)
        fscale = np.ones(8, np.float32)          # per frame: power-of-two scale of its domain
        for fi in range(nf):
            g = g0 + fi
            u = int(np.searchsorted(row_off, g, side="right") - 1)
            lo_, hi_ = row_off[u], row_off[u + 1]
            fmax = np.float32(0)
            for tt in range(8):
                src = g + tt - 3
                if lo_ <= src < hi_:
                    v[fi * FS: fi * FS + BINS, tt] = mag[src]
                    fmax = max(fmax, np.abs(mag[src]).max())
            fscale[fi] = frame_scale(max(fmax, bias_ref))
            v[fi * FS: fi * FS + BINS] *= fscale[fi]
        for fi in range(nf, 8):
            fscale[fi] = frame_scale(bias_ref)
        rscale = fscale[np.minimum(fi_of, 7)]    # per row
        amax = max(amax, float(np.abs(v).max()))
        h, l = split(v)
        flat[FR + LEAD: FR + LEAD + ROWS] = h
        flat[FR + LO16 + LEAD: FR + LO16 + LEAD + ROWS] = l
        saved = {}
        outp = np.full((2, ROWS), np.nan, np.float32)
        for s, st in enumerate(lay["steps"]):
            NP = st["np"]
            rows_b = (3 if st["final"] else 2) * NP
            tile_h = st["tile_bytes"] // 2
            D = np.zeros((ROWS, 64), np.float32)
            for t in range(TILES):
                sl = slice(128 * t, 128 * t + 128)
                for u in range(st["units"]):
                    off16, lbo16 = lay["units"][st["unit_base"] + u]
                    tile = halfs[st["w_off"] // 2 + u * tile_h: st["w_off"] // 2 + (u + 1) * tile_h].reshape(2, rows_b, 8)
                    B = np.concatenate([tile[0], tile[1]], axis=1).astype(np.float32)   # [rows_b][16]
                    for odd in range(2 if st["final"] else 1):
                        if odd and t in (0, TILES - 1):
                            continue
                        start = FR + LEAD + 128 * t + off16 + odd * 2 * P16
                        a_hi = np.concatenate([flat[start: start + 128], flat[start + lbo16: start + lbo16 + 128]],
                                              axis=1).astype(np.float32)
                        a_lo = np.concatenate([flat[LO16 + start: LO16 + start + 128],
                                               flat[LO16 + start + lbo16: LO16 + start + lbo16 + 128]], axis=1).astype(np.float32)
                        if not st["final"]:
                            pa = a_hi @ B.T
                            if u == 0:
                                D[sl, :rows_b] = pa
                            else:
                                D[sl, :rows_b] += pa
                            D[sl, NP:2 * NP] += a_lo @ B[:NP].T      # both cross products carry 2^11
                        else:
                            c0 = odd * NP
                            pa = a_hi @ B[:NP].T
                            if u == 0:
                                D[sl, c0:c0 + NP] = pa
                            else:
                                D[sl, c0:c0 + NP] += pa
                            D[sl, c0:c0 + NP] += a_lo @ B[2 * NP:].T      # scaled residual x (Whi 2^-11)
                            D[sl, c0:c0 + NP] += a_hi @ B[NP:2 * NP].T    # hi x Wlo (not scaled)
            if not st["final"]:
                L = table[s]
                cout = L["cout"]
                cg = (cout + 7) // 8
                x = (D[:, :cg * 8] + D[:, NP:NP + cg * 8] * lo_inv).astype(np.float32)
                x = (x + bias[s * 32: s * 32 + cg * 8][None, :] * rscale[:, None]).astype(np.float32)
                if L["skip"] is not None and not L["skip_after_act"]:
                    x = (x + saved[L["skip"]] * aux[s]).astype(np.float32)
                if L["act"]:
                    x = np.maximum(x, 0)
                if L["skip"] is not None and L["skip_after_act"]:
                    x = (x + saved[L["skip"]] * aux[s]).astype(np.float32)
                x = np.where(valid[:, None], x, 0).astype(np.float32)
                amax = max(amax, float(np.abs(x).max()))
                saved[scopes[s]] = x[:, :cg * 8]      # FP32 rows, written from the registers
                h, l = split(x)
                last = s == nl - 2
                odd_row = (fi_of % 2 == 1)[:, None]
                for g in range(cg):
                    hg, lg = h[:, 8 * g: 8 * g + 8], l[:, 8 * g: 8 * g + 8]
                    if not last:
                        flat[FR + g * P16 + LEAD: FR + g * P16 + LEAD + ROWS] = hg
                        flat[FR + LO16 + g * P16 + LEAD: FR + LO16 + g * P16 + LEAD + ROWS] = lg
                    else:   # even-frame copy in planes g, odd-frame copy in planes g + 2, zeros in the other one
                        z = np.zeros_like(hg)
                        flat[FR + g * P16 + LEAD: FR + g * P16 + LEAD + ROWS] = np.where(odd_row, z, hg)
                        flat[FR + LO16 + g * P16 + LEAD: FR + LO16 + g * P16 + LEAD + ROWS] = np.where(odd_row, z, lg)
                        flat[FR + (g + 2) * P16 + LEAD: FR + (g + 2) * P16 + LEAD + ROWS] = np.where(odd_row, hg, z)
                        flat[FR + LO16 + (g + 2) * P16 + LEAD: FR + LO16 + (g + 2) * P16 + LEAD + ROWS] = np.where(odd_row, lg, z)
            else:
                # E[r][n] belongs to output row r - n; per 32-row block the diagonal sums are split into the part that
                # stays in the block (outp[0]) and the part that falls into the previous block (outp[1]); a sum is
                # stored only on a valid row of a frame of the accumulator's parity
                for t in range(TILES):
                    for q in range(4):
                        r0 = 128 * t + 32 * q
                        for odd in range(2):
                            if odd and t in (0, TILES - 1):
                                continue
                            own = np.zeros(32, np.float32)
                            prev = np.zeros(32, np.float32)
                            for n in range(NP):
                                for lane in range(32):
                                    d = lane - n
                                    if d >= 0:
                                        own[d] += D[r0 + lane, odd * NP + n]
                                    else:
                                        prev[d + 32] += D[r0 + lane, odd * NP + n]
                            for m in range(32):
                                ra, rb = r0 + m, r0 + m - 32
                                if valid[ra] and fi_of[ra] % 2 == odd:
                                    outp[0, ra] = own[m]
                                if rb >= 0 and valid[rb] and fi_of[rb] % 2 == odd:
                                    outp[1, rb] = prev[m]
        bias_f = bias[(nl - 1) * 32]
        for fi in range(nf):
            inv = np.float32(aux[nl - 1]) / fscale[fi]    # a power of two
            pred[g0 + fi] = (outp[0, fi * FS: fi * FS + BINS] + outp[1, fi * FS: fi * FS + BINS]) * inv + bias_f
    return pred, amax
