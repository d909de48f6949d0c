"""CPU checks of the index algebra of K1 / K3's warp FFT (csrc/rced_fft.cuh, csrc/rced_istft.cu): the per-lane twiddle
table, and the staging buffer of K3's kept outputs (which lane stores what where, what each lane reads back, and that neither
the stores nor the loads meet on a shared-memory bank).  The constants are read from the sources."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fullycnnspeechenhancement_b200", "csrc")


def _src(name):
    with open(os.path.join(CSRC, name)) as f:
        return f.read()


def bitrev5(l):
    return int("{:05b}".format(l)[::-1], 2)


def test_k3_staging_of_the_kept_outputs_is_a_bijection_without_bank_conflicts():
    src = _src("rced_istft.cu")
    row = int(re.search(r"constexpr int kZzRow = (\d+);", src).group(1))
    assert "s_zz[warp][kZzRow * b + (br & 7)] = q;" in src and "v = s_zz[warp][kZzRow * (lane & 3) + (lane >> 2)];" in src
    assert "if ((br >> 3) == half)" in src
    for half in (0, 1):
        where = {}                                   # output index n (relative to 32 * half) -> slot
        for b in range(4):
            slots = []
            for lane in range(32):
                br = bitrev5(lane)
                if (br >> 3) != half:
                    continue
                n = 4 * br + b - 32 * half           # the transform leaves lane l with outputs 4 bitrev5(l) + b
                assert 0 <= n < 32
                where[n] = row * b + (br & 7)
                slots.append(where[n])
            # one store instruction: its 8 active lanes write 8 consecutive 16-byte slots = 128 contiguous bytes
            assert sorted(slots) == list(range(row * b, row * b + 8))
        assert sorted(where) == list(range(32)) and len(set(where.values())) == 32
        for lane in range(32):                        # lane l reads output 32 * half + l
            assert row * (lane & 3) + (lane >> 2) == where[lane]
        for q in range(4):                            # a 16-byte load is served a quarter-warp at a time: 8 slots, 32 banks
            slots = [row * (lane & 3) + (lane >> 2) for lane in range(8 * q, 8 * q + 8)]
            assert len({s % 8 for s in slots}) == 8
    assert 4 * row >= row * 3 + 8                     # the buffer holds the last row


def test_per_lane_twiddle_table_holds_the_strided_entries():
    src = _src("rced_fft.cuh")
    assert "if (row < 3) return 2 * lane * (row + 1);" in src and "return (lane & (s - 1)) * (128 / s);" in src
    tw = np.exp(-2j * np.pi * np.arange(256) / 256)
    table = np.zeros((8, 32), complex)
    for row in range(8):
        for lane in range(32):
            if row < 3:
                idx = 2 * lane * (row + 1)
                one = False
            else:
                s = 16 >> (row - 3)
                idx = (lane & (s - 1)) * (128 // s)
                one = (lane & s) == 0                 # lanes that only add hold 1
            assert 0 <= idx < 256
            table[row, lane] = 1.0 if one else tw[idx]
    # rows 0..2: the twiddles between the radix-4 step and the 32-point transforms, W128^(lane b)
    for b in (1, 2, 3):
        assert np.allclose(table[b - 1], np.exp(-2j * np.pi * np.arange(32) * b / 128))
    # rows 3..7: stage s multiplies the lanes with bit s by W_(2s)^(lane mod s)
    for r, s in enumerate((16, 8, 4, 2, 1)):
        lanes = np.arange(32)
        want = np.where((lanes & s) != 0, np.exp(-2j * np.pi * (lanes & (s - 1)) / (2 * s)), 1.0)
        assert np.allclose(table[3 + r], want)
    assert np.allclose(table[7], 1.0)                 # the last stage's twiddle is 1 everywhere: the kernels skip it
