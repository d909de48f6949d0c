"""numpy lane-level emulator of the warp FFT used by csrc/rced_fft.cuh (development aid).

Mirrors the device code step by step -- 4 complex values per lane, radix-4 in registers,
five shfl_xor radix-2 stages -- so the index algebra can be checked against numpy.fft
without a GPU.  Run: python tools/fft_emulator.py
"""
import numpy as np


def brev5(l):
    return int('{:05b}'.format(l)[::-1], 2)


def fft128_warp(z, inverse=False):
    """z: complex[128] natural order.  Returns Z natural order, computed the way the warp does."""
    tw = np.exp(-2j * np.pi * np.arange(256) / 256)
    if inverse:
        tw = np.conj(tw)
    lanes = np.arange(32)
    v = np.stack([z[lanes + 32 * a] for a in range(4)], axis=0)          # v[a][lane]
    s0, s1, s2, s3 = v[0] + v[2], v[0] - v[2], v[1] + v[3], v[1] - v[3]
    mi = 1j if inverse else -1j
    y = np.stack([s0 + s2, s1 + mi * s3, s0 - s2, s1 - mi * s3], axis=0)  # y[b][lane]
    for b in range(4):
        y[b] = y[b] * tw[(2 * lanes * b) % 256]
    for s in (16, 8, 4, 2, 1):
        hi = (lanes & s) != 0
        # the per-lane twiddle table of the kernels: the stage's twiddle on the lanes with bit s, 1 on the others,
        # so that every lane runs the same two steps t = o -+ v, v = t * w
        w = np.where(hi, tw[(lanes & (s - 1)) * (128 // s)], 1.0)
        sg = np.where(hi, -1.0, 1.0)
        for b in range(4):
            o = y[b][lanes ^ s]                                            # shfl_xor
            y[b] = (sg * y[b] + o) * w
    Z = np.zeros(128, complex)
    for l in range(32):
        for b in range(4):
            Z[4 * brev5(l) + b] = y[b][l]
    return Z


def rfft256_via128(x):
    """real x[256] -> X[0..128] as K1 does it."""
    tw = np.exp(-2j * np.pi * np.arange(256) / 256)
    z = x[0::2] + 1j * x[1::2]
    Z = fft128_warp(z)
    X = np.zeros(129, complex)
    for k in range(129):
        Zk, Zn = Z[k & 127], Z[(128 - k) & 127]
        E = 0.5 * (Zk + np.conj(Zn))
        O = -0.5j * (Zk - np.conj(Zn))
        X[k] = E + tw[k] * O
    return X


def half_irfft(A):
    """A[0..128] Hermitian half spectrum (A[0], A[128] real) -> z[128] with
    r[2n] = Re z[n]/128, r[2n+1] = Im z[n]/128, r = irfft(A, 256)."""
    tw = np.exp(-2j * np.pi * np.arange(256) / 256)
    Z = np.zeros(128, complex)
    for k in range(128):
        Ak, An = A[k], A[128 - k]
        E = 0.5 * (Ak + np.conj(An))
        O = 0.5 * (Ak - np.conj(An)) * np.conj(tw[k])
        Z[k] = E + 1j * O
    return fft128_warp(Z, inverse=True)


def irfft512_first256(Y):
    """Y[0..128] arbitrary complex -> np.fft.irfft(Y, 512)[:256] as K3 does it."""
    k = np.arange(129)
    Ae = Y.astype(complex).copy()
    Ae[0] = Y[0].real
    Ae[128] = 2.0 * Y[128].real
    Ao = Y * np.exp(1j * np.pi * k / 256)
    Ao[0] = Y[0].real
    Ao[128] = -2.0 * Y[128].imag
    ze, zo = half_irfft(Ae), half_irfft(Ao)
    y = np.zeros(256)
    for n in range(64):
        y[4 * n + 0] = ze[n].real / 256
        y[4 * n + 1] = zo[n].real / 256
        y[4 * n + 2] = ze[n].imag / 256
        y[4 * n + 3] = zo[n].imag / 256
    return y


def irfft256_full(Y):
    A = Y.astype(complex).copy()
    A[0] = Y[0].real
    A[128] = Y[128].real
    z = half_irfft(A)
    y = np.zeros(256)
    y[0::2] = z.real / 128
    y[1::2] = z.imag / 128
    return y


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    z = rng.normal(size=128) + 1j * rng.normal(size=128)
    print("fft128  err", np.abs(fft128_warp(z) - np.fft.fft(z)).max())
    print("ifft128 err", np.abs(fft128_warp(z, True) - np.fft.ifft(z) * 128).max())
    x = rng.normal(size=256)
    print("rfft256 err", np.abs(rfft256_via128(x) - np.fft.rfft(x, 256)).max())
    Y = rng.normal(size=129) + 1j * rng.normal(size=129)
    print("irfft512[:256] err", np.abs(irfft512_first256(Y) - np.fft.irfft(Y, 512)[:256]).max())
    print("irfft256 err", np.abs(irfft256_full(Y) - np.fft.irfft(Y, 256)).max())
