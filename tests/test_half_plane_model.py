"""CPU model (numpy, float64) of the NEXT step for odd colour planes: a lone real
plane restored through HALF a complex plane instead of a half-empty one.  Not product code -- it pins the index
algebra the CUDA passes will have to follow, against the direct pipeline of SURVEY.md Appendix A:

  pass 1  rows y and y' = y + Rl/2 of a rank's slab share one complex row transform; their Hermitian spectra are
          untangled; columns 1 .. Cp/2-1 go to a half-width plane, the two REAL columns 0 and Cp/2 to a 2-column
          side plane;
  pass 2  ordinary column transform -> Wiener factor -> inverse on the half plane (left half of Wf) and on the
          side plane (columns 0 and Cp/2 of Wf);
  pass 3  the row spectrum of the restored (real) rows is Hermitian again: the upper half is rebuilt by reading
          the lower half in reverse order and conjugating; rows y and y' are packed into one inverse transform.
"""
import numpy as np
import pytest


def wiener_factor(Rp, Cp, rng, K=0.01):
    psf = np.zeros((Rp, Cp))
    psf[:5, :7] = rng.random((5, 7)) / 10  # real kernel -> Hermitian spectrum
    H = np.fft.fft2(psf)
    return np.conj(H) / (np.abs(H) ** 2 + K)


def direct(g, wf):
    """Unscaled inverse of the filtered spectrum (fft_serial.cpp:176-229 semantics)."""
    return np.real(np.fft.ifft2(np.fft.fft2(g) * wf)) * g.size


def half_plane(g, wf, world=1):
    Rp, Cp = g.shape
    Rl, h = Rp // world, Cp // 2
    assert Rl % 2 == 0
    HP = np.zeros((Rp, h), np.complex128)   # columns 1 .. h-1 used
    SP = np.zeros((Rp, 2), np.complex128)   # columns 0 and Cp/2 (real numbers)
    k = np.arange(h + 1)
    for rank in range(world):               # pass 1: row pairs stay inside a rank's row slab
        for y in range(rank * Rl, rank * Rl + Rl // 2):
            y2 = y + Rl // 2
            Z = np.fft.fft(g[y] + 1j * g[y2])
            Zm = np.conj(Z[(-k) % Cp])       # conj(Z[N - k]): the one extra shared-memory exchange of pass 1
            Xa, Xb = (Z[k] + Zm) / 2, (Z[k] - Zm) / 2j
            for r, X in ((y, Xa), (y2, Xb)):
                HP[r, 1:h] = X[1:h]
                assert abs(X[0].imag) < 1e-9 and abs(X[h].imag) < 1e-9
                SP[r, 0], SP[r, 1] = X[0].real, X[h].real
    # pass 2: columns (any column kernel: a column never meets its mirror)
    HP2 = np.fft.ifft(np.fft.fft(HP, axis=0) * wf[:, :h], axis=0) * Rp
    SP2 = np.fft.ifft(np.fft.fft(SP, axis=0) * wf[:, [0, h]], axis=0) * Rp
    assert np.abs(SP2.imag).max() < 1e-9 * max(1.0, np.abs(SP2).max())  # columns 0 and Cp/2 stay real
    out = np.zeros((Rp, Cp))
    for rank in range(world):               # pass 3
        for y in range(rank * Rl, rank * Rl + Rl // 2):
            y2 = y + Rl // 2
            Z = np.zeros(Cp, np.complex128)
            for r, w in ((y, 1.0), (y2, 1j)):
                X = np.zeros(Cp, np.complex128)
                X[0], X[h] = SP2[r, 0].real, SP2[r, 1].real
                X[1:h] = HP2[r, 1:h]
                X[h + 1:] = np.conj(HP2[r, 1:h][::-1])  # reverse-order load + conjugate, no exchange
                Z += w * X
            z = np.fft.ifft(Z) * Cp
            out[y], out[y2] = z.real, z.imag
    return out


@pytest.mark.parametrize("Rp,Cp,world", [(8, 8, 1), (16, 32, 1), (64, 16, 2), (32, 64, 8)])
def test_half_plane_scheme_equals_direct(Rp, Cp, world):
    rng = np.random.default_rng(Rp * 131 + Cp)
    g = rng.random((Rp, Cp))
    wf = wiener_factor(Rp, Cp, rng)
    want = direct(g, wf)
    got = half_plane(g, wf, world)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
