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


# ---------------------------------------------------------------------------------------------
# Thread-level model of the kernels as built (csrc/passes_impl.cuh, ROW_OUT_HALF / ROW_IN_HALF): registers v[m] of
# thread t hold points t + T*m (T = N/16), the mirror exchange goes through shared-memory words exactly as indexed in
# the kernel, the column pass leaves conj(IFFT) behind and the inverse row pass is a FORWARD transform of conj(Z).
# ---------------------------------------------------------------------------------------------
E, H8 = 16, 8


def kernel_untangle(Z):
    """ROW_OUT_HALF after fft_forward: Z[N] -> (X1[k], X2[k]) for k < N/2 and the two Nyquist reals."""
    N = Z.size
    T = N // E
    v = Z.reshape(E, T).T.copy()                      # v[t][m] = Z[t + T*m]
    ex = np.full(H8 * T + 1, np.nan + 0j)
    for t in range(T):
        for j in range(H8):
            ex[j * T + t] = v[t, H8 + j]
    ex[H8 * T] = v[0, 0]
    X1 = np.zeros(N // 2, complex)
    X2 = np.zeros(N // 2, complex)
    for t in range(T):
        for m in range(H8):
            z, zm = v[t, m], ex[(H8 - m) * T - t]
            k = t + T * m
            X1[k] = complex(0.5 * (z.real + zm.real), 0.5 * (z.imag - zm.imag))
            X2[k] = complex(0.5 * (z.imag + zm.imag), 0.5 * (zm.real - z.real))
    return X1, X2, v[0, H8].real, v[0, H8].imag


def kernel_tangle(A, B, nyqA, nyqB):
    """ROW_IN_HALF: stored half rows A = conj(Y_y[k]), B = conj(Y_{y+D}[k]) (k < N/2) + Nyquist entries -> the N inputs of
    the forward transform, conj(Z) with Z = Y_y + i Y_{y+D}."""
    N = 2 * A.size
    T = N // E
    v = np.full((T, E), np.nan + 0j)
    ex = np.full(H8 * T + 1, np.nan + 0j)
    for t in range(T):
        for m in range(H8):
            k = t + T * m
            a, b = A[k], B[k]
            v[t, m] = complex(a.real + b.imag, a.imag - b.real)
            ex[(H8 - m) * T - t] = complex(a.real - b.imag, -(a.imag + b.real))
    for t in range(T):
        for j in range(H8):
            v[t, H8 + j] = ex[j * T + t]
    v[0, H8] = complex(nyqA.real + nyqB.imag, nyqA.imag - nyqB.real)
    assert not np.isnan(v).any()
    return v.T.reshape(N)                              # point t + T*m


@pytest.mark.parametrize("N", [64, 128, 512, 2048])
def test_kernel_index_algebra_untangle_tangle(N):
    rng = np.random.default_rng(N)
    x1, x2 = rng.random(N), rng.random(N)
    X1, X2, n1, n2 = kernel_untangle(np.fft.fft(x1 + 1j * x2))
    F1, F2 = np.fft.fft(x1), np.fft.fft(x2)
    assert np.abs(X1 - F1[: N // 2]).max() < 1e-9 and np.abs(X2 - F2[: N // 2]).max() < 1e-9
    assert abs(n1 - F1[N // 2].real) < 1e-9 and abs(n2 - F2[N // 2].real) < 1e-9
    # inverse side: any Hermitian row spectra Y1, Y2 (of real rows r1, r2), stored conjugated
    r1, r2 = rng.random(N), rng.random(N)
    Y1, Y2 = np.fft.fft(r1), np.fft.fft(r2)
    vin = kernel_tangle(np.conj(Y1[: N // 2]), np.conj(Y2[: N // 2]), np.conj(Y1[N // 2]), np.conj(Y2[N // 2]))
    out = np.fft.fft(vin)                              # the kernel's forward transform
    assert np.abs(out.real - N * r1).max() < 1e-8 * N  # plane row y   = v.x
    assert np.abs(-out.imag - N * r2).max() < 1e-8 * N  # plane row y+D = -v.y


@pytest.mark.parametrize("H,W,world", [(64, 64, 1), (50, 70, 1), (37, 128, 2), (128, 256, 4), (5, 64, 1), (9, 64, 2)])
def test_kernel_half_plane_pipeline(H, W, world):
    """Whole lone-plane path with the kernels' conventions: pass-1 pairing D1 = ceil(rows_local/2) inside a rank, rows >= H
    never stored, column pass = conj(IFFT(FFT * Wf)), pass-3 pairing D3 = Rl/2; against the direct pipeline."""
    rng = np.random.default_rng(H * 7 + W)
    Rp = 1 << int(np.ceil(np.log2(max(H, 2))))
    Cp = 1 << int(np.ceil(np.log2(W)))
    g = np.zeros((Rp, Cp))
    g[:H, :W] = rng.random((H, W))
    wf = wiener_factor(Rp, Cp, rng)
    want = direct(g, wf)
    Rl, h = Rp // world, Cp // 2
    HP = np.zeros((Rp, h), complex)
    NQ = np.zeros(Rp, complex)
    for rank in range(world):
        row0 = rank * Rl
        rows_local = max(0, min(row0 + Rl, H) - row0)
        D1 = (rows_local + 1) // 2
        for r in range(D1):
            second = g[row0 + r + D1] if r + D1 < rows_local else np.zeros(Cp)
            X1, X2, n1, n2 = kernel_untangle(np.fft.fft(g[row0 + r] + 1j * second))
            HP[row0 + r], NQ[row0 + r] = X1, n1
            if row0 + r + D1 < H:
                HP[row0 + r + D1], NQ[row0 + r + D1] = X2, n2
    HP2 = np.conj(np.fft.ifft(np.fft.fft(HP, axis=0) * wf[:, :h], axis=0) * Rp)
    NQ2 = np.conj(np.fft.ifft(np.fft.fft(NQ) * wf[:, h]) * Rp)
    out = np.zeros((Rp, Cp))
    D3 = Rl // 2
    for rank in range(world):
        for r in range(D3):
            y, y2 = rank * Rl + r, rank * Rl + r + D3
            o = np.fft.fft(kernel_tangle(HP2[y], HP2[y2], NQ2[y], NQ2[y2]))
            out[y], out[y2] = o.real, -o.imag
    assert np.abs(out - want).max() < 1e-9 * np.abs(want).max()
