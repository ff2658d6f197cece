"""Thread-level numpy model of the 32-points-per-thread long-row core (csrc/fft_mid.cuh): the same register ownership
(thread t holds points t + T*m), the same three Stockham stages (32.32.16 for 16384, 32.16.16 for 8192), the same shared-memory
word addresses (word(idx) = idx + (idx >> 5)) and the same twiddle tables (stage 2: [r-1][t mod 32]; stage 3: butterfly 0 only,
butterfly b derived by the root W_32^{r b}).  Pins the index algebra on the CPU: result against numpy's FFT, every exchange a
bijection, and no shared-memory bank conflicts for the 64-bit accesses of a half warp."""
import numpy as np
import pytest

E = 32


def skew(idx):
    return idx + (idx >> 5)


def bank_conflict_free(words):
    """words: (threads, accesses) float2 word addresses, one column per instruction.  64-bit accesses are served per half warp:
    the 16 lanes of a half warp must hit 16 different 8-byte bank pairs (word mod 16)."""
    T = words.shape[0]
    for col in range(words.shape[1]):
        w = words[:, col].reshape(T // 16, 16) % 16
        if not all(len(set(row)) == 16 for row in w):
            return False
    return True


def mid_fft(x):
    N = x.size
    T = N // E
    R2 = 32 if N == 16384 else 16
    R3 = 16
    NB2, NB3 = E // R2, E // R3
    t = np.arange(T)
    v = x.reshape(E, T).T.copy()                     # v[t, m] = x[t + T m]
    ex = np.full(skew(N), np.nan + 0j)
    # stage 1: radix 32 over the thread's own points, natural order out
    v = np.fft.fft(v, axis=1)
    w = 33 * t[:, None] + np.arange(E)[None, :]      # outputs 32 t + q -> word 33 t + q
    assert np.array_equal(w, skew(32 * t[:, None] + np.arange(E)[None, :]))
    assert bank_conflict_free(w)
    ex[w] = v
    r = skew(t)[:, None] + (T + T // 32) * np.arange(E)[None, :]
    assert np.array_equal(r, skew(t[:, None] + T * np.arange(E)[None, :]))
    assert bank_conflict_free(r)
    assert len(np.unique(w)) == N and set(np.unique(r)) == set(np.unique(w))
    v = ex[r]
    # stage 2: radix R2, sub-length 32, twiddle exp(-2 pi i r k / (32 R2)), k = t mod 32 for every butterfly
    tw2 = np.exp(-2j * np.pi * np.arange(R2)[None, :] * (t & 31)[:, None] / (32 * R2))
    ex[:] = np.nan
    allw = []
    for b in range(NB2):
        xb = v[:, b + NB2 * np.arange(R2)] * tw2
        xb = np.fft.fft(xb, axis=1)
        v[:, b + NB2 * np.arange(R2)] = xb
        j = t + b * T
        base = (j >> 5) * (32 * R2) + (j & 31)
        w = skew(base)[:, None] + 33 * np.arange(R2)[None, :]
        assert np.array_equal(w, skew(base[:, None] + 32 * np.arange(R2)[None, :]))
        assert bank_conflict_free(w)
        ex[w] = xb
        allw.append(w)
    assert len(np.unique(np.concatenate(allw))) == N
    v = ex[r]
    assert not np.isnan(v).any()
    # stage 3: radix 16, sub-length 32 R2, k = t + b T; table of butterfly 0 times the root W_32^{r b}
    tw3 = np.exp(-2j * np.pi * np.arange(R3)[None, :] * t[:, None] / N)
    for b in range(NB3):
        root = np.exp(-2j * np.pi * np.arange(R3) * b / 32)
        xb = v[:, b + NB3 * np.arange(R3)] * root[None, :] * tw3
        v[:, b + NB3 * np.arange(R3)] = np.fft.fft(xb, axis=1)
    return v.T.reshape(N)                            # X[t + T m] = v[t, m]


@pytest.mark.parametrize("N", [8192, 16384])
def test_mid_core_model_equals_fft(N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    got = mid_fft(x)
    want = np.fft.fft(x)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()


def test_mid_core_exchange_fits_the_buffer():
    # N + N/32 words hold every address the kernel touches; the half-plane forms additionally park N/2 + 1 values in it
    for N in (8192, 16384):
        assert skew(N - 1) < N + N // 32
        assert N // 2 + 1 <= N + N // 32
