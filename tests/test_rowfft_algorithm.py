"""The row-FFT hybrid algorithm of csrc/rltv_stencil_fft.cuh, restated in numpy (float64) and checked on CPU against
the plain definitions in oracle/rl_mm_oracle.py.

This pins the MATH the CUDA kernels implement -- segment geometry (128-sample rows, 112 / 96 valid outputs, 16-byte
aligned box start P4), two real rows per complex FFT, tap spectra, the direct vertical MAC over spectra, and for the
PSF gradient the frequency-domain accumulation, the untangling A[k] = (C[k] + conj(C[-k]))/2 and the K-lag inverse
DFT -- independently of any GPU.  The kernels themselves are tested against the same definitions in test_gpu_parity.py.
"""
import numpy as np
import pytest

from oracle import rl_mm_oracle as orc

FFT_N = 128


def cfg(K):
    """Mirror of FftCfg<K> / GradkFftCfg<K>."""
    P = K // 2
    return dict(P=P, P4=(P + 3) & ~3, TWO=112 if K <= 17 else 96, HB=24, GHB=40 if K <= 17 else 32)


def tap_spectra(w):
    """k_psf_spectrum: Wc[ky][k] = 1/128 sum_kx w[ky][kx] exp(+2 pi i (kx - P) k / 128)."""
    K = w.shape[0]
    P = K // 2
    k = np.arange(FFT_N)
    ph = np.exp(2j * np.pi * np.outer(np.arange(K) - P, k) / FFT_N)        # [kx][k]
    return (w @ ph) / FFT_N                                                  # [ky][k]


def box(a, y0, x0, h, w):
    """TMA box: rows [y0, y0+h), columns [x0, x0+w) of `a`, zero outside."""
    out = np.zeros((h, w))
    ys, xs = max(y0, 0), max(x0, 0)
    ye, xe = min(y0 + h, a.shape[0]), min(x0 + w, a.shape[1])
    if ye > ys and xe > xs:
        out[ys - y0:ye - y0, xs - x0:xe - x0] = a[ys:ye, xs:xe]
    return out


def conv_rowfft(inp, w, K):
    """k_conv_fft: out[Y][X] = sum w[ky][kx] inp[Y-P+ky][X-P+kx] (zero outside), all of inp's domain."""
    c = cfg(K)
    P, P4, TWO, HB = c["P"], c["P4"], c["TWO"], c["HB"]
    H, W = inp.shape
    Wc = tap_spectra(w)
    out = np.zeros((H, W))
    for Y0 in range(0, H, 2 * HB):
        for X0 in range(0, W, TWO):
            tile = box(inp, Y0 - P, X0 - P4, 2 * HB + K - 1, FFT_N)
            Z = np.fft.fft(tile[:HB + K - 1] + 1j * tile[HB:2 * HB + K - 1], axis=1)       # two real rows per FFT
            O = np.stack([sum(Wc[ky] * Z[y + ky] for ky in range(K)) for y in range(HB)])
            o = np.fft.ifft(O, axis=1) * FFT_N                                             # unnormalised inverse
            blk = np.concatenate([o.real, o.imag])[:, P4:P4 + TWO]                          # rows y and y + HB
            h, wd = min(2 * HB, H - Y0), min(TWO, W - X0)
            out[Y0:Y0 + h, X0:X0 + wd] = blk[:h, :wd]
    return out


def gradk_rowfft(u, e_pad, K):
    """k_gradk_fft + k_gradk_fft_finish: gk'[dy][dx] = sum_{Y,X} e[Y][X] u[Y-P+dy][X-P+dx] on the padded domain."""
    c = cfg(K)
    P, P4, TWO, HB = c["P"], c["P4"], c["TWO"], c["GHB"]
    H, W = u.shape
    C = np.zeros((K, FFT_N), complex)
    for Y0 in range(0, H, 2 * HB):
        for X0 in range(0, W, TWO):
            ut = box(u, Y0 - P, X0 - P4, 2 * HB + K - 1, FFT_N)
            et = box(e_pad, Y0, X0 - P4, 2 * HB, FFT_N)
            et[:, :P4] = 0.0
            et[:, P4 + TWO:] = 0.0                                                          # only the valid columns
            Zu = np.fft.fft(ut[:HB + K - 1] + 1j * ut[HB:2 * HB + K - 1], axis=1)
            Ze = np.fft.fft(et[:HB] + 1j * et[HB:], axis=1)
            for dy in range(K):
                C[dy] += (np.conj(Ze) * Zu[dy:dy + HB]).sum(axis=0)
    A = 0.5 * (C + np.conj(C[:, (-np.arange(FFT_N)) % FFT_N]))                             # drop the packed cross term
    k = np.arange(FFT_N)
    lag = np.exp(2j * np.pi * np.outer(k, np.arange(K) - P) / FFT_N)                        # [k][dx]
    return (A @ lag).real / FFT_N


@pytest.mark.parametrize("K,M,N", [(9, 60, 130), (15, 101, 240), (17, 49, 113), (25, 70, 100), (31, 97, 97)])
def test_row_fft_hybrid_equals_the_definitions(K, M, N):
    rng = np.random.default_rng(K)
    P = K // 2
    u = rng.random((M + K - 1, N + K - 1))
    psf = rng.random((K, K))
    psf /= psf.sum()
    image = rng.random((M, N))
    # forward blur (true convolution = correlation with rot180(psf)), residual on the image, zero on the ring
    blur = conv_rowfft(u, orc.rot180(psf), K)[P:P + M, P:P + N]
    e_ref = orc.conv2(u, psf, "valid") - image
    assert np.allclose(blur - image, e_ref, atol=1e-12)
    e_pad = np.zeros_like(u)
    e_pad[P:P + M, P:P + N] = e_ref
    # adjoint: full convolution with rot180(psf) = centred correlation with psf on the padded domain
    g = conv_rowfft(e_pad, psf, K)
    assert np.allclose(g, orc.conv2(e_ref, orc.rot180(psf), "full"), atol=1e-12)
    # PSF gradient, then the K-1-q flip of k_psf_update / rltv_stage_gradk
    gkp = gradk_rowfft(u, e_pad, K)
    gk_ref = orc.conv2(orc.rot180(u), e_ref, "valid")
    assert np.allclose(gkp[::-1, ::-1], gk_ref, rtol=1e-10, atol=1e-9)


def gradk_fused_rowfft(u_band, img_band, w_fwd, K, own0, own1, row0, M, N):
    """k_gradk_fft<K, FUSED> on one row band: the residual is formed per tile from the u spectra the kernel already
    holds (vertical MAC with the forward tap spectra, inverse FFT, minus image, mask), then accumulated as above.
    u_band / img_band: the band's local rows of the padded arrays (image at columns P.., zero ring); [own0, own1) the
    owned local rows; row0 the band's first row in the frame.  Returns (gk' partial sums, residual on the owned rows)."""
    c = cfg(K)
    P, P4, TWO, HB = c["P"], c["P4"], c["TWO"], c["GHB"]
    W = u_band.shape[1]
    nown = own1 - own0
    Wc = tap_spectra(w_fwd)
    C = np.zeros((K, FFT_N), complex)
    err = np.zeros((nown, W))
    n = np.arange(FFT_N)
    for Y0 in range(0, nown, 2 * HB):
        for X0 in range(0, W, TWO):
            ut = box(u_band, own0 + Y0 - P, X0 - P4, 2 * HB + K - 1, FFT_N)
            it = box(img_band[own0:own1], Y0, X0 - P4, 2 * HB, FFT_N)
            Zu = np.fft.fft(ut[:HB + K - 1] + 1j * ut[HB:2 * HB + K - 1], axis=1)
            O = np.stack([sum(Wc[ky] * Zu[y + ky] for ky in range(K)) for y in range(HB)])
            o = np.fft.ifft(O, axis=1) * FFT_N
            blur = np.concatenate([o.real, o.imag])                                          # [2 HB][128]
            X = X0 - P4 + n
            colin = (n >= P4) & (n < P4 + TWO) & (X >= P) & (X < P + N)
            rl = Y0 + np.arange(2 * HB)
            gy = row0 + own0 + rl
            rowin = (rl < nown) & (gy >= P) & (gy < P + M)
            e = np.where(rowin[:, None] & colin[None, :], blur - it, 0.0)
            valid = (n >= P4) & (n < P4 + TWO) & (X < W)
            for r in np.nonzero(rl < nown)[0]:
                err[rl[r], X[valid]] = e[r, valid]
            Ze = np.fft.fft(e[:HB] + 1j * e[HB:], axis=1)
            for dy in range(K):
                C[dy] += (np.conj(Ze) * Zu[dy:dy + HB]).sum(axis=0)
    A = 0.5 * (C + np.conj(C[:, (-n) % FFT_N]))
    lag = np.exp(2j * np.pi * np.outer(n, np.arange(K) - P) / FFT_N)
    return (A @ lag).real / FFT_N, err


@pytest.mark.parametrize("K,M,N,world", [(15, 130, 150, 1), (15, 130, 150, 3), (11, 200, 90, 4), (17, 90, 230, 2)])
def test_fused_residual_and_psf_gradient_per_band(K, M, N, world):
    """Whole frame and row bands (plan_bands: 2P halo rows, owned rows tile the frame): the bands' partial sums add
    up to the PSF gradient of the definition, and each band leaves the residual of exactly its owned rows."""
    from image_cases_studies_b200.distributed import plan_bands
    rng = np.random.default_rng(10 * K + world)
    P = K // 2
    u = rng.random((M + K - 1, N + K - 1))
    psf = rng.random((K, K))
    psf /= psf.sum()
    image = rng.random((M, N))
    img_pad = np.zeros_like(u)
    img_pad[P:P + M, P:P + N] = image
    e_ref = orc.conv2(u, psf, "valid") - image
    e_pad = np.zeros_like(u)
    e_pad[P:P + M, P:P + N] = e_ref
    bands, _ = plan_bands(M, K, world, None)
    total = np.zeros((K, K))
    for lo, hi, olo, ohi in bands:
        gkp, err = gradk_fused_rowfft(u[lo:hi], img_pad[lo:hi], orc.rot180(psf), K, olo - lo, ohi - lo, lo, M, N)
        assert np.allclose(err, e_pad[olo:ohi], atol=1e-12)
        total += gkp
    gk_ref = orc.conv2(orc.rot180(u), e_ref, "valid")
    assert np.allclose(total[::-1, ::-1], gk_ref, rtol=1e-10, atol=1e-9)
