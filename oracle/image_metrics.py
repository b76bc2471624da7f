"""TEST INFRASTRUCTURE (oracle): CPU restatement of the image metric of the reference's generate().

    trainer.py:514-526, tester.py:236-241
        G_gray = rgb2gray(G.clip(0, 255).astype(np.uint8)); x_gray = rgb2gray(((x+1)*127.5).clip(0, 255).astype(np.uint8))
        ssim(G_gray, x_gray, data_range=x_gray.max() - x_gray.min(), multichannel=False)

`ssim` / `rgb2gray` come from scikit-image (trainer.py:15-16), an un-vendored dependency without a pinned version that
is not installed here.  Restated from its published default path (skimage.measure.compare_ssim, later
skimage.metrics.structural_similarity; Wang, Bovik, Sheikh, Simoncelli 2004): 7x7 uniform window
(scipy.ndimage.uniform_filter), K1 = 0.01, K2 = 0.03, sample covariance (N/(N-1), N = 49), float64, mean over the SSIM
map cropped by (7-1)//2 pixels; rgb2gray = 0.2125 R + 0.7154 G + 0.0721 B on the image scaled to [0,1] (img_as_float).
Parity with a real scikit-image run is unpinned; tests pin this file against a brute-force window loop and against the
closed-form cases (identical images -> 1, constant shift)."""
import numpy as np
from scipy.ndimage import uniform_filter


def rgb2gray_u8(img):
    """uint8 [..., 3] -> float64 [...] in [0,1] (skimage.color.rgb2gray after img_as_float)."""
    f = np.asarray(img, dtype=np.uint8).astype(np.float64) / 255.0
    return f[..., 0] * 0.2125 + f[..., 1] * 0.7154 + f[..., 2] * 0.0721


def ssim_gray(X, Y, data_range, win_size=7, K1=0.01, K2=0.03):
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1.0)
    f = lambda a: uniform_filter(a, size=win_size)
    ux, uy = f(X), f(Y)
    uxx, uyy, uxy = f(X * X), f(Y * Y), f(X * Y)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean())


def ssim_generate(G_u8, x_u8):
    """Per-sample SSIM of generate(): uint8 [n,h,w,3] pairs -> float64 [n]."""
    out = []
    for g, x in zip(G_u8, x_u8):
        gg, xg = rgb2gray_u8(g), rgb2gray_u8(x)
        out.append(ssim_gray(gg, xg, data_range=xg.max() - xg.min()))
    return np.asarray(out)


def ssim_gray_bruteforce(X, Y, data_range):
    """The same quantity with explicit 7x7 window loops (independent of scipy's filter); small images only."""
    X = np.asarray(X, np.float64)
    Y = np.asarray(Y, np.float64)
    H, W = X.shape
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    tot = 0.0
    for y in range(H - 6):
        for x in range(W - 6):
            a, b = X[y:y + 7, x:x + 7].ravel(), Y[y:y + 7, x:x + 7].ravel()
            ux, uy = a.mean(), b.mean()
            vx, vy = a.var(ddof=1), b.var(ddof=1)
            vxy = ((a - ux) * (b - uy)).sum() / 48.0
            tot += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return tot / ((H - 6) * (W - 6))
