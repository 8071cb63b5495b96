"""Deterministic synthetic test images (numpy only; no reference data needed on the GPU box)."""
import numpy as np


def blob_image(w, h, n_blobs=None, seed=0, noise=4.0):
    """Smooth blobs of random size / anisotropy / contrast on a grey canvas + mild noise, u8-valued f32."""
    rng = np.random.default_rng(seed)
    n_blobs = n_blobs or max(20, int(5e-3 * w * h))
    img = np.full((h, w), 128.0)
    for _ in range(n_blobs):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        s1 = np.exp(rng.uniform(np.log(1.5), np.log(12.0)))
        s2 = s1 / rng.uniform(1.0, 3.0)
        th = rng.uniform(0, np.pi)
        amp = rng.uniform(25, 90) * rng.choice([-1.0, 1.0])
        r = int(4 * s1) + 1
        x0, x1, y0, y1 = max(0, int(cx) - r), min(w, int(cx) + r + 1), max(0, int(cy) - r), min(h, int(cy) + r + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dx, dy = xx - cx, yy - cy
        u = np.cos(th) * dx + np.sin(th) * dy
        v = -np.sin(th) * dx + np.cos(th) * dy
        img[y0:y1, x0:x1] += amp * np.exp(-0.5 * ((u / s1) ** 2 + (v / s2) ** 2))
    img += rng.uniform(-noise, noise, size=img.shape)
    return np.floor(np.clip(img, 0, 255) + 0.5).clip(0, 255).astype(np.float32)


def warp_image(img, H, seed=1, noise=2.0):
    """img resampled through H (maps source -> destination), bilinear, border 128, + gaussian noise."""
    h, w = img.shape
    Hi = np.linalg.inv(H)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    d = Hi[2, 0] * xx + Hi[2, 1] * yy + Hi[2, 2]
    sx = (Hi[0, 0] * xx + Hi[0, 1] * yy + Hi[0, 2]) / d
    sy = (Hi[1, 0] * xx + Hi[1, 1] * yy + Hi[1, 2]) / d
    x0 = np.floor(sx).astype(int); y0 = np.floor(sy).astype(int)
    fx, fy = sx - x0, sy - y0
    ok = (x0 >= 0) & (y0 >= 0) & (x0 < w - 1) & (y0 < h - 1)
    x0c, y0c = np.clip(x0, 0, w - 2), np.clip(y0, 0, h - 2)
    a = img.astype(np.float64)
    v = (a[y0c, x0c] * (1 - fx) * (1 - fy) + a[y0c, x0c + 1] * fx * (1 - fy) + a[y0c + 1, x0c] * (1 - fx) * fy + a[y0c + 1, x0c + 1] * fx * fy)
    out = np.where(ok, v, 128.0)
    out += np.random.default_rng(seed).normal(0, noise, size=out.shape)
    return np.floor(np.clip(out, 0, 255) + 0.5).clip(0, 255).astype(np.float32)


def gt_homography(w, h):
    """C * T(37,-21) * R(12 deg) * S(0.93) * P(2e-5, -1e-5) * C^-1."""
    C_ = np.array([[1, 0, w / 2.0], [0, 1, h / 2.0], [0, 0, 1]])
    T = np.array([[1, 0, 37.0], [0, 1, -21.0], [0, 0, 1]])
    a = np.deg2rad(12.0)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    S = np.diag([0.93, 0.93, 1.0])
    P = np.array([[1, 0, 0], [0, 1, 0], [2e-5, -1e-5, 1]])
    return C_ @ T @ R @ S @ P @ np.linalg.inv(C_)


def random_descriptors(n, seed=0, dup_of=None, dup_frac=0.5, jitter=6):
    """RootSIFT-like u8 descriptors; a fraction are noisy copies of rows of `dup_of` (=> real matches)."""
    rng = np.random.default_rng(seed)
    g = rng.gamma(0.3, 1.0, size=(n, 128))
    g = g / g.sum(1, keepdims=True)
    d = np.clip(np.floor(512.0 * np.sqrt(g) + 0.5), 0, 255)
    src = np.full(n, -1)
    if dup_of is not None and len(dup_of):
        k = int(n * dup_frac)
        rows = rng.choice(n, k, replace=False)
        src_rows = rng.integers(0, len(dup_of), k)
        d[rows] = np.clip(dup_of[src_rows].astype(np.int64) + rng.integers(-jitter, jitter + 1, size=(k, 128)), 0, 255)
        src[rows] = src_rows
    return d.astype(np.uint8), src
