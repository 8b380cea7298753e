"""Seeded synthetic image sequences and patch layouts.

Modelled on the reference's Examples/cpp/generateSyntheticSeq.cc:251-400 (one frame warped by a
random-walk SSM) and on the pre-processing MTF applies before the AM sees a frame
(5x5 Gaussian, sigma 3, on the CV_32FC1 image: Utilities/include/mtf/Utilities/preprocUtils.h:67-78,
Config/include/mtf/Config/parameters.h:229-235).  Pure NumPy, host side, not part of the timed path.
"""
import numpy as np


def gaussian_blur5(img, sigma=3.0):
    """Separable 5-tap Gaussian with reflect-101 borders (cv::GaussianBlur(5x5, sigma) equivalent)."""
    k = np.exp(-0.5 * (np.arange(-2, 3) / sigma) ** 2)
    k /= k.sum()
    out = img.astype(np.float64)
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (2, 2)
        p = np.pad(out, pad, mode="reflect")
        acc = np.zeros_like(out)
        for i in range(5):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(i, i + out.shape[axis])
            acc += k[i] * p[tuple(sl)]
        out = acc
    return out


def make_frame0(h=1024, w=1024, seed=1234, n_waves=64):
    """Sum of random-phase sinusoids + smoothed white noise, blurred, scaled to [0, 255], float32."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w))
    for _ in range(n_waves):
        wavelength = np.exp(rng.uniform(np.log(12.0), np.log(240.0)))
        theta = rng.uniform(0, 2 * np.pi)
        phase = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(0.3, 1.0) * wavelength ** 0.3
        k = 2 * np.pi / wavelength
        img += amp * np.sin(k * (np.cos(theta) * xx + np.sin(theta) * yy) + phase)
    # piecewise-constant "objects" (soft-edged after the blur) give the histogram-based AMs
    # a multi-modal intensity distribution inside every patch
    step = rng.uniform(-1.0, 1.0, size=(h // 16 + 2, w // 16 + 2))
    step = np.kron(step, np.ones((16, 16)))[:h, :w]
    img += 0.8 * img.std() * step
    noise = rng.standard_normal((h, w))
    for _ in range(3):
        noise = gaussian_blur5(noise, 1.5)
    img += 0.35 * img.std() * noise / noise.std()
    img = gaussian_blur5(img, 3.0)
    img -= img.min()
    img *= 255.0 / img.max()
    return img.astype(np.float32)


def bilinear(img, x, y):
    """fp64 bilinear sampling with clamped coordinates (used only to synthesise frames)."""
    h, w = img.shape
    x = np.clip(x, 0, w - 1.000001)
    y = np.clip(y, 0, h - 1.000001)
    lx = np.floor(x).astype(np.int64)
    ly = np.floor(y).astype(np.int64)
    dx = x - lx
    dy = y - ly
    im = img.astype(np.float64)
    return (im[ly, lx] * (1 - dx) * (1 - dy) + im[ly, lx + 1] * dx * (1 - dy) +
            im[ly + 1, lx] * (1 - dx) * dy + im[ly + 1, lx + 1] * dx * dy)


def dlt4(src, dst):
    """4-point homography src -> dst (2x4 arrays), NumPy SVD."""
    A = []
    for i in range(4):
        x, y = src[0, i], src[1, i]
        u, v = dst[0, i], dst[1, i]
        A.append([0, 0, 0, -x, -y, -1, v * x, v * y, v])
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y, -u])
    _, _, vt = np.linalg.svd(np.asarray(A, dtype=np.float64))
    H = vt[-1].reshape(3, 3)
    return H / H[2, 2]


def warp_frame(img, G):
    """frame_t(p) = frame_0(G^-1 p): the scene content moves by G."""
    h, w = img.shape
    Gi = np.linalg.inv(G)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    d = Gi[2, 0] * xx + Gi[2, 1] * yy + Gi[2, 2]
    sx = (Gi[0, 0] * xx + Gi[0, 1] * yy + Gi[0, 2]) / d
    sy = (Gi[1, 0] * xx + Gi[1, 1] * yy + Gi[1, 2]) / d
    return bilinear(img, sx, sy).astype(np.float32)


def make_sequence(n_frames=8, h=1024, w=1024, seed=1234, walk_seed=5678, sigma=1.0, noise_sigma=0.0):
    """Returns (frames[list of float32 h x w], warps[list of 3x3]); warps[t] maps frame-0 coords to frame t.

    Per-frame global homography random walk: the image corners are jittered by N(0, sigma) px each
    frame (generateSyntheticSeq.cc:318-352 does the same through the SSM sampler)."""
    f0 = make_frame0(h, w, seed)
    rng = np.random.default_rng(walk_seed)
    base = np.array([[0, w - 1, w - 1, 0], [0, 0, h - 1, h - 1]], dtype=np.float64)
    cur = base.copy()
    frames, warps = [f0], [np.eye(3)]
    for _ in range(1, n_frames):
        cur = cur + rng.normal(0, sigma, size=(2, 4))
        G = dlt4(base, cur)
        fr = warp_frame(f0, G)
        if noise_sigma > 0:
            fr = (fr + rng.normal(0, noise_sigma, size=fr.shape)).astype(np.float32)
        frames.append(fr)
        warps.append(G)
    return frames, warps


def make_patches(n_patches, side=49.0, h=1024, w=1024, seed=42, margin=40.0):
    """Axis-aligned square boxes on a jittered lattice, centres >= margin + side/2 from the border.

    side = 49.0 makes the 50x50 grid land on integer pixels (exercises the dx == 0 rule of
    imgUtils.h:103-104); side = 52.3 is the generic sub-pixel case.  Returns (P, 2, 4) corners,
    columns UL, UR, LR, LL (SM/include/mtf/SM/SearchMethod.h:19)."""
    rng = np.random.default_rng(seed)
    g = int(np.ceil(np.sqrt(n_patches)))
    lo = margin + side / 2
    xs = np.linspace(lo, w - 1 - lo, g)
    ys = np.linspace(lo, h - 1 - lo, g)
    step = min(xs[1] - xs[0], ys[1] - ys[0]) if g > 1 else 0.0
    out = np.empty((n_patches, 2, 4))
    integer = float(side).is_integer()
    for i in range(n_patches):
        cx = xs[i % g] + rng.uniform(-0.25, 0.25) * step
        cy = ys[i // g] + rng.uniform(-0.25, 0.25) * step
        if integer:
            cx = np.round(cx - side / 2) + side / 2
            cy = np.round(cy - side / 2) + side / 2
        cx = min(max(cx, lo), w - 1 - lo)
        cy = min(max(cy, lo), h - 1 - lo)
        x0, y0 = cx - side / 2, cy - side / 2
        if integer:
            x0, y0 = np.round(x0), np.round(y0)
        out[i, 0] = [x0, x0 + side, x0 + side, x0]
        out[i, 1] = [y0, y0, y0 + side, y0 + side]
    return out


def warp_corners(G, corners):
    """Apply a 3x3 homography to (..., 2, 4) corners."""
    c = np.asarray(corners, dtype=np.float64)
    x, y = c[..., 0, :], c[..., 1, :]
    d = G[2, 0] * x + G[2, 1] * y + G[2, 2]
    return np.stack([(G[0, 0] * x + G[0, 1] * y + G[0, 2]) / d,
                     (G[1, 0] * x + G[1, 1] * y + G[1, 2]) / d], axis=-2)
