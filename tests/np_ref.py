"""Independent NumPy restatement of the leaf functions of the path, written from the reference sources (not from
the oracle) so that the oracle is not self-certifying."""
import numpy as np


def pix_vals(img, pts, overflow=128.0):
    """utils::getPixVal<Linear, Constant> (Utilities/include/mtf/Utilities/imgUtils.h:91-113), vectorised"""
    h, w = img.shape
    x, y = pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64)
    out = np.full(len(pts), overflow)
    inb = ~((x < 0) | (x >= w) | (y < 0) | (y >= h))
    xs, ys = np.where(inb, x, 0.0), np.where(inb, y, 0.0)
    lx, ly = xs.astype(np.int64), ys.astype(np.int64)            # (int) truncation; coordinates are >= 0 here
    dx, dy = xs - lx, ys - ly
    ux, uy = np.where(dx == 0, lx, lx + 1), np.where(dy == 0, ly, ly + 1)
    ok = inb & (ux < w) & (uy < h)
    ux, uy = np.minimum(ux, w - 1), np.minimum(uy, h - 1)
    im = img.astype(np.float64)
    v = im[ly, lx] * (1 - dx) * (1 - dy) + im[ly, ux] * dx * (1 - dy) + im[uy, lx] * (1 - dx) * dy + im[uy, ux] * dx * dy
    out[ok] = v[ok]
    return out


def img_grad(img, pts, eps=1e-8):
    """utils::getImgGrad (Utilities/src/imgUtils.cc:233-254): central finite differences of the interpolant"""
    e = np.array([eps, 0.0]); f = np.array([0.0, eps])
    gx = (pix_vals(img, pts + e) - pix_vals(img, pts - e)) * (1.0 / (2 * eps))
    gy = (pix_vals(img, pts + f) - pix_vals(img, pts - f)) * (1.0 / (2 * eps))
    return np.stack([gx, gy], 1)


def homography_dlt(src, dst):
    """utils::computeHomographyDLT (Utilities/src/warpUtils.cc:171-223): null vector of the 8 x 9 system"""
    A = np.zeros((8, 9))
    for i in range(4):
        x, y, u, v = src[0, i], src[1, i], dst[0, i], dst[1, i]
        A[2 * i] = [0, 0, 0, -x, -y, -1, v * x, v * y, v]
        A[2 * i + 1] = [x, y, 1, 0, 0, 0, -u * x, -u * y, -u]
    h = np.linalg.svd(A)[2][-1]
    return (h / h[8]).reshape(3, 3)


def warp_from_state(ssm, s):
    if ssm == "homography":                                       # SSM/src/Homography.cc:94-107
        return np.array([[1 + s[0], s[1], s[2]], [s[3], 1 + s[4], s[5]], [s[6], s[7], 1.0]])
    return np.array([[1 + s[2], s[3], s[0]], [s[4], 1 + s[5], s[1]], [0, 0, 1.0]])   # SSM/src/Affine.cc:117-131


def ssd(I0, It):
    return -0.5 * np.sum((It - I0) ** 2)                          # AM/src/SSDBase.cc:94


def ncc(I0, It):
    a, b = I0 - I0.mean(), It - It.mean()                         # AM/src/NCC.cc:139-152
    return (a @ b) / (np.linalg.norm(a) * np.linalg.norm(b))
