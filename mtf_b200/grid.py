"""GridTracker<SSM> (SM/src/GridTracker.cc) on the batch library: grid_size_x x grid_size_y patch trackers as ONE batch context,
the warp of the whole region estimated from their centroids on the device (mtfb_grid_estimate: RANSAC / LMedS / least squares +
LM refinement in one launch), the few host-side steps of GridTracker::update around it (the 4 corners of the region, the cell
layout) in NumPy.  Host work per frame: 8 doubles down, the P x 8 cell corners up.

Implemented: Homography / Affine as the grid's SSM with its default non-normalised initialisation, reset_at_each_frame 0 / 1,
dyn_patch_size, patch_centroid_inside, forward-backward error estimation (fb_err_thresh > 0, fb_reinit: GridTracker.cc:292-343;
cells in one process).  Not implemented (raise): pyramidal cells, heterogeneous cells."""
import numpy as np

from . import api


def norm_unit_square_pts(resx, resy):
    """utils::getNormUnitSquarePts (warpUtils.cc:15-33): (2, resx*resy) points row by row, (2, 4) corners"""
    xs = np.linspace(-0.5, 0.5, resx) if resx > 1 else np.array([-0.5])
    ys = np.linspace(-0.5, 0.5, resy) if resy > 1 else np.array([-0.5])
    pts = np.empty((2, resx * resy))
    pts[0] = np.tile(xs, resy)
    pts[1] = np.repeat(ys, resx)
    corners = np.array([[-0.5, 0.5, 0.5, -0.5], [-0.5, -0.5, 0.5, 0.5]])
    return pts, corners


def homography_dlt(src, dst):
    """utils::computeHomographyDLT (warpUtils.cc:183-220) for 4 point pairs: the homography with H[2, 2] = 1.  The reference takes
    the null vector of the 8 x 9 system from its SVD and divides by its last entry; the 8 x 8 system solved here has the same
    solution."""
    src = np.asarray(src, dtype=np.float64).reshape(2, 4)
    dst = np.asarray(dst, dtype=np.float64).reshape(2, 4)
    x, y, u, v = src[0], src[1], dst[0], dst[1]
    A = np.zeros((8, 8))
    A[0::2, 0], A[0::2, 1], A[0::2, 2], A[0::2, 6], A[0::2, 7] = x, y, 1.0, -u * x, -u * y
    A[1::2, 3], A[1::2, 4], A[1::2, 5], A[1::2, 6], A[1::2, 7] = x, y, 1.0, -v * x, -v * y
    b = np.empty(8)
    b[0::2], b[1::2] = u, v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


_NORM_CACHE = {}


def pts_from_corners(corners, resx, resy):
    """utils::getPtsFromCorners (warpUtils.cc:34-60): the regular grid of the unit square under the 4-corner homography"""
    key = (resx, resy)
    if key not in _NORM_CACHE:
        pts, nc = norm_unit_square_pts(resx, resy)
        _NORM_CACHE[key] = (np.vstack([pts, np.ones(pts.shape[1])]), nc)
    pts_hm, nc = _NORM_CACHE[key]
    q = homography_dlt(nc, corners) @ pts_hm
    return q[:2] / q[2]


def apply_warp_to_corners(ssm, corners, state_update):
    """ssm.applyWarpToCorners (Homography.cc / Affine.cc:371-380): the warp of a state applied to 4 corners"""
    s = np.asarray(state_update, dtype=np.float64)
    if ssm == "homography":
        W = np.array([[1 + s[0], s[1], s[2]], [s[3], 1 + s[4], s[5]], [s[6], s[7], 1.0]])
    else:
        # Affine::getWarpFromState (Affine.cc:127-141): (tx, ty, a, b, c, d)
        W = np.array([[1 + s[2], s[3], s[0]], [s[4], 1 + s[5], s[1]], [0, 0, 1.0]])
    q = W @ np.vstack([np.asarray(corners, dtype=np.float64).reshape(2, 4), np.ones(4)])
    return q[:2] / q[2]


class GridTracker:
    """mirror of GridTracker<SSM>: setImage / initialize / update / setRegion / getRegion"""

    def __init__(self, cell_params, grid_size_x=10, grid_size_y=10, patch_size_x=10, patch_size_y=10, reset_at_each_frame=1,
                 dyn_patch_size=0, patch_centroid_inside=True, fb_err_thresh=0, fb_reinit=0, enable_pyr=0, ssm="homography",
                 est_params=None, seed=1, cells=None, shard=None, gather=None, upload=None, device_layout=True):
        """cells / shard / gather / upload: the cells split over several processes (one GPU each).  cells = this process's
        BatchTracker of the cells [shard[0], shard[1]); gather() -> a device array (.data_ptr()) of the current corners of ALL
        cells (n x 8) on this device; upload(ndarray) -> the same kind of object for the regions the cells were reset to.  Every
        process then runs the estimation on all the centroids (same seed, same result) and resets its own cells."""
        # defaults: GridTracker.h:8-23
        if fb_err_thresh > 0 and gather is not None:
            raise api.MTFError(2, "GridTracker: forward-backward error estimation is not implemented for cells split over processes")
        if enable_pyr:
            raise api.MTFError(2, "GridTracker: pyramidal patch trackers are not implemented")
        if ssm not in ("homography", "affine"):
            raise api.MTFError(2, "GridTracker: the grid's SSM must be homography or affine")
        self.gx, self.gy = int(grid_size_x), int(grid_size_y)
        n = self.gx * self.gy
        self.shard = (0, n) if shard is None else (int(shard[0]), int(shard[1]))
        if (cells.P if cells is not None else cell_params.n_patches) != self.shard[1] - self.shard[0]:
            # GridTracker.cc:126-131
            raise api.MTFError(1, "GridTracker :: Mismatch between grid dimensions and no. of trackers")
        self.gather, self.upload, self._d_prev = gather, upload, None
        self.patch_size_x, self.patch_size_y = float(patch_size_x), float(patch_size_y)
        self.reset_at_each_frame = int(reset_at_each_frame)
        self.reinit_at_each_frame = self.reset_at_each_frame == 1          # GridTracker.cc:138
        self.dyn_patch_size, self.patch_centroid_inside = int(dyn_patch_size), bool(patch_centroid_inside)
        # forward-backward error estimation (GridTracker.cc:186-189, 292-343)
        self.device_layout = bool(device_layout)       # False: the host-side layout (NumPy), the reference for the device one
        self.fb_err_thresh, self.fb_reinit = float(fb_err_thresh), bool(fb_reinit)
        self.enable_fb_err_est = self.fb_err_thresh > 0
        self._curr_img = self._prev_img = None
        self.fb_err_mask = None
        # GridTrackerParams::updateRes (GridTracker.cc:85-93)
        self.resx, self.resy = (self.gx + 1, self.gy + 1) if (self.dyn_patch_size or self.patch_centroid_inside) else (self.gx, self.gy)
        self.ssm = ssm
        self.est_params = est_params if est_params is not None else api.make_est_params()
        self.seed = int(seed)
        self.frame = 0
        self.cells = cells if cells is not None else api.BatchTracker(cell_params)
        if self.gather is None:
            self.cells.grid_enable()
        self.n_trackers = n
        self.corners = None
        self.pts = None
        self.ssm_update = None
        self.pix_mask = np.ones(n, dtype=np.uint8)
        self.last_estimate = None
        self._ids = None

    # the region's sample points after ssm.setCorners / ssm.initialize (ProjectiveBase.cc:27-39, Homography.cc:50-71)
    def _set_corners(self, corners):
        self.corners = np.array(corners, dtype=np.float64).reshape(2, 4)
        self.pts = pts_from_corners(self.corners, self.resx, self.resy)

    def cell_corners(self):
        """GridTracker::resetTrackers (GridTracker.cc:345-392): the region of every patch tracker, (P, 2, 4)"""
        if self._ids is None:
            t = np.arange(self.n_trackers)
            r, c = t // self.gx, t % self.gx
            w = self.gx + 1
            self._ids = np.stack([r * w + c, r * w + c + 1, (r + 1) * w + c + 1, (r + 1) * w + c], axis=0)      # (4, P)
        pc = None
        if self.pts is None:                       # (the device-side layout keeps only the region's corners on the host)
            self.pts = pts_from_corners(self.corners, self.resx, self.resy)
        if self.resx == self.gx + 1:
            pc = self.pts.take(self._ids, axis=1)                                                            # (2, 4, P)
        # (with resx = grid_size_x the reference still indexes the (grid_size + 1)-wide table, GridTracker.cc:361-371, and then
        # overwrites the result: only the fixed-size branch is meaningful there)
        out = np.empty((2, 4, self.n_trackers))
        if not self.dyn_patch_size:
            if self.patch_centroid_inside:
                cen = (pc[:, 0] + pc[:, 1] + pc[:, 2] + pc[:, 3]) / 4.0                                      # (2, P)
            else:
                cen = self.pts[:, :self.n_trackers]
            x0, y0 = cen[0] - self.patch_size_x / 2.0, cen[1] - self.patch_size_y / 2.0
            x1, y1 = x0 + self.patch_size_x, y0 + self.patch_size_y
            out[0, 0] = x0; out[0, 3] = x0; out[0, 1] = x1; out[0, 2] = x1
            out[1, 0] = y0; out[1, 1] = y0; out[1, 2] = y1; out[1, 3] = y1
        else:
            out[:] = pc
        out = np.ascontiguousarray(out.transpose(2, 0, 1))
        return out

    def _reset_trackers(self, reinit):
        cells = self.cell_corners()
        if self.gather is not None:
            self._d_prev = self.upload(cells.reshape(-1, 8))
        mine = cells[self.shard[0]:self.shard[1]]
        if reinit:
            self.cells.initialize(mine)
        else:
            self.cells.setRegion(mine)

    def setImage(self, img):
        self._curr_img = img
        self.cells.setImage(img)

    @staticmethod
    def _clone(img):
        return img.clone() if hasattr(img, "clone") else np.array(img, copy=True)

    def initialize(self, corners, img=None):
        if img is not None:
            self.setImage(img)
        self._set_corners(corners)
        self._reset_trackers(True)
        if self.enable_fb_err_est:
            self._prev_img = self._clone(self._curr_img)                       # GridTracker.cc:241-243

    @staticmethod
    def _centroids(regions):
        """utils::getCentroid of every region as cv::Point2f: the mean of the four corners in double, rounded to float"""
        c = np.asarray(regions, dtype=np.float64).reshape(-1, 2, 4)
        return np.stack([(c[:, 0, 0] + c[:, 0, 1] + c[:, 0, 2] + c[:, 0, 3]) / 4.0,
                         (c[:, 1, 0] + c[:, 1, 1] + c[:, 1, 2] + c[:, 1, 3]) / 4.0], axis=1).astype(np.float32)

    def _backward_estimation(self, ep):
        """GridTracker::backwardEstimation (GridTracker.cc:292-343): every cell tracks back into the previous frame from where
        it is now; cells that do not return to where they started (squared distance > fb_err_thresh) are left out of the
        estimation.  Batched: one backward update of all the cells instead of a loop over trackers."""
        prev_pts, _ = self.cells.grid_pts()                                    # prev_pts of the last reset / frame
        loc = self.cells.getRegion()
        curr_pts = self._centroids(loc)
        if self.fb_reinit:
            self.cells.initialize(loc)
        self.cells.setImage(self._prev_img)
        self.cells.update()
        fb_prev_pts = self._centroids(self.cells.getRegion())
        self.cells.setImage(self._curr_img)
        self.cells.setRegion(loc)
        # cv::Point2f coordinates: the difference is a float operation, its square a double one (GridTracker.cc:310-313)
        d = (fb_prev_pts - prev_pts).astype(np.float64)
        first = ~((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) > self.fb_err_thresh)
        mask = first.copy()
        late = []
        if int(mask.sum()) < ep.n_model_pts:
            for t in range(self.n_trackers):
                if mask[t]:
                    continue
                mask[t] = True; late.append(t)
                if int(mask.sum()) == ep.n_model_pts:
                    break
        order = np.array(list(np.flatnonzero(first)) + late, dtype=np.int64)
        est = dict(self.cells.estimate_warp_from_pts(self.ssm, prev_pts[order], curr_pts[order], ep))
        pix_mask = np.zeros(self.n_trackers, dtype=np.uint8)
        pix_mask[np.flatnonzero(mask)] = np.asarray(est["mask"])[:int(mask.sum())]    # walked in tracker order (GridTracker.cc:336-342)
        self.fb_err_mask = mask
        est["mask_est"], est["mask"], est["order"] = est["mask"], pix_mask, order
        return est

    def setRegion(self, corners):
        self._set_corners(corners)
        self._reset_trackers(self.reinit_at_each_frame)

    def update(self, img=None):
        if img is not None:
            self.setImage(img)
        self.cells.update()
        return self.finish_update()

    def finish_update(self):
        """GridTracker::update after the cells' own update (GridTracker.cc:265-285)"""
        # every frame's estimator is a new object with a fresh seed in the reference (SSMEstimator.cc:22-24); here seed + frame
        self.frame += 1
        ep = api.EstParams.from_buffer_copy(self.est_params)
        ep.seed = self.seed + self.frame
        if (self.gather is None and not self.enable_fb_err_est and self.reinit_at_each_frame and not self.dyn_patch_size
                and self.patch_centroid_inside and self.device_layout):
            # the shipped grid (reset_at_each_frame = 1, fixed patch size, centroids inside): estimate, region, cell layout and
            # re-initialisation without a host hop (mtfb_grid_advance); the host keeps its copy of the region
            est, region = self.cells.grid_advance(self.ssm, ep, self.gx, self.gy, self.patch_size_x, self.patch_size_y, self.corners)
            self.last_estimate = est
            self.ssm_update, self.pix_mask = est["state_update"], est["mask"]
            self.corners, self.pts = region, None
            return self.getRegion()
        if self.enable_fb_err_est:
            est = self._backward_estimation(ep)
            self._prev_img = self._clone(self._curr_img)
        elif self.gather is None:
            est = self.cells.grid_estimate(self.ssm, ep)
        else:
            curr = self.gather()
            est = self.cells.estimate_warp_from_corners_device(self.ssm, self._d_prev.data_ptr(), curr.data_ptr(), self.n_trackers, ep)
        self.last_estimate = est
        self.ssm_update, self.pix_mask = est["state_update"], est["mask"]
        self._set_corners(apply_warp_to_corners(self.ssm, self.corners, self.ssm_update))
        if self.reset_at_each_frame:
            self._reset_trackers(self.reinit_at_each_frame)
        elif self.enable_fb_err_est:
            pass                                 # the backward pass's setRegion left prev_pts = the cells' current centroids
        elif self.gather is None:
            self.cells.grid_commit()
        else:
            self._d_prev = curr.clone()
        return self.getRegion()

    def getRegion(self):
        return self.corners.copy()

    def close(self):
        self.cells.close()
