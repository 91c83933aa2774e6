"""Oracle (test infrastructure): the slice of `filterpy.kalman.KalmanFilter` (filterpy 1.4.x, absent from this image and
from /root/reference) that aether/utils/postprocess_utils.py:751-844 `smooth_trajectory` uses.  PARITY UNPINNED: restated
from the published algorithm -- predict: x <- F x, P <- F P F^T + Q; update: y = z - H x, S = H P H^T + R,
K = P H^T S^-1, x <- x + K y, P <- (I - K H) P (I - K H)^T + K R K^T (Joseph form); defaults x = 0 (column vector),
P = Q = I, R = I, F = I, H = 0.  Only used by tests/golden/_reference_shim.py as the `filterpy` stand-in when the
reference's own smoothing code is executed to produce goldens."""
import numpy as np


class KalmanFilter:
    def __init__(self, dim_x, dim_z, dim_u=0):
        self.dim_x, self.dim_z = dim_x, dim_z
        self.x = np.zeros((dim_x, 1))
        self.P = np.eye(dim_x)
        self.Q = np.eye(dim_x)
        self.F = np.eye(dim_x)
        self.H = np.zeros((dim_z, dim_x))
        self.R = np.eye(dim_z)

    def predict(self):
        self.x = self.F @ self.x
        self.P = self.F @ self.P @ self.F.T + self.Q

    def update(self, z):
        z = np.asarray(z, dtype=float)
        if self.x.ndim == 2:
            z = z.reshape(self.dim_z, 1)
        innovation = z - self.H @ self.x
        pht = self.P @ self.H.T
        gain = pht @ np.linalg.inv(self.H @ pht + self.R)
        self.x = self.x + gain @ innovation
        ikh = np.eye(self.dim_x) - gain @ self.H
        self.P = ikh @ self.P @ ikh.T + gain @ self.R @ gain.T
