"""Race tracks as natural cubic splines (host side, numpy).

Mirrors envs/tracks/tracks.py:14-64 (get_track: centre / inner / outer splines sharing the centre line's
arc-length knots, shifted so the centre starts at the origin) and the coefficient fit of
envs/tracks/torch_spline.py:52-108, :240-301, :385-441 (natural cubic spline through a tridiagonal solve for
the knot derivatives).  The fit runs once per track on the host; the kernels only evaluate the tables
(ilqc_b200/csrc/models.cuh: spline_locate / spline_eval / spline_deriv).
"""
import functools
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "track_points.npz")
LOOP_TRACKS = ("simple", "complex", "circle")  # handle_termination='loop' (tracks.py:55-56)
TRACK_NAMES = ("simple", "complex", "circle", "line", "bend")


def _thomas(rhs, upper, diag, lower):
    """Tridiagonal solve, forward elimination + back substitution (torch_spline.py:385-441)."""
    n = diag.shape[0]
    new_d = np.empty(n)
    new_b = np.empty(n)
    new_d[0] = diag[0]
    new_b[0] = rhs[0]
    for i in range(1, n):
        w = lower[i - 1] / new_d[i - 1]
        new_d[i] = diag[i] - w * upper[i - 1]
        new_b[i] = rhs[i] - w * new_b[i - 1]
    out = np.empty(n)
    out[n - 1] = new_b[n - 1] / new_d[n - 1]
    for i in range(n - 2, -1, -1):
        out[i] = (new_b[i] - upper[i] * out[i + 1]) / new_d[i]
    return out


def natural_cubic_spline_coeffs(t, x):
    """Coefficients (a, b, c, d) of the natural cubic spline through (t_i, x_i); x has shape (n,) ; each
    returned array has shape (n-1,): value on segment i is a+(b+(c+d f)f)f with f = t - t_i."""
    t = np.asarray(t, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n == 2:
        return x[:1], (x[1:] - x[:1]) / (t[1:] - t[:1]), np.zeros(1), np.zeros(1)
    h_inv = 1.0 / (t[1:] - t[:-1])
    h_inv2 = h_inv ** 2
    three_dx = 3.0 * (x[1:] - x[:-1])
    six_dx = 2.0 * three_dx
    scaled = three_dx * h_inv2
    diag = np.empty(n)
    diag[:-1] = h_inv
    diag[-1] = 0.0
    diag[1:] += h_inv
    diag *= 2.0
    rhs = np.empty(n)
    rhs[:-1] = scaled
    rhs[-1] = 0.0
    rhs[1:] += scaled
    m = _thomas(rhs, h_inv, diag, h_inv)  # derivatives at the knots
    a = x[:-1]
    b = m[:-1]
    two_c = (six_dx * h_inv - 4.0 * m[:-1] - 2.0 * m[1:]) * h_inv
    three_d = (-six_dx * h_inv + 3.0 * (m[:-1] + m[1:])) * h_inv2
    return a, b, two_c / 2.0, three_d / 3.0


class Track:
    """knots [n+1], coef [n_seg, 3 (centre, inner, outer), 2 (x, y), 4 (a, b, c, d)], loop flag."""

    def __init__(self, name, knots, coef, loop):
        self.name, self.knots, self.coef, self.loop = name, knots, coef, loop
        self.length = float(knots[-1])
        self.n_seg = coef.shape[0]

    # host-side evaluation (numpy), same semantics as torch_spline.py:341-375; used for x0 / obstacle set-up
    def _locate(self, t):
        if self.loop:
            t = np.remainder(t, self.length)
        else:
            # smooth_min(t, L) away from the 1e-6 blend zone around L (torch_utils.py:72); the host only
            # evaluates the splines at t = 0 and L/20 (initial heading, obstacle)
            t = -(self.length - t) + self.length if self.length - t > 1e-6 else self.length
        idx = int(np.clip(np.searchsorted(self.knots, t, side="left") - 1, 0, self.n_seg - 1))
        return t - self.knots[idx], idx

    def evaluate(self, t, which=0):
        f, i = self._locate(float(t))
        c = self.coef[i, which]
        return c[:, 0] + (c[:, 1] + (c[:, 2] + c[:, 3] * f) * f) * f

    def derivative(self, t, which=0):
        f, i = self._locate(float(t))
        c = self.coef[i, which]
        return c[:, 1] + (2.0 * c[:, 2] + 3.0 * c[:, 3] * f) * f


@functools.lru_cache(maxsize=None)
def get_track(name):
    if name not in TRACK_NAMES:
        raise NotImplementedError(f"track {name}")
    data = np.load(_DATA)
    lines = []
    start = None
    knots = None
    for suffix in ("", "_i", "_o"):
        line = np.stack([data[f"{name}_X{suffix}"], data[f"{name}_Y{suffix}"]], axis=1).astype(np.float64)
        if start is None:
            start = line[0].copy()
        line = line - start
        if knots is None:
            seg = np.sqrt(np.sum(np.diff(line, axis=0) ** 2, axis=1))
            knots = np.insert(seg, 0, 0).cumsum()
        lines.append(line)
    n_seg = knots.shape[0] - 1
    coef = np.empty((n_seg, 3, 2, 4))
    for s, line in enumerate(lines):
        for ch in range(2):
            a, b, c, d = natural_cubic_spline_coeffs(knots, line[:, ch])
            coef[:, s, ch, 0], coef[:, s, ch, 1], coef[:, s, ch, 2], coef[:, s, ch, 3] = a, b, c, d
    return Track(name, knots, np.ascontiguousarray(coef), name in LOOP_TRACKS)
