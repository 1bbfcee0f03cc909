"""Host-side geometries for collocation-point generation and boundary filtering.

Only what sits either side of the hot path: where points come from and which rows a BC owns.
Behaviour follows the DeepXDE calls the reference makes:
`random_points` uses the global NumPy RNG ('pseudo'), which the resamplers rely on
(/root/reference/AI4XDE/algorithm/RAD.py:25, R3Sampling.py:97).  Quasi-random distributions
(DeepXDE delegates to scikit-optimize, absent here) fall back to a van der Corput/Hammersley
construction with the same uniform law.
"""
import itertools

import numpy as np

from . import config


def _real():
    return config.real(np)


def _radical_inverse(count, base):
    idx = np.arange(1, count + 1)
    out = np.zeros(count)
    f = 1.0 / base
    while np.any(idx > 0):
        out += f * (idx % base)
        idx //= base
        f /= base
    return out


_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19)


def sample(n_samples, dimension, sampler="pseudo"):
    """n_samples x dimension points in the unit cube."""
    if sampler == "pseudo":
        return np.random.random(size=(n_samples, dimension)).astype(_real())
    if sampler in ("LHS", "Halton", "Hammersley", "Sobol"):
        cols = []
        bases = list(_PRIMES[:dimension])
        if sampler == "Hammersley" and dimension > 1:
            cols.append((np.arange(n_samples) + 0.5) / n_samples)
            bases = bases[: dimension - 1]
        cols += [_radical_inverse(n_samples, b) for b in bases]
        return np.stack(cols, axis=1).astype(_real())
    raise ValueError(f"{sampler} sampling is not available.")


class Geometry:
    dim = 0

    def inside(self, x):
        raise NotImplementedError

    def on_boundary(self, x):
        raise NotImplementedError

    def boundary_normal(self, x):
        raise NotImplementedError

    def random_points(self, n, random="pseudo"):
        raise NotImplementedError

    def random_boundary_points(self, n, random="pseudo"):
        raise NotImplementedError

    def uniform_points(self, n, boundary=True):
        return self.random_points(n)

    def uniform_boundary_points(self, n):
        return self.random_boundary_points(n)

    def periodic_point(self, x, component):
        raise NotImplementedError

    def __sub__(self, other):
        return CSGDifference(self, other)


class Interval(Geometry):
    dim = 1

    def __init__(self, l, r):
        self.l, self.r = l, r
        self.bbox = (np.array([l]), np.array([r]))
        self.diam = r - l

    def inside(self, x):
        return ((x >= self.l) & (x <= self.r)).reshape(-1)

    def on_boundary(self, x):
        return np.isclose(x, self.l).any(axis=-1) | np.isclose(x, self.r).any(axis=-1)

    def boundary_normal(self, x):
        return np.isclose(x, self.r).astype(_real()) - np.isclose(x, self.l).astype(_real())

    def uniform_points(self, n, boundary=True):
        if boundary:
            return np.linspace(self.l, self.r, num=n, dtype=_real())[:, None]
        return np.linspace(self.l, self.r, num=n + 1, endpoint=False, dtype=_real())[1:, None]

    def random_points(self, n, random="pseudo"):
        return (self.diam * sample(n, 1, random) + self.l).astype(_real())

    def uniform_boundary_points(self, n):
        if n == 1:
            return np.array([[self.l]], dtype=_real())
        return np.concatenate([np.full((n // 2, 1), self.l), np.full((n - n // 2, 1), self.r)]).astype(_real())

    def random_boundary_points(self, n, random="pseudo"):
        if n == 2:
            return np.array([[self.l], [self.r]], dtype=_real())
        return np.random.choice([self.l, self.r], n)[:, None].astype(_real())

    def periodic_point(self, x, component=0):
        out = np.copy(x)
        out[np.isclose(x, self.l)] = self.r
        out[np.isclose(x, self.r)] = self.l
        return out


class TimeDomain(Interval):
    def __init__(self, t0, t1):
        super().__init__(t0, t1)
        self.t0, self.t1 = t0, t1

    def on_initial(self, t):
        return np.isclose(t, self.t0).reshape(-1)


class Hypercube(Geometry):
    def __init__(self, xmin, xmax):
        if len(xmin) != len(xmax):
            raise ValueError("Dimensions of xmin and xmax do not match.")
        self.xmin = np.array(xmin, dtype=_real())
        self.xmax = np.array(xmax, dtype=_real())
        if np.any(self.xmin >= self.xmax):
            raise ValueError("xmin >= xmax")
        self.dim = len(xmin)
        self.side_length = self.xmax - self.xmin
        self.bbox = (self.xmin, self.xmax)
        self.diam = np.linalg.norm(self.side_length)
        self.volume = np.prod(self.side_length)

    def inside(self, x):
        return np.all(x >= self.xmin, axis=-1) & np.all(x <= self.xmax, axis=-1)

    def on_boundary(self, x):
        touching = np.isclose(x, self.xmin).any(axis=-1) | np.isclose(x, self.xmax).any(axis=-1)
        return self.inside(x) & touching

    def boundary_normal(self, x):
        nrm = np.isclose(x, self.xmax).astype(_real()) - np.isclose(x, self.xmin).astype(_real())
        corner = np.count_nonzero(nrm, axis=-1) > 1
        if corner.any():
            nrm[corner] /= np.linalg.norm(nrm[corner], axis=-1, keepdims=True)
        return nrm

    def uniform_points(self, n, boundary=True):
        dx = (self.volume / n) ** (1 / self.dim)
        axes = []
        for i in range(self.dim):
            ni = int(np.ceil(self.side_length[i] / dx))
            if boundary:
                axes.append(np.linspace(self.xmin[i], self.xmax[i], num=ni, dtype=_real()))
            else:
                axes.append(np.linspace(self.xmin[i], self.xmax[i], num=ni + 1, endpoint=False, dtype=_real())[1:])
        return np.array(list(itertools.product(*axes)))

    def random_points(self, n, random="pseudo"):
        return self.side_length * sample(n, self.dim, random) + self.xmin

    def random_boundary_points(self, n, random="pseudo"):
        x = sample(n, self.dim, random)
        face = np.random.randint(self.dim, size=n)
        rows = np.arange(n)
        x[rows, face] = np.round(x[rows, face])
        return self.side_length * x + self.xmin

    def periodic_point(self, x, component):
        y = np.copy(x)
        lo = np.isclose(y[:, component], self.xmin[component])
        hi = np.isclose(y[:, component], self.xmax[component])
        y[lo, component] = self.xmax[component]
        y[hi, component] = self.xmin[component]
        return y


class Rectangle(Hypercube):
    def __init__(self, xmin, xmax):
        super().__init__(xmin, xmax)
        self.perimeter = 2 * np.sum(self.side_length)
        self.area = np.prod(self.side_length)

    def _walk(self, s):
        """Map arc length s (0 at the bottom-left corner, counter-clockwise) to boundary points."""
        w, h = self.side_length
        l1, l2, l3 = w, w + h, 2 * w + h
        out = np.empty((len(s), 2))
        for i, l in enumerate(s):
            if l < l1:
                out[i] = (self.xmin[0] + l, self.xmin[1])
            elif l < l2:
                out[i] = (self.xmax[0], self.xmin[1] + l - l1)
            elif l < l3:
                out[i] = (self.xmax[0] - l + l2, self.xmax[1])
            else:
                out[i] = (self.xmin[0], self.xmax[1] - l + l3)
        return out

    def uniform_boundary_points(self, n):
        nx, ny = np.ceil(n / self.perimeter * self.side_length).astype(int)
        bottom = np.stack([np.linspace(self.xmin[0], self.xmax[0], nx, endpoint=False), np.full(nx, self.xmin[1])], 1)
        right = np.stack([np.full(ny, self.xmax[0]), np.linspace(self.xmin[1], self.xmax[1], ny, endpoint=False)], 1)
        top = np.stack([np.linspace(self.xmin[0], self.xmax[0], nx + 1)[1:], np.full(nx, self.xmax[1])], 1)
        left = np.stack([np.full(ny, self.xmin[0]), np.linspace(self.xmin[1], self.xmax[1], ny + 1)[1:]], 1)
        return np.concatenate([bottom, right, top, left]).astype(_real())

    def random_boundary_points(self, n, random="pseudo"):
        w, h = self.side_length
        u = np.ravel(sample(n + 2, 1, random))
        # drop draws that land on the two corners where the normal flips, then keep n
        u = u[~np.isclose(u, w / self.perimeter)]
        u = u[~np.isclose(u, (2 * w + h) / self.perimeter)]
        return self._walk(u[:n] * self.perimeter).astype(_real())


class Cuboid(Hypercube):
    pass


class Disk(Geometry):
    dim = 2

    def __init__(self, center, radius):
        self.center = np.array(center, dtype=_real())
        self.radius = radius
        self.bbox = (self.center - radius, self.center + radius)
        self.diam = 2 * radius

    def inside(self, x):
        return np.linalg.norm(x - self.center, axis=-1) <= self.radius

    def on_boundary(self, x):
        return np.isclose(np.linalg.norm(x - self.center, axis=-1), self.radius)

    def boundary_normal(self, x):
        v = x - self.center
        l = np.linalg.norm(v, axis=-1, keepdims=True)
        return v / l * np.isclose(l, self.radius)

    def random_points(self, n, random="pseudo"):
        rng = sample(n, 2, random)
        r, theta = np.sqrt(rng[:, 0]), 2 * np.pi * rng[:, 1]
        pts = np.stack([r * np.cos(theta), r * np.sin(theta)], axis=1)
        return (self.radius * pts + self.center).astype(_real())

    def uniform_boundary_points(self, n):
        theta = np.linspace(0, 2 * np.pi, num=n, endpoint=False)
        return (self.radius * np.stack([np.cos(theta), np.sin(theta)], 1) + self.center).astype(_real())

    def random_boundary_points(self, n, random="pseudo"):
        theta = 2 * np.pi * sample(n, 1, random)
        return (self.radius * np.hstack([np.cos(theta), np.sin(theta)]) + self.center).astype(_real())


class Sphere(Geometry):
    """dde.geometry.Sphere (the ball of the reference's 3-D fractional Poisson case, cases/PDECases.py:2346).
    Pseudo-random interior points: radius U^(1/3) times a normalised Gaussian direction, U drawn before the
    directions; boundary points: normalised Gaussian directions."""
    dim = 3

    def __init__(self, center, radius):
        self.center = np.array(center, dtype=_real())
        self.radius = radius
        self.bbox = (self.center - radius, self.center + radius)
        self.diam = 2 * radius

    def _distance(self, x):
        return np.linalg.norm(x - self.center, axis=-1)

    def inside(self, x):
        return self._distance(x) <= self.radius

    def on_boundary(self, x):
        return np.isclose(self._distance(x), self.radius)

    def boundary_normal(self, x):
        v = x - self.center
        l = np.linalg.norm(v, axis=-1, keepdims=True)
        return v / l * np.isclose(l, self.radius)

    def _directions(self, n, random, skip=0):
        if random == "pseudo":
            g = np.random.normal(size=(n, self.dim))
        else:
            from scipy import stats

            g = stats.norm.ppf(sample(n, self.dim + skip, random)[:, skip:])
        return g / np.linalg.norm(g, axis=1, keepdims=True)

    def random_points(self, n, random="pseudo"):
        if random == "pseudo":
            u = np.random.rand(n, 1)
            d = self._directions(n, random)
        else:
            u = sample(n, self.dim + 1, random)[:, 0:1]
            d = self._directions(n, random, skip=1)
        return (self.radius * u ** (1 / self.dim) * d + self.center).astype(_real())

    def random_boundary_points(self, n, random="pseudo"):
        return (self.radius * self._directions(n, random) + self.center).astype(_real())


class Polygon(Geometry):
    """Simple polygon given by its vertices in order (either orientation; stored clockwise as DeepXDE does) --
    `dde.geometry.Polygon([[0, 0], [1, 0], [1, -1], [-1, -1], [-1, 1], [0, 1]])` is the L-shaped domain of
    /root/reference/AI4XDE/cases/PoissonCases.py:515-516.  `inside` is the winding-number test, `on_boundary`
    the closed-segment test |x - v_i| + |x - v_{i+1}| = |v_{i+1} - v_i| (np.isclose), `boundary_normal` the
    outward unit normal of the segment a point lies on; points come from rejection sampling in the bounding
    box (interior) and from arc-length sampling of the perimeter (boundary)."""

    dim = 2

    def __init__(self, vertices):
        v = np.array(vertices, dtype=_real())
        if v.ndim != 2 or v.shape[1] != 2 or len(v) < 3:
            raise ValueError("Polygon needs an (n >= 3, 2) array of vertices")
        if len(v) == 3:
            raise ValueError("The polygon is a triangle. Use Triangle instead.")
        area2 = float(np.sum(v[:, 0] * np.roll(v[:, 1], -1) - np.roll(v[:, 0], -1) * v[:, 1]))
        if area2 > 0:                                  # counter-clockwise input -> store clockwise
            v = np.flipud(v)
        self.area = abs(area2) / 2
        self.vertices = v
        self.nvertices = len(v)
        self.bbox = (np.min(v, axis=0), np.max(v, axis=0))
        self.diam = float(max(np.linalg.norm(a - b) for a in v for b in v))
        self.segments = np.roll(v, -1, axis=0) - v     # segment i: v_i -> v_{i+1}
        self.lengths = np.linalg.norm(self.segments, axis=1)
        self.perimeter = float(self.lengths.sum())
        # clockwise orientation: the outward normal is the segment direction rotated by +90 degrees
        n = np.stack([-self.segments[:, 1], self.segments[:, 0]], axis=1)
        self.normal = n / np.linalg.norm(n, axis=1, keepdims=True)

    def inside(self, x):
        x = np.atleast_2d(x)
        a, b = self.vertices, np.roll(self.vertices, -1, axis=0)
        wn = np.zeros(len(x), dtype=int)
        for p, q in zip(a, b):
            left = (q[0] - p[0]) * (x[:, 1] - p[1]) - (x[:, 0] - p[0]) * (q[1] - p[1])
            up = (p[1] <= x[:, 1]) & (q[1] > x[:, 1]) & (left > 0)
            down = (p[1] > x[:, 1]) & (q[1] <= x[:, 1]) & (left < 0)
            wn += up.astype(int) - down.astype(int)
        return (wn != 0) | self.on_boundary(x)

    def _segment_mask(self, x):
        x = np.atleast_2d(x)
        a, b = self.vertices, np.roll(self.vertices, -1, axis=0)
        l1 = np.linalg.norm(x[:, None, :] - a[None], axis=-1)
        l2 = np.linalg.norm(x[:, None, :] - b[None], axis=-1)
        return np.isclose(l1 + l2, self.lengths[None])

    def on_boundary(self, x):
        return self._segment_mask(x).any(axis=1)

    def boundary_normal(self, x):
        m = self._segment_mask(x)
        first = np.argmax(m, axis=1)                   # a corner takes the normal of its first segment
        return (self.normal[first] * m.any(axis=1, keepdims=True)).astype(_real())

    def random_points(self, n, random="pseudo"):
        out = np.empty((0, 2), dtype=_real())
        lo, hi = self.bbox
        while len(out) < n:
            cand = sample(n, 2, "pseudo" if len(out) else random) * (hi - lo) + lo
            cand = cand[self.inside(cand) & ~self.on_boundary(cand)]
            out = np.vstack((out, cand.astype(_real())))
        return out[:n]

    def _at_arclength(self, s):
        edges = np.concatenate(([0.0], np.cumsum(self.lengths)))
        i = np.clip(np.searchsorted(edges, s, side="right") - 1, 0, self.nvertices - 1)
        t = (s - edges[i]) / self.lengths[i]
        return (self.vertices[i] + t[:, None] * self.segments[i]).astype(_real())

    def uniform_boundary_points(self, n):
        return self._at_arclength(np.linspace(0, self.perimeter, num=n, endpoint=False))

    def random_boundary_points(self, n, random="pseudo"):
        return self._at_arclength(np.ravel(sample(n, 1, random)) * self.perimeter)


class CSGDifference(Geometry):
    """geom1 - geom2 (e.g. rectangle minus disk, /root/reference/AI4XDE/cases/PDECases.py:679)."""

    def __init__(self, geom1, geom2):
        if geom1.dim != geom2.dim:
            raise ValueError("geometries differ in dimension")
        self.geom1, self.geom2 = geom1, geom2
        self.dim = geom1.dim
        self.bbox = geom1.bbox
        self.diam = geom1.diam

    def inside(self, x):
        return self.geom1.inside(x) & ~self.geom2.inside(x)

    def on_boundary(self, x):
        return (self.geom1.on_boundary(x) & ~self.geom2.inside(x)) | (
            self.geom1.inside(x) & self.geom2.on_boundary(x))

    def boundary_normal(self, x):
        a = (self.geom1.on_boundary(x) & ~self.geom2.inside(x))[:, None] * self.geom1.boundary_normal(x)
        b = (self.geom1.inside(x) & self.geom2.on_boundary(x))[:, None] * -self.geom2.boundary_normal(x)
        return a + b

    def _fill(self, n, draw, keep):
        out = np.empty((n, self.dim), dtype=_real())
        have = 0
        while have < n:
            cand = draw(n)
            cand = cand[keep(cand)]
            take = min(len(cand), n - have)
            out[have : have + take] = cand[:take]
            have += take
        return out

    def random_points(self, n, random="pseudo"):
        return self._fill(n, lambda m: self.geom1.random_points(m, random=random),
                          lambda c: ~self.geom2.inside(c))

    def random_boundary_points(self, n, random="pseudo"):
        def draw(m):
            a = self.geom1.random_boundary_points(m, random=random)
            a = a[~self.geom2.inside(a)]
            b = self.geom2.random_boundary_points(m, random=random)
            b = b[self.geom1.inside(b)]
            return np.random.permutation(np.concatenate([a, b]))

        return self._fill(n, draw, lambda c: np.ones(len(c), bool))


class GeometryXTime:
    def __init__(self, geometry, timedomain):
        self.geometry, self.timedomain = geometry, timedomain
        self.dim = geometry.dim + timedomain.dim

    def on_boundary(self, x):
        return self.geometry.on_boundary(x[:, :-1])

    def on_initial(self, x):
        return self.timedomain.on_initial(x[:, -1:])

    def boundary_normal(self, x):
        nrm = self.geometry.boundary_normal(x[:, :-1])
        return np.hstack([nrm, np.zeros((len(nrm), 1))])

    def _box(self):
        if isinstance(self.geometry, Interval):
            return Hypercube([self.geometry.l, self.timedomain.t0], [self.geometry.r, self.timedomain.t1])
        if isinstance(self.geometry, Hypercube):
            return Hypercube(np.append(self.geometry.xmin, self.timedomain.t0),
                             np.append(self.geometry.xmax, self.timedomain.t1))
        return None

    def random_points(self, n, random="pseudo"):
        box = self._box()
        if box is not None:
            return box.random_points(n, random=random)
        x = self.geometry.random_points(n, random=random)
        t = np.random.permutation(self.timedomain.random_points(n, random=random))
        return np.hstack((x, t))

    def uniform_points(self, n, boundary=True):
        span = np.prod(self.geometry.bbox[1] - self.geometry.bbox[0])
        nx = int(np.ceil((n * span / self.timedomain.diam) ** 0.5))
        nt = int(np.ceil(n / nx))
        x = self.geometry.uniform_points(nx, boundary=boundary)
        if boundary:
            t = self.timedomain.uniform_points(nt, boundary=True)[:, 0]
        else:
            t = np.linspace(self.timedomain.t1, self.timedomain.t0, num=nt, endpoint=False, dtype=_real())
        return np.vstack([np.hstack((x, np.full((len(x), 1), ti, dtype=_real()))) for ti in t])

    def random_boundary_points(self, n, random="pseudo"):
        x = self.geometry.random_boundary_points(n, random=random)
        t = np.random.permutation(self.timedomain.random_points(n, random=random))
        return np.hstack((x, t))

    def uniform_boundary_points(self, n):
        if self.geometry.dim == 1:
            nx = 2
        else:
            sides = self.geometry.bbox[1] - self.geometry.bbox[0]
            s = 2 * sum(a * b for a, b in itertools.combinations(sides, 2))
            nx = int((n * s / self.timedomain.diam) ** 0.5)
        nt = int(np.ceil(n / nx))
        x = self.geometry.uniform_boundary_points(nx)
        t = np.linspace(self.timedomain.t1, self.timedomain.t0, num=nt, endpoint=False, dtype=_real())
        return np.vstack([np.hstack((x, np.full((len(x), 1), ti, dtype=_real()))) for ti in t])

    def random_initial_points(self, n, random="pseudo"):
        x = self.geometry.random_points(n, random=random)
        return np.hstack((x, np.full((n, 1), self.timedomain.t0, dtype=_real())))

    def uniform_initial_points(self, n):
        x = self.geometry.uniform_points(n, True)
        return np.hstack((x, np.full((len(x), 1), self.timedomain.t0, dtype=_real())))

    def periodic_point(self, x, component):
        return np.hstack([self.geometry.periodic_point(x[:, :-1], component), x[:, -1:]])
