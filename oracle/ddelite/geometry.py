"""Geometry samplers/filters restated from DeepXDE semantics.

Only the shapes the BASELINE configs and their neighbours use: Interval, TimeDomain,
Hypercube/Rectangle/Cuboid, Disk, GeometryXTime.  'pseudo' sampling is
np.random.random (the global legacy RNG, which the resamplers depend on:
/root/reference/AI4XDE/algorithm/RAD.py:25).  Quasi-random names ('Hammersley', ...)
come from scikit-optimize in DeepXDE; skopt is absent here, so they use a plain
van-der-Corput/Hammersley construction (same distribution, not the same points).
"""
import itertools

import numpy as np

from . import config


def _primes(n):
    out, c = [], 2
    while len(out) < n:
        if all(c % p for p in out):
            out.append(c)
        c += 1
    return out


def _van_der_corput(n, base, start=1):
    seq = np.zeros(n)
    for idx in range(n):
        k, f, v = idx + start, 1.0 / base, 0.0
        while k > 0:
            v += f * (k % base)
            k //= base
            f /= base
        seq[idx] = v
    return seq


def sample(n_samples, dimension, sampler="pseudo"):
    if sampler == "pseudo":
        return np.random.random(size=(n_samples, dimension)).astype(config.real(np))
    if sampler in ("LHS", "Halton", "Hammersley", "Sobol"):
        cols = []
        primes = _primes(dimension)
        if sampler == "Hammersley" and dimension > 1:
            cols.append((np.arange(n_samples) + 0.5) / n_samples)
            primes = primes[: dimension - 1]
        for p in primes:
            cols.append(_van_der_corput(n_samples, p))
        return np.stack(cols, axis=1).astype(config.real(np))
    raise ValueError("f{sampler} sampling is not available.")


class Geometry:
    def __init__(self, dim, bbox, diam):
        self.dim = dim
        self.bbox = bbox
        self.diam = min(diam, np.linalg.norm(bbox[1] - bbox[0]))
        self.idstr = type(self).__name__

    def uniform_points(self, n, boundary=True):
        return self.random_points(n)

    def uniform_boundary_points(self, n):
        return self.random_boundary_points(n)

    def periodic_point(self, x, component):
        raise NotImplementedError

    def __sub__(self, other):
        return CSGDifference(self, other)


class CSGDifference(Geometry):
    """geom1 - geom2 (DeepXDE csg.CSGDifference), used at /root/reference/AI4XDE/cases/PDECases.py:679,831."""

    def __init__(self, geom1, geom2):
        if geom1.dim != geom2.dim:
            raise ValueError("{} - {} failed (dimensions do not match).".format(geom1.idstr, geom2.idstr))
        super().__init__(geom1.dim, geom1.bbox, geom1.diam)
        self.geom1, self.geom2 = geom1, geom2

    def inside(self, x):
        return np.logical_and(self.geom1.inside(x), ~self.geom2.inside(x))

    def on_boundary(self, x):
        return np.logical_or(
            np.logical_and(self.geom1.on_boundary(x), ~self.geom2.inside(x)),
            np.logical_and(self.geom1.inside(x), self.geom2.on_boundary(x)),
        )

    def boundary_normal(self, x):
        return np.logical_and(self.geom1.on_boundary(x), ~self.geom2.inside(x))[:, np.newaxis] * self.geom1.boundary_normal(x) \
            + np.logical_and(self.geom1.inside(x), self.geom2.on_boundary(x))[:, np.newaxis] * -self.geom2.boundary_normal(x)

    def random_points(self, n, random="pseudo"):
        x = np.empty(shape=(n, self.dim), dtype=config.real(np))
        i = 0
        while i < n:
            tmp = self.geom1.random_points(n, random=random)
            tmp = tmp[~self.geom2.inside(tmp)]
            if len(tmp) > n - i:
                tmp = tmp[: n - i]
            x[i : i + len(tmp)] = tmp
            i += len(tmp)
        return x

    def random_boundary_points(self, n, random="pseudo"):
        x = np.empty(shape=(n, self.dim), dtype=config.real(np))
        i = 0
        while i < n:
            g1 = self.geom1.random_boundary_points(n, random=random)
            g1 = g1[~self.geom2.inside(g1)]
            g2 = self.geom2.random_boundary_points(n, random=random)
            g2 = g2[self.geom1.inside(g2)]
            tmp = np.random.permutation(np.concatenate((g1, g2)))
            if len(tmp) > n - i:
                tmp = tmp[: n - i]
            x[i : i + len(tmp)] = tmp
            i += len(tmp)
        return x


class Interval(Geometry):
    def __init__(self, l, r):
        super().__init__(1, (np.array([l]), np.array([r])), r - l)
        self.l, self.r = l, r

    def inside(self, x):
        return np.logical_and(self.l <= x, x <= self.r).flatten()

    def on_boundary(self, x):
        return np.any(np.isclose(x, [self.l, self.r]), axis=-1)

    def boundary_normal(self, x):
        return -np.isclose(x, self.l).astype(config.real(np)) + np.isclose(x, self.r).astype(
            config.real(np)
        )

    def uniform_points(self, n, boundary=True):
        if boundary:
            return np.linspace(self.l, self.r, num=n, dtype=config.real(np))[:, None]
        return np.linspace(self.l, self.r, num=n + 1, endpoint=False, dtype=config.real(np))[
            1:, None
        ]

    def random_points(self, n, random="pseudo"):
        x = sample(n, 1, random)
        return (self.diam * x + self.l).astype(config.real(np))

    def uniform_boundary_points(self, n):
        if n == 1:
            return np.array([[self.l]]).astype(config.real(np))
        xl = np.full((n // 2, 1), self.l).astype(config.real(np))
        xr = np.full((n - n // 2, 1), self.r).astype(config.real(np))
        return np.vstack((xl, xr))

    def random_boundary_points(self, n, random="pseudo"):
        if n == 2:
            return np.array([[self.l], [self.r]]).astype(config.real(np))
        return np.random.choice([self.l, self.r], n)[:, None].astype(config.real(np))

    def periodic_point(self, x, component=0):
        tmp = np.copy(x)
        tmp[np.isclose(x, self.l)] = self.r
        tmp[np.isclose(x, self.r)] = self.l
        return tmp


class TimeDomain(Interval):
    def __init__(self, t0, t1):
        super().__init__(t0, t1)
        self.t0, self.t1 = t0, t1

    def on_initial(self, t):
        return np.isclose(t, self.t0).flatten()


class Hypercube(Geometry):
    def __init__(self, xmin, xmax):
        if len(xmin) != len(xmax):
            raise ValueError("Dimensions of xmin and xmax do not match.")
        self.xmin = np.array(xmin, dtype=config.real(np))
        self.xmax = np.array(xmax, dtype=config.real(np))
        if np.any(self.xmin >= self.xmax):
            raise ValueError("xmin >= xmax")
        self.side_length = self.xmax - self.xmin
        super().__init__(len(xmin), (self.xmin, self.xmax), np.linalg.norm(self.side_length))
        self.volume = np.prod(self.side_length)

    def inside(self, x):
        return np.logical_and(np.all(x >= self.xmin, axis=-1), np.all(x <= self.xmax, axis=-1))

    def on_boundary(self, x):
        _on = np.logical_or(
            np.any(np.isclose(x, self.xmin), axis=-1), np.any(np.isclose(x, self.xmax), axis=-1)
        )
        return np.logical_and(self.inside(x), _on)

    def boundary_normal(self, x):
        _n = -np.isclose(x, self.xmin).astype(config.real(np)) + np.isclose(x, self.xmax)
        idx = np.count_nonzero(_n, axis=-1) > 1
        if np.any(idx):
            l = np.linalg.norm(_n[idx], axis=-1, keepdims=True)
            _n[idx] /= l
        return _n

    def uniform_points(self, n, boundary=True):
        dx = (self.volume / n) ** (1 / self.dim)
        xi = []
        for i in range(self.dim):
            ni = int(np.ceil(self.side_length[i] / dx))
            if boundary:
                xi.append(np.linspace(self.xmin[i], self.xmax[i], num=ni, dtype=config.real(np)))
            else:
                xi.append(
                    np.linspace(
                        self.xmin[i], self.xmax[i], num=ni + 1, endpoint=False, dtype=config.real(np)
                    )[1:]
                )
        return np.array(list(itertools.product(*xi)))

    def random_points(self, n, random="pseudo"):
        x = sample(n, self.dim, random)
        return (self.xmax - self.xmin) * x + self.xmin

    def random_boundary_points(self, n, random="pseudo"):
        x = sample(n, self.dim, random)
        rand_dim = np.random.randint(self.dim, size=n)
        x[np.arange(n), rand_dim] = np.round(x[np.arange(n), rand_dim])
        return (self.xmax - self.xmin) * x + self.xmin

    def periodic_point(self, x, component):
        y = np.copy(x)
        _on_xmin = np.isclose(y[:, component], self.xmin[component])
        _on_xmax = np.isclose(y[:, component], self.xmax[component])
        y[:, component][_on_xmin] = self.xmax[component]
        y[:, component][_on_xmax] = self.xmin[component]
        return y


class Rectangle(Hypercube):
    def __init__(self, xmin, xmax):
        super().__init__(xmin, xmax)
        self.perimeter = 2 * np.sum(self.xmax - self.xmin)
        self.area = np.prod(self.xmax - self.xmin)

    def uniform_boundary_points(self, n):
        nx, ny = np.ceil(n / self.perimeter * (self.xmax - self.xmin)).astype(int)
        xbot = np.hstack(
            (
                np.linspace(self.xmin[0], self.xmax[0], num=nx, endpoint=False)[:, None],
                np.full([nx, 1], self.xmin[1]),
            )
        )
        yrig = np.hstack(
            (
                np.full([ny, 1], self.xmax[0]),
                np.linspace(self.xmin[1], self.xmax[1], num=ny, endpoint=False)[:, None],
            )
        )
        xtop = np.hstack(
            (
                np.linspace(self.xmin[0], self.xmax[0], num=nx + 1)[1:, None],
                np.full([nx, 1], self.xmax[1]),
            )
        )
        ylef = np.hstack(
            (
                np.full([ny, 1], self.xmin[0]),
                np.linspace(self.xmin[1], self.xmax[1], num=ny + 1)[1:, None],
            )
        )
        return np.vstack((xbot, yrig, xtop, ylef)).astype(config.real(np))

    def random_boundary_points(self, n, random="pseudo"):
        l1 = self.xmax[0] - self.xmin[0]
        l2 = l1 + self.xmax[1] - self.xmin[1]
        l3 = l2 + l1
        u = np.ravel(sample(n + 2, 1, random))
        u = u[np.logical_not(np.isclose(u, l1 / self.perimeter))]
        u = u[np.logical_not(np.isclose(u, l3 / self.perimeter))]
        u = u[:n]
        u *= self.perimeter
        x = []
        for l in u:
            if l < l1:
                x.append([self.xmin[0] + l, self.xmin[1]])
            elif l < l2:
                x.append([self.xmax[0], self.xmin[1] + l - l1])
            elif l < l3:
                x.append([self.xmax[0] - l + l2, self.xmax[1]])
            else:
                x.append([self.xmin[0], self.xmax[1] - l + l3])
        return np.vstack(x).astype(config.real(np))


class Cuboid(Hypercube):
    pass


class Disk(Geometry):
    def __init__(self, center, radius):
        self.center = np.array(center, dtype=config.real(np))
        self.radius = radius
        super().__init__(2, (self.center - radius, self.center + radius), 2 * radius)
        self._r2 = radius**2

    def inside(self, x):
        return np.linalg.norm(x - self.center, axis=-1) <= self.radius

    def on_boundary(self, x):
        return np.isclose(np.linalg.norm(x - self.center, axis=-1), self.radius)

    def boundary_normal(self, x):
        _n = x - self.center
        l = np.linalg.norm(_n, axis=-1, keepdims=True)
        _n = _n / l * np.isclose(l, self.radius)
        return _n

    def random_points(self, n, random="pseudo"):
        rng = sample(n, 2, random)
        r, theta = rng[:, 0], 2 * np.pi * rng[:, 1]
        x, y = np.cos(theta), np.sin(theta)
        return (self.radius * (np.sqrt(r) * np.vstack((x, y))).T + self.center).astype(
            config.real(np)
        )

    def uniform_boundary_points(self, n):
        theta = np.linspace(0, 2 * np.pi, num=n, endpoint=False)
        X = np.vstack((np.cos(theta), np.sin(theta))).T
        return (self.radius * X + self.center).astype(config.real(np))

    def random_boundary_points(self, n, random="pseudo"):
        u = sample(n, 1, random)
        theta = 2 * np.pi * u
        X = np.hstack((np.cos(theta), np.sin(theta)))
        return (self.radius * X + self.center).astype(config.real(np))


class Hypersphere(Geometry):
    """[DeepXDE-upstream `geometry_nd.Hypersphere`, restated; parity unpinned against the absent package]: ball of
    any dimension; 'pseudo' interior points are U^(1/dim) * normalised Gaussian vectors with U drawn first."""

    def __init__(self, center, radius):
        self.center = np.array(center, dtype=config.real(np))
        self.radius = radius
        super().__init__(len(center), (self.center - radius, self.center + radius), 2 * radius)
        self._r2 = radius**2

    def inside(self, x):
        return np.linalg.norm(x - self.center, axis=-1) <= self.radius

    def on_boundary(self, x):
        return np.isclose(np.linalg.norm(x - self.center, axis=-1), self.radius)

    def boundary_normal(self, x):
        _n = x - self.center
        l = np.linalg.norm(_n, axis=-1, keepdims=True)
        return _n / l * np.isclose(l, self.radius)

    def random_points(self, n, random="pseudo"):
        if random == "pseudo":
            U = np.random.rand(n, 1)
            X = np.random.normal(size=(n, self.dim))
        else:
            from scipy import stats

            rng = sample(n, self.dim + 1, random)
            U, X = rng[:, 0:1], stats.norm.ppf(rng[:, 1:])
        X = X / np.linalg.norm(X, axis=1, keepdims=True)
        X = U ** (1 / self.dim) * X
        return (self.radius * X + self.center).astype(config.real(np))

    def random_boundary_points(self, n, random="pseudo"):
        if random == "pseudo":
            X = np.random.normal(size=(n, self.dim))
        else:
            from scipy import stats

            X = stats.norm.ppf(sample(n, self.dim, random))
        X = X / np.linalg.norm(X, axis=1, keepdims=True)
        return (self.radius * X + self.center).astype(config.real(np))


class Sphere(Hypersphere):
    """Constructed by the reference at /root/reference/AI4XDE/cases/PDECases.py:2346 (`Sphere([0, 0, 0], 1)`)."""


class Polygon(Geometry):
    """Simple polygon given by its vertices in order (either orientation; stored clockwise as DeepXDE does) --
    `dde.geometry.Polygon([[0, 0], [1, 0], [1, -1], [-1, -1], [-1, 1], [0, 1]])` is the L-shaped domain of
    /root/reference/AI4XDE/cases/PoissonCases.py:515-516.  `inside` is the winding-number test, `on_boundary`
    the closed-segment test |x - v_i| + |x - v_{i+1}| = |v_{i+1} - v_i| (np.isclose), `boundary_normal` the
    outward unit normal of the segment a point lies on; points come from rejection sampling in the bounding
    box (interior) and from arc-length sampling of the perimeter (boundary)."""

    dim = 2

    def __init__(self, vertices):
        v = np.array(vertices, dtype=config.real(np))
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
        return (self.normal[first] * m.any(axis=1, keepdims=True)).astype(config.real(np))

    def random_points(self, n, random="pseudo"):
        out = np.empty((0, 2), dtype=config.real(np))
        lo, hi = self.bbox
        while len(out) < n:
            cand = sample(n, 2, "pseudo" if len(out) else random) * (hi - lo) + lo
            cand = cand[self.inside(cand) & ~self.on_boundary(cand)]
            out = np.vstack((out, cand.astype(config.real(np))))
        return out[:n]

    def _at_arclength(self, s):
        edges = np.concatenate(([0.0], np.cumsum(self.lengths)))
        i = np.clip(np.searchsorted(edges, s, side="right") - 1, 0, self.nvertices - 1)
        t = (s - edges[i]) / self.lengths[i]
        return (self.vertices[i] + t[:, None] * self.segments[i]).astype(config.real(np))

    def uniform_boundary_points(self, n):
        return self._at_arclength(np.linspace(0, self.perimeter, num=n, endpoint=False))

    def random_boundary_points(self, n, random="pseudo"):
        return self._at_arclength(np.ravel(sample(n, 1, random)) * self.perimeter)


class GeometryXTime:
    def __init__(self, geometry, timedomain):
        self.geometry = geometry
        self.timedomain = timedomain
        self.dim = geometry.dim + timedomain.dim

    def on_boundary(self, x):
        return self.geometry.on_boundary(x[:, :-1])

    def on_initial(self, x):
        return self.timedomain.on_initial(x[:, -1:])

    def boundary_normal(self, x):
        _n = self.geometry.boundary_normal(x[:, :-1])
        return np.hstack([_n, np.zeros((len(_n), 1))])

    def uniform_points(self, n, boundary=True):
        nx = int(
            np.ceil(
                (n * np.prod(self.geometry.bbox[1] - self.geometry.bbox[0]) / self.timedomain.diam)
                ** 0.5
            )
        )
        nt = int(np.ceil(n / nx))
        x = self.geometry.uniform_points(nx, boundary=boundary)
        nx = len(x)
        if boundary:
            t = self.timedomain.uniform_points(nt, boundary=True)
        else:
            t = np.linspace(
                self.timedomain.t1, self.timedomain.t0, num=nt, endpoint=False, dtype=config.real(np)
            )[:, None]
        xt = []
        for ti in t:
            xt.append(np.hstack((x, np.full([nx, 1], ti[0]))))
        return np.vstack(xt)

    def random_points(self, n, random="pseudo"):
        if isinstance(self.geometry, Interval):
            geom = Hypercube(
                [self.geometry.l, self.timedomain.t0], [self.geometry.r, self.timedomain.t1]
            )
            return geom.random_points(n, random=random)
        if isinstance(self.geometry, Hypercube):
            geom = Hypercube(
                np.append(self.geometry.xmin, self.timedomain.t0),
                np.append(self.geometry.xmax, self.timedomain.t1),
            )
            return geom.random_points(n, random=random)
        x = self.geometry.random_points(n, random=random)
        t = self.timedomain.random_points(n, random=random)
        t = np.random.permutation(t)
        return np.hstack((x, t))

    def uniform_boundary_points(self, n):
        if self.geometry.dim == 1:
            nx = 2
        else:
            s = 2 * sum(map(lambda l: l[0] * l[1], itertools.combinations(self.geometry.bbox[1] - self.geometry.bbox[0], 2)))
            nx = int((n * s / self.timedomain.diam) ** 0.5)
        nt = int(np.ceil(n / nx))
        x = self.geometry.uniform_boundary_points(nx)
        nx = len(x)
        t = np.linspace(
            self.timedomain.t1, self.timedomain.t0, num=nt, endpoint=False, dtype=config.real(np)
        )
        xt = []
        for ti in t:
            xt.append(np.hstack((x, np.full([nx, 1], ti))))
        return np.vstack(xt)

    def random_boundary_points(self, n, random="pseudo"):
        x = self.geometry.random_boundary_points(n, random=random)
        t = self.timedomain.random_points(n, random=random)
        t = np.random.permutation(t)
        return np.hstack((x, t))

    def uniform_initial_points(self, n):
        x = self.geometry.uniform_points(n, True)
        t = self.timedomain.t0
        return np.hstack((x, np.full([len(x), 1], t, dtype=config.real(np))))

    def random_initial_points(self, n, random="pseudo"):
        x = self.geometry.random_points(n, random=random)
        t = self.timedomain.t0
        return np.hstack((x, np.full([n, 1], t, dtype=config.real(np))))

    def periodic_point(self, x, component):
        xp = self.geometry.periodic_point(x[:, :-1], component)
        return np.hstack([xp, x[:, -1:]])
