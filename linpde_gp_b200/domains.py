"""Domains of the problem set-up code around the hot path (``linpde_gp.domains``): intervals, points, boxes and
Cartesian products with their boundaries and uniform grids.

These classes are host-side set-up only (``domains/``: O(N) work); they exist so that the
reference's experiment and test scripts (``experiments/0000 - 0003``, ``tests/linpde_gp/problems/test_heat.py``) run
against this package with nothing but their imports changed.  Behaviour follows src/linpde_gp/domains/_domain.py:13-55,
_interval.py:14-83, _point.py:11-51, _box.py:17-114, _cartesian_product.py:12-127, _asdomain.py:10-22; a ``Box``'s uniform
grid is a ``TensorProductGrid`` (``_box.py:81-114``), which is what routes gridded collocation points to the
Kronecker assembly kernel."""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np

from .functions import _as_shape


class Domain:
    def __init__(self, shape, dtype=np.double):
        if not np.issubdtype(dtype, np.floating):
            raise TypeError("The dtype of a domain must be a sub dtype of `np.floating`")
        self._shape = _as_shape(shape)
        self._dtype = np.dtype(dtype)

    dtype = property(lambda self: self._dtype)
    shape = property(lambda self: self._shape)
    ndims = property(lambda self: len(self._shape))
    size = property(lambda self: int(np.prod(self._shape)) if self._shape else 1)

    @property
    def boundary(self):  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def volume(self):  # pragma: no cover - abstract
        raise NotImplementedError

    __hash__ = None


class Point(Domain):
    """A single point (the boundary parts of an interval; a collapsed factor of a box)."""

    def __init__(self, point):
        self._point = np.asarray(point, dtype=np.double)
        super().__init__(self._point.shape, self._point.dtype)

    @property
    def boundary(self):
        return (self,)

    @property
    def volume(self):
        return np.zeros((), dtype=self.dtype)

    def __repr__(self):
        return f"<Point {self._point} with shape={self.shape} and dtype={self.dtype}>"

    def __array__(self, dtype=None, copy=None):
        return np.array(self._point, dtype=dtype, copy=True)

    def __float__(self):
        if self.ndims > 1:
            raise NotImplementedError()
        return float(self._point)

    def __contains__(self, item):
        arr = np.asarray(item, dtype=self.dtype)
        return arr.shape == self.shape and bool(np.all(self._point == arr))

    def __eq__(self, other):
        return isinstance(other, Point) and self.shape == other.shape and bool(np.all(self._point == other._point))


class Interval(Domain, Sequence):
    """Closed interval ``[lower_bound, upper_bound]``; iterating yields the two bounds."""

    def __init__(self, lower_bound, upper_bound, dtype=np.double):
        self._lower_bound = np.dtype(dtype).type(lower_bound)
        self._upper_bound = np.dtype(dtype).type(upper_bound)
        if self._lower_bound > self._upper_bound:
            raise ValueError("The lower bound must not be larger than the upper bound")
        super().__init__((), dtype)

    def __len__(self):
        return 2

    def __getitem__(self, idx):
        if idx in (0, -2):
            return self._lower_bound
        if idx in (1, -1):
            return self._upper_bound
        raise IndexError(f"Index {idx} is out of range")

    def __iter__(self):
        yield self._lower_bound
        yield self._upper_bound

    @property
    def boundary(self):
        return (Point(self._lower_bound), Point(self._upper_bound))

    @property
    def volume(self):
        return self._upper_bound - self._lower_bound

    def __repr__(self):
        return f"<Interval {[self._lower_bound, self._upper_bound]} with shape={self.shape} and dtype={self.dtype}>"

    def __contains__(self, item):
        arr = np.asarray(item, dtype=self.dtype)
        return arr.shape == self.shape and bool(self._lower_bound <= arr <= self._upper_bound)

    def __eq__(self, other):
        return isinstance(other, Interval) and tuple(self) == tuple(other)

    def uniform_grid(self, shape, inset=0.0) -> np.ndarray:
        shape, inset = _as_shape(shape), np.asarray(inset)
        if len(shape) != 1 or inset.ndim != 0:
            raise ValueError("an interval takes a single number of grid points and a scalar inset")
        return np.linspace(self._lower_bound + inset, self._upper_bound - inset, shape[0])


class CartesianProduct(Domain):
    """Product of scalar / vector domains; its boundary replaces one factor at a time by a part of that factor's
    boundary (``_cartesian_product.py:75-86``)."""

    def __init__(self, *domains):
        self._domains = tuple(asdomain(d) for d in domains)
        if any(d.ndims > 1 for d in self._domains):
            raise ValueError("factors must be scalar or vector domains")
        dtype = self._domains[0].dtype if self._domains else np.dtype(np.double)
        if any(d.dtype != dtype for d in self._domains):
            raise ValueError("all factors must share one dtype")
        super().__init__((sum(d.shape[0] if d.ndims == 1 else 1 for d in self._domains),), dtype)

    factors = property(lambda self: self._domains)

    def _as_box(self):
        if not all(isinstance(d, (Interval, Box, Point)) for d in self._domains):
            return None
        bounds = []
        for d in self._domains:
            if isinstance(d, Interval):
                bounds.append(tuple(d))
            elif isinstance(d, Box):
                bounds.extend(tuple(iv) for iv in d.bounds)
            else:
                bounds.extend((c, c) for c in np.atleast_1d(np.asarray(d)))
        return Box(np.asarray(bounds, dtype=self.dtype))

    @property
    def boundary(self):
        return tuple(CartesianProduct(*self._domains[:i], part, *self._domains[i + 1:])
                     for i, factor in enumerate(self._domains) for part in factor.boundary)

    @property
    def volume(self):
        out = np.ones((), dtype=self.dtype)
        for d in self._domains:
            out = out * d.volume
        return out

    def __contains__(self, item):
        box = self._as_box()
        if box is None:
            raise NotImplementedError
        return item in box

    def __eq__(self, other):
        return (isinstance(other, CartesianProduct) and len(self) == len(other)
                and all(a == b for a, b in zip(self._domains, other._domains)))

    def __len__(self):
        return len(self._domains)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return self._domains[idx]
        return CartesianProduct(*self._domains[idx])

    def __iter__(self):
        return iter(self._domains)

    def __repr__(self):
        inner = "".join(f"\n  - {d!r}" for d in self._domains)
        return f"<CartesianProduct of{inner}\nwith shape={self.shape} and dtype={self.dtype}>"

    def uniform_grid(self, shape, inset=0.0):
        box = self._as_box()
        if box is None:
            raise NotImplementedError
        return box.uniform_grid(shape, inset=inset)


class Box(CartesianProduct):
    """Axis-aligned box given by ``bounds`` of shape ``(D, 2)``; collapsed axes (lower == upper) are points."""

    def __init__(self, bounds):
        bounds = np.array(bounds, copy=True)
        if not (bounds.ndim == 2 and bounds.shape[-1] == 2):
            raise ValueError(f"`bounds` must have shape (D, 2), but an object of shape {bounds.shape} was given.")
        if not np.issubdtype(bounds.dtype, np.floating):
            raise TypeError(f"The dtype of `bounds` must be a sub dtype of `np.floating`, but {bounds.dtype} was given.")
        if not np.all(bounds[:, 0] <= bounds[:, 1]):
            raise ValueError("The lower bounds must not be larger than the upper bounds.")
        bounds.flags.writeable = False
        self._bounds = bounds
        self._interior_idcs = np.nonzero(bounds[:, 0] != bounds[:, 1])[0]
        super().__init__(*(Interval(lo, hi, dtype=bounds.dtype) if lo != hi else Point(lo) for lo, hi in bounds))

    bounds = property(lambda self: self._bounds)

    def _as_box(self):
        return self

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return super().__getitem__(idx)
        return Box(self._bounds[idx, :])

    def __repr__(self):
        return (f"<Box {' x '.join(str(list(b)) for b in self._bounds)} with shape={self.shape} and "
                f"dtype={self.dtype}>")

    def __contains__(self, item):
        arr = np.asarray(item, dtype=self.dtype)
        return arr.shape == self.shape and bool(np.all((self._bounds[:, 0] <= arr) & (arr <= self._bounds[:, 1])))

    def __eq__(self, other):
        if isinstance(other, Box):
            return self._bounds.shape == other._bounds.shape and bool(np.all(self._bounds == other._bounds))
        return CartesianProduct.__eq__(self, other)

    def uniform_grid(self, shape, inset=0.0):
        """``TensorProductGrid`` with ``shape`` points along the non-collapsed axes (one point on collapsed ones), each
        axis shrunk by ``inset`` at both ends."""
        from .randprocs.covfuncs import TensorProductGrid  # pylint: disable=import-outside-toplevel

        n_int = len(self._interior_idcs)
        shape = _as_shape(shape)
        if len(shape) == 1 and n_int > 1:
            shape = shape * n_int
        if len(shape) != n_int:
            raise ValueError(f"expected {n_int} grid sizes, got {len(shape)}")
        counts = np.ones(len(self._bounds), dtype=int)
        insets = np.zeros(len(self._bounds))
        counts[self._interior_idcs] = shape
        insets[self._interior_idcs] = np.broadcast_to(inset, n_int)
        axes = []
        for (lo, hi), n, ins in zip(self._bounds, counts, insets):
            axes.append(np.linspace(lo + ins, hi - ins, int(n)) if lo != hi else np.array([lo]))
        return TensorProductGrid(*axes, indexing="ij")


def asdomain(arg) -> Domain:
    """Domains pass through; ``(a, b)`` with scalars is an interval, with two equal-length vectors a box."""
    if isinstance(arg, Domain):
        return arg
    if isinstance(arg, (Sequence, np.ndarray)) and len(arg) == 2:
        if all(np.ndim(b) == 0 for b in arg):
            return Interval(float(arg[0]), float(arg[1]))
        if np.ndim(arg[0]) == 1 and np.shape(arg[0]) == np.shape(arg[1]):
            return Box(np.stack([np.asarray(arg[0], dtype=np.double), np.asarray(arg[1], dtype=np.double)], axis=1))
    raise ValueError(f"Could not convert {arg} to a domain")
