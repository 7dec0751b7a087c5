"""A minimal stand-in for `affine.Affine` (what rasterio.transform.from_origin returns at
neilpy/neilpy.py:1141): six coefficients, `~t`, `t * (x, y)`, indexing, and the 9-tuple
protocol rasterio accepts as `transform=`.  When the real `affine` package is importable
the public API returns a genuine `affine.Affine` instead (see api._make_transform).

The arithmetic of __invert__ and __mul__ is the `affine` package's, operation for
operation: the binning kernel receives the six inverse coefficients computed here and
applies them in the same order, which is what makes the cell index of every point
bit-identical to the reference.
"""
from __future__ import annotations

from collections import namedtuple


class Affine(namedtuple('Affine', ('a', 'b', 'c', 'd', 'e', 'f', 'g', 'h', 'i'))):
    __slots__ = ()

    def __new__(cls, a, b, c, d, e, f, g=0.0, h=0.0, i=1.0):
        return super().__new__(cls, float(a), float(b), float(c), float(d), float(e), float(f),
                               float(g), float(h), float(i))

    @classmethod
    def from_origin(cls, west, north, xsize, ysize):
        """rasterio.transform.from_origin = Affine.translation(west, north) * Affine.scale(xsize, -ysize)."""
        return cls(xsize, 0.0, west, 0.0, -ysize, north)

    @property
    def coeffs(self):
        return tuple(self[:6])

    def to_gdal(self):
        return (self.c, self.a, self.b, self.f, self.d, self.e)

    def __invert__(self):
        a, b, c, d, e, f = self[:6]
        idet = 1.0 / (a * e - b * d)
        ra = e * idet
        rb = -b * idet
        rd = -d * idet
        re = a * idet
        return Affine(ra, rb, -c * ra - f * rb, rd, re, -c * rd - f * re)

    def __mul__(self, other):
        sa, sb, sc, sd, se, sf = self[:6]
        if isinstance(other, Affine):
            oa, ob, oc, od, oe, of = other[:6]
            return Affine(sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
                          sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)
        vx, vy = other
        return (vx * sa + vy * sb + sc, vx * sd + vy * se + sf)

    def __repr__(self):
        return 'Affine(%r, %r, %r,\n       %r, %r, %r)' % tuple(self[:6])
