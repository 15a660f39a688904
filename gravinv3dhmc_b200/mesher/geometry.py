"""Cell types handed out by the meshes (same attribute names as the reference's
mesher/geometry.py:71-78 Prism and :117-125 Tesseroid)."""
import numpy as np


class GeometricElement:
    def __init__(self, props=None):
        self.props = {} if props is None else dict(props)

    def addprop(self, prop, value):
        self.props[prop] = value


class Prism(GeometricElement):
    """Right rectangular prism: x1,x2 (north), y1,y2 (east), z1,z2 (down)."""

    def __init__(self, x1, x2, y1, y2, z1, z2, props=None):
        super().__init__(props)
        self.x1, self.x2 = float(x1), float(x2)
        self.y1, self.y2 = float(y1), float(y2)
        self.z1, self.z2 = float(z1), float(z2)

    def get_bounds(self):
        return [self.x1, self.x2, self.y1, self.y2, self.z1, self.z2]

    def center(self):
        return np.array([0.5 * (self.x1 + self.x2), 0.5 * (self.y1 + self.y2),
                         0.5 * (self.z1 + self.z2)])

    def __str__(self):
        names = [("x1", self.x1), ("x2", self.x2), ("y1", self.y1), ("y2", self.y2),
                 ("z1", self.z1), ("z2", self.z2)]
        names.extend((p, self.props[p]) for p in sorted(self.props))
        return " | ".join("%s:%g" % (n, v) for n, v in names)


class Tesseroid(GeometricElement):
    """Spherical prism: w,e,s,n in degrees, top/bottom heights (m) above the mean earth radius."""

    def __init__(self, w, e, s, n, top, bottom, props=None):
        super().__init__(props)
        self.w, self.e = float(w), float(e)
        self.s, self.n = float(s), float(n)
        self.top, self.bottom = float(top), float(bottom)

    def get_bounds(self):
        return [self.w, self.e, self.s, self.n, self.top, self.bottom]

    def __str__(self):
        names = [("w", self.w), ("e", self.e), ("s", self.s), ("n", self.n), ("top", self.top),
                 ("bottom", self.bottom)]
        names.extend((p, self.props[p]) for p in sorted(self.props))
        return " | ".join("%s:%g" % (n, v) for n, v in names)
