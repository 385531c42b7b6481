"""Minimal stand-in for `shapely.geometry.Polygon` (shapely is not installed in this image).

TEST INFRASTRUCTURE ONLY.  Used solely to let the *unmodified* reference modules import and run
in the build container (oracle/gen_golden.py).  It supports exactly what the reference's hot
path uses on convex quads: Polygon(points).buffer(0), .area, .intersection(other).area
(src/utils/iou_rotated_boxes_utils.py:24-31,91,118-120).  Arithmetic is fp64 like GEOS; results
labelled "stand-in" wherever they are reported.  Pure Python, deliberately independent of
oracle/rbox_oracle.c so the two can check each other.
"""
import sys
import types


class Polygon:
    def __init__(self, pts=()):
        self.pts = [(float(x), float(y)) for x, y in pts]

    def buffer(self, _d):
        return self

    @property
    def area(self):
        p = self.pts
        if len(p) < 3:
            return 0.0
        s = 0.0
        for i in range(len(p)):
            x0, y0 = p[i]
            x1, y1 = p[(i + 1) % len(p)]
            s += x0 * y1 - x1 * y0
        return abs(s) * 0.5

    def _signed(self):
        p = self.pts
        return sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p)))

    def intersection(self, other):
        subj = list(self.pts)
        clip = list(other.pts)
        if len(subj) < 3 or len(clip) < 3:
            return Polygon()
        if other._signed() < 0:
            clip = clip[::-1]
        for k in range(len(clip)):
            ax, ay = clip[k]
            bx, by = clip[(k + 1) % len(clip)]
            out = []

            def side(p):
                return (bx - ax) * (p[1] - ay) - (by - ay) * (p[0] - ax)

            for i in range(len(subj)):
                s, t = subj[i], subj[(i + 1) % len(subj)]
                ds, dt = side(s), side(t)
                if ds >= 0:
                    out.append(s)
                if (ds > 0 > dt) or (ds < 0 < dt):
                    u = ds / (ds - dt)
                    out.append((s[0] + u * (t[0] - s[0]), s[1] + u * (t[1] - s[1])))
            subj = out
            if not subj:
                break
        return Polygon(subj)


def install():
    """Register the stand-in as `shapely` / `shapely.geometry` unless real shapely imports."""
    try:
        import shapely.geometry  # noqa: F401
        return "shapely"
    except Exception:
        pass
    m = types.ModuleType("shapely")
    g = types.ModuleType("shapely.geometry")
    g.Polygon = Polygon
    m.geometry = g
    sys.modules["shapely"] = m
    sys.modules["shapely.geometry"] = g
    return "stand-in"
