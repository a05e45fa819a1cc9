"""Derives the cross-face neighbour table of the cube map (spec switch E11-alt, seamless filtering) from the reference's
face table alone (models/modules/NVDIFFREC/util.py:94-101, restated in oracle/raster_ref.py) with exact rational
arithmetic, and prints it as the constant that oracle/raster_ref.py (CUBE_WRAP) and texgs_common.cuh (CUBE_WRAP_TABLE)
hold. tests/test_oracle.py re-derives it and checks both copies.

For face f and side s (0: x < 0, 1: x >= R, 2: y < 0, 3: y >= R) a tap one texel outside the face, at position k
along the edge, is the texel of face f' that touches the same edge at the same place: (x', y') with each coordinate one
of 0, R-1, k, R-1-k (codes 0..3)."""
from fractions import Fraction as Fr


def cube_to_dir(f, x, y):
    return [(1, -y, -x), (-1, -y, x), (x, 1, y), (x, -1, -y), (x, -y, 1), (-x, -y, -1)][f]


def dir_to_face(d):
    x, y, z = d
    ax, ay, az = abs(x), abs(y), abs(z)
    if ax >= ay and ax >= az:
        f, sx, sy, m = (1 if x < 0 else 0), (z if x < 0 else -z), -y, ax
    elif ay >= az:
        f, sx, sy, m = (3 if y < 0 else 2), x, (-z if y < 0 else z), ay
    else:
        f, sx, sy, m = (5 if z < 0 else 4), (-x if z < 0 else x), -y, az
    return f, sx / m, sy / m


def derive(R=8):
    table = []
    for f in range(6):
        row = []
        for side in range(4):
            pts = []
            for k in (1, R - 2):
                x, y = {0: (-1, k), 1: (R, k), 2: (k, -1), 3: (k, R)}[side]
                sx, sy = Fr(2 * x + 1, R) - 1, Fr(2 * y + 1, R) - 1
                nf, sxn, syn = dir_to_face(cube_to_dir(f, sx, sy))
                xi, yi = int((sxn + 1) * R / 2), int((syn + 1) * R / 2)      # floor: both are positive
                pts.append((nf, min(xi, R - 1), min(yi, R - 1)))
            (nf, xa, ya), (nf2, xb, yb) = pts
            assert nf == nf2 != f

            def code(a, b):
                for c, (va, vb) in enumerate(((0, 0), (R - 1, R - 1), (1, R - 2), (R - 2, 1))):
                    if (a, b) == (va, vb):
                        return c
                raise AssertionError((a, b))
            row.append((nf, code(xa, xb), code(ya, yb)))
        table.append(row)
    return table


if __name__ == "__main__":
    t = derive()
    assert t == derive(32)
    print("CUBE_WRAP =", t)
    print("{" + ", ".join("{" + ", ".join("{%d, %d, %d}" % e for e in row) + "}" for row in t) + "}")
