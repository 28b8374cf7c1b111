"""The detect tail's labelling as an executable specification, on the CPU.

`oat_b200/csrc/tail_fast.cuh` obtains the external contours of a mask and their polygon moments without tracing
borders: union-find over row RUNS (8-connected foreground runs, 4-connected "candidate" background runs between a
row's first and last foreground pixel; background that reaches the row-exterior is linked to node 0), holes join the
foreground beside them, and every 2x2 cell of pixel centres with 4 (3) pixels inside adds area 1 (1/2) -- summed as
exact integers 2*m00, 6*m10, 6*m01.  For busy masks the work is split (band_prelabel / tail_label_phase): every BAND
of rows builds the run table of its own rows, merges inside the band, and sums the moments of the runs whose cells lie
inside the band, per run and per band-local root; the labelling CTA then merges only the first and last row of every
band and takes the bands' sums -- except the last row of a band, the filled holes, and the runs beside or above a
filled hole, whose band sums are replaced.

This file transcribes that split in plain Python and holds it, on random masks with rings, nested shapes, U shapes,
worms and speckle, against the oracle's border-following contours + Green's-theorem moments (which
tests/test_oracle_vs_cv2.py pins to cv2.findContours / cv2.moments).  The CUDA code itself is held to the same oracle,
and to cv2, by tests/test_gpu_resident.py::test_prelabelled_bands_equal_the_synchronous_tail on the GPU."""
import numpy as np
import pytest

import oracle


class _UF:
    """parents only ever decrease (tail_fast.cuh: suf_find / suf_union / suf_link_up)"""

    def __init__(self, n):
        self.p = list(range(n))

    def find(self, x):
        while self.p[x] != x:
            self.p[x] = self.p[self.p[x]]
            x = self.p[x]
        return x

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a != b:
            self.p[max(a, b)] = min(a, b)


def _row_runs(row, inner):
    """foreground runs and candidate background runs (inner rows only: background between the first and last foreground pixel)"""
    xs = np.flatnonzero(row)
    if xs.size == 0:
        return [], [], None
    lo, hi = int(xs[0]), int(xs[-1])
    fg, bg = [], []
    x = lo
    while x <= hi:
        v = row[x]
        e = x
        while e + 1 <= hi and row[e + 1] == v:
            e += 1
        (fg if v else bg).append((x, e))
        x = e + 1
    return fg, (bg if inner else []), (lo, hi)


def _cell_sums(G, y, s, e):
    """2*m00, 6*m10, 6*m01 of the cells owned by pixels [s, e] of row y (run_cell_sums): a cell belongs to its top-left pixel,
    or to its top-right one when the top-left is background; the last row owns nothing"""
    rows, cols = G.shape
    if y >= rows - 1:
        return 0, 0, 0

    def g(yy, xx):
        return bool(G[yy, xx]) if 0 <= xx < cols and 0 <= yy < rows else False

    t00 = t10 = t01 = 0
    for x in range(s, e + 1):
        tl, tr, bl, br = g(y, x), g(y, x + 1), g(y + 1, x), g(y + 1, x + 1)
        assert tl
        if tr and bl and br:
            t00, t10, t01 = t00 + 2, t10 + 6 * x + 3, t01 + 6 * y + 3
        elif (not tr) and bl and br:
            t00, t10, t01 = t00 + 1, t10 + 3 * x + 1, t01 + 3 * y + 2
        elif tr and (not bl) and br:
            t00, t10, t01 = t00 + 1, t10 + 3 * x + 2, t01 + 3 * y + 1
        elif tr and bl and (not br):
            t00, t10, t01 = t00 + 1, t10 + 3 * x + 1, t01 + 3 * y + 1
        # owner = top-right pixel of the cell to its left, whose top-left is background
        if not g(y, x - 1) and g(y + 1, x - 1) and g(y + 1, x):
            t00, t10, t01 = t00 + 1, t10 + 3 * x - 1, t01 + 3 * y + 2
    return t00, t10, t01


def prelabelled_contours(mask, R):
    """-> {first pixel index: (2*m00, 6*m10, 6*m01)} per external contour, computed the way the bands + the labelling CTA do"""
    rows, cols = mask.shape
    M = mask.astype(bool)
    per_row = [_row_runs(M[y], 0 < y < rows - 1) for y in range(rows)]
    ext = [r[2] for r in per_row]
    # global table: 0 = exterior, foreground runs in raster order, then the candidates in raster order
    F = [(y, s, e) for y in range(rows) for (s, e) in per_row[y][0]]
    B = [(y, s, e) for y in range(rows) for (s, e) in per_row[y][1]]
    nF, nB = len(F), len(B)
    fid = {}
    for i, (y, s, e) in enumerate(F):
        fid.setdefault(y, []).append(1 + i)
    bid = {}
    for i, (y, s, e) in enumerate(B):
        bid.setdefault(y, []).append(1 + nF + i)
    run = {1 + i: r for i, r in enumerate(F)}
    run.update({1 + nF + i: r for i, r in enumerate(B)})

    def merge_up(uf, i, allowed_rows):
        """tail_label_phase step 4 / the band's merges for run i against row y - 1 (if that row is in allowed_rows), and
        the exterior test against the neighbour rows in allowed_rows"""
        y, s, e = run[i]
        if i <= nF:
            if y - 1 in allowed_rows:
                for k in fid.get(y - 1, []):
                    if run[k][2] >= s - 1 and run[k][1] <= e + 1:
                        uf.union(i, k)
            return
        if y - 1 in allowed_rows:
            for k in bid.get(y - 1, []):
                if run[k][2] >= s and run[k][1] <= e:
                    uf.union(i, k)
        for yy in (y - 1, y + 1):
            if yy not in allowed_rows:
                continue
            if yy == 0 or yy == rows - 1:
                isext = not M[yy, s:e + 1].all()
            else:
                isext = ext[yy] is None or s < ext[yy][0] or e > ext[yy][1]
            if isext:
                uf.union(i, 0)

    # ---- bands: merges inside the band, sums of the runs whose cells lie inside the band ----
    uf = _UF(1 + nF + nB)
    band_sum = {}   # foreground run -> its own sums (hole filling not known yet)
    nbands = (rows + R - 1) // R
    for b in range(nbands):
        y0, y1 = b * R, min(b * R + R, rows) - 1
        inside = set(range(y0, y1 + 1))
        ids = [i for y in range(y0, y1 + 1) for i in fid.get(y, []) + bid.get(y, [])]
        for i in ids:
            merge_up(uf, i, inside)
        for i in ids:
            y, s, e = run[i]
            if i <= nF and y < y1:
                band_sum[i] = _cell_sums(M, y, s, e)
    local_root = {i: uf.find(i) for i in range(1, 1 + nF + nB)}
    agg = {}
    for i, t in band_sum.items():
        a = agg.setdefault(local_root[i], [0, 0, 0])
        for k in range(3):
            a[k] += t[k]
    # (a band-local root lies in the band of its runs: nothing has been merged across a border yet)
    assert all((run[r][0] // R) == (run[i][0] // R) for i, r in local_root.items() if r != 0)

    # ---- labelling CTA: band borders only ----
    every = set(range(rows))
    for b in range(nbands):
        for y in {b * R, min(b * R + R, rows) - 1}:
            for i in fid.get(y, []) + bid.get(y, []):
                merge_up(uf, i, every)
    # holes join the foreground beside them; the runs whose cells they change are marked
    G = M.copy()
    dirty, holes = set(), []
    for i in range(1 + nF, 1 + nF + nB):
        if uf.find(i) == 0:
            continue
        y, s, e = run[i]
        holes.append(i)
        for k in fid[y]:
            if run[k][2] == s - 1 or run[k][1] == e + 1:
                uf.union(i, k)
                dirty.add(k)
        for k in fid.get(y - 1, []):
            if run[k][2] >= s - 1 and run[k][1] <= e + 1:
                dirty.add(k)
        G[y, s:e + 1] = True
    acc = {}

    def add(i, t):
        a = acc.setdefault(uf.find(i), [0, 0, 0])
        for k in range(3):
            a[k] += t[k]

    for i in range(1, 1 + nF):
        acc.setdefault(uf.find(i), [0, 0, 0])
        y, s, e = run[i]
        last_row_of_band = (y % R == R - 1) or (y == rows - 1)
        if last_row_of_band:                       # (a) its cells reach into the next band
            add(i, _cell_sums(G, y, s, e))
        elif i in dirty:                           # (b) a filled hole changed them: replace the band's sums
            new, old = _cell_sums(G, y, s, e), band_sum[i]
            add(i, tuple(n - o for n, o in zip(new, old)))
    for i in holes:                                # (b) the filled holes themselves
        y, s, e = run[i]
        add(i, _cell_sums(G, y, s, e))
    for r, t in agg.items():                       # (c) the bands' sums, one set per band-local root
        add(r, tuple(t))
    out = {}
    for r, t in acc.items():
        assert 1 <= r <= nF
        y, s, _ = run[r]
        out[y * cols + s] = tuple(t)
    return out


def _scene(rng, rows, cols):
    m = np.zeros((rows, cols), bool)
    yy, xx = np.mgrid[0:rows, 0:cols]
    for _ in range(rng.integers(1, 5)):      # discs, rings, rings with a disc inside
        cy, cx, r = rng.integers(0, rows), rng.integers(0, cols), rng.integers(2, max(3, rows // 3))
        d2 = (yy - cy) ** 2 + (xx - cx) ** 2
        kind = rng.integers(3)
        if kind == 0:
            m |= d2 <= r * r
        else:
            m |= (d2 <= r * r) & (d2 >= (r - rng.integers(1, 3)) ** 2)
            if kind == 2:
                m |= d2 <= (r // 3) ** 2
    for _ in range(rng.integers(0, 3)):      # U shapes
        y0, x0 = rng.integers(0, rows - 4), rng.integers(0, cols - 4)
        y1, x1 = min(rows - 1, y0 + rng.integers(3, rows // 2)), min(cols - 1, x0 + rng.integers(3, cols // 2))
        box = np.zeros_like(m)
        box[y0:y1 + 1, x0:x1 + 1] = True
        box[y0 + 1:y1, x0 + 1:x1] = False
        side = rng.integers(4)
        if side == 0:
            box[y0, x0 + 1:x1] = False
        elif side == 1:
            box[y1, x0 + 1:x1] = False
        elif side == 2:
            box[y0 + 1:y1, x0] = False
        else:
            box[y0 + 1:y1, x1] = False
        m |= box
    for _ in range(rng.integers(0, 3)):      # worms
        y, x = int(rng.integers(0, rows)), int(rng.integers(0, cols))
        for _ in range(rows):
            m[y, x] = True
            y = int(np.clip(y + rng.integers(-1, 2), 0, rows - 1))
            x = int(np.clip(x + rng.integers(-1, 2), 0, cols - 1))
    if rng.integers(2):                      # speckle
        m |= rng.random(m.shape) < rng.choice([0.02, 0.1, 0.4])
    if rng.integers(3) == 0:                 # something on the frame border
        m[0, cols // 4:cols // 2] = True
        m[rows - 1, cols // 3:] = True
        m[rows // 4:, 0] = True
        m[:rows // 2, cols - 1] = True
    return m


@pytest.mark.parametrize("shape,R", [((24, 37), 4), ((40, 64), 8), ((33, 50), 32), ((50, 41), 5), ((64, 96), 8)])
def test_prelabelled_labelling_model_matches_border_following(shape, R):
    rows, cols = shape
    rng = np.random.default_rng(rows * 1000 + cols + R)
    for trial in range(25):
        m = _scene(rng, rows, cols)
        if not m.any():
            continue
        got = prelabelled_contours(m, R)
        want = {fi: (m00, m10, m01) for (fi, npts, m00, m10, m01) in oracle.external_contours(m.astype(np.uint8) * 255)}
        assert sorted(got) == sorted(want), (trial, sorted(set(got) ^ set(want))[:5])
        for fi, (t00, t10, t01) in got.items():
            m00, m10, m01 = want[fi]
            assert abs(t00 / 2 - m00) < 1e-9 and abs(t10 / 6 - m10) < 1e-6 and abs(t01 / 6 - m01) < 1e-6, (trial, fi, (t00, t10, t01), want[fi])
