"""CPU check of the arithmetic argument behind the tile kernels' pair classification (csrc/raster.cuh: FaceTile,
classify_fast): the three edge functions are evaluated in affine form with two FMAs, and only pairs with
min_i |t_i| < eps = 1e-6 max_i (|a_i| (1 + |x_anchor|) + |b_i| (1 + |y_anchor|)) are re-classified with PyTorch3D's bit-exact
unfused EdgeFunctionForward.  Emulated here in float32 (FMA = exact product and sum in float64, one rounding) on millions of
random (pixel, face) pairs, including faces far off screen and pixels within fractions of an ulp-scale distance of an edge:
  * outside the ambiguous band the affine value has the sign of the exact edge function (strict-inside is bit-identical);
  * |t_affine - t_exact| stays below eps / 2;
  * a pair the affine form calls saturated (u >= 17.5 sigma) is saturated exactly (1 - sigmoid(d/sigma) == 0 in fp32 needs
    d/sigma >= 16.64), and a pair it rejects (u >= 1.02 blur on a violated edge) lies at least blur away from that edge line.
"""
import numpy as np

f32 = np.float32
SIGMA = f32(1e-7)
BLUR = f32(9.210240366975849e-7)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def _edge_exact(px, py, ax, ay, bx, by):
    """EdgeFunctionForward with every operation rounded to float32 (raster.cuh: edge_fn)."""
    return ((px - ax) * (by - ay)).astype(f32) - ((py - ay) * (bx - ax)).astype(f32)


def _cases(rng, n, spread):
    """n random triangles of a few pixels (512^2 image: pixel pitch 2/512) around centres in [-spread, spread]^2 and, per
    triangle, one pixel placed near an edge at a log-uniform distance (down to 1e-9 NDC) or anywhere in the bbox."""
    c = rng.uniform(-spread, spread, size=(n, 1, 2))
    size = 10 ** rng.uniform(-3.0, -1.0, size=(n, 1, 1))
    v = (c + size * rng.normal(size=(n, 3, 2))).astype(f32)
    k = rng.integers(0, 3, size=n)
    a, b = v[np.arange(n), k], v[np.arange(n), (k + 1) % 3]
    t = rng.uniform(-0.2, 1.2, size=(n, 1))
    on_edge = a + t * (b - a)
    nrm = np.stack([-(b - a)[:, 1], (b - a)[:, 0]], 1)
    nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    dist = (10 ** rng.uniform(-9.0, -2.0, size=(n, 1))) * rng.choice([-1.0, 1.0], size=(n, 1))
    p_near = on_edge + dist * nrm
    p_any = c[:, 0] + size[:, 0] * rng.normal(size=(n, 2))
    p = np.where(rng.random((n, 1)) < 0.7, p_near, p_any).astype(f32)
    return v, p


def test_affine_edge_functions_bound_and_decisions():
    rng = np.random.default_rng(0)
    total = amb_total = 0
    for spread in (0.9, 1.0, 3.0, 8.0):
        v, p = _cases(rng, 400_000, spread)
        x0, y0, x1, y1, x2, y2 = v[:, 0, 0], v[:, 0, 1], v[:, 1, 0], v[:, 1, 1], v[:, 2, 0], v[:, 2, 1]
        px, py = p[:, 0], p[:, 1]
        area = _edge_exact(x2, y2, x0, y0, x1, y1) + f32(1e-8)
        keep = np.abs(area) > 1e-8
        sg = np.where(area > 0, f32(1), f32(-1))
        # stage_face: deltas rounded as edge_fn rounds them; coefficients; eps
        d12x, d12y = x2 - x1, y2 - y1
        d20x, d20y = x0 - x2, y0 - y2
        d01x, d01y = x1 - x0, y1 - y0
        coef = []
        for (dx, dy, ax, ay) in ((d12x, d12y, x1, y1), (d20x, d20y, x2, y2), (d01x, d01y, x0, y0)):
            a_, b_ = sg * dy, -sg * dx
            c_ = sg * ((ay * dx).astype(f32) - (ax * dy).astype(f32)).astype(f32)  # worst case: no FMA in the staging
            m_ = np.abs(dy) * (1 + np.abs(ax)) + np.abs(dx) * (1 + np.abs(ay))
            il = (f32(1) / (dx * dx + dy * dy)).astype(f32)
            coef.append((a_, b_, c_, m_.astype(f32), il, dx, dy))
        eps = f32(1e-6) * np.maximum(np.maximum(coef[0][3], coef[1][3]), coef[2][3])
        t_fast = [_fma(a_, px, _fma(b_, py, c_)) for (a_, b_, c_, _, _, _, _) in coef]
        t_exact = [sg * _edge_exact(px, py, x1, y1, x2, y2), sg * _edge_exact(px, py, x2, y2, x0, y0),
                   sg * _edge_exact(px, py, x0, y0, x1, y1)]
        ambiguous = np.minimum(np.minimum(np.abs(t_fast[0]), np.abs(t_fast[1])), np.abs(t_fast[2])) < eps
        ok = keep & ~ambiguous
        total += int(ok.sum()); amb_total += int((keep & ambiguous).sum())
        for tf, te in zip(t_fast, t_exact):
            assert np.all(np.abs(tf - te)[keep] <= 0.5 * eps[keep]), float(np.max((np.abs(tf - te) / eps)[keep]))
            assert np.all((tf > 0)[ok] == (te > 0)[ok])           # strict-inside decided identically
        # saturation / rejection thresholds against the exact line distances (float64 geometry as the arbiter)
        inside = (t_fast[0] > 0) & (t_fast[1] > 0) & (t_fast[2] > 0)
        u = [tf * tf * il for tf, (_, _, _, _, il, _, _) in zip(t_fast, coef)]
        sat = inside & (u[0] >= f32(17.5) * SIGMA) & (u[1] >= f32(17.5) * SIGMA) & (u[2] >= f32(17.5) * SIGMA) & ok
        far = f32(1.02) * BLUR
        d2_exact = []
        for te, (_, _, _, _, _, dx, dy) in zip(t_exact, coef):
            d2_exact.append(te.astype(np.float64) ** 2 / (dx.astype(np.float64) ** 2 + dy.astype(np.float64) ** 2))
        dmin = np.minimum(np.minimum(d2_exact[0], d2_exact[1]), d2_exact[2])
        assert np.all(dmin[sat] >= 16.7 * float(SIGMA)), float(dmin[sat].min() / SIGMA) if sat.any() else None
        for i in range(3):
            rej = ok & ~inside & ~(t_fast[i] > 0) & (u[i] >= far)
            assert np.all(d2_exact[i][rej] >= float(BLUR))      # dist to the triangle >= dist to this edge line >= blur
    assert total > 1_000_000 and amb_total > 100_000           # both sides of the ambiguous band were exercised


def test_ambiguous_pairs_are_rare_on_a_pixel_grid():
    """With pixel centres on the 512^2 grid (not adversarially close to edges) the exact re-classification is needed for a
    small fraction of the in-bbox pairs only."""
    rng = np.random.default_rng(1)
    n, S = 400_000, 512
    c = rng.uniform(-0.9, 0.9, size=(n, 1, 2))
    v = (c + (10.0 / S) * rng.normal(size=(n, 3, 2))).astype(f32)                   # ~5 px faces
    j = np.floor((1 - (c[:, 0] + (6.0 / S) * rng.normal(size=(n, 2)))) * S / 2 - 0.5)  # a pixel index near the face
    p = (1 - (2 * j + 1) / S).astype(f32)                                           # its centre in NDC (raster.cuh: pix_to_ndc)
    x0, y0, x1, y1, x2, y2 = v[:, 0, 0], v[:, 0, 1], v[:, 1, 0], v[:, 1, 1], v[:, 2, 0], v[:, 2, 1]
    px, py = p[:, 0], p[:, 1]
    inb = (px <= np.maximum(np.maximum(x0, x1), x2)) & (px >= np.minimum(np.minimum(x0, x1), x2)) & \
          (py <= np.maximum(np.maximum(y0, y1), y2)) & (py >= np.minimum(np.minimum(y0, y1), y2))
    tmin, eps = None, None
    for (ax, ay, bx, by) in ((x1, y1, x2, y2), (x2, y2, x0, y0), (x0, y0, x1, y1)):
        dx, dy = bx - ax, by - ay
        t = np.abs(_fma(dy, px, _fma(-dx, py, ((ay * dx).astype(f32) - (ax * dy).astype(f32)).astype(f32))))
        m = np.abs(dy) * (1 + np.abs(ax)) + np.abs(dx) * (1 + np.abs(ay))
        tmin = t if tmin is None else np.minimum(tmin, t)
        eps = m if eps is None else np.maximum(eps, m)
    frac = float(((tmin < f32(1e-6) * eps) & inb).sum()) / max(int(inb.sum()), 1)
    assert inb.sum() > 50_000 and frac < 0.02, frac


def _pix_to_ndc(idx, S):
    i = (S - 1 - idx).astype(f32)
    return f32(-1.0) + (f32(2.0) * i + f32(1.0)) / f32(S)


def _point_tri_dist2(p, v):
    """Exact-enough (float64) squared distance from points p (n,2) to triangles v (n,3,2) and strict inside flags."""
    p = p.astype(np.float64); v = v.astype(np.float64)
    d2 = np.full(len(p), np.inf)
    for i in range(3):
        a, b = v[:, i], v[:, (i + 1) % 3]
        ab = b - a
        t = np.clip(((p - a) * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-300), 0, 1)
        q = a + t[:, None] * ab
        d2 = np.minimum(d2, ((p - q) ** 2).sum(1))
    def e(a, b):
        return (p[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]) - (p[:, 1] - a[:, 1]) * (b[:, 0] - a[:, 0])
    e0, e1, e2 = e(v[:, 1], v[:, 2]), e(v[:, 2], v[:, 0]), e(v[:, 0], v[:, 1])
    inside = ((e0 > 0) & (e1 > 0) & (e2 > 0)) | ((e0 < 0) & (e1 < 0) & (e2 < 0))
    return d2, inside


def test_tile_edge_test_of_the_binning_is_conservative():
    """csrc/raster.cu: a face is left out of an 8x8 tile's list when all four corner pixel centres lie outside one edge
    line by more than sqrt(1.05 blur) (+1e-5 guard).  Emulated in float32; for every (face, tile) pair the rule drops, NO
    pixel centre of the tile may be inside the face or within sqrt(blur) of it (silhouette view), resp. inside it (light view)."""
    rng = np.random.default_rng(2)
    S, T = 512, 8
    n = 60_000
    c = rng.uniform(-0.95, 0.95, size=(n, 1, 2))
    size = 10 ** rng.uniform(-2.6, -1.2, size=(n, 1, 1))           # faces of ~1 to ~30 pixels
    v = (c + size * rng.normal(size=(n, 3, 2))).astype(f32)
    x = v[..., 0]; y = v[..., 1]
    for radius in (f32(np.sqrt(float(BLUR))), f32(0.0)):
        area = _edge_exact(x[:, 2], y[:, 2], x[:, 0], y[:, 0], x[:, 1], y[:, 1]) + f32(1e-8)
        sgn = np.where(area > 0, f32(1), f32(-1))
        m = f32(1.05) * radius * radius + f32(1e-10)
        # tiles of the bbox range (float32 arithmetic as face_tile_range)
        hs = f32(0.5 * S)
        xmin, xmax = x.min(1) - radius, x.max(1) + radius
        ymin, ymax = y.min(1) - radius, y.max(1) + radius
        jlo = np.clip(np.ceil((1 - xmax) * hs - f32(0.5) - f32(0.01)), 0, S - 1); jhi = np.clip(np.floor((1 - xmin) * hs - f32(0.5) + f32(0.01)), 0, S - 1)
        ilo = np.clip(np.ceil((1 - ymax) * hs - f32(0.5) - f32(0.01)), 0, S - 1); ihi = np.clip(np.floor((1 - ymin) * hs - f32(0.5) + f32(0.01)), 0, S - 1)
        valid = (jlo <= jhi) & (ilo <= ihi) & (np.abs(area) > 1e-8)
        tx = (jlo // T + rng.integers(0, 4, n)).astype(np.int64)    # one random tile of (a superset of) each face's range
        ty = (ilo // T + rng.integers(0, 4, n)).astype(np.int64)
        valid &= (tx <= jhi // T) & (ty <= ihi // T)
        xa, xb = _pix_to_ndc(tx * T, S), _pix_to_ndc(tx * T + T - 1, S)
        ya, yb = _pix_to_ndc(ty * T, S), _pix_to_ndc(ty * T + T - 1, S)
        reject = np.zeros(n, bool)
        for (i0, i1) in ((1, 2), (2, 0), (0, 1)):
            ax, ay, bx, by = x[:, i0], y[:, i0], x[:, i1], y[:, i1]
            dx, dy = bx - ax, by - ay
            ea, eb, fa, fb = (xa - ax) * dy, (xb - ax) * dy, (ya - ay) * dx, (yb - ay) * dx
            mx = np.maximum(np.maximum(sgn * (ea - fa), sgn * (ea - fb)), np.maximum(sgn * (eb - fa), sgn * (eb - fb)))
            thr = m * (dx * dx + dy * dy)
            reject |= (mx < 0) & (mx * mx >= thr)
        rej = np.nonzero(valid & reject)[0]
        assert len(rej) > 3000
        # every pixel centre of every rejected (face, tile) pair
        jj, ii = np.meshgrid(np.arange(T), np.arange(T))
        for dj, di in zip(jj.ravel(), ii.ravel()):
            pxy = np.stack([_pix_to_ndc(tx[rej] * T + dj, S), _pix_to_ndc(ty[rej] * T + di, S)], 1)
            d2, inside = _point_tri_dist2(pxy, v[rej])
            assert not inside.any()
            if radius > 0:
                assert np.all(d2 >= float(BLUR)), float(d2.min() / float(BLUR))


def test_fast_inverse_depth_orders_faces_like_the_exact_depth():
    """Visibility candidates are ordered by the fast inverse depth q = sum(t_i / z_i) / sum(t_i) (raster.cuh: fast_inv_depth)
    and only pairs within 1e-5 relative of each other are re-decided with PyTorch3D's exact perspective-correct depth
    (vis_update).  Emulated in float32: q must equal 1 / pz_exact to well within that tolerance, so the nearest face chosen
    outside the tie band is the one PyTorch3D's z-ordering picks."""
    rng = np.random.default_rng(3)
    n = 500_000
    c = rng.uniform(-0.9, 0.9, size=(n, 1, 2))
    v = (c + 0.02 * rng.normal(size=(n, 3, 2))).astype(f32)
    z = rng.uniform(0.3, 1.5, size=(n, 1)).astype(f32) + (0.05 * rng.normal(size=(n, 3))).astype(f32)   # metres, hand-sized relief
    w = rng.dirichlet((1.0, 1.0, 1.0), size=n)                                                           # a point strictly inside
    p = (w[:, :, None] * v.astype(np.float64)).sum(1).astype(f32)
    x0, y0, x1, y1, x2, y2 = v[:, 0, 0], v[:, 0, 1], v[:, 1, 0], v[:, 1, 1], v[:, 2, 0], v[:, 2, 1]
    px, py = p[:, 0], p[:, 1]
    area = _edge_exact(x2, y2, x0, y0, x1, y1) + f32(1e-8)
    e0, e1, e2 = _edge_exact(px, py, x1, y1, x2, y2), _edge_exact(px, py, x2, y2, x0, y0), _edge_exact(px, py, x0, y0, x1, y1)
    w0, w1, w2 = e0 / area, e1 / area, e2 / area
    inside = (w0 > 0) & (w1 > 0) & (w2 > 0) & (np.abs(area) > 1e-6)
    z0, z1, z2 = z[:, 0], z[:, 1], z[:, 2]
    # BarycentricPerspectiveCorrectionForward + the depth interpolation, every operation rounded to float32 (eval_pair_full)
    t0, t1, t2 = (w0 * z1) * z2, (z0 * w1) * z2, (z0 * z1) * w2
    den = np.maximum((t0 + t1) + t2, f32(1e-8))
    b0, b1, b2 = t0 / den, t1 / den, t2 / den
    pz = (b0 * z0 + b1 * z1) + b2 * z2
    q = ((e0 * (f32(1) / z0) + e1 * (f32(1) / z1)) + e2 * (f32(1) / z2)) / ((e0 + e1) + e2)
    rel = np.abs(q.astype(np.float64) * pz.astype(np.float64) - 1.0)[inside]
    assert inside.sum() > 300_000
    assert float(rel.max()) < 2.5e-6, float(rel.max())   # 4x inside the 1e-5 tie band
