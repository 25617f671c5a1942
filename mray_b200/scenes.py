"""Deterministic procedural scenes for the BASELINE.json configs (numpy only, no GPU).

The reference's Sponza-class scenes are an external download that is not available offline
(SURVEY.md §8c), so config 2/3/5 use fixed generators:

* ``arcade_mesh``   — "Sponza-scale" closed hall with two rows of columns, arches, displaced
                      floor/walls and blob clutter; ~260 K triangles, coordinates in [-20,20]^3.
* ``random_soup``   — uniformly random small triangles in the unit cube (BVH build sweep, config 5).
* ``pinhole_rays``  — deterministic pixel-centre primary rays (config 2 batch A).
* ``cornell_box``   — triangles-only Cornell box (config 1).

All functions return C-contiguous float32 / uint32 arrays.
"""
from __future__ import annotations

import numpy as np


def _grid(nu: int, nv: int):
    """Index list of an (nu+1) x (nv+1) vertex grid -> 2*nu*nv triangles (CCW in u,v)."""
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).ravel()
    b = a + (nv + 1)
    c = b + 1
    d = a + 1
    return np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3)


def _patch(origin, eu, ev, nu, nv, disp=None, normal=None):
    u, v = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="ij")
    p = (np.asarray(origin, np.float64)[None, None, :]
         + u[..., None] * np.asarray(eu, np.float64) + v[..., None] * np.asarray(ev, np.float64))
    if disp is not None:
        n = np.asarray(normal, np.float64)
        p = p + disp(u, v)[..., None] * n
    return p.reshape(-1, 3), _grid(nu, nv)


def _cylinder(center, radius, y0, y1, nseg, nh, flute=0.0):
    th = np.linspace(0, 2 * np.pi, nseg + 1)
    h = np.linspace(y0, y1, nh + 1)
    T, H = np.meshgrid(th, h, indexing="ij")
    r = radius * (1.0 + flute * np.cos(12 * T)) * (1.0 + 0.04 * np.cos(2 * np.pi * (H - y0) / (y1 - y0)))
    p = np.stack([center[0] + r * np.cos(T), H, center[1] + r * np.sin(T)], -1)
    return p.reshape(-1, 3), _grid(nseg, nh)


def _sphere(center, radius, nu, nv, bump, rng):
    th = np.linspace(0, 2 * np.pi, nu + 1)
    ph = np.linspace(1e-3, np.pi - 1e-3, nv + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    k = rng.integers(2, 6, size=3)
    r = radius * (1.0 + bump * np.sin(k[0] * T) * np.sin(k[1] * P) + 0.5 * bump * np.cos(k[2] * P))
    p = np.stack([center[0] + r * np.sin(P) * np.cos(T),
                  center[1] + r * np.cos(P),
                  center[2] + r * np.sin(P) * np.sin(T)], -1)
    return p.reshape(-1, 3), _grid(nu, nv)


def arcade_mesh(target_tris: int = 262144, seed: int = 1234):
    """Closed hall (40 x 24 x 40, inside [-20,20]^3) with colonnades. Returns (positions[V,3] f32,
    indices[T,3] u32). ``target_tris`` scales the tessellation; the default gives ~260 K."""
    rng = np.random.default_rng(seed)
    s = max(0.05, np.sqrt(target_tris / 288000.0))  # 288000: measured count at s = 1
    parts = []

    def q(n):
        return max(2, int(round(n * s)))

    ph = rng.uniform(0, 2 * np.pi, size=8)

    def bumps(amp, f):
        return lambda u, v: amp * (np.sin(f * u * 2 * np.pi + ph[0]) * np.cos(f * v * 2 * np.pi + ph[1])
                                   + 0.35 * np.sin(3.1 * f * u * 2 * np.pi + ph[2]) * np.sin(2.7 * f * v * 2 * np.pi + ph[3]))

    X, Y0, Y1, Z = 19.5, -12.0, 12.0, 19.5
    # floor / ceiling / 4 walls (normals pointing inside)
    parts.append(_patch([-X, Y0, -Z], [0, 0, 2 * Z], [2 * X, 0, 0], q(128), q(128), bumps(0.18, 9), [0, 1, 0]))
    parts.append(_patch([-X, Y1, -Z], [2 * X, 0, 0], [0, 0, 2 * Z], q(64), q(64), bumps(0.25, 5), [0, -1, 0]))
    parts.append(_patch([-X, Y0, -Z], [0, Y1 - Y0, 0], [0, 0, 2 * Z], q(72), q(96), bumps(0.22, 7), [1, 0, 0]))
    parts.append(_patch([X, Y0, -Z], [0, 0, 2 * Z], [0, Y1 - Y0, 0], q(96), q(72), bumps(0.22, 7), [-1, 0, 0]))
    parts.append(_patch([-X, Y0, -Z], [2 * X, 0, 0], [0, Y1 - Y0, 0], q(96), q(72), bumps(0.22, 6), [0, 0, 1]))
    parts.append(_patch([-X, Y0, Z], [0, Y1 - Y0, 0], [2 * X, 0, 0], q(72), q(96), bumps(0.22, 6), [0, 0, -1]))
    # two storeys of colonnades along z, plus arches (half tori) between neighbouring columns
    col_x = [-9.0, 9.0]
    col_z = np.linspace(-16.0, 16.0, 9)
    for cx in col_x:
        for cz in col_z:
            parts.append(_cylinder((cx, cz), 0.85, Y0, 0.0, q(48), q(40), flute=0.05))
            parts.append(_cylinder((cx, cz), 0.6, 0.8, 8.0, q(40), q(32), flute=0.04))
        for z0, z1 in zip(col_z[:-1], col_z[1:]):
            # arch: swept circle along a half circle in the (z,y) plane
            nu, nv = q(28), q(12)
            a = np.linspace(0, np.pi, nu + 1)
            b = np.linspace(0, 2 * np.pi, nv + 1)
            A, B = np.meshgrid(a, b, indexing="ij")
            R, r = 0.5 * (z1 - z0), 0.35
            zc = 0.5 * (z0 + z1)
            p = np.stack([cx + r * np.cos(B),
                          0.0 + (R + r * np.sin(B)) * np.sin(A) * 0.45,
                          zc - (R + r * np.sin(B)) * np.cos(A)], -1).reshape(-1, 3)
            parts.append((p, _grid(nu, nv)))
        # gallery floor slab above the arches
        parts.append(_patch([cx - 1.4, 0.75, -17.0], [0, 0, 34.0], [2.8, 0, 0], q(96), q(8), bumps(0.03, 11), [0, 1, 0]))
    # clutter: bumpy blobs on the floor and hanging from the ceiling
    for _ in range(14):
        c = np.array([rng.uniform(-7, 7), rng.uniform(Y0 + 1.0, Y0 + 2.5), rng.uniform(-17, 17)])
        parts.append(_sphere(c, rng.uniform(0.7, 1.6), q(40), q(28), 0.08, rng))
    for _ in range(6):
        c = np.array([rng.uniform(-6, 6), rng.uniform(5.0, 9.0), rng.uniform(-15, 15)])
        parts.append(_sphere(c, rng.uniform(0.5, 1.2), q(32), q(24), 0.12, rng))
    # hanging curtains (thin displaced sheets) across the nave
    for zc in np.linspace(-12, 12, 4):
        parts.append(_patch([-7.5, 1.0, zc], [15.0, 0, 0], [0, 9.5, 0], q(72), q(40),
                            lambda u, v: 0.35 * np.sin(14 * np.pi * u + ph[4]) * (1.0 - 0.6 * v), [0, 0, 1]))

    pos, idx, off = [], [], 0
    for p, i in parts:
        pos.append(p)
        idx.append(i + off)
        off += p.shape[0]
    positions = np.ascontiguousarray(np.concatenate(pos).astype(np.float32))
    indices = np.ascontiguousarray(np.concatenate(idx).astype(np.uint32))
    assert np.abs(positions).max() <= 20.0
    return positions, indices


def random_soup(n_tris: int, seed: int | None = None, size: float = 0.01):
    """Uniformly random small triangles in the unit cube (config 5 build sweep; seed defaults to N)."""
    rng = np.random.default_rng(n_tris if seed is None else seed)
    c = rng.random((n_tris, 1, 3), dtype=np.float32)
    d = (rng.random((n_tris, 3, 3), dtype=np.float32) - 0.5) * np.float32(size)
    positions = np.ascontiguousarray(np.clip(c + d, 0.0, 1.0).reshape(-1, 3).astype(np.float32))
    indices = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    return positions, indices


def pinhole_rays(width: int, height: int, eye, gaze, up, fov_y_deg: float,
                 t_min: float = 0.0, t_max: float = np.float32(3.4028235e38)):
    """Deterministic pixel-centre primary rays, RayGMem layout [pos(3), tMin, dir(3), tMax] float32
    (Tracer/TracerTypes.h:L276-283). Row-major, pixel (0,0) first."""
    eye = np.asarray(eye, np.float64)
    w = np.asarray(gaze, np.float64) - eye
    w /= np.linalg.norm(w)
    u = np.cross(w, np.asarray(up, np.float64))
    u /= np.linalg.norm(u)
    v = np.cross(u, w)
    th = np.tan(np.deg2rad(fov_y_deg) * 0.5)
    aspect = width / height
    px = ((np.arange(width) + 0.5) / width * 2.0 - 1.0) * th * aspect
    py = (1.0 - (np.arange(height) + 0.5) / height * 2.0) * th
    PX, PY = np.meshgrid(px, py, indexing="xy")
    d = w[None, None, :] + PX[..., None] * u + PY[..., None] * v
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.empty((height * width, 8), np.float32)
    rays[:, 0:3] = eye.astype(np.float32)
    rays[:, 3] = t_min
    rays[:, 4:7] = d.reshape(-1, 3).astype(np.float32)
    rays[:, 7] = t_max
    return rays


ARCADE_CAMERA = dict(eye=(0.0, -4.0, 18.5), gaze=(1.5, -3.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=60.0)


def pcg32_floats(seeds: np.ndarray, n: int):
    """n uniform floats in [0,1) per seed from PCG32 (XSH-RR 64/32, the generator family of the
    reference's PermutedCG32, Tracer/Random.h:L763-771). Deterministic ray-batch helper only."""
    mult = np.uint64(6364136223846793005)
    inc = np.uint64(1442695040888963407)
    state = (seeds.astype(np.uint64) + inc) * mult + inc
    out = np.empty((seeds.shape[0], n), np.float32)
    with np.errstate(over="ignore"):
        for k in range(n):
            old = state
            state = old * mult + inc
            xs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
            rot = (old >> np.uint64(59)).astype(np.uint32)
            r = (xs >> rot) | (xs << ((np.uint32(32) - rot) & np.uint32(31)))
            out[:, k] = np.minimum(r.astype(np.float64) * 2.0 ** -32, np.float64(np.nextafter(np.float32(1), np.float32(0))))
    return out


def ao_rays(rays: np.ndarray, prim: np.ndarray, t: np.ndarray, positions: np.ndarray,
            indices: np.ndarray, t_max: float, seed_offset: int = 0):
    """Config 2 batch B: from each primary hit one cosine-hemisphere direction about the geometric
    normal (flipped towards the viewer), PCG32 seeded by pixel index. Misses produce tMax = 0 rays
    (never hit anything). Origin is offset by 1e-3 along the normal."""
    n = rays.shape[0]
    hit = prim != 0xFFFFFFFF
    pr = np.where(hit, prim, 0).astype(np.int64)
    p0 = positions[indices[pr, 0]].astype(np.float64)
    p1 = positions[indices[pr, 1]].astype(np.float64)
    p2 = positions[indices[pr, 2]].astype(np.float64)
    nrm = np.cross(p1 - p0, p2 - p0)
    ln = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = nrm / np.where(ln > 0, ln, 1.0)
    d = rays[:, 4:7].astype(np.float64)
    nrm = np.where((np.sum(nrm * d, 1, keepdims=True) > 0), -nrm, nrm)
    xi = pcg32_floats(np.arange(n, dtype=np.uint64) + np.uint64(seed_offset), 2).astype(np.float64)
    r = np.sqrt(xi[:, 0])
    phi = 2 * np.pi * xi[:, 1]
    lx, ly, lz = r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0.0, 1.0 - xi[:, 0]))
    a = np.where(np.abs(nrm[:, 0:1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    tx = np.cross(a, nrm)
    tx /= np.linalg.norm(tx, axis=1, keepdims=True)
    ty = np.cross(nrm, tx)
    wdir = lx[:, None] * tx + ly[:, None] * ty + lz[:, None] * nrm
    wdir /= np.linalg.norm(wdir, axis=1, keepdims=True)
    tt = np.where(hit, t, 0.0).astype(np.float64)
    org = rays[:, 0:3].astype(np.float64) + tt[:, None] * d + 1e-3 * nrm
    out = np.empty((n, 8), np.float32)
    out[:, 0:3] = org.astype(np.float32)
    out[:, 3] = 0.0
    out[:, 4:7] = wdir.astype(np.float32)
    out[:, 7] = np.where(hit, np.float32(t_max), np.float32(0.0))
    return out


def cornell_box():
    """Triangles-only Cornell box after the reference's documented scene
    (Docs/markdown/scene/mrayScene.md:L310-553): walls from a unit plane, two boxes instead of the
    spheres. Returns dict with positions, indices, per-triangle material id
    (0 white, 1 red, 2 green, 3 light) and the camera."""
    quads = []

    def quad(a, b, c, d, m):
        quads.append((np.array([a, b, c, d], np.float32), m))

    # room spans x,z in [-1,1], y in [0,2]
    quad([-1, 0, 1], [1, 0, 1], [1, 0, -1], [-1, 0, -1], 0)     # floor
    quad([-1, 2, -1], [1, 2, -1], [1, 2, 1], [-1, 2, 1], 0)     # ceiling
    quad([-1, 0, -1], [1, 0, -1], [1, 2, -1], [-1, 2, -1], 0)   # back
    quad([-1, 0, 1], [-1, 0, -1], [-1, 2, -1], [-1, 2, 1], 1)   # left (red)
    quad([1, 0, -1], [1, 0, 1], [1, 2, 1], [1, 2, -1], 2)       # right (green)
    quad([-0.25, 1.98, -0.25], [0.25, 1.98, -0.25], [0.25, 1.98, 0.25], [-0.25, 1.98, 0.25], 3)  # light

    def box(cx, cz, sx, sy, sz, ang):
        ca, sa = np.cos(ang), np.sin(ang)
        def P(x, y, z):
            return [cx + ca * x * sx - sa * z * sz, y * sy, cz + sa * x * sx + ca * z * sz]
        v = [P(-1, 0, -1), P(1, 0, -1), P(1, 0, 1), P(-1, 0, 1), P(-1, 1, -1), P(1, 1, -1), P(1, 1, 1), P(-1, 1, 1)]
        for f in ([4, 7, 6, 5], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [0, 3, 2, 1]):
            quad(v[f[0]], v[f[1]], v[f[2]], v[f[3]], 0)

    box(-0.35, -0.3, 0.3, 1.2, 0.3, 0.3)
    box(0.4, 0.35, 0.3, 0.6, 0.3, -0.3)
    pos, idx, mat = [], [], []
    for k, (qv, m) in enumerate(quads):
        pos.append(qv)
        idx += [[4 * k, 4 * k + 1, 4 * k + 2], [4 * k, 4 * k + 2, 4 * k + 3]]
        mat += [m, m]
    return dict(positions=np.ascontiguousarray(np.concatenate(pos)),
                indices=np.array(idx, np.uint32), material=np.array(mat, np.uint32),
                albedo=np.array([[0.725, 0.71, 0.68], [0.63, 0.065, 0.05], [0.14, 0.45, 0.091], [0, 0, 0]], np.float32),
                radiance=np.array([68.0, 48.0, 16.0], np.float32),
                camera=dict(eye=(0.0, 1.0, 6.8), gaze=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=19.5))


def arcade_materials(positions, indices, material_count: int = 64, seed: int = 1234):
    """Config 3 material assignment for ``arcade_mesh``: ``material_count`` Lambert albedos assigned to
    contiguous triangle blocks, plus one emissive patch (~200 ceiling triangles around the nave centre).
    Returns (indices sorted by material, prim_ranges, light_or_mat_keys, albedo[material_count,3], radiance[3],
    tri_material[int32, -1 = light]) with the triangles reordered so that each material is one prim range."""
    rng = np.random.default_rng(seed)
    n = indices.shape[0]
    block = max(1, n // (material_count * 8))
    mat = ((np.arange(n) // block) % material_count).astype(np.int32)
    cen = positions[indices].mean(axis=1)
    nrm = np.cross(positions[indices[:, 1]] - positions[indices[:, 0]], positions[indices[:, 2]] - positions[indices[:, 0]])
    ceiling = (cen[:, 1] > 11.0) & (nrm[:, 1] < 0)
    d2 = cen[:, 0] ** 2 + cen[:, 2] ** 2
    cand = np.nonzero(ceiling)[0]
    light = cand[np.argsort(d2[cand])[:200]]
    mat[light] = -1
    order = np.argsort(np.where(mat < 0, material_count, mat), kind="stable")
    idx = np.ascontiguousarray(indices[order])
    m = mat[order]
    ranges, keys = [], []
    for k in list(range(material_count)) + [-1]:
        w = np.nonzero(m == k)[0]
        if w.size == 0:
            continue
        ranges.append([int(w[0]), int(w[-1]) + 1])
        keys.append((0x80000000 | 0) if k < 0 else k)
    albedo = (0.25 + 0.6 * rng.random((material_count, 3))).astype(np.float32)
    radiance = np.array([40.0, 36.0, 30.0], np.float32)
    return idx, np.array(ranges, np.uint32), np.array(keys, np.uint32), albedo, radiance, m


def instanced_field(seed: int = 99, instance_count: int = 1000, target_tris: int = 10_000_000, material_count: int = 64,
                    light_panels: int = 16):
    """BASELINE config 4: ``instance_count`` instances ((T)Single, random rotation * scale * translation) of 10
    distinct meshes of 1 K .. 100 K triangles (bumpy spheres / fluted columns, log-spaced sizes; the smaller meshes
    get more instances so that the instanced total is ~``target_tris``), one of ``material_count`` Lambert
    materials per INSTANCE (round-robin), a ground plane and ``light_panels`` emissive quads above the field.

    Returns dict(meshes=[(positions, indices)], instances=[(mesh, 3x4 transform or None, material or -1 for the
    light)], albedo[material_count,3], radiance[3], camera, triangles_instanced)."""
    rng = np.random.default_rng(seed)
    sizes = np.round(1000.0 * 10.0 ** (2.0 * np.arange(10) / 9.0)).astype(int)
    meshes = []
    for k, s in enumerate(sizes):
        if k % 2 == 0:
            nv = max(4, int(np.sqrt(s / 4.0))); nu = max(8, int(round(s / (2.0 * nv))))
            p, i = _sphere((0.0, 0.0, 0.0), 1.0, nu, nv, 0.08 + 0.02 * k, rng)
        else:
            nh = max(4, int(np.sqrt(s / 8.0))); nseg = max(8, int(round(s / (2.0 * nh))))
            p, i = _cylinder((0.0, 0.0), 0.5, -1.5, 1.5, nseg, nh, flute=0.05)
        meshes.append((np.ascontiguousarray(p, np.float32), np.ascontiguousarray(i, np.uint32)))
    tris = np.array([m[1].shape[0] for m in meshes], np.float64)
    # instance counts n_k ~ tris_k^-a with sum n_k = instance_count and sum n_k tris_k ~ target_tris (bisection on a)
    lo, hi = 0.0, 3.0
    for _ in range(60):
        a = 0.5 * (lo + hi)
        w = tris ** -a; n = instance_count * w / w.sum()
        if (n * tris).sum() > target_tris: lo = a
        else: hi = a
    n = np.maximum(1, np.floor(instance_count * w / w.sum())).astype(int)
    n[0] += instance_count - n.sum()
    which = np.repeat(np.arange(10), n); rng.shuffle(which)
    extent = 60.0
    instances = []
    for k, m in enumerate(which):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w_, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                      [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                      [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
        s = rng.uniform(0.8, 2.2)
        t = np.array([rng.uniform(-extent, extent), rng.uniform(1.0, 14.0), rng.uniform(-extent, extent)])
        instances.append((int(m), np.hstack([R * s, t[:, None]]), k % material_count))
    # ground (two triangles, identity) and light panels (one quad mesh, instanced)
    g = 1.5 * extent
    meshes.append((np.array([[-g, 0, g], [g, 0, g], [g, 0, -g], [-g, 0, -g]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)))
    instances.append((10, None, 0))
    meshes.append((np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)))
    for k in range(light_panels):
        t = np.array([rng.uniform(-extent, extent), 24.0, rng.uniform(-extent, extent)])
        instances.append((11, np.hstack([np.eye(3) * 5.0, t[:, None]]), -1))
    albedo = (0.25 + 0.6 * rng.random((material_count, 3))).astype(np.float32)
    camera = dict(eye=(0.0, 22.0, 1.45 * extent), gaze=(0.0, 4.0, 0.0), up=(0.0, 1.0, 0.0), fov_y_deg=50.0)
    total = int(sum(meshes[m][1].shape[0] for m, _, _ in instances))
    return dict(meshes=meshes, instances=instances, albedo=albedo, radiance=np.array([60.0, 54.0, 45.0], np.float32),
                camera=camera, triangles_instanced=total)


def cornell_textures(seed: int = 5):
    """UV0 + albedo textures for ``cornell_box`` (every quad owns its 4 vertices): the white material reads an 8x8
    fp32 RGBA texture (bilinear, wrap) over uv in [-0.75, 1.75]^2 — wrap-around and the negative-texel path of the
    reference's filter both occur —, the red one a 4x4 unorm8 RGBA texture (nearest, clamp), green stays constant.
    Returns (vertex_uvs[V, 2], textures, albedo_texture[4])."""
    rng = np.random.default_rng(seed)
    c = cornell_box()
    nq = c["positions"].shape[0] // 4
    corner = np.array([[-0.75, -0.75], [1.75, -0.75], [1.75, 1.75], [-0.75, 1.75]], np.float32)
    uvs = np.tile(corner, (nq, 1)).astype(np.float32)
    t0 = (0.1 + 0.8 * rng.random((8, 8, 4))).astype(np.float32)   # RGBA: TracerI textures are pushed as 4-channel pixels
    t0[..., 3] = 1.0
    t1 = rng.integers(40, 230, size=(4, 4, 4), dtype=np.uint8)
    t1[..., 1:3] //= 4          # keep it reddish
    textures = [dict(data=t0, interp="Linear", edge="Wrap"), dict(data=t1, interp="Nearest", edge="Clamp")]
    return uvs, textures, np.array([0, 1, -1, -1], np.int32)


def cornell_mirror():
    """``cornell_box`` with the tall box turned into a perfect mirror: material 4 = (Mt)Reflect (triangles 12..23).
    Returns the cornell dict with `material` updated, a 5-row albedo table and `material_type` (0 Lambert, 1 Reflect)
    per material id (id 3 is the light)."""
    c = cornell_box()
    m = c["material"].copy()
    m[12:24] = 4
    c["material"] = m
    c["albedo"] = np.concatenate([c["albedo"], np.zeros((1, 3), np.float32)])
    c["material_type"] = np.array([0, 0, 0, 0, 1], np.uint8)
    return c


def cornell_glossy():
    """``cornell_box`` with the tall box made of (Mt)Unreal (a rough gold-like metal) and the short box of (Mt)Refract
    (glass, air outside): material 4 = Unreal (triangles 12..23), material 5 = Refract (triangles 24..35). Returns the
    cornell dict with `material` updated, a 6-row albedo table, `material_type` (mrb_material_type per material id; id 3 is
    the light) and `material_params` [6, 8] (Unreal: roughness, specular, metallic; Refract: cauchyFront xyz, -, cauchyBack xyz)."""
    c = cornell_box()
    m = c["material"].copy()
    m[12:24] = 4
    m[24:36] = 5
    c["material"] = m
    c["albedo"] = np.concatenate([c["albedo"], np.array([[0.95, 0.64, 0.30], [0, 0, 0]], np.float32)])
    c["material_type"] = np.array([0, 0, 0, 0, 3, 2], np.uint8)
    mp = np.zeros((6, 8), np.float32)
    mp[4, :3] = [0.35, 0.5, 0.8]                 # roughness, specular, metallic
    mp[5, :3] = [1.0, 0.0, 0.0]                  # cauchyFront: air
    mp[5, 4:7] = [1.5046, 0.0042, 0.0]           # cauchyBack: BK7-like glass
    c["material_params"] = mp
    return c


def icosphere(subdivisions: int = 1):
    """Unit icosphere: (positions [V, 3] float64, indices [T, 3]); 20 * 4^subdivisions triangles, shared vertices."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []
        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    return np.array(v), np.array(f, np.uint32)


def cornell_sphere():
    """``cornell_box`` with the short box replaced by an 80-triangle sphere with SMOOTH vertex normals (the only fixture
    whose shading normals differ from the geometric ones: Triangle::GenerateSurface's interpolated tangent frames).
    Returns the cornell dict plus `normals` [V, 3] (face normals on the quads, radial on the sphere)."""
    c = cornell_box()
    keep = np.ones(c["indices"].shape[0], bool); keep[24:36] = False
    idx = c["indices"][keep]; mat = c["material"][keep]
    pos = c["positions"].astype(np.float64)
    fn = np.cross(pos[c["indices"][:, 1]] - pos[c["indices"][:, 0]], pos[c["indices"][:, 2]] - pos[c["indices"][:, 0]])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    nrm = np.zeros_like(pos)
    for k in range(3):
        nrm[c["indices"][:, k]] = fn                      # quads: unshared vertices, every corner takes its face's normal
    sv, sf = icosphere(1)
    centre, radius = np.array([0.4, 0.35, 0.35]), 0.35
    base = pos.shape[0]
    pos = np.concatenate([pos, centre + radius * sv]); nrm = np.concatenate([nrm, sv])
    idx = np.concatenate([idx, sf + base]); mat = np.concatenate([mat, np.zeros(sf.shape[0], np.uint32)])
    c["positions"] = np.ascontiguousarray(pos, np.float32); c["indices"] = np.ascontiguousarray(idx, np.uint32)
    c["material"] = mat.astype(np.uint32); c["normals"] = np.ascontiguousarray(nrm, np.float32)
    return c


def cornell_open(keep_light: bool = False):
    """``cornell_box`` with the ceiling removed (and, unless keep_light, the area light too): the room is lit through the
    opening by the BOUNDARY light — the skysphere test scene. Same dict layout as cornell_box."""
    c = cornell_box()
    keep = np.ones(c["indices"].shape[0], bool)
    keep[2:4] = False                  # ceiling quad = triangles 2, 3
    if not keep_light:
        keep[10:12] = False            # light quad = triangles 10, 11
    c["indices"] = np.ascontiguousarray(c["indices"][keep])
    c["material"] = np.ascontiguousarray(c["material"][keep])
    return c


def sky_texture(width: int = 32, height: int = 16, seed: int = 9):
    """A small HDR environment map (RGBA fp32, bilinear, wrap): a blue-ish gradient brighter towards the top rows, low-
    amplitude noise, and a 2x2-texel "sun" three orders of magnitude brighter — so that importance sampling of the
    luminance distribution matters. Row 0 is v = 0 (towards -Y on the latitude-longitude map)."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float32) + 0.5) / height
    t = np.zeros((height, width, 4), np.float32)
    t[..., 0] = (0.15 + 0.6 * v)[:, None]
    t[..., 1] = (0.2 + 0.9 * v)[:, None]
    t[..., 2] = (0.3 + 1.6 * v)[:, None]
    t[..., :3] *= (0.85 + 0.3 * rng.random((height, width, 1))).astype(np.float32)
    sy, sx = int(height * 0.8), int(width * 0.3)
    t[sy:sy + 2, sx:sx + 2, :3] = np.array([900.0, 800.0, 600.0], np.float32)
    t[..., 3] = 1.0
    return dict(data=np.ascontiguousarray(t), interp="Linear", edge="Wrap")


def cornell_alpha(seed: int = 21):
    """``cornell_box`` plus a pane (material 4, bluish Lambert) standing in front of the boxes whose surface carries an ALPHA
    MAP: an 8x8 single-channel fp32 texture (nearest, clamp) with fully transparent, fully opaque and fractional texels, so
    camera rays, bounce rays and NEE shadow rays all meet the stochastic alpha test. Returns the cornell dict with
    `material` / `albedo` extended, `uvs` [V, 2] (the pane spans uv [0, 1]^2, everything else (0, 0)), `alpha_texture`
    and `alpha_map` (per material id: -1 or 0)."""
    rng = np.random.default_rng(seed)
    c = cornell_box()
    pane = np.array([[-0.8, 0.2, 0.6], [0.8, 0.2, 0.6], [0.8, 1.6, 0.6], [-0.8, 1.6, 0.6]], np.float32)
    nv = c["positions"].shape[0]
    c["positions"] = np.ascontiguousarray(np.concatenate([c["positions"], pane]))
    c["indices"] = np.ascontiguousarray(np.concatenate([c["indices"], np.array([[nv, nv + 1, nv + 2], [nv, nv + 2, nv + 3]], np.uint32)]))
    c["material"] = np.concatenate([c["material"], np.array([4, 4], np.uint32)])
    c["albedo"] = np.concatenate([c["albedo"], np.array([[0.15, 0.25, 0.7]], np.float32)])
    uvs = np.zeros((nv + 4, 2), np.float32)
    uvs[nv:] = [[0, 0], [1, 0], [1, 1], [0, 1]]
    a = rng.choice(np.array([0.0, 0.0, 1.0, 1.0, 0.25, 0.5, 0.75], np.float32), size=(8, 8)).astype(np.float32)
    c["uvs"] = uvs
    c["alpha_texture"] = dict(data=np.ascontiguousarray(a), interp="Nearest", edge="Clamp")
    c["alpha_map"] = np.array([-1, -1, -1, -1, 0], np.int32)
    return c


def mray_scene_json(c, width: int, height: int, near_far=(0.01, 1000.0), light_material: int = 3, material_types=None,
                    boundary=None, textures=None, alpha_map=None, uvs=None, cull_back_face: bool = False, albedo_texture=None) -> str:
    """A scene dict (cornell_box layout: positions, indices, per-triangle material id, albedo, radiance, camera) as a scene file
    of the reference's JSON format (Docs/markdown/scene/mrayScene.md): one in-node indexed triangle primitive per material id,
    one surface each, the `light_material` batch as a Primitive light. boundary: None (Null light) or dict(type=
    "Skysphere_Spherical"|"Skysphere_CoOcta", radiance=[r, g, b] | texture=scene texture id); textures: list of
    dict(id=, file=, ...) texture nodes; alpha_map / albedo_texture: per material id a scene texture id or None."""
    import json
    pos, idx, mat = c["positions"], c["indices"], c["material"]
    prims, mats, surfaces = [], [], []
    for m in np.unique(mat):
        tris = idx[mat == m]
        used, inv = np.unique(tris.ravel(), return_inverse=True)
        node = {"id": int(m), "type": "Triangle", "tag": "nodeTriangleIndexed",
                "position": pos[used].astype(float).round(9).tolist(), "index": inv.reshape(-1, 3).astype(int).tolist()}
        if c.get("normals") is not None:
            node["normal"] = np.asarray(c["normals"])[used].astype(float).tolist()
        if uvs is not None:
            node["uv"] = np.asarray(uvs)[used].astype(float).tolist()
        prims.append(node)
        if int(m) == light_material:
            continue
        kind = "Lambert" if material_types is None else material_types[int(m)]
        mnode = {"id": int(m), "type": kind}
        if kind in ("Lambert", "Unreal"):
            mnode["albedo"] = [float(x) for x in c["albedo"][int(m)]]
            if albedo_texture is not None and albedo_texture[int(m)] is not None:
                mnode["albedo"] = {"texture": int(albedo_texture[int(m)])}
        mats.append(mnode)
        s = {"transform": 0, "material": int(m), "primitive": int(m), "cullBackFace": bool(cull_back_face)}
        if alpha_map is not None and alpha_map[int(m)] is not None:
            s["alphaMap"] = {"texture": int(alpha_map[int(m)])}
        surfaces.append(s)
    cam = c["camera"]
    lights = [{"id": 0, "type": "Null"}]
    light_surfaces = []
    if light_material in np.unique(mat):
        lights.append({"id": 1, "type": "Primitive", "primitive": int(light_material), "radiance": [float(x) for x in c["radiance"]]})
        light_surfaces.append({"light": 1, "transform": 0})
    b_light = 0
    if boundary is not None:
        node = {"id": 2, "type": boundary["type"]}
        node["radiance"] = {"texture": int(boundary["texture"])} if "texture" in boundary else [float(x) for x in boundary["radiance"]]
        lights.append(node); b_light = 2
    scene = {
        "Cameras": [{"id": 0, "type": "Pinhole", "isFovX": False, "fov": float(cam["fov_y_deg"]), "aspect": width / height,
                     "planes": [float(near_far[0]), float(near_far[1])], "gaze": list(map(float, cam["gaze"])),
                     "position": list(map(float, cam["eye"])), "up": list(map(float, cam["up"]))}],
        "Lights": lights, "Mediums": [{"id": 0, "type": "Vacuum"}], "Transforms": [{"id": 0, "type": "Identity"}],
        "Materials": mats, "Primitives": prims, "Textures": textures or [],
        "Boundary": {"medium": 0, "light": b_light, "transform": 0},
        "Surfaces": surfaces, "LightSurfaces": light_surfaces, "CameraSurfaces": [{"camera": 0}],
    }
    return json.dumps(scene, indent=1)


def write_pfm(path: str, image) -> None:
    """[h, w] or [h, w, 3] float image -> PFM ("Pf" / "PF", little endian, row 0 first = the bottom row of the format)."""
    a = np.ascontiguousarray(image, np.float32)
    with open(path, "wb") as f:
        f.write(("Pf" if a.ndim == 2 else "PF").encode() + f"\n{a.shape[1]} {a.shape[0]}\n-1.0\n".encode())
        f.write(a.tobytes())


def read_pfm(path: str):
    with open(path, "rb") as f:
        magic = f.readline().strip().decode()
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        a = np.frombuffer(f.read(), "<f4" if scale < 0 else ">f4")
    return a.reshape(h, w, 3).astype(np.float32) if magic == "PF" else a.reshape(h, w).astype(np.float32)


def cornell_normal_map(size: int = 16):
    """``cornell_box`` whose white material (id 0: floor, ceiling, back wall, boxes) carries a NORMAL MAP: a size x size fp32
    RGBA texture of tangent-space normals (a smooth egg-crate bump, tilted up to ~35 degrees; bilinear, wrap) addressed by
    per-quad UVs that tile it twice. Returns the cornell dict with `uvs` [V, 2], `normals` (flat per-vertex normals: every
    quad owns its vertices), `normal_texture` and `normal_map` (per material id: -1 or 0)."""
    c = cornell_box()
    nq = c["positions"].shape[0] // 4
    c["uvs"] = np.tile(np.array([[0, 0], [2, 0], [2, 2], [0, 2]], np.float32), (nq, 1))
    p, i = c["positions"], c["indices"]
    fn = np.cross(p[i[:, 1]] - p[i[:, 0]], p[i[:, 2]] - p[i[:, 0]])
    fn /= np.linalg.norm(fn, axis=1, keepdims=True)
    nrm = np.zeros_like(p)
    nrm[i[:, 0]] = fn; nrm[i[:, 1]] = fn; nrm[i[:, 2]] = fn
    c["normals"] = nrm.astype(np.float32)
    g = (np.arange(size, dtype=np.float32) + 0.5) / size * 2.0 * np.pi
    x, y = np.meshgrid(g, g)
    n = np.stack([0.7 * np.cos(x), 0.7 * np.cos(y), np.ones_like(x)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    t = np.concatenate([n, np.ones((size, size, 1))], axis=-1).astype(np.float32)
    c["normal_texture"] = dict(data=np.ascontiguousarray(t), interp="Linear", edge="Wrap")
    c["normal_map"] = np.array([0, -1, -1, -1], np.int32)
    return c


MIP_LEVEL_COLOURS = np.array([[0.80, 0.80, 0.80], [0.85, 0.25, 0.20], [0.20, 0.80, 0.25], [0.20, 0.30, 0.85], [0.85, 0.80, 0.15], [0.80, 0.20, 0.80]],
                             np.float32)


def mip_debug_texture(size: int = 32, seed: int = 9):
    """An fp32 RGBA texture with EXPLICIT mip levels down to 1 x 1, every level its own colour (MIP_LEVEL_COLOURS) under +-0.1 of
    per-texel noise: which level a read takes — and how two levels blend — shows in the image, which a filtered chain (whose
    levels all average to the same colour) would hide. Bilinear, wrap."""
    rng = np.random.default_rng(seed)
    levels = []
    k = 0
    while True:
        n = max(size >> k, 1)
        lv = np.ones((n, n, 4), np.float32)
        lv[..., :3] = MIP_LEVEL_COLOURS[k % len(MIP_LEVEL_COLOURS)] + (rng.random((n, n, 3), dtype=np.float32) - 0.5) * 0.2
        levels.append(np.clip(lv, 0.02, 0.98).astype(np.float32))
        if n == 1:
            break
        k += 1
    return dict(data=levels[0], mips=levels[1:], interp="Linear", edge="Wrap")


def cornell_mips(kind: str = "explicit", tile: float = 96.0):
    """Mip-mapped albedo on the white material of a Cornell box (floor, ceiling, back wall, tall / short box), UV0 tiling the texture
    `tile` times across every quad so that the ray cone's footprint spans several texels (the reference's host backend takes the mip
    level from the gradient in UV units, so only a strongly tiled texture leaves level 0 there).
      "explicit"      : cornell_box + mip_debug_texture (explicit levels, distinct colours)
      "gen_glossy"    : cornell_glossy (a (Mt)Refract and a (Mt)Unreal box: refracted / reflected cones) + a 64 x 64 unorm8 noise
                        texture whose chain is GENERATED (TracerParameters.genMips, Gaussian radius 2)
      "sphere_mirror" : cornell_sphere with the smooth-normal sphere made a (Mt)Reflect mirror (curvature widens the reflected cone)
                        + mip_debug_texture
    Returns the scene dict with `uvs`, `textures`, `albedo_texture` (per material id) and, for "gen_glossy", `gen_mips`."""
    if kind == "gen_glossy":
        c = cornell_glossy()
    elif kind == "sphere_mirror":
        c = cornell_sphere()
        m = c["material"].copy()
        sphere = np.arange(m.shape[0]) >= 24          # the sphere's triangles follow the 24 kept box triangles
        m[sphere] = 4
        c["material"] = m
        c["albedo"] = np.concatenate([c["albedo"], np.zeros((1, 3), np.float32)])
        c["material_type"] = np.array([0, 0, 0, 0, 1], np.uint8)
    else:
        c = cornell_box()
    nq = 18                                            # cornell_box: 18 quads own its 72 vertices (a sphere's vertices follow them)
    uv = np.zeros((c["positions"].shape[0], 2), np.float32)
    uv[:4 * nq] = np.tile(np.array([[0, 0], [tile, 0], [tile, tile], [0, tile]], np.float32), (nq, 1))
    c["uvs"] = uv
    if kind == "gen_glossy":
        rng = np.random.default_rng(12)
        t = rng.integers(0, 256, size=(64, 64, 4), dtype=np.uint8)
        # low-frequency structure, so that coarse levels differ from each other (pure noise filters to flat grey)
        g = (np.arange(64) // 16)
        t[..., 0] = np.clip(t[..., 0] // 2 + 40 * ((g[:, None] + g[None, :]) % 3), 0, 255)
        t[..., 1] = np.clip(t[..., 1] // 2 + 50 * ((g[:, None] * 2 + g[None, :]) % 3), 0, 255)
        t[..., 3] = 255
        c["textures"] = [dict(data=np.ascontiguousarray(t), interp="Linear", edge="Wrap")]
        c["gen_mips"] = ("Gaussian", 2.0)
    else:
        c["textures"] = [mip_debug_texture()]
    at = np.full(c["albedo"].shape[0], -1, np.int32); at[0] = 0
    c["albedo_texture"] = at
    return c
