"""Geometry sources for the scene layer: a small OBJ importer that honours the reference's ModelData contract
(src/scene/Models.cpp:117-175 — the reference delegates to Assimp with Triangulate | GenSmoothNormals |
JoinIdenticalVertices | CalcTangentSpace | FlipUVs, first mesh only) and deterministic procedural meshes used
where the reference's assets are absent from its snapshot (stanford_dragon / stanford_bunny / max_planck, see
.MISSING_LARGE_BLOBS) or cannot travel to the GPU box.

Importer rules (Assimp is an un-pinned third-party dependency; parity is unpinned, the rule is ours):
  * faces are fan-triangulated in file order  -> gl_PrimitiveID order;
  * one output vertex per distinct (v, vt, vn) triple, numbered by first use (JoinIdenticalVertices);
  * V is flipped (v' = 1 - v, FlipUVs); missing normals are generated as area-weighted smooth normals;
  * tangent / bitangent from the UV derivatives of the adjacent triangles, accumulated per output vertex;
    meshes without UVs get zero tangents exactly like the reference (src/scene/Models.cpp:144-152), which makes
    Disney / normal-mapped shading NaN on them — a warning is emitted.
All ModelData use indices == tbnsIndices == texIndices, as the reference's importer does (Models.cpp:163-171).
"""
import warnings

import numpy as np

from .scene import ModelData


def _tangent_frames(pos, uv, nrm, tris):
    """Per-vertex (T, B, N) columns from UV derivatives. pos (n,3), uv (n,2) or None, nrm (n,3), tris (m,3)."""
    n = pos.shape[0]
    T = np.zeros((n, 3), np.float64)
    B = np.zeros((n, 3), np.float64)
    if uv is not None and uv.shape[0] == n:
        p0, p1, p2 = pos[tris[:, 0]], pos[tris[:, 1]], pos[tris[:, 2]]
        w0, w1, w2 = uv[tris[:, 0]], uv[tris[:, 1]], uv[tris[:, 2]]
        e1, e2 = (p1 - p0).astype(np.float64), (p2 - p0).astype(np.float64)
        d1, d2 = (w1 - w0).astype(np.float64), (w2 - w0).astype(np.float64)
        det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
        ok = np.abs(det) > 1e-20
        r = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        t = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) * r[:, None]
        b = (e2 * d1[:, 0:1] - e1 * d2[:, 0:1]) * r[:, None]
        for k in range(3):
            np.add.at(T, tris[:, k], t)
            np.add.at(B, tris[:, k], b)

        def nz(v):
            l = np.linalg.norm(v, axis=1, keepdims=True)
            return np.where(l > 0, v / np.where(l > 0, l, 1.0), 0.0)
        T, B = nz(T), nz(B)
    tb = np.zeros((n, 3, 3), np.float32)
    tb[:, 0, :] = T
    tb[:, 1, :] = B
    tb[:, 2, :] = nrm
    return tb


def make_model(pos, uv, nrm, tris, tbn=None):
    """Pack arrays into a ModelData. pos (n,3); uv (n,2) or None; nrm (n,3); tris (m,3) uint32."""
    pos = np.ascontiguousarray(pos, np.float32)
    tris = np.ascontiguousarray(tris, np.uint32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    v4 = np.concatenate([pos, np.ones((pos.shape[0], 1), np.float32)], axis=1)
    if tbn is None:
        tbn = _tangent_frames(pos, uv, nrm, tris)
    idx = tris.reshape(-1)
    tc = np.zeros((0, 2), np.float32) if uv is None else np.ascontiguousarray(uv, np.float32)
    return ModelData(vertices=v4, indices=idx.copy(), tbns=np.ascontiguousarray(tbn, np.float32),
                     tbnsIndices=idx.copy(), texCoords=tc, texIndices=idx.copy())


def load_obj(path):
    """Minimal OBJ reader following the rules in the module docstring (first object/mesh only semantics are
    approximated by reading every face in the file, which is what the reference's single-mesh assets contain)."""
    vs, vts, vns = [], [], []
    keymap = {}
    out_pos, out_uv, out_nrm, tris = [], [], [], []
    has_uv = has_n = True
    with open(path, "r") as f:
        for line in f:
            s = line.split("#", 1)[0].split()
            if not s:
                continue
            if s[0] == "v":
                vs.append((float(s[1]), float(s[2]), float(s[3])))
            elif s[0] == "vt":
                vts.append((float(s[1]), float(s[2])))
            elif s[0] == "vn":
                vns.append((float(s[1]), float(s[2]), float(s[3])))
            elif s[0] == "f":
                corner = []
                for tok in s[1:]:
                    parts = tok.split("/")
                    vi = int(parts[0])
                    ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    vi = vi - 1 if vi > 0 else len(vs) + vi
                    ti = (ti - 1 if ti > 0 else len(vts) + ti) if ti else -1
                    ni = (ni - 1 if ni > 0 else len(vns) + ni) if ni else -1
                    if ti < 0:
                        has_uv = False
                    if ni < 0:
                        has_n = False
                    key = (vi, ti, ni)
                    if key not in keymap:
                        keymap[key] = len(out_pos)
                        out_pos.append(vs[vi])
                        out_uv.append((vts[ti][0], 1.0 - vts[ti][1]) if ti >= 0 else (0.0, 0.0))
                        out_nrm.append(vns[ni] if ni >= 0 else (0.0, 0.0, 0.0))
                    corner.append(keymap[key])
                for k in range(1, len(corner) - 1):
                    tris.append((corner[0], corner[k], corner[k + 1]))
    pos = np.array(out_pos, np.float32).reshape(-1, 3)
    tri = np.array(tris, np.uint32).reshape(-1, 3)
    uv = np.array(out_uv, np.float32).reshape(-1, 2) if has_uv else None
    if has_n:
        nrm = np.array(out_nrm, np.float32).reshape(-1, 3)
    else:
        nrm = smooth_normals(pos, tri)
    if uv is None:
        warnings.warn(f"{path}: no texture coordinates; tangents are zero and Disney/normal-mapped shading will be NaN "
                      "on this mesh (same as the reference, src/scene/Models.cpp:144-152)")
    return make_model(pos, uv, nrm, tri)


def smooth_normals(pos, tris):
    n = np.zeros(pos.shape, np.float64)
    fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]]).astype(np.float64)
    for k in range(3):
        np.add.at(n, tris[:, k], fn)
    l = np.linalg.norm(n, axis=1, keepdims=True)
    return (n / np.where(l > 0, l, 1.0)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# procedural meshes
# ---------------------------------------------------------------------------------------------------------------
def quad(p0, p1, p2, p3, uv=((0, 0), (1, 0), (1, 1), (0, 1)), normal=None):
    """One quad as two triangles (0,1,2), (0,2,3)."""
    pos = np.array([p0, p1, p2, p3], np.float32)
    if normal is None:
        nv = np.cross(pos[1] - pos[0], pos[2] - pos[0])
        normal = nv / np.linalg.norm(nv)
    nrm = np.tile(np.array(normal, np.float32), (4, 1))
    return make_model(pos, np.array(uv, np.float32), nrm, np.array([[0, 1, 2], [0, 2, 3]], np.uint32))


def cornell_box():
    """The classic closed box of models/cornell_box.obj: 8 corner positions of [-1,1] x [0,2] x [-1,1], six quads with
    inward normals, each mapped to one cell of a 2 x 3 texture atlas (see cornell_texture()). Face and corner order
    follow the asset so primitive ids agree: ceiling, z=+1 wall, x=-1 wall (red), floor, x=+1 wall (green), z=-1 wall."""
    c = {1: (1, 2, -1), 2: (1, 0, -1), 3: (1, 2, 1), 4: (1, 0, 1), 5: (-1, 2, -1), 6: (-1, 0, -1), 7: (-1, 2, 1),
         8: (-1, 0, 1)}
    faces = [((1, 3, 7, 5), (0, -1, 0), (0, 0)), ((4, 8, 7, 3), (0, 0, -1), (1, 0)), ((8, 6, 5, 7), (1, 0, 0), (0, 1)),
             ((6, 8, 4, 2), (0, 1, 0), (1, 1)), ((2, 4, 3, 1), (-1, 0, 0), (0, 2)), ((6, 2, 1, 5), (0, 0, 1), (1, 2))]
    pos, uv, nrm, tris = [], [], [], []
    third = [0.0, 0.333333, 0.666667, 1.0]
    for corners, n, (cx, cy) in faces:
        u0, u1 = 0.5 * cx, 0.5 * cx + 0.5
        v0, v1 = third[cy], third[cy + 1]
        cell = [(u0, v0), (u1, v0), (u1, v1), (u0, v1)]
        b = len(pos)
        for k, ci in enumerate(corners):
            pos.append(c[ci])
            uv.append((cell[k][0], 1.0 - cell[k][1]))     # FlipUVs
            nrm.append(n)
        tris += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    return make_model(np.array(pos, np.float32), np.array(uv, np.float32), np.array(nrm, np.float32),
                      np.array(tris, np.uint32))


def cornell_texture(width=1024, height=1536):
    """Stand-in for textures/cornell_texture.png (2 x 3 atlas: white / white, red / white, green / white), stored the way
    the reference stores file textures: rows flipped vertically on load (src/graphics/Image.cpp:14)."""
    img = np.full((height, width, 4), 255, np.uint8)
    h3, w2 = height // 3, width // 2
    img[h3:2 * h3, :w2, :3] = (255, 63, 63)
    img[2 * h3:, :w2, :3] = (119, 203, 63)
    return np.ascontiguousarray(img[::-1])


def cornell_light():
    """models/cornell_light.obj: a 0.47 x 0.38 panel just under the ceiling (y = 1.989) facing -y; two triangles
    (4,6,3), (4,5,6) of the asset's vertex list."""
    v = {3: (-0.24, 1.989, 0.16), 4: (-0.24, 1.989, -0.22), 5: (0.23, 1.989, -0.22), 6: (0.23, 1.989, 0.16)}
    t = {1: (0.0, 0.0), 2: (1.0, 0.0), 3: (0.0, 1.0), 4: (1.0, 1.0)}
    order = [(4, 1), (6, 4), (3, 3), (5, 2)]
    pos = np.array([v[a] for a, _ in order], np.float32)
    uv = np.array([(t[b][0], 1.0 - t[b][1]) for _, b in order], np.float32)
    nrm = np.tile(np.array((0, -1, 0), np.float32), (4, 1))
    return make_model(pos, uv, nrm, np.array([[0, 1, 2], [0, 3, 1]], np.uint32))


def uv_sphere(segments=32, rings=16, radius=1.0):
    pos, uv, nrm, tris = [], [], [], []
    for r in range(rings + 1):
        th = np.pi * r / rings
        for s in range(segments + 1):
            ph = 2 * np.pi * s / segments
            n = (np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph))
            pos.append((radius * n[0], radius * n[1], radius * n[2]))
            nrm.append(n)
            uv.append((s / segments, r / rings))
    w = segments + 1
    for r in range(rings):
        for s in range(segments):
            a, b, c, d = r * w + s, r * w + s + 1, (r + 1) * w + s + 1, (r + 1) * w + s
            if r != 0:
                tris.append((a, b, c))
            if r != rings - 1:
                tris.append((a, c, d))
    return make_model(np.array(pos, np.float32), np.array(uv, np.float32), np.array(nrm, np.float32),
                      np.array(tris, np.uint32))


def showroom(fillet_segments=12, span=24):
    """Stand-in for models/showroom.obj: an open cyclorama over [-1,1]^2 — floor at y = 0, walls at x = +1 and z = +1
    rising to y ~ 1.98, joined to the floor by quarter-circle fillets; no ceiling and no -x / -z walls, so paths can
    escape to the sky exactly as in the reference asset."""
    rf, top = 0.3, 1.977
    prof = [(-1.0, 0.0)]
    for i in range(fillet_segments + 1):
        a = -np.pi / 2 + (np.pi / 2) * i / fillet_segments
        prof.append((1.0 - rf + rf * np.cos(a), rf + rf * np.sin(a)))
    prof.append((1.0, top))
    prof = np.array(prof)
    # profile normal (pointing towards -axis / +y side = into the room)
    d = np.gradient(prof, axis=0)
    pn = np.stack([-d[:, 1], d[:, 0]], axis=1)
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    pos, uv, nrm, tris = [], [], [], []

    def sweep(axis):
        b0 = len(pos)
        m = prof.shape[0]
        for j in range(span + 1):
            s = -1.0 + 2.0 * j / span
            for i in range(m):
                a, y = prof[i]
                if axis == 0:
                    pos.append((a, y, s)); nrm.append((pn[i, 0], pn[i, 1], 0.0))
                else:
                    pos.append((s, y, a)); nrm.append((0.0, pn[i, 1], pn[i, 0]))
                uv.append((j / span, i / (m - 1)))
        for j in range(span):
            for i in range(m - 1):
                a, b, c, e = b0 + j * m + i, b0 + j * m + i + 1, b0 + (j + 1) * m + i + 1, b0 + (j + 1) * m + i
                if axis == 0:
                    tris.extend([(a, c, b), (a, e, c)])
                else:
                    tris.extend([(a, b, c), (a, c, e)])
    sweep(0)
    sweep(1)
    pos, nrm = np.array(pos, np.float32), np.array(nrm, np.float32)
    tri = np.array(tris, np.uint32)
    # make every geometric normal agree with the smooth normal (faces look into the room)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    flip = (fn * nrm[tri[:, 0]]).sum(axis=1) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    return make_model(pos, np.array(uv, np.float32), nrm, tri)


def _hash_u32(x):
    x = np.asarray(x, np.uint64) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    return (x ^ (x >> np.uint64(16))) & np.uint64(0xFFFFFFFF)


def value_noise3(p, seed):
    """Trilinear value noise on the integer lattice, integer-hash based (bit-reproducible everywhere)."""
    pf = np.floor(p)
    f = p - pf
    f = f * f * (3.0 - 2.0 * f)
    i = pf.astype(np.int64)
    out = np.zeros(p.shape[0], np.float64)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                h = _hash_u32((i[:, 0] + dx) * 73856093 ^ (i[:, 1] + dy) * 19349663 ^ (i[:, 2] + dz) * 83492791 ^ seed)
                v = h.astype(np.float64) / 4294967295.0
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                out += v * w
    return out


def torus_knot(p=3, q=7, n_along=10627, n_ring=41, tube=0.13, seed=0xD2A60, displacement=0.035, fit=None):
    """Stanford-dragon stand-in (asset absent, .MISSING_LARGE_BLOBS): a (p,q) torus-knot tube displaced by 3 octaves of
    value noise. n_along x n_ring quads x 2 = 871,414 triangles with the defaults — the Stanford dragon's count.
    UVs run along / around the tube (every mesh used with the Disney BSDF must carry UVs)."""
    t = np.linspace(0.0, 2.0 * np.pi, n_along, endpoint=False)

    def curve(t):
        r = 1.0 + 0.45 * np.cos(q * t)
        return np.stack([r * np.cos(p * t), 0.45 * np.sin(q * t), r * np.sin(p * t)], axis=1)
    c = curve(t)
    dt = 1e-4
    tan = curve(t + dt) - curve(t - dt)
    tan /= np.linalg.norm(tan, axis=1, keepdims=True)
    acc = curve(t + dt) - 2 * c + curve(t - dt)
    nb = acc - (acc * tan).sum(axis=1, keepdims=True) * tan
    nb /= np.linalg.norm(nb, axis=1, keepdims=True)
    bb = np.cross(tan, nb)
    a = np.linspace(0.0, 2.0 * np.pi, n_ring, endpoint=False)
    ca, sa = np.cos(a)[None, :, None], np.sin(a)[None, :, None]
    radial = nb[:, None, :] * ca + bb[:, None, :] * sa
    base = c[:, None, :] + tube * radial
    flat = base.reshape(-1, 3)
    disp = np.zeros(flat.shape[0])
    amp, freq = 1.0, 6.0
    for o in range(3):
        disp += amp * (value_noise3(flat * freq + 17.0 * o, seed + o) - 0.5)
        amp *= 0.5
        freq *= 2.1
    pos = flat + radial.reshape(-1, 3) * (displacement * disp)[:, None]
    if fit is not None:      # fit = (center xyz, max extent)
        lo, hi = pos.min(0), pos.max(0)
        s = fit[1] / (hi - lo).max()
        pos = (pos - 0.5 * (lo + hi)) * s + np.array(fit[0])
    iu = np.arange(n_along)[:, None]
    iv = np.arange(n_ring)[None, :]
    a0 = (iu * n_ring + iv).reshape(-1)
    a1 = (((iu + 1) % n_along) * n_ring + iv).reshape(-1)
    a2 = (((iu + 1) % n_along) * n_ring + (iv + 1) % n_ring).reshape(-1)
    a3 = (iu * n_ring + (iv + 1) % n_ring).reshape(-1)
    tris = np.stack([np.stack([a0, a2, a1], 1), np.stack([a0, a3, a2], 1)], axis=1).reshape(-1, 3).astype(np.uint32)
    pos32 = pos.astype(np.float32)
    nrm = smooth_normals(pos32, tris)
    # orient outwards
    out = (nrm * radial.reshape(-1, 3)).sum(axis=1) < 0
    if out.mean() > 0.5:
        tris = tris[:, [0, 2, 1]]
        nrm = -nrm
    uv = np.stack([np.broadcast_to(iu / n_along * 64.0, (n_along, n_ring)).reshape(-1),
                   np.broadcast_to(iv / n_ring, (n_along, n_ring)).reshape(-1)], axis=1).astype(np.float32)
    # analytic tangent frame: T along the curve, B around the ring
    T = np.broadcast_to(tan[:, None, :], (n_along, n_ring, 3)).reshape(-1, 3)
    B = np.cross(nrm, T)
    tbn = np.zeros((pos32.shape[0], 3, 3), np.float32)
    tbn[:, 0, :] = T
    tbn[:, 1, :] = B / np.maximum(np.linalg.norm(B, axis=1, keepdims=True), 1e-20)
    tbn[:, 2, :] = nrm
    return make_model(pos32, uv, nrm, tris, tbn=tbn)


def subdivided_blob(levels=5, seed=0x9A8, displacement=0.25, radius=1.0):
    """Noise-displaced icosphere (Max-Planck / bunny stand-in): 20 * 4^levels triangles, spherical UVs."""
    phi = (1 + 5 ** 0.5) / 2
    v = [(-1, phi, 0), (1, phi, 0), (-1, -phi, 0), (1, -phi, 0), (0, -1, phi), (0, 1, phi), (0, -1, -phi), (0, 1, -phi),
         (phi, 0, -1), (phi, 0, 1), (-phi, 0, -1), (-phi, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    pos = np.array(v, np.float64)
    pos /= np.linalg.norm(pos, axis=1, keepdims=True)
    tri = np.array(f, np.int64)
    for _ in range(levels):
        edges = {}
        plist = [pos]
        n = pos.shape[0]
        newtri = []

        def mid(a, b):
            nonlocal n
            k = (a, b) if a < b else (b, a)
            if k not in edges:
                m = pos[a] + pos[b]
                plist.append((m / np.linalg.norm(m))[None, :])
                edges[k] = n
                n += 1
            return edges[k]
        for a, b, c in tri:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            newtri += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        pos = np.concatenate(plist, axis=0)
        tri = np.array(newtri, np.int64)
    d = np.zeros(pos.shape[0])
    amp, freq = 1.0, 2.0
    for o in range(3):
        d += amp * (value_noise3(pos * freq + 11.0 * o, seed + o) - 0.5)
        amp *= 0.5
        freq *= 2.0
    p = (pos * (radius * (1.0 + displacement * d))[:, None]).astype(np.float32)
    tri = tri.astype(np.uint32)
    nrm = smooth_normals(p, tri)
    uv = np.stack([np.arctan2(pos[:, 2], pos[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(pos[:, 1], -1, 1)) / np.pi],
                  axis=1).astype(np.float32)
    T = np.cross(np.array([0.0, 1.0, 0.0])[None, :], nrm)
    tl = np.linalg.norm(T, axis=1, keepdims=True)
    T = np.where(tl > 1e-6, T / np.maximum(tl, 1e-20), np.array([1.0, 0.0, 0.0])[None, :])
    B = np.cross(nrm, T)
    tbn = np.zeros((p.shape[0], 3, 3), np.float32)
    tbn[:, 0, :], tbn[:, 1, :], tbn[:, 2, :] = T, B, nrm
    return make_model(p, uv, nrm, tri, tbn=tbn)
