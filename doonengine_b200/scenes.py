"""Procedural maps of BASELINE.json's configs, as pure functions of integer coordinates.

Every generator yields whole chunks -- (map_pos, voxels[8,8,8,2] uint32 indexed [x][y][z] -> (normal word,
albedo word)) -- so the same stream can be fed to the CUDA engine, to the oracle and to the reference host code
through their common ``set_chunk`` call.  Only numpy integer / float32 arithmetic is used, so every rank of a
multi-GPU run builds the identical map.

    terrain(tiles, seed)        config 2 / 4: fbm height field, solid below the surface, 1 % emissive surface voxels
    sparse_balls(tiles, ...)    config 3: ~20 % of the tiles hold a solid ball (diffuse / glossy / emissive)
    dense_corridors(tiles)      config 5: everything solid, mirror-like material, a lattice of empty corridors
    mixed_materials(tiles)      small test scene with every material kind incl. glass (refraction in the draw pass)
"""
import numpy as np

EMPTY = np.uint32(0xFFFFFFFF)


def pcg_hash(x):
    """pcg-style 32-bit integer hash, vectorised (uint32 in, uint32 out)."""
    x = np.asarray(x, dtype=np.uint64)
    state = (x * np.uint64(747796405) + np.uint64(2891336453)) & np.uint64(0xFFFFFFFF)
    shift = ((state >> np.uint64(28)) + np.uint64(4)) & np.uint64(31)
    word = (((state >> shift) ^ state) * np.uint64(277803737)) & np.uint64(0xFFFFFFFF)
    return (((word >> np.uint64(22)) ^ word) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def hash3(x, y, z, seed=0):
    x = np.asarray(x, dtype=np.uint64)
    y = np.asarray(y, dtype=np.uint64)
    z = np.asarray(z, dtype=np.uint64)
    k = (x * np.uint64(73856093)) ^ (y * np.uint64(19349663)) ^ (z * np.uint64(83492791)) ^ np.uint64(seed)
    return pcg_hash(k & np.uint64(0xFFFFFFFF))


def compress_normal(n):
    """DN_compress_voxel's normal packing (reference voxel.c:1293-1297): byte = (int(c*255)+255)/2, c clamped to [-1,1]."""
    c = np.clip(np.asarray(n, dtype=np.float32), np.float32(-1.0), np.float32(1.0))
    b = ((c * np.float32(255.0)).astype(np.int32) + 255) // 2
    return b.astype(np.uint32)


def normal_word(material, n):
    b = compress_normal(n)
    return (np.asarray(material, dtype=np.uint32) << np.uint32(24)) | (b[..., 0] << np.uint32(16)) | (b[..., 1] << np.uint32(8)) | b[..., 2]


def albedo_word(r, g, b):
    return (np.asarray(r, np.uint32) << np.uint32(24)) | (np.asarray(g, np.uint32) << np.uint32(16)) | (np.asarray(b, np.uint32) << np.uint32(8))


def max_normalise(v):
    """normals are max-component-normalised, not unit (reference voxelShapes.c:145-146)."""
    v = np.asarray(v, dtype=np.float32)
    m = np.max(np.abs(v), axis=-1, keepdims=True)
    m = np.where(m == 0, np.float32(1.0), m)
    return (v / m).astype(np.float32)


def default_materials():
    """material table used by the synthetic scenes; fields as DNmaterial (voxel.h:82-94)."""
    from . import MATERIAL_DT
    m = np.zeros(256, MATERIAL_DT)
    m["opacity"] = 1.0
    m["refractIndex"] = 1.0
    # 0 diffuse, 1 mirror-like (reflects sky), 2 emissive, 3 glossy, 4 glass, 5 specular-heavy (config 5)
    m[1]["specular"], m[1]["reflectType"], m[1]["shininess"] = 1.0, 1, 100
    m[2]["emissive"] = 1
    m[3]["specular"], m[3]["reflectType"], m[3]["shininess"] = 0.7, 0, 3
    m[4]["opacity"], m[4]["refractIndex"] = 0.5, 1.52
    m[5]["specular"], m[5]["reflectType"], m[5]["shininess"] = 0.8, 1, 3
    return m


# ---------------------------------------------------------------------------------------------------------------
def _value_noise(px, pz, seed):
    """bilinear value noise on the integer lattice, float32."""
    x0 = np.floor(px).astype(np.int64)
    z0 = np.floor(pz).astype(np.int64)
    fx = (px - x0).astype(np.float32)
    fz = (pz - z0).astype(np.float32)

    def lat(ix, iz):
        return (hash3(ix & 0xFFFF, 0, iz & 0xFFFF, seed) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)

    sx = fx * fx * (np.float32(3.0) - np.float32(2.0) * fx)
    sz = fz * fz * (np.float32(3.0) - np.float32(2.0) * fz)
    a = lat(x0, z0) * (1 - sx) + lat(x0 + 1, z0) * sx
    b = lat(x0, z0 + 1) * (1 - sx) + lat(x0 + 1, z0 + 1) * sx
    return (a * (1 - sz) + b * sz).astype(np.float32)


def terrain_height(nx, nz, seed=1234, base=96.0, amp=160.0, octaves=5, scale=None):
    """height in voxels of column (x, z): base + amp * fbm(x/nx', z/nz') (SURVEY.md 8d C2 uses 96 + 160 fbm on 512^2)."""
    xs, zs = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(nz, dtype=np.float32), indexing="ij")
    s = np.float32(scale if scale else max(nx, nz))
    f = np.zeros((nx, nz), np.float32)
    ampl, freq, norm = np.float32(1.0), np.float32(4.0), np.float32(0.0)
    for o in range(octaves):
        f += ampl * _value_noise(xs / s * freq, zs / s * freq, seed + o)
        norm += ampl
        ampl *= np.float32(0.5)
        freq *= np.float32(2.0)
    f /= norm
    return np.float32(base) + np.float32(amp) * f


def terrain(tiles=(64, 64, 64), seed=1234, base=None, amp=None):
    """config 2 map: solid below an fbm height field.  Yields (map_pos, voxels)."""
    tx, ty, tz = tiles
    nx, ny, nz = tx * 8, ty * 8, tz * 8
    base = (ny * 96.0 / 512.0) if base is None else base
    amp = (ny * 160.0 / 512.0) if amp is None else amp
    h = terrain_height(nx, nz, seed, base, amp)
    hi = np.clip(np.floor(h).astype(np.int32), 1, ny)            # voxels y < hi are solid

    # per-column normal from the height gradient
    gx = np.zeros_like(h)
    gz = np.zeros_like(h)
    gx[1:-1, :] = (h[2:, :] - h[:-2, :]) * np.float32(0.5)
    gz[:, 1:-1] = (h[:, 2:] - h[:, :-2]) * np.float32(0.5)
    n = max_normalise(np.stack([-gx, np.ones_like(h), -gz], axis=-1))
    nbytes = compress_normal(n)                                    # [nx, nz, 3]
    nword_xz = (nbytes[..., 0] << np.uint32(16)) | (nbytes[..., 1] << np.uint32(8)) | nbytes[..., 2]

    ys = np.arange(ny, dtype=np.int64)
    for cz in range(tz):
        for cx in range(tx):
            hcol = hi[cx * 8:cx * 8 + 8, cz * 8:cz * 8 + 8]       # [x, z]
            layers = min(ty, (int(hcol.max()) + 7) // 8)
            xs = np.arange(cx * 8, cx * 8 + 8, dtype=np.int64)
            zs = np.arange(cz * 8, cz * 8 + 8, dtype=np.int64)
            yy = ys[:layers * 8]
            # the whole column of chunks at once: arrays are [x, y, z] with y spanning `layers` chunks
            solid = yy[None, :, None] < hcol[:, None, :]
            X, Y, Z = np.meshgrid(xs, yy, zs, indexing="ij")
            hsh = hash3(X, Y, Z, seed)
            surface = (Y >= hcol[:, None, :] - 1)
            emissive = surface & ((hsh % np.uint32(100)) == 0)
            mat = np.where(emissive, np.uint32(2), np.uint32(0))
            # height bands: sand / grass / rock / snow, jittered per voxel
            t = Y.astype(np.float32) / np.float32(ny)
            jit = ((hsh >> np.uint32(8)) & np.uint32(31)).astype(np.int32) - 16
            r = np.where(t < 0.22, 194, np.where(t < 0.34, 86, np.where(t < 0.46, 120, 235))) + jit
            g = np.where(t < 0.22, 178, np.where(t < 0.34, 152, np.where(t < 0.46, 112, 238))) + jit
            b = np.where(t < 0.22, 128, np.where(t < 0.34, 66, np.where(t < 0.46, 104, 240))) + jit
            r, g, b = (np.clip(c, 32, 240).astype(np.uint32) for c in (r, g, b))
            col = np.empty((8, layers * 8, 8, 2), np.uint32)
            col[..., 0] = np.where(solid, (mat << np.uint32(24)) | nword_xz[cx * 8:cx * 8 + 8, None, cz * 8:cz * 8 + 8], EMPTY)
            col[..., 1] = albedo_word(r, g, b)
            for cy in range(layers):
                if solid[:, cy * 8:cy * 8 + 8, :].any():
                    yield (cx, cy, cz), np.ascontiguousarray(col[:, cy * 8:cy * 8 + 8])


def terrain_camera(tiles=(64, 64, 64)):
    """camera of config 2 scaled to the map: just outside the -x/-z corner, looking down the diagonal."""
    ty = tiles[1]
    return dict(camPos=(-2.0, ty * 40.0 / 64.0, -2.0), camOrient=(25.0, 45.0, 0.0), camFOV=90.0)


# ---------------------------------------------------------------------------------------------------------------
_BALL = None


def _ball_offsets():
    global _BALL
    if _BALL is None:
        c = np.arange(8, dtype=np.float32) - np.float32(3.5)
        X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
        _BALL = (X, Y, Z, np.sqrt(X * X + Y * Y + Z * Z))
    return _BALL


def sparse_balls(tiles=(256, 256, 256), fill=0.2, seed=99, zrange=None):
    """config 3 map: tile occupied iff hash < fill; an occupied tile is a solid ball of radius 2.5..4 voxels."""
    tx, ty, tz = tiles
    X, Y, Z, R = _ball_offsets()
    n = max_normalise(np.stack([X, Y, Z], axis=-1))
    threshold = np.uint32(min(0xFFFFFFFF, int(fill * 4294967296.0)))
    z0, z1 = zrange if zrange else (0, tz)
    xs = np.arange(tx, dtype=np.int64)
    for cz in range(z0, z1):
        for cy in range(ty):
            hrow = hash3(xs, cy, cz, seed)
            for cx in np.nonzero(hrow < threshold)[0]:
                hv = int(pcg_hash(np.uint32(hrow[cx])))
                radius = np.float32(2.5 + 0.5 * (hv & 3))
                kind = (hv >> 2) % 10
                mat = 0 if kind < 7 else (3 if kind < 9 else 2)
                solid = R <= radius
                hh = hash3(np.arange(512).reshape(8, 8, 8), hv & 0xFFFF, 7, seed)
                r = 32 + (hh & np.uint32(0xFF)) % np.uint32(209)
                g = 32 + ((hh >> np.uint32(8)) & np.uint32(0xFF)) % np.uint32(209)
                b = 32 + ((hh >> np.uint32(16)) & np.uint32(0xFF)) % np.uint32(209)
                v = np.empty((8, 8, 8, 2), np.uint32)
                v[..., 0] = np.where(solid, normal_word(np.uint32(mat), n), EMPTY)
                v[..., 1] = albedo_word(r, g, b)
                yield (int(cx), cy, cz), v


def sparse_camera(tiles=(256, 256, 256)):
    tx, ty, tz = tiles
    return dict(camPos=(tx / 2.0, ty / 2.0, -4.0), camOrient=(0.0, 0.0, 0.0), camFOV=90.0)


# ---------------------------------------------------------------------------------------------------------------
def dense_corridors(tiles=(128, 128, 128), period=4, material=5, seed=5):
    """config 5 map: every voxel solid with a specular material, except empty corridor tiles (one tile wide, every
    `period` tiles along each axis) that make all chunks reachable."""
    tx, ty, tz = tiles
    idx = np.arange(512).reshape(8, 8, 8)
    for cz in range(tz):
        for cy in range(ty):
            for cx in range(tx):
                on_axis = (cx % period == 1) + (cy % period == 1) + (cz % period == 1)
                if on_axis >= 2:
                    continue  # corridor tile
                hh = hash3(idx, cx + 131 * cy, cz, seed)
                r = 32 + (hh & np.uint32(0xFF)) % np.uint32(209)
                g = 32 + ((hh >> np.uint32(8)) & np.uint32(0xFF)) % np.uint32(209)
                b = 32 + ((hh >> np.uint32(16)) & np.uint32(0xFF)) % np.uint32(209)
                # normals point away from the chunk centre (faces of the corridor walls)
                X, Y, Z, _ = _ball_offsets()
                n = max_normalise(np.stack([X, Y, Z], axis=-1))
                v = np.empty((8, 8, 8, 2), np.uint32)
                v[..., 0] = normal_word(np.uint32(material), n)
                v[..., 1] = albedo_word(r, g, b)
                yield (cx, cy, cz), v


def dense_camera(tiles=(128, 128, 128), period=4):
    tx, ty, tz = tiles
    return dict(camPos=(1.5, 1.5, 0.5), camOrient=(5.0, 8.0, 0.0), camFOV=90.0)


# ---------------------------------------------------------------------------------------------------------------
def mixed_materials(tiles=(6, 4, 6), seed=3):
    """small test scene: a diffuse floor with a mirror block, an emissive pillar, a glossy ball and a glass slab."""
    tx, ty, tz = tiles
    nx, ny, nz = tx * 8, ty * 8, tz * 8
    mat = np.full((nx, ny, nz), 255, np.uint32)
    nrm = np.zeros((nx, ny, nz, 3), np.float32)
    X, Y, Z = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")

    floor = Y < 5 + ((hash3(X // 3, 0, Z // 3, seed) & np.uint32(3)) == 0)
    mat[floor] = 0
    nrm[floor] = (0.0, 1.0, 0.0)

    box = (X >= 10) & (X < 18) & (Y >= 5) & (Y < 14) & (Z >= 10) & (Z < 20)
    mat[box] = 1
    c = np.stack([X - 13.5, Y - 9.0, Z - 14.5], axis=-1).astype(np.float32)
    nrm[box] = max_normalise(c[box])

    pillar = (np.abs(X - 30) <= 1) & (Y >= 5) & (Y < 22) & (np.abs(Z - 12) <= 1)
    mat[pillar] = 2
    nrm[pillar] = max_normalise(np.stack([X - 30.0, np.zeros_like(X, dtype=np.float32), Z - 12.0 + 0.25], axis=-1).astype(np.float32)[pillar])

    d = np.stack([X - 24.0, Y - 12.0, Z - 30.0], axis=-1).astype(np.float32)
    ball = np.sqrt((d * d).sum(-1)) <= 6.5
    mat[ball] = 3
    nrm[ball] = max_normalise(d[ball])

    slab = (X >= 6) & (X < 20) & (Y >= 5) & (Y < 18) & (Z >= 28) & (Z < 31)
    mat[slab] = 4
    nrm[slab] = (0.0, 0.0, -1.0)

    h = hash3(X, Y, Z, seed)
    r = np.where(mat == 0, 96 + (h & np.uint32(63)), np.where(mat == 4, 120, 200 + (h & np.uint32(31))))
    g = np.where(mat == 0, 128 + ((h >> np.uint32(6)) & np.uint32(63)), np.where(mat == 4, 200, 180 + ((h >> np.uint32(5)) & np.uint32(31))))
    b = np.where(mat == 0, 80 + ((h >> np.uint32(12)) & np.uint32(63)), np.where(mat == 4, 230, 160 + ((h >> np.uint32(10)) & np.uint32(31))))
    nw = np.where(mat == 255, EMPTY, normal_word(mat, nrm))
    aw = albedo_word(r, g, b)
    for cz in range(tz):
        for cy in range(ty):
            for cx in range(tx):
                sl = (slice(cx * 8, cx * 8 + 8), slice(cy * 8, cy * 8 + 8), slice(cz * 8, cz * 8 + 8))
                if (mat[sl] == 255).all():
                    continue
                v = np.empty((8, 8, 8, 2), np.uint32)
                v[..., 0] = nw[sl]
                v[..., 1] = aw[sl]
                yield (cx, cy, cz), v


def mixed_camera(tiles=(6, 4, 6)):
    return dict(camPos=(-0.6, 2.6, -0.8), camOrient=(22.0, 42.0, 0.0), camFOV=90.0)


# ---------------------------------------------------------------------------------------------------------------
# native twins of sparse_balls / dense_corridors (csrc/scenegen.c -> libdoon_scenes.so): whole z-slabs of chunks at a time,
# bit-identical to the numpy generators above (tests/test_abi_host.py), fast enough for the full-size configs 3 and 5
_native = None


def native_lib():
    global _native
    if _native is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdoon_scenes.so")
        if not os.path.exists(path):
            raise RuntimeError("%s is missing: run `python -m doonengine_b200.build`" % path)
        L = C.CDLL(path)
        L.dnscene_sparse_balls.restype = C.c_size_t
        L.dnscene_sparse_balls.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.dnscene_dense_corridors.restype = C.c_size_t
        L.dnscene_dense_corridors.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _native = L
    return _native


def _normal24_table():
    X, Y, Z, _ = _ball_offsets()
    b = compress_normal(max_normalise(np.stack([X, Y, Z], axis=-1)))
    return np.ascontiguousarray(((b[..., 0] << np.uint32(16)) | (b[..., 1] << np.uint32(8)) | b[..., 2]).astype(np.uint32).reshape(512))


def native_slabs(scene, tiles, slab=1, fill=0.2, seed=None, period=4, material=5):
    """yields (positions int32 [n,3], voxels uint32 [n,8,8,8,2]) for `slab` z-layers of tiles at a time, in the order of the numpy
    generators (z, then y, then x).  scene: "sparse" | "dense"."""
    L = native_lib()
    t = np.array(tiles, np.uint32)
    n24 = _normal24_table()
    dist = np.ascontiguousarray(_ball_offsets()[3].astype(np.float32).reshape(512))
    for z0 in range(0, tiles[2], slab):
        z1 = min(tiles[2], z0 + slab)
        if scene == "sparse":
            threshold = min(0xFFFFFFFF, int(fill * 4294967296.0))
            sd = 99 if seed is None else seed
            call = lambda p, v, cap: L.dnscene_sparse_balls(t.ctypes.data, threshold, sd, z0, z1, dist.ctypes.data, n24.ctypes.data, p, v, cap)
        elif scene == "dense":
            sd = 5 if seed is None else seed
            call = lambda p, v, cap: L.dnscene_dense_corridors(t.ctypes.data, period, material, sd, z0, z1, n24.ctypes.data, p, v, cap)
        else:
            raise ValueError(scene)
        n = int(call(None, None, 0))
        if n == 0:
            continue
        pos = np.empty((n, 3), np.int32)
        vox = np.empty((n, 8, 8, 8, 2), np.uint32)
        call(pos.ctypes.data, vox.ctypes.data, n)
        yield pos, vox


def native_count(scene, tiles, **kw):
    """number of chunks the native generator will produce (cheap: hashes only for "sparse", arithmetic for "dense")."""
    L = native_lib()
    t = np.array(tiles, np.uint32)
    if scene == "sparse":
        threshold = min(0xFFFFFFFF, int(kw.get("fill", 0.2) * 4294967296.0))
        return int(L.dnscene_sparse_balls(t.ctypes.data, threshold, kw.get("seed", 99), 0, tiles[2], None, None, None, None, 0))
    return int(L.dnscene_dense_corridors(t.ctypes.data, kw.get("period", 4), kw.get("material", 5), kw.get("seed", 5), 0, tiles[2], None, None, None, 0))


def build_native(engine, scene, tiles, materials=None, slab=2, **params):
    """the large maps through the bulk path: native generator -> engine.set_chunks, a few z-layers at a time."""
    engine.materials()[:] = default_materials() if materials is None else materials
    n = 0
    for pos, vox in native_slabs(scene, tiles, slab=slab):
        n += engine.set_chunks(pos, vox)
    if params:
        engine.set_params(**params)
    return n


def build(engine, chunks, materials=None, **params):
    """feed a chunk stream and a material table to any engine (CUDA, oracle or reference); returns the chunk count."""
    if materials is None:
        materials = default_materials()
    engine.materials()[:] = materials
    n = 0
    for pos, vox in chunks:
        engine.set_chunk(pos, vox)
        n += 1
    if params:
        engine.set_params(**params)
    return n
