"""ctypes bindings for the oracle -- TEST INFRASTRUCTURE ONLY.

Two engines with one interface (the frame protocol of SURVEY.md 8d):

  OracleEngine   oracle/liboracle.so       restated host (host_cpu.c) + Oracle B (shader_cpu.c)
  RefEngine      oracle/_ref/libdoon_ref.so the reference's OWN voxel.c behind the fake-GL shim
                                            (fake_gl.c), with Oracle B executing the dispatches

Nothing under doonengine_b200/ may import this module; only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() do.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libdoon_ref.so")

HANDLE_DT = np.dtype([("flags", "<u4"), ("lastUsed", "<u4"), ("voxelIndex", "<u4")])
CHUNK_DT = np.dtype([("pos", "<i4", 3), ("numIndirectSamples", "<u4"), ("partialCounts", "<u4", 3),
                     ("bitMask", "<u4", 16), ("pad", "<u4")])
VOXEL_DT = np.dtype([("normal", "<u4"), ("albedo", "<u4"), ("specLight", "<u4"), ("diffuseLight", "<u4")])
MATERIAL_DT = np.dtype([("pad", "<f4", 2), ("emissive", "<u4"), ("opacity", "<f4"), ("refractIndex", "<f4"),
                        ("specular", "<f4"), ("reflectType", "<u4"), ("shininess", "<u4")])
HIT_DT = np.dtype([("status", "<i4"), ("mapIndex", "<u4"), ("localIndex", "<u4"), ("recordIndex", "<u4")])
HOST_CHUNK_DT = np.dtype([("pos", "<i4", 3), ("updated", "u1"), ("_pad", "u1", 3), ("numVoxels", "<u4"),
                          ("numVoxelsGpu", "<u4"), ("voxels", "<u4", (8, 8, 8, 2))])
assert HANDLE_DT.itemsize == 12 and CHUNK_DT.itemsize == 96 and VOXEL_DT.itemsize == 16
assert MATERIAL_DT.itemsize == 32 and HOST_CHUNK_DT.itemsize == 4120

COUNTER_FIELDS = ("rays", "tiles", "chunks", "voxelSteps", "records", "voxelsLit", "pixels")


class OrbCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_FIELDS]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_FIELDS}


class OrbUniforms(C.Structure):
    _fields_ = [("mapSize", C.c_uint32 * 3), ("useCubemap", C.c_uint32),
                ("skyGradientBot", C.c_float * 3), ("skyGradientTop", C.c_float * 3),
                ("sunStrength", C.c_float * 3), ("ambientStrength", C.c_float * 3),
                ("viewMode", C.c_uint32), ("composeRasterized", C.c_uint32),
                ("invViewMat", C.c_float * 16), ("invCenteredViewMat", C.c_float * 16),
                ("invProjectionMat", C.c_float * 16),
                ("time", C.c_float), ("numDiffuseSamples", C.c_uint32), ("maxDiffuseSamples", C.c_uint32),
                ("diffuseBounceLimit", C.c_uint32), ("specularBounceLimit", C.c_uint32),
                ("sunDir", C.c_float * 3), ("shadowSoftness", C.c_float), ("camPos", C.c_float * 3)]


class Params(C.Structure):
    """camera .. sky block shared by every engine (voxel.h:120-138)."""
    _fields_ = [("camPos", C.c_float * 3), ("camOrient", C.c_float * 3), ("camFOV", C.c_float),
                ("camViewMode", C.c_uint32),
                ("sunDir", C.c_float * 3), ("sunStrength", C.c_float * 3), ("ambientLightStrength", C.c_float * 3),
                ("diffuseBounceLimit", C.c_uint32), ("specBounceLimit", C.c_uint32), ("shadowSoftness", C.c_float),
                ("skyGradientBot", C.c_float * 3), ("skyGradientTop", C.c_float * 3)]


PARAM_NAMES = [f[0] for f in Params._fields_]


class DNvolume(C.Structure):
    """the reference's public struct, voxel.h:97-142 (232 bytes on x86-64)."""
    _fields_ = [("glMapBufferID", C.c_uint32), ("glChunkBufferID", C.c_uint32), ("glVoxelBufferID", C.c_uint32),
                ("mapSize", C.c_uint32 * 3),
                ("chunkCap", C.c_size_t), ("nextChunk", C.c_size_t), ("voxelCap", C.c_size_t),
                ("numVoxelNodes", C.c_size_t), ("numLightingRequests", C.c_size_t), ("lightingRequestCap", C.c_size_t),
                ("map", C.c_void_p), ("chunks", C.c_void_p), ("materials", C.c_void_p),
                ("lightingRequests", C.c_void_p), ("gpuVoxelLayout", C.c_void_p),
                ("camPos", C.c_float * 3), ("camOrient", C.c_float * 3), ("camFOV", C.c_float),
                ("camViewMode", C.c_uint32),
                ("sunDir", C.c_float * 3), ("sunStrength", C.c_float * 3), ("ambientLightStrength", C.c_float * 3),
                ("diffuseBounceLimit", C.c_uint32), ("specBounceLimit", C.c_uint32), ("shadowSoftness", C.c_float),
                ("useCubemap", C.c_bool), ("glCubemapTex", C.c_uint32),
                ("skyGradientBot", C.c_float * 3), ("skyGradientTop", C.c_float * 3),
                ("frameNum", C.c_uint32), ("lastTime", C.c_float)]


assert C.sizeof(DNvolume) == 232


def build(force=False):
    """compile liboracle.so and (when /root/reference exists) _ref/libdoon_ref.so."""
    if force or not os.path.exists(LIB_ORACLE) or (os.path.isdir("/root/reference") and not (os.path.exists(LIB_REF) and os.path.exists(os.path.join(HERE, "_ref", "libglsl_ref.so")))):
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))
    return LIB_ORACLE


def have_ref():
    return os.path.exists(LIB_REF)


def set_num_threads(n=0):
    """OpenMP team size of every oracle dispatch in this process (one libgomp serves liboracle.so and libdoon_ref.so).
    n <= 0: every online processor -- torchrun exports OMP_NUM_THREADS=1, which the reference arm of bench.py must not inherit."""
    build()
    L = C.CDLL(LIB_ORACLE)
    L.orb_set_num_threads.argtypes = [C.c_int]
    L.orb_num_threads.restype = C.c_int
    L.orb_set_num_threads(int(n))
    return int(L.orb_num_threads())


def _view(ptr, dtype, count):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    buf = (C.c_char * (dtype.itemsize * count)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=count)


def set_params(target, **kw):
    for k, v in kw.items():
        cur = getattr(target, k)
        if hasattr(cur, "__len__"):
            for i, x in enumerate(v):
                cur[i] = x
        else:
            setattr(target, k, v)


def get_params(src):
    out = {}
    for k in PARAM_NAMES:
        v = getattr(src, k)
        out[k] = tuple(v) if hasattr(v, "__len__") else v
    return out


def fill_local_indices(hits, map_view, chunk_view):
    """the reference's shaders, run as they are (GlslEngine / RefEngine(glsl=True)), report a first hit as (tile, record index); the
    voxel's index inside its chunk follows from the chunk's bit mask: the record's rank among the chunk's records = the rank of the
    voxel's bit among the set bits (get_voxel_index, voxelShared.comp:150-169)."""
    flat = hits.reshape(-1)
    for k in np.nonzero((flat["status"] == 2) & (flat["localIndex"] == 0xFFFFFFFF))[0]:
        tile = int(flat["mapIndex"][k])
        rank = int(flat["recordIndex"][k]) - int(map_view["voxelIndex"][tile])
        bits = np.unpackbits(chunk_view["bitMask"][tile].astype("<u4").view(np.uint8), bitorder="little")
        flat["localIndex"][k] = int(np.nonzero(bits)[0][rank])
    return hits


class _EngineBase:
    """shared helpers: state export in a layout-independent form."""

    def num_tiles(self):
        sx, sy, sz = self.map_size
        return sx * sy * sz

    def export_state(self):
        """dict mapIndex -> (state, visible, header(pos, samples, partialCounts, bitMask), records[n,4])"""
        m = self.map_view()
        ch = self.chunk_view()
        vx = self.voxel_view()
        out = {}
        for idx in np.nonzero(m["flags"] & 3)[0]:
            idx = int(idx)
            fl = int(m["flags"][idx])
            if (fl & 3) != 2:
                out[idx] = (fl & 3, (fl >> 2) & 1, None, None)
                continue
            h = ch[idx]
            n = int(sum(bin(int(w)).count("1") for w in h["bitMask"]))
            base = int(m["voxelIndex"][idx])
            rec = np.stack([vx["normal"][base:base + n], vx["albedo"][base:base + n],
                            vx["specLight"][base:base + n], vx["diffuseLight"][base:base + n]], axis=1).copy()
            hdr = (tuple(int(x) for x in h["pos"]), int(h["numIndirectSamples"]),
                   tuple(int(x) for x in h["partialCounts"]), tuple(int(x) for x in h["bitMask"]))
            out[idx] = (2, (fl >> 2) & 1, hdr, rec)
        return out

    def frame(self, w, h, time, num_diffuse=1, max_diffuse=1000, split=1, aspect=None, want_hits=False):
        """draw -> sync(READ_WRITE) -> update_lighting (main.c:503-505)."""
        res = self.draw(w, h, aspect=aspect, want_hits=want_hits)
        self.sync(2, split)
        self.update_lighting(num_diffuse, max_diffuse, time)
        return res


class OracleEngine(_EngineBase):
    def __init__(self, map_size=None, min_chunks=256, voxvol=None):
        build()
        L = C.CDLL(LIB_ORACLE)
        self.L = L
        L.orh_create.restype = C.c_void_p
        L.orh_create.argtypes = [C.c_uint32] * 4
        L.orh_load_voxvol.restype = C.c_void_p
        L.orh_load_voxvol.argtypes = [C.c_char_p, C.c_uint32]
        for name in ("orh_map", "orh_gpu_chunks", "orh_voxels", "orh_requests", "orh_materials", "orh_cpu_flags",
                     "orh_cpu_chunk_index", "orh_cpu_chunks", "orh_draw_counters", "orh_light_counters"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orh_voxel_top", "orh_num_requests", "orh_chunk_cap", "orh_upload_bytes"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_void_p]
        L.orh_destroy.argtypes = [C.c_void_p]
        L.orh_set_voxel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orh_set_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orh_sync.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orh_view_projection.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.orh_draw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orh_update_lighting.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.orh_get_params.argtypes = [C.c_void_p, C.c_void_p]
        L.orh_light_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orh_light_commit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orh_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.orh_draw_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orh_light_uniforms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.orh_reset_counters.argtypes = [C.c_void_p]
        L.orh_compress_voxel.argtypes = [C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_num_threads.restype = C.c_int
        if voxvol is not None:
            self.v = L.orh_load_voxvol(os.fsencode(voxvol), min_chunks)
            if not self.v:
                raise IOError("cannot load %s" % voxvol)
        else:
            self.v = L.orh_create(map_size[0], map_size[1], map_size[2], min_chunks)
        ms = _view(self.v, np.dtype("<u4"), 3)
        self.map_size = (int(ms[0]), int(ms[1]), int(ms[2]))

    def close(self):
        if self.v:
            self.L.orh_destroy(self.v)
            self.v = None

    # --- parameters ---
    def params(self):
        p = Params()
        self.L.orh_get_params(self.v, C.byref(p))
        return p

    def set_params(self, **kw):
        p = self.params()
        set_params(p, **kw)
        self.L.orh_set_params(self.v, C.byref(p))

    def get_params(self):
        return get_params(self.params())

    def materials(self):
        return _view(self.L.orh_materials(self.v), MATERIAL_DT, 256)

    # --- edits ---
    def set_voxel(self, map_pos, chunk_pos, normal_word, albedo_word):
        mp = (C.c_int32 * 3)(*map_pos)
        cp = (C.c_int32 * 3)(*chunk_pos)
        self.L.orh_set_voxel(self.v, mp, cp, int(normal_word), int(albedo_word))

    def set_chunk(self, map_pos, voxels):
        """voxels: uint32 [8,8,8,2] indexed [x][y][z] -> (normal word, albedo word)"""
        mp = (C.c_int32 * 3)(*map_pos)
        a = np.ascontiguousarray(voxels, dtype=np.uint32)
        assert a.shape == (8, 8, 8, 2)
        self.L.orh_set_chunk(self.v, mp, a.ctypes.data)

    def compress_voxel(self, material, normal, albedo):
        out = (C.c_uint32 * 2)()
        self.L.orh_compress_voxel(material, (C.c_float * 3)(*normal), (C.c_uint8 * 3)(*albedo), out)
        return int(out[0]), int(out[1])

    # --- frame ---
    def sync(self, op=2, split=1):
        self.L.orh_sync(self.v, op, split)

    def view_projection(self, aspect, near=0.1, far=100.0):
        view = np.zeros(16, np.float32)
        proj = np.zeros(16, np.float32)
        self.L.orh_view_projection(self.v, aspect, near, far, view.ctypes.data, proj.ctypes.data)
        return view, proj

    def draw(self, w, h, aspect=None, want_hits=False):
        view, proj = self.view_projection(aspect if aspect is not None else h / w)
        img = np.zeros((h, w, 4), np.float32)
        hits = np.zeros((h, w), HIT_DT) if want_hits else None
        self.L.orh_draw(self.v, w, h, view.ctypes.data, proj.ctypes.data, img.ctypes.data,
                        hits.ctypes.data if want_hits else None)
        return (img, hits) if want_hits else img

    def update_lighting(self, num_diffuse=1, max_diffuse=1000, time=1.0):
        self.L.orh_update_lighting(self.v, num_diffuse, max_diffuse, C.c_float(time))

    def light_compute(self, num_diffuse, max_diffuse, time, first, count, staging, propagate):
        """phase 1 over requests [first, first+count): fills staging (uint32[96*R]) and ORs propagate (uint8[tiles])."""
        self.L.orh_light_compute(self.v, num_diffuse, max_diffuse, C.c_float(time), first, count, staging.ctypes.data, propagate.ctypes.data)

    def light_commit(self, staging, propagate):
        self.L.orh_light_commit(self.v, staging.ctypes.data, propagate.ctypes.data)

    def draw_uniforms(self, aspect):
        view, proj = self.view_projection(aspect)
        u = OrbUniforms()
        self.L.orh_draw_uniforms(self.v, view.ctypes.data, proj.ctypes.data, C.byref(u))
        return u

    # --- state ---
    def map_view(self):
        return _view(self.L.orh_map(self.v), HANDLE_DT, self.num_tiles())

    def chunk_view(self):
        return _view(self.L.orh_gpu_chunks(self.v), CHUNK_DT, self.num_tiles())

    def voxel_view(self):
        return _view(self.L.orh_voxels(self.v), VOXEL_DT, self.L.orh_voxel_top(self.v))

    def requests(self):
        return _view(self.L.orh_requests(self.v), np.dtype("<u4"), self.L.orh_num_requests(self.v)).copy()

    def counters(self):
        d = OrbCounters.from_address(self.L.orh_draw_counters(self.v)).as_dict()
        l = OrbCounters.from_address(self.L.orh_light_counters(self.v)).as_dict()
        return {"draw": d, "light": l, "upload_bytes": int(self.L.orh_upload_bytes(self.v))}

    def reset_counters(self):
        self.L.orh_reset_counters(self.v)

    def num_threads(self):
        return int(self.L.orb_num_threads())


LIB_GLSL = os.path.join(HERE, "_ref", "libglsl_ref.so")


def have_glsl():
    return os.path.exists(LIB_GLSL)


class OrbBuffers(C.Structure):
    _fields_ = [("map", C.c_void_p), ("chunks", C.c_void_p), ("voxels", C.c_void_p), ("materials", C.c_void_p)]


class GlslEngine(OracleEngine):
    """the restated host (host_cpu.c) with the reference's OWN shaders as its device: every dispatch runs the text of
    assets/shaders/voxel{Shared,Lighting,Draw}.comp compiled as C++ (oracle/glsl/, oracle/_ref/libglsl_ref.so) instead of the hand
    restatement shader_cpu.c.  Exists only where /root/reference does; used to PIN shader_cpu.c (tests/test_glsl_pin.py)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        if not have_glsl():
            build()
        if not have_glsl():
            raise RuntimeError("oracle/_ref/libglsl_ref.so is not built and /root/reference is absent")
        G = C.CDLL(LIB_GLSL)
        G.glsl_draw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        G.glsl_light.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
        self.G = G

    def _buffers(self):
        L = self.L
        return OrbBuffers(L.orh_map(self.v), L.orh_gpu_chunks(self.v), L.orh_voxels(self.v), L.orh_materials(self.v))

    def draw(self, w, h, aspect=None, want_hits=False):
        view, proj = self.view_projection(aspect if aspect is not None else h / w)
        u = OrbUniforms()
        self.L.orh_draw_uniforms(self.v, view.ctypes.data, proj.ctypes.data, C.byref(u))
        img = np.zeros((h, w, 4), np.float32)
        hits = np.zeros((h, w), HIT_DT) if want_hits else None
        buf = self._buffers()
        self.G.glsl_draw(C.byref(buf), C.byref(u), w, h, img.ctypes.data, hits.ctypes.data if want_hits else None)
        if want_hits:
            fill_local_indices(hits, self.map_view(), self.chunk_view())
        return (img, hits) if want_hits else img

    def update_lighting(self, num_diffuse=1, max_diffuse=1000, time=1.0):
        u = OrbUniforms()
        self.L.orh_light_uniforms(self.v, num_diffuse, max_diffuse, C.c_float(time), C.byref(u))
        buf = self._buffers()
        self.G.glsl_light(C.byref(buf), C.byref(u), self.L.orh_requests(self.v), self.L.orh_num_requests(self.v), self.L.orh_voxel_top(self.v))


_MSG_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_char_p)


class RefEngine(_EngineBase):
    """the reference's own host code (voxel.c) with Oracle B as its GPU."""

    def __init__(self, map_size=None, min_chunks=256, voxvol=None, resident=True, glsl=False):
        """glsl=True: the dispatches run the reference's own shaders (oracle/_ref/libglsl_ref.so) instead of the restatement --
        reference host + reference shaders, nothing restated."""
        if not have_ref():
            build()
        if not have_ref():
            raise RuntimeError("oracle/_ref/libdoon_ref.so is not built and /root/reference is absent")
        L = C.CDLL(LIB_REF)
        self.L = L
        L.fgl_init()
        L.fgl_install_message_sink()

        class DNuvec3(C.Structure):
            _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32)]

        class DNivec3(C.Structure):
            _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]

        class DNcompressedVoxel(C.Structure):
            _fields_ = [("normal", C.c_uint32), ("albedo", C.c_uint32)]

        self.DNivec3, self.DNcompressedVoxel = DNivec3, DNcompressedVoxel
        L.DN_init.restype = C.c_bool
        L.DN_create_volume.restype = C.POINTER(DNvolume)
        L.DN_create_volume.argtypes = [DNuvec3, C.c_uint]
        L.DN_load_volume.restype = C.POINTER(DNvolume)
        L.DN_load_volume.argtypes = [C.c_char_p, C.c_uint]
        L.DN_delete_volume.argtypes = [C.POINTER(DNvolume)]
        L.DN_sync_gpu.argtypes = [C.POINTER(DNvolume), C.c_int, C.c_int]
        L.DN_update_lighting.argtypes = [C.POINTER(DNvolume), C.c_int, C.c_int, C.c_float]
        L.DN_set_compressed_voxel.argtypes = [C.POINTER(DNvolume), DNivec3, DNivec3, DNcompressedVoxel]
        L.fgl_set_chunk.argtypes = [C.POINTER(DNvolume), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.fgl_draw.argtypes = [C.POINTER(DNvolume), C.c_uint, C.c_void_p, C.c_void_p]
        L.fgl_set_view_projection.argtypes = [C.POINTER(DNvolume), C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.fgl_create_texture.restype = C.c_uint
        L.fgl_create_texture.argtypes = [C.c_int, C.c_int]
        L.fgl_texture_pixels.restype = C.c_void_p
        L.fgl_texture_pixels.argtypes = [C.c_uint]
        L.fgl_texture_hits.restype = C.c_void_p
        L.fgl_texture_hits.argtypes = [C.c_uint]
        L.fgl_delete_texture.argtypes = [C.c_uint]
        L.fgl_buffer_ptr.restype = C.c_void_p
        L.fgl_buffer_ptr.argtypes = [C.c_uint]
        L.fgl_buffer_size.restype = C.c_size_t
        L.fgl_buffer_size.argtypes = [C.c_uint]
        L.fgl_request_all_unloaded.restype = C.c_size_t
        L.fgl_request_all_unloaded.argtypes = [C.POINTER(DNvolume)]
        L.fgl_mark_all_visible.restype = C.c_size_t
        L.fgl_mark_all_visible.argtypes = [C.POINTER(DNvolume)]
        L.fgl_collect_uniforms.argtypes = [C.c_int, C.c_void_p]
        L.fgl_get_counters.argtypes = [C.c_int, C.c_void_p]
        L.fgl_last_message.restype = C.c_char_p
        L.fgl_upload_bytes.restype = C.c_size_t
        L.fgl_binding.restype = C.c_uint
        L.fgl_binding.argtypes = [C.c_uint]
        L.fgl_set_device.argtypes = [C.c_void_p, C.c_void_p]
        if glsl:
            if not have_glsl():
                raise RuntimeError("oracle/_ref/libglsl_ref.so is not built")
            self.G = C.CDLL(LIB_GLSL)
            L.fgl_set_device(C.cast(self.G.glsl_draw, C.c_void_p), C.cast(self.G.glsl_light, C.c_void_p))
        else:
            L.fgl_set_device(None, None)
        if not L.DN_init():
            raise RuntimeError("reference DN_init failed under the shim")
        if voxvol is not None:
            self.vol = L.DN_load_volume(os.fsencode(voxvol), min_chunks)
        else:
            self.vol = L.DN_create_volume(DNuvec3(*map_size), min_chunks)
        if not self.vol:
            raise RuntimeError("reference volume creation failed")
        self.resident = resident
        self.map_size = tuple(int(x) for x in self.vol.contents.mapSize)
        self._tex = {}

    def close(self):
        if self.vol:
            for t in self._tex.values():
                self.L.fgl_delete_texture(t)
            self.L.DN_delete_volume(self.vol)
            self.vol = None

    def set_params(self, **kw):
        set_params(self.vol.contents, **kw)

    def get_params(self):
        return get_params(self.vol.contents)

    def materials(self):
        return _view(self.vol.contents.materials, MATERIAL_DT, 256)

    def set_voxel(self, map_pos, chunk_pos, normal_word, albedo_word):
        self.L.DN_set_compressed_voxel(self.vol, self.DNivec3(*map_pos), self.DNivec3(*chunk_pos),
                                       self.DNcompressedVoxel(int(normal_word), int(albedo_word)))

    def set_chunk(self, map_pos, voxels):
        a = np.ascontiguousarray(voxels, dtype=np.uint32)
        assert a.shape == (8, 8, 8, 2)
        self.L.fgl_set_chunk(self.vol, int(map_pos[0]), int(map_pos[1]), int(map_pos[2]), a.ctypes.data)

    def sync(self, op=2, split=1):
        """DN_sync_gpu; in resident mode a WRITE sync is followed by the pre-warm of
        SURVEY.md Appendix A: tiles that reached state 1 are forced to 3 (what a ray
        would do, voxelShared.comp:462-466) and uploaded by a second WRITE-only pass
        that does not touch frameNum-dependent state (lightingSplit passed through)."""
        self.L.DN_sync_gpu(self.vol, op, split)
        if self.resident and op != 0:
            if self.L.fgl_request_all_unloaded(self.vol):
                saved = self.vol.contents.frameNum
                self.vol.contents.frameNum = (saved - 1) % (1 << 32)
                n_req = self.vol.contents.numLightingRequests
                self.L.DN_sync_gpu(self.vol, 1, split)
                self.vol.contents.frameNum = saved
                self.vol.contents.numLightingRequests = n_req

    def view_projection(self, aspect, near=0.1, far=100.0):
        view = np.zeros(16, np.float32)
        proj = np.zeros(16, np.float32)
        self.L.fgl_set_view_projection(self.vol, aspect, near, far, view.ctypes.data, proj.ctypes.data)
        return view, proj

    def _texture(self, w, h):
        if (w, h) not in self._tex:
            self._tex[(w, h)] = self.L.fgl_create_texture(w, h)
        return self._tex[(w, h)]

    def draw(self, w, h, aspect=None, want_hits=False):
        view, proj = self.view_projection(aspect if aspect is not None else h / w)
        tex = self._texture(w, h)
        px = _view(self.L.fgl_texture_pixels(tex), np.dtype("<f4"), w * h * 4)
        px[:] = 0
        self.L.fgl_draw(self.vol, tex, view.ctypes.data, proj.ctypes.data)
        img = px.reshape(h, w, 4).copy()
        if want_hits:
            hits = _view(self.L.fgl_texture_hits(tex), HIT_DT, w * h).reshape(h, w).copy()
            if getattr(self, "G", None) is not None:
                fill_local_indices(hits, self.map_view(), self.chunk_view())
            return img, hits
        return img

    def update_lighting(self, num_diffuse=1, max_diffuse=1000, time=1.0):
        self.L.DN_update_lighting(self.vol, num_diffuse, max_diffuse, C.c_float(time))

    def step_map(self, direction, origin, max_steps):
        """the reference's OWN DN_step_map (voxel.c:1195-1272) on its CPU map: (hit, cell, face normal, (material, normal xyz, albedo rgb))"""
        if not hasattr(self, "_step_types"):
            class V3(C.Structure):
                _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

            class Col(C.Structure):
                _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)]

            class Vox(C.Structure):
                _fields_ = [("material", C.c_uint8), ("normal", V3), ("albedo", Col)]

            self._step_types = (V3, Vox)
            self.L.DN_step_map.restype = C.c_bool
            self.L.DN_step_map.argtypes = [C.POINTER(DNvolume), V3, V3, C.c_uint, C.POINTER(self.DNivec3), C.POINTER(Vox), C.POINTER(self.DNivec3)]
        V3, Vox = self._step_types
        hp, hn, hv = self.DNivec3(), self.DNivec3(), Vox()
        ok = bool(self.L.DN_step_map(self.vol, V3(*[float(x) for x in direction]), V3(*[float(x) for x in origin]), int(max_steps), C.byref(hp), C.byref(hv), C.byref(hn)))
        return ok, (hp.x, hp.y, hp.z), (hn.x, hn.y, hn.z), (hv.material, (hv.normal.x, hv.normal.y, hv.normal.z), (hv.albedo.r, hv.albedo.g, hv.albedo.b))

    def uniforms(self, program):
        u = OrbUniforms()
        self.L.fgl_collect_uniforms(program, C.byref(u))
        return u

    def map_view(self):
        return _view(self.L.fgl_buffer_ptr(self.vol.contents.glMapBufferID), HANDLE_DT, self.num_tiles())

    def chunk_view(self):
        return _view(self.L.fgl_buffer_ptr(self.vol.contents.glChunkBufferID), CHUNK_DT, self.num_tiles())

    def voxel_view(self):
        return _view(self.L.fgl_buffer_ptr(self.vol.contents.glVoxelBufferID), VOXEL_DT,
                     self.vol.contents.voxelCap + 512)

    def requests(self):
        return _view(self.vol.contents.lightingRequests, np.dtype("<u4"), self.vol.contents.numLightingRequests).copy()

    def host_chunks(self):
        return _view(self.vol.contents.chunks, HOST_CHUNK_DT, self.vol.contents.chunkCap)

    def counters(self):
        d, l = OrbCounters(), OrbCounters()
        self.L.fgl_get_counters(2, C.byref(d))
        self.L.fgl_get_counters(1, C.byref(l))
        return {"draw": d.as_dict(), "light": l.as_dict(), "upload_bytes": int(self.L.fgl_upload_bytes())}

    def reset_counters(self):
        self.L.fgl_reset_counters()

    def last_message(self):
        return self.L.fgl_last_message().decode()
