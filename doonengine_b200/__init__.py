"""doonengine_b200 -- B200 (sm_100a) CUDA back end for DoonEngine's per-voxel lighting + ray-cast draw path.

The product is the C-ABI shared library ``libdoon_b200.so`` (the DN_* API of include/DoonEngine/voxel.h plus the
additive DN_b200_* calls of include/DoonEngine/b200.h).  This package is the thin Python host above it:

    lib()          ctypes handle with every prototype declared; raises if the library is not built
    Engine         one volume driven through the C ABI (create/load, edit, sync, draw, update_lighting, state export)

Nothing here computes lighting or pixels: without a CUDA device ``Engine`` raises.  The oracle (oracle/) is test
infrastructure and is never imported from this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DN_B200_LIB") or os.path.join(HERE, "libdoon_b200.so")  # $DN_B200_LIB: an experimental build variant


# ---- C structs (include/DoonEngine/*.h) ----
class DNuvec3(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("z", C.c_uint32)]


class DNivec3(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32)]


class DNvec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class DNcolor(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)]


class DNvoxel(C.Structure):
    _fields_ = [("material", C.c_uint8), ("normal", DNvec3), ("albedo", DNcolor)]


class DNcompressedVoxel(C.Structure):
    _fields_ = [("normal", C.c_uint32), ("albedo", C.c_uint32)]


class DNmat4(C.Structure):
    """64-byte column-major matrix.  The C type is 16-byte aligned and DN_draw takes two of them BY VALUE; both
    land on the stack at 16-byte-aligned offsets (they are the only stack arguments), so ctypes' 4-byte
    alignment of this Structure is harmless."""
    _fields_ = [("m", (C.c_float * 4) * 4)]


class DNvolume(C.Structure):
    """public volume struct, include/DoonEngine/voxel.h (232 bytes, same as reference voxel.h:97-142)."""
    _fields_ = [("glMapBufferID", C.c_uint32), ("glChunkBufferID", C.c_uint32), ("glVoxelBufferID", C.c_uint32),
                ("mapSize", DNuvec3),
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


class DNb200counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "tiles", "chunks", "voxelSteps", "records", "voxelsLit", "pixels")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class DNb200stats(C.Structure):
    _fields_ = [("chunksUploaded", C.c_uint64), ("chunksRemoved", C.c_uint64), ("bytesUploaded", C.c_uint64),
                ("residentChunks", C.c_uint64), ("residentRecords", C.c_uint64),
                ("slotCap", C.c_uint64), ("recordCap", C.c_uint64), ("voxelsLit", C.c_uint64),
                ("lastDrawMs", C.c_float), ("lastCompactMs", C.c_float), ("lastUploadMs", C.c_float),
                ("lastLightMs", C.c_float), ("lastCommitMs", C.c_float),
                ("lightLaunchesWarp", C.c_uint64), ("lightLaunchesFlat", C.c_uint64), ("nsPerCtaWarp", C.c_float), ("nsPerCtaFlat", C.c_float),
                ("lastScanHostMs", C.c_float), ("lastPackHostMs", C.c_float), ("lastEnqueueHostMs", C.c_float),
                ("lightLaunchesWave", C.c_uint64), ("nsPerCtaWave", C.c_float), ("lastWavePasses", C.c_uint32), ("pad0", C.c_uint32),
                ("nodeSplits", C.c_uint64), ("nodeMerges", C.c_uint64), ("usedNodes", C.c_uint64), ("freeNodes", C.c_uint64), ("recordTop", C.c_uint64),
                ("lightLaunchesSpread", C.c_uint64), ("nsPerCtaSpread", C.c_float), ("pad1", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DNb200peerBuffers(C.Structure):
    """where one replica's exchange buffers are mapped in this process (include/DoonEngine/b200.h)."""
    _fields_ = [("staging", C.c_void_p), ("mailbox", C.c_void_p), ("visible", C.c_void_p), ("propagate", C.c_void_p),
                ("stagingRequestCap", C.c_uint64)]


PEER_AUTO, PEER_MANUAL = 0, 1

assert C.sizeof(DNvolume) == 232 and C.sizeof(DNvoxel) == 20 and C.sizeof(DNmat4) == 64

MESSAGE_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_char_p)
MESSAGE_TYPES = ("CPU_MEMORY", "GPU_MEMORY", "SHADER", "FILE_IO")
MESSAGE_SEVERITIES = ("NOTE", "ERROR", "FATAL")

# numpy views of the device / host layouts
HOST_CHUNK_DT = np.dtype([("pos", "<i4", 3), ("updated", "u1"), ("_pad", "u1", 3), ("numVoxels", "<u4"),
                          ("numVoxelsGpu", "<u4"), ("voxels", "<u4", (8, 8, 8, 2))])
HOST_HANDLE_DT = np.dtype([("flag", "u1"), ("_pad", "u1", 3), ("chunkIndex", "<u4")])
MATERIAL_DT = np.dtype([("pad", "<f4", 2), ("emissive", "<u4"), ("opacity", "<f4"), ("refractIndex", "<f4"),
                        ("specular", "<f4"), ("reflectType", "<u4"), ("shininess", "<u4")])
SLOT_DT = np.dtype([("mask", "<u4", 16), ("voxelBase", "<u4"), ("numVoxels", "<u4"), ("numSamples", "<u4"),
                    ("matIds", "<u4"), ("prefix", "<u2", 16), ("pos", "<i4", 3), ("bbox", "<u4")])
VOXEL_DT = np.dtype([("material", "u1"), ("normal", "<f4", 3), ("albedo", "u1", 3)], align=True)  # DNvoxel (voxel.h:44-52), 20 bytes
HIT_DT = np.dtype([("status", "<i4"), ("mapIndex", "<u4"), ("localIndex", "<u4"), ("recordIndex", "<u4")])
assert HOST_CHUNK_DT.itemsize == 4120 and HOST_HANDLE_DT.itemsize == 8 and MATERIAL_DT.itemsize == 32 and SLOT_DT.itemsize == 128

ARRAY_TILE_SLOTS, ARRAY_VISIBLE, ARRAY_SLOTS, ARRAY_RECORDS, ARRAY_REQUESTS, ARRAY_STAGING, ARRAY_PROPAGATE = range(7)
DN_READ, DN_WRITE, DN_READ_WRITE = 0, 1, 2

PARAM_NAMES = ("camPos", "camOrient", "camFOV", "camViewMode", "sunDir", "sunStrength", "ambientLightStrength",
               "diffuseBounceLimit", "specBounceLimit", "shadowSoftness", "skyGradientBot", "skyGradientTop")

_PROTOTYPES = {
    # reference API (include/DoonEngine/voxel.h)
    "DN_init": (C.c_bool, []),
    "DN_quit": (None, []),
    "DN_create_volume": (C.POINTER(DNvolume), [DNuvec3, C.c_uint]),
    "DN_delete_volume": (None, [C.POINTER(DNvolume)]),
    "DN_load_volume": (C.POINTER(DNvolume), [C.c_char_p, C.c_uint]),
    "DN_save_volume": (C.c_bool, [C.c_char_p, C.POINTER(DNvolume)]),
    "DN_set_view_projection_matrices": (None, [C.POINTER(DNvolume), C.c_float, C.c_float, C.c_float, C.POINTER(DNmat4), C.POINTER(DNmat4)]),
    "DN_draw": (None, [C.POINTER(DNvolume), C.c_uint32, DNmat4, DNmat4, C.c_int, C.c_int]),
    "DN_update_lighting": (None, [C.POINTER(DNvolume), C.c_int, C.c_int, C.c_float]),
    "DN_sync_gpu": (None, [C.POINTER(DNvolume), C.c_int, C.c_int]),
    "DN_add_chunk": (C.c_int, [C.POINTER(DNvolume), DNivec3]),
    "DN_remove_chunk": (None, [C.POINTER(DNvolume), DNivec3]),
    "DN_set_map_size": (C.c_bool, [C.POINTER(DNvolume), DNuvec3]),
    "DN_set_max_chunks": (C.c_bool, [C.POINTER(DNvolume), C.c_size_t]),
    "DN_set_max_voxels_gpu": (C.c_bool, [C.POINTER(DNvolume), C.c_size_t]),
    "DN_set_max_lighting_requests": (C.c_bool, [C.POINTER(DNvolume), C.c_size_t]),
    "DN_in_map_bounds": (C.c_bool, [C.POINTER(DNvolume), DNivec3]),
    "DN_in_chunk_bounds": (C.c_bool, [DNivec3]),
    "DN_get_voxel": (DNvoxel, [C.POINTER(DNvolume), DNivec3, DNivec3]),
    "DN_get_compressed_voxel": (DNcompressedVoxel, [C.POINTER(DNvolume), DNivec3, DNivec3]),
    "DN_set_voxel": (None, [C.POINTER(DNvolume), DNivec3, DNivec3, DNvoxel]),
    "DN_set_compressed_voxel": (None, [C.POINTER(DNvolume), DNivec3, DNivec3, DNcompressedVoxel]),
    "DN_remove_voxel": (None, [C.POINTER(DNvolume), DNivec3, DNivec3]),
    "DN_does_chunk_exist": (C.c_bool, [C.POINTER(DNvolume), DNivec3]),
    "DN_does_voxel_exist": (C.c_bool, [C.POINTER(DNvolume), DNivec3, DNivec3]),
    "DN_step_map": (C.c_bool, [C.POINTER(DNvolume), DNvec3, DNvec3, C.c_int, C.POINTER(DNivec3), C.POINTER(DNvoxel), C.POINTER(DNivec3)]),
    "DN_separate_position": (None, [DNivec3, C.POINTER(DNivec3), C.POINTER(DNivec3)]),
    "DN_cam_dir": (DNvec3, [DNvec3]),
    "DN_compress_voxel": (DNcompressedVoxel, [DNvoxel]),
    "DN_decompress_voxel": (DNvoxel, [DNcompressedVoxel]),
    # additive API (include/DoonEngine/b200.h)
    "DN_b200_device_count": (C.c_int, []),
    "DN_b200_set_device": (C.c_bool, [C.c_int]),
    "DN_b200_set_stream": (None, [C.c_void_p]),
    "DN_b200_synchronize": (C.c_bool, []),
    "DN_b200_create_framebuffer": (C.c_uint32, [C.c_int, C.c_int]),
    "DN_b200_delete_framebuffer": (None, [C.c_uint32]),
    "DN_b200_framebuffer_size": (C.c_bool, [C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "DN_b200_framebuffer_device_ptr": (C.c_void_p, [C.c_uint32]),
    "DN_b200_read_framebuffer": (C.c_bool, [C.c_uint32, C.c_void_p, C.c_size_t]),
    "DN_b200_clear_framebuffer": (C.c_bool, [C.c_uint32, C.c_float]),
    "DN_b200_read_framebuffer_async": (C.c_bool, [C.c_uint32, C.c_void_p, C.c_size_t]),
    "DN_b200_read_framebuffer_rows_async": (C.c_bool, [C.c_uint32, C.POINTER(DNvolume), C.c_void_p, C.c_size_t]),
    "DN_b200_host_register": (C.c_bool, [C.c_void_p, C.c_size_t]),
    "DN_b200_host_unregister": (C.c_bool, [C.c_void_p]),
    "DN_b200_wait_framebuffer": (C.c_bool, []),
    "DN_b200_wait_framebuffer_read": (C.c_bool, [C.c_uint32]),
    "DN_b200_capture_hits": (C.c_bool, [C.c_uint32, C.c_bool]),
    "DN_b200_read_hits": (C.c_bool, [C.c_uint32, C.c_void_p, C.c_size_t]),
    "DN_b200_set_light_kernel": (None, [C.c_int]),
    "DN_b200_get_light_kernel": (C.c_int, []),
    "DN_b200_set_flat_tuning": (None, [C.c_int, C.c_int, C.c_int]),
    "DN_b200_fetch_lighting_requests": (C.c_size_t, [C.POINTER(DNvolume)]),
    "DN_b200_array_bytes": (C.c_size_t, [C.POINTER(DNvolume), C.c_int]),
    "DN_b200_download": (C.c_size_t, [C.POINTER(DNvolume), C.c_int, C.c_void_p, C.c_size_t]),
    "DN_b200_array_device_ptr": (C.c_void_p, [C.POINTER(DNvolume), C.c_int]),
    "DN_b200_enable_counters": (C.c_bool, [C.POINTER(DNvolume), C.c_bool]),
    "DN_b200_read_counters": (C.c_bool, [C.POINTER(DNvolume), C.POINTER(DNb200counters), C.c_bool]),
    "DN_b200_get_stats": (None, [C.POINTER(DNvolume), C.POINTER(DNb200stats)]),
    "DN_b200_enable_timing": (None, [C.c_bool]),
    "DN_b200_kernel_launches": (C.c_uint64, []),
    "DN_b200_touch_tile": (None, [C.POINTER(DNvolume), DNivec3]),
    "DN_b200_rescan": (None, [C.POINTER(DNvolume)]),
    "DN_b200_pack_chunk": (C.c_int, [C.POINTER(DNvolume), DNivec3, C.c_void_p, C.c_void_p]),
    "DN_b200_set_voxels": (C.c_size_t, [C.POINTER(DNvolume), C.c_size_t, C.c_void_p, C.c_void_p]),
    "DN_b200_lighting_request_count": (C.c_size_t, [C.POINTER(DNvolume)]),
    "DN_b200_set_exact_sync": (None, [C.POINTER(DNvolume), C.c_bool]),
    "DN_b200_set_max_frames_in_flight": (None, [C.POINTER(DNvolume), C.c_int]),
    "DN_b200_mirror_voxel_layout": (C.c_bool, [C.POINTER(DNvolume)]),
    "DN_b200_step_map_batch": (C.c_size_t, [C.POINTER(DNvolume), C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "DN_b200_set_wave_slots": (None, [C.c_uint32]),
    "DN_b200_set_chunks": (C.c_size_t, [C.POINTER(DNvolume), C.c_size_t, C.c_void_p, C.c_void_p]),
    "DN_b200_save_lighting": (C.c_bool, [C.POINTER(DNvolume), C.c_char_p]),
    "DN_b200_load_lighting": (C.c_int, [C.POINTER(DNvolume), C.c_char_p]),
    "DN_b200_set_shard": (C.c_bool, [C.POINTER(DNvolume), C.c_int, C.c_int]),
    "DN_b200_light_compute": (C.c_bool, [C.POINTER(DNvolume), C.c_int, C.c_int, C.c_float]),
    "DN_b200_light_commit": (C.c_bool, [C.POINTER(DNvolume)]),
    "DN_b200_staging_slice_bytes": (C.c_size_t, [C.POINTER(DNvolume)]),
    "DN_b200_or_bitmap": (C.c_bool, [C.POINTER(DNvolume), C.c_int, C.c_void_p]),
    "DN_b200_peer_prepare": (C.c_bool, [C.POINTER(DNvolume), C.c_size_t, C.POINTER(DNb200peerBuffers)]),
    "DN_b200_ipc_export": (C.c_bool, [C.c_void_p, C.c_void_p]),
    "DN_b200_ipc_open": (C.c_void_p, [C.c_void_p]),
    "DN_b200_ipc_close": (C.c_bool, [C.c_void_p]),
    "DN_b200_peer_attach": (C.c_bool, [C.POINTER(DNvolume), C.c_int, C.c_int, C.POINTER(DNb200peerBuffers), C.c_int]),
    "DN_b200_peer_detach": (None, [C.POINTER(DNvolume)]),
    "DN_b200_peer_capacity_ok": (C.c_bool, [C.POINTER(DNvolume)]),
    "DN_b200_framebuffer_set_mirror": (C.c_bool, [C.c_uint32, C.c_void_p]),
    "DN_b200_peer_exchange_visible": (C.c_bool, [C.POINTER(DNvolume)]),
    "DN_b200_peer_barrier_status": (C.c_bool, [C.POINTER(DNvolume), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
}

_lib = None
_messages = []
_callback_keepalive = None


def _on_message(mtype, severity, text):
    _messages.append((MESSAGE_TYPES[mtype] if 0 <= mtype < 4 else str(mtype),
                      MESSAGE_SEVERITIES[severity] if 0 <= severity < 3 else str(severity),
                      text.decode(errors="replace") if text else ""))
    if len(_messages) > 4096:
        del _messages[:2048]


def lib():
    """the loaded C-ABI library, with prototypes; raises if it has not been built (no fallback of any kind)."""
    global _lib, _callback_keepalive
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -m doonengine_b200.build` (needs nvcc); there is no fallback path" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _callback_keepalive = MESSAGE_CB(_on_message)
    C.c_void_p.in_dll(L, "g_DN_message_callback").value = C.cast(_callback_keepalive, C.c_void_p).value
    _lib = L
    return L


def messages(clear=False):
    """(type, severity, text) tuples the library has reported through g_DN_message_callback."""
    out = list(_messages)
    if clear:
        del _messages[:]
    return out


_initialised = False


def init(device=None):
    """DN_init(); raises when there is no usable CUDA device."""
    global _initialised
    L = lib()
    if _initialised:
        return
    if device is not None:
        L.DN_b200_set_device(int(device))
    if not L.DN_init():
        raise RuntimeError("DN_init failed: %s" % (messages()[-1][2] if messages() else "no CUDA device"))
    _initialised = True


def _view(ptr, dtype, count):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (dtype.itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class Engine:
    """One DNvolume behind the C ABI, with the frame protocol of SURVEY.md 8d.

    The method names and argument meanings mirror the reference calls they wrap (DN_set_compressed_voxel,
    DN_sync_gpu, DN_draw, DN_update_lighting ...); the oracle's engines expose the same interface so that the
    parity tests read the same for all three.
    """

    def __init__(self, map_size=None, min_chunks=256, voxvol=None, host_only=False, device=None):
        self.L = lib()
        if not host_only:
            init(device)
        self.host_only = host_only
        if voxvol is not None:
            self.vol = self.L.DN_load_volume(os.fsencode(voxvol), min_chunks)
        else:
            self.vol = self.L.DN_create_volume(DNuvec3(*map_size), min_chunks)
        if not self.vol:
            raise RuntimeError("volume creation failed: %s" % (messages()[-1][2] if messages() else "?"))
        ms = self.vol.contents.mapSize
        self.map_size = (int(ms.x), int(ms.y), int(ms.z))
        self._fb = {}

    # ---- lifetime ----
    def close(self):
        if self.vol:
            for fb in self._fb.values():
                self.L.DN_b200_delete_framebuffer(fb)
            self._fb = {}
            self.L.DN_delete_volume(self.vol)
            self.vol = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def num_tiles(self):
        sx, sy, sz = self.map_size
        return sx * sy * sz

    # ---- parameters (public DNvolume fields) ----
    def set_params(self, **kw):
        v = self.vol.contents
        for k, val in kw.items():
            cur = getattr(v, k)
            if hasattr(cur, "__len__"):
                for i, x in enumerate(val):
                    cur[i] = x
            else:
                setattr(v, k, val)

    def get_params(self):
        v = self.vol.contents
        out = {}
        for k in PARAM_NAMES:
            val = getattr(v, k)
            out[k] = tuple(val) if hasattr(val, "__len__") else val
        return out

    def materials(self):
        return _view(self.vol.contents.materials, MATERIAL_DT, 256)

    def host_chunks(self):
        return _view(self.vol.contents.chunks, HOST_CHUNK_DT, self.vol.contents.chunkCap)

    def host_map(self):
        return _view(self.vol.contents.map, HOST_HANDLE_DT, self.num_tiles())

    # ---- edits ----
    def set_voxel(self, map_pos, chunk_pos, normal_word, albedo_word):
        self.L.DN_set_compressed_voxel(self.vol, DNivec3(*map_pos), DNivec3(*chunk_pos), DNcompressedVoxel(int(normal_word), int(albedo_word)))

    def remove_voxel(self, map_pos, chunk_pos):
        self.L.DN_remove_voxel(self.vol, DNivec3(*map_pos), DNivec3(*chunk_pos))

    def set_chunk(self, map_pos, voxels):
        """bulk form of 512 DN_set_compressed_voxel calls; voxels: uint32 [8,8,8,2] indexed [x][y][z] -> (normal, albedo).

        The chunk is written straight into vol->chunks (the struct is public, reference voxel.h:114-116) with the
        bookkeeping DN_set_compressed_voxel would have done, then announced with DN_b200_touch_tile."""
        a = np.ascontiguousarray(voxels, dtype=np.uint32)
        assert a.shape == (8, 8, 8, 2)
        solid = (a[..., 0] >> 24) != 255
        n = int(solid.sum())
        mp = DNivec3(*map_pos)
        if n == 0:
            if self.L.DN_does_chunk_exist(self.vol, mp):
                self.L.DN_remove_chunk(self.vol, mp)
            return
        if not self.L.DN_does_chunk_exist(self.vol, mp):
            if self.L.DN_add_chunk(self.vol, mp) < 0:
                raise MemoryError("DN_add_chunk failed")
        sx, sy, _ = self.map_size
        idx = map_pos[0] + sx * (map_pos[1] + sy * map_pos[2])
        ci = int(self.host_map()["chunkIndex"][idx])
        ch = self.host_chunks()
        ch["voxels"][ci] = a
        ch["voxels"][ci][..., 0][~solid] = 0xFFFFFFFF
        ch["numVoxels"][ci] = n
        ch["updated"][ci] = 1
        self.L.DN_b200_touch_tile(self.vol, mp)

    def set_chunks(self, positions, voxels):
        """bulk form of set_chunk: positions int32 [n,3] (tiles), voxels uint32 [n,8,8,8,2]; returns the chunks present afterwards."""
        p = np.ascontiguousarray(positions, dtype=np.int32)
        v = np.ascontiguousarray(voxels, dtype=np.uint32)
        assert p.ndim == 2 and p.shape[1] == 3 and v.shape == (p.shape[0], 8, 8, 8, 2)
        return int(self.L.DN_b200_set_chunks(self.vol, p.shape[0], p.ctypes.data, v.ctypes.data))

    def set_voxels(self, positions, voxels):
        """bulk edits: positions int32 [n,3] in voxel units, voxels uint32 [n,2] (normal word, albedo word); material 255 removes."""
        p = np.ascontiguousarray(positions, dtype=np.int32)
        v = np.ascontiguousarray(voxels, dtype=np.uint32)
        assert p.ndim == 2 and p.shape[1] == 3 and v.shape == (p.shape[0], 2)
        return int(self.L.DN_b200_set_voxels(self.vol, p.shape[0], p.ctypes.data, v.ctypes.data))

    def step_map_batch(self, dirs, origins, max_steps):
        """DN_step_map for many rays at once on the device map (DN_b200_step_map_batch).  dirs / origins: float32 [n,3] (origins in
        chunk units).  Returns dict(hit uint8 [n], pos int32 [n,3], normal int32 [n,3], voxel structured [n] (DNvoxel fields))."""
        d = np.ascontiguousarray(dirs, dtype=np.float32)
        o = np.ascontiguousarray(origins, dtype=np.float32)
        assert d.ndim == 2 and d.shape[1] == 3 and o.shape == d.shape
        n = d.shape[0]
        hit = np.zeros(n, np.uint8)
        pos = np.zeros((n, 3), np.int32)
        normal = np.zeros((n, 3), np.int32)
        voxel = np.zeros(n, VOXEL_DT)
        hits = self.L.DN_b200_step_map_batch(self.vol, n, d.ctypes.data, o.ctypes.data, int(max_steps), pos.ctypes.data, voxel.ctypes.data, normal.ctypes.data, hit.ctypes.data)
        assert hits == int(hit.sum())
        return dict(hit=hit, pos=pos, normal=normal, voxel=voxel)

    def step_map(self, direction, origin, max_steps):
        """one DN_step_map call (CPU map, voxel.c:1195-1272): (hit, pos, normal, voxel fields) like one row of step_map_batch."""
        hp, hn, hv = DNivec3(), DNivec3(), DNvoxel()
        ok = bool(self.L.DN_step_map(self.vol, DNvec3(*[float(x) for x in direction]), DNvec3(*[float(x) for x in origin]), int(max_steps), C.byref(hp), C.byref(hv), C.byref(hn)))
        return ok, (hp.x, hp.y, hp.z), (hn.x, hn.y, hn.z), hv

    def pack_chunk(self, map_pos):
        """(slot header as SLOT_DT scalar, records uint32 [n,4]) of the chunk at map_pos as the next writing sync would upload it;
        host-only.  None if the tile has no chunk."""
        slot = np.zeros(1, SLOT_DT)
        rec = np.zeros((512, 4), np.uint32)
        n = self.L.DN_b200_pack_chunk(self.vol, DNivec3(*map_pos), slot.ctypes.data, rec.ctypes.data)
        return None if n < 0 else (slot[0], rec[:n].copy())

    def compress_voxel(self, material, normal, albedo):
        r = self.L.DN_compress_voxel(DNvoxel(material, DNvec3(*normal), DNcolor(*albedo)))
        return int(r.normal), int(r.albedo)

    # ---- frame ----
    def sync(self, op=DN_READ_WRITE, split=1):
        self.L.DN_sync_gpu(self.vol, op, split)

    def view_projection(self, aspect, near=0.1, far=100.0):
        view, proj = DNmat4(), DNmat4()
        self.L.DN_set_view_projection_matrices(self.vol, aspect, near, far, C.byref(view), C.byref(proj))
        return view, proj

    def view_projection_arrays(self, aspect, near=0.1, far=100.0):
        view, proj = self.view_projection(aspect, near, far)
        return (np.frombuffer(view, dtype=np.float32, count=16).copy(), np.frombuffer(proj, dtype=np.float32, count=16).copy())

    def framebuffer(self, w, h):
        if (w, h) not in self._fb:
            fb = self.L.DN_b200_create_framebuffer(w, h)
            if not fb:
                raise MemoryError("DN_b200_create_framebuffer failed")
            self._fb[(w, h)] = fb
        return self._fb[(w, h)]

    def draw_async(self, w, h, aspect=None):
        """DN_draw into the cached framebuffer of that size; returns its handle without waiting."""
        view, proj = self.view_projection(aspect if aspect is not None else h / w)
        fb = self.framebuffer(w, h)
        self.L.DN_draw(self.vol, fb, view, proj, -1, -1)
        return fb

    def read_framebuffer(self, fb, out=None):
        w, h = C.c_int(), C.c_int()
        self.L.DN_b200_framebuffer_size(fb, C.byref(w), C.byref(h))
        if out is None:
            out = np.empty((h.value, w.value, 4), np.float32)
        if not self.L.DN_b200_read_framebuffer(fb, out.ctypes.data, out.nbytes):
            raise RuntimeError("DN_b200_read_framebuffer failed")
        return out

    def draw(self, w, h, aspect=None, want_hits=False):
        fb = self.framebuffer(w, h)
        self.L.DN_b200_clear_framebuffer(fb, 0.0)
        self.L.DN_b200_capture_hits(fb, bool(want_hits))
        self.draw_async(w, h, aspect)
        img = self.read_framebuffer(fb)
        if want_hits:
            hits = np.zeros((h, w), HIT_DT)
            if not self.L.DN_b200_read_hits(fb, hits.ctypes.data, hits.size):
                raise RuntimeError("DN_b200_read_hits failed")
            return img, hits
        return img

    def update_lighting(self, num_diffuse=1, max_diffuse=1000, time=1.0):
        self.L.DN_update_lighting(self.vol, num_diffuse, max_diffuse, C.c_float(time))

    def frame(self, w, h, time, num_diffuse=1, max_diffuse=1000, split=1, aspect=None, want_hits=False):
        """draw -> sync(READ_WRITE) -> update_lighting (reference main.c:503-505)."""
        res = self.draw(w, h, aspect=aspect, want_hits=want_hits)
        self.sync(DN_READ_WRITE, split)
        self.update_lighting(num_diffuse, max_diffuse, time)
        return res

    def synchronize(self):
        if not self.L.DN_b200_synchronize():
            raise RuntimeError("CUDA error: %s" % (messages()[-1][2] if messages() else "?"))

    # ---- state ----
    def download(self, which, dtype):
        n = self.L.DN_b200_array_bytes(self.vol, which)
        out = np.zeros(n // np.dtype(dtype).itemsize, dtype=dtype)
        if n and self.L.DN_b200_download(self.vol, which, out.ctypes.data, out.nbytes) != n:
            raise RuntimeError("DN_b200_download failed: %s" % (messages()[-1][2] if messages() else "?"))
        return out

    def requests(self):
        n = self.L.DN_b200_fetch_lighting_requests(self.vol)
        return _view(self.vol.contents.lightingRequests, np.dtype("<u4"), n).copy()

    def num_requests(self):
        """exact length of the request list of the last reading sync (waits for the device if it is still in flight)"""
        return int(self.L.DN_b200_lighting_request_count(self.vol))

    def export_state(self):
        """dict mapIndex -> (state, visible, header(pos, samples, partialCounts, bitMask), records[n,4]) in the same
        layout-independent form as the oracle engines' export_state()."""
        tile_slot = self.download(ARRAY_TILE_SLOTS, np.uint32)
        vis = self.download(ARRAY_VISIBLE, np.uint32)
        slots = self.download(ARRAY_SLOTS, SLOT_DT)
        rec = self.download(ARRAY_RECORDS, np.uint32).reshape(-1, 4)
        out = {}
        for idx in np.nonzero(tile_slot)[0]:
            idx = int(idx)
            s = slots[int(tile_slot[idx]) - 1]
            n = int(s["numVoxels"])
            base = int(s["voxelBase"])
            visible = int((vis[idx >> 5] >> (idx & 31)) & 1)
            pref = s["prefix"]
            hdr = (tuple(int(x) for x in s["pos"]), int(s["numSamples"]), (int(pref[4]), int(pref[8]), int(pref[12])),
                   tuple(int(x) for x in s["mask"]))
            out[idx] = (2, visible, hdr, rec[base:base + n].copy())
        return out

    def counters(self, reset=True):
        c = DNb200counters()
        self.L.DN_b200_read_counters(self.vol, C.byref(c), reset)
        return c.as_dict()

    def enable_counters(self, on=True):
        return bool(self.L.DN_b200_enable_counters(self.vol, on))

    def stats(self):
        s = DNb200stats()
        self.L.DN_b200_get_stats(self.vol, C.byref(s))
        return s.as_dict()
