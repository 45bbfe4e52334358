"""One process per GPU: the map is replicated, the lighting-request list and the screen rows are sharded.

Two exchange mechanisms:

  exchange="peer" (default on GPUs)   the replicas map each other's exchange buffers (cudaIpc handles, swapped once through
      torch.distributed) and from then on the plain frame calls run SPMD with NO collective on the frame path: the
      lighting kernel stores its staged words into every replica's staging array over NVLink, the draw kernel mirrors
      its pixels into the root's framebuffer, a device-side barrier kernel separates the phases and a merge kernel ORs
      the peers' visible / propagate bitmaps (csrc/peer.cu, csrc/light.cu, csrc/draw.cu).  Rows and request CTAs are
      interleaved over the ranks (rank, rank + world, ...).
  exchange="collective"               host-driven all-gathers of contiguous slices through torch.distributed (NCCL on GPUs, gloo
      in the CPU tests, where the oracle stands in for the device):

Per frame, on every rank (collective mode):

    draw            each rank ray-casts its band of 16-pixel rows          (DN_draw with DN_b200_set_shard)
      exchange      all-gather the framebuffer bands; OR the ranks' visible bitmaps together
    sync            every rank compacts the (now identical) visible bitmap itself -> identical request lists
    light compute   rank r lights requests [r*ceil(R/N), (r+1)*ceil(R/N))  (DN_b200_light_compute)
      exchange      all-gather the staged lit words (96 uint32 per request); OR the propagate bitmaps
    light commit    every rank scatters ALL staged words into its replica   (DN_b200_light_commit)

Because the lighting kernel only reads pre-dispatch state (snapshot semantics, oracle.h N1-N3) the result is
bit-identical for any number of ranks -- tests/test_shard_gloo.py checks that on CPU with the oracle as the "GPU",
tests/test_parity_gpu.py::test_sharded_equals_unsharded on the device.

The collectives go through torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests); the tensors
alias the library's own device buffers (no staging copies besides the padded framebuffer band).
"""
import ctypes as C

import numpy as np


def request_slice(total, rank, world):
    """(first, count, per_rank) of the contiguous request slice a rank lights in collective mode; mirrors csrc/engine.cpp
    slice_len(): ceil(total / world) rounded up to whole 4-request CTAs."""
    per = ((total + world - 1) // world + 3) & ~3
    first = min(total, per * rank)
    last = min(total, per * (rank + 1))
    return first, last - first, per


def peer_ctas(total, rank, world):
    """the 4-request CTAs replica `rank` lights in peer mode: rank, rank + world, ... (csrc/engine.cpp light_compute)."""
    return range(rank, (total + 3) // 4, world)


def peer_rows(group_rows, rank, world):
    """the 16-pixel group rows replica `rank` draws in peer mode: rank, rank + world, ... (csrc/engine.cpp DN_draw)."""
    return range(rank, group_rows, world)


def row_band(group_rows, rank, world):
    """(begin, end, per_rank) in units of 16-pixel rows; mirrors csrc/engine.cpp DN_draw."""
    per = (group_rows + world - 1) // world
    return min(group_rows, per * rank), min(group_rows, per * (rank + 1)), per


def gather_slices(dist, buf, rank, world, slice_len):
    """in-place all-gather: buf is [world * slice_len]; rank r contributed buf[r*slice_len:(r+1)*slice_len]."""
    if slice_len == 0 or world == 1:
        return
    mine = buf[rank * slice_len:(rank + 1) * slice_len].clone()
    dist.all_gather_into_tensor(buf[:world * slice_len], mine)


def gather_bands(dist, torch, image, rank, world, band_len):
    """all-gather of framebuffer bands; image is the flat framebuffer, its length need not be a multiple of band_len."""
    if world == 1:
        return
    total = image.numel()
    padded = torch.empty(band_len * world, dtype=image.dtype, device=image.device)
    lo, hi = min(total, rank * band_len), min(total, (rank + 1) * band_len)
    mine = torch.zeros(band_len, dtype=image.dtype, device=image.device)
    mine[:hi - lo].copy_(image[lo:hi])
    dist.all_gather_into_tensor(padded, mine)
    image.copy_(padded[:total])


def or_reduce_bitmaps(dist, torch, bitmap, world, or_into=None):
    """bitmap |= every other rank's bitmap.  `or_into(ptr)` merges a gathered device bitmap with the library's own
    kernel (DN_b200_or_bitmap); without it (CPU tensors in the tests) the OR is done by torch."""
    if world == 1:
        return
    n = bitmap.numel()
    gathered = torch.empty(n * world, dtype=bitmap.dtype, device=bitmap.device)
    dist.all_gather_into_tensor(gathered, bitmap.clone())
    for r in range(world):
        part = gathered[r * n:(r + 1) * n]
        if or_into is not None:
            or_into(part.data_ptr())
        else:
            bitmap |= part
    return gathered


class _DevicePtr:
    """exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


class ShardedEngine:
    """drives one replica (a doonengine_b200.Engine) as rank `rank` of `world`."""

    def __init__(self, engine, rank, world, torch, dist, device, exchange="peer"):
        from . import ARRAY_PROPAGATE, ARRAY_STAGING, ARRAY_VISIBLE
        self.e, self.rank, self.world, self.torch, self.dist, self.device = engine, rank, world, torch, dist, device
        self.L = engine.L
        self.A_VISIBLE, self.A_STAGING, self.A_PROPAGATE = ARRAY_VISIBLE, ARRAY_STAGING, ARRAY_PROPAGATE
        self.exchange = exchange if world > 1 else "none"
        self._opened = []     # peer mappings to close
        self._mirrors = {}    # fb -> opened root image
        if self.exchange == "peer":
            self._attach()
        elif not self.L.DN_b200_set_shard(engine.vol, rank, world):
            raise ValueError("bad shard %d/%d" % (rank, world))

    # ---- peer memory plumbing: done once (and again only if the staging arrays must grow) ----
    def _all_gather_bytes(self, payload):
        """every rank's `payload` (bytes, same length everywhere), as a list indexed by rank."""
        t = self.torch.tensor(list(payload), dtype=self.torch.uint8, device=self.device)
        out = self.torch.empty(len(payload) * self.world, dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        raw = bytes(out.cpu().numpy().tobytes())
        return [raw[r * len(payload):(r + 1) * len(payload)] for r in range(self.world)]

    def _attach(self, request_cap=0):
        from . import PEER_AUTO, DNb200peerBuffers
        L, e = self.L, self.e
        mine = DNb200peerBuffers()
        if not L.DN_b200_peer_prepare(e.vol, request_cap, C.byref(mine)):
            raise RuntimeError("DN_b200_peer_prepare failed")
        fields = ("staging", "mailbox", "visible", "propagate")
        payload = b""
        for f in fields:
            h = C.create_string_buffer(64)
            if not L.DN_b200_ipc_export(getattr(mine, f), h):
                raise RuntimeError("DN_b200_ipc_export(%s) failed" % f)
            payload += h.raw
        payload += int(mine.stagingRequestCap).to_bytes(8, "little")
        table = (DNb200peerBuffers * self.world)()
        for r, raw in enumerate(self._all_gather_bytes(payload)):
            if r == self.rank:
                table[r] = mine
                continue
            for i, f in enumerate(fields):
                ptr = L.DN_b200_ipc_open(raw[64 * i:64 * (i + 1)])
                if not ptr:
                    raise RuntimeError("DN_b200_ipc_open(%s of rank %d) failed: %s" % (f, r, _last_message()))
                self._opened.append(ptr)
                setattr(table[r], f, ptr)
            table[r].stagingRequestCap = int.from_bytes(raw[256:264], "little")
        if not L.DN_b200_peer_attach(e.vol, self.rank, self.world, table, PEER_AUTO):
            raise RuntimeError("DN_b200_peer_attach failed: %s" % _last_message())
        self.dist.barrier()  # nobody posts into a mailbox that is not attached yet

    def _detach(self):
        self.L.DN_b200_peer_detach(self.e.vol)
        self.dist.barrier()  # every replica has stopped using the mappings
        for fb in list(self._mirrors):
            self.L.DN_b200_framebuffer_set_mirror(fb, None)
        for ptr in self._opened:
            self.L.DN_b200_ipc_close(ptr)
        self._opened, self._mirrors = [], {}

    def mirror_framebuffer(self, fb, root=0):
        """peer mode: every pixel this rank draws into `fb` is also stored into the root's framebuffer of the same size."""
        if self.exchange != "peer":
            return
        L = self.L
        h = C.create_string_buffer(64)
        if self.rank == root and not L.DN_b200_ipc_export(L.DN_b200_framebuffer_device_ptr(fb), h):
            raise RuntimeError("DN_b200_ipc_export(framebuffer) failed")
        raw = self._all_gather_bytes(h.raw)[root]
        if self.rank != root:
            ptr = L.DN_b200_ipc_open(raw)
            if not ptr:
                raise RuntimeError("DN_b200_ipc_open(framebuffer) failed: %s" % _last_message())
            self._opened.append(ptr)
            self._mirrors[fb] = ptr
            L.DN_b200_framebuffer_set_mirror(fb, ptr)
        self.dist.barrier()

    def close(self):
        if self.exchange == "peer" and self.e.vol:
            self._detach()

    def barrier_status(self):
        ep, to = C.c_uint64(), C.c_uint32()
        self.L.DN_b200_peer_barrier_status(self.e.vol, C.byref(ep), C.byref(to))
        return int(ep.value), int(to.value)

    # ---- helpers of the collective mode ----
    def _tensor(self, ptr, nbytes):
        return self.torch.as_tensor(_DevicePtr(ptr, nbytes), device=self.device)

    def _bitmap(self, which):
        nbytes = self.L.DN_b200_array_bytes(self.e.vol, which)
        return self._tensor(self.L.DN_b200_array_device_ptr(self.e.vol, which), nbytes)

    # ---- frame ----
    def draw(self, fb, view, proj):
        """DN_draw of this rank's rows + exchange.  Afterwards every rank holds all visible bits; the whole image is on every
        rank (collective mode) or on the root of a mirrored framebuffer (peer mode)."""
        L, e = self.L, self.e
        L.DN_draw(e.vol, fb, view, proj, -1, -1)
        if self.exchange != "collective":
            return
        w, h = C.c_int(), C.c_int()
        L.DN_b200_framebuffer_size(fb, C.byref(w), C.byref(h))
        _, _, per = row_band(h.value // 16, self.rank, self.world)
        image = self._tensor(L.DN_b200_framebuffer_device_ptr(fb), w.value * h.value * 16)
        gather_bands(self.dist, self.torch, image, self.rank, self.world, per * 16 * w.value * 16)
        or_reduce_bitmaps(self.dist, self.torch, self._bitmap(self.A_VISIBLE), self.world,
                          or_into=lambda p: L.DN_b200_or_bitmap(e.vol, self.A_VISIBLE, p))

    def sync(self, op, split=1):
        self.L.DN_sync_gpu(self.e.vol, op, split)

    def light_compute(self, num_diffuse, max_diffuse, time):
        if self.exchange == "peer" and not self.L.DN_b200_peer_capacity_ok(self.e.vol):
            # the request list outgrew the staging arrays (identically on every rank): remap with room to spare
            self._detach()
            self._attach(request_cap=0)  # 0: sized from what is resident now, with room to spare (DN_b200_peer_prepare)
        if not self.L.DN_b200_light_compute(self.e.vol, num_diffuse, max_diffuse, C.c_float(time)):
            raise RuntimeError("DN_b200_light_compute failed: %s" % _last_message())

    def light_exchange(self):
        if self.exchange != "collective":
            return  # peer mode: the lighting kernel has already stored into every replica
        L, e = self.L, self.e
        slice_bytes = L.DN_b200_staging_slice_bytes(e.vol)
        if slice_bytes:
            staging = self._tensor(L.DN_b200_array_device_ptr(e.vol, self.A_STAGING), slice_bytes * self.world)
            gather_slices(self.dist, staging, self.rank, self.world, slice_bytes)
        or_reduce_bitmaps(self.dist, self.torch, self._bitmap(self.A_PROPAGATE), self.world,
                          or_into=lambda p: L.DN_b200_or_bitmap(e.vol, self.A_PROPAGATE, p))

    def light_commit(self):
        if not self.L.DN_b200_light_commit(self.e.vol):
            raise RuntimeError("DN_b200_light_commit failed")

    def update_lighting(self, num_diffuse=1, max_diffuse=1000, time=1.0):
        self.light_compute(num_diffuse, max_diffuse, time)
        self.light_exchange()
        self.light_commit()


def _last_message():
    from . import messages
    m = messages()
    return m[-1][2] if m else "?"
