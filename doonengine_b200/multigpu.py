"""One process per GPU: the map is replicated, the lighting-request list and the screen rows are sharded.

Per frame, on every rank:

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
    """(first, count, per_rank) of the contiguous request slice a rank lights; mirrors csrc/engine.cpp light_compute."""
    per = (total + world - 1) // world
    first = min(total, per * rank)
    last = min(total, per * (rank + 1))
    return first, last - first, per


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

    def __init__(self, engine, rank, world, torch, dist, device):
        from . import ARRAY_PROPAGATE, ARRAY_STAGING, ARRAY_VISIBLE
        self.e, self.rank, self.world, self.torch, self.dist, self.device = engine, rank, world, torch, dist, device
        self.L = engine.L
        self.A_VISIBLE, self.A_STAGING, self.A_PROPAGATE = ARRAY_VISIBLE, ARRAY_STAGING, ARRAY_PROPAGATE
        if not self.L.DN_b200_set_shard(engine.vol, rank, world):
            raise ValueError("bad shard %d/%d" % (rank, world))

    def _tensor(self, ptr, nbytes):
        return self.torch.as_tensor(_DevicePtr(ptr, nbytes), device=self.device)

    def _bitmap(self, which):
        nbytes = self.L.DN_b200_array_bytes(self.e.vol, which)
        return self._tensor(self.L.DN_b200_array_device_ptr(self.e.vol, which), nbytes)

    def draw(self, fb, view, proj):
        """DN_draw of this rank's band + exchange; afterwards every rank holds the whole image and all visible bits."""
        L, e = self.L, self.e
        L.DN_draw(e.vol, fb, view, proj, -1, -1)
        if self.world == 1:
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
        if not self.L.DN_b200_light_compute(self.e.vol, num_diffuse, max_diffuse, C.c_float(time)):
            raise RuntimeError("DN_b200_light_compute failed")

    def light_exchange(self):
        if self.world == 1:
            return
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
