"""Render group: N ranks, one process and one context per GPU of a single NVSwitch box.

Paths are independent and the scene is replicated, so there is no collective on the data path while rendering;
the only exchange is the combination of the per-rank accumulators at read-out and the 32-byte edit records.
The reference is single-GPU (one GL context, renderer/renderer.cpp:556-645); this module is the new build's
plumbing around it and uses torch.distributed for the transport (NCCL over NVLink on the GPUs, gloo in the CPU tests).

  TILES    64x64-pixel tiles dealt round-robin (tile t -> rank t % N, the rule of vt_set_partition / wf_item_pixel).
           Every rank uses global pixel coordinates and global sampleCount, so each pixel's running average has the
           bits a single GPU would produce; ranks hold zeros outside their tiles, and the exchange is a SUM whose
           every addend but one is +0.0 -- exact.
  SAMPLES  rank r renders sampleCount = p * N + r for its p-th local pass and keeps a float4 SUM; the exchange is a
           SUM-reduce followed by one division by the total pass count on the destination rank (fp-sum order differs
           from the reference's running average: equal within 1e-5 relative).
"""
import os

import numpy as np

PART_NONE, PART_TILES, PART_SAMPLES = 0, 1, 2
TILE = 64


class RenderGroup:
    def __init__(self, mode, rank=0, world=1, process_group=None):
        if mode not in (PART_NONE, PART_TILES, PART_SAMPLES) or not (0 <= rank < world):
            raise ValueError("bad render group (mode %r, rank %r of %r)" % (mode, rank, world))
        self.mode = PART_NONE if world == 1 else mode
        self.rank, self.world, self.pg = rank, world, process_group

    # ---- partition rules (mirrored by csrc/vt_api.cu vt_render and csrc/vt_wavefront.cuh wf_item_pixel) ------------
    def sample_index(self, local_pass):
        """Global `sampleCount` uniform of this rank's local_pass-th pass."""
        return local_pass * self.world + self.rank if self.mode == PART_SAMPLES else local_pass

    def owns_tile(self, tile):
        return self.mode != PART_TILES or tile % self.world == self.rank

    def tile_mask(self, width, height):
        """(H, W) bool: pixels this rank renders."""
        tx = (width + TILE - 1) // TILE
        ys, xs = np.mgrid[0:height, 0:width]
        tiles = (xs // TILE) + (ys // TILE) * tx
        return np.ones((height, width), bool) if self.mode != PART_TILES else (tiles % self.world) == self.rank

    def apply(self, renderer):
        """Tell a host Renderer (voxeltoy_b200.host.Renderer) which share of the frame / samples it owns."""
        renderer.setPartition(self.mode, self.rank, self.world)

    # ---- exchange -------------------------------------------------------------------------------------------------
    def combine(self, accum, local_passes, dst=0, out=None):
        """Combine the per-rank accumulators (torch tensor, (H, W, 4) float32, on any device the process group's backend
        supports). Returns the finished image on rank `dst` (None elsewhere). `accum` itself is left untouched so that
        progressive rendering can continue.
        SAMPLES: SUM-reduce + one division (`out`: optional staging buffer). TILES: every rank contributes only the tiles it
        owns -- a compact (tiles_per_rank, 64, 64, 4) buffer, W*H*16/world bytes -- through one all_gather; no SUM of zeros."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return accum
        if self.mode == PART_TILES:
            H, W = int(accum.shape[0]), int(accum.shape[1])
            ty, tx = (H + TILE - 1) // TILE, (W + TILE - 1) // TILE
            per_rank = (ty * tx + self.world - 1) // self.world
            padded = torch.zeros((ty * TILE, tx * TILE, 4), dtype=accum.dtype, device=accum.device)
            padded[:H, :W] = accum
            tiles = padded.view(ty, TILE, tx, TILE, 4).permute(0, 2, 1, 3, 4).reshape(ty * tx, TILE, TILE, 4)
            mine = torch.zeros((per_rank, TILE, TILE, 4), dtype=accum.dtype, device=accum.device)
            own = tiles[self.rank::self.world]
            mine[:own.shape[0]] = own
            gathered = torch.empty((self.world * per_rank, TILE, TILE, 4), dtype=accum.dtype, device=accum.device)
            dist.all_gather_into_tensor(gathered, mine, group=self.pg)
            if self.rank != dst:
                return None
            full = torch.empty_like(tiles)
            for r in range(self.world):
                n_r = full[r::self.world].shape[0]
                full[r::self.world] = gathered[r * per_rank: r * per_rank + n_r]
            return full.view(ty, tx, TILE, TILE, 4).permute(0, 2, 1, 3, 4).reshape(ty * TILE, tx * TILE, 4)[:H, :W].contiguous()
        buf = out if out is not None else torch.empty_like(accum)
        buf.copy_(accum)
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=self.pg)
        if self.rank != dst:
            return None
        buf.div_(float(local_passes * self.world))
        return buf

    def broadcast_action(self, record, src=0):
        """Edit / pick requests are issued on one rank (Renderer::requestAction, renderer/actions.cpp:5-18) and must reach
        every replica of the scene: 8 floats (x, y, dx, dy, action, restart, 0, 0)."""
        import torch
        import torch.distributed as dist
        t = torch.as_tensor(np.asarray(record, np.float32).reshape(8).copy())
        if self.world > 1:
            backend = dist.get_backend(self.pg)
            if backend == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=src, group=self.pg)
        return t.cpu().numpy()


class DeviceGroup:
    """Thin binding of the C ABI's render groups (include/voxeltoy_b200.h, vt_group_*): the partition, the NCCL / peer-memory
    combination of the accumulators and the replication of edits live in libvoxeltoy_b200.so (csrc/vt_group.inl).

      DeviceGroup.adopt(contexts, mode)                    one process, several contexts (rank = position)
      DeviceGroup.join(context, id128, rank, world, mode)  one process per GPU; id128 = DeviceGroup.unique_id() of rank 0,
                                                           handed round by the launcher's own plumbing
    """
    EXCHANGE_NCCL, EXCHANGE_PEER = 0, 1

    def __init__(self, handle, lib, contexts):
        self.h, self.lib, self.contexts = handle, lib, contexts

    _nccl_preloaded = False

    @classmethod
    def _lib(cls):
        """The library binds NCCL at run time by soname (dlopen("libnccl.so.2")). In a Python process the copy that must win is
        the one PyTorch ships (a later `import torch` would otherwise be handed the older system library under the same
        soname), so it is loaded first when it exists."""
        from . import _capi
        if not cls._nccl_preloaded:
            cls._nccl_preloaded = True
            import ctypes
            import glob
            import site
            import sys
            if "torch" not in sys.modules:
                roots = list(site.getsitepackages()) + [p for p in sys.path if p.endswith("site-packages")]
                for root in roots:
                    hits = glob.glob(os.path.join(root, "nvidia", "nccl", "lib", "libnccl.so.2"))
                    if hits:
                        try:
                            ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)
                        except OSError:
                            pass
                        break
        return _capi.load()

    @classmethod
    def adopt(cls, contexts, mode):
        import ctypes as C
        from . import _capi
        lib = cls._lib()
        arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
        h = C.c_void_p()
        rc = lib.vt_group_adopt(len(contexts), arr, int(mode), C.byref(h))
        if rc != 0:
            raise _capi.VtError("vt_group_adopt failed with status %d" % rc)
        return cls(h, lib, list(contexts))

    @classmethod
    def unique_id(cls):
        import ctypes as C
        from . import _capi
        buf = (C.c_ubyte * 128)()
        rc = cls._lib().vt_group_unique_id(buf)
        if rc != 0:
            raise _capi.VtError("vt_group_unique_id failed with status %d (is libnccl.so.2 loadable?)" % rc)
        return bytes(buf)

    @classmethod
    def join(cls, context, id128, rank, world, mode):
        import ctypes as C
        from . import _capi
        lib = cls._lib()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(id128))
        h = C.c_void_p()
        rc = lib.vt_group_join(context.h, buf, int(rank), int(world), int(mode), C.byref(h))
        if rc != 0:
            raise _capi.VtError("vt_group_join failed with status %d: %s" % (rc, lib.vt_last_error(context.h).decode()))
        return cls(h, lib, [context])

    def _ck(self, rc):
        if rc != 0:
            from . import _capi
            raise _capi.VtError("status %d: %s" % (rc, self.lib.vt_group_last_error(self.h).decode()))

    def close(self):
        if self.h:
            self.lib.vt_group_destroy(self.h)
            self.h = None

    def size(self):
        return self.lib.vt_group_size(self.h)

    def set_exchange(self, exchange):
        self._ck(self.lib.vt_group_set_exchange(self.h, int(exchange)))

    def exchange(self):
        return self.lib.vt_group_get_exchange(self.h)

    def render(self, first_sample, n_passes):
        self._ck(self.lib.vt_group_render(self.h, int(first_sample), int(n_passes)))

    def reset_accumulation(self):
        self._ck(self.lib.vt_group_reset_accumulation(self.h))

    def sync(self):
        self._ck(self.lib.vt_group_sync(self.h))

    def begin_combine(self):
        self._ck(self.lib.vt_group_begin_combine(self.h))

    def wait_combine(self):
        """Stream-level wait: the contexts' streams continue after the exchange (no host synchronisation)."""
        self._ck(self.lib.vt_group_wait_combine(self.h))

    def result_device_ptr(self):
        return self.lib.vt_group_result_device_ptr(self.h)

    def end_combine(self, out=None, want=True):
        """Waits for the exchange. In the process that holds rank 0 returns the frame ((H, W, 4) float32; `out`: a host array
        or pinned tensor's numpy view to fill); with want=False the frame stays on the device."""
        import ctypes as C
        ptr = None
        if want:
            c0 = self.contexts[0]
            if out is None:
                out = np.empty((c0.height, c0.width, 4), np.float32)
            ptr = out.ctypes.data_as(C.POINTER(C.c_float))
        self._ck(self.lib.vt_group_end_combine(self.h, ptr))
        return out

    def read_average(self, out=None):
        self.begin_combine()
        return self.end_combine(out)

    def last_exchange_ms(self):
        import ctypes as C
        ms = C.c_float()
        self._ck(self.lib.vt_group_last_exchange_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def exchange_bytes(self):
        return int(self.lib.vt_group_exchange_bytes(self.h))

    def pick(self, px, py):
        self._ck(self.lib.vt_group_pick(self.h, float(px), float(py)))

    def pick_focal(self, px, py):
        self._ck(self.lib.vt_group_pick_focal(self.h, float(px), float(py)))

    def add_voxel(self, mx, my):
        self._ck(self.lib.vt_group_add_voxel(self.h, float(mx), float(my)))

    def remove_voxel(self):
        self._ck(self.lib.vt_group_remove_voxel(self.h))

    def broadcast(self, record, root=0):
        """<= 256 bytes (a numpy array, modified in place on the other ranks) from the process holding rank `root`."""
        import ctypes as C
        a = np.ascontiguousarray(record)
        self._ck(self.lib.vt_group_broadcast(self.h, a.ctypes.data_as(C.c_void_p), a.nbytes, int(root)))
        return a
