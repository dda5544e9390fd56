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
        progressive rendering can continue; pass `out` to reuse a staging buffer."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return accum
        buf = out if out is not None else torch.empty_like(accum)
        buf.copy_(accum)
        dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM, group=self.pg)
        if self.rank != dst:
            return None
        if self.mode == PART_SAMPLES:
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
