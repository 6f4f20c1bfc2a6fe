"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed for
the plumbing (NCCL over NVLink on the box, gloo in CPU tests).

Every (pixel, sample) is independent (pathtracer.py:360-632 reads no neighbour), so the path
shards with NO data-path collective; the only exchange is the merge of the float4 accumulation
buffer, once per frame batch:
  * sample sharding: rank r renders sample indices r, r+N, ...; merge = all-reduce(sum)
  * tile sharding:   rank r renders the 8x4 tiles with tile_id % N == r (others stay zero);
                     merge = all-reduce(sum) as well (disjoint support => exact, bit-identical
                     to the unsharded image)."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_samples(renderer, rank, world):
    """Weak scaling: per-rank work stays `spp` per accumulate(), sample indices interleave."""
    renderer.set_sample_shard(rank, world)


def shard_tiles(renderer, rank, world):
    """Strong scaling of one frame: interleaved 8x4 tiles (balanced sky / geometry load)."""
    renderer.set_tile_shard(rank, world)


def merge_accumulation(accum, group=None):
    """One collective per frame batch: sum the [H, W, 4] accumulation buffers (rgb sums and the
    per-pixel sample count in w) of all ranks, in place."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    return accum


def mean_from_accumulation(accum):
    """[H, W, 4] sums -> mean radiance (rgb / w), w kept."""
    import torch

    w = accum[..., 3:4]
    out = accum.clone()
    out[..., :3] = torch.where(w > 0, accum[..., :3] / w.clamp_min(1.0), torch.zeros_like(accum[..., :3]))
    return out


class PeerMerge:
    """Fused alternative to merge_accumulation + fetch_image for the displaying rank: rank 0 maps
    the other ranks' accumulation buffers (CUDA IPC over NVLink) and sums them inside the tonemap
    kernel (`vrt_fetch_ldr_merged`). Usage per frame batch, on every rank:

        r.accumulate(spp); pm.ready()          # barrier: all partial sums are complete
        if rank == 0: img = pm.fetch_image()   # one kernel: peer reads + reduce + tonemap
        pm.release()                           # barrier: peers may reset / continue
    """

    def __init__(self, renderer, group=None):
        import torch.distributed as dist

        self.r, self.group = renderer, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        handles = [None] * self.world
        dist.all_gather_object(handles, renderer.accum_ipc_handle(), group=group)
        self.peers = []
        if self.rank == 0:
            self.peers = [renderer.open_peer_accum(h) for k, h in enumerate(handles) if k != 0]

    def ready(self):
        import torch.distributed as dist

        dist.barrier(group=self.group)

    release = ready

    def fetch_image(self, out=None):
        return self.r.fetch_image_merged(self.peers, out)

    def close(self):
        for p in self.peers:
            self.r.close_peer_accum(p)
        self.peers = []
