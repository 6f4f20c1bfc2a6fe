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


def balanced_row_cuts(row_cost, world):
    """Cut positions (world + 1 tile-row indices) that split the per-tile-row costs into `world` contiguous strips of
    nearly equal total cost. Deterministic: every rank computes the same cuts from the same costs."""
    import numpy as np

    c = np.concatenate([[0.0], np.cumsum(np.asarray(row_cost, np.float64))])
    rows = len(row_cost)
    cuts = [0]
    for k in range(1, world):
        target = c[-1] * k / world
        r = int(np.searchsorted(c, target))
        r = min(max(r, cuts[-1] + 1), rows - (world - k))  # every strip keeps at least one row
        cuts.append(r)
    cuts.append(rows)
    return cuts


def shard_rows(renderer, rank, world, balance=True):
    """Contiguous strips of tile rows: the partition for the ReSTIR mode (one reservoir chain over all GPUs; each rank
    renders a 24-pixel halo around its rows for the spatial pass). With balance=True (call after prepare_data) the
    strips are cut by cost instead of by height: a primary-hit pass counts the geometry pixels of every tile row (sky
    pixels cost the resampling passes next to nothing), and the halo rows a strip has to render are charged to it."""
    if not balance or world == 1:
        renderer.set_row_shard(rank, world)
        return None
    import numpy as np

    hits = renderer.trace_primary()
    H = hits.shape[0]
    geo = ((hits["flags"] & 255) > 0).reshape(H // 4, 4, -1).sum(axis=(1, 2)).astype(np.float64)
    cost = geo + 0.05 * hits.shape[1] * 4          # a sky pixel still costs a primary ray and a pass-through
    cuts = balanced_row_cuts(cost, world)
    renderer.set_row_range(cuts[rank], cuts[rank + 1] - cuts[rank])
    return cuts


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


class FusedMerge:
    """Per-frame merge of the sample- or tile-sharded accumulation buffers WITHOUT an all-reduce: a fused
    reduce-scatter + tonemap over NVLink peer memory, spread evenly over the ranks (SURVEY.md §8e "fused
    variant"). Rank r owns the contiguous pixel slice [r·npx/N, (r+1)·npx/N): one kernel
    (`vrt_merge_slice`) adds the N partial sums of that slice — its own from HBM, the others through
    CUDA-IPC peer mappings — applies the tonemap and stores the pixels straight into the displaying
    rank's image buffer (a peer store). NVLink traffic per rank and frame: (N-1)/N x 16 B/pixel read +
    16 B/pixel/N written, against 2 (N-1)/N x 16 B/pixel for an all-reduce, and the only collective
    left is a 4-byte NCCL all-reduce used as a stream-ordered barrier.

    Batches alternate between the two accumulation / image slots of the library, so batch k+1 renders
    while the peers still read batch k. Per step, on every rank:

        fm.begin(k)                 # slot k & 1, reset_framebuffer (deferred: the kernel overwrites)
        r.accumulate(spp)
        fm.merge()                  # barrier k (all partial sums of batch k complete) + merge of the own slice
        if rank == 0: fm.copy_previous(pinned)   # image of batch k-1: complete since barrier k

    Hazards and what orders them (all in stream order, no host synchronisation):
      * peers read slot s of batch k only after barrier k;
      * batch k+2 overwrites slot s on a rank only after barrier k+1, which completes after every rank's
        merge of batch k;
      * rank 0 copies image slot s of batch k after barrier k+1 (every rank's merge k precedes it) and
        joins barrier k+2 only after that copy (`stream_wait_copy`), so the peers' stores of batch k+2
        cannot overtake it.
    """

    def __init__(self, renderer, group=None, display_rank=0):
        import torch
        import torch.distributed as dist

        self.r, self.group, self.dst = renderer, group, display_rank
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        W, H = renderer.image_res
        npx = W * H
        self.first = self.rank * npx // self.world
        self.count = (self.rank + 1) * npx // self.world - self.first
        # Set-up is collective and must fail on EVERY rank or on none (a rank that raised before a collective would
        # leave the others waiting in it): local failures are gathered first, then raised everywhere.
        mine, err = [], None
        try:
            for s in (0, 1):
                renderer.set_accum_slot(s)
                mine.append((renderer.accum_ipc_handle(), renderer.out_ipc_handle()))
            renderer.set_accum_slot(0)
        except Exception as e:  # noqa: BLE001
            mine, err = None, repr(e)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if any(h is None for h in everyone):
            raise RuntimeError("FusedMerge: a rank could not export its buffers (%s)" % (err or "see the other ranks"))
        self.peer_accum = [[], []]   # per slot: the other ranks' accumulation buffers, in rank order
        self.peer_out = [None, None]  # per slot: the displaying rank's image buffer (None on that rank itself)
        try:
            for s in (0, 1):
                for k, h in enumerate(everyone):
                    if k != self.rank:
                        self.peer_accum[s].append(renderer.open_peer_accum(h[s][0]))
                if self.rank != display_rank:
                    self.peer_out[s] = renderer.open_peer_accum(everyone[display_rank][s][1])
        except Exception as e:  # noqa: BLE001  (no peer access between two GPUs, for one)
            err = repr(e)
        oks = [None] * self.world
        dist.all_gather_object(oks, err, group=group)
        if any(x is not None for x in oks):
            for s in (0, 1):
                for p in self.peer_accum[s]:
                    renderer.close_peer_accum(p)
                if self.peer_out[s]:
                    renderer.close_peer_accum(self.peer_out[s])
            raise RuntimeError("FusedMerge: peer mapping failed on a rank (%s)" % next(x for x in oks if x is not None))
        self._token = torch.zeros(1, device=torch.device("cuda", renderer.device))
        self.slot, self.k = 0, -1

    def begin(self, k, reset=True):
        """Batch k goes to slot k & 1. reset=False keeps what the slot has accumulated (and the ReSTIR reservoir history,
        which reset_framebuffer drops): the slot then holds the running sum of every second batch."""
        self.k, self.slot = k, k & 1
        self.r.set_accum_slot(self.slot)
        if reset:
            self.r.reset_framebuffer()

    def barrier(self):
        """Stream-ordered barrier: a 4-byte NCCL all-reduce on the current torch stream."""
        import torch.distributed as dist

        dist.all_reduce(self._token, group=self.group)

    def merge(self):
        if self.rank == self.dst:
            self.r.stream_wait_copy()
        self.barrier()
        self.r.merge_slice(self.peer_accum[self.slot], self.first, self.count, ldr_dst=self.peer_out[self.slot])

    def copy_previous(self, out_pinned):
        """Displaying rank, after merge() of batch k: image of batch k-1 -> pinned host memory (copy engine)."""
        self.r.set_accum_slot(self.slot ^ 1)
        self.r.copy_image_async(out_pinned)
        self.r.set_accum_slot(self.slot)

    def finish(self, out_pinned=None):
        """After the last batch: one more barrier so every rank's last merge has landed; the displaying rank
        then copies the last image."""
        self.barrier()
        if self.rank == self.dst and out_pinned is not None:
            self.r.copy_image_async(out_pinned)
            self.r.wait_image()

    def close(self):
        """Unmap the peers' buffers; collective: nobody frees a buffer a peer still has mapped."""
        import torch.distributed as dist

        for s in (0, 1):
            for p in self.peer_accum[s]:
                self.r.close_peer_accum(p)
            if self.peer_out[s]:
                self.r.close_peer_accum(self.peer_out[s])
        self.peer_accum, self.peer_out = [[], []], [None, None]
        dist.barrier(group=self.group)
