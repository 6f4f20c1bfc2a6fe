"""N > 1 host logic on CPU: world_size-2 gloo. Each rank renders its sample shard (engine: the
CPU oracle — test infrastructure standing in for the per-rank CUDA context) and the product's
merge helper (voxel_rt2_b200.parallel) all-reduces the accumulation buffers; the merged image
must equal the unsharded render. The same helper runs over NCCL in bench.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import scenes
    from oracle.binding import OracleRenderer
    from voxel_rt2_b200 import parallel
    from voxel_rt2_b200.materials import material_table

    dist.init_process_group("gloo", rank=rank, world_size=world)
    R, res, spp = 32, (64, 32), 4
    o = OracleRenderer(dx=2.0 / R, image_res=res, grid_res=R, sky_res=0, seed=3, materials=material_table())
    o.n_threads = 2
    o.set_voxels(*scenes.random_grid(R, 0.3, 9))
    o.set_background_color((0.5, 0.6, 0.7))
    o.set_directional_light((1, 1, 1), 0.1, (1, 1, 1))
    r, w = parallel.rank_world()
    if mode == "sample":
        parallel.shard_samples(o, r, w)
        o.prepare_data()
        o.accumulate(spp // w)
    else:
        parallel.shard_tiles(o, r, w)
        o.prepare_data()
        o.accumulate(spp)
    h = o.fetch_hdr()
    accum = torch.from_numpy(np.concatenate([h[..., :3] * h[..., 3:4], h[..., 3:4]], axis=-1).copy())
    if mode == "tile":
        # pixels outside this rank's tiles were never rendered
        assert float((accum[..., 3] == 0).float().mean()) > 0.4
    parallel.merge_accumulation(accum)
    mean = parallel.mean_from_accumulation(accum)
    if rank == 0:
        q.put(mean.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["sample", "tile"])
def test_world_size_2_gloo_merge_equals_unsharded(oracle, mode):
    import torch.multiprocessing as mp

    import scenes
    from voxel_rt2_b200.materials import material_table

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "sample" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    R, res, spp = 32, (64, 32), 4
    o = oracle.OracleRenderer(dx=2.0 / R, image_res=res, grid_res=R, sky_res=0, seed=3, materials=material_table())
    o.set_voxels(*scenes.random_grid(R, 0.3, 9))
    o.set_background_color((0.5, 0.6, 0.7))
    o.set_directional_light((1, 1, 1), 0.1, (1, 1, 1))
    o.prepare_data()
    o.accumulate(spp)
    full = o.fetch_hdr()
    assert np.array_equal(merged[..., 3], full[..., 3])
    if mode == "tile":
        assert np.allclose(merged[..., :3], full[..., :3], rtol=1e-6, atol=1e-7)
    else:
        # same sample set, summed in a different order (running means vs sums)
        assert np.allclose(merged[..., :3], full[..., :3], rtol=2e-5, atol=1e-6)


def _fm_worker(rank, world, port, q):
    """FusedMerge set-up with a stand-in renderer: rank 1 cannot export its buffers. Both ranks must raise
    (a rank that failed alone before the handle exchange would leave the other waiting in the collective), and the
    pixel slices of a healthy set-up must tile the frame."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from voxel_rt2_b200 import parallel

    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Fake:
        image_res, device = (30, 7), 0

        def __init__(self, broken):
            self.broken = broken

        def set_accum_slot(self, s):
            pass

        def accum_ipc_handle(self):
            if self.broken:
                raise RuntimeError("no IPC here")
            return b"a" * 64

        def out_ipc_handle(self):
            return b"o" * 64

        def open_peer_accum(self, h):
            raise RuntimeError("no peer access")

        def close_peer_accum(self, p):
            pass

    results = []
    for broken_rank in (1, None):  # export fails on rank 1; then: export fine, peer mapping fails everywhere
        try:
            parallel.FusedMerge(Fake(rank == broken_rank))
            results.append("constructed")
        except RuntimeError as e:
            results.append(str(e)[:40])
    q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


def test_fused_merge_setup_fails_on_every_rank_or_none():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7
    procs = [ctx.Process(target=_fm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert got[rank][0].startswith("FusedMerge: a rank could not export")
        assert got[rank][1].startswith("FusedMerge: peer mapping failed")
    # pixel slices tile the frame for any world size
    for world in (1, 2, 3, 8):
        npx = 1920 * 1080
        cuts = [r * npx // world for r in range(world + 1)]
        assert cuts[0] == 0 and cuts[-1] == npx and all(b > a for a, b in zip(cuts, cuts[1:]))


def test_balanced_row_cuts():
    """parallel.balanced_row_cuts: contiguous strips of nearly equal cost, every strip non-empty, deterministic."""
    from voxel_rt2_b200.parallel import balanced_row_cuts

    cost = np.r_[np.ones(100), np.full(100, 50.0), np.ones(70)]
    cuts = balanced_row_cuts(cost, 8)
    assert cuts[0] == 0 and cuts[-1] == 270 and all(b > a for a, b in zip(cuts, cuts[1:]))
    per = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
    assert max(per) <= 1.15 * sum(per) / 8
    assert balanced_row_cuts(np.ones(24), 5) == [0, 5, 10, 15, 20, 24]
    z = balanced_row_cuts(np.zeros(10), 4)
    assert z[0] == 0 and z[-1] == 10 and all(b > a for a, b in zip(z, z[1:]))
    assert balanced_row_cuts(cost, 1) == [0, 270]
