"""N-GPU check of the reference-facing entry point, run under torchrun (2+ GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/scene_multi_gpu_check.py

Scene.finish() sharded over the ranks (tiles, samples, ReSTIR in row strips with a 24-pixel halo = ONE reservoir chain
over all GPUs, ReSTIR with one chain per GPU) against the same Scene rendered on one GPU by rank 0: the tile-sharded
and the row-sharded ReSTIR images must be IDENTICAL, the sample-sharded one equal up to float re-association of the
per-pixel sums, the per-GPU-chain ReSTIR one a valid image with the same mean within noise. Also checks parallel.FusedMerge against all-reduce + tonemap on a double-buffered
sequence of batches (the protocol bench.py uses)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import scenes  # noqa: E402
import voxel_rt2_b200 as vrt  # noqa: E402
from voxel_rt2_b200 import parallel  # noqa: E402
from voxel_rt2_b200.scene import Scene  # noqa: E402


def make_scene(mode, shard):
    os.environ["VRT_RES"], os.environ["VRT_GRID"], os.environ["VRT_MODE"], os.environ["VRT_SHARD"] = "512x256", "64", mode, shard
    os.environ["VRT_SEED"] = "3"
    s = Scene(voxel_edges=0.06, exposure=2.0)
    mat, col = scenes.material_zoo(64)
    s.voxel_material[:], s.voxel_color[:] = mat, col
    s.set_floor(-0.6, (0.8, 0.8, 0.8))
    s.set_directional_light((1, 1, 0.3), 0.05, (1.0, 0.95, 0.9))
    s.set_background_color((0.3, 0.4, 0.6))
    return s


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {}
    # ---- Scene.finish on N GPUs vs one GPU
    for mode, shard, spp in (("pt", "tiles", 16), ("pt", "samples", 16), ("restir", "rows", 8), ("restir", "samples", 8)):
        img = make_scene(mode, shard).finish(spp=spp, out="")
        if rank == 0:
            save = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE")}
            one = make_scene(mode, shard).finish(spp=spp, out="")
            os.environ.update(save)
            key = "%s_%s" % (mode, shard)
            out[key + "_identical"] = bool(np.array_equal(img, one))
            out[key + "_max_abs_diff"] = float(np.abs(img - one).max())
            out[key + "_mean_ratio"] = float(img[..., :3].mean() / one[..., :3].mean())
        dist.barrier()
    # ---- sharded sky precompute (rows of the tables split over the ranks + all-gather) vs the whole precompute on one GPU
    def sky(shard):
        g = vrt.Renderer(dx=2.0 / 32, image_res=(64, 32), grid_res=32, sky_res=256, cloud_passes=2, seed=2, device=local)
        g.set_voxels(*scenes.random_grid(32, 0.2, 3))
        g.set_directional_light((1, 1, -1), 0.025, (1.3, 1.2, 1.2))
        g.set_use_physical_sky(True, True)
        if shard:
            g.set_sky_shard(rank, world)
        g.prepare_data()
        return g.get_sky_tables(), g.stats()["sky_precompute_ms"]

    (sa, ta), ms_shard = sky(True)
    (sb, tb), ms_full = sky(False)
    same_sky = bool(np.array_equal(sa, sb) and np.array_equal(ta, tb))
    flags = [None] * world
    dist.all_gather_object(flags, same_sky)
    if rank == 0:
        out["sharded_sky_identical_on_every_rank"] = bool(all(flags))
        out["sky_precompute_ms_sharded_vs_full"] = [ms_shard, ms_full]
    # ---- FusedMerge over a sequence of double-buffered batches vs all-reduce + tonemap
    R, W, H = 64, 512, 256
    r = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=0, seed=7, device=local)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    r.set_voxels(*scenes.random_grid(R, 0.3, 5))
    r.set_directional_light((1, 1, 0.5), 0.05, (1.2, 1.1, 1.0))
    r.set_background_color((0.3, 0.4, 0.6))
    r.set_sample_shard(rank, world)
    r.prepare_data()
    hosts = [torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(4)]
    with torch.cuda.stream(stream):
        fm = parallel.FusedMerge(r)
        for k in range(4):
            fm.begin(k)
            r.current_spp = 4 * k  # a different sample set per batch (reset_framebuffer rewinds the sample counter)
            r.accumulate(4)
            fm.merge()
            if rank == 0 and k > 0:
                fm.copy_previous(hosts[k - 1])
        fm.finish(hosts[3] if rank == 0 else None)
    torch.cuda.synchronize()
    fused = [h.copy() for h in hosts]
    # the same batches through NCCL all-reduce + the plain tonemap pass
    r.set_accum_slot(0)
    ok = []
    for k in range(4):
        with torch.cuda.stream(stream):
            r.reset_framebuffer()
            r.current_spp = 4 * k
            r.accumulate(4)
            acc = r.accum_tensor()
            dist.all_reduce(acc)
        torch.cuda.synchronize()
        if rank == 0:
            ref = r.fetch_image()
            ok.append(float(np.abs(ref - fused[k]).max()))
        dist.barrier()
    fm.close()
    if rank == 0:
        out["fused_vs_allreduce_max_abs_diff_per_batch"] = ok
        out["world"] = world
        print(json.dumps(out))
        assert out["pt_tiles_identical"] and out["restir_rows_identical"] and out["sharded_sky_identical_on_every_rank"], out
        assert out["pt_samples_max_abs_diff"] < 1e-5, out
        assert abs(out["restir_samples_mean_ratio"] - 1.0) < 0.05, out
        assert max(ok) < 1e-6, out
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
