"""BASELINE.json config 5: 3840x2160, 1024 spp converged render of the largest example scene
(example4's sphere, 319 489 voxels; rebuilt here procedurally: x.x < n*n/2, example4.py:13-17),
tile-sharded over the GPUs of one box (interleaved 8x4 tiles, tile_id % N == rank), merged at the end by
the fused peer-memory gather + tonemap (parallel.FusedMerge: every rank moves W*H/N float4; the partial
buffers have disjoint support, so the sum IS the gather). bench.py runs the same thing as its
`other_configs.config5` leg at N = 8; this script also writes the image.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/render_4k_tiles.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import voxel_rt2_b200 as vrt  # noqa: E402
from voxel_rt2_b200 import parallel  # noqa: E402


def sphere_scene(R=128, n=60):
    i = np.arange(-R // 2, R // 2)
    x, y, z = np.meshgrid(i, i, i, indexing="ij")
    inside = (x * x + y * y + z * z < n * n * 0.5) & (np.abs(x) < n) & (np.abs(y) < n) & (np.abs(z) < n)
    mat = inside.astype(np.int8)
    col = np.zeros((R, R, R, 3), np.uint8)
    col[inside] = (229, 76, 76)  # u8(0.9*255), u8(0.3*255)
    return mat, col


def main():
    rank, world = parallel.rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    spp = int(os.environ.get("VRT_SPP", "1024"))
    W, H = 3840, 2160
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = vrt.Renderer(dx=1 / 64, image_res=(W, H), grid_res=128, sky_res=0, exposure=1.0, seed=5, device=local)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    mat, col = sphere_scene()
    r.set_voxels(mat, col)
    r.set_directional_light((1, 1, 1), 0.1, (1, 1, 1))      # example4.py:6
    r.set_background_color((0.3, 0.4, 0.6))                  # example4.py:7
    parallel.shard_tiles(r, rank, world)
    r.prepare_data()
    fm = parallel.FusedMerge(r) if world > 1 else None
    host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy() if rank == 0 else None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        if fm:
            fm.begin(0)
        done = 0
        while done < spp:
            n = min(64, spp - done)
            r.accumulate(n)
            done += n
        if fm:
            fm.merge()
            fm.finish(host)
        elif rank == 0:
            r.fetch_image_async(host)
            r.wait_image()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        img = host
        out = {"config": "config5: 3840x2160, %d spp, example4 sphere (%d voxels), tile-sharded x%d + fused peer-memory gather" % (spp, int((mat > 0).sum()), world),
               "seconds": dt, "paths_per_s": W * H * spp / dt, "alpha_min": float(img[..., 3].min()), "mean_ldr": float(img[..., :3].mean())}
        print(json.dumps(out))
        os.makedirs("gpurun_out", exist_ok=True)
        try:
            from voxel_rt2_b200.scene import save_image

            save_image(img, "gpurun_out/config5_4k.jpg")
        except Exception as e:
            print("no image:", e)
    if fm:
        fm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
