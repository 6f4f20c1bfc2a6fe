"""N-GPU check of the fused peer-read merge + tonemap (vrt_fetch_ldr_merged) against the NCCL
all-reduce + tonemap path, plus their timings. Run under torchrun on one NVLink box:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_merge_check.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import scenes  # noqa: E402
import voxel_rt2_b200 as vrt  # noqa: E402
from voxel_rt2_b200 import parallel  # noqa: E402


def main():
    rank, world = parallel.rank_world()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R, W, H, spp = 128, 1920, 1080, 4
    r = vrt.Renderer(dx=2.0 / R, image_res=(W, H), grid_res=R, sky_res=0, seed=2, device=local)
    stream = torch.cuda.Stream()
    r.set_stream(stream.cuda_stream)
    r.set_voxels(*scenes.random_grid(R, 0.3, 7))
    r.set_floor(-1e5, (1, 1, 1))
    r.set_directional_light((1, 1, 1), 0.05, (1.2, 1.1, 1.0))
    r.set_background_color((0.3, 0.4, 0.6))
    parallel.shard_samples(r, rank, world)
    r.prepare_data()
    accum = r.accum_tensor()
    pm = parallel.PeerMerge(r)
    host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy()
    res = {}
    # --- fused: accumulate, barrier, rank 0 reads peers inside the tonemap kernel
    with torch.cuda.stream(stream):
        r.accumulate(spp)
        pm.ready()
        if rank == 0:
            img_peer = pm.fetch_image().copy()
        pm.release()
        # --- NCCL: all-reduce the same partial sums, then tonemap
        dist.all_reduce(accum)
        torch.cuda.synchronize()
        img_nccl = r.fetch_image()
    if rank == 0:
        res["max_abs_diff"] = float(np.abs(img_peer - img_nccl).max())
        res["identical"] = bool(np.array_equal(img_peer, img_nccl))
    # --- timing of the two merge paths (per frame batch: accumulate + merge + LDR image on rank 0's host)
    for mode in ("peer", "nccl"):
        times = []
        for it in range(12):
            with torch.cuda.stream(stream):
                r.reset_framebuffer()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                r.accumulate(spp)
                if mode == "peer":
                    pm.ready()
                    if rank == 0:
                        pm.fetch_image(host)
                    pm.release()
                else:
                    dist.all_reduce(accum)
                    if rank == 0:
                        r._check(r._lib.vrt_fetch_ldr(r._h, host.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float))))
                    torch.cuda.synchronize()
                    dist.barrier()
                times.append(time.perf_counter() - t0)
        res["ms_per_batch_" + mode] = 1e3 * float(np.median(times[2:]))
        if rank == 0 and mode == "peer":
            res["merge_kernel_ms"] = r.stats()["last_resolve_ms"]
    if rank == 0:
        res["n_gpus"] = world
        res["spp_per_gpu"] = spp
        print(json.dumps(res))
    pm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
