import os, sys, json
ROOT=os.getcwd(); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"tests"))
import torch, torch.distributed as dist
import voxel_rt2_b200 as vrt
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for sky in (3840, 1024):
    o = bench.leg_config4_rows(vrt, torch, dist, rank, world, local, 1920, 1080, sky, frames=int(os.environ.get("FRAMES","24")))
    if rank == 0: print("LEG sky", sky, "no_clocks", os.environ.get("VRT_BENCH_NO_CLOCKS"), "ms_per_frame", o["ms_per_frame"], o["clocks"])
dist.barrier(); dist.destroy_process_group()
