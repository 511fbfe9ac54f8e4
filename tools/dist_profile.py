"""Kernel timeline of one distributed training step (run under torchrun, >= 2 GPUs): prints, for rank 0, every
device kernel / memcpy of one step with its start offset, duration and stream, plus the idle gaps on the main
stream -- the tool used to see which collectives and copies are exposed.  Not a benchmark (profiler attached)."""
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import gscodec_studio_b200 as gs  # noqa: E402
from gscodec_studio_b200.compression_simulation import CompressionSimulation  # noqa: E402
from gscodec_studio_b200.losses import photometric_loss  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    raw, viewmat, K, target, _ = bench.build_workload(rank, world, 1 if "SMALL" in os.environ else 3)
    params = {k: v.to(dev).requires_grad_(True) for k, v in raw.items()}
    viewmat, K, target = viewmat.to(dev), K.to(dev), target.to(dev)
    sim = CompressionSimulation(entropy_model_enable=False, device=dev)

    def step():
        for p in params.values():
            p.grad = None
        new, _ = sim.simulate_compression(params, 100, fuse_activations=True)
        colors = (new["sh0"], new["shN"])
        rc, _, _ = gs.rasterization(new["means"], new["quats"], new["scales_act"], new["opacities_act"], colors, viewmat, K,
                                    bench.W_IMG, bench.H_IMG, sh_degree=3, packed=False, near_plane=0.01, far_plane=100.0,
                                    radius_clip=3.0, distributed=world > 1)
        photometric_loss(rc, target, ssim_lambda=0.0).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        evs.sort(key=lambda e: e.time_range.start)
        # keep the middle step: split at the first fake-quant kernel of each step
        starts = [i for i, e in enumerate(evs) if "fake_quantize" in e.name and "fwd" in e.name]
        per_step = len(starts) // 3
        lo, hi = starts[per_step], starts[2 * per_step]
        t0 = evs[lo].time_range.start
        prev_end = t0
        print(f"step span {(evs[hi].time_range.start - t0) / 1e3:.3f} ms, {hi - lo} device activities")
        for e in evs[lo:hi]:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            gap = e.time_range.start - prev_end
            print(f"{s / 1e3:8.3f} ms  {d:7.1f} us  gap {gap:7.1f}  {e.name[:90]}")
            prev_end = max(prev_end, e.time_range.end)
        # host side of the same step: top-level CPU ops / python functions, start relative to the step's first kernel
        t1 = evs[hi].time_range.start
        cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.cpu_parent is None]
        cpu.sort(key=lambda e: e.time_range.start)
        print("---- host timeline (top-level ops; offsets relative to the first kernel of the step)")
        for e in cpu:
            s0 = e.time_range.start - t0
            if -3000 <= s0 <= (t1 - t0):
                d = e.time_range.end - e.time_range.start
                if d >= 15:
                    print(f"{s0 / 1e3:8.3f} ms  {d:7.1f} us  {e.name[:100]}")
    # host profile of 20 further steps (rank 0): which Python functions own the host time of a distributed step
    import cProfile
    import pstats

    torch.cuda.synchronize()
    dist.barrier()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        step()
    pr.disable()
    torch.cuda.synchronize()
    if rank == 0:
        print("---- cProfile, 20 steps, by own time")
        pstats.Stats(pr).sort_stats("tottime").print_stats(30)
        print("---- cProfile, 20 steps, cumulative")
        pstats.Stats(pr).sort_stats("cumulative").print_stats(50)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
