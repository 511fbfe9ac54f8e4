"""Times the collectives the distributed render path uses (run under torchrun)."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 1006065
n_even = N // world
counts = [len(range(r, N, world)) for r in range(world)]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n


for width, name in ((48, "shN+sh0 [n,16,3]"), (11, "geometry 11 floats")):
    x_even = torch.randn(n_even, width, device=dev)
    out_even = torch.empty(world * n_even, width, device=dev)
    t_ag = timeit(lambda: dist.all_gather_into_tensor(out_even, x_even))
    g_even = torch.randn(world * n_even, width, device=dev)
    r_even = torch.empty(n_even, width, device=dev)
    t_rs = timeit(lambda: dist.reduce_scatter_tensor(r_even, g_even))
    x_un = torch.randn(counts[rank], width, device=dev)
    out_un = torch.empty(N, width, device=dev)
    views, s = [], 0
    for c in counts:
        views.append(out_un[s:s + c]); s += c
    t_agu = timeit(lambda: dist.all_gather(views, x_un))
    g_un = torch.randn(N, width, device=dev)
    gviews, s = [], 0
    for c in counts:
        gviews.append(g_un[s:s + c]); s += c
    r_un = torch.empty(counts[rank], width, device=dev)
    t_rsu = timeit(lambda: dist.reduce_scatter(r_un, gviews))
    t_ar = timeit(lambda: dist.all_reduce(g_un))
    if rank == 0:
        mb = N * width * 4 / 1e6
        print(f"{name}: {mb:.0f} MB total | all_gather_into_tensor {t_ag:.3f} ms | reduce_scatter_tensor {t_rs:.3f} ms | "
              f"uneven all_gather(list) {t_agu:.3f} ms | uneven reduce_scatter(list) {t_rsu:.3f} ms | all_reduce {t_ar:.3f} ms")
dist.destroy_process_group()
