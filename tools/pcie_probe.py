"""Host<->device copy rates of the e2e path's traffic pattern (80 MB H2D + 40 MB D2H per tick and rank, pinned buffers), alone
and with all ranks copying at once.  Names the limiter of the end-to-end scaling:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
Rank r pins itself to its slice of the cores like bench.py does, uses cuda:r, and prints one JSON line on rank 0."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist


def rate(fn, reps=10):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    try:
        from bench import pin_rank_to_cores
        cores = pin_rank_to_cores(local, world)
    except Exception:
        cores = []
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    x = torch.empty(80 * 2**20 // 8, dtype=torch.float64, pin_memory=True).fill_(1.0)
    y = torch.empty_like(x, device="cuda")
    z = torch.empty(40 * 2**20 // 8, dtype=torch.float64, pin_memory=True).fill_(1.0)
    w = torch.empty_like(z, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d(): y.copy_(x, non_blocking=True)
    def d2h(): z.copy_(w, non_blocking=True)
    def both():
        with torch.cuda.stream(s1): y.copy_(x, non_blocking=True)
        with torch.cuda.stream(s2): z.copy_(w, non_blocking=True)

    for f in (h2d, d2h, both):
        rate(f, 3)
    out = {}
    for mode in ("alone", "all_ranks"):
        res = {}
        for name, f, mb in (("h2d_80MB", h2d, 80), ("d2h_40MB", d2h, 40), ("both", both, 120)):
            if world > 1:
                dist.barrier()
            if mode == "alone" and rank != 0:
                dt = float("nan")
            else:
                dt = rate(f)
            if world > 1:
                dist.barrier()
            t = torch.tensor([dt if dt == dt else 0.0], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = {"ms": round(float(t.item()) * 1e3, 3), "GB_per_s_per_rank": round(mb * 2**20 / 1e9 / float(t.item()), 1)}
        out[mode] = res
    if rank == 0:
        print(json.dumps({"world": world, "cores_of_rank0": len(cores), "copies": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
