#!/usr/bin/env python
"""Soak of the in-kernel sum exchange (ticket + tagged mailbox words) under torchrun: a few hundred thousand attempts of the
device-resident loop and of the host-driven fused loop on small shards (latency-dominated: the exchange is most of an attempt), then
every rank must report the same number of attempts, the same t and dt to the last bit, and no time-out."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist

    import numericalnim_b200 as nn
    from numericalnim_b200 import distributed as D

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = D.init_context(local)
    n = (1 << 17) * world
    off, ln = D.shard_range(n, rank, world)
    i = np.arange(off, off + ln, dtype=np.float64)
    lam = nn.GpuVector.from_local(n, 0.1 + 9.9 * i / (n - 1), ctx)
    y0 = nn.GpuVector.from_local(n, 1.0 + 0.5 * np.sin(2 * np.pi * i / n), ctx)
    opts = nn.newODEoptions(absTol=1e-9, relTol=1e-9, dtMax=1e-3, dtMin=1e-8)   # tight tolerance + small dtMax: many steps, some rejections
    ok = True
    for name, devloop, steps in (("device loop", 1, 300000), ("host-driven fused loop", 0, 30000)):
        ctx.set("device_loop", devloop)
        s = nn.Solver("dopri54", nn.rhsDiagLinear(lam), y0, 1e12, opts)
        t0 = time.time()
        done, _ = s.advance(steps)
        ctx.synchronize()
        wall = time.time() - t0
        t, dt_next, err, _ = s.state()
        st = s.stats()
        mine = torch.tensor([float(done), float(st["attempts"]), t, dt_next, err], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        same = all(torch.equal(a.view(torch.int64), allr[0].view(torch.int64)) for a in allr)
        ok = ok and same and done == steps
        if rank == 0:
            print(f"[soak world={world}] {name}: steps={done} attempts={st['attempts']} rejected={st['rejected']} collectives={st['collectives']} "
                  f"t={t:.6f} {1e6 * wall / max(1, st['attempts']):.2f} us/attempt identical_on_all_ranks={same}", flush=True)
        s.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
