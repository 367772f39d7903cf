#!/usr/bin/env python
"""2-GPU check of the CFG split (run under torchrun --nproc-per-node 2): the split loop (each rank evaluates one CFG
branch, one ncclAllGather of epsilon per step inside the captured graph) must reproduce the single-GPU batched loop.
Prints one JSON line from rank 0; exits non-zero on mismatch."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from minsdtf_b200 import dist as D, synth
    from minsdtf_b200.engine import Engine
    from minsdtf_b200.scheduler import Scheduler, timestep_embedding

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    eng.load_state_dict(synth.make_state_dict("unet"), "unet")
    plan = D.setup_cfg_split(eng, rank, world, device=torch.device("cuda", local))
    B, h, steps = 2, 32, 4
    lat = synth.latents(B, h, h, seed=77 + plan.pair)
    ctx, unc = synth.context(B, 77, seed=78 + plan.pair), synth.uncond_context(B, 77)
    s = Scheduler(active_tcd=False)
    s.set_timesteps(25)
    ts = [int(t) for t in s.timesteps[:steps]]
    coefs = s.coefficients(ts, 7.5, 0.7)
    temb = np.stack([timestep_embedding(t) for t in ts])
    out = {}
    for graph in (False, True):
        t0 = time.perf_counter()
        split = eng.denoise(lat, ctx, unc, temb, coefs, decode=False, use_cuda_graph=graph, cfg_split=True)
        t_split = time.perf_counter() - t0
        whole = eng.denoise(lat, ctx, unc, temb, coefs, decode=False, use_cuda_graph=graph, cfg_split=False)
        err = float(np.abs(split - whole).max())
        # both members of the pair must hold the same latent
        mine = torch.from_numpy(split).cuda()
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)
        pair_err = float((got[plan.rank] - got[plan.partner]).abs().max())
        out["graph" if graph else "eager"] = {"max_abs_vs_unsplit": err, "pair_replica_diff": pair_err, "split_s": round(t_split, 3)}
    ok = all(v["max_abs_vs_unsplit"] <= 1e-5 and v["pair_replica_diff"] == 0.0 for v in out.values())
    flags = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flags)
    if rank == 0:
        print(json.dumps({"cfg_split_check": out, "world": world, "ok": flags.item() == 0}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()
    sys.exit(0 if flags.item() == 0 else 1)


if __name__ == "__main__":
    main()
