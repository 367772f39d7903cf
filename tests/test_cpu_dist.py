"""CPU tests of the multi-GPU host logic (world size 2 and 4 over gloo): prompt sharding, CFG-pair planning, the
NCCL-id exchange plumbing, and that the split step (each rank computes one CFG branch, all-gather, redundant update)
reproduces the single-process CFG combine + scheduler step of the oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minsdtf_b200 import dist as D
from oracle.scheduler_oracle import OracleScheduler, cfg_combine


def test_shard_is_a_balanced_partition():
    for n in (0, 1, 7, 8, 9, 64):
        for parts in (1, 2, 3, 8):
            got = []
            for i in range(parts):
                s = D.shard(n, parts, i)
                got.extend(range(n)[s])
                assert 0 <= (s.stop - s.start) - n // parts <= 1
            assert got == list(range(n))
    with pytest.raises(ValueError):
        D.shard(8, 2, 2)


def test_cfg_split_plan_pairs_and_errors():
    p = [D.cfg_split_plan(r, 8) for r in range(8)]
    assert [q.pair for q in p] == [0, 0, 1, 1, 2, 2, 3, 3]
    assert [q.branch for q in p] == [0, 1] * 4
    assert all(p[q.partner].partner == q.rank and p[q.partner].pair == q.pair for q in p)
    assert all(q.leader == 2 * q.pair and q.n_pairs == 4 for q in p)
    for bad in ((0, 1), (0, 3), (4, 4)):
        with pytest.raises(ValueError):
            D.cfg_split_plan(*bad)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = D.cfg_split_plan(rank, world)
        # 1. id exchange: every pair ends up with its leader's 128 bytes, distinct between pairs
        uid = D.exchange_pair_id(plan, lambda: bytes([plan.pair + 1]) * 128)
        assert uid == bytes([plan.pair + 1]) * 128
        # 2. prompts are sharded over pairs; both members of a pair see the same shard
        prompts = np.arange(8)
        mine = prompts[D.shard(len(prompts), plan.n_pairs, plan.pair)]
        # 3. split step == unsplit step.  "UNet" = a deterministic function of (latent, branch)
        rng = np.random.default_rng(1234 + plan.pair)
        latent = rng.standard_normal((len(mine), 8, 8, 4)).astype(np.float32)
        eps_u = np.sin(latent * 1.7 + 0.3).astype(np.float32)
        eps_c = np.cos(latent * 0.9 - 0.2).astype(np.float32)
        pair_group = None
        for p in range(plan.n_pairs):  # new_group must be called by all ranks for every group
            g = dist.new_group([2 * p, 2 * p + 1])
            if p == plan.pair:
                pair_group = g

        def all_gather(x):
            t = torch.from_numpy(x)
            outs = [torch.empty_like(t) for _ in range(2)]
            dist.all_gather(outs, t, group=pair_group)
            return [o.numpy() for o in outs]

        got_u, got_c = D.split_step_reference(eps_c if plan.branch else eps_u, plan, all_gather)
        assert np.array_equal(got_u, eps_u) and np.array_equal(got_c, eps_c)
        sch = OracleScheduler(False)
        sch.set_timesteps(25)
        t = int(sch.timesteps[3])
        sch._step_index = None
        want = sch.step(cfg_combine(eps_u, eps_c, 7.5, 0.7), t, latent)
        sch2 = OracleScheduler(False)
        sch2.set_timesteps(25)
        got = sch2.step(cfg_combine(got_u, got_c, 7.5, 0.7), t, latent)
        assert np.array_equal(np.asarray(want), np.asarray(got))
        np.save(os.path.join(out_dir, f"lat_{rank}.npy"), np.asarray(got))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cfg_split_over_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for p in range(world // 2):  # both members of a pair hold the same updated latent (replicated state)
        a, b = np.load(tmp_path / f"lat_{2 * p}.npy"), np.load(tmp_path / f"lat_{2 * p + 1}.npy")
        assert np.array_equal(a, b)
