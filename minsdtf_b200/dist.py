"""Host-side rank logic for multi-GPU runs (one process per GPU, launched by torchrun).

The reference is single-device; two partitionings fall out of its loop (§8e):

* data parallel over prompts — images never interact (even `rescale_noise_cfg` reduces per sample,
  stable_diffusion.py:309-310): `shard(n, world, rank)` gives each rank its slice, no data-path collective;
* 2-way CFG split — the uncond / cond evaluations of one step (:454-457) are independent given the latent: GPUs are
  grouped in pairs, the even member evaluates the unconditional branch, the odd one the conditional branch, and the
  engine exchanges the two epsilons with one ncclAllGather per step (`sdtf_comm_init`, csrc/comm.cuh).  Pairs are
  data parallel among themselves.

torch.distributed is only the plumbing that ships the 128-byte NCCL unique id from the even to the odd member of a pair
(gloo on CPU in the tests, nccl on the GPU box); the per-step exchange itself never goes through Python.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np


def shard(n_items: int, parts: int, index: int) -> slice:
    """Contiguous, balanced slice `index` of `parts` over range(n_items) (earlier parts get the remainder)."""
    if not (0 <= index < parts):
        raise ValueError(f"index {index} outside [0, {parts})")
    base, rem = divmod(n_items, parts)
    start = index * base + min(index, rem)
    return slice(start, start + base + (1 if index < rem else 0))


@dataclass(frozen=True)
class CfgSplitPlan:
    """Where rank `rank` of `world` sits when GPUs are paired for the CFG split."""
    rank: int
    world: int
    pair: int          # data-parallel index of the pair
    n_pairs: int
    branch: int        # 0: unconditional branch, 1: conditional branch (= NCCL rank inside the pair)
    partner: int       # global rank of the other member

    @property
    def leader(self) -> int:  # global rank of the pair's even member (generates the NCCL id)
        return self.pair * 2


def cfg_split_plan(rank: int, world: int) -> CfgSplitPlan:
    if world < 2 or world % 2:
        raise ValueError(f"the CFG split pairs GPUs: world size must be even and >= 2, got {world}")
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    return CfgSplitPlan(rank=rank, world=world, pair=rank // 2, n_pairs=world // 2, branch=rank % 2, partner=rank ^ 1)


def exchange_pair_id(plan: CfgSplitPlan, make_id, device=None) -> bytes:
    """Ship a 128-byte id from each pair's leader to its partner over torch.distributed (default group must be
    initialised; works with gloo and nccl).  `make_id()` is called on leaders only.  Every rank takes part in every
    pair's broadcast-group creation, as `new_group` requires."""
    import torch
    import torch.distributed as dist

    groups = [dist.new_group([2 * p, 2 * p + 1]) for p in range(plan.n_pairs)]
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if plan.rank == plan.leader:
        raw = bytes(make_id())
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=plan.leader, group=groups[plan.pair])
    return bytes(buf.cpu().numpy().tobytes())


def nccl_library_path():
    """The torch-bundled libnccl.so.2 if present (what torch.distributed itself uses), else None (engine falls back to
    $SDTF_NCCL_LIB / the default soname)."""
    try:
        import os
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def setup_cfg_split(engine, rank: int, world: int, device=None) -> CfgSplitPlan:
    """Pair up, exchange the NCCL id, and create the engine's 2-rank communicator.  Returns the plan."""
    plan = cfg_split_plan(rank, world)
    lib = nccl_library_path()
    uid = exchange_pair_id(plan, lambda: engine.comm_unique_id(lib), device=device)
    engine.comm_init(uid, plan.branch, 2, lib)
    return plan


def split_step_reference(eps_branch: np.ndarray, plan: CfgSplitPlan, all_gather) -> tuple:
    """What the engine's split step does with this rank's epsilon, as plain NumPy (used by the CPU tests with a gloo
    all-gather standing in for ncclAllGather): returns (eps_uncond, eps_cond) as every member of the pair sees them."""
    parts = all_gather(np.ascontiguousarray(eps_branch, np.float32))
    return parts[0], parts[1]
