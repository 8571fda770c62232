"""Chain sharding for the multi-GPU SAPG run.

One process per GPU.  `total_chains` independent MYULA chains are split into
contiguous blocks, rank r owning chains [r*n_local, (r+1)*n_local).  At every
outer iteration each rank contributes the per-chain sums of its chains; all
ranks receive all of them IN GLOBAL CHAIN ORDER and reduce them in that fixed
order, so the theta / sigma^2 / PSF-parameter updates are bit-identical on every
rank and independent of the number of GPUs.  This generalises the reference's
size-1 mini-batch average `G_b = mean(g_b)` (SAPG_algorithm_moffat.m:170-173).

On the GPU the exchange is an ncclAllGather enqueued on the compute stream by
libsbd (`sbd_comm_init`); this module holds the host-side logic - the
partition, the NCCL unique-id rendezvous through torch.distributed, and the
same gather-in-chain-order combine for CPU (gloo) tests.
"""
import ctypes as C

import numpy as np

from ._lib import lib, SbdError, SBD_NCCL_ID_BYTES


class ChainShard:
    def __init__(self, total_chains, world_size=1, rank=0):
        if total_chains % world_size != 0:
            raise ValueError("total_chains must be a multiple of world_size (equal shards)")
        self.total_chains, self.world_size, self.rank = int(total_chains), int(world_size), int(rank)
        self.n_local = self.total_chains // self.world_size
        self.chain_offset = self.rank * self.n_local

    def local_chains(self):
        return range(self.chain_offset, self.chain_offset + self.n_local)

    def owner(self, chain):
        return chain // self.n_local

    # -- rendezvous ---------------------------------------------------------
    def init_engine_comm(self, engine):
        """Create the NCCL communicator inside libsbd for this rank.  The unique
        id is generated on rank 0 and broadcast with torch.distributed."""
        if self.world_size == 1:
            return
        import torch
        import torch.distributed as dist
        buf = C.create_string_buffer(SBD_NCCL_ID_BYTES)
        if self.rank == 0:
            rc = lib.sbd_comm_unique_id(buf)
            if rc != 0:
                raise SbdError(rc, lib.sbd_last_error(None).decode())
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        rc = lib.sbd_comm_init(engine._h, self.world_size, self.rank, raw)
        if rc != 0:
            raise SbdError(rc, lib.sbd_last_error(engine._h).decode())

    # -- host-side combine (same semantics as the device all-gather) -------
    def combine(self, local_vectors):
        """local_vectors: list (length n_local) of equal-length 1-D arrays, one
        per local chain.  Returns the sum over ALL chains, accumulated in global
        chain order."""
        local = np.stack([np.asarray(v, dtype=np.float64) for v in local_vectors], 0)
        if self.world_size == 1:
            allv = local
        else:
            import torch
            import torch.distributed as dist
            mine = torch.from_numpy(np.ascontiguousarray(local))
            parts = [torch.empty_like(mine) for _ in range(self.world_size)]
            dist.all_gather(parts, mine)
            allv = torch.cat(parts, 0).numpy()
        tot = np.zeros(allv.shape[1])
        for ch in range(allv.shape[0]):          # fixed order
            tot = tot + allv[ch]
        return tot
