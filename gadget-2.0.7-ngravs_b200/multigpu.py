"""Multi-GPU plumbing of the tree-force step (one process per GPU, torch.distributed): every rank owns 1/N of the particle
records, one all-gather per step replicates them (NCCL over NVLink on GPUs, gloo in the CPU tests), every rank builds the
same tree and walks an equal slice of the tree-ordered active targets.  No reduction is needed: a rank
writes only the accelerations of its own targets."""
import torch
import torch.distributed as dist


def slice_bounds(n, rank, nranks):
    """[lo, hi) of `n` items for `rank`: the same arithmetic as g2_stage_walk (csrc/g2_walk.cu): n*rank/nranks."""
    return (n * rank) // nranks, (n * (rank + 1)) // nranks


def owner_slice(n, rank, nranks):
    """Particle records owned (uploaded) by `rank`: equal chunks of ceil(n/nranks), the last one short."""
    per = (n + nranks - 1) // nranks
    return min(n, rank * per), min(n, (rank + 1) * per), per


def pack_records(pos, mass, ptype, oldacc=None, active=None):
    """(n, 8) float32 tensor of 32-byte g2gpu_particle records: x, y, z, m, type (int32 bits), OldAcc, active (int32 bits), pad."""
    n = pos.shape[0]
    rec = torch.zeros((n, 8), dtype=torch.float32, device=pos.device)
    rec[:, 0:3] = pos
    rec[:, 3] = mass
    irec = rec.view(torch.int32)
    irec[:, 4] = ptype.to(torch.int32)
    if oldacc is not None:
        rec[:, 5] = oldacc
    irec[:, 6] = 1 if active is None else active.to(torch.int32)
    return rec


class ParticleExchange:
    """Persistent buffers for the per-step all-gather of the particle records (ONE collective per step)."""

    def __init__(self, n, device, world):
        self.n, self.world = n, world
        self.per = (n + world - 1) // world
        npad = self.per * world
        self.g_rec = torch.zeros((npad, 8), dtype=torch.float32, device=device)
        self.s_rec = torch.zeros((self.per, 8), dtype=torch.float32, device=device)

    def set_local(self, rec):
        self.s_rec[: rec.shape[0]].copy_(rec)

    def gather(self):
        """returns a view of the first n gathered records"""
        if self.world > 1:
            dist.all_gather_into_tensor(self.g_rec, self.s_rec)
        else:
            self.g_rec.copy_(self.s_rec)
        return self.g_rec[: self.n]

    def bytes_per_step(self):
        return self.per * self.world * 32
