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


class ParticleExchange:
    """Persistent buffers for the per-step all-gather of (x,y,z,m), type and OldAcc."""

    def __init__(self, n, device, world):
        self.n, self.world = n, world
        self.per = (n + world - 1) // world
        npad = self.per * world
        self.g_pm = torch.zeros((npad, 4), dtype=torch.float32, device=device)
        self.g_type = torch.zeros(npad, dtype=torch.int32, device=device)
        self.g_old = torch.zeros(npad, dtype=torch.float32, device=device)
        self.s_pm = torch.zeros((self.per, 4), dtype=torch.float32, device=device)
        self.s_type = torch.zeros(self.per, dtype=torch.int32, device=device)
        self.s_old = torch.zeros(self.per, dtype=torch.float32, device=device)

    def set_local(self, pm, ptype, oldacc):
        k = pm.shape[0]
        self.s_pm[:k].copy_(pm)
        self.s_type[:k].copy_(ptype)
        self.s_old[:k].copy_(oldacc)

    def gather(self):
        """one collective per array; returns views of the first n gathered records"""
        if self.world > 1:
            dist.all_gather_into_tensor(self.g_pm, self.s_pm)
            dist.all_gather_into_tensor(self.g_type, self.s_type)
            dist.all_gather_into_tensor(self.g_old, self.s_old)
        else:
            self.g_pm.copy_(self.s_pm)
            self.g_type.copy_(self.s_type)
            self.g_old.copy_(self.s_old)
        return self.g_pm[: self.n], self.g_type[: self.n], self.g_old[: self.n]

    def bytes_per_step(self):
        return self.per * self.world * (16 + 4 + 4)
