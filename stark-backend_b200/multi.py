"""Multi-GPU plumbing of the prover: one process per GPU (torch.distributed), independent proofs per
rank ("replicas only"); only 32-byte commitments and timings cross ranks.
reference: OpenVM scales by proving independent segments on separate devices
(benchmarks/synthetic/README.md:23); there is no data-path collective in a single proof yet."""
import numpy as np
import torch
import torch.distributed as dist


def assign_proofs(n_proofs, world, rank):
    """Round-robin assignment of independent proofs (segments / AIR groups) to ranks."""
    return [i for i in range(n_proofs) if i % world == rank]


def all_gather_commitments(root_words, device=None):
    """root_words: uint32[8] commitment of this rank -> list of every rank's commitment, in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.asarray(root_words, dtype=np.uint32)]
    t = torch.from_numpy(np.ascontiguousarray(root_words, dtype=np.uint32).view(np.int32)).clone()
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy().view(np.uint32) for o in out]


def max_over_ranks(value, device=None):
    """Device-timed durations are reported as the maximum over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
