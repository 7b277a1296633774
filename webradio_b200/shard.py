"""Sharding of tuners / receivers across GPUs (one process per GPU).

Receivers are independent -- each reference `Receiver` owns a private block chain and private
state (reference src/radio.cxx:62-90); the only shared datum is the read-only tuner block.  So the
batch shards by ASSIGNMENT: whole tuners (front-ends) go to ranks round-robin and every receiver
follows its tuner, which puts each IQ block on exactly one GPU and needs no collective on the data
path.  torch.distributed is used only off the path: a barrier around the timed
region and a MAX-reduce of the per-rank device time.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass
class Shard:
    rank: int
    tuners: List[int] = field(default_factory=list)        # global tuner ids on this rank
    receivers: List[int] = field(default_factory=list)     # global receiver ids on this rank
    local_stream: List[int] = field(default_factory=list)  # per local receiver: index into `tuners`


def assign(receiver_tuner: List[int], world_size: int) -> List[Shard]:
    """receiver_tuner[r] = tuner id receiver r listens to.  Returns one Shard per rank.

    Tuners are dealt round-robin in order of first appearance; a rank's receivers keep their
    global order.  Every receiver lands on exactly one rank, next to its tuner's stream.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = []
    for t in receiver_tuner:
        if t not in order:
            order.append(t)
    owner = {t: i % world_size for i, t in enumerate(order)}
    shards = [Shard(rank=k) for k in range(world_size)]
    for t in order:
        shards[owner[t]].tuners.append(t)
    for r, t in enumerate(receiver_tuner):
        s = shards[owner[t]]
        s.receivers.append(r)
        s.local_stream.append(s.tuners.index(t))
    return shards


def weak_scaling_shard(n_streams: int, n_receivers: int, rank: int, world_size: int) -> Shard:
    """bench.py's layout (weak scaling): every rank runs its own copy of the workload -- n_streams
    tuners with n_receivers / n_streams receivers each -- so rank k owns the contiguous global
    tuner ids k*n_streams ... (k+1)*n_streams - 1 and the receivers that follow them."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    per = n_receivers // n_streams
    return Shard(rank,
                 list(range(rank * n_streams, (rank + 1) * n_streams)),
                 list(range(rank * n_receivers, (rank + 1) * n_receivers)),
                 [r // per for r in range(n_receivers)])


def reduce_max_ms(ms: float, device=None) -> float:
    """MAX over ranks of a per-rank elapsed time (the job finishes when the slowest rank does)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank: float, world_size: int, max_ms: float) -> float:
    """Whole-job units per second: what all ranks processed divided by the slowest rank's time."""
    return units_per_rank * world_size / (max_ms * 1e-3)
