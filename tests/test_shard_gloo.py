"""N > 1 host logic on CPU: tuners are dealt to ranks, every receiver follows its tuner, no rank
needs another rank's data, and the job time is the MAX over ranks (torch.distributed, gloo,
world_size 2).  The per-rank "compute" here is the CPU oracle -- test infrastructure standing in
for the GPU bank, which the -m gpu tests cover; what is under test is the partition and the
reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from webradio_b200 import shard, synth


def test_assign_partitions_every_receiver_once():
    tuner_of = [0, 0, 1, 2, 1, 0, 3, 3, 2, 4]
    for world in (1, 2, 3, 8):
        shards = shard.assign(tuner_of, world)
        seen = sorted(r for s in shards for r in s.receivers)
        assert seen == list(range(len(tuner_of)))
        for s in shards:
            for r, ls in zip(s.receivers, s.local_stream):
                assert s.tuners[ls] == tuner_of[r]          # the receiver sits next to its tuner
        owners = {}
        for s in shards:
            for t in s.tuners:
                assert t not in owners                       # a tuner block goes to exactly one GPU
                owners[t] = s.rank
        sizes = [len(s.tuners) for s in shards]
        assert max(sizes) - min(sizes) <= 1                  # round-robin balance


def test_weak_scaling_layout():
    s = shard.weak_scaling_shard(16, 1024, rank=3, world_size=8)
    assert s.tuners == list(range(48, 64)) and len(s.receivers) == 1024
    assert s.receivers[0] == 3 * 1024 and s.local_stream[:65] == [0] * 64 + [1]
    with pytest.raises(ValueError):
        shard.weak_scaling_shard(1, 64, rank=2, world_size=2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tuner_of, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wro
        fs, F, n1, d1, n2, d2 = 2400000, 4000, 64, 10, 64, 5
        t1 = wro.lowpass_design(n1, 80000, fs)
        t2 = wro.lowpass_design(n2, 8000, fs // d1)
        ifs = synth.receiver_ifs(len(tuner_of), fs)
        mine = shard.assign(tuner_of, world)[rank]
        # each rank only ever touches the streams of ITS tuners
        blocks = {t: [synth.lattice_noise(F, stream=t, start=b * F) for b in range(2)] for t in mine.tuners}
        audio = {}
        for r, ls in zip(mine.receivers, mine.local_stream):
            rx = wro.Rx(fs, int(ifs[r]), t1, d1, r % 4 if r % 4 != 1 else 0, t2, d2)
            audio[r] = np.concatenate([rx.process(x) for x in blocks[mine.tuners[ls]]])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{str(r): a for r, a in audio.items()})
        # timing reduction: the job takes as long as its slowest rank
        ms = shard.reduce_max_ms(10.0 * (rank + 1))
        assert ms == 10.0 * world
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path, wro):
    tuner_of = [0, 0, 1, 1, 2, 2, 2, 0]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), tuner_of, str(tmp_path)), nprocs=world, join=True)
    got = {}
    for k in range(world):
        z = np.load(tmp_path / f"rank{k}.npz")
        for r in z.files:
            assert int(r) not in got
            got[int(r)] = z[r]
    assert sorted(got) == list(range(len(tuner_of)))
    # unsharded reference run
    fs, F, n1, d1, n2, d2 = 2400000, 4000, 64, 10, 64, 5
    t1 = wro.lowpass_design(n1, 80000, fs)
    t2 = wro.lowpass_design(n2, 8000, fs // d1)
    ifs = synth.receiver_ifs(len(tuner_of), fs)
    for r, t in enumerate(tuner_of):
        rx = wro.Rx(fs, int(ifs[r]), t1, d1, r % 4 if r % 4 != 1 else 0, t2, d2)
        want = np.concatenate([rx.process(synth.lattice_noise(F, stream=t, start=b * F)) for b in range(2)])
        assert np.array_equal(got[r].view(np.uint32), want.view(np.uint32)), f"receiver {r}"


def test_job_throughput():
    assert shard.job_throughput(1e6, 8, 2.0) == pytest.approx(4e9)
