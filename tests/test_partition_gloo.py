"""CPU, world_size 2 over gloo: the multi-GPU plumbing of bench.py / the engine -- utterances are
partitioned across ranks with no data-path collective; the only communication is the timing
barrier and the max-over-ranks reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fullycnnspeechenhancement_b200.engine import num_frames, partition_utterances


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                      # same manifest on every rank
    lengths = rng.integers(16000, 64001, 257)
    mine = partition_utterances(lengths, world)[rank]
    frames = int(num_frames(lengths[mine]).sum())
    # "work": pretend time proportional to frames, report max over ranks like bench.py does
    t = torch.tensor([frames * 1e-6], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.tolist(), frames))
    if rank == 0:
        q.put((gathered, float(t.item()), lengths.tolist()))
    dist.destroy_process_group()


def test_two_rank_partition_is_complete_and_balanced():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax, lengths = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = sorted(gathered[0][0] + gathered[1][0])
    assert idx == list(range(len(lengths)))                       # every utterance exactly once
    loads = [gathered[0][1], gathered[1][1]]
    assert abs(loads[0] - loads[1]) <= int(num_frames(np.array(lengths)).max())
    assert abs(tmax - max(loads) * 1e-6) < 1e-12                  # max over ranks
