"""World-size-2 gloo test of the multi-GPU host logic (no GPU): sharding plan, per-rank synthesis of a shard with a
stand-in synthesiser, gather in input order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_synth(ids_batch):
    # deterministic stand-in: "waveform" depends only on the ids, like the real graph with zero noise scales
    return [np.cumsum(np.asarray(ids, np.float32)) * 0.5 for ids in ids_batch]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from piper_b200 import dist as pdist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    ids_list = [rng.integers(0, 100, n).tolist() for n in (381, 113, 193, 197, 151, 165, 307, 5)]
    out = pdist.synthesize_sharded(_fake_synth, ids_list, rank, world, gather_to=0)
    if rank == 0:
        ok = len(out) == len(ids_list) and all(np.array_equal(o, e) for o, e in zip(out, _fake_synth(ids_list)))
        q.put(bool(ok))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_synthesis_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_single_process_path_keeps_order():
    from piper_b200 import dist as pdist
    ids_list = [[3, 1, 2], [9] * 10, [4, 4]]
    out = pdist.synthesize_sharded(_fake_synth, ids_list, 0, 1)
    assert all(np.array_equal(o, e) for o, e in zip(out, _fake_synth(ids_list)))
