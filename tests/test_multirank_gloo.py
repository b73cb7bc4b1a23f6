"""N>1 host logic on CPU: world_size-2 gloo processes partition 11 sessions, run the ORACLE on their own sessions (the
checker stands in for the device path, which needs a GPU) and gather the statistics.  Checks: the partition is a
disjoint cover, per-session results do not depend on which rank owned the session (sessions are independent -> no
collective on the data path), and the gathered throughput is total audio over the slowest rank's time."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from freeze_omni_b200 import sharding
from freeze_omni_b200.config import load_path_config
from freeze_omni_b200.weights import make_adapter_state, make_encoder_state
from oracle import freeze_omni_oracle as O

N_SESSIONS, N_CHUNKS = 11, 3


def _session_outputs(cfg, esd, asd, sid):
    g = torch.Generator().manual_seed(100 + sid)
    s = O.StreamSession(cfg, esd, asd)
    outs = []
    for _ in range(N_CHUNKS):
        pcm = 0.05 * torch.randn(cfg.samples_per_chunk, generator=g) * 32768
        _, enc, y = s.step_pcm(pcm, 1.0)
        outs.append(float(y.double().sum()))
    return outs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = load_path_config("tiny")
    esd, asd = make_encoder_state(cfg, 3), make_adapter_state(cfg, 3)
    mine = sharding.partition(range(N_SESSIONS), world, rank)
    res = {sid: _session_outputs(cfg, esd, asd, sid) for sid in mine}
    local = {"audio_seconds": len(mine) * N_CHUNKS * 0.16, "session_chunks": len(mine) * N_CHUNKS,
             "max_elapsed_s": 1.0 + rank}
    tot = sharding.gather_stats(local)
    q.put((rank, mine, res, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_stats():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(sum((g[1] for g in got), []))
    assert owned == list(range(N_SESSIONS))                              # disjoint cover
    cfg = load_path_config("tiny")
    esd, asd = make_encoder_state(cfg, 3), make_adapter_state(cfg, 3)
    for _, mine, res, tot in got:
        for sid in mine[:2]:                                              # independent of the owner rank
            assert np.allclose(res[sid], _session_outputs(cfg, esd, asd, sid), rtol=0, atol=1e-4)
        assert tot["session_chunks"] == N_SESSIONS * N_CHUNKS
        assert abs(tot["audio_seconds"] - N_SESSIONS * N_CHUNKS * 0.16) < 1e-9
        assert tot["max_elapsed_s"] == 2.0                                # slowest rank
        assert abs(sharding.throughput(tot) - N_SESSIONS * N_CHUNKS * 0.16 / 2.0) < 1e-9


def test_partition_properties():
    for world in (1, 2, 4, 8):
        ids = list(range(1024))
        parts = [sharding.partition(ids, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == ids
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        own = sharding.owner(ids, world)
        assert all(own[s] == r for r, p in enumerate(parts) for s in p)
    assert sharding.assign_least_loaded([3, 1, 2]) == 1
