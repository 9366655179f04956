"""CPU tests of the N > 1 host logic: the row partition and the handle exchange, with a world_size-2 gloo process group."""
import os
import socket

import pytest


def test_row_partition_covers_all_rows():
    from superfluid_dynamics_b200 import api
    for N in (2048, 4096, 5000, 16384, 65536, 65536 + 300):
        for G in (1, 2, 4, 8):
            ranges = [api.comm_row_range(N, r, G) for r in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == N
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            for lo, hi in ranges:
                assert lo % 256 == 0 or lo == N                  # ranges start on a 256-row cell
                assert hi % 256 == 0 or hi == N
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) == sizes[0] and max(sizes) - min(sizes[:-1] or sizes) <= 256 * ((N // 256) // G + 1)
    with pytest.raises(ValueError):
        api.comm_row_range(4096, 3, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from superfluid_dynamics_b200 import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = bytes([rank]) * 64
    got = api.exchange_handles(blob)
    lo, hi = api.comm_row_range(4096, rank, world)
    q.put((rank, [g[0] for g in got], [len(g) for g in got], lo, hi))
    dist.destroy_process_group()


def test_handle_exchange_world_size_2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1] and res[1][1] == [0, 1]          # every rank sees every handle, in rank order
    assert res[0][2] == [64, 64]
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 2048, 2048, 4096)


def test_ensemble_partition_and_packing():
    import numpy as np
    from superfluid_dynamics_b200 import api
    for B in (0, 1, 7, 1024, 1025):
        for G in (1, 2, 3, 8):
            ranges = [api.ensemble_member_range(B, r, G) for r in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            sizes = [hi - lo for lo, hi in ranges]
            assert all(a1 == b0 for (_, a1), (b0, _) in zip(ranges, ranges[1:]))
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        api.ensemble_member_range(8, 2, 2)
    N = 4
    members = [np.arange(2 * N) + 100 * m + 0j for m in range(3)]
    packed = api.ensemble_state(members, N)
    assert packed.shape == (3 * 2 * N,)
    for m in range(3):
        assert np.array_equal(packed[m * N:(m + 1) * N], members[m][:N])
        assert np.array_equal(packed[3 * N + m * N:3 * N + (m + 1) * N], members[m][N:])
