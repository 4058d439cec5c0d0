"""CPU (gloo, world_size 2..3) tests of the multi-rank host logic: slice bounds, the
tagged carry-slot words, the carry each rank derives from the gathered totals.  The
per-slice scan itself is stood in for by the oracle here (the CUDA path is covered by the
gpu tests); what is checked is that sharding + carry reproduces the unsharded scan."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_total: int, exclusive: bool, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpuprefixsums_b200.sharded import carry_from_totals, pack_slot, slice_bounds, unpack_slot
        from tests import oracle_lib

        oracle = oracle_lib.load()
        lo, hi = slice_bounds(n_total, world, rank)
        x = oracle.random_u32(hi - lo, seed=123, first_index=lo)  # each rank generates only its slice
        # pass 1: local total, published as a tagged word; exchange = all_gather of G words
        _, total = oracle.scan_u32(x)
        tag = 7
        word = torch.tensor([pack_slot(total, tag) - (1 << 64) if pack_slot(total, tag) >= 1 << 63 else pack_slot(total, tag)],
                            dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, word)
        vals, tags = zip(*(unpack_slot(int(w.item())) for w in gathered))
        assert all(t == tag for t in tags)
        carry = carry_from_totals(list(vals), rank)
        # pass 2: scan seeded with the carry
        out, _ = oracle.scan_u32(x, exclusive=exclusive, carry=carry)
        q.put((rank, lo, hi, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,exclusive", [(2, 100_003, False), (2, 65_536, True), (3, 12_345, False)])
def test_sharded_scan_equals_unsharded(world, n_total, exclusive, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, exclusive, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.random_u32(n_total, seed=123)
    want, _ = oracle.scan_u32(full, exclusive=exclusive)
    got = np.empty_like(want)
    covered = 0
    for rank, lo, hi, out in sorted(parts):
        got[lo:hi] = out
        covered += hi - lo
    assert covered == n_total
    assert np.array_equal(got, want)


def test_slice_bounds_partition_and_alignment():
    from gpuprefixsums_b200.sharded import slice_bounds

    for n in [0, 1, 5, 4096, 100_003, 1 << 20, (1 << 33)]:
        for world in [1, 2, 3, 4, 8]:
            prev = 0
            for r in range(world):
                lo, hi = slice_bounds(n, world, r)
                assert lo == prev and lo <= hi and lo % 4 == 0 or lo == n
                prev = hi
            assert prev == n


def test_slot_words_roundtrip():
    from gpuprefixsums_b200.sharded import carry_from_totals, pack_slot, unpack_slot

    for v, t in [(0, 1), (0xFFFFFFFF, 0xFFFFFFFE), (123456789, 42)]:
        assert unpack_slot(pack_slot(v, t)) == (v, t)
    assert carry_from_totals([0xFFFFFFFF, 2, 5], 2) == 1  # wraps mod 2^32
    assert carry_from_totals([9, 9, 9], 0) == 0
