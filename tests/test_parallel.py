"""World-size-2 gloo tests (CPU) of the sharding logic used on the N>1 path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_is_a_partition():
    from marconet_b200.parallel import shard_range
    for n in (0, 1, 5, 16, 17, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_generator(styles, labels, noise):
    """Deterministic stand-in with TSPGAN's contract and memory formats (channels_last views)."""
    n = labels.shape[0]
    base = (labels.float().reshape(n, 1, 1, 1) + styles.sum(dim=1).reshape(n, 1, 1, 1))
    img = (base + torch.arange(3 * 8 * 8).reshape(1, 3, 8, 8)).contiguous(memory_format=torch.channels_last)
    f64 = (base * 2 + torch.arange(4 * 4 * 4).reshape(1, 4, 4, 4)).contiguous(memory_format=torch.channels_last)
    f32 = (base * 3 + torch.arange(6 * 2 * 2).reshape(1, 6, 2, 2)).contiguous(memory_format=torch.channels_last)
    return img, f64, f32


def _worker(rank, world, port, n_chars, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from marconet_b200.parallel import generate_priors_sharded, shard_lines
        g = torch.Generator().manual_seed(5)
        styles = torch.randn(n_chars, 16, generator=g)
        labels = torch.randint(0, 100, (n_chars, 1), generator=g)
        full = _fake_generator(styles, labels, None)
        got = generate_priors_sharded(_fake_generator, styles, labels)
        good = all(torch.equal(a, b) for a, b in zip(full, got))
        if n_chars % 4 == 0:   # pipelined variant: 2 sub-chunks per rank, async all-gathers
            got2 = generate_priors_sharded(_fake_generator, styles, labels, pipeline_chunks=2)
            good &= all(torch.equal(a, b) for a, b in zip(full, got2))
        good &= all(a.permute(0, 2, 3, 1).is_contiguous() for a in got)
        lines = list(shard_lines(5))
        good &= lines == ([0, 1, 2] if rank == 0 else [3, 4])
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chars", [16, 5, 1])
def test_char_sharded_generation_world2_gloo(n_chars):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, n_chars, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def _owner_worker(rank, world, port, n_chars, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from marconet_b200.parallel import exchange_bytes_per_rank, generate_priors_for_owners
        g = torch.Generator().manual_seed(9)
        styles = torch.randn(n_chars, 16, generator=g)
        labels = torch.randint(0, 100, (n_chars, 1), generator=g)
        full = _fake_generator(styles, labels, None)
        own = slice(rank * n_chars // world, (rank + 1) * n_chars // world)
        f64, f32 = generate_priors_for_owners(_fake_generator, styles, labels)
        good = torch.equal(f64, full[1][own]) and torch.equal(f32, full[2][own])          # natural order, owned characters only
        good &= f64.permute(0, 2, 3, 1).is_contiguous()
        img, = generate_priors_for_owners(_fake_generator, styles, labels, keep=(0,))
        good &= torch.equal(img, full[0][own])
        loc = generate_priors_for_owners(_fake_generator, styles, labels, exchange=False)
        good &= loc[0].shape[0] == n_chars // world
        b = exchange_bytes_per_rank(n_chars, world, bytes_per_char=10)
        good &= b["all_to_all"] * world == b["all_gather"]
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chars", [4, 16])
def test_owner_only_exchange_world2_gloo(n_chars):
    """Block-cyclic generation + all-to-all: every rank ends up with exactly its own lines' priors, in order."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_owner_worker, args=(2, port, n_chars, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


def test_owner_blocks_rejects_ragged_counts():
    from marconet_b200.parallel import owner_blocks
    assert owner_blocks(1024, 8) == 16
    with pytest.raises(ValueError):
        owner_blocks(100, 8)
