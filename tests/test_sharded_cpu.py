"""CPU tests of the multi-GPU host logic (world_size-2 and -3 gloo groups, no GPU): the slab partition and the halo-row
exchange of spatial_sbml_b200/sharded.py — the same function the NCCL path calls on device tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatial_sbml_b200.sharded import exchange_rows, partition


def test_partition_covers_every_row_once():
    for n in (1, 2, 7, 22, 301, 8192):
        for world in (1, 2, 3, 4, 8):
            if world > n:
                continue
            parts = partition(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and b > a
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_wide(rank, world, port, n_rows, unit, width, out):
    """The exchange of `width` edge rows (4: what the apron schedule and the host step of a slab send once per step)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = partition(n_rows, world)[rank]
        held_lo, held_hi = max(0, lo - width), min(n_rows, hi + width)
        own_lo, own_hi = lo - held_lo, hi - held_lo
        base = unit + 4
        t = torch.full((base + (held_hi - held_lo + 2) * unit,), -1.0, dtype=torch.float64)
        for r in range(own_lo, own_hi):
            t[base + r * unit: base + (r + 1) * unit] = torch.arange(unit, dtype=torch.float64) + 1000.0 * (held_lo + r)
        for req in exchange_rows([t], unit, base, own_lo, own_hi, rank, world, width=width):
            req.wait()
        ok = True
        for r in range(held_hi - held_lo):
            g = held_lo + r
            want = torch.arange(unit, dtype=torch.float64) + 1000.0 * g
            ok = ok and bool(torch.equal(t[base + r * unit: base + (r + 1) * unit], want))  # owned rows and all halo rows
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width", [(2, 4), (3, 4), (3, 2)])
def test_wide_halo_exchange_over_gloo(world, width):
    port = _free_port()
    out = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker_wide, args=(world, port, 41, 24, width, out), nprocs=world, join=True)
    assert list(out) == [1] * world


def _worker(rank, world, port, n_rows, unit, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = partition(n_rows, world)[rank]
        held_lo = max(0, lo - 2)
        held_hi = min(n_rows, hi + 2)
        own_lo, own_hi = lo - held_lo, hi - held_lo
        base = unit + 4  # one guard row + front pad, as in the device layout
        # global field: row r holds r*1000 + column; every rank fills only the rows it owns
        fields = []
        for f in range(3):
            t = torch.full((base + (held_hi - held_lo + 2) * unit,), -1.0, dtype=torch.float64)
            for r in range(own_lo, own_hi):
                g = held_lo + r
                t[base + r * unit: base + (r + 1) * unit] = torch.arange(unit, dtype=torch.float64) + 1000.0 * g + 1e6 * f
            fields.append(t)
        for req in exchange_rows(fields, unit, base, own_lo, own_hi, rank, world):
            req.wait()
        ok = True
        for f, t in enumerate(fields):
            for r, g in ((own_lo - 1, lo - 1), (own_hi, hi)):
                if g < 0 or g >= n_rows:
                    continue
                want = torch.arange(unit, dtype=torch.float64) + 1000.0 * g + 1e6 * f
                ok = ok and bool(torch.equal(t[base + r * unit: base + (r + 1) * unit], want))
            # nothing but the two halo rows may have changed outside the owned range
            for r in range(-1, held_hi - held_lo + 1):
                if own_lo - 1 <= r <= own_hi:
                    continue
                ok = ok and bool((t[base + r * unit: base + (r + 1) * unit] == -1.0).all())
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_over_gloo(world):
    port = _free_port()
    out = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, port, 23, 40, out), nprocs=world, join=True)
    assert list(out) == [1] * world
