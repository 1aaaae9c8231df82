"""CPU: the multi-GPU host logic (Morton partition + guard-zone exchange plan) with REAL two-process
communication over torch.distributed / gloo.  Each rank fills the blocks it owns with a known function
of (global block, field, i, j), packs the strips its plan says to send, exchanges them, unpacks into its
ghost blocks, and then checks every cell the stage kernel would read through its neighbour table (the
two-cell halo of every owned block, corners included, periodic wrap) against the true neighbour."""
import os
import socket
import sys
import numpy as np
import pytest
import mara3_b200 as m3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def known(global_block, N):
    q, i, j = np.meshgrid(np.arange(3), np.arange(N), np.arange(N), indexing="ij")
    return global_block * 1000.0 + q * 100.0 + i + j / 64.0


def region(N, di, dj):
    """Cells of the SOURCE block that face back towards the block needing them (partition.hpp)."""
    rows = slice(N - 2, N) if di < 0 else (slice(0, 2) if di > 0 else slice(0, N))
    cols = slice(N - 2, N) if dj < 0 else (slice(0, 2) if dj > 0 else slice(0, N))
    return rows, cols


def worker(rank, world, port, cfg, q):
    import torch
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        s = m3.Solver(cfg, host_only=True, rank=rank, nranks=world)
        N, owned, L = s.block_size, s.num_blocks, s.num_local_blocks
        l2g = s.local_to_global
        U = np.full((L, 3, N, N), np.nan)
        for b in range(owned):
            U[b] = known(l2g[b], N)

        # pack -> exchange -> unpack, following the plan in order
        ops, recv_bufs = [], {}
        for peer in range(world):
            if peer == rank:
                continue
            send = s.halo_plan(peer, True)
            recv = s.halo_plan(peer, False)
            if len(send):
                parts = []
                for b, di, dj in send:
                    assert b < owned
                    rows, cols = region(N, di, dj)
                    parts.append(U[b][:, rows, cols].reshape(-1))
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(np.concatenate(parts)), peer))
            if len(recv):
                n = sum(3 * (2 if di else N) * (2 if dj else N) for _, di, dj in recv)
                recv_bufs[peer] = torch.empty(n, dtype=torch.float64)
                ops.append(dist.P2POp(dist.irecv, recv_bufs[peer], peer))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for peer, buf in recv_bufs.items():
            buf, at = buf.numpy(), 0
            for b, di, dj in s.halo_plan(peer, False):
                assert b >= owned
                rows, cols = region(N, di, dj)
                shape = U[b][:, rows, cols].shape
                U[b][:, rows, cols] = buf[at:at + int(np.prod(shape))].reshape(shape)
                at += int(np.prod(shape))

        # every halo cell the stage kernels read is now the true neighbour's value
        full = m3.Solver(cfg, host_only=True)                       # whole tree on one rank: ground truth
        table, full_table = s.neighbor_table, full.neighbor_table
        faces, full_faces = s.face_neighbor_table, full.face_neighbor_table
        checked = 0
        for b in range(owned):
            g = l2g[b]
            if (table[b] >= 0).all():
                # fused kernel: two-cell strips / corners of the eight same-level neighbours
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        if di == 0 and dj == 0:
                            continue
                        n_local, n_global = table[b, di + 1, dj + 1], full_table[g, di + 1, dj + 1]
                        assert n_local >= 0 and l2g[n_local] == n_global
                        rows, cols = region(N, di, dj)
                        assert np.array_equal(U[n_local][:, rows, cols], known(n_global, N)[:, rows, cols])
                        checked += 1
            else:
                # any-tree kernels: whole face neighbours (layer 1) and whole face neighbours of those (layer 2)
                assert (full_table[g] < 0).any()
                for side in range(4):
                    assert faces[b, side, 0] == full_faces[g, side, 0]
                    for l1, g1 in zip(faces[b, side, 1:], full_faces[g, side, 1:]):
                        assert (l1 < 0) == (g1 < 0)
                        if l1 < 0:
                            continue
                        assert l2g[l1] == g1 and np.array_equal(U[l1], known(g1, N))
                        for side2 in range(4):
                            assert faces[l1, side2, 0] == full_faces[g1, side2, 0]
                            for l2, g2 in zip(faces[l1, side2, 1:], full_faces[g1, side2, 1:]):
                                assert (l2 < 0) == (g2 < 0)
                                if l2 >= 0:
                                    assert l2g[l2] == g2 and np.array_equal(U[l2], known(g2, N))
                checked += 8
        first, count = s.first_block, owned
        counts = [None] * world
        dist.all_gather_object(counts, (first, count, L - owned, checked))
        dist.destroy_process_group()
        q.put((rank, "ok", counts))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "error: " + traceback.format_exc(), None))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("cfg,world", [
    (dict(depth=3, block_size=8, focus_factor=1e3), 2),      # 64 blocks, 8x8
    (dict(depth=2, block_size=6, focus_factor=1e3), 2),      # 16 blocks, 4x4: wrap neighbours are everywhere
    (dict(depth=3, block_size=4, focus_factor=1e3), 3),      # uneven ranges: 21 / 21 / 22
    (dict(depth=4, block_size=8), 2),                        # nested tree (64 leaves, levels 2-4): whole-block ghosts at the jumps
    (dict(depth=5, block_size=4), 3),                        # deeper nesting on three ranks
])
def test_exchange_plan_with_gloo(cfg, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, cfg, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, counts in results:
        assert status == "ok", status
    counts = results[0][2]
    total = sum(c[1] for c in counts)
    assert total == m3.Solver(cfg, host_only=True).num_blocks
    assert [c[0] for c in counts] == [sum(x[1] for x in counts[:k]) for k in range(world)]      # contiguous Morton ranges
    assert max(c[1] for c in counts) - min(c[1] for c in counts) <= 1                            # balanced
    assert all(c[2] > 0 and c[3] == 8 * c[1] for c in counts)                                    # ghosts exist; all 8 neighbours checked


def test_partition_is_identity_on_one_rank():
    s = m3.Solver(dict(depth=3, block_size=8), host_only=True)
    assert s.num_blocks == s.num_global_blocks == s.num_local_blocks and s.first_block == 0
    assert np.array_equal(s.local_to_global, np.arange(s.num_blocks))
    assert len(s.halo_plan(0, True)) == 0


def test_nested_trees_store_whole_ghost_blocks():
    s = m3.Solver(dict(depth=4, block_size=8), host_only=True, rank=0, nranks=2)
    plan = s.halo_plan(1, False)
    assert s.num_blocks == 32 and len(plan) > 0
    assert ((plan[:, 1] == 0) & (plan[:, 2] == 0)).any()         # whole blocks for the any-tree kernels
    assert (plan[:, 0] >= s.num_blocks).all()
