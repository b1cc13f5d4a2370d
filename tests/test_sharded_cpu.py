"""CPU (gloo, world_size 2 and 4): host-side logic of the multi-GPU commit -- column / coset-block partition, the
coefficient all-gather, the top-layer node exchange and finishing layers -- with the CPU oracle standing in for the
device engine.  The same orchestration function drives the GPU engine in bench.py / tests -m gpu."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from plonky2_gpu_b200 import sharded  # noqa: E402


def test_partition_functions():
    assert sharded.column_shard(135, 8, 0) == (0, 17, 17)
    assert sharded.column_shard(135, 8, 7) == (119, 135, 17)
    assert sum(c1 - c0 for c0, c1, _ in (sharded.column_shard(135, 8, r) for r in range(8))) == 135
    assert sharded.column_shard(3, 4, 3) == (3, 3, 1)  # more ranks than columns: empty shard
    assert sharded.block_shard(3, 8, 5) == (5, 1)
    assert sharded.block_shard(3, 2, 1) == (4, 4)
    with pytest.raises(ValueError):
        sharded.block_shard(1, 4, 0)
    with pytest.raises(ValueError):
        sharded.block_shard(3, 3, 0)
    # C2: 2^20 rows, rate 3, cap 4 on 8 ranks -> every rank reaches the cap (two cap entries each)
    assert sharded.local_top_layer(20, 3, 4, 8) == 19
    # cap_height 0 on 4 ranks: ranks stop two layers below the root
    assert sharded.local_top_layer(5, 3, 0, 4) == 6


def test_node_index_matches_oracle_layout():
    rng = np.random.default_rng(3)
    leaves = rng.integers(0, oracle.ORDER, size=(64, 6), dtype=np.uint64)
    cap_height = 2
    digests, cap = oracle.merkle_tree(leaves, cap_height)
    sub_log, sub_d = 6 - cap_height, digests.shape[0] >> cap_height
    layer = [oracle.hash_or_noop(l) for l in leaves]
    for l in range(sub_log + 1):
        for Q, h in enumerate(layer):
            where, idx = sharded.node_index(sub_log, sub_d, l, Q)
            got = cap[idx] if where == "cap" else digests[idx]
            assert np.array_equal(got, h)
        layer = [oracle.two_to_one(layer[2 * q], layer[2 * q + 1]) for q in range(len(layer) // 2)]


class OracleEngine:
    """CPU stand-in for GpuEngine: same interface, numpy/torch-CPU arrays, computed by the oracle."""

    def ifft_columns(self, t, ncols, n_log):
        a = t.numpy().view(np.uint64)
        for c in range(ncols):
            a[c] = oracle.ifft(a[c])

    def commit_blocks(self, coeffs, num_polys, n_log, rate_bits, cap_height, b0, bcount):
        co = coeffs.numpy().view(np.uint64)[:num_polys]
        full = oracle.batch_from_coeffs(co, rate_bits, cap_height)
        n, N = 1 << n_log, 1 << (n_log + rate_bits)
        sub_log = n_log + rate_bits - cap_height
        # keep only what this rank may legitimately know: its leaves and the digests derived from them
        mask_d = np.zeros(full.digests.shape[0], dtype=bool)
        sub_d = full.digests.shape[0] >> cap_height if cap_height <= n_log + rate_bits else 0
        cap = np.zeros_like(full.cap)
        leaf0, leaf1 = b0 * n, (b0 + bcount) * n
        l = 0
        top = 0
        while True:
            for Q in range(leaf0 >> l, leaf1 >> l):
                where, idx = sharded.node_index(sub_log, sub_d, l, Q)
                if where == "cap":
                    cap[idx] = full.cap[idx]
                else:
                    mask_d[idx] = True
            top = l
            if l == sub_log or (leaf0 % (1 << (l + 1))) or (leaf1 % (1 << (l + 1))):
                break
            l += 1
        digests = np.where(mask_d[:, None], full.digests, np.uint64(0xDEADBEEF))
        return {"leaves": full.leaves[leaf0:leaf1], "digests": digests, "cap": cap, "sub_log": sub_log, "sub_d": sub_d,
                "N": N, "top": top, "coeffs": co.copy(), "leaf0": leaf0}

    def export_nodes(self, b, layer, first, count):
        out = np.empty((count, 4), dtype=np.uint64)
        for i in range(count):
            where, idx = sharded.node_index(b["sub_log"], b["sub_d"], layer, first + i)
            out[i] = b["cap"][idx] if where == "cap" else b["digests"][idx]
        return torch.from_numpy(out.view(np.int64))

    def import_nodes(self, b, layer, first, count, t):
        a = t.numpy().view(np.uint64)
        for i in range(count):
            where, idx = sharded.node_index(b["sub_log"], b["sub_d"], layer, first + i)
            (b["cap"] if where == "cap" else b["digests"])[idx] = a[i]

    def finish_layers(self, b, from_layer):
        l = from_layer
        while l < b["sub_log"]:
            l += 1
            for Q in range(b["N"] >> l):
                kids = []
                for ch in (2 * Q, 2 * Q + 1):
                    where, idx = sharded.node_index(b["sub_log"], b["sub_d"], l - 1, ch)
                    kids.append(b["cap"][idx] if where == "cap" else b["digests"][idx])
                where, idx = sharded.node_index(b["sub_log"], b["sub_d"], l, Q)
                (b["cap"] if where == "cap" else b["digests"])[idx] = oracle.two_to_one(kids[0], kids[1])

    def cap(self, b):
        return b["cap"]

    def commit_begin(self, num_polys, n_log, rate_bits, cap_height, b0, bcount):
        return {"args": (num_polys, n_log, rate_bits, cap_height, b0, bcount), "cols": [], "done": 0}

    def commit_absorb(self, h, cols, col0, ncols):
        assert col0 == h["done"] and (ncols % 8 == 0 or col0 + ncols == h["args"][0])   # the C ABI's contract
        h["cols"].append(cols.numpy().view(np.uint64)[:ncols].copy())
        h["done"] += ncols

    def commit_finish(self, h):
        num_polys, n_log, rate_bits, cap_height, b0, bcount = h["args"]
        assert h["done"] == num_polys
        co = np.concatenate(h["cols"])
        return self.commit_blocks(torch.from_numpy(co.view(np.int64)), num_polys, n_log, rate_bits, cap_height, b0, bcount)

    def pack_open_rows(self, b, idx, slots, Q, leaf_len, layers):
        packed = np.zeros((Q, leaf_len + 4 * layers), dtype=np.uint64)
        for x, slot in zip(idx, slots):
            packed[slot, :leaf_len] = b["leaves"][x - b["leaf0"]]
            for l in range(layers):   # MerkleTree::prove (merkle_tree.rs:392-440): sibling of the layer-l ancestor
                where, i = sharded.node_index(b["sub_log"], b["sub_d"], l, (x >> l) ^ 1)
                assert where == "digests"
                packed[slot, leaf_len + 4 * l: leaf_len + 4 * l + 4] = b["digests"][i]
        return torch.from_numpy(packed.view(np.int64))

    def to_numpy(self, t):
        return t.numpy().view(np.uint64)


def test_exchange_schedule():
    # 135 columns on 8 ranks: rounds of 8, 8, 16, 32, 64, 7 columns; rank r contributes `per` consecutive columns of each
    assert sharded.exchange_schedule(135, 8) == [(0, 8, 1), (8, 8, 1), (16, 16, 2), (32, 32, 4), (64, 64, 8), (128, 7, 1)]
    assert sharded.exchange_schedule(135, 2)[:4] == [(0, 8, 4), (8, 8, 4), (16, 16, 8), (32, 16, 8)]
    assert sharded.exchange_schedule(5, 4) == [(0, 5, 2)]
    for P, world in [(135, 8), (135, 4), (135, 2), (234, 8), (400, 8), (9, 2), (5, 4), (16, 8)]:
        sched = sharded.exchange_schedule(P, world)
        assert [r[0] for r in sched] == [sum(x[1] for x in sched[:j]) for j in range(len(sched))] and sum(r[1] for r in sched) == P
        assert all(w % 8 == 0 for _, w, _ in sched[:-1])                      # sponge-rate aligned except the last round
        covered = []
        for j, (col0, width, per) in enumerate(sched):
            for r in range(world):
                rows, layout = sharded.local_layout(P, world, r)
                row0, c0, c1 = layout[j]
                assert c1 - c0 <= per and row0 == sum(x[2] for x in sched[:j])
                covered += list(range(c0, c1))
        assert covered == list(range(P))                                      # rank-major order inside a round == column order


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_log, P, rate_bits, cap_height, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle.set_threads(1)
        rng = np.random.default_rng(99)  # same matrix on every rank; each takes its own columns
        values = rng.integers(0, oracle.ORDER, size=(P, 1 << n_log), dtype=np.uint64)
        c0, c1, cmax = sharded.column_shard(P, world, rank)
        shard = np.zeros((cmax, 1 << n_log), dtype=np.uint64)
        shard[: c1 - c0] = values[c0:c1]
        comm = sharded.TorchComm(dist)
        batch = sharded.sharded_commit_from_values(OracleEngine(), comm, torch.from_numpy(shard.view(np.int64)), P, n_log,
                                                   rate_bits, cap_height)
        ref = oracle.batch_from_values(values, rate_bits, cap_height)
        ok = np.array_equal(batch["cap"], ref.cap) and np.array_equal(batch["coeffs"], ref.coeffs)
        n = 1 << n_log
        b0, bc = sharded.block_shard(rate_bits, world, rank)
        ok = ok and np.array_equal(batch["leaves"], ref.leaves[b0 * n:(b0 + bc) * n])
        # FRI query openings: rows + Merkle paths of leaves owned by different ranks reach every rank
        N = n << rate_bits
        idx = [0, N - 1, N // 2, (N // 2 - 1) % N, 1 % N]
        rows, sibs = sharded.sharded_open_rows(OracleEngine(), comm, batch, idx, n_log, rate_bits, cap_height, P)
        for x, r, sb in zip(idx, rows, sibs):
            ok = ok and np.array_equal(r, ref.leaves[x]) and oracle.merkle_verify(r, x, ref.cap, sb)
            ok = ok and np.array_equal(sb, oracle.merkle_prove(ref.digests, N, cap_height, x))
        # pipelined variant (growing exchange rounds, one all-gather per round): same cap, coefficients, leaves
        blocks = sharded.pack_local(values, P, world, rank)
        if P > 4:
            pb = sharded.sharded_commit_from_values_pipelined(OracleEngine(), comm, torch.from_numpy(blocks.view(np.int64)), P, n_log,
                                                              rate_bits, cap_height)
            ok = ok and np.array_equal(pb["cap"], ref.cap) and np.array_equal(pb["coeffs"], ref.coeffs)
            ok = ok and np.array_equal(pb["leaves"], ref.leaves[b0 * n:(b0 + bc) * n])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_log,P,rate_bits,cap_height", [(2, 4, 5, 3, 4), (2, 3, 3, 1, 0), (4, 3, 7, 3, 1), (4, 4, 2, 2, 0)])
def test_sharded_commit_gloo(world, n_log, P, rate_bits, cap_height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_log, P, rate_bits, cap_height, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
