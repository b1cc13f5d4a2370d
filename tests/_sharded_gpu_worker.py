"""torchrun worker for tests/test_gpu_sharded.py: sharded commit over WORLD_SIZE GPUs vs the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import plonky2_gpu_b200 as p2b  # noqa: E402
from plonky2_gpu_b200 import sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = p2b.Context(local)
    engine, comm = sharded.GpuEngine(ctx), sharded.TorchComm(dist)
    ok = True
    for (n_log, P, rate_bits, cap_height) in [(10, 20, 3, 4), (11, 135, 3, 4), (9, 5, 3, 0), (8, 3, 1, 0), (12, 9, 2, 1)]:
        if world > (1 << rate_bits):
            continue
        rng = np.random.default_rng(1234 + n_log)
        values = rng.integers(0, oracle.ORDER, size=(P, 1 << n_log), dtype=np.uint64)
        c0, c1, cmax = sharded.column_shard(P, world, rank)
        shard = np.zeros((cmax, 1 << n_log), dtype=np.uint64)
        shard[: c1 - c0] = values[c0:c1]
        t = torch.from_numpy(shard.view(np.int64)).cuda()
        b = sharded.sharded_commit_from_values(engine, comm, t, P, n_log, rate_bits, cap_height)
        ref = oracle.batch_from_values(values, rate_bits, cap_height)
        n = 1 << n_log
        b0, bc = sharded.block_shard(rate_bits, world, rank)
        good = np.array_equal(b.cap(), ref.cap) and np.array_equal(b.polynomials(), ref.coeffs)
        good = good and np.array_equal(b.leaves(), ref.leaves[b0 * n:(b0 + bc) * n])
        # an opened row of this shard verifies against the (global) cap
        idx = [b0 * n, (b0 + bc) * n - 1]
        rows, sibs = b.open_rows(idx)
        for r, s, i in zip(rows, sibs, idx):
            good = good and oracle.merkle_verify(r, i, b.cap(), s)
        # FRI query openings across the shards (rows owned by different ranks, gathered with one all-reduce)
        N = n << rate_bits
        qidx = [0, N - 1, N // 2, N // 2 - 1, 1, N // 4 + 3]
        qrows, qsibs = sharded.sharded_open_rows(engine, comm, b, qidx, n_log, rate_bits, cap_height, P)
        for x, r, sb in zip(qidx, qrows, qsibs):
            good = good and np.array_equal(r, ref.leaves[x]) and np.array_equal(sb, oracle.merkle_prove(ref.digests, N, cap_height, x))
        # pipelined variant: growing exchange rounds, exchange overlapped with LDE + progressive leaf hashing
        if P > 4:
            blocks = sharded.pack_local(values, P, world, rank)
            tb = torch.from_numpy(blocks.view(np.int64)).cuda()
            pb = sharded.sharded_commit_from_values_pipelined(engine, comm, tb, P, n_log, rate_bits, cap_height)
            good = good and np.array_equal(pb.cap(), ref.cap) and np.array_equal(pb.polynomials(), ref.coeffs)
            good = good and np.array_equal(pb.leaves(), ref.leaves[b0 * n:(b0 + bc) * n])
            prow, psib = pb.open_rows(idx)
            for r, s, i in zip(prow, psib, idx):
                good = good and oracle.merkle_verify(r, i, pb.cap(), s)
            pb.close()
        if not good:
            print("rank %d MISMATCH for" % rank, (n_log, P, rate_bits, cap_height), flush=True)
        ok = ok and good
        b.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if int(flag[0]) == 1 else "SHARDED_FAIL", flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
