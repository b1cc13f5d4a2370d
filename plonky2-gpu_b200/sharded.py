"""Multi-GPU commit: one process per GPU, coset blocks / cap sub-trees sharded over the ranks.

Partition (strong scaling of ONE commit over G ranks, G a power of two, G <= 2^rate_bits):
  * inverse NTT is per column            -> rank r transforms columns [c0_r, c1_r) of the value matrix,
  * one exchange step                    -> all-gather of the coefficient columns (NCCL over NVLink),
  * LDE + leaf hashing + digest layers   -> rank r owns coset blocks [r*R/G, (r+1)*R/G), i.e. the contiguous leaf range
                                            [r*N/G, (r+1)*N/G); no communication,
  * cap                                  -> all-gather of each rank's top-layer nodes (32 B each); if those are not yet
                                            the cap entries (2^cap_height < G) every rank finishes the few top layers.
The reference has no multi-GPU path (cudaSetDevice(0) only in dead code, cuda/plonky2_gpu.cu:63); the partition follows
from PolynomialBatch::from_values' own structure: `values.into_par_iter().map(ifft)` (fri/oracle.rs:717-721) and
fill_digests_buf's independent cap sub-trees (hash/merkle_tree.rs:232-243).

The orchestration is written against two small interfaces so that the host-side logic is testable on CPU with gloo:
  engine : ifft_columns / commit_blocks / export_nodes / import_nodes / finish_layers / cap / pack_open_rows   (GpuEngine = the C ABI)
  comm   : rank, world, all_gather_columns / all_gather_nodes / all_reduce_sum                                  (TorchComm = torch.distributed)
"""
import ctypes as C
import math

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# host-side planning (pure functions)
# ---------------------------------------------------------------------------------------------------------------
def column_shard(num_polys, world, rank):
    """Columns [c0, c1) transformed by `rank`.  Every rank but the last gets ceil(P / world) columns so that the
    all-gathered (padded) buffer holds the P real columns contiguously at its start."""
    cmax = -(-num_polys // world)
    c0 = min(rank * cmax, num_polys)
    c1 = min(c0 + cmax, num_polys)
    return c0, c1, cmax


def block_shard(rate_bits, world, rank):
    """Coset blocks [b0, b0 + count) owned by `rank` (block b = leaves [b*n, (b+1)*n))."""
    R = 1 << rate_bits
    if world < 1 or world & (world - 1):
        raise ValueError("world size must be a power of two")
    if world > R:
        raise ValueError("world size %d exceeds the 2^rate_bits = %d coset blocks" % (world, R))
    per = R // world
    return rank * per, per


def local_top_layer(n_log, rate_bits, cap_height, world):
    """Highest digest layer a rank can compute from its own leaves (0 = leaf digests)."""
    sub_log = n_log + rate_bits - cap_height
    span_log = n_log + rate_bits - int(math.log2(world))  # log2(leaves per rank)
    return min(sub_log, span_log)


def node_index(sub_log, sub_digests, layer, Q):
    """Index (in hashes) of node Q of `layer` inside the reference's digest buffer (merkle_tree.rs:46-54, 424-435);
    layer == sub_log addresses cap[Q] instead (returned as ('cap', Q))."""
    if layer == sub_log:
        return ("cap", Q)
    bits = sub_log - layer
    tree, q = Q >> bits, Q & ((1 << bits) - 1)
    return ("digests", tree * sub_digests + 2 * (((q >> 1) << (layer + 1)) + (1 << layer) - 1) + (q & 1))


# ---------------------------------------------------------------------------------------------------------------
# orchestration
# ---------------------------------------------------------------------------------------------------------------
def sharded_commit_from_values(engine, comm, values_shard, num_polys, n_log, rate_bits, cap_height, on_coeffs_ready=None,
                               host_values=None):
    """values_shard: this rank's columns [c0, c1) of the value matrix, engine-native array [cmax][n] (rows beyond
    c1 - c0 are padding).  Returns the engine's batch handle; afterwards engine.cap(batch) is the full cap on every
    rank and the batch holds this rank's leaves / digests.
    on_coeffs_ready(values_shard): called as soon as this rank's coefficient columns are final (the reference keeps the
    coefficients host-side, fri/oracle.rs:403-407: a caller starts its device-to-host copy here, overlapped with the
    exchange, the LDE and the tree).
    host_values: this rank's value columns in pinned HOST memory (engine-native [cmax][n]); when given they are uploaded
    into values_shard in column groups and each group's inverse NTT starts as soon as it has landed."""
    rank, world = comm.rank, comm.world
    c0, c1, cmax = column_shard(num_polys, world, rank)
    b0, bcount = block_shard(rate_bits, world, rank)
    if host_values is not None:
        engine.ifft_columns_from_host(host_values, values_shard, c1 - c0, n_log)   # upload in column groups, each transformed as it lands
    else:
        engine.ifft_columns(values_shard, c1 - c0, n_log)             # in place
    if on_coeffs_ready is not None:
        on_coeffs_ready(values_shard)
    coeffs_all = comm.all_gather_columns(values_shard, cmax, n_log)   # [world * cmax][n]; first num_polys rows are real
    batch = engine.commit_blocks(coeffs_all, num_polys, n_log, rate_bits, cap_height, b0, bcount)
    top = local_top_layer(n_log, rate_bits, cap_height, world)
    count = ((1 << (n_log + rate_bits)) // world) >> top               # top-layer nodes per rank
    mine = engine.export_nodes(batch, top, rank * count, count)        # [count][4]
    everyone = comm.all_gather_nodes(mine, count)                      # [world * count][4], rank-major == node order
    engine.import_nodes(batch, top, 0, world * count, everyone)
    engine.finish_layers(batch, top)
    return batch


# ---------------------------------------------------------------------------------------------------------------
# pipelined variant: the exchange of one column group overlaps the LDE + leaf hashing of the previous one
# ---------------------------------------------------------------------------------------------------------------
SPONGE_RATE = 8


def exchange_schedule(num_polys, world):
    """Rounds of the pipelined exchange: [(col0, width, per_rank)].  Round j delivers the `width` CONSECUTIVE columns
    [col0, col0 + width) -- the order in which the leaves' sponges absorb them (hashing.rs:81-104), so every width but the
    last is a multiple of the sponge rate 8 -- of which rank r contributes [col0 + r * per_rank, col0 + (r + 1) * per_rank)
    (clipped to the round).  Widths grow 8, 8, 16, 32, ... up to 8 * world: the first rounds are small so that hashing
    starts after ONE column per rank has been uploaded, transformed, exchanged and extended (at 8 ranks the pipeline
    prologue drops from 64 columns to 8), the later ones large so that launches and collectives stay few.
    Mirrored in C++ by mgpu_schedule() (csrc/mgpu.cuh)."""
    rounds, col0, w = [], 0, SPONGE_RATE
    while col0 < num_polys:
        width = min(w, num_polys - col0)
        rounds.append((col0, width, -(-width // world)))
        col0 += width
        if len(rounds) >= 2:
            w = min(2 * w, SPONGE_RATE * world)
    return rounds


def local_layout(num_polys, world, rank):
    """This rank's slice of every round: (rows_total, [(row0, c0, c1)] per round) -- its columns [c0, c1) of round j sit at
    rows [row0, row0 + per_rank_j) of its local buffer [rows_total][n] (zero rows where the round is ragged)."""
    row, out = 0, []
    for col0, width, per in exchange_schedule(num_polys, world):
        c0 = min(col0 + rank * per, col0 + width)
        c1 = min(c0 + per, col0 + width)
        out.append((row, c0, c1))
        row += per
    return row, out


def pack_local(values, num_polys, world, rank):
    """The rank's local buffer (numpy uint64 [rows_total][n]) from the full value matrix [P][n] (tests / examples)."""
    rows, layout = local_layout(num_polys, world, rank)
    buf = np.zeros((rows, values.shape[1]), dtype=np.uint64)
    for row0, c0, c1 in layout:
        buf[row0: row0 + (c1 - c0)] = values[c0:c1]
    return buf


def sharded_commit_from_values_pipelined(engine, comm, values_local, num_polys, n_log, rate_bits, cap_height,
                                         on_coeffs_ready=None, host_values=None):
    """Same result as sharded_commit_from_values with the exchange hidden behind the hashing.
    values_local: this rank's columns in the layout of local_layout() (engine-native [rows_total][n]).  Steps: inverse NTT
    of the local columns; one asynchronous all-gather per round (NCCL runs them back to back on its own stream); as soon as
    round j has landed its consecutive coefficient columns are LDE'd into the leaf rows of this rank's coset blocks and
    absorbed by the leaves' sponges (p2b_commit_blocks_absorb) while round j + 1 is still in flight; digest layers and the
    top-layer node exchange as in the unpipelined flow."""
    rank, world = comm.rank, comm.world
    sched = exchange_schedule(num_polys, world)
    rows_total, layout = local_layout(num_polys, world, rank)
    b0, bcount = block_shard(rate_bits, world, rank)
    batch = engine.commit_begin(num_polys, n_log, rate_bits, cap_height, b0, bcount)
    pending = []

    def absorb(j):
        cols = comm.wait(pending[j])                                    # [world * per_j][n]: the round's columns in order
        col0, width, _ = sched[j]
        engine.commit_absorb(batch, cols, col0, width)

    # software pipeline over the rounds: (upload +) inverse NTT of this rank's slice of round j -> its all-gather starts ->
    # LDE + absorb of round j - 1 is enqueued, so the GPU hashes round j - 1 while round j is exchanged and j + 1 uploads
    if host_values is None:
        engine.ifft_columns(values_local, rows_total, n_log)            # device-resident input: all local rows at once (zero rows stay zero)
    for j, (row0, c0, c1) in enumerate(layout):
        per = sched[j][2]
        blk = values_local[row0:row0 + per]
        if host_values is not None:
            engine.ifft_columns_from_host(host_values[row0:row0 + per], blk, per, n_log, groups=1)
        pending.append(comm.all_gather_async(blk))
        if j >= 1:
            absorb(j - 1)
    if on_coeffs_ready is not None:
        on_coeffs_ready(values_local)
    absorb(len(sched) - 1)
    batch = engine.commit_finish(batch)
    top = local_top_layer(n_log, rate_bits, cap_height, world)
    count = ((1 << (n_log + rate_bits)) // world) >> top
    mine_nodes = engine.export_nodes(batch, top, rank * count, count)
    everyone = comm.all_gather_nodes(mine_nodes, count)
    engine.import_nodes(batch, top, 0, world * count, everyone)
    engine.finish_layers(batch, top)
    return batch


def sharded_open_rows(engine, comm, batch, indices, n_log, rate_bits, cap_height, leaf_len):
    """FRI query openings over a sharded batch (fri/prover.rs:187-216: `t.get(x_index)`, `t.prove(x_index)` for every
    query): the rank that owns leaf x (contiguous range [rank*N/G, (rank+1)*N/G)) gathers the row and its Merkle path --
    its own digests below the local top layer, the layers above are complete on every rank after
    sharded_commit_from_values -- and one all-reduce (sum; every query has exactly one owner, the others contribute
    zeros) hands all rows and paths to every rank.  Returns (rows [Q][leaf_len], siblings [Q][layers][4]) as uint64."""
    rank, world = comm.rank, comm.world
    per = (1 << (n_log + rate_bits)) // world
    layers = n_log + rate_bits - cap_height
    width = leaf_len + 4 * layers
    mine = [k for k, x in enumerate(indices) if x // per == rank]
    packed = engine.pack_open_rows(batch, [indices[k] for k in mine], mine, len(indices), leaf_len, layers)  # [Q][width]
    total = comm.all_reduce_sum(packed)
    a = engine.to_numpy(total).reshape(len(indices), width)
    return a[:, :leaf_len].copy(), a[:, leaf_len:].reshape(len(indices), layers, 4).copy()


# ---------------------------------------------------------------------------------------------------------------
# GPU engine / torch.distributed communicator
# ---------------------------------------------------------------------------------------------------------------
class GpuEngine:
    """The C ABI (libplonky2_b200.so) on torch CUDA tensors (torch is plumbing: device memory + NCCL).

    No host synchronisation between the phases: work is handed between torch's current stream (where NCCL collectives are
    ordered) and the library's stream with CUDA events in both directions (`_sync_in` / `_sync_out`); tensors the library
    reads asynchronously are kept referenced until `synchronize()`.  The caller waits once, at the end of a commit
    (`engine.synchronize()`), or implicitly through a getter that copies to the host."""

    def __init__(self, ctx):
        import torch
        from . import lib, _check, PolynomialBatch
        self.torch, self.lib, self._check, self._PB, self.ctx = torch, lib(), _check, PolynomialBatch, ctx
        self.lib.p2b_ctx_stream.restype = C.c_void_p
        self._lib_stream = torch.cuda.ExternalStream(self.lib.p2b_ctx_stream(ctx.handle))
        self._copy_stream = torch.cuda.Stream()
        self._keep = []

    def synchronize(self):
        """Wait for everything enqueued so far and release the tensors the library was still reading."""
        self.ctx.synchronize()
        self.torch.cuda.current_stream().synchronize()
        self._copy_stream.synchronize()
        self._keep = []

    def _sync_in(self, *tensors):
        """torch-side producers (NCCL, copies) -> the library's stream: a device-side wait, the host does not block."""
        ev = self.torch.cuda.Event()
        ev.record(self.torch.cuda.current_stream())
        self._lib_stream.wait_event(ev)
        self._keep.extend(tensors)

    def _sync_out(self):
        """the library's results -> torch's current stream (the next collective is ordered behind them)."""
        ev = self.torch.cuda.Event()
        ev.record(self._lib_stream)
        self.torch.cuda.current_stream().wait_event(ev)

    def ifft_columns(self, t, ncols, n_log):
        if ncols == 0:
            return
        self._sync_in(t)
        self._check(self.lib.p2b_ifft_batch(self.ctx.handle, t.data_ptr(), t.data_ptr(), n_log, ncols))
        self._sync_out()

    def ifft_columns_from_host(self, host_t, t, ncols, n_log, groups=8):
        """H2D on a copy stream in column groups; the library's stream waits for each group's event and transforms it."""
        if ncols == 0:
            return
        torch = self.torch
        self._sync_in(t)
        self._copy_stream.wait_stream(torch.cuda.current_stream())
        for g in range(groups):
            a, b = ncols * g // groups, ncols * (g + 1) // groups
            if a == b:
                continue
            with torch.cuda.stream(self._copy_stream):
                t[a:b].copy_(host_t[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._lib_stream.wait_event(ev)
            self._check(self.lib.p2b_ifft_batch(self.ctx.handle, t[a:b].data_ptr(), t[a:b].data_ptr(), n_log, b - a))
        self._sync_out()

    def commit_blocks(self, coeffs, num_polys, n_log, rate_bits, cap_height, b0, bcount):
        self._sync_in(coeffs)
        h = C.c_void_p()
        self._check(self.lib.p2b_commit_blocks(self.ctx.handle, coeffs.data_ptr(), n_log, num_polys, rate_bits, cap_height,
                                               None, b0, bcount, C.byref(h)))
        return self._PB(self.ctx, h.value)

    def commit_begin(self, num_polys, n_log, rate_bits, cap_height, b0, bcount):
        h = C.c_void_p()
        self._check(self.lib.p2b_commit_blocks_begin(self.ctx.handle, n_log, num_polys, rate_bits, cap_height, b0, bcount, C.byref(h)))
        return h

    def commit_absorb(self, h, cols, col0, ncols):
        self._sync_in(cols)      # ordered behind the all-gather that produced `cols`; the previous group's kernels keep running
        self._check(self.lib.p2b_commit_blocks_absorb(h, cols.data_ptr(), col0, ncols))

    def commit_finish(self, h):
        self._check(self.lib.p2b_commit_blocks_finish(h))
        return self._PB(self.ctx, h.value)

    def export_nodes(self, batch, layer, first, count):
        out = self.torch.empty((count, 4), dtype=self.torch.int64, device="cuda")
        self._sync_in(out)
        self._check(self.lib.p2b_batch_export_nodes(batch.handle, layer, first, count, out.data_ptr()))
        self._sync_out()
        return out

    def import_nodes(self, batch, layer, first, count, t):
        self._sync_in(t)
        self._check(self.lib.p2b_batch_import_nodes(batch.handle, layer, first, count, t.data_ptr()))

    def finish_layers(self, batch, from_layer):
        self._check(self.lib.p2b_batch_finish_layers(batch.handle, from_layer))
        self._sync_out()

    def cap(self, batch):
        return batch.cap()

    def pack_open_rows(self, batch, idx, slots, Q, leaf_len, layers):
        packed = np.zeros((Q, leaf_len + 4 * layers), dtype=np.uint64)
        if idx:
            rows, sibs = batch.open_rows(idx)      # open_rows_kernel gathers rows + paths on the device
            packed[slots, :leaf_len] = rows
            packed[slots, leaf_len:] = sibs.reshape(len(idx), -1)
        return self.torch.from_numpy(packed.view(np.int64)).cuda()

    def to_numpy(self, t):
        return t.cpu().numpy().view(np.uint64)


class TorchComm:
    """torch.distributed (NCCL on GPUs, gloo on CPU) behind the two collectives the path needs."""

    def __init__(self, dist=None):
        import torch
        import torch.distributed as d
        self.torch, self.dist = torch, dist or d
        self.rank = self.dist.get_rank() if self.dist.is_initialized() else 0
        self.world = self.dist.get_world_size() if self.dist.is_initialized() else 1

    def _gather(self, t):
        if self.world == 1:
            return t
        out = self.torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def all_gather_columns(self, t, cmax, n_log):
        return self._gather(t)

    def all_gather_nodes(self, t, count):
        return self._gather(t)

    def all_gather_async(self, t):
        """Starts the all-gather of `t` (same shape on every rank); returns a handle for wait()."""
        if self.world == 1:
            return (t, None)
        out = self.torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        work = self.dist.all_gather_into_tensor(out, t.contiguous(), async_op=True)
        return (out, work)

    def wait(self, handle):
        out, work = handle
        if work is not None:
            work.wait()
        return out

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t
