"""Row-sharded multi-GPU driver: one process per GPU, torch.distributed for the plumbing.

SURVEY.md 8(e): every output row depends on one row of A (or H) and all of B (Q, H^T), so the path shards by rows
with one broadcast in and one gather out.  The reference's analogue is `limits` (src/workdivision.cpp:16-89), an
even split by row COUNT across OpenMP threads; here the split is by the per-row cost from the flop-counting pass
(spgemm_b200_row_costs / spgemm_b200_partition), which matters on power-law inputs and for the upper-triangle
triple product whose rows get cheaper towards the bottom.

    rank 0 holds the host operands
      -> dist.broadcast of the CSR arrays (NCCL over NVLink on GPUs, gloo in the CPU tests)
      -> every rank: partition (identical on all ranks), compute its row block with the CUDA library
      -> gather of the row blocks to rank 0 (dense: contiguous slabs; sparse: variable-size indices/values)

The compute callback is injected so that the host-side logic (partition, broadcast packing, gather offsets) is
testable on CPU with the gloo backend and the oracle as the per-rank compute (tests/test_distributed_cpu.py).
"""
import os

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------------
# partition (host mirror of spgemm_b200_partition, used where costs are already on the host)
def partition_by_cost(costs, parts):
    """Contiguous row bounds (len parts+1) so that every part carries ~sum(costs+1)/parts."""
    c = np.asarray(costs, dtype=np.float64) + 1.0
    total = c.sum()
    cum = np.cumsum(c)
    bounds = np.zeros(parts + 1, dtype=np.int64)
    for p in range(1, parts):
        bounds[p] = int(np.searchsorted(cum, total * p / parts, side="left")) + 1
    bounds[parts] = len(c)
    bounds = np.minimum(np.maximum.accumulate(bounds), len(c))
    return bounds


def host_row_costs(a, b, kind, upper_only):
    """Per-row cost on the host (numpy): products of A rows against B; for the triple product
    a P1_i + (b + c) P2_i (n - i)/n with (a, b, c) = (4, 1, 2): the one-panel case of k_triple_costs
    (csrc/analysis.cu; the CPU tests only need a deterministic, monotone stand-in)."""
    blen = np.diff(b.indptr).astype(np.int64)
    rows = np.repeat(np.arange(a.shape[0]), np.diff(a.indptr))
    p1 = np.bincount(rows, weights=blen[a.indices], minlength=a.shape[0]).astype(np.float64)
    if kind != "triple":
        return p1
    ht_len = np.bincount(a.indices, minlength=a.shape[1]).astype(np.int64)      # nnz of rows of H^T
    # sum over (i,j) in H of sum_{c in Q_j} nnz(H^T_c): cost of row j of Q first, then gather by H's columns
    # (cumulative sum differenced at the row pointers: np.add.reduceat rejects the index nnz that trailing empty
    #  rows of Q put into indptr[:-1])
    cum = np.concatenate([[0], np.cumsum(ht_len[b.indices])])
    qcost = (cum[b.indptr[1:]] - cum[b.indptr[:-1]]).astype(np.float64)
    p2 = np.bincount(rows, weights=qcost[a.indices], minlength=a.shape[0])
    keep = 1.0
    if upper_only:
        n = a.shape[0]
        keep = (n - np.arange(n)) / max(1, n)
    return 4.0 * p1 + 3.0 * p2 * keep


# ------------------------------------------------------------------------------------------------------
# broadcast of a CSR matrix from rank 0: one 24-byte header + ONE packed byte buffer per matrix
def _packed_layout(rows, nnz):
    """Byte offsets of indptr | indices | data inside the packed buffer (data 8-byte aligned) and its size."""
    o_ptr = 0
    o_idx = o_ptr + 4 * (rows + 1)
    o_val = (o_idx + 4 * nnz + 7) & ~7
    return o_ptr, o_idx, o_val, o_val + 8 * nnz


def broadcast_csr(x, device):
    """rank 0 passes a scipy CSR, the others None; returns (shape, indptr, indices, data) tensors on `device`.
    Two collectives per matrix: the header (rows, cols, nnz) and the three arrays packed into one byte buffer
    (round 1 sent a header and three arrays and synchronised the host in between)."""
    meta = torch.zeros(3, dtype=torch.int64)
    if dist.get_rank() == 0:
        meta[:] = torch.tensor([x.shape[0], x.shape[1], x.nnz], dtype=torch.int64)
    meta = meta.to(device)
    dist.broadcast(meta, src=0)
    rows, cols, nnz = (int(v) for v in meta.tolist())
    o_ptr, o_idx, o_val, size = _packed_layout(rows, nnz)
    buf = torch.empty(max(size, 8), dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:                     # three copies straight into their slices of the packed buffer
        buf[o_ptr:o_idx].view(torch.int32).copy_(torch.from_numpy(np.ascontiguousarray(x.indptr, dtype=np.int32)))
        if nnz:
            buf[o_idx:o_idx + 4 * nnz].view(torch.int32).copy_(torch.from_numpy(np.ascontiguousarray(x.indices, dtype=np.int32)))
            buf[o_val:o_val + 8 * nnz].view(torch.float64).copy_(torch.from_numpy(np.ascontiguousarray(x.data, dtype=np.float64)))
    dist.broadcast(buf, src=0)
    indptr = buf[o_ptr:o_idx].view(torch.int32)
    indices = buf[o_idx:o_idx + 4 * nnz].view(torch.int32)
    data = buf[o_val:o_val + 8 * nnz].view(torch.float64)
    return (rows, cols), indptr, indices, data


def tensors_to_scipy(shape, indptr, indices, data):
    m = sp.csr_matrix(shape)
    m.indptr, m.indices, m.data = indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy()
    return m


# ------------------------------------------------------------------------------------------------------
# gathers to rank 0
def gather_dense_rows(local, bounds, ncols, device):
    """local: (rows_of_this_rank, ncols) float64 tensor on `device`.  Rank 0 returns the stacked matrix."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        full = torch.empty((int(bounds[-1]), ncols), dtype=torch.float64, device=device)
        full[int(bounds[0]):int(bounds[1])] = local
        ops = [dist.P2POp(dist.irecv, full[int(bounds[r]):int(bounds[r + 1])], r) for r in range(1, world)
               if bounds[r + 1] > bounds[r]]
        if ops:                                   # one grouped launch (ncclGroupStart/End), not serialized recvs
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return full
    if local.numel():
        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), 0)]):
            q.wait()
    return None


def gather_csr_rows(local_indptr, local_indices, local_data, bounds, device):
    """Every rank passes its block's LOCAL indptr (int64, starts at 0), indices, data.  Rank 0 returns the
    stitched (indptr int64, indices int32, data float64) -- the parallel replacement of the reference's serial
    stitch (src/sparse_sparse_sparse.cpp:265-291)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    nnz_local = torch.tensor([int(local_indices.numel())], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, nnz_local)
    counts = [int(c.item()) for c in counts]
    offs = np.concatenate([[0], np.cumsum(counts)])
    if rank == 0:
        rows = int(bounds[-1])
        indptr = torch.zeros(rows + 1, dtype=torch.int64, device=device)
        indices = torch.empty(int(offs[-1]), dtype=torch.int32, device=device)
        data = torch.empty(int(offs[-1]), dtype=torch.float64, device=device)
        indptr[int(bounds[0]):int(bounds[1]) + 1] = local_indptr
        indices[:counts[0]] = local_indices
        data[:counts[0]] = local_data
        ops, ptr_bufs = [], {}
        for r in range(1, world):
            r0, r1 = int(bounds[r]), int(bounds[r + 1])
            if r1 > r0:
                ptr_bufs[r] = torch.empty(r1 - r0 + 1, dtype=torch.int64, device=device)
                ops.append(dist.P2POp(dist.irecv, ptr_bufs[r], r))
            if counts[r]:
                ops.append(dist.P2POp(dist.irecv, indices[int(offs[r]):int(offs[r + 1])], r))
                ops.append(dist.P2POp(dist.irecv, data[int(offs[r]):int(offs[r + 1])], r))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        for r, buf in ptr_bufs.items():
            indptr[int(bounds[r]) + 1:int(bounds[r + 1]) + 1] = buf[1:] + int(offs[r])
        return indptr, indices, data
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    ops = []
    if r1 > r0:
        ops.append(dist.P2POp(dist.isend, local_indptr.contiguous(), 0))
    if counts[rank]:
        ops.append(dist.P2POp(dist.isend, local_indices.contiguous(), 0))
        ops.append(dist.P2POp(dist.isend, local_data.contiguous(), 0))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return None


# ------------------------------------------------------------------------------------------------------
def host_partition(a_t, b_t, kind, upper_only, world):
    """Partition from the broadcast operands with numpy (CPU tests; the GPU path uses cuda_partition)."""
    a_h, b_h = tensors_to_scipy(*a_t), tensors_to_scipy(*b_t)
    return partition_by_cost(host_row_costs(a_h, b_h, kind, upper_only), world)


def multiply_sharded(matrix_a, matrix_b, kind, upper_only, device, compute_block, partition=host_partition):
    """The whole sharded product.  rank 0 passes scipy CSR operands (others None).

    kind          : "sparse" | "dense" | "triple"
    compute_block : f(a_tensors, b_tensors, r0, r1) -> dense (r1-r0, ncols) float64 tensor on `device`, or for
                    kind == "sparse" a tuple (indptr int64[r1-r0+1], indices int32, data float64) of tensors.
                    a_tensors / b_tensors are (shape, indptr, indices, data).
    Returns on rank 0 the full result (dense tensor, or (indptr, indices, data) tensors); None elsewhere.
    """
    world = dist.get_world_size()
    a_t = broadcast_csr(matrix_a, device)
    b_t = broadcast_csr(matrix_b, device)
    # identical partition on every rank: the cost pass runs on the broadcast operands (deterministic integers)
    bounds = partition(a_t, b_t, kind, upper_only, world)
    r0, r1 = int(bounds[dist.get_rank()]), int(bounds[dist.get_rank() + 1])
    block = compute_block(a_t, b_t, r0, r1)
    if kind == "sparse":
        return gather_csr_rows(block[0], block[1], block[2], bounds, device)
    ncols = a_t[0][0] if kind == "triple" else b_t[0][1]
    return gather_dense_rows(block, bounds, ncols, device)


# ------------------------------------------------------------------------------------------------------
# GPU compute callback + the end-to-end leg of bench.py at N > 1
def cuda_compute_block(kind, upper_only):
    """compute_block for multiply_sharded that runs the CUDA library on this rank's GPU, borrowing the broadcast
    tensors (no copies) and launching on torch's current stream so that NCCL and the kernels are ordered."""
    from . import device as dev

    def run(a_t, b_t, r0, r1):
        # same stream as torch (ordered with the broadcasts before and the gather after) -- and, belt and braces,
        # a host-side wait on both ends: the operands must have landed, the block must be complete
        torch.cuda.current_stream().synchronize()
        dev.set_stream(torch.cuda.current_stream().cuda_stream)
        (ash, ap, ai, av), (bsh, bp, bi, bv) = a_t, b_t
        A, B = _wrap(dev, a_t), _wrap(dev, b_t)
        if kind == "dense":
            out = torch.empty((r1 - r0, bsh[1]), dtype=torch.float64, device=ap.device)
            dev.spgemm_dense(A, B, upper_only, r0, r1, out=out.data_ptr())
            dev.synchronize()
            return out
        if kind == "triple":
            out = torch.empty((r1 - r0, ash[0]), dtype=torch.float64, device=ap.device)
            dev.triple_product(A, B, None, upper_only, r0, r1, out=out.data_ptr())
            dev.synchronize()
            return out
        res = dev.spgemm_csr(A, B, upper_only, r0, r1)
        p, i, v = res.device_ptrs()
        nnz, rows = res.nnz, r1 - r0
        # copy out of the library's pool into torch tensors (device to device) before the handle is freed
        indptr = torch.empty(rows + 1, dtype=torch.int64, device=ap.device)
        indices = torch.empty(nnz, dtype=torch.int32, device=ap.device)
        data = torch.empty(nnz, dtype=torch.float64, device=ap.device)
        for dst, src, nbytes in ((indptr, p, (rows + 1) * 8), (indices, i, nnz * 4), (data, v, nnz * 8)):
            dev.copy_on_device(dst.data_ptr(), src, nbytes)
        dev.synchronize()
        res.free()
        return indptr, indices, data

    return run


def _wrap(dev, t):
    shape, p, i, v = t
    return dev.DeviceMatrix.wrap(shape, int(i.numel()), p.data_ptr(), i.data_ptr(), v.data_ptr(), keep=t)


def cuda_partition(a_t, b_t, kind, upper_only, world):
    """Flop-balanced bounds from the GPU cost pass (spgemm_b200_row_costs + spgemm_b200_partition)."""
    from . import device as dev
    from .matrix_ops import matrix_ops
    torch.cuda.current_stream().synchronize()            # the broadcast operands have landed
    dev.set_stream(torch.cuda.current_stream().cuda_stream)
    A, B = _wrap(dev, a_t), _wrap(dev, b_t)
    if kind == "triple":
        Ht = A.transpose()
        costs, _ = dev.row_costs(A, Ht, B, upper_only)
    else:
        costs, _ = dev.row_costs(A, B, None, upper_only, dense_cols=(b_t[0][1] if kind == "dense" else 0))
    bounds = dev.partition_rows(costs, a_t[0][0], world,
                                tail_indptr=a_t[1].cpu().numpy() if kind == "triple" else None)
    matrix_ops.get_lib().spgemm_b200_device_free(costs)
    return bounds.astype(np.int64)


# ------------------------------------------------------------------------------------------------------
# fused compute + gather: every rank's kernel writes its rows straight into rank 0's buffer over NVLink
_peer_cache = {}


def multiply_sharded_peer(matrix_a, matrix_b, kind, upper_only, device):
    """Dense / triple product with the gather fused into the compute kernels: rank 0 owns an IPC-exported result
    buffer, the other ranks map it, and the row-range kernels of every rank store (and reduce) directly into
    their row slab of it -- peer stores over NVLink, no staging buffer and no separate gather collective.
    Returns a device.SharedDense on rank 0 (valid until the next call with another shape), None elsewhere."""
    from . import device as dev
    assert kind in ("dense", "triple")
    rank, world = dist.get_rank(), dist.get_world_size()
    a_t = broadcast_csr(matrix_a, device)
    b_t = broadcast_csr(matrix_b, device)
    rows = a_t[0][0]
    ncols = rows if kind == "triple" else b_t[0][1]
    key = (rows, ncols)
    if key not in _peer_cache:                       # IPC mapping is expensive: once per result shape
        for old in _peer_cache.values():
            old.close()
        _peer_cache.clear()
        handle = torch.zeros(64, dtype=torch.uint8, device=device)
        if rank == 0:
            buf = dev.SharedDense(rows, ncols)
            handle.copy_(torch.frombuffer(bytearray(buf.export()), dtype=torch.uint8))
        dist.broadcast(handle, src=0)
        if rank != 0:
            buf = dev.SharedDense.open(bytes(handle.cpu().numpy().tobytes()), rows, ncols)
        _peer_cache[key] = buf
    buf = _peer_cache[key]
    bounds = cuda_partition(a_t, b_t, kind, upper_only, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    torch.cuda.current_stream().synchronize()
    dev.set_stream(torch.cuda.current_stream().cuda_stream)
    A, B = _wrap(dev, a_t), _wrap(dev, b_t)
    if r1 > r0:
        if kind == "dense":
            dev.spgemm_dense(A, B, upper_only, r0, r1, out=buf.row_ptr(r0))
        else:
            dev.triple_product(A, B, None, upper_only, r0, r1, out=buf.row_ptr(r0))
    dev.synchronize()                                # this rank's peer stores are complete
    dist.barrier()                                   # ... and so are everybody else's
    return buf if rank == 0 else None


def bench_e2e(args, w, flops, rank, world, csr_bytes):
    """e2e at N GPUs: rank 0 starts from HOST operands and ends with a HOST result; every step does
    H2D on rank 0 (inside broadcast_csr) -> NCCL broadcast -> sharded compute -> gather to rank 0 -> D2H."""
    import time
    device = torch.device("cuda", torch.cuda.current_device())
    kind = w["kind"]
    upper = kind == "triple" or bool(w["kwargs"].get("symmetric"))
    a = w["a"] if rank == 0 else None
    b = w["b"] if rank == 0 else None
    from . import device as dev
    from .matrix_ops import _result_array
    fn = cuda_compute_block(kind, upper)
    times, d2h = [], 0
    for it in range(2 + args.steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if kind in ("dense", "triple") and os.environ.get("SPGEMM_B200_GATHER", "peer") == "peer":
            out = multiply_sharded_peer(a, b, kind, upper, device)          # gather fused into the kernels
            gather = "peer stores over NVLink from inside the compute kernels (CUDA IPC)"
            if rank == 0:
                h = _result_array(out.shape, np.float64)
                if upper and out.shape[0] == out.shape[1]:
                    dev.copy_upper_to_host(h, out.ptr)
                else:
                    dev.copy_to_host(h, out.ptr)
                d2h = int(dev.last_stats()["bytes_d2h"]) or h.nbytes
                del h
        else:
            out = multiply_sharded(a, b, kind, upper, device, fn, partition=cuda_partition)
            gather = "grouped NCCL isend/irecv of row blocks to rank 0"
            if rank == 0:
                # device -> pinned host arrays from the library's cache (what the single-GPU API returns too)
                outs = out if kind == "sparse" else (out,)
                d2h, host = 0, []
                for t in outs:
                    h = _result_array(tuple(t.shape), {torch.int64: np.int64, torch.int32: np.int32,
                                                       torch.float64: np.float64}[t.dtype])
                    if kind != "sparse" and upper and t.shape[0] == t.shape[1]:
                        dev.copy_upper_to_host(h, t.data_ptr())
                    else:
                        dev.copy_to_host(h, t.data_ptr())
                    host.append(h)
                    d2h += h.nbytes
                del host
        torch.cuda.synchronize()
        dist.barrier()
        if it >= 2:
            times.append(time.perf_counter() - t0)
        del out
    t = torch.tensor([float(np.mean(times))], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    if rank != 0:
        return None
    return {"value": flops / sec / 1e9, "unit": "GFLOP/s", "ms_per_step": sec * 1e3,
            "h2d_bytes_per_step": int(csr_bytes(w["a"]) + csr_bytes(w["b"])), "d2h_bytes_per_step": int(d2h),
            "gather": gather,
            "timing": "host wall clock on rank 0 incl. H2D, NCCL broadcast, sharded kernels, gather to rank 0, D2H"}
