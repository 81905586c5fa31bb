"""World-size-2 gloo tests of the multi-GPU host logic (partition, CSR broadcast, row-block gathers), with the
oracle as the per-rank compute.  Runs on CPU: `python -m pytest tests -m "not gpu"`."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(kind):
    rng = np.random.default_rng(11)
    if kind == "triple":
        h = sp.random(57, 300, density=0.05, format='csr', random_state=rng)
        offs = list(range(-3, 4))
        q = sp.diags([np.full(300 - abs(o), np.exp(-abs(o) / 2.0)) for o in offs], offs, format='csr')
        return h, sp.csr_matrix(q)
    a = sp.random(90, 70, density=0.08, format='csr', random_state=rng)
    b = sp.random(70, 90, density=0.08, format='csr', random_state=rng)
    return a, b


def _oracle_block(kind, upper):
    from oracle import port
    from sparse_matrix_mult_b200.distributed import tensors_to_scipy

    def run(a_t, b_t, r0, r1):
        a, b = tensors_to_scipy(*a_t), tensors_to_scipy(*b_t)
        if kind == "dense":
            full = port.spgemm_dense(a, b, upper)
            return torch.from_numpy(full[r0:r1].copy())
        if kind == "triple":
            full = port.triple_product(a, b, 0)
            return torch.from_numpy(full[r0:r1].copy())
        c = port.spgemm_csr(a, b, upper)
        c.sort_indices()
        blk = c[r0:r1]
        return (torch.from_numpy(blk.indptr.astype(np.int64)), torch.from_numpy(blk.indices.astype(np.int32)),
                torch.from_numpy(blk.data.astype(np.float64)))
    return run


def _worker(rank, world, port, kind, upper, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparse_matrix_mult_b200 import distributed as sd
    a, b = _inputs(kind)
    res = sd.multiply_sharded(a if rank == 0 else None, b if rank == 0 else None, kind, upper,
                              torch.device("cpu"), _oracle_block(kind, upper))
    if rank == 0:
        if kind == "sparse":
            np.savez(out_path, indptr=res[0].numpy(), indices=res[1].numpy(), data=res[2].numpy())
        else:
            np.savez(out_path, dense=res.numpy())
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,upper,world", [("dense", False, 2), ("dense", True, 2), ("sparse", False, 2),
                                              ("sparse", True, 3), ("triple", True, 2)])
def test_sharded_product_matches_oracle(tmp_path, kind, upper, world):
    from oracle import port
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), kind, upper, out), nprocs=world, join=True)
    got = np.load(out)
    a, b = _inputs(kind)
    if kind == "dense":
        assert np.array_equal(got["dense"], port.spgemm_dense(a, b, upper))
    elif kind == "triple":
        assert np.array_equal(got["dense"], port.triple_product(a, b, 0))
    else:
        want = port.spgemm_csr(a, b, upper)
        want.sort_indices()
        assert np.array_equal(got["indptr"], want.indptr.astype(np.int64))
        assert np.array_equal(got["indices"], want.indices)
        assert np.array_equal(got["data"], want.data)


def test_partition_by_cost_balances_and_covers():
    from sparse_matrix_mult_b200.distributed import host_row_costs, partition_by_cost
    rng = np.random.default_rng(3)
    costs = rng.pareto(1.2, size=5000) * 100            # heavy tailed, like R-MAT row costs
    for parts in (1, 2, 4, 8):
        b = partition_by_cost(costs, parts)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0)
        loads = np.add.reduceat(costs + 1, b[:-1])[:parts] if parts > 1 else [costs.sum() + 5000]
        assert max(loads) <= (costs.sum() + 5000) / parts + costs.max() + 1
    # more parts than rows: trailing parts are empty, everything still covered
    b = partition_by_cost(np.ones(3), 8)
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0)
    # the reference's `limits` splits by row count; on uniform costs both agree up to rounding
    from oracle import port
    even = partition_by_cost(np.full(10, 7.0), 3)
    assert [(int(even[i]), int(even[i + 1]) - 1) for i in range(3)] in ([(0, 3), (4, 6), (7, 9)], [(0, 2), (3, 6), (7, 9)],
                                                                       [(0, 3), (4, 7), (8, 9)], [(0, 3), (4, 6), (7, 9)])
    assert port.limits(10, 3) == [(0, 3), (4, 6), (7, 9)]
    # triple-product cost model decreases with the row index when only the upper triangle is computed
    h = sp.random(200, 500, density=0.05, format='csr', random_state=rng)
    q = sp.identity(500, format='csr')
    c_up = host_row_costs(h, q, "triple", True)
    c_full = host_row_costs(h, q, "triple", False)
    assert c_up.sum() < c_full.sum() and c_up[:50].mean() > c_up[-50:].mean()
