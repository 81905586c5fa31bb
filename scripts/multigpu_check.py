"""Run under torchrun with N >= 2 GPUs: the row-sharded multi-GPU paths against the oracle on reduced configs.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multigpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

from helpers import assert_csr_equal, assert_dense_equal
from oracle import port
from sparse_matrix_mult_b200 import device as dev
from sparse_matrix_mult_b200 import distributed as sd
from sparse_matrix_mult_b200 import synthetic


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev.init(local)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    device = torch.device("cuda", local)
    for name in ("cfg2s", "cfg3s", "cfg1s", "cfg4r10"):
        w = synthetic.workload(name)
        kind = w["kind"]
        upper = kind == "triple" or bool(w["kwargs"].get("symmetric"))
        a = w["a"] if rank == 0 else None
        b = w["b"] if rank == 0 else None
        out = sd.multiply_sharded(a, b, kind, upper, device, sd.cuda_compute_block(kind, upper), partition=sd.cuda_partition)
        if rank == 0:
            want = port.sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
            if kind == "sparse":
                got = sp.csr_matrix(want.shape)
                got.indptr, got.indices, got.data = (t.cpu().numpy() for t in out)
                assert_csr_equal(got, want, name + " nccl gather")
            else:
                assert_dense_equal(out.cpu().numpy(), want, name + " nccl gather")
            print(name, "nccl gather ok", flush=True)
        if kind != "sparse":
            buf = sd.multiply_sharded_peer(a, b, kind, upper, device)
            if rank == 0:
                host = np.empty(buf.shape)
                dev.copy_to_host(host, buf.ptr)
                assert_dense_equal(host, want, name + " peer gather")
                print(name, "peer (fused) gather ok", flush=True)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multigpu check ok")


if __name__ == "__main__":
    main()
