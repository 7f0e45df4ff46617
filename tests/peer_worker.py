"""Worker of test_peer_sum_between_processes: one "rank" = one process holding a share of the C1 families; the ranks
exchange CUDA IPC handles through files and every evaluation flagged WHALE_PEER_SUM must return the full-batch golden
log-likelihood and gradient on every rank (src/core.jl:54,63: the families are i.i.d. terms of a sum).
usage: peer_worker.py <rank> <world> <dir>"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    from conftest import load_golden, golden_model, golden_fams
    from whale_jl_b200 import lib as wlib
    L = wlib.Lib()
    ndev = L.L.whale_device_count()
    assert ndev > 0
    L.check(L.L.whale_set_device(rank % ndev))
    g = load_golden("c1_example1")
    n = len(g["ll_fam"][0])
    mine = [f for f in range(n) if f % world == rank]
    mh = L.model_create(golden_model(g))
    dh = L.data_create(mh, golden_fams(g, mine))
    h = L.peer_export(dh, rank, world)
    with open(os.path.join(d, f"h{rank}.tmp"), "wb") as fh:
        fh.write(h)
    os.rename(os.path.join(d, f"h{rank}.tmp"), os.path.join(d, f"h{rank}.bin"))
    for q in range(world):
        if q == rank:
            continue
        p = os.path.join(d, f"h{q}.bin")
        t0 = time.time()
        while not os.path.exists(p):
            assert time.time() - t0 < 120, "peer handle did not appear"
            time.sleep(0.01)
        L.peer_import(dh, q, open(p, "rb").read())
    assert L.L.whale_peer_ready(dh) == 1
    import ctypes as C
    f64p = C.POINTER(C.c_double)
    pl = np.ascontiguousarray(g["m_pleaf"])
    for rep in range(2):
        for xi, x in enumerate(g["xs"]):
            for want_grad in (True, False):
                xx = np.ascontiguousarray(x)
                ll = C.c_double()
                grad = np.zeros(len(x))
                flags = (wlib.WANT_GRAD if want_grad else 0) | wlib.PEER_SUM
                L.check(L.L.whale_logpdf_grad(mh, dh, xx.ctypes.data_as(f64p), pl.ctypes.data_as(f64p), 1, flags, C.byref(ll),
                                              grad.ctypes.data_as(f64p) if want_grad else None, None, None))
                want = g["tot_root"][xi]
                assert abs(ll.value - want) <= 1e-9 * abs(want), (rank, xi, ll.value, want)
                if want_grad:
                    wg = g["grad_root"][xi]
                    assert np.allclose(grad, wg, rtol=1e-9, atol=1e-9 * np.abs(wg).max()), (rank, xi)
                with open(os.path.join(d, f"ll_r{rank}_{rep}_{xi}_{int(want_grad)}.txt"), "w") as fh:
                    fh.write(ll.value.hex())
    # the exchange alone (whale_peer_sum_async) on a caller-owned device buffer: 1 + P doubles, here (rank + 1) * (j + 1)
    import torch
    torch.cuda.set_device(rank % ndev)
    P = len(g["xs"][0])
    buf = (rank + 1) * torch.arange(1, 2 + P, device="cuda", dtype=torch.float64)
    L.check(L.L.whale_peer_sum_async(dh, buf.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = sum(r + 1 for r in range(world)) * np.arange(1, 2 + P, dtype=np.float64)
    assert np.array_equal(buf.cpu().numpy(), want), (rank, buf.cpu().numpy()[:4], want[:4])
    L.L.whale_data_destroy(dh)
    L.L.whale_model_destroy(mh)
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
