"""The N>1 path on CPU: world_size-2 gloo processes each evaluate their shard of families (through the
emulation build, there is no GPU here) and all-reduce the (1+P) result; it must equal the full-batch golden."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT, load_golden
from whale_jl_b200 import dist as wdist


def test_lpt_shards_partition_and_balance():
    rng = np.random.default_rng(0)
    work = rng.lognormal(3, 1, 1000)
    parts = [wdist.shard(work, r, 8) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(1000))
    loads = np.array([work[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.02


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        import numpy as np, torch, torch.distributed as dist
        from conftest import load_golden, golden_model, golden_fams
        from whale_jl_b200 import lib as wlib, dist as wdist
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        L = wlib.Lib({os.path.join(ROOT, 'tests', 'emu', 'libwhalecuda_emu.so')!r})
        g = load_golden("const_wgdturing")
        F = len(g["f_clade_off"]) - 1
        work = np.diff(g["f_split_off"][g["f_clade_off"]]).astype(float)
        mine = wdist.shard(work, rank, world)
        mh = L.model_create(golden_model(g))
        dh = L.data_create(mh, golden_fams(g, mine))
        ll, grad, _, _ = L.logpdf_grad(mh, dh, g["xs"][1], g["m_pleaf"], 1, want_grad=True)
        t = torch.tensor([ll, *grad], dtype=torch.float64)
        wdist.allreduce_sum(t)
        if rank == 0:
            want = np.concatenate([[g["tot_root"][1]], g["grad_root"][1]])
            assert np.allclose(t.numpy(), want, rtol=1e-9, atol=1e-9 * np.abs(want).max()), (t.numpy(), want)
            print("OK", len(mine), F)
        dist.destroy_process_group()
    """))
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
