"""CPU tests (gloo, world_size 2) of the multi-GPU scheme: all-gather of particle-record slices, replicated tree, sliced
targets.  Each rank runs the CPU oracle restatement on the gathered set for ITS slice only; the union of the slices must be
bit-identical to a single-rank run."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    ROOT = sys.argv[1]
    for p in (os.path.join(ROOT, "gadget-2.0.7-ngravs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import g2test, multigpu
    from portrun import PortOracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 6001
    pos, mass, ptype = g2test.hernquist(n, seed=3)
    lo, hi, per = multigpu.owner_slice(n, rank, world)
    ex = multigpu.ParticleExchange(n, torch.device("cpu"), world)
    full = multigpu.pack_records(torch.from_numpy(pos), torch.from_numpy(mass), torch.from_numpy(ptype.astype(np.int32)))
    ex.set_local(full[lo:hi])
    g_rec = ex.gather()
    assert torch.equal(g_rec, full)
    g_pm, g_type = g_rec[:, :4], g_rec.view(torch.int32)[:, 4].contiguous()
    # every rank: same tree from the gathered set, walk only its slice of the targets
    o = PortOracle(int(1.1 * n) + 64, softening=g2test.SOFT_NP, gravity=g2test.GRAV_D2)
    o.load(g_pm[:, :3].numpy(), g_pm[:, 3].numpy(), g_type.numpy())
    o.domain()
    nn = o.treebuild()
    tlo, thi = multigpu.slice_bounds(n, rank, world)
    dt, cost = o.walk_threads(1, tlo, thi)
    p = o.particles()
    acc = torch.zeros((n, 3), dtype=torch.float64)
    acc[tlo:thi] = torch.from_numpy(p["accd"][tlo:thi])
    cnt = torch.zeros(n, dtype=torch.float64)
    cnt[tlo:thi] = torch.from_numpy(p["cost"][tlo:thi].astype(np.float64))
    # the tree potential shards the same way (every particle is a target, potential.c:86): a rank walks its slice only
    o.set_potential_laws("newtonian", "plummer")
    o.set_opening(0.5, 0.005)
    pot = torch.zeros(n, dtype=torch.float64)
    pot[tlo:thi] = torch.from_numpy(o.potential_targets(np.arange(tlo, thi, dtype=np.int32)).astype(np.float64))
    dist.all_reduce(acc)       # slices are disjoint: the sum is a concatenation
    dist.all_reduce(cnt)
    dist.all_reduce(pot)
    nodes = torch.tensor([nn]); allnodes = [torch.zeros_like(nodes) for _ in range(world)]
    dist.all_gather(allnodes, nodes)
    assert all(int(x) == nn for x in allnodes)
    if rank == 0:
        np.save(sys.argv[2], np.concatenate([acc.numpy(), cnt.numpy()[:, None], pot.numpy()[:, None]], axis=1))
    dist.destroy_process_group()
''')


def test_two_ranks_reproduce_single_rank(tmp_path):
    sys.path[:0] = [os.path.join(ROOT, "gadget-2.0.7-ngravs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import g2test
    import multigpu
    from portrun import PortOracle
    # slices tile the range exactly for awkward sizes
    for n in (1, 7, 6001, 16777216):
        for world in (1, 2, 3, 8):
            b = [multigpu.slice_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            ow = [multigpu.owner_slice(n, r, world) for r in range(world)]
            assert sum(h - l for l, h, _ in ow) == n
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "out.npy"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script), ROOT, str(out)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    got = np.load(out)
    n = 6001
    pos, mass, ptype = g2test.hernquist(n, seed=3)
    o = PortOracle(int(1.1 * n) + 64, softening=g2test.SOFT_NP, gravity=g2test.GRAV_D2)
    o.load(pos, mass, ptype)
    o.domain()
    o.treebuild()
    o.walk_threads(1)
    p = o.particles()
    assert np.array_equal(got[:, :3], p["accd"])
    assert np.array_equal(got[:, 3], p["cost"].astype(np.float64))
    o.set_potential_laws("newtonian", "plummer")
    assert np.array_equal(got[:, 4], o.potential(nthreads=2).astype(np.float64))
