"""N>1 host logic on CPU: world_size-2 gloo run of the ray sharding + terminal all-reduce.  The tracer itself
needs a GPU, so the per-rank trace is done by the CPU oracle here (checker role); what is tested is that
sharded + reduced results equal the unsharded ones and do not depend on the number of ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as H
    from robast_b200 import configs, sharding
    oracle = H.load_oracle()
    mgr, _k = configs.schmidt_cassegrain()  # stochastic config: exercises the global Philox ray ids
    ex = mgr.ExportScene()
    b, e = sharding.shard_range(n, rank, world)
    rays = H.make_rays(oracle, configs.beam(4, 0.05), b, e - b)
    H.trace_with(oracle.orc_trace, ex, rays, H.opts(seed=77, ray_id_offset=b))
    f = rays.status == 3
    hist, _, _ = np.histogram2d(rays.pos[f][:, 0], rays.pos[f][:, 1], bins=(20, 20), range=((-0.1, 0.1), (-0.8, 0.2)))
    t_hist = torch.from_numpy(hist.astype(np.int64))
    t_cnt = torch.from_numpy(np.bincount(rays.status, minlength=6).astype(np.int64))
    t_mom = torch.tensor([f.sum(), rays.pos[f][:, 0].sum(), rays.pos[f][:, 1].sum()], dtype=torch.float64)
    sharding.reduce_results([t_hist, t_cnt, t_mom], dist)
    if rank == 0:
        np.savez(os.path.join(out_dir, "w%d.npz" % world), hist=t_hist.numpy(), cnt=t_cnt.numpy(), mom=t_mom.numpy())
    # rank-ordered concatenation reproduces the single-rank order (mirrors the ordered Merge, src/AOpticsManager.cxx:559-562)
    gathered = [None] * world
    dist.all_gather_object(gathered, rays.status.tolist())
    if rank == 0:
        np.save(os.path.join(out_dir, "status_w%d.npy" % world), np.concatenate([np.array(g) for g in gathered]))
    dist.destroy_process_group()


def test_shard_range_matches_reference_split():
    from robast_b200.sharding import shard_range
    for n, w in ((10, 3), (100040004, 8), (7, 8), (16, 4)):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert all(r[k][1] - r[k][0] == n // w for k in range(w - 1))


@pytest.mark.timeout(600)
def test_two_rank_gloo_matches_single_rank(tmp_path):
    n = 6000
    for world, port in ((1, 29611), (2, 29612)):
        mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "w1.npz"), np.load(tmp_path / "w2.npz")
    assert (a["hist"] == b["hist"]).all() and (a["cnt"] == b["cnt"]).all() and a["cnt"].sum() == n
    assert np.allclose(a["mom"], b["mom"], rtol=1e-12)
    assert (np.load(tmp_path / "status_w1.npy") == np.load(tmp_path / "status_w2.npy")).all()
