# per-call latency of AOpticsManager::TraceNonSequential on small batches through the mirror classes (the regime of
# tutorials/HexOkumuraCone.C: 80 000 calls of 1000 rays, and of the MINUIT loops of Optimize.C)
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import robast_b200 as R
from robast_b200 import configs
for name, (mgr, keep) in (("okumura_pgon", configs.okumura_cone("pgon")), ("davies_cotton", configs.davies_cotton())):
    for n in (1000, 10000, 100000):
        side = 10. if name.startswith("okumura") else 1400.
        tr = R.TGeoTranslation("t", 0, 0, 5. if name.startswith("okumura") else 3200.)
        for depth in (-1, 0):
            mgr.SetHistoryDepth(depth)
            def one():
                rays = R.ARayShooter.RandomSquare(400e-7, side, n, None, tr, R.TVector3(0, 0, -1))
                t0 = time.perf_counter()
                mgr.TraceNonSequential(rays)
                return time.perf_counter() - t0, rays
            for _ in range(3): one()
            reps = 30
            tt = sorted(one()[0] for _ in range(reps))
            print("%-14s n=%6d history=%2d  median %.0f us per call  (%.2e rays/s)" % (name, n, depth, 1e6 * tt[reps // 2], n / tt[reps // 2]), flush=True)
