"""The N>1 path on CPU: two processes over gloo.  Checks the chain sharding (no collective on the sampling path), that a
chain's results do not depend on which rank runs it (the per-rank "sampler" here is the CPU oracle -- there is no GPU in
this box), the max-over-ranks timing reduction, and the warmup-only pooled-statistics all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rainier_b200 import dist as rdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.rainier_py.binding import OracleModel, default_config
    from rainier_b200 import abi

    rir = open(os.path.join(ROOT, "rainier_b200", "models", "funnel10.rir"), "rb").read()
    cfg = default_config()
    cfg.sampler, cfg.n_steps = abi.RN_SAMPLER_HMC, 3
    cfg.mass_tuner = abi.RN_MASS_IDENTITY
    cfg.warmup_iterations, cfg.iterations = 40, 10
    seeds = np.arange(total) + 100
    mine = rdist.seeds_for_rank(seeds, rank, world)
    local = OracleModel(rir, []).sample(cfg, seeds=mine)["samples"]
    full = rdist.gather_samples(local, total)
    # pooled window statistics: every rank contributes its chains' sums
    q = local[:, -1, :]  # last draw of every local chain
    stats = torch.tensor(np.concatenate([[q.shape[0]], q.sum(axis=0), (q * q).sum(axis=0)]), dtype=torch.float64)
    rdist.allreduce_window_stats(stats)
    tmax = rdist.max_over_ranks(1.0 + rank)
    # multi-start MAP: starts sharded like chains, the best converged start is the job's answer on every rank
    from oracle.rainier_py.optimizer import lbfgs
    srir = open(os.path.join(ROOT, "rainier_b200", "models", "eight_schools.rir"), "rb").read()
    om = OracleModel(srir, [])
    starts = np.random.default_rng(5).normal(size=(total, 10)) * 0.7
    lo, hi = rdist.chain_block(total, rank, world)
    res = [lbfgs(om.density_batch, 10, x0=x, max_evals=300) for x in starts[lo:hi]]
    xb, fb, owner = rdist.best_start([r["x"] for r in res], [r["f"] for r in res], [r["info"] for r in res])
    if rank == 0:
        np.save(os.path.join(out_dir, "best.npy"), np.concatenate([xb, [fb, owner]]))
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full)
        np.save(os.path.join(out_dir, "stats.npy"), stats.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), np.array([tmax]))
    dist.barrier()
    dist.destroy_process_group()


def test_chain_blocks_partition_the_job():
    for total in (1, 7, 8, 1000, 4097):
        for world in (1, 2, 3, 8):
            blocks = [rdist.chain_block(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo(tmp_path):
    from oracle.rainier_py.binding import OracleModel, default_config
    from rainier_b200 import abi

    total, world = 13, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    # the same job in one process
    rir = open(os.path.join(ROOT, "rainier_b200", "models", "funnel10.rir"), "rb").read()
    cfg = default_config()
    cfg.sampler, cfg.n_steps = abi.RN_SAMPLER_HMC, 3
    cfg.mass_tuner = abi.RN_MASS_IDENTITY
    cfg.warmup_iterations, cfg.iterations = 40, 10
    ref = OracleModel(rir, []).sample(cfg, seeds=np.arange(total) + 100)["samples"]
    assert np.array_equal(full, ref), "a chain's samples must not depend on the rank layout"
    stats = np.load(tmp_path / "stats.npy")
    q = ref[:, -1, :]
    assert stats[0] == total
    assert np.allclose(stats[1:11], q.sum(axis=0), rtol=1e-13) and np.allclose(stats[11:], (q * q).sum(axis=0), rtol=1e-13)
    var = rdist.pooled_variance(stats[0], stats[1:11], stats[11:])
    assert np.allclose(var, q.var(axis=0), rtol=1e-9)
    # what the library does at a window end: per-chain Welford statistics combined around the pooled mean
    draws = ref[:, -6:, :]  # a window of 6 draws per chain
    w = rdist.combine_welford(6, draws.mean(axis=1), ((draws - draws.mean(axis=1, keepdims=True)) ** 2).sum(axis=1))
    assert np.allclose(w, draws.reshape(-1, 10).var(axis=0), rtol=1e-12)
    shifted = draws + 1e9  # |mean| >> sd: the sums-of-squares form loses everything, the combined form nothing
    w2 = rdist.combine_welford(6, shifted.mean(axis=1), ((shifted - shifted.mean(axis=1, keepdims=True)) ** 2).sum(axis=1))
    assert np.allclose(w2, w, rtol=1e-6)
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
    # multi-start MAP over 2 ranks == the same starts in one process
    from oracle.rainier_py.optimizer import lbfgs
    om = OracleModel(open(os.path.join(ROOT, "rainier_b200", "models", "eight_schools.rir"), "rb").read(), [])
    starts = np.random.default_rng(5).normal(size=(total, 10)) * 0.7
    res = [lbfgs(om.density_batch, 10, x0=x, max_evals=300) for x in starts]
    xb, fb, _ = rdist.best_start([r["x"] for r in res], [r["f"] for r in res], [r["info"] for r in res])
    best = np.load(tmp_path / "best.npy")
    assert np.array_equal(best[:10], xb) and best[10] == fb
    lo, hi = rdist.chain_block(total, int(best[11]), world)
    assert any(np.array_equal(r["x"], xb) for r in res[lo:hi])
