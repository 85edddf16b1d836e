"""world_size-2 tests of the N>1 host logic on CPU (gloo): sharding of the column range, the MAX
reduction of the step time and the MIN/MAX/SUM reduction of the validation norms.  The per-rank
"kernel" here is the CPU oracle (the checker) -- the point is the plumbing around it, which bench.py
uses unchanged with the CUDA kernel on N GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ngptot, nproma, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import pyoracle
    from dwarf_p_cloudsc_b200 import params, shard, state, validate
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cols, scalars, _, klev = state.load_input()
        p, k = params.params_from_scalars(scalars)
        first, n = shard.weak_shard(rank, world, ngptot)
        F = state.blocked_fields(cols, n, nproma, klev, first_column=first)
        orc = pyoracle.best_oracle()
        orc.run(F, p, k, n, nproma, klev, scalars["PTSPHY"], 1)
        # every rank validates its shard against the replicated reference, phase-shifted like its inputs
        ref = state.load_reference()
        R = {name: state.expand(a, n, nproma, first_column=first) for name, a in ref.items()}
        local = validate.validate(F, R, n)
        glob = shard.reduce_field_errors(local, world * n, dist)
        t = shard.reduce_max(10.0 + rank, dist)          # "step time" of this rank
        total = shard.reduce_sum(float(n), dist)
        checksum = {name: float(np.abs(state.unblock(F[name], n)).sum()) for name in ("TENDENCY_LOC_T", "PFPLSN")}
        q.put((rank, t, total, [tuple(e) for e in glob], checksum, first))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_shard_reduce_and_validate(built):
    import torch.multiprocessing as mp
    from dwarf_p_cloudsc_b200 import params, state, validate
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    world, ngptot, nproma = 2, 150, 32          # 150 is not a multiple of 100 nor of 32: phases differ per rank
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, ngptot, nproma, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (r0, t0, tot0, g0, c0, f0), (r1, t1, tot1, g1, c1, f1) = res
    assert (f0, f1) == (0, 150)
    assert t0 == t1 == 11.0                      # MAX over ranks
    assert tot0 == tot1 == 300.0
    assert g0 == g1                              # every rank holds the same global table
    errs = [validate.FieldError(*e) for e in g0]
    assert validate.all_pass(errs) and abs(errs[0].avgpgp - errs[0].errsum / 300.0) < 1e-30

    # the two shards together are exactly one 300-column run
    cols, scalars, _, klev = state.load_input()
    p, k = params.params_from_scalars(scalars)
    F = state.blocked_fields(cols, 300, 150, klev)      # NPROMA 150: block b == rank b's column range
    pyoracle.best_oracle().run(F, p, k, 300, 150, klev, scalars["PTSPHY"], 1)
    for name in ("TENDENCY_LOC_T", "PFPLSN"):
        u = np.abs(state.unblock(F[name], 300))
        assert np.isclose(u[..., :150].sum(), c0[name], rtol=1e-13) and np.isclose(u[..., 150:].sum(), c1[name], rtol=1e-13)


def test_strong_shard_covers_all_blocks_once():
    from dwarf_p_cloudsc_b200 import shard
    for ngptot, nproma, world in [(163840, 128, 8), (1000, 128, 3), (100, 16, 8), (257, 32, 4), (5, 8, 4)]:
        nblocks = ngptot // nproma + min(ngptot % nproma, 1)
        seen, cols = [], 0
        for r in range(world):
            b0, nb, nc = shard.strong_shard(r, world, ngptot, nproma)
            seen += list(range(b0, b0 + nb))
            cols += nc
        assert seen == list(range(nblocks)) and cols == ngptot, (ngptot, nproma, world)
    assert shard.weak_shard(3, 8, 163840) == (3 * 163840, 163840)
    with pytest.raises(ValueError):
        shard.weak_shard(8, 8, 10)
