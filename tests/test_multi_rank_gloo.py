"""N > 1 host path on CPU: two ranks (gloo) each take their depth-balanced BED-interval shard and only the reads that
overlap it, call their loci independently (the oracle stands in for the GPU here -- this is a test), and rank 0
concatenates the rows in BED order.  No data-path collective: the only communication is the final gather, as in the
reference's ``[p.get() for p in results]`` (smCounter.py:685)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

IVS = [("chr1", 500, 530), ("chr1", 700, 745), ("chr2", 100, 120), ("chr1", 900, 910), ("chr2", 300, 360), ("chr1", 720, 735)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_for(soa, ivs, refs, prm):
    from oracle import smcounter_oracle as orc
    from smcounter_b200.soa import soa_to_records
    from smcounter_b200.targets import loc_list
    idx = orc.ReadIndex(soa_to_records(soa, orc.Read))
    return [orc.vc(idx, c, p, prm.minBQ, prm.minMQ, prm.mtDepth, prm.rpb, prm.hpLen, prm.mismatchThr, prm.mtDrop, prm.maxMT,
                   prm.primerDist, refs) for (c, p) in loc_list(ivs)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from smcounter_b200.caller import VcParams
    from smcounter_b200.shard import interleave_rows, plan_shards, reads_for_intervals
    from smcounter_b200.synth import SynthSpec, make_panel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    spec = SynthSpec(umis_per_locus=14, rpb=2.5, snv_every=25, snv_vaf=0.2, indel_every=35, indel_vaf=0.2, depth_sigma=0.6)
    soa, refs, _ = make_panel(IVS, spec, seed=21)
    prm = VcParams(mtDepth=14, rpb=2.5)
    plan = plan_shards(soa, IVS, soa.chroms, world)
    mine = [IVS[k] for k in plan[rank][0]]
    sub = soa.select(reads_for_intervals(soa, mine, soa.chroms))
    rows = _rows_for(sub, mine, refs, prm)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        merged = interleave_rows(plan, IVS, gathered)
        whole = _rows_for(soa, IVS, refs, prm)
        q.put((merged == whole, len(whole), [len(g) for g in gathered], sub.n, soa.n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process_rows():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n, per_rank, n_sub, n_all = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n == sum(e - s for _, s, e in IVS) == sum(per_rank)
    assert min(per_rank) > 0 and n_sub < n_all          # both ranks worked, each on a strict subset of the reads
