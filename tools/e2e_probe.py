"""Where one end-to-end batch spends its host time: cProfile over GpuCaller.call() + device_hp_flags() on one bench batch,
for the 4-bit and the 2-bit base encodings.  Usage (GPU box): python tools/e2e_probe.py [intervals]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from smcounter_b200.caller import GpuCaller, LocusResults, VcParams
from smcounter_b200.rows import device_hp_flags


def main():
    n_iv = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    for bits in ("4", "2"):
        os.environ["SMC_BENCH_SEQ_BITS"] = bits
        args = bench.parse_args(["--batches", "1", "--intervals", str(n_iv)])
        (ivs, soa, refs, loci, bed_order, _), = bench.make_batches(args, 0, 1)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        for f in ("ref_id", "pos", "flag", "mapq", "nm", "l_seq", "n_cigar", "umi", "frag_id", "seq", "qual", "cigar", "store_lo", "store_len", "qual_lut"):
            setattr(soa, f, pin(getattr(soa, f)))
        if soa.seq_exc is not None:
            soa.seq_exc = tuple(pin(a) for a in soa.seq_exc)
        prm = VcParams(**bench.WORKLOADS[args.workload][2])
        c = GpuCaller(prm, 0)
        res = c.call(soa, loci)
        out = c.download(LocusResults(loci.n, max(int(c.timings()["n_dyn"]), 16), pinned=True))

        def one():
            r = c.call(soa, loci, out=out)
            device_hp_flags(c, r, soa, loci, soa.chroms, refs, prm.hpLen)
        for _ in range(3):
            one()
        t0 = time.perf_counter()
        for _ in range(10):
            one()
        dt = (time.perf_counter() - t0) / 10
        tm = c.timings()
        print("seq_bits=%s reads=%d exc=%d: %.2f ms per batch; h2d %.2f device %.2f d2h %.2f launches %s" % (
            bits, soa.n, 0 if soa.seq_exc is None else len(soa.seq_exc[0]), dt * 1e3, tm["ms_h2d"], tm["ms_total_device"], tm["ms_d2h"], tm.get("pipe_launches")))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(10):
            one()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
        c.close()


if __name__ == "__main__":
    main()
