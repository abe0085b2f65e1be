"""Generates tests/golden/ref_rows.json: the 45-column rows printed by the REFERENCE'S OWN vc() (/root/reference/smCounter.py,
run through oracle/ref_build.py + oracle/ref_shims.py with Python-2 container order) on the hand-written parity cases, a few
loci at the depths of the BASELINE.json configurations and a few cases of the fuzz corpus.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_ref_rows.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from fuzz import CASES, DEEP_CASES, case_inputs, fuzz_case     # noqa: E402
from oracle import ref_build, ref_shims                         # noqa: E402
from oracle import smcounter_oracle as orc                      # noqa: E402
from smcounter_b200.soa import soa_to_records                   # noqa: E402
from smcounter_b200.synth import make_panel                     # noqa: E402
from smcounter_b200.targets import loc_list                     # noqa: E402

FUZZ = (101, 104, 107, 110)


def main():
    ref = ref_build.load("py2")
    cases = {}
    todo = [(n,) + case_inputs(n) for n in sorted(CASES) + sorted(DEEP_CASES)]
    for seed in FUZZ:
        ivs, spec, prm = fuzz_case(seed)
        todo.append(("fuzz%d" % seed, ivs, spec, prm, seed))
    for (name, ivs, spec, prm, seed) in todo:
        soa, refs, _ = make_panel(ivs, spec, seed=seed)
        bam = ref_shims.register_bam("mem.bam", soa_to_records(soa, orc.Read))
        fa = ref_shims.register_fasta("mem.fa", refs)
        cases[name] = [ref.vc(bam, chrom, pos, prm.minBQ, prm.minMQ, prm.mtDepth, prm.rpb, prm.hpLen, prm.mismatchThr, prm.mtDrop,
                              prm.maxMT, prm.primerDist, fa) for (chrom, pos) in loc_list(ivs)]
        print(name, len(cases[name]), "rows")
    with open(ref_build.REF_SRC, "rb") as fh:
        sha = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(HERE, "ref_rows.json"), "w") as fh:
        json.dump({"made_by": "tests/golden/make_ref_rows.py", "source": ref_build.REF_SRC, "source_sha256": sha, "cases": cases}, fh, indent=0)


if __name__ == "__main__":
    main()
