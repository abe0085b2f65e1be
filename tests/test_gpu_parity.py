"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar: every integer tally bit-exact, PI / Fisher p within 1e-9 relative, rows byte-identical."""
import pytest

from smcounter_b200.caller import VcParams
from smcounter_b200.synth import SynthSpec

pytestmark = pytest.mark.gpu

CASES = {
    "snv_basic": (dict(umis_per_locus=60, rpb=3.0, snv_every=50, snv_vaf=0.1), dict(mtDepth=60, rpb=3.0), [("chr1", 1000, 1200), ("chr2", 500, 560)], 7),
    "indel_heavy": (dict(umis_per_locus=50, rpb=3.0, snv_every=40, snv_vaf=0.08, indel_every=30, indel_vaf=0.08), dict(mtDepth=50, rpb=3.0), [("chr1", 2000, 2180)], 11),
    "mtdrop1_rpb8": (dict(umis_per_locus=40, rpb=8.6, snv_every=60, snv_vaf=0.05, indel_every=70, indel_vaf=0.03), dict(mtDepth=40, rpb=8.6, mtDrop=1, hpLen=8), [("chr17", 41243700, 41243860)], 20170410),
    "low_rpb": (dict(umis_per_locus=120, rpb=1.2, snv_every=30, snv_vaf=0.2, n_frac=0.01), dict(mtDepth=120, rpb=1.2), [("chrX", 100, 260)], 3),
    "strict_bq_mq": (dict(umis_per_locus=60, rpb=4.0, snv_every=45, snv_vaf=0.5, lowmapq_frac=0.2, softclip_frac=0.3), dict(mtDepth=60, rpb=2.0, minBQ=31, minMQ=50, mismatchThr=3.0, primerDist=5), [("chr3", 700, 900)], 5),
    "ragged_intervals": (dict(umis_per_locus=30, rpb=3.0, snv_every=25, snv_vaf=0.9, indel_every=45, indel_vaf=0.4), dict(mtDepth=30, rpb=3.0), [("chr1", 100, 101), ("chr1", 140, 173), ("chr1", 173, 175), ("chr2", 5, 70), ("chr1", 150, 160)], 13),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_case(name):
    from helpers import run_case
    spec_kw, prm_kw, intervals, seed = CASES[name]
    problems, stats, _ = run_case(intervals, SynthSpec(**spec_kw), VcParams(**prm_kw), seed)
    print(name, stats)
    assert not problems, "\n".join(problems)
    assert stats["events"] > 0
