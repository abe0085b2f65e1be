"""Output writers of main() (reference smCounter.py:742-749, 787-901): ``<outPrefix>.smCounter.all.txt`` (45 columns),
``.smCounter.cut.txt`` (14 columns) and ``.smCounter.cut.vcf``, byte-compatible with the reference.

Input rows are the 45-field tab-joined strings of vc() after the repeat filters (FILTER already 'PASS' or tags).
"""
from __future__ import annotations

import math

from .rows import headerAll, headerVariants

_IDX = {h: i for i, h in enumerate(headerAll)}

_VCF_META = (
    '##fileformat=VCFv4.2',
    '##reference=GRCh37',
    '##INFO=<ID=TYPE,Number=1,Type=String,Description="Variant type: SNP or INDEL">',
    '##INFO=<ID=DP,Number=1,Type=Integer,Description="Total read depth">',
    '##INFO=<ID=MT,Number=1,Type=Integer,Description="Total MT depth">',
    '##INFO=<ID=UMT,Number=1,Type=Integer,Description="Filtered MT depth">',
    '##INFO=<ID=PI,Number=1,Type=Float,Description="Variant prediction index">',
    '##INFO=<ID=THR,Number=1,Type=Integer,Description="Variant prediction index minimum threshold">',
    '##INFO=<ID=VMT,Number=1,Type=Integer,Description="Variant MT depth">',
    '##INFO=<ID=VMF,Number=1,Type=Float,Description="Variant MT fraction">',
    '##INFO=<ID=VSM,Number=1,Type=Integer,Description="Variant strong MT depth">',
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
    '##FORMAT=<ID=AD,Number=.,Type=Integer,Description="Filtered allelic MT depths for the ref and alt alleles">',
    '##FORMAT=<ID=VF,Number=1,Type=Float,Description="Variant MT fraction, same as VMF">',
    '##FILTER=<ID=RepT,Description="Variant in simple tandem repeat region, as defined by Tandem Repeats Finder">',
    '##FILTER=<ID=RepS,Description="Variant in simple repeat region, as defined by RepeatMasker">',
    '##FILTER=<ID=LowC,Description="Variant in low complexity region, as defined by RepeatMasker">',
    '##FILTER=<ID=SL,Description="Variant in micro-satelite region, as defined by RepeatMasker">',
    '##FILTER=<ID=HP,Description="Inside or flanked by homopolymer region">',
    '##FILTER=<ID=LM,Description="Low coverage (fewer than 5 MTs)">',
    '##FILTER=<ID=LSM,Description="Fewer than 2 strong MTs">',
    '##FILTER=<ID=SB,Description="Strand bias">',
    '##FILTER=<ID=LowQ,Description="Low base quality (mean < 22)">',
    '##FILTER=<ID=MM,Description="Too many genome reference mismatches in reads (default threshold is 6.5 per 100 bases)">',
    '##FILTER=<ID=DP,Description="Too many discordant read pairs">',
    '##FILTER=<ID=R1CP,Description="Variants are clustered at the end of R1 reads">',
    '##FILTER=<ID=R2CP,Description="Variants are clustered at the end of R2 reads">',
    '##FILTER=<ID=PrimerCP,Description="Variants are clustered immediately after the primer, possible enzyme initiation error">',
)


def vcf_header(outPrefix: str) -> str:
    """smCounter.py:788-817; the sample column is named after outPrefix."""
    cols = ('#CHROM', 'POS', 'ID', 'REF', 'ALT', 'QUAL', 'FILTER', 'INFO', 'FORMAT', outPrefix)
    return "\n".join(_VCF_META) + "\n" + "\t".join(cols) + "\n"


def pi_threshold(mtDepth: int, threshold_arg: int = 0) -> int:
    """smCounter.py:820: cutoff for about 20 FP/Mb unless given."""
    return int(math.ceil(14.0 + 0.012 * mtDepth)) if threshold_arg == 0 else threshold_arg


def called_lines(fields, threshold):
    """(vcfLine, shortLine) for one row, or None when it is not called (smCounter.py:840-891)."""
    PI = fields[_IDX['PI']]
    if len(PI) == 0:
        return None
    ALT = fields[_IDX['ALT']]
    QUAL = str(int(float(PI)))                      # truncated PI, VCF phred-like tradition (:847)
    if not (int(QUAL) >= threshold and ALT != 'DEL'):
        return None
    g = lambda k: fields[_IDX[k]]
    CHROM, POS, REF, TYPE, DP, MT, UMT = g('CHROM'), g('POS'), g('REF'), g('TYPE'), g('DP'), g('MT'), g('UMT')
    VMT, VMF, VSM, FILTER = g('VMT'), g('VMF'), g('VSM'), g('FILTER')
    THR = str(threshold)
    INFO = ';'.join(('TYPE=' + TYPE, 'DP=' + DP, 'MT=' + MT, 'UMT=' + UMT, 'PI=' + PI, 'THR=' + THR, 'VMT=' + VMT,
                     'VMF=' + VMF, 'VSM=' + VSM))
    alts = ALT.split(",")
    if len(alts) == 2:                              # :868-878 genotype hack
        genotype = '1/2'
    elif len(alts) != 1:
        raise Exception("error hacking genotype field for " + str(alts))
    elif CHROM == "chrY" or CHROM == "chrM":
        genotype = '1'
    elif float(VMF) > 0.95:
        genotype = '1/1'
    else:
        genotype = '0/1'
    AD = str(int(UMT) - int(VMT)) + "," + VMT
    if len(alts) == 2:
        AD += ",1"
    SAMPLE = ":".join((genotype, AD, VMF))
    vcfLine = '\t'.join((CHROM, POS, '.', REF, ALT, QUAL, FILTER, INFO, 'GT:AD:VF', SAMPLE)) + '\n'
    shortLine = '\t'.join((CHROM, POS, REF, ALT, TYPE, DP, MT, UMT, PI, THR, VMT, VMF, VSM, FILTER)) + '\n'
    return vcfLine, shortLine


def render_outputs(rows, outPrefix, mtDepth, threshold_arg=0):
    """(threshold, all_txt, cut_txt, cut_vcf) as strings."""
    threshold = pi_threshold(mtDepth, threshold_arg)
    outAll = ['\t'.join(headerAll) + '\n']
    outVariants = ['\t'.join(headerVariants) + '\n']
    outVcf = [vcf_header(outPrefix)]
    for line in rows:
        outAll.append(line + "\n")
        r = called_lines(line.split('\t'), threshold)
        if r is not None:
            outVcf.append(r[0])
            outVariants.append(r[1])
    return threshold, "".join(outAll), "".join(outVariants), "".join(outVcf)


def write_outputs(rows, outPrefix, mtDepth, threshold_arg=0):
    """Writes the three files next to ``outPrefix`` (smCounter.py:823-901) and returns the threshold (:909)."""
    threshold, a, c, v = render_outputs(rows, outPrefix, mtDepth, threshold_arg)
    with open(outPrefix + '.smCounter.all.txt', 'w') as fh:
        fh.write(a)
    with open(outPrefix + '.smCounter.cut.txt', 'w') as fh:
        fh.write(c)
    with open(outPrefix + '.smCounter.cut.vcf', 'w') as fh:
        fh.write(v)
    return threshold
