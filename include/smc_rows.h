/*
 * smc_rows.h -- C ABI of the host-side output stage (compiled into libsmc_bamio.so next to the BAM decoder).
 *
 * Turns the per-locus results of smc_call_batch (smc_out, include/smc_b200.h) into the text the reference writes:
 *   * the 45-column row of vc()                      smCounter.py:575-600 (Python-2 round() / str() semantics, :552-573
 *                                                    bi-allelic resolution, FILTER assembly of filterVariants() :184-269);
 *   * the repeat filters of main()                   smCounter.py:751-785 (first TRF / RepeatMasker region with
 *                                                    locL < pos <= locR, PASS / strip(';'));
 *   * the called-variant lines of the three writers  smCounter.py:832-891 (cut.txt row, VCF row with the GT / AD hack).
 * One pass per row, rows spread over host threads; nothing is re-parsed from strings.  Everything that needs strings the
 * device does not have (names of the non-ACGT alleles, contig names) comes in from the caller.
 */
#ifndef SMC_ROWS_H
#define SMC_ROWS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smc_rows_in {
    int64_t         n_loci;         /* loci of the batch: every per-locus array below has this many entries (stride of the 2-D ones) */
    int64_t         n_rows;         /* rows to emit */
    const int64_t  *order;          /* locus index of row k (BED order, duplicates allowed); NULL = 0 .. n_loci-1 */
    const int32_t  *ref_id;         /* smc_loci */
    const int32_t  *pos0;
    const uint8_t  *ref_base;
    int32_t         n_chroms;
    const char *const *chroms;      /* contig names by ref_id */
    /* smc_out arrays of the batch */
    const int32_t  *loc;            /* [SMC_NLOC][n_loci] */
    const int32_t  *cnt;            /* [SMC_NFIXED][SMC_NCNT][n_loci] */
    const double   *pi;             /* [SMC_NFIXED][n_loci] */
    const int32_t  *alt_allele;
    const int32_t  *second_allele;
    const uint32_t *fl1;
    const uint32_t *fl2;
    const uint8_t  *biallelic;
    int64_t         n_dyn;
    const int32_t  *dyn_cnt;        /* [n_dyn][SMC_NCNT] */
    const double   *dyn_pi;
    const char     *dyn_names;      /* the reference's allele string ('N', 'INS|A|AT', 'DEL|AC|A') of dynamic row j is           */
    const int64_t  *dyn_name_off;   /* dyn_names[dyn_name_off[j] .. dyn_name_off[j+1]); needed for rows that are an ALT candidate */
    const uint8_t  *hp1;            /* isHPorLowComp() of candidate 0 / 1 per locus (smc_hp_lowcomp): bit 0 homopolymer, bit 1 low */
    const uint8_t  *hp2;            /* complexity, bit 7 "computed"; may be NULL when no candidate reaches that test               */
    int32_t         finalize;       /* 0: rows as vc() returns them (FILTER still ';TAG;TAG;'); 1: main()'s post-processing too     */
    int32_t         threshold;      /* finalize: PI cut-off of the cut.txt / cut.vcf rows (smCounter.py:820, :850) */
    /* finalize: repeat regions in the order the reference scans them (per contig, sorted by start) */
    int64_t         n_trf;
    const int32_t  *trf_chrom;      /* index into chroms */
    const int64_t  *trf_lo;
    const int64_t  *trf_hi;
    int64_t         n_rm;
    const int32_t  *rm_chrom;
    const int64_t  *rm_lo;
    const int64_t  *rm_hi;
    const char     *rm_tags;        /* tag string of region r ('RepS;LowC;'): rm_tags[rm_tag_off[r] .. rm_tag_off[r+1]) */
    const int64_t  *rm_tag_off;
    int32_t         threads;        /* host threads (<= 0: all) */
    int32_t         reserved0;
} smc_rows_in;

typedef struct smc_rows_out {
    char    *all;        /* the rows, each terminated by '\n' */
    int64_t *all_off;    /* n_rows + 1 offsets into all */
    char    *cut;        /* finalize: 14-column rows of the called variants ('' for the others) */
    int64_t *cut_off;
    char    *vcf;        /* finalize: VCF rows of the called variants */
    int64_t *vcf_off;
    int64_t  bad_row;    /* on SMC_ROWS_E_STATUS / SMC_ROWS_E_HP / SMC_ROWS_E_NAME: the row that failed */
    uint32_t bad_status; /* its device status bits */
} smc_rows_out;

#define SMC_ROWS_OK        0
#define SMC_ROWS_E_ARG    -1
#define SMC_ROWS_E_STATUS -2    /* a locus carries a device status that has no row (needs a down-sampling mask, overflow) */
#define SMC_ROWS_E_HP     -3    /* a candidate reaches the HP / LowC test but no flags were supplied for it */
#define SMC_ROWS_E_NAME   -4    /* a dynamic allele is reported but has no name */
#define SMC_ROWS_E_MEM    -5

/* Buffers of *out are allocated by the library; release them with smc_rows_free.  Thread safe (no global state). */
int  smc_rows_emit(const smc_rows_in *in, smc_rows_out *out);
void smc_rows_free(smc_rows_out *out);

#ifdef __cplusplus
}
#endif
#endif
