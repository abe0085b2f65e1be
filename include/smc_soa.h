/*
 * smc_soa.h -- C ABI of the host-side batch packer (compiled into libsmc_bamio.so next to the BAM decoder).
 *
 * Between the decoded BAM (smc_bam_reads, include/smc_bamio.h: every read of the file, 32-bit scalars, one byte per
 * quality, BAM's 4-bit bases) and smc_call_batch (include/smc_b200.h) sits what the reference does implicitly when
 * pysam.pileup() hands vc() the reads of ONE locus (smCounter.py:275,316-349): pick the reads of a batch of target
 * intervals.  Here that is one threaded pass that gathers the selected reads AND writes them in the compact wire
 * encodings of smc_reads_soa (ABI v3: 16-bit nm / l_seq / store_lo / store_len, 2- / 4-bit quality codes behind a
 * codebook, 2-bit bases with a side list of the non-ACGT ones), straight into buffers the caller owns -- pinned host
 * memory, so the upload is one DMA per array.  Fragment ids are renumbered densely in their old relative order
 * (the contract of smc_reads_soa.frag_id).
 *
 *   smc_soa_qual_hist   once per decoded BAM: which quality values occur (the codebook is valid for every batch of it)
 *   smc_soa_ref_end     once per decoded BAM: reference end of every read (interval look-ups)
 *   smc_soa_pack_begin  sizes of the batch (prefix sums over the selected reads)
 *   smc_soa_pack_fill   the pass itself
 *   smc_soa_pack_end    releases the handle (and the exception list it owns)
 */
#ifndef SMC_SOA_H
#define SMC_SOA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Plain reads: the arrays of smc_bam_reads / smc_reads_soa with scalar_bits 32, qual_bits 8, seq_bits 4, offsets given. */
typedef struct smc_soa_view {
    int64_t         n_reads;
    const int32_t  *ref_id;
    const int32_t  *pos;
    const uint16_t *flag;
    const uint8_t  *mapq;
    const int32_t  *nm;
    const int32_t  *l_seq;
    const int64_t  *seq_off;
    const int64_t  *qual_off;
    const int64_t  *cigar_off;
    const uint16_t *n_cigar;
    const uint64_t *umi;
    const uint32_t *frag_id;
    const uint8_t  *seq;
    const uint8_t  *qual;
    const uint32_t *cigar;
    const int32_t  *store_lo;       /* both NULL: reads stored whole */
    const int32_t  *store_len;
} smc_soa_view;

typedef struct smc_soa_pack_opts {
    int32_t  scalar_bits;           /* 8, 16 or 32 (8 / 16 need every nm / l_seq / store_lo / store_len < 256 / < 65536: checked) */
    int32_t  qual_bits;             /* 2, 4 or 8 */
    int32_t  seq_bits;              /* 2 or 4 */
    int32_t  threads;               /* <= 0: all */
    int32_t  ref_id_bits;           /* 8 (every ref_id < 256: checked) or 32 */
    int32_t  umi_bits;              /* 32 (every barcode code < 2^32: checked) or 64 */
    uint8_t  code_of[256];          /* qual_bits != 8: code of every phred value that occurs (others must not occur: checked) */
} smc_soa_pack_opts;

typedef struct smc_soa_pack_sizes {
    int64_t n_reads;
    int64_t seq_bytes;
    int64_t qual_bytes;
    int64_t n_cigar_words;
} smc_soa_pack_sizes;

/* Output buffers, sized from smc_soa_pack_sizes: per-read arrays hold n_reads entries (nm .. store_len: uint8, uint16 or int32
 * by scalar_bits; store_lo / store_len may be NULL when the view has none), seq / qual / cigar the byte / word counts.
 * seq_poff (optional, n_reads + 1 entries) receives the byte offset of every read inside seq (host-side look-ups only). */
typedef struct smc_soa_pack_bufs {
    void     *ref_id;               /* uint8 or int32 by ref_id_bits */
    int32_t  *pos;
    uint16_t *flag;
    uint8_t  *mapq;
    void     *nm;
    void     *l_seq;
    void     *store_lo;
    void     *store_len;
    uint16_t *n_cigar;
    void     *umi;                  /* uint32 or uint64 by umi_bits */
    uint32_t *frag_id;
    uint8_t  *seq;
    uint8_t  *qual;
    uint32_t *cigar;
    int64_t  *seq_poff;
} smc_soa_pack_bufs;

/* The non-ACGT bases of a seq_bits == 2 batch, sorted by (read, base index); owned by the handle. */
typedef struct smc_soa_pack_exc {
    int64_t         n;
    const uint32_t *read;
    const uint32_t *pos;
    const uint8_t  *nib;
} smc_soa_pack_exc;

typedef struct smc_soa_pack smc_soa_pack;

#define SMC_SOA_OK       0
#define SMC_SOA_E_ARG   -1
#define SMC_SOA_E_RANGE -2      /* a scalar does not fit 16 bits / a quality has no code / idx not ascending or out of range */
#define SMC_SOA_E_MEM   -3

int  smc_soa_qual_hist(const smc_soa_view *v, int threads, uint64_t hist[256]);
/* 0-based exclusive reference end of every read (pos + the reference bases its CIGAR consumes: M D N = X); what the host
 * needs to find the reads of an interval. */
int  smc_soa_ref_end(const smc_soa_view *v, int threads, int64_t *ref_end);
/* The same pass, plus what an interval look-up over the reads needs to know: stats[0] = 1 when the reads are in BAM coordinate
 * order ((ref_id, pos) never decreases), stats[1] = the longest reference span of a read. */
int  smc_soa_order_stats(const smc_soa_view *v, int threads, int64_t *ref_end, int64_t stats[2]);
/* idx: ascending read indices of the batch, or NULL = all reads of the view. */
int  smc_soa_pack_begin(const smc_soa_view *v, const int64_t *idx, int64_t n_idx, const smc_soa_pack_opts *opts,
                        smc_soa_pack **out, smc_soa_pack_sizes *sizes);
int  smc_soa_pack_fill(smc_soa_pack *h, const smc_soa_pack_bufs *bufs, smc_soa_pack_exc *exc);
void smc_soa_pack_end(smc_soa_pack *h);

#ifdef __cplusplus
}
#endif
#endif
