// K1 (read prep / CIGAR walk) and K3 (tile pileup + fragment merge + per-barcode posterior) kernels.
//
// Thread mapping of the pileup kernel ("the transpose"): one warp owns a tile of 32 consecutive target loci,
// LANE = LOCUS.  The warp streams the tile's reads in (barcode, fragment, BAM index) order; every lane applies
// the read to its own locus.  Barcode and fragment boundaries are therefore warp-uniform, every counter is
// lane-private (no atomics in the loop), and the reference's order-dependent semantics (first read of a
// fragment defines its base, discordant mates delete the fragment, a third read may recreate it --
// smCounter.py:467-479) are reproduced by a plain per-lane state machine.
#pragma once
#include "smc_common.cuh"

// ------------------------------------------------------------------------------------------------------------
// K1: per-read preparation, in srank order (thread s handles read perm[s]).
// Restates smCounter.py:327-356 (mapq, NM, nIndel, leftSP, mismatchPer100b) and the htslib column membership
// pos <= p < reference_end, once per read instead of once per pileup event.
// ------------------------------------------------------------------------------------------------------------
struct PrepArgs {
    int64_t n_reads;
    const uint32_t* perm;         // srank -> read index
    const uint32_t* urank;        // per srank
    const uint32_t* frank;
    const int32_t* ref_id; const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const int32_t* nm;
    const int32_t* l_seq; const int64_t* seq_off; const int64_t* qual_off; const int64_t* cigar_off;
    const uint16_t* n_cigar; const uint32_t* cigar;
    const uint64_t* loci_key; int64_t n_loci;
    int minMQ; int primerDist; double mismatchThr;
    ReadRec* recs; uint32_t* ntiles; uint32_t* gflags;
};

#define GF_DYN_FULL   1u
#define GF_BAD_READ   2u     // l_seq / clip length beyond the 16-bit record fields

__global__ void __launch_bounds__(256) k_read_prep(PrepArgs A) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n_reads) return;
    uint32_t r = A.perm[s];
    uint32_t ncig = A.n_cigar[r];
    int64_t co = A.cigar_off[r];
    int32_t lseq = A.l_seq[r];
    int reflen = 0, nindel = 0, leftSP = 0, lead = 0, trail = 0, n_refops = 0;
    bool simple = true, in_lead = true;
    uint32_t c4[4] = {0, 0, 0, 0};
    for (uint32_t k = 0; k < ncig; ++k) {
        uint32_t cw = A.cigar[co + k];
        if (k < 4) c4[k] = cw;
        uint32_t op = cw & 15u; int len = (int)(cw >> 4);
        if (op == 1 || op == 2) nindel += len;                 // smCounter.py:343-344
        if (k == 0 && op == 4) leftSP = len;                   // :345-346
        if (op == 0 || op == 7 || op == 8) { reflen += len; ++n_refops; in_lead = false; trail = 0; }
        else if (op == 2 || op == 3) { reflen += len; simple = false; in_lead = false; trail = 0; }
        else if (op == 4) { if (in_lead) lead += len; else trail += len; }
        else if (op == 5) { simple = false; }
        else { simple = false; in_lead = false; trail = 0; }   // I, P
    }
    if (n_refops != 1) simple = false;
    int alnlen = lseq - lead - trail;                          // query_alignment_length
    int nmv = A.nm[r];
    int mismatch = nmv - nindel; if (mismatch < 0) mismatch = 0;                       // :352
    double mm100 = lseq > 0 ? (100.0 * (double)mismatch) / (double)lseq : 0.0;          // :356
    uint32_t fl = A.flag[r];
    bool ok = ((int)A.mapq[r] >= A.minMQ) && (mm100 <= A.mismatchThr);
    int32_t start = A.pos[r];
    int64_t lo = 0, hi = 0;
    if (!(fl & 0x4u) && reflen > 0) {
        uint64_t k0 = ((uint64_t)(uint32_t)A.ref_id[r] << 32) | (uint32_t)start;
        uint64_t k1 = ((uint64_t)(uint32_t)A.ref_id[r] << 32) | (uint32_t)(start + reflen);
        lo = lower_bound_u64(A.loci_key, A.n_loci, k0);
        hi = lower_bound_u64(A.loci_key, A.n_loci, k1);
    }
    if (lseq > 65535 || leftSP > 65535 || alnlen < 0 || alnlen > 65535) { atomicOr(A.gflags, GF_BAD_READ); lo = hi = 0; }
    ReadRec rec;
    rec.start = start; rec.lo = (int32_t)lo; rec.hi = (int32_t)hi;
    const bool rev = fl & 0x10u, r2 = fl & 0x80u;
    rec.meta = (ok ? RM_OK : 0u) | (rev ? RM_REVERSE : 0u) | (r2 ? RM_READ2 : 0u) | (simple ? RM_SIMPLE : 0u) | (ncig << 8);
    rec.seq_off = (uint32_t)A.seq_off[r]; rec.qual_off = (uint32_t)A.qual_off[r]; rec.cigar_off = (uint32_t)co;
    rec.urank = A.urank[s]; rec.frank = A.frank[s];
    rec.read_idx = r; rec.gspan = simple ? (uint32_t)(hi - lo) : 0u;
    if (simple) {
        // One aligned run: d = p - start is the distance from the alignment start, alnlen - d from its end.
        //   R1: distToBcEnd = rev ? alnlen - d : d;   R2: distToBcEnd = rev ? d : alnlen - d, distToPrimerEnd = rev ? alnlen - d : d
        // "<= X from the start" is the window [start, start + X], "<= X from the end" is [start + alnlen - X, inf).
        rec.sp_aln = (uint32_t)(leftSP - start);
        const uint32_t EMPTY_LO = 0x7fffffffu;
        auto from_start = [&](int X, uint32_t& wlo, uint32_t& wspan) { if (X < 0) { wlo = EMPTY_LO; wspan = 0; } else { wlo = (uint32_t)start; wspan = (uint32_t)X; } };
        auto from_end = [&](int X, uint32_t& wlo, uint32_t& wspan) { wlo = (uint32_t)(start + alnlen - X); wspan = 0x7fffffffu; };
        const bool bc_at_start = (r2 == rev);                 // R1 fwd / R2 rev measure the barcode end from the start
        if (bc_at_start) from_start(20, rec.cig[0], rec.cig[1]); else from_end(20, rec.cig[0], rec.cig[1]);
        if (!r2) { rec.cig[2] = EMPTY_LO; rec.cig[3] = 0; }
        else if (rev) from_end(A.primerDist, rec.cig[2], rec.cig[3]);
        else from_start(A.primerDist, rec.cig[2], rec.cig[3]);
    } else {
        rec.sp_aln = (uint32_t)leftSP | ((uint32_t)alnlen << 16);
        rec.cig[0] = c4[0]; rec.cig[1] = c4[1]; rec.cig[2] = c4[2]; rec.cig[3] = c4[3];
    }
    const uint4* src = reinterpret_cast<const uint4*>(&rec);
    uint4* dst = reinterpret_cast<uint4*>(&A.recs[s]);
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
    A.ntiles[s] = hi > lo ? (uint32_t)(((hi - 1) >> 5) - (lo >> 5) + 1) : 0u;
}

// Expansion of reads into (tile, read) events -- the only "event" that is ever materialised: one 12-byte row per
// (read x 32-locus tile) instead of one per (read x locus).
__global__ void __launch_bounds__(256)
k_expand(const ReadRec* __restrict__ recs, const uint32_t* __restrict__ ev_off, int64_t n_reads,
         uint64_t* __restrict__ ev_key, uint32_t* __restrict__ ev_val) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_reads) return;
    int32_t lo = recs[s].lo, hi = recs[s].hi;
    if (hi <= lo) return;
    uint32_t t0 = (uint32_t)lo >> 5, t1 = (uint32_t)(hi - 1) >> 5;
    uint32_t o = ev_off[s];
    for (uint32_t t = t0; t <= t1; ++t, ++o) { ev_key[o] = t; ev_val[o] = (uint32_t)s; }
}

// ------------------------------------------------------------------------------------------------------------
// K3: tile pileup (v6)
//
// A warp owns one unit = a run of <= `chunk` tile events of one 32-locus tile, cut at barcode boundaries; LANE = LOCUS.
// It works in batches of 32 reads:
//   stage   : 32 ReadRec (64 B each, gathered through ev_read[]) -> shared memory, 4 x 128-bit loads per lane; barcode /
//             fragment boundaries and "simple read" flags become warp-uniform bit masks (ballots);
//   pass A  : "gather + tally" -- for each staged read every lane computes the query position of ITS locus, loads the
//             base nibble and the quality, and adds the event to the order-independent tallies (cvg, alleleCnt,
//             forward, lowQ, R1/R2 end-distance counts) held in registers as 4 x 8-bit fields (A, C, T, G) per word.
//             The reads of a batch are independent, so the loads are issued K3_GATHER reads at a time: this is where
//             all the DRAM/L2 latency of the kernel is, and it is hidden by ILP plus the other warps' pass B.
//             What the ordered pass still needs is packed into a 16-bit event code in shared memory;
//   pass B  : "merge" -- the order-dependent state machine (fragment merge :467-479, per-barcode posterior, consensus)
//             walks the codes.  Register counters are spilled to the 16-bit shared-memory counters every 224 events;
//             those go to the global 32-bit accumulators every 49 152 events and at the end of the unit.
// Everything rare is out of line (__noinline__) so that the hot loop stays small in the instruction cache: reads with
// indels / hard clips / several aligned runs (per-event CIGAR walk), non-ACGT bases, barcodes showing several alleles.
// ------------------------------------------------------------------------------------------------------------
#ifndef K3_WARPS
#define K3_WARPS 4
#endif
#ifndef K3_MINBLOCKS
#define K3_MINBLOCKS 4
#endif
#ifndef K3_GATHER
#define K3_GATHER 4        // reads per gather group in pass A
#endif
#define NF SMC_NFIXED
#define NSLOT 7            // per-barcode allele slots: 5 fixed + 2 dynamic
// Per-lane shared-memory counters of the fixed alleles, two 16-bit counters per word:
enum { KW_ALLELE_FWD = 0,   // alleleCnt | forwardCnt << 16
       KW_R1,               // len(r1BcEndPos) | #<=20 << 16
       KW_R2,               // len(r2BcEndPos) | #<=20 << 16
       KW_LOWQ_R2P,         // lowQReads | #r2PrimerEndPos<=primerDist << 16
       KW_PAIR,             // concordPairCnt | discordPairCnt << 16
       KW_MT,               // MTCnt | strongMTCnt << 16
       K3_NW };
#define K3_FLUSH_EVERY 49152u      // tile events between shared -> global flushes (each event adds at most 1 to a field)
#define K3_REG_FLUSH   224u        // tile events between register -> shared flushes (8-bit fields)
#define K3_STAGE_WORDS 512                     // 32 ReadRec
#define K3_CODE_WORDS  512                     // 32 x 32 event codes, 16 bit
#define K3_FC_WORDS    (NF * K3_NW * 32)
#define K3_LIMB_WORDS  (NF * 4 * 32)           // 128-bit fixed-point PI accumulator per fixed allele and lane
#define K3_UCNT_WORDS  (NSLOT * 32)
#define K3_UPROD_WORDS (NSLOT * 64)
#define K3_WARP_WORDS  (K3_STAGE_WORDS + K3_CODE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS + K3_UCNT_WORDS + K3_UPROD_WORDS)
#define K3_TAB_BYTES   (2048 + 64)             // block-wide tables: 10^(-q/10) (256 doubles), nibble -> counter field increment
#define K3_SMEM_BYTES  (K3_TAB_BYTES + K3_WARPS * K3_WARP_WORDS * 4)

// event code: 16 bits from the gather pass (simple reads), a few more from the out-of-line CIGAR walk
#define EC_NIB_SH   8u            // bits 8-11: BAM nibble of the base (regular events)
#define EC_COVERED  (1u << 12)
#define EC_DYN      (1u << 13)    // regular base that is not A/C/G/T (N / IUPAC): dynamic allele row
#define EC_INC      (1u << 14)    // event passes incCond (:431): it enters bcDict
#define EC_REGULAR  (1u << 15)    // (slow path) a plain base, not an indel start / in-deletion event
#define EC_LE20     (1u << 16)    // (slow path) distance to the barcode end <= 20
#define EC_PLE      (1u << 17)    // (slow path) R2 and distance to the primer end <= primerDist

// Allele ids inside the merge state machine ("mid"): A/C/G/T = their BAM nibble (1, 2, 4, 8), in-deletion 'DEL' = 16,
// dynamic row e = 32 + e.  Slots / smc_out allele references (A0 C1 DEL2 T3 G4, 5 + row) are derived once per fragment.
#define MID_DEL 16u
#define MID_DYN 32u
__device__ __forceinline__ uint32_t mid_to_aid(uint32_t mid) {
    if (mid >= MID_DYN) return NF + (mid - MID_DYN);
    return mid == MID_DEL ? (uint32_t)SMC_A_DEL : ((0x3410u >> (4 * (__ffs(mid) - 1))) & 15u);   // nibble 1,2,4,8 -> slot 0,1,4,3
}

// per-lane flags
#define LF_FRAG_SEEN EC_COVERED   // a read of the open fragment covers the locus (same bit as EC_COVERED: one OR per event)
#define LF_UMI_SEEN  (1u << 0)
#define LF_UMI_BC    (1u << 1)    // the open barcode is in bcDict (a read passed incCond)
#define LF_F_EXISTS  (1u << 2)
#define LF_F_PAIRED  (1u << 3)

struct K3Args {
    const ReadRec* recs; const uint32_t* ev_read; const uint32_t* tile_off; const uint32_t* unit_off;
    uint32_t n_tiles; uint32_t chunk;
    const int32_t* loci_pos; int64_t n_loci;
    const uint8_t* seq; const uint8_t* qual; const uint32_t* cigar;
    const double* bqtab;            // [256]  10^(-bq/10), host glibc pow (smCounter.py:469)
    const double* pcrtab;           // [3][(nmax+1)(nmax+2)/2]  10^(-6 (cnt+.5)/(n+.5k)), k = 4,5,6 (smCounter.py:80-81)
    int pcr_nmax;
    int minBQ, mtDrop, primerDist; double smt;
    const int32_t* keep_idx; const int64_t* keep_off; const uint64_t* keep_umi; const uint64_t* umi_of_urank;
    int32_t* loc; int32_t* cnt; unsigned long long* limb;
    unsigned long long* dkey; uint32_t dmask; uint32_t* drep_read; int32_t* drep_qpos; int32_t* dlen;
    int32_t* dcnt; unsigned long long* dlimb; uint8_t* diskey; uint32_t* dcount; uint32_t* gflags;
    // optional: list the barcodes of bcDict for flagged loci (down-sampling support)
    const int32_t* list_idx; uint32_t* list_count; const int64_t* list_off; uint64_t* list_umi; uint32_t* list_first; int64_t list_cap;
};

// The dynamic-allele table, passed BY VALUE to the out-of-line helpers (taking the address of the kernel parameter block
// would copy all of it to local memory).
struct DynTab {
    unsigned long long* dkey; uint32_t dmask; uint32_t* drep_read; int32_t* drep_qpos; int32_t* dlen;
    int32_t* dcnt; unsigned long long* dlimb; uint8_t* diskey; uint32_t* dcount; uint32_t* gflags;
};
__device__ __forceinline__ DynTab dyn_tab(const K3Args& A) {
    DynTab T; T.dkey = A.dkey; T.dmask = A.dmask; T.drep_read = A.drep_read; T.drep_qpos = A.drep_qpos; T.dlen = A.dlen;
    T.dcnt = A.dcnt; T.dlimb = A.dlimb; T.diskey = A.diskey; T.dcount = A.dcount; T.gflags = A.gflags;
    return T;
}

// BAM nibble of A, C, G, T (1, 2, 4, 8) -> field A0 C1 T2 G3 of the packed register counters
__device__ __forceinline__ uint32_t nib_field(uint32_t nib) { return (0x20310u >> (2u * nib)) & 3u; }
__device__ __forceinline__ bool nib_is_acgt(uint32_t nib) { return (0x0116u >> nib) & 1u; }

// open-addressing table of the non-ACGT/DEL alleles
__device__ __noinline__ uint32_t dyn_lookup(DynTab T, unsigned long long key, uint32_t rep_read, int rep_qpos, int len) {
    uint32_t h = hash64to32(key) & T.dmask;
    for (uint32_t probe = 0; probe <= T.dmask; ++probe) {
        unsigned long long cur = __ldcg(&T.dkey[h]);
        if (cur == key) return h;
        if (cur == DYN_EMPTY) {
            unsigned long long prev = atomicCAS(&T.dkey[h], DYN_EMPTY, key);
            if (prev == DYN_EMPTY) {
                T.drep_read[h] = rep_read; T.drep_qpos[h] = rep_qpos; T.dlen[h] = len;
                uint32_t c = atomicAdd(T.dcount, 1u);
                if (2ull * (c + 1ull) > (unsigned long long)T.dmask + 1ull) atomicOr(T.gflags, GF_DYN_FULL);
                return h;
            }
            if (prev == key) return h;
        }
        h = (h + 1) & T.dmask;
    }
    atomicOr(T.gflags, GF_DYN_FULL);
    return 0;
}

// A non-negative double < 2^20 as a 128-bit fixed-point integer with LSB 2^-108 (exact: the terms are 0 or >= 2^-55).
__device__ __forceinline__ void pi_fixed128(double l, unsigned long long& lo, unsigned long long& hi) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(l);
    const int e = (int)((bits >> 52) & 0x7ffull);
    const unsigned long long m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    const int sh = e - 967;                             // value = m * 2^(e-1075) = (m << sh) * 2^-108
    if (e == 0 || sh <= -53) { lo = hi = 0; return; }
    if (sh <= 0) { lo = m >> (-sh); hi = 0; }
    else if (sh < 64) { lo = m << sh; hi = m >> (64 - sh); }
    else { lo = 0; hi = m << (sh - 64); }
}
__device__ __forceinline__ void add128(unsigned long long& lo, unsigned long long& hi, unsigned long long alo, unsigned long long ahi) {
    lo += alo; hi += ahi + (lo < alo ? 1ull : 0ull);
}
__device__ __forceinline__ void sub128(unsigned long long& lo, unsigned long long& hi, unsigned long long blo, unsigned long long bhi) {
    const unsigned long long borrow = lo < blo ? 1ull : 0ull;
    lo -= blo; hi -= bhi + borrow;
}
// 128-bit value -> three 44-bit carry-save limbs of the global accumulators (v = l0 + l1*2^44 + l2*2^88)
__device__ __forceinline__ void split_limbs(unsigned long long lo, unsigned long long hi, unsigned long long& a0, unsigned long long& a1,
                                            unsigned long long& a2) {
    const unsigned long long M44 = (1ull << 44) - 1ull;
    a0 = lo & M44;
    a1 = ((lo >> 44) | (hi << 20)) & M44;
    a2 = hi >> 24;
}

struct LaneState {
    // locus-level
    int cvg, allFrag, allMT, usedFrag, nBC, usedMT, mt3, mt5, mt7, mt10;
    uint32_t keymask, status;
    unsigned long long pad_lo, pad_hi;      // PI terms that go to all of A, C, G, T (single-allele barcodes, see umi_finalize)
    // hot counters, 4 x 8-bit fields (A, C, T, G)
    uint32_t r_allele, r_fwd, r_lowq, r_r1tot, r_r1le, r_r2tot, r_r2le, r_r2ple, r_conc;
    uint32_t flags;                         // LF_*
    // barcode-level
    int n; uint32_t exist; double Q, rightP; uint32_t last_aid; uint32_t udyn0, udyn1; int ndyn;
    uint32_t first_read;                    // BAM index of the barcode's first passing read at this locus (listing only)
    // fragment-level
    uint32_t f_mid; int f_bq;
};

#define FCW(w, a)   fc[((a) * K3_NW + (w)) * 32 + lane]
#define LIMB(a)     limb[(a) * 32 + lane]
#define UCNT(s)     ucnt[(s) * 32 + lane]
#define UPROD(s)    uprod[(s) * 32 + lane]

// counter update for an allele that may be dynamic: `word`/`add` address the packed shared-memory counter of a fixed
// allele, c_lo / c_hi are the smc_out counter indices the low / high half stand for.
__device__ __forceinline__ void bump(int32_t* dcnt, int* fc, int lane, uint32_t aid, int word, uint32_t add, int c_lo, int c_hi) {
    if (aid < NF) FCW(word, aid) += (int)add;
    else {
        int32_t* row = dcnt + (size_t)(aid - NF) * SMC_NCNT;
        if (add & 0xffffu) atomicAdd(&row[c_lo], 1);
        if (add >> 16) atomicAdd(&row[c_hi], 1);
    }
}

// registers -> shared-memory counters
__device__ __forceinline__ void flush_regs(int* fc, int lane, LaneState& S) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const int a = f + (f >> 1);                                     // A0 C1 T3 G4
        const uint32_t al = (S.r_allele >> (8 * f)) & 255u, fw = (S.r_fwd >> (8 * f)) & 255u;
        const uint32_t t1 = (S.r_r1tot >> (8 * f)) & 255u, l1 = (S.r_r1le >> (8 * f)) & 255u;
        const uint32_t t2 = (S.r_r2tot >> (8 * f)) & 255u, l2 = (S.r_r2le >> (8 * f)) & 255u;
        const uint32_t lq = (S.r_lowq >> (8 * f)) & 255u, pl = (S.r_r2ple >> (8 * f)) & 255u;
        const uint32_t cc = (S.r_conc >> (8 * f)) & 255u;
        if (al) FCW(KW_ALLELE_FWD, a) += (int)(al | (fw << 16));
        if (t1) FCW(KW_R1, a) += (int)(t1 | (l1 << 16));
        if (t2) FCW(KW_R2, a) += (int)(t2 | (l2 << 16));
        if (lq | pl) FCW(KW_LOWQ_R2P, a) += (int)(lq | (pl << 16));
        if (cc) FCW(KW_PAIR, a) += (int)cc;
    }
    S.r_allele = S.r_fwd = S.r_lowq = S.r_r1tot = S.r_r1le = S.r_r2tot = S.r_r2le = S.r_r2ple = S.r_conc = 0;
}

// order-independent tallies of a regular A/C/G/T event (smCounter.py:423-459) into the register counters;
// `one` = the field increment of the base (0 when the event does not count), `onei` = the same if it passes incCond
__device__ __forceinline__ void tally_regular(LaneState& S, uint32_t one, uint32_t onei, bool reverse, bool read2, bool lowq, bool le20, bool ple) {
    S.r_allele += one;                                               // :459
    if (!reverse) S.r_fwd += one;                                    // :454-457
    if (lowq) S.r_lowq += one;                                       // :428-429
    const uint32_t onel = le20 ? onei : 0u;
    if (!read2) { S.r_r1tot += onei; S.r_r1le += onel; }             // :432-441
    else { S.r_r2tot += onei; S.r_r2le += onel; }                    // :442-452
    if (ple) S.r_r2ple += onei;                                      // the primer window of an R1 read is empty
}

// Tallies of one pileup event whose base is not A/C/G/T (N / IUPAC, smCounter.py:423-457 with that key): rare, so it goes
// straight to the dynamic-allele row with atomics.  Returns the row | (nibble == N) << 31.
__device__ __noinline__ uint32_t dyn_base_event(DynTab T, const uint8_t* seq_read, uint32_t locus, uint32_t read_idx, int qpos,
                                                uint32_t flags /* 1 fwd 2 lowq 4 inc 8 r2 16 le20 32 ple */) {
    const uint32_t sb = __ldg(seq_read + (qpos >> 1));
    const uint32_t nib = (qpos & 1) ? (sb & 15u) : (sb >> 4);
    const uint32_t e = dyn_lookup(T, dyn_make_key(locus, SMC_K_BASE, nib, 0ull), read_idx, qpos, 0);
    int32_t* row = T.dcnt + (size_t)e * SMC_NCNT;
    atomicAdd(&row[SMC_C_ALLELE], 1);
    if (flags & 1u) atomicAdd(&row[SMC_C_FWD], 1);
    if (flags & 2u) atomicAdd(&row[SMC_C_LOWQ], 1);
    if (flags & 4u) {
        if (!(flags & 8u)) { atomicAdd(&row[SMC_C_R1TOT], 1); if (flags & 16u) atomicAdd(&row[SMC_C_R1LE], 1); }
        else {
            atomicAdd(&row[SMC_C_R2TOT], 1);
            if (flags & 16u) atomicAdd(&row[SMC_C_R2LE], 1);
            if (flags & 32u) atomicAdd(&row[SMC_C_R2PLE], 1);
        }
    }
    return e | (nib == 15u ? 0x80000000u : 0u);
}

// Pileup event of a read that is NOT one plain aligned run (indels, hard clips, ...): htslib resolve_cigar2 for the
// lane's position p, then the allele classification of smCounter.py:371-457.  Returns {event code, x}:
//   regular base          : EC_REGULAR, nibble in the code (EC_DYN when it is not A/C/G/T; then x = query position)
//   inside a deletion     : x = MID_DEL, bq = minBQ (:416-421)
//   insertion / deletion start (:371-411): x = MID_DYN + row; alleleCnt and strand are tallied here
__device__ __noinline__ uint2 slow_event(DynTab T, const uint32_t* rw /* staged ReadRec */, const uint32_t* __restrict__ cigar,
                                         const uint8_t* __restrict__ seqp, const uint8_t* __restrict__ qualp, int32_t p, int32_t Li,
                                         int minBQ, int primerDist) {
    const uint32_t meta = rw[3];
    const int32_t start = (int32_t)rw[6], lo = (int32_t)rw[0], hi = (int32_t)rw[7];
    if (!(Li >= lo && Li < hi)) return make_uint2(0u, 0u);
    const bool reverse = meta & RM_REVERSE, read2 = meta & RM_READ2;
    const uint32_t ncig = meta >> 8;
    const int leftSP = (int)(rw[2] & 0xffffu), alnlen = (int)(rw[2] >> 16);
    const uint32_t seq_off = rw[4], qual_off = rw[5], cigar_off = rw[11];
    int qpos = 0, indel = 0; bool isdel = false;
    {
        int x = start, y = 0;
        for (uint32_t k = 0; k < ncig; ++k) {
            const uint32_t cw = k < 4 ? rw[12 + k] : __ldg(&cigar[cigar_off + k]);
            const uint32_t op = cw & 15u; const int len = (int)(cw >> 4);
            if (op == 0 || op == 7 || op == 8 || op == 2 || op == 3) {
                if (p < x + len) {                                           // the op that covers p
                    isdel = (op == 2 || op == 3);
                    qpos = isdel ? y : y + (p - x);
                    if (p == x + len - 1 && k + 1 < ncig) {                   // peek the next op
                        const uint32_t c2 = (k + 1) < 4 ? rw[12 + k + 1] : __ldg(&cigar[cigar_off + k + 1]);
                        const uint32_t op2 = c2 & 15u; const int l2 = (int)(c2 >> 4);
                        if (op2 == 2) indel = -l2;
                        else if (op2 == 1) indel = l2;
                        else if (op2 == 6 && k + 2 < ncig) {
                            int l3 = 0;
                            for (uint32_t kk = k + 2; kk < ncig; ++kk) {
                                const uint32_t c3 = kk < 4 ? rw[12 + kk] : __ldg(&cigar[cigar_off + kk]);
                                const uint32_t op3 = c3 & 15u;
                                if (op3 == 1) l3 += (int)(c3 >> 4);
                                else if (op3 == 2 || op3 == 0 || op3 == 3 || op3 == 7 || op3 == 8) break;
                            }
                            if (l3 > 0) indel = l3;
                        }
                    }
                    break;
                }
                x += len;
                if (op == 0 || op == 7 || op == 8) y += len;
            } else if (op == 1 || op == 4) y += len;
        }
    }
    if (indel == 0 && isdel) {                                                 // :416-421
        const bool inc = (meta & RM_OK);                                       // bq = minBQ passes the quality gate
        return make_uint2(((uint32_t)minBQ & 255u) | EC_COVERED | (inc ? EC_INC : 0u), MID_DEL);
    }
    const uint32_t sb = __ldg(seqp + ((size_t)seq_off + (size_t)(qpos >> 1)));
    const uint32_t nib = (qpos & 1) ? (sb & 15u) : (sb >> 4);
    const uint32_t bq = __ldg(qualp + ((size_t)qual_off + (size_t)qpos));
    const bool lowq = (int)bq < minBQ;
    const bool inc = !lowq && (meta & RM_OK);                                  // :378,400,431
    uint32_t code = bq | (nib << EC_NIB_SH) | EC_COVERED | (inc ? EC_INC : 0u);
    if (indel == 0) {                                                          // :423-457 regular base
        const int d = qpos - leftSP;
        const int da = reverse ? alnlen - d : d, db = reverse ? d : alnlen - d;
        if ((read2 ? db : da) <= 20) code |= EC_LE20;
        if (read2 && da <= primerDist) code |= EC_PLE;
        code |= EC_REGULAR;
        if (nib_is_acgt(nib)) return make_uint2(code, 0u);
        return make_uint2(code | EC_DYN, (uint32_t)qpos);
    }
    // insertion start (:371-389) or deletion start (:392-411): alleleCnt and strand only
    unsigned long long key; int len;
    if (indel > 0) {
        len = indel;
        unsigned long long payload;
        if (len <= 8) {
            unsigned long long nibs = 0;
            for (int t = 0; t < len; ++t) {
                const int qq = qpos + 1 + t;
                const uint32_t b2 = __ldg(seqp + ((size_t)seq_off + (size_t)(qq >> 1)));
                nibs |= (unsigned long long)((qq & 1) ? (b2 & 15u) : (b2 >> 4)) << (28 - 4 * t);
            }
            payload = ((unsigned long long)len << 32) | nibs;
        } else {
            uint32_t hsh = 2166136261u ^ (uint32_t)len;
            for (int t = 0; t < len; ++t) {
                const int qq = qpos + 1 + t;
                const uint32_t b2 = __ldg(seqp + ((size_t)seq_off + (size_t)(qq >> 1)));
                hsh = (hsh ^ ((qq & 1) ? (b2 & 15u) : (b2 >> 4))) * 16777619u;
            }
            payload = (15ull << 32) | hsh;
        }
        key = dyn_make_key((uint32_t)Li, SMC_K_INS, nib, payload);
    } else {
        len = -indel;
        key = dyn_make_key((uint32_t)Li, SMC_K_DEL, nib, (unsigned long long)len);
    }
    const uint32_t e = dyn_lookup(T, key, rw[10], qpos, len);
    atomicAdd(&T.dcnt[(size_t)e * SMC_NCNT + SMC_C_ALLELE], 1);
    if (!reverse) atomicAdd(&T.dcnt[(size_t)e * SMC_NCNT + SMC_C_FWD], 1);
    return make_uint2(code, MID_DYN + e);
}

// first use of the per-barcode shared-memory arrays: a barcode that has shown a single allele so far keeps its state in
// registers only (its product over fragments IS rightP, its count IS n)
__device__ __forceinline__ void umi_materialize(int lane, int* ucnt, double* uprod, uint32_t exist, int n, double rightP) {
    const int s0 = __ffs(exist) - 1;
    UCNT(s0) = n; UPROD(s0) = rightP;
}

__device__ __forceinline__ void fragment_finalize(const double* bqtab_s, int lane, int* ucnt, double* uprod, LaneState& S) {
    if (S.flags & LF_FRAG_SEEN) { S.allFrag++; S.flags = (S.flags & ~LF_FRAG_SEEN) | LF_UMI_SEEN; }      // :463-464
    if (!(S.flags & LF_F_EXISTS)) return;
    const double p = (S.flags & LF_F_PAIRED) ? bqtab_s[S.f_bq] : 0.1;  // smCounter.py:65-68
    S.flags &= ~LF_F_EXISTS;
    const uint32_t aid = mid_to_aid(S.f_mid);
    int slot = (int)aid;
    if (aid >= NF) {
        const uint32_t e = aid - NF;
        if (S.ndyn > 0 && S.udyn0 == e) slot = 5;
        else if (S.ndyn > 1 && S.udyn1 == e) slot = 6;
        else if (S.ndyn == 0) { S.udyn0 = e; S.ndyn = 1; slot = 5; }
        else if (S.ndyn == 1) { S.udyn1 = e; S.ndyn = 2; slot = 6; }
        else { S.status |= SMC_ST_UMI_OVERFLOW; slot = 5; }
    }
    const double q1 = 1.0 - p;
    const uint32_t bit = 1u << slot;
    if (S.exist == 0) S.exist = bit;
    else {
        const bool multi = (S.exist & (S.exist - 1u)) != 0u;
        if (multi || S.exist != bit) {
            if (!multi) umi_materialize(lane, ucnt, uprod, S.exist, S.n, S.rightP);
            if (!(S.exist & bit)) { S.exist |= bit; UCNT(slot) = 0; UPROD(slot) = S.Q; }
            uint32_t m = S.exist;
            while (m) {                                                  // :70-74
                int s = __ffs(m) - 1; m &= m - 1;
                UPROD(s) = __dmul_rn(UPROD(s), s == slot ? q1 : p);
            }
            UCNT(slot) += 1;
        }
    }
    S.Q = __dmul_rn(S.Q, p);
    S.rightP = __dmul_rn(S.rightP, q1);                              // :77
    S.n += 1;
    S.last_aid = aid;
}

// PCR prior outside the host-built table (barcodes with > pcr_nmax fragments or > 6 distinct alleles): device pow()
__device__ __noinline__ double pcr_slow(int cnt, double denom) {
    return pow(10.0, -6.0 * (((double)cnt + 0.5) / denom));
}

__device__ __forceinline__ double neg_log10_1m(double p) {              // smCounter.py:509-510
    const double x = 1.0 - p;
    return x > 0.0 ? -log10(x) : 16.0;
}

// -log10(x) for x = fl(1 - p) when p is tiny: with q = 1 - x (exact, Sterbenz) the series q + q^2/2 + ... + q^6/6 is within
// 2e-16 relative of -ln(x) for q < 2^-10, so the result agrees with log10(x) to the last bit or two -- and it costs a
// sixth of the library call.  (The posterior of a padded allele is ~1e-5: every barcode takes this branch once.)
__device__ __forceinline__ double neg_log10_1m_small(double p) {
    const double x = 1.0 - p;
    const double q = 1.0 - x;
    if (!(q < 0.0009765625)) return x > 0.0 ? -log10(x) : 16.0;
    double s = fma(q, 1.0 / 6.0, 0.2);
    s = fma(s, q, 0.25);
    s = fma(s, q, 1.0 / 3.0);
    s = fma(s, q, 0.5);
    s = fma(s, q, 1.0);
    return (s * q) * 0.43429448190325182765;             // 1 / ln(10)
}

// calProb + consensus for a barcode that shows several alleles, a DEL / dynamic allele, or more fragments than the prior
// table holds (smCounter.py:26-98, 506-523) -- the general form; the per-barcode arrays are in shared memory.
// Returns the finalDict keys it touched among the fixed alleles.
__device__ __noinline__ uint32_t umi_general(DynTab T, const double* __restrict__ pcrtab, int pcr_nmax, double smt, int lane, int n,
                                             uint32_t exist, double rightP, int ndyn, uint32_t udyn0, uint32_t udyn1, uint32_t last_aid,
                                             int* fc, ulonglong2* limb, int* ucnt, double* uprod) {
    uint32_t keymask = 0;
    // canonical order of the dynamic slots = ascending allele key
    if (ndyn == 2 && __ldcg(&T.dkey[udyn0]) > __ldcg(&T.dkey[udyn1])) {
        uint32_t t = udyn0; udyn0 = udyn1; udyn1 = t;
        int c5 = UCNT(5), c6 = UCNT(6); double p5 = UPROD(5), p6 = UPROD(6);
        uint32_t b5 = (exist >> 5) & 1u, b6 = (exist >> 6) & 1u;
        UCNT(5) = c6; UCNT(6) = c5; UPROD(5) = p6; UPROD(6) = p5;
        exist = (exist & 0x1fu) | (b6 << 5) | (b5 << 6);
    }
    int k = __popc(exist);
    uint32_t pad = 0;                             // :49-54  pad with A, T, G, C until 4
    if (k < 4 && !((exist >> SMC_A_A) & 1u)) { pad |= 1u << SMC_A_A; ++k; }
    if (k < 4 && !((exist >> SMC_A_T) & 1u)) { pad |= 1u << SMC_A_T; ++k; }
    if (k < 4 && !((exist >> SMC_A_G) & 1u)) { pad |= 1u << SMC_A_G; ++k; }
    if (k < 4 && !((exist >> SMC_A_C) & 1u)) { pad |= 1u << SMC_A_C; ++k; }
    const uint32_t uniq = exist | pad;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    // ---- PCR prior of every allele of the barcode (:79-81): table lookups (host glibc pow) or pow()
    const bool tab = (k <= 6 && n <= pcr_nmax);
    const double* trow = pcrtab + ((size_t)(k - 4) * ((size_t)(pcr_nmax + 1) * (pcr_nmax + 2) / 2) + (size_t)n * (n + 1) / 2);
    const double denom = (double)n + 0.5 * (double)k;
    const double pcr_pad = tab ? __ldg(trow) : pcr_slow(0, denom);
    double m1 = pad ? pcr_pad : INF, m2 = INF; int arg1 = -1;      // smallest / second smallest prior and its slot
    double tpad = rightP;                                         // :88-91
    for (uint32_t m = exist; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        const int c = UCNT(s);
        const double v = tab ? __ldg(trow + c) : pcr_slow(c, denom);
        tpad = __dmul_rn(tpad, v);
        if (v < m1) { m2 = m1; m1 = v; arg1 = s; } else if (v < m2) m2 = v;
    }
    // ---- likelihood of each present allele (:86), stored over its (no longer needed) product; pads all share tpad
    const double PCR_NO_ERROR = 1.0 - 3e-5;                       // smCounter.py:20
    for (uint32_t m = exist; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        const double minp = (s == arg1) ? m2 : m1;                // min over the OTHER members of uniq
        UPROD(s) = __dadd_rn(__dmul_rn(PCR_NO_ERROR, UPROD(s)), __dmul_rn(rightP, minp));
    }
    double sumP = 0.0;                                            // :93, in canonical slot order
    for (uint32_t m = uniq; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        sumP = __dadd_rn(sumP, ((exist >> s) & 1u) ? UPROD(s) : tpad);
    }
    // ---- posterior -> -log10(1-p) (:96, :509-510), PI accumulation (:512), consensus (:514-523).
    // One loop over the members of uniq in slot order; the pads share one value (computed at the first pad).
    double best = -1.0, l_pad = 0.0; int nbest = 0, cons = -1; bool have_pad = false;
    unsigned long long plo = 0, phi = 0;
#pragma unroll 1
    for (uint32_t m = uniq; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        const bool is_pad = (pad >> s) & 1u;
        double l; unsigned long long lo, hi;
        if (is_pad && have_pad) { l = l_pad; lo = plo; hi = phi; }
        else {
            l = neg_log10_1m(sumP <= 0.0 ? 0.0 : (is_pad ? tpad : UPROD(s)) / sumP);
            pi_fixed128(l, lo, hi);
            if (is_pad) { have_pad = true; l_pad = l; plo = lo; phi = hi; }
        }
        if (s < NF) {
            ulonglong2 v = LIMB(s);
            add128(v.x, v.y, lo, hi);
            LIMB(s) = v;
            keymask |= 1u << s;
        } else {
            const uint32_t e = s == 5 ? udyn0 : udyn1;
            unsigned long long a0, a1, a2;
            split_limbs(lo, hi, a0, a1, a2);
            if (a0) atomicAdd(&T.dlimb[(size_t)e * 3 + 0], a0);
            if (a1) atomicAdd(&T.dlimb[(size_t)e * 3 + 1], a1);
            if (a2) atomicAdd(&T.dlimb[(size_t)e * 3 + 2], a2);
            T.diskey[e] = 1;
        }
        if (l > best) { best = l; nbest = 1; cons = s; }
        else if (l == best) nbest++;
    }
    if (nbest == 1) {                                             // :515-519
        const uint32_t aid = cons < NF ? (uint32_t)cons : NF + (cons == 5 ? udyn0 : udyn1);
        bump(T.dcnt, fc, lane, aid, KW_MT, best > smt ? 0x10001u : 1u, SMC_C_MT, SMC_C_STRONG);
    } else if (n == 1) {                                          // :521-523
        bump(T.dcnt, fc, lane, last_aid, KW_MT, 1u, SMC_C_MT, SMC_C_STRONG);
    }
    return keymask;
}

// calProb + the per-barcode part of vc() (smCounter.py:26-98, 506-532) for the lane's locus.
template <bool LIST>
__device__ __forceinline__ void umi_finalize(const K3Args& A, int lane, int ki, int li, uint32_t urank, int* fc, ulonglong2* limb,
                                             int* ucnt, double* uprod, LaneState& S) {
    if (S.flags & LF_UMI_SEEN) S.allMT++;
    bool used = S.flags & LF_UMI_BC;
    if (used) {
        S.nBC++;
        if (ki >= 0) {                                   // down-sampling mask (smCounter.py:496-500)
            unsigned long long u = A.umi_of_urank[urank];
            int64_t lo = A.keep_off[ki], hi = A.keep_off[ki + 1];
            int64_t pos = lower_bound_u64((const uint64_t*)A.keep_umi + lo, hi - lo, u);
            used = (pos < hi - lo) && (A.keep_umi[lo + pos] == u);
        }
        if (LIST && li >= 0) {
            uint32_t slot = atomicAdd(&A.list_count[li], 1u);
            int64_t o = A.list_off[li] + slot;
            if (o < A.list_off[li + 1] && o < A.list_cap) { A.list_umi[o] = A.umi_of_urank[urank]; A.list_first[o] = S.first_read; }
        }
    }
    if (used) {
        const int n = S.n;
        S.usedMT++; S.usedFrag += n;
        S.mt3 += n >= 3; S.mt5 += n >= 5; S.mt7 += n >= 7; S.mt10 += n >= 10;
        const bool multi = (S.exist & (S.exist - 1u)) != 0u;
        const uint32_t ACGT = (1u << SMC_A_A) | (1u << SMC_A_T) | (1u << SMC_A_G) | (1u << SMC_A_C);
        if (n <= A.mtDrop) {                              // :28-32 -> four zeros, a 4-way tie (:514-523)
            S.keymask |= ACGT;
            if (n == 1) bump(A.dcnt, fc, lane, S.last_aid, KW_MT, 1u, SMC_C_MT, SMC_C_STRONG);
        } else if (!multi && (S.exist & ACGT) && n <= A.pcr_nmax) {
            // ---- fast path: every fragment of the barcode shows the same base a0 in {A,C,G,T}; uniq = {A,C,G,T} (:49-54).
            // Same operations in the same order as umi_general (prodP[a0] == rightP, one present allele, three pads
            // sharing one value), with all intermediates in registers.
            const int a0 = __ffs(S.exist) - 1;
            const double rightP = S.rightP;
            const double* trow = A.pcrtab + (size_t)n * (n + 1) / 2;                       // k = 4
            const double v_e = __ldg(trow + n), v_pad = __ldg(trow);                       // :79-81
            const double tpad = __dmul_rn(rightP, v_e);                                    // :88-91
            const double PCR_NO_ERROR = 1.0 - 3e-5;
            const double t_e = __dadd_rn(__dmul_rn(PCR_NO_ERROR, rightP), __dmul_rn(rightP, v_pad));   // :86 (min over the others = a pad's prior)
            const int posidx = a0 - (a0 > SMC_A_DEL ? 1 : 0);                              // rank of a0 in slot order A C T G
            double sumP = posidx == 0 ? t_e : tpad;                                        // :93 (0.0 + x == x)
            sumP = __dadd_rn(sumP, posidx == 1 ? t_e : tpad);
            sumP = __dadd_rn(sumP, posidx == 2 ? t_e : tpad);
            sumP = __dadd_rn(sumP, posidx == 3 ? t_e : tpad);
            const bool pos = sumP > 0.0;
            const double l_pad = neg_log10_1m_small(pos ? tpad / sumP : 0.0);              // :96, :509-510
            const double l_e = neg_log10_1m(pos ? t_e / sumP : 0.0);
            // PI: l_pad goes to all four bases through the register accumulator, a0 gets the difference (mod 2^128)
            unsigned long long plo, phi, elo, ehi;
            pi_fixed128(l_pad, plo, phi);
            pi_fixed128(l_e, elo, ehi);
            add128(S.pad_lo, S.pad_hi, plo, phi);
            sub128(elo, ehi, plo, phi);
            ulonglong2 v = LIMB(a0);
            add128(v.x, v.y, elo, ehi);
            LIMB(a0) = v;
            S.keymask |= ACGT;
            // consensus (:514-523): three pads tie at l_pad; a0 wins iff l_e > l_pad
            if (l_e > l_pad) FCW(KW_MT, a0) += (l_e > A.smt) ? 0x10001 : 1;
            else if (n == 1) FCW(KW_MT, a0) += 1;
        } else {
            if (!multi) umi_materialize(lane, ucnt, uprod, S.exist, n, S.rightP);
            S.keymask |= umi_general(dyn_tab(A), A.pcrtab, A.pcr_nmax, A.smt, lane, n, S.exist, S.rightP, S.ndyn, S.udyn0, S.udyn1,
                                     S.last_aid, fc, limb, ucnt, uprod);
        }
    }
    S.n = 0; S.exist = 0; S.Q = 1.0; S.rightP = 1.0; S.ndyn = 0; S.flags &= ~(LF_UMI_SEEN | LF_UMI_BC); S.first_read = 0xffffffffu;
}

// first event index >= x (x > tb) at which the barcode changes, or te
__device__ __forceinline__ uint32_t chunk_boundary(const K3Args& A, uint32_t x, uint32_t te, int lane) {
    while (x < te) {
        uint32_t i = x + lane;
        bool b = true;
        if (i < te) b = A.recs[A.ev_read[i]].urank != A.recs[A.ev_read[i - 1]].urank;
        uint32_t m = __ballot_sync(FULL_MASK, b);
        if (m) { uint32_t r = x + (__ffs(m) - 1); return r < te ? r : te; }
        x += 32;
    }
    return te;
}

// add the lane's packed shared-memory counters to the global 32-bit accumulators and clear them
__device__ __noinline__ void flush_counters(int32_t* cnt, size_t nl, int* fc, int lane, int64_t L) {
    const int lo_idx[K3_NW] = {SMC_C_ALLELE, SMC_C_R1TOT, SMC_C_R2TOT, SMC_C_LOWQ, SMC_C_CONCORD, SMC_C_MT};
    const int hi_idx[K3_NW] = {SMC_C_FWD, SMC_C_R1LE, SMC_C_R2LE, SMC_C_R2PLE, SMC_C_DISCORD, SMC_C_STRONG};
#pragma unroll
    for (int a = 0; a < NF; ++a) {
#pragma unroll
        for (int w = 0; w < K3_NW; ++w) {
            const uint32_t v = (uint32_t)FCW(w, a);
            if (v) {
                FCW(w, a) = 0;
                const int lo = (int)(v & 0xffffu), hi = (int)(v >> 16);
                if (lo) atomicAdd(&cnt[((size_t)a * SMC_NCNT + lo_idx[w]) * nl + L], lo);
                if (hi) atomicAdd(&cnt[((size_t)a * SMC_NCNT + hi_idx[w]) * nl + L], hi);
                if (w == KW_ALLELE_FWD && a != SMC_A_DEL && lo - hi) atomicAdd(&cnt[((size_t)a * SMC_NCNT + SMC_C_REV) * nl + L], lo - hi);
            }
        }
    }
}

template <bool LIST>
__global__ void __launch_bounds__(K3_WARPS * 32, K3_MINBLOCKS) k_pileup_t(const K3Args A) {
    extern __shared__ __align__(16) uint32_t smem[];
    double* bqtab_s = reinterpret_cast<double*>(smem);
    uint32_t* one_lut = smem + 512;                                      // nibble -> 1 << 8 * field (0 for non-ACGT)
    for (int i = threadIdx.x; i < 256; i += K3_WARPS * 32) bqtab_s[i] = __ldg(&A.bqtab[i]);
    if (threadIdx.x < 16) one_lut[threadIdx.x] = nib_is_acgt(threadIdx.x) ? (1u << (8u * nib_field(threadIdx.x))) : 0u;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* ws = smem + K3_TAB_BYTES / 4 + (size_t)w * K3_WARP_WORDS;
    uint16_t* codes = reinterpret_cast<uint16_t*>(ws + K3_STAGE_WORDS);
    int* fc = (int*)(ws + K3_STAGE_WORDS + K3_CODE_WORDS);
    ulonglong2* limb = (ulonglong2*)(ws + K3_STAGE_WORDS + K3_CODE_WORDS + K3_FC_WORDS);
    int* ucnt = (int*)(ws + K3_STAGE_WORDS + K3_CODE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS);
    double* uprod = (double*)(ws + K3_STAGE_WORDS + K3_CODE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS + K3_UCNT_WORDS);

    const uint32_t unit = blockIdx.x * K3_WARPS + w;
    if (unit >= A.unit_off[A.n_tiles]) return;
    const uint32_t tile = (uint32_t)upper_slot_u32(A.unit_off, (int64_t)A.n_tiles + 1, unit);
    const uint32_t c = unit - A.unit_off[tile];
    const uint32_t tb = A.tile_off[tile], te = A.tile_off[tile + 1];
    uint32_t eb = tb + c * A.chunk, ee = tb + (c + 1) * A.chunk;
    eb = c == 0 ? tb : chunk_boundary(A, eb, te, lane);
    ee = ee >= te ? te : chunk_boundary(A, ee, te, lane);
    if (eb >= ee) return;

    const int64_t L = (int64_t)tile * 32 + lane;
    const bool lane_valid = L < A.n_loci;
    const int32_t p = lane_valid ? A.loci_pos[L] : 0;
    const int32_t Li = lane_valid ? (int32_t)L : -1;                     // -1 is never inside a read's [lo, hi)
    const int ki = (lane_valid && A.keep_idx) ? A.keep_idx[L] : -1;
    const int li = (LIST && lane_valid) ? A.list_idx[L] : -1;

    for (int i = lane; i < K3_FC_WORDS + K3_LIMB_WORDS; i += 32) ws[K3_STAGE_WORDS + K3_CODE_WORDS + i] = 0;
    __syncwarp();

    LaneState S;
    S.cvg = S.allFrag = S.allMT = S.usedFrag = S.nBC = S.usedMT = S.mt3 = S.mt5 = S.mt7 = S.mt10 = 0;
    S.keymask = 0; S.status = 0; S.pad_lo = S.pad_hi = 0;
    S.r_allele = S.r_fwd = S.r_lowq = S.r_r1tot = S.r_r1le = S.r_r2tot = S.r_r2le = S.r_r2ple = S.r_conc = 0;
    S.flags = 0;
    S.n = 0; S.exist = 0; S.Q = 1.0; S.rightP = 1.0; S.last_aid = 0; S.udyn0 = S.udyn1 = 0; S.ndyn = 0;
    S.first_read = 0xffffffffu;
    S.f_mid = 0; S.f_bq = 0;

    uint32_t carry_urank = 0xffffffffu, carry_frank = 0xffffffffu;       // barcode / fragment of the last read of the previous batch
    const int minBQ = A.minBQ;
    const uint8_t* __restrict__ seqp = A.seq;
    const uint8_t* __restrict__ qualp = A.qual;
    uint32_t since_flush = 0, since_reg_flush = 0;

    for (uint32_t base = eb; base < ee; base += 32) {
        const int nb = (int)min(32u, ee - base);
        // ---------------- stage the next 32 read records in shared memory (4 x 128-bit loads per lane; slots past the
        // end of the unit get an empty record); boundaries and per-read flags as warp-uniform bit masks
        uint32_t my_urank = 0xfffffffdu, my_frank = 0xfffffffdu, my_meta = 0;
        {
            uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0, r2 = r0, r3 = r0;
            if (lane < nb) {
                const uint4* src = reinterpret_cast<const uint4*>(&A.recs[A.ev_read[base + lane]]);
                r0 = __ldg(src); r1 = __ldg(src + 1); r2 = __ldg(src + 2); r3 = __ldg(src + 3);
                my_urank = r2.x; my_frank = r2.y; my_meta = r0.w;
            }
            uint4* dst = reinterpret_cast<uint4*>(ws + lane * 16);
            dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
        }
        uint32_t pu = __shfl_up_sync(FULL_MASK, my_urank, 1), pf = __shfl_up_sync(FULL_MASK, my_frank, 1);
        if (lane == 0) { pu = carry_urank; pf = carry_frank; }
        const bool first_ever = (base == eb) && lane == 0;               // nothing is open before the first read of the unit
        const uint32_t valid = nb == 32 ? FULL_MASK : ((1u << nb) - 1u);
        const uint32_t fragmask = __ballot_sync(FULL_MASK, my_frank != pf && !first_ever) & valid;
        const uint32_t umimask = __ballot_sync(FULL_MASK, my_urank != pu && !first_ever) & valid;
        const uint32_t simplemask = __ballot_sync(FULL_MASK, my_meta & RM_SIMPLE) & valid;
        const uint32_t batch_prev_urank = carry_urank;                    // barcode that is open when this batch starts
        carry_urank = __shfl_sync(FULL_MASK, my_urank, nb - 1); carry_frank = __shfl_sync(FULL_MASK, my_frank, nb - 1);
        since_flush += 32; since_reg_flush += 32;
        if (since_reg_flush > K3_REG_FLUSH) { flush_regs(fc, lane, S); since_reg_flush = 32; }
        if (since_flush > K3_FLUSH_EVERY) {
            flush_regs(fc, lane, S);
            if (lane_valid) flush_counters(A.cnt, (size_t)A.n_loci, fc, lane, L);
            since_flush = 32; since_reg_flush = 32;
        }
        __syncwarp();
        // ---------------- pass A: gather base + quality of my locus, K3_GATHER reads at a time, and tally (simple reads only;
        // the gather span of any other record is 0)
#pragma unroll 1
        for (int g = 0; g < nb; g += K3_GATHER) {
            uint32_t sbv[K3_GATHER], bqv[K3_GATHER], fl[K3_GATHER];
#pragma unroll
            for (int u = 0; u < K3_GATHER; ++u) {
                const uint32_t* rw = ws + (g + u) * 16;
                const uint4 qa = *reinterpret_cast<const uint4*>(rw);        // lo gspan qk meta
                const uint2 qb = *reinterpret_cast<const uint2*>(rw + 4);    // seq_off qual_off
                const uint4 qw = *reinterpret_cast<const uint4*>(rw + 12);   // le_lo le_span ple_lo ple_span
                const bool cov = (uint32_t)(Li - (int32_t)qa.x) < qa.y;
                const uint32_t qpos = (uint32_t)(p + (int32_t)qa.z);
                sbv[u] = 0; bqv[u] = 0;
                if (cov) {
                    sbv[u] = __ldg(seqp + (qb.x + (qpos >> 1)));
                    bqv[u] = __ldg(qualp + (qb.y + qpos));
                }
                const bool le20 = (uint32_t)(p - (int32_t)qw.x) <= qw.y;
                const bool ple = (uint32_t)(p - (int32_t)qw.z) <= qw.w;
                // bits 0-3 RM_*, 4 covered, 5 le20, 6 ple, 7 odd query position
                fl[u] = (qa.w & 15u) | (cov ? 16u : 0u) | (le20 ? 32u : 0u) | (ple ? 64u : 0u) | ((qpos & 1u) << 7);
            }
#pragma unroll
            for (int u = 0; u < K3_GATHER; ++u) {
                const uint32_t f = fl[u];
                const uint32_t nib = (f & 128u) ? (sbv[u] & 15u) : (sbv[u] >> 4);   // 0 when not covered
                const uint32_t bq = bqv[u];
                const bool cov = f & 16u;
                const uint32_t one = one_lut[nib];                                   // 0 when not covered or not A/C/G/T
                const bool lowq = (int)bq < minBQ;
                const bool inc = cov && !lowq && (f & RM_OK);                        // :431
                S.cvg += cov ? 1 : 0;                                                // :368
                tally_regular(S, one, inc ? one : 0u, f & RM_REVERSE, f & RM_READ2, lowq, f & 32u, f & 64u);
                codes[(g + u) * 32 + lane] = (uint16_t)(bq | (nib << EC_NIB_SH) | (cov ? EC_COVERED : 0u) | ((cov && !one) ? EC_DYN : 0u) |
                                                        (inc ? EC_INC : 0u));
            }
        }
        __syncwarp();
        // ---------------- pass B: the order-dependent part
#pragma unroll 1
        for (int j = 0; j < nb; ++j) {
            const uint32_t* rw = ws + j * 16;
            // ---- barcode / fragment boundaries (warp uniform)
            if ((fragmask >> j) & 1u) fragment_finalize(bqtab_s, lane, ucnt, uprod, S);
            if ((umimask >> j) & 1u) umi_finalize<LIST>(A, lane, ki, li, j ? rw[8 - 16] : batch_prev_urank, fc, limb, ucnt, uprod, S);
            uint32_t cd, mid;
            const bool simple = (simplemask >> j) & 1u;
            if (simple) {
                cd = codes[j * 32 + lane];
                mid = (cd >> EC_NIB_SH) & 15u;
            } else {                                                             // rare: per-event CIGAR walk, out of line
                const uint2 ev = slow_event(dyn_tab(A), rw, A.cigar, seqp, qualp, p, Li, minBQ, A.primerDist);
                cd = ev.x; mid = ev.y;
                if (cd & EC_COVERED) {
                    S.cvg++;                                                     // :368
                    if ((cd & EC_REGULAR) && !(cd & EC_DYN)) {
                        mid = (cd >> EC_NIB_SH) & 15u;
                        const uint32_t one = one_lut[mid];
                        tally_regular(S, one, (cd & EC_INC) ? one : 0u, rw[3] & RM_REVERSE, rw[3] & RM_READ2, (int)(cd & 255u) < minBQ,
                                      cd & EC_LE20, cd & EC_PLE);
                    } else if (!(cd & EC_REGULAR) && mid == MID_DEL) FCW(KW_ALLELE_FWD, SMC_A_DEL) += 1;   // alleleCnt only (:416-421, :459)
                }
            }
            S.flags |= cd & EC_COVERED;                                          // LF_FRAG_SEEN (:463-464)
            bool isN = false;
            if (cd & EC_DYN) {                                                   // rare: N / IUPAC base -> dynamic allele row
                const uint32_t meta = rw[3];
                const int qpos = simple ? p + (int32_t)rw[2] : (int)mid;
                const bool le20 = simple ? (uint32_t)(p - (int32_t)rw[12]) <= rw[13] : (cd & EC_LE20) != 0u;
                const bool ple = simple ? (uint32_t)(p - (int32_t)rw[14]) <= rw[15] : (cd & EC_PLE) != 0u;
                const uint32_t fl = ((meta & RM_REVERSE) ? 0u : 1u) | ((int)(cd & 255u) < minBQ ? 2u : 0u) | ((cd & EC_INC) ? 4u : 0u) |
                                    ((meta & RM_READ2) ? 8u : 0u) | (le20 ? 16u : 0u) | (ple ? 32u : 0u);
                const uint32_t e = dyn_base_event(dyn_tab(A), seqp + rw[4], (uint32_t)Li, rw[10], qpos, fl);
                mid = MID_DYN + (e & 0x7fffffffu); isN = e >> 31;
            }
            if (cd & EC_INC) {                                                   // :467-479
                const int bq = (int)(cd & 255u);
                S.flags |= LF_UMI_BC;
                if (LIST) S.first_read = min(S.first_read, rw[10]);
                if (!(S.flags & LF_F_EXISTS)) { S.flags = (S.flags | LF_F_EXISTS) & ~LF_F_PAIRED; S.f_mid = mid; S.f_bq = bq; }
                else if (mid == S.f_mid || isN) {
                    S.f_bq = min(S.f_bq, bq); S.flags |= LF_F_PAIRED;
                    if (mid == S.f_mid) {
                        if (mid < MID_DEL) S.r_conc += one_lut[mid];
                        else bump(A.dcnt, fc, lane, mid_to_aid(mid), KW_PAIR, 1u, SMC_C_CONCORD, SMC_C_DISCORD);
                    }
                } else { S.flags &= ~LF_F_EXISTS; bump(A.dcnt, fc, lane, mid_to_aid(mid), KW_PAIR, 0x10000u, SMC_C_CONCORD, SMC_C_DISCORD); }
            }
        }
        __syncwarp();
    }
    // ---- close the last fragment and barcode of the unit
    fragment_finalize(bqtab_s, lane, ucnt, uprod, S);
    umi_finalize<LIST>(A, lane, ki, li, carry_urank, fc, limb, ucnt, uprod, S);
    // ---- flush the lane's locus to the per-locus accumulators ([field][locus] layout: coalesced across lanes)
    flush_regs(fc, lane, S);
    if (lane_valid) {
        const size_t nl = (size_t)A.n_loci;
        flush_counters(A.cnt, nl, fc, lane, L);
#pragma unroll
        for (int a = 0; a < NF; ++a) {
            ulonglong2 v = LIMB(a);
            if (a != SMC_A_DEL) add128(v.x, v.y, S.pad_lo, S.pad_hi);
            unsigned long long a0, a1, a2;
            split_limbs(v.x, v.y, a0, a1, a2);
            if (a0) atomicAdd(&A.limb[((size_t)a * 3 + 0) * nl + L], a0);
            if (a1) atomicAdd(&A.limb[((size_t)a * 3 + 1) * nl + L], a1);
            if (a2) atomicAdd(&A.limb[((size_t)a * 3 + 2) * nl + L], a2);
        }
        int32_t* loc = A.loc;
        if (S.cvg) atomicAdd(&loc[SMC_L_CVG * nl + L], S.cvg);
        if (S.allFrag) atomicAdd(&loc[SMC_L_ALLFRAG * nl + L], S.allFrag);
        if (S.allMT) atomicAdd(&loc[SMC_L_ALLMT * nl + L], S.allMT);
        if (S.usedFrag) atomicAdd(&loc[SMC_L_USEDFRAG * nl + L], S.usedFrag);
        if (S.nBC) atomicAdd(&loc[SMC_L_NBC * nl + L], S.nBC);
        if (S.usedMT) atomicAdd(&loc[SMC_L_USEDMT * nl + L], S.usedMT);
        if (S.mt3) atomicAdd(&loc[SMC_L_MT3 * nl + L], S.mt3);
        if (S.mt5) atomicAdd(&loc[SMC_L_MT5 * nl + L], S.mt5);
        if (S.mt7) atomicAdd(&loc[SMC_L_MT7 * nl + L], S.mt7);
        if (S.mt10) atomicAdd(&loc[SMC_L_MT10 * nl + L], S.mt10);
        if (S.keymask) atomicOr(&loc[SMC_L_KEYMASK * nl + L], (int)S.keymask);
        if (S.status) atomicOr(&loc[SMC_L_STATUS * nl + L], (int)S.status);
    }
}
