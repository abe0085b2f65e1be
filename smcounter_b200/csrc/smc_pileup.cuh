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
    int minMQ; double mismatchThr;
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
    rec.meta = (ok ? RM_OK : 0u) | ((fl & 0x10u) ? RM_REVERSE : 0u) | ((fl & 0x80u) ? RM_READ2 : 0u) |
               (simple ? RM_SIMPLE : 0u) | (ncig << 8);
    rec.sp_aln = (uint32_t)leftSP | ((uint32_t)alnlen << 16);
    rec.seq_off = (uint32_t)A.seq_off[r]; rec.qual_off = (uint32_t)A.qual_off[r]; rec.cigar_off = (uint32_t)co;
    rec.urank = A.urank[s]; rec.frank = A.frank[s];
    rec.cig[0] = c4[0]; rec.cig[1] = c4[1]; rec.cig[2] = c4[2]; rec.cig[3] = c4[3];
    rec.read_idx = r; rec.pad = 0;
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
// K3: tile pileup
// ------------------------------------------------------------------------------------------------------------
#ifndef K3_WARPS
#define K3_WARPS 4
#endif
#ifndef K3_MINBLOCKS
#define K3_MINBLOCKS 4
#endif
#define NF SMC_NFIXED
#define NSLOT 7            // per-barcode allele slots: 5 fixed + 2 dynamic
// Per-lane shared-memory counters of the fixed alleles, two 16-bit counters per word (flushed to the global 32-bit
// accumulators before any of them can reach 65536):
enum { KW_ALLELE_FWD = 0,   // alleleCnt | forwardCnt << 16
       KW_R1,               // len(r1BcEndPos) | #<=20 << 16
       KW_R2,               // len(r2BcEndPos) | #<=20 << 16
       KW_LOWQ_R2P,         // lowQReads | #r2PrimerEndPos<=primerDist << 16
       KW_PAIR,             // concordPairCnt | discordPairCnt << 16
       KW_MT,               // MTCnt | strongMTCnt << 16
       K3_NW };
#define K3_FLUSH_EVERY 49152u      // tile events between counter flushes (each event adds at most 1 to a field)
#define K3_STAGE_WORDS 512
#define K3_FC_WORDS    (NF * K3_NW * 32)
#define K3_LIMB_WORDS  (NF * 3 * 64)
#define K3_UCNT_WORDS  (NSLOT * 32)
#define K3_UPROD_WORDS (NSLOT * 64)
#define K3_UT_WORDS    (NSLOT * 64)
#define K3_WARP_WORDS  (K3_STAGE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS + K3_UCNT_WORDS + K3_UPROD_WORDS + K3_UT_WORDS)
#define K3_SMEM_BYTES  (K3_WARPS * K3_WARP_WORDS * 4)

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

// BAM nibble of A, C, G, T -> fixed slot (A0 C1 T3 G4), anything else -> -1
__device__ __forceinline__ int nib_to_fixed(uint32_t nib) {
    // nib 1,2,4,8 -> ffs 1,2,3,4 -> slots 0,1,4,3
    return (__popc(nib) == 1) ? (int)((0x3410u >> (4 * (__ffs(nib) - 1))) & 15u) : -1;
}

// open-addressing table of the non-ACGT/DEL alleles; arguments by value so that the kernel parameter block never has
// to be spilled to local memory for this (rare) call
__device__ __noinline__ uint32_t dyn_lookup(unsigned long long* dkey, uint32_t dmask, uint32_t* drep_read, int32_t* drep_qpos,
                                            int32_t* dlen, uint32_t* dcount, uint32_t* gflags, unsigned long long key,
                                            uint32_t rep_read, int rep_qpos, int len) {
    uint32_t h = hash64to32(key) & dmask;
    for (uint32_t probe = 0; probe <= dmask; ++probe) {
        unsigned long long cur = __ldcg(&dkey[h]);
        if (cur == key) return h;
        if (cur == DYN_EMPTY) {
            unsigned long long prev = atomicCAS(&dkey[h], DYN_EMPTY, key);
            if (prev == DYN_EMPTY) {
                drep_read[h] = rep_read; drep_qpos[h] = rep_qpos; dlen[h] = len;
                uint32_t c = atomicAdd(dcount, 1u);
                if (2ull * (c + 1ull) > (unsigned long long)dmask + 1ull) atomicOr(gflags, GF_DYN_FULL);
                return h;
            }
            if (prev == key) return h;
        }
        h = (h + 1) & dmask;
    }
    atomicOr(gflags, GF_DYN_FULL);
    return 0;
}

// split a non-negative double < 2^20 into three 44-bit limbs of a fixed-point number with LSB 2^-108
__device__ __forceinline__ void pi_limbs(double l, unsigned long long& a0, unsigned long long& a1, unsigned long long& a2) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(l);
    int e = (int)((bits >> 52) & 0x7ffull);
    unsigned long long m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    int sh = e - 967;                                   // value = m * 2^(e-1075) = (m << sh) * 2^-108
    if (e == 0 || sh <= -53) { a0 = a1 = a2 = 0; return; }
    unsigned long long lo, hi;
    if (sh <= 0) { lo = m >> (-sh); hi = 0; }
    else if (sh < 64) { lo = m << sh; hi = m >> (64 - sh); }
    else { lo = 0; hi = m << (sh - 64); }
    const unsigned long long M44 = (1ull << 44) - 1ull;
    a0 = lo & M44;
    a1 = ((lo >> 44) | (hi << 20)) & M44;
    a2 = hi >> 24;
}

struct LaneState {
    // locus-level
    int cvg, allFrag, allMT, usedFrag, nBC, usedMT, mt3, mt5, mt7, mt10;
    uint32_t keymask, status;
    // barcode-level
    int n; uint32_t exist; double Q, rightP; uint32_t last_aid; uint32_t udyn0, udyn1; int ndyn;
    bool umi_seen, umi_bc; uint32_t first_read;     // BAM index of the barcode's first passing read at this locus
    // fragment-level
    bool frag_seen, f_exists, f_paired; uint32_t f_aid; int f_bq;
};

#define FCW(w, a)   fc[((a) * K3_NW + (w)) * 32 + lane]
#define LIMB(a, j)  limb[((a) * 3 + (j)) * 32 + lane]
#define UCNT(s)     ucnt[(s) * 32 + lane]
#define UPROD(s)    uprod[(s) * 32 + lane]
#define UT(s)       ut[(s) * 32 + lane]

// counter update for an allele that may be dynamic: `word`/`add` address the packed shared-memory counter of a fixed
// allele, c_lo / c_hi are the smc_out counter indices the low / high half stand for.
__device__ __forceinline__ void bump(const K3Args& A, int* fc, int lane, uint32_t aid, int word, uint32_t add, int c_lo, int c_hi) {
    if (aid < NF) FCW(word, aid) += (int)add;
    else {
        int32_t* row = A.dcnt + (size_t)(aid - NF) * SMC_NCNT;
        if (add & 0xffffu) atomicAdd(&row[c_lo], 1);
        if (add >> 16) atomicAdd(&row[c_hi], 1);
    }
}

__device__ __forceinline__ void fragment_finalize(const K3Args& A, int lane, int* ucnt, double* uprod, LaneState& S) {
    if (S.frag_seen) { S.allFrag++; S.frag_seen = false; }
    if (!S.f_exists) return;
    S.f_exists = false;
    int slot;
    if (S.f_aid < NF) slot = (int)S.f_aid;
    else {
        uint32_t e = S.f_aid - NF;
        if (S.ndyn > 0 && S.udyn0 == e) slot = 5;
        else if (S.ndyn > 1 && S.udyn1 == e) slot = 6;
        else if (S.ndyn == 0) { S.udyn0 = e; S.ndyn = 1; slot = 5; }
        else if (S.ndyn == 1) { S.udyn1 = e; S.ndyn = 2; slot = 6; }
        else { S.status |= SMC_ST_UMI_OVERFLOW; slot = 5; }
    }
    double p = S.f_paired ? __ldg(&A.bqtab[S.f_bq]) : 0.1;          // smCounter.py:65-68
    double q1 = 1.0 - p;
    if (!((S.exist >> slot) & 1u)) { S.exist |= 1u << slot; UCNT(slot) = 0; UPROD(slot) = S.Q; }
    uint32_t m = S.exist;
    while (m) {                                                      // :70-74
        int s = __ffs(m) - 1; m &= m - 1;
        UPROD(s) = __dmul_rn(UPROD(s), s == slot ? q1 : p);
    }
    UCNT(slot) += 1;
    S.Q = __dmul_rn(S.Q, p);
    S.rightP = __dmul_rn(S.rightP, q1);                              // :77
    S.n += 1;
    S.last_aid = S.f_aid;
}

__device__ __forceinline__ void pi_add_limbs(const K3Args& A, int lane, unsigned long long* limb, LaneState& S, int slot,
                                             unsigned long long a0, unsigned long long a1, unsigned long long a2) {
    if (slot < NF) {
        LIMB(slot, 0) += a0; LIMB(slot, 1) += a1; LIMB(slot, 2) += a2;
        S.keymask |= 1u << slot;
    } else {
        uint32_t e = slot == 5 ? S.udyn0 : S.udyn1;
        if (a0) atomicAdd(&A.dlimb[(size_t)e * 3 + 0], a0);
        if (a1) atomicAdd(&A.dlimb[(size_t)e * 3 + 1], a1);
        if (a2) atomicAdd(&A.dlimb[(size_t)e * 3 + 2], a2);
        A.diskey[e] = 1;
    }
}

// PCR prior outside the host-built table (barcodes with > pcr_nmax fragments or > 6 distinct alleles): device pow()
__device__ __noinline__ double pcr_slow(int cnt, double denom) {
    return pow(10.0, -6.0 * (((double)cnt + 0.5) / denom));
}

__device__ __forceinline__ uint32_t slot_to_aid(const LaneState& S, int slot) {
    return slot < NF ? (uint32_t)slot : NF + (slot == 5 ? S.udyn0 : S.udyn1);
}

// calProb + the per-barcode part of vc() (smCounter.py:26-98, 506-532) for the lane's locus.
// The heavy FP64 work (PCR prior lookup, division, log10) runs over the COMPACTED list of alleles present in the
// barcode, so lanes whose loci have different reference bases still execute the same instructions; only the cheap,
// order-sensitive sums walk the slots in canonical order.
__device__ __forceinline__ void umi_finalize(const K3Args& A, int lane, int64_t L, uint32_t urank, int* fc, unsigned long long* limb,
                                             int* ucnt, double* uprod, double* ut, LaneState& S) {
    if (S.umi_seen) S.allMT++;
    bool used = S.umi_bc;
    if (used) {
        S.nBC++;
        int ki = A.keep_idx ? A.keep_idx[L] : -1;
        if (ki >= 0) {                                   // down-sampling mask (smCounter.py:496-500)
            unsigned long long u = A.umi_of_urank[urank];
            int64_t lo = A.keep_off[ki], hi = A.keep_off[ki + 1];
            int64_t pos = lower_bound_u64((const uint64_t*)A.keep_umi + lo, hi - lo, u);
            used = (pos < hi - lo) && (A.keep_umi[lo + pos] == u);
        }
        if (A.list_idx) {
            int li = A.list_idx[L];
            if (li >= 0) {
                uint32_t slot = atomicAdd(&A.list_count[li], 1u);
                int64_t o = A.list_off[li] + slot;
                if (o < A.list_off[li + 1] && o < A.list_cap) { A.list_umi[o] = A.umi_of_urank[urank]; A.list_first[o] = S.first_read; }
            }
        }
    }
    if (used) {
        const int n = S.n;
        S.usedMT++; S.usedFrag += n;
        S.mt3 += n >= 3; S.mt5 += n >= 5; S.mt7 += n >= 7; S.mt10 += n >= 10;
        if (n <= A.mtDrop) {                              // :28-32 -> four zeros, a 4-way tie (:514-523)
            S.keymask |= (1u << SMC_A_A) | (1u << SMC_A_T) | (1u << SMC_A_G) | (1u << SMC_A_C);
            if (n == 1) bump(A, fc, lane, S.last_aid, KW_MT, 1u, SMC_C_MT, SMC_C_STRONG);
        } else {
            // canonical order of the dynamic slots = ascending allele key
            if (S.ndyn == 2 && __ldcg(&A.dkey[S.udyn0]) > __ldcg(&A.dkey[S.udyn1])) {
                uint32_t t = S.udyn0; S.udyn0 = S.udyn1; S.udyn1 = t;
                int c5 = UCNT(5), c6 = UCNT(6); double p5 = UPROD(5), p6 = UPROD(6);
                uint32_t b5 = (S.exist >> 5) & 1u, b6 = (S.exist >> 6) & 1u;
                UCNT(5) = c6; UCNT(6) = c5; UPROD(5) = p6; UPROD(6) = p5;
                S.exist = (S.exist & 0x1fu) | (b6 << 5) | (b5 << 6);
            }
            const uint32_t exist = S.exist;
            int k = __popc(exist);
            uint32_t pad = 0;                             // :49-54  pad with A, T, G, C until 4
            if (k < 4 && !((exist >> SMC_A_A) & 1u)) { pad |= 1u << SMC_A_A; ++k; }
            if (k < 4 && !((exist >> SMC_A_T) & 1u)) { pad |= 1u << SMC_A_T; ++k; }
            if (k < 4 && !((exist >> SMC_A_G) & 1u)) { pad |= 1u << SMC_A_G; ++k; }
            if (k < 4 && !((exist >> SMC_A_C) & 1u)) { pad |= 1u << SMC_A_C; ++k; }
            const uint32_t uniq = exist | pad;
            const double rightP = S.rightP;
            const double INF = __longlong_as_double(0x7ff0000000000000ll);
            // ---- PCR prior of every allele of the barcode (:79-81): table lookups (host glibc pow) or pow()
            const bool tab = (k <= 6 && n <= A.pcr_nmax);
            const double* trow = A.pcrtab + ((size_t)(k - 4) * ((size_t)(A.pcr_nmax + 1) * (A.pcr_nmax + 2) / 2) + (size_t)n * (n + 1) / 2);
            const double denom = (double)n + 0.5 * (double)k;
            const double pcr_pad = tab ? __ldg(trow) : pcr_slow(0, denom);
            double m1 = pad ? pcr_pad : INF, m2 = INF; int arg1 = -1;      // smallest / second smallest prior and its slot
            double tpad = rightP;                                         // :88-91
            for (uint32_t m = exist; m; m &= m - 1) {
                const int s = __ffs(m) - 1;
                const int c = UCNT(s);
                const double v = tab ? __ldg(trow + c) : pcr_slow(c, denom);
                UT(s) = v;
                tpad = __dmul_rn(tpad, v);
                if (v < m1) { m2 = m1; m1 = v; arg1 = s; } else if (v < m2) m2 = v;
            }
            // ---- likelihood of each present allele (:86); pads all share tpad
            const double PCR_NO_ERROR = 1.0 - 3e-5;                       // smCounter.py:20
            for (uint32_t m = exist; m; m &= m - 1) {
                const int s = __ffs(m) - 1;
                const double minp = (s == arg1) ? m2 : m1;                // min over the OTHER members of uniq
                UT(s) = __dadd_rn(__dmul_rn(PCR_NO_ERROR, UPROD(s)), __dmul_rn(rightP, minp));
            }
            double sumP = 0.0;                                            // :93, in canonical slot order
            for (uint32_t m = uniq; m; m &= m - 1) {
                const int s = __ffs(m) - 1;
                sumP = __dadd_rn(sumP, ((exist >> s) & 1u) ? UT(s) : tpad);
            }
            // ---- posterior -> -log10(1-p) (:96, :509-510), PI accumulation (:512), consensus (:514-523)
            double best = -1.0; int nbest = 0, cons = -1;
            if (pad) {
                const double p = sumP <= 0.0 ? 0.0 : tpad / sumP;
                const double x = 1.0 - p;
                const double l = x > 0.0 ? -log10(x) : 16.0;
                unsigned long long a0, a1, a2;
                pi_limbs(l, a0, a1, a2);
                for (uint32_t m = pad; m; m &= m - 1) pi_add_limbs(A, lane, limb, S, __ffs(m) - 1, a0, a1, a2);
                best = l; nbest = __popc(pad); cons = __ffs(pad) - 1;
            }
            for (uint32_t m = exist; m; m &= m - 1) {
                const int s = __ffs(m) - 1;
                const double p = sumP <= 0.0 ? 0.0 : UT(s) / sumP;
                const double x = 1.0 - p;
                const double l = x > 0.0 ? -log10(x) : 16.0;
                unsigned long long a0, a1, a2;
                pi_limbs(l, a0, a1, a2);
                pi_add_limbs(A, lane, limb, S, s, a0, a1, a2);
                if (l > best) { best = l; nbest = 1; cons = s; }
                else if (l == best) nbest++;
            }
            if (nbest == 1) {                                             // :515-519
                bump(A, fc, lane, slot_to_aid(S, cons), KW_MT, best > A.smt ? 0x10001u : 1u, SMC_C_MT, SMC_C_STRONG);
            } else if (n == 1) {                                          // :521-523
                bump(A, fc, lane, S.last_aid, KW_MT, 1u, SMC_C_MT, SMC_C_STRONG);
            }
        }
    }
    S.n = 0; S.exist = 0; S.Q = 1.0; S.rightP = 1.0; S.ndyn = 0; S.umi_seen = false; S.umi_bc = false; S.first_read = 0xffffffffu;
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
__device__ __forceinline__ void flush_counters(const K3Args& A, int* fc, int lane, int64_t L, bool lane_valid) {
    if (!lane_valid) return;
    const size_t nl = (size_t)A.n_loci;
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
                if (lo) atomicAdd(&A.cnt[((size_t)a * SMC_NCNT + lo_idx[w]) * nl + L], lo);
                if (hi) atomicAdd(&A.cnt[((size_t)a * SMC_NCNT + hi_idx[w]) * nl + L], hi);
                if (w == KW_ALLELE_FWD && a != SMC_A_DEL && lo - hi) atomicAdd(&A.cnt[((size_t)a * SMC_NCNT + SMC_C_REV) * nl + L], lo - hi);
            }
        }
    }
}

__global__ void __launch_bounds__(K3_WARPS * 32, K3_MINBLOCKS) k_pileup(const K3Args A) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* ws = smem + (size_t)w * K3_WARP_WORDS;
    int* fc = (int*)(ws + K3_STAGE_WORDS);
    unsigned long long* limb = (unsigned long long*)(ws + K3_STAGE_WORDS + K3_FC_WORDS);
    int* ucnt = (int*)(ws + K3_STAGE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS);
    double* uprod = (double*)(ws + K3_STAGE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS + K3_UCNT_WORDS);
    double* ut = (double*)(ws + K3_STAGE_WORDS + K3_FC_WORDS + K3_LIMB_WORDS + K3_UCNT_WORDS + K3_UPROD_WORDS);

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
    const int32_t Li = (int32_t)L;

    for (int i = lane; i < K3_FC_WORDS + K3_LIMB_WORDS; i += 32) ws[K3_STAGE_WORDS + i] = 0;
    __syncwarp();

    LaneState S;
    S.cvg = S.allFrag = S.allMT = S.usedFrag = S.nBC = S.usedMT = S.mt3 = S.mt5 = S.mt7 = S.mt10 = 0;
    S.keymask = 0; S.status = 0;
    S.n = 0; S.exist = 0; S.Q = 1.0; S.rightP = 1.0; S.last_aid = 0; S.udyn0 = S.udyn1 = 0; S.ndyn = 0;
    S.umi_seen = S.umi_bc = false; S.first_read = 0xffffffffu;
    S.frag_seen = S.f_exists = S.f_paired = false; S.f_aid = 0; S.f_bq = 0;

    uint32_t prev_urank = 0xffffffffu, prev_frank = 0xffffffffu;
    bool first = true;
    const int minBQ = A.minBQ;
    uint32_t since_flush = 0;

    for (uint32_t base = eb; base < ee; base += 32) {
        {   // stage the next 32 read records in shared memory (4 x 128-bit loads per lane)
            uint32_t e = base + lane;
            if (e < ee) {
                const uint4* src = reinterpret_cast<const uint4*>(&A.recs[A.ev_read[e]]);
                uint4* dst = reinterpret_cast<uint4*>(ws + lane * 16);
                uint4 r0 = __ldg(src), r1 = __ldg(src + 1), r2 = __ldg(src + 2), r3 = __ldg(src + 3);
                dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
            }
        }
        since_flush += 32;
        if (since_flush > K3_FLUSH_EVERY) { flush_counters(A, fc, lane, L, lane_valid); since_flush = 0; }
        __syncwarp();
        // the last batch runs one extra (sentinel) iteration that only closes the open fragment and barcode
        const int cntj = (int)min(32u, ee - base) + (base + 32 >= ee ? 1 : 0);
        for (int j = 0; j < cntj; ++j) {
            const bool sentinel = base + (uint32_t)j >= ee;
            const uint32_t* rw = ws + (j & 31) * 16;
            const uint4 q0 = *reinterpret_cast<const uint4*>(rw);        // start lo hi meta
            const uint4 q1 = *reinterpret_cast<const uint4*>(rw + 4);    // sp_aln seq_off qual_off cigar_off
            const uint2 q2 = *reinterpret_cast<const uint2*>(rw + 8);    // urank frank
            const int32_t start = (int32_t)q0.x, lo = (int32_t)q0.y, hi = (int32_t)q0.z;
            const uint32_t meta = q0.w;
            const uint32_t urank = sentinel ? 0xfffffffeu : q2.x, frank = sentinel ? 0xfffffffeu : q2.y;
            // ---- barcode / fragment boundaries (warp uniform)
            if (!first) {
                if (frank != prev_frank) fragment_finalize(A, lane, ucnt, uprod, S);
                if (urank != prev_urank) umi_finalize(A, lane, L, prev_urank, fc, limb, ucnt, uprod, ut, S);
            }
            first = false; prev_urank = urank; prev_frank = frank;
            if (sentinel) break;
            // ---- does the read cover my locus?
            const bool covered = lane_valid && Li >= lo && Li < hi;
            int qpos = 0, indel = 0; bool isdel = false;
            const uint32_t ncig = meta >> 8;
            const int leftSP = (int)(q1.x & 0xffffu), alnlen = (int)(q1.x >> 16);
            if (meta & RM_SIMPLE) {
                qpos = leftSP + (p - start);
            } else {
                // htslib resolve_cigar2: find the reference-consuming op that covers p
                int x = start, y = 0; bool found = false;
                for (uint32_t k = 0; k < ncig; ++k) {
                    uint32_t cw = k < 4 ? rw[10 + k] : __ldg(&A.cigar[q1.w + k]);
                    uint32_t op = cw & 15u; int len = (int)(cw >> 4);
                    bool refop = (op == 0 || op == 7 || op == 8 || op == 2 || op == 3);
                    if (refop) {
                        if (covered && !found && p < x + len) {
                            found = true;
                            isdel = (op == 2 || op == 3);
                            qpos = isdel ? y : y + (p - x);
                            if (p == x + len - 1 && k + 1 < ncig) {       // peek the next op
                                uint32_t c2 = (k + 1) < 4 ? rw[10 + k + 1] : __ldg(&A.cigar[q1.w + k + 1]);
                                uint32_t op2 = c2 & 15u; int l2 = (int)(c2 >> 4);
                                if (op2 == 2) indel = -l2;
                                else if (op2 == 1) indel = l2;
                                else if (op2 == 6 && k + 2 < ncig) {
                                    int l3 = 0;
                                    for (uint32_t kk = k + 2; kk < ncig; ++kk) {
                                        uint32_t c3 = kk < 4 ? rw[10 + kk] : __ldg(&A.cigar[q1.w + kk]);
                                        uint32_t op3 = c3 & 15u;
                                        if (op3 == 1) l3 += (int)(c3 >> 4);
                                        else if (op3 == 2 || op3 == 0 || op3 == 3 || op3 == 7 || op3 == 8) break;
                                    }
                                    if (l3 > 0) indel = l3;
                                }
                            }
                        }
                        x += len;
                        if (op == 0 || op == 7 || op == 8) y += len;
                    } else if (op == 1 || op == 4) y += len;
                    if (__all_sync(FULL_MASK, found || !covered)) break;
                }
            }
            if (covered) {
                S.cvg++;                                                   // smCounter.py:368
                const bool reverse = meta & RM_REVERSE, read2 = meta & RM_READ2;
                uint32_t aid; int bq; bool regular = false, isN = false;
                if (indel == 0 && isdel) {                                 // :416-421
                    aid = SMC_A_DEL; bq = minBQ;
                    FCW(KW_ALLELE_FWD, SMC_A_DEL) += 1;
                } else {
                    const uint32_t sb = __ldg(&A.seq[(size_t)q1.y + (qpos >> 1)]);
                    const uint32_t nib = (qpos & 1) ? (sb & 15u) : (sb >> 4);
                    bq = (int)__ldg(&A.qual[(size_t)q1.z + qpos]);
                    const int fa = nib_to_fixed(nib);
                    if (indel == 0 && fa >= 0) {                           // :423-457 regular base, A/C/G/T
                        aid = (uint32_t)fa; regular = true;
                        FCW(KW_ALLELE_FWD, fa) += reverse ? 1 : 0x10001;
                    } else {
                        // dynamic allele: N / IUPAC base, insertion start (:371-389) or deletion start (:392-411)
                        unsigned long long key; int len = 0;
                        if (indel > 0) {
                            len = indel;
                            unsigned long long payload;
                            if (len <= 8) {
                                unsigned long long nibs = 0;
                                for (int t = 0; t < len; ++t) {
                                    int qq = qpos + 1 + t;
                                    uint32_t b2 = __ldg(&A.seq[(size_t)q1.y + (qq >> 1)]);
                                    nibs |= (unsigned long long)((qq & 1) ? (b2 & 15u) : (b2 >> 4)) << (28 - 4 * t);
                                }
                                payload = ((unsigned long long)len << 32) | nibs;
                            } else {
                                uint32_t hsh = 2166136261u ^ (uint32_t)len;
                                for (int t = 0; t < len; ++t) {
                                    int qq = qpos + 1 + t;
                                    uint32_t b2 = __ldg(&A.seq[(size_t)q1.y + (qq >> 1)]);
                                    hsh = (hsh ^ ((qq & 1) ? (b2 & 15u) : (b2 >> 4))) * 16777619u;
                                }
                                payload = (15ull << 32) | hsh;
                            }
                            key = dyn_make_key((uint32_t)Li, SMC_K_INS, nib, payload);
                        } else if (indel < 0) {
                            len = -indel;
                            key = dyn_make_key((uint32_t)Li, SMC_K_DEL, nib, (unsigned long long)len);
                        } else {
                            key = dyn_make_key((uint32_t)Li, SMC_K_BASE, nib, 0ull);
                            regular = true; isN = (nib == 15u);
                        }
                        const uint32_t e = dyn_lookup(A.dkey, A.dmask, A.drep_read, A.drep_qpos, A.dlen, A.dcount, A.gflags, key, rw[14], qpos, len);
                        aid = NF + e;
                        atomicAdd(&A.dcnt[(size_t)e * SMC_NCNT + SMC_C_ALLELE], 1);
                        if (!reverse) atomicAdd(&A.dcnt[(size_t)e * SMC_NCNT + SMC_C_FWD], 1);
                    }
                }
                const bool inc = (bq >= minBQ) && (meta & RM_OK);          // :378,400,421,431
                if (regular) {
                    if (bq < minBQ) bump(A, fc, lane, aid, KW_LOWQ_R2P, 1u, SMC_C_LOWQ, SMC_C_R2PLE);          // :428-429
                    if (inc) {                                                                                  // :432-452
                        const int d = qpos - leftSP;
                        if (!read2) {
                            const int dist = reverse ? alnlen - d : d;
                            bump(A, fc, lane, aid, KW_R1, dist <= 20 ? 0x10001u : 1u, SMC_C_R1TOT, SMC_C_R1LE);
                        } else {
                            const int dbc = reverse ? d : alnlen - d;
                            const int dpr = reverse ? alnlen - d : d;
                            bump(A, fc, lane, aid, KW_R2, dbc <= 20 ? 0x10001u : 1u, SMC_C_R2TOT, SMC_C_R2LE);
                            if (dpr <= A.primerDist) bump(A, fc, lane, aid, KW_LOWQ_R2P, 0x10000u, SMC_C_LOWQ, SMC_C_R2PLE);
                        }
                    }
                }
                S.umi_seen = true; S.frag_seen = true;                     // :463-464
                if (inc) {                                                 // :467-479
                    S.umi_bc = true; S.first_read = min(S.first_read, rw[14]);
                    if (!S.f_exists) { S.f_exists = true; S.f_aid = aid; S.f_bq = bq; S.f_paired = false; }
                    else if (aid == S.f_aid || isN) {
                        S.f_bq = min(S.f_bq, bq); S.f_paired = true;
                        if (aid == S.f_aid) bump(A, fc, lane, aid, KW_PAIR, 1u, SMC_C_CONCORD, SMC_C_DISCORD);
                    } else { S.f_exists = false; bump(A, fc, lane, aid, KW_PAIR, 0x10000u, SMC_C_CONCORD, SMC_C_DISCORD); }
                }
            }
        }
        __syncwarp();
    }
    // ---- flush the lane's locus to the per-locus accumulators ([field][locus] layout: coalesced across lanes)
    flush_counters(A, fc, lane, L, lane_valid);
    if (lane_valid) {
        const size_t nl = (size_t)A.n_loci;
#pragma unroll
        for (int a = 0; a < NF; ++a) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                unsigned long long v = LIMB(a, j);
                if (v) atomicAdd(&A.limb[((size_t)a * 3 + j) * nl + L], v);
            }
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
