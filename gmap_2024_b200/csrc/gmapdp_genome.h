/* gmapdp_genome.h -- the reference's compressed genome read in place, and its MaxEnt splice-site model,
 * for the host (shim) and the device (kernels) alike: SURVEY.md section 8 row A13.
 *
 * Genome.  Genomecomp_T blocks as gmap_build writes them and genome.c reads them (uncompress_mmap, genome.c:8985-9060):
 * three 32-bit words per 32 nt -- high, low, flags.  Nucleotide j of a block is the 2-bit code (A C G T = 0 1 2 3)
 * at bits 2j of `low' (j < 16) or bits 2(j-16) of `high'; a set bit j of `flags' makes it 'N'.
 * A segment as Genome_get_segment_right / _left deliver it (genome.c:11023, :11079) is
 *     seg[i] = chr( G[p0 + i] )                     revcomp == false
 *     seg[i] = complement( chr( G[p0 - i] ) )       revcomp == true   (the reference reverses and complements in place)
 * with '*' for positions outside [chroffset, chrhigh): one start coordinate and a direction per side.
 *
 * MaxEnt (maxent_hr.c).  The four probabilities of a genomic position are products of table entries indexed by
 * 2-bit windows of the genome (flags ignored, as in the reference), turned into odds / (1 + odds):
 *   donor        (Maxent_hr_donor_prob :27357, generic form donor_prob_plus :24757)       window at splice_pos - 3, 9 nt
 *   acceptor     (Maxent_hr_acceptor_prob :27433, acceptor_prob_plus :25197)              window at splice_pos - 20, 23 nt
 *   antidonor    (Maxent_hr_antidonor_prob :27512, donor_prob_minus :26046)               window at splice_pos - 6, 9 nt
 *   antiacceptor (Maxent_hr_antiacceptor_prob :27586, acceptor_prob_minus :26484)         window at splice_pos - 3, 23 nt
 * and 0.0 when the window would start left of chroffset.  The tables are the model's parameters: data of the
 * reference (static arrays of maxent_hr.c), handed over by the caller at set-up (gmapdp_maxent_upload); this file holds
 * only the index arithmetic.  IEEE double multiplications in the reference's order, one division: bit-identical on
 * both sides (no contraction is possible: there is no multiply-add in the expression).
 */
#ifndef GMAPDP_GENOME_H
#define GMAPDP_GENOME_H

#include <stdint.h>
#include "../../include/gmapdp_b200.h"

#if defined(__CUDACC__)
#define GDPG_HD __host__ __device__ __forceinline__
#else
#define GDPG_HD inline
#endif

/* the 14 tables in one array of doubles (gmapdp_maxent_tables gives the pointers; this is the packed copy) */
enum {
  GDP_ME_DONOR_P = 0, GDP_ME_ACC1_P = 16384, GDP_ME_ACC2_P = 2 * 16384, GDP_ME_ACC3_P = 3 * 16384, GDP_ME_ACC467_P = 4 * 16384,
  GDP_ME_ACC589_P = 5 * 16384,
  GDP_ME_DONOR_M = 6 * 16384, GDP_ME_ACC1_M = 7 * 16384, GDP_ME_ACC2_M = 8 * 16384, GDP_ME_ACC3_M = 9 * 16384, GDP_ME_ACC467_M = 10 * 16384,
  GDP_ME_ACC589_M = 11 * 16384,
  GDP_ME_DONOR_DI_P = 12 * 16384, GDP_ME_ACC_DI_P = 12 * 16384 + 16, GDP_ME_DONOR_DI_M = 12 * 16384 + 32, GDP_ME_ACC_DI_M = 12 * 16384 + 48,
  GDP_ME_NDOUBLES = 12 * 16384 + 64
};

/* probability kinds, in the order of tests/harness.py's refdrv_maxent */
enum { GDP_ME_DONOR = 0, GDP_ME_ACCEPTOR = 1, GDP_ME_ANTIDONOR = 2, GDP_ME_ANTIACCEPTOR = 3 };

struct GdpGenome {
  const uint32_t *blocks;	/* high, low, flags per 32 nt */
  uint64_t nwords;		/* words readable at `blocks' */
};

GDPG_HD uint32_t gdp_genome_word (const GdpGenome &g, uint64_t w) { return w < g.nwords ? g.blocks[w] : 0u; }

/* chr(G[pos]) */
GDPG_HD int gdp_genome_char (const GdpGenome &g, uint32_t pos) {
  const uint64_t ptr = (uint64_t) (pos >> 5) * 3u;
  const uint32_t j = pos & 31u;
  const uint32_t flags = gdp_genome_word(g,ptr + 2);
  if ((flags >> j) & 1u) return 'N';
  const uint32_t w = (j < 16u) ? gdp_genome_word(g,ptr + 1) >> (2u * j) : gdp_genome_word(g,ptr) >> (2u * (j - 16u));
  return (int) ((0x54474341u >> (8u * (w & 3u))) & 0xffu);	/* "ACGT" */
}

/* seg[i] of a segment that starts at p0 and runs in direction dir (+1, or -1 = complemented), see above */
GDPG_HD int gdp_segment_char (const GdpGenome &g, uint32_t p0, int dir, int i, uint32_t chroffset, uint32_t chrhigh) {
  const int64_t pos = (int64_t) p0 + (int64_t) dir * i;
  if (pos < (int64_t) chroffset || pos >= (int64_t) chrhigh) return '*';
  const int ch = gdp_genome_char(g,(uint32_t) pos);
  if (dir > 0 || ch == 'N') return ch;
  return ch == 'A' ? 'T' : (ch == 'C' ? 'G' : (ch == 'G' ? 'C' : 'A'));
}

/* the 2-bit codes of the 32 nt that start at startpos, nt k at bits 2k (flags ignored) */
GDPG_HD uint64_t gdp_genome_window (const GdpGenome &g, uint32_t startpos) {
  const uint64_t ptr = (uint64_t) (startpos >> 5) * 3u;
  const uint32_t shift = startpos & 31u;
  const uint64_t cur = (uint64_t) gdp_genome_word(g,ptr + 1) | ((uint64_t) gdp_genome_word(g,ptr) << 32);
  if (shift == 0u) return cur;
  const uint64_t nxt = (uint64_t) gdp_genome_word(g,ptr + 4) | ((uint64_t) gdp_genome_word(g,ptr + 3) << 32);
  return (cur >> (2u * shift)) | (nxt << (64u - 2u * shift));
}

/* The evaluation as a sequence of steps, one large table per step (gdp_maxent_prob below runs them back to back; the
   look-ups of the steps are independent loads, only the multiplications are ordered): window start = splice_pos - margin(kind); pass p multiplies the running odds by its table entry (and,
   where the reference does, by the dinucleotide entry right after it); the last pass returns odds / (1 + odds). */
GDPG_HD int gdp_maxent_margin (int kind) { return kind == GDP_ME_DONOR ? 3 : (kind == GDP_ME_ACCEPTOR ? 20 : (kind == GDP_ME_ANTIDONOR ? 6 : 3)); }
GDPG_HD int gdp_maxent_npasses (int kind) { return (kind == GDP_ME_DONOR || kind == GDP_ME_ANTIDONOR) ? 1 : 5; }
/* offset (in doubles) of the large table pass `pass' of `kind' reads */
GDPG_HD int gdp_maxent_table (int kind, int pass) {
  if (kind == GDP_ME_DONOR) return GDP_ME_DONOR_P;
  if (kind == GDP_ME_ANTIDONOR) return GDP_ME_DONOR_M;
  if (kind == GDP_ME_ACCEPTOR) return pass == 0 ? GDP_ME_ACC1_P : (pass == 1 ? GDP_ME_ACC2_P : (pass == 2 ? GDP_ME_ACC3_P : (pass == 3 ? GDP_ME_ACC467_P : GDP_ME_ACC589_P)));
  return pass == 0 ? GDP_ME_ACC1_M : (pass == 1 ? GDP_ME_ACC2_M : (pass == 2 ? GDP_ME_ACC3_M : (pass == 3 ? GDP_ME_ACC467_M : GDP_ME_ACC589_M)));
}
/* big = the pass's large table (wherever it lives), di = the four 16-entry dinucleotide tables (the 64 doubles at
   GDP_ME_DONOR_DI_P of the packed array) */
#define GDP_DI_DONOR_P 0
#define GDP_DI_ACC_P 16
#define GDP_DI_DONOR_M 32
#define GDP_DI_ACC_M 48
GDPG_HD double gdp_maxent_step (int kind, int pass, uint64_t W, const double *big, const double *di, double odds) {
  const uint32_t seq = (uint32_t) W;
  if (kind == GDP_ME_DONOR) return big[(seq & 0x3Fu) | ((seq >> 4) & 0x3FC0u)] * di[GDP_DI_DONOR_P + ((seq >> 6) & 0x0Fu)];
  if (kind == GDP_ME_ANTIDONOR) return big[(seq & 0xFFu) | ((seq >> 4) & 0x3F00u)] * di[GDP_DI_DONOR_M + ((seq >> 8) & 0x0Fu)];
  if (kind == GDP_ME_ACCEPTOR) {
    if (pass == 0) return big[seq & 0x3FFFu];						/* 7-mer at +0 */
    if (pass == 1) return odds * big[(uint32_t) (W >> 14) & 0x3FFFu];			/* 7-mer at +7 */
    if (pass == 2) {									/* 9-mer at +14: 4 nt, skip 2, 3 nt */
      const uint32_t s9 = (uint32_t) (W >> 28);
      odds *= big[(s9 & 0xFFu) | ((s9 >> 4) & 0x3F00u)];
      return odds * di[GDP_DI_ACC_P + ((s9 >> 8) & 0x0Fu)];
    }
    if (pass == 3) return odds * big[(uint32_t) (W >> 8) & 0x3FFFu];			/* 7-mer at +4 */
    return odds * big[(uint32_t) (W >> 22) & 0x3FFFu];					/* 7-mer at +11 */
  }
  if (pass == 0) return big[(uint32_t) (W >> 32) & 0x3FFFu];				/* 7-mer at +16 */
  if (pass == 1) return odds * big[(uint32_t) (W >> 18) & 0x3FFFu];			/* 7-mer at +9 */
  if (pass == 2) {									/* 9-mer at +0: 3 nt, skip 2, 4 nt */
    odds *= big[(seq & 0x3Fu) | ((seq >> 4) & 0x3FC0u)];
    return odds * di[GDP_DI_ACC_M + ((seq >> 6) & 0x0Fu)];
  }
  if (pass == 3) return odds * big[(uint32_t) (W >> 24) & 0x3FFFu];			/* 7-mer at +12 */
  return odds * big[(uint32_t) (W >> 10) & 0x3FFFu];					/* 7-mer at +5 */
}

GDPG_HD double gdp_maxent_prob (int kind, const GdpGenome &g, const double *T, uint32_t splice_pos, uint32_t chroffset) {
  const uint32_t margin = (uint32_t) gdp_maxent_margin(kind);
  if (splice_pos < chroffset + margin) return 0.0;
  const uint64_t W = gdp_genome_window(g,splice_pos - margin);
  const int np = gdp_maxent_npasses(kind);
  double odds = 0.0;
  for (int p = 0; p < np; p++) odds = gdp_maxent_step(kind,p,W,T + gdp_maxent_table(kind,p),T + GDP_ME_DONOR_DI_P,odds);
  return odds / (1 + odds);
}

#endif
