/* Bench / test infrastructure (not product code): a fast deterministic generator of the BASELINE.json
 * config-2 workload -- "1M synthetic DP boxes, 50-2000 bp per side, single/cdna/genome(splice)/end5/end3
 * modes" -- inside the reference's shape envelope:
 *   mode = i % 5 : single, genome, cdna, end5, end3
 *   single : glength ~ U[50,2000]; half of them rlength ~ glength (mutated copy), half rlength ~ U[50,2000]
 *            independent (wide band); extraband_single = max(6,|r-g|), widebandp = true (stage3.c:9070-9081)
 *   genome : rlength ~ U[50,1990], glengthL = glengthR = rlength + 8 (stage3.c:9530), planted GT..AG /
 *            GC..AG / CT..AC / none = 65/12/12/11 %, extraband_paired 14, finalp 10 %, halfp 0
 *   cdna   : glength ~ U[50,200] (the bridge is O(g^2 band^2)), rlengthL = rlengthR = glength + 8 (stage3.c:9276)
 *   end5/3 : rlength ~ U[50,1990], glength = rlength + 10, endalign GAP / INDELS / NOGAPS in thirds
 *   per-base error 0.1 / 1 / 5 / 15 % (60 % substitution, 20 % deletion, 20 % insertion), defect_rate equal to it
 * Box i depends only on (seed, i).
 *
 * Decorated stratum (`small' bit 1; SURVEY.md section 8d's input decorations): 0.3 % N in the genomic segments, 0.2 %
 * lower-case and 0.1 % IUPAC characters in the query, and 5 % of the single-gap boxes tandem repeats (unit 16-24 nt, the
 * query rotated by 7-9 units of the band so that the matching diagonal sits below it: what exposes the 8-bit stripe
 * artefact F11).
 *
 * Resident-genome form (benchgen_resident_begin; bench.py's default): the genomic segments of the boxes are laid out,
 * each between 32-nt pads, in one synthetic genome in the reference's compressed block layout; boxes are queued as
 * COORDINATES into it (GmapDP_batch_next_coords) and the splice-site probabilities of genome gaps are MaxEnt values
 * (gmapdp_genome.h arithmetic over SYNTHETIC tables: log-normal factors, consensus dinucleotides favoured -- the
 * reference's own tables are not available to the product) instead of the u^3 arrays.  benchgen_make then fills
 * left_probs / right_probs with exactly those values (the pads make them independent of the neighbours), so the CPU
 * checkers see the same box.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <math.h>
#include <vector>
#include "../../include/gmapdp_shim.h"
#include "../../gmap_2024_b200/csrc/gmapdp_genome.h"

#define BG_MAXSEQ 2304

extern "C" {

typedef struct benchgen_box {
  int mode, dynprogindex, jump_late_p, extraband, widebandp, endalign, require_pos_score_p;
  int cdna_direction, finalp, halfp, maxpeelback;
  int rlength, rlengthR, glength, glengthR;	/* genome: glength = glengthL; cdna: rlength = rlengthL */
  int roffset, rev_roffsetR, goffset, rev_goffsetR, querylength;
  double defect_rate;
  char query[2 * BG_MAXSEQ], queryuc[2 * BG_MAXSEQ];
  char gsegL[BG_MAXSEQ], gsegR[BG_MAXSEQ];
  double left_probs[BG_MAXSEQ], right_probs[BG_MAXSEQ];
} benchgen_box;

}

namespace {
struct Rng {
  uint64_t s;
  explicit Rng (uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) { next(); next(); }
  uint64_t next () { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  int below (int n) { return (int) (next() % (uint64_t) n); }
  int range (int lo, int hi) { return lo + below(hi - lo + 1); }	/* inclusive */
  double unit () { return (double) (next() >> 11) * (1.0 / 9007199254740992.0); }
};

const char NT[4] = {'A','C','G','T'};

void rand_dna (Rng &r, char *out, int n) {
  int i = 0;
  while (i < n) { uint64_t w = r.next(); for (int k = 0; k < 32 && i < n; k++, w >>= 2) out[i++] = NT[w & 3]; }
}

/* mutated copy of src[0..n) into dst (capacity cap); returns its length */
int mutate (Rng &r, const char *src, int n, double e, char *dst, int cap) {
  int m = 0;
  const uint64_t thr = (uint64_t) (e * 4294967296.0);
  for (int i = 0; i < n && m < cap - 2; i++) {
    uint64_t w = r.next();
    if ((w & 0xffffffffu) < thr) {
      int k = (int) ((w >> 32) % 10);
      if (k < 6) dst[m++] = NT[(w >> 40) & 3];
      else if (k < 8) { }
      else { dst[m++] = src[i]; dst[m++] = NT[(w >> 40) & 3]; }
    } else dst[m++] = src[i];
  }
  return m;
}

void probs (Rng &r, double *p, int n, int hot) {
  for (int i = 0; i < n; i++) { double u = r.unit(); p[i] = u * u * u; }
  for (int k = hot - 1; k <= hot + 1; k++) if (k >= 0 && k < n) p[k] = 0.9 + 0.0999 * r.unit();
}
}

/* ---- resident genome ------------------------------------------------------------------------------ */
namespace {
const int PAD = 32;
/* consensus contexts of the planted sites: donor = 3 exon nt + GT + 4 intron nt; acceptor = 18 intron nt + AG + 3 exon nt;
   the minus-strand twins are their reverse complements */
const char BG_DONOR9[] = "CAGGTAAGT";
const char BG_ANTIDONOR9[] = "ACTTACCTG";
const char BG_ACC23[] = "TTTCTTTTCTTTTTTTCCAGGAA";
const char BG_ANTIACC23[] = "TTCCTGGAAAAAAAGAAAAGAAA";
struct Resident {
  bool on = false;
  std::vector<double> packed;		/* GDP_ME_NDOUBLES, the layout of gmapdp_genome.h */
  gmapdp_maxent_tables tables;
  uint32_t *blocks = NULL; size_t cap_words = 0; uint64_t used_nt = 0;
} RES;

void synth_tables (uint64_t seed) {
  Rng r(seed ^ 0x4D41584Eull);
  RES.packed.assign(GDP_ME_NDOUBLES,0.0);
  auto normal = [&]() { double u1 = r.unit(), u2 = r.unit(); if (u1 < 1e-12) u1 = 1e-12; return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2); };
  for (int t = 0; t < 12; t++) {
    const bool donor = (t == 0 || t == 6);
    const double mu = donor ? -3.0 : -0.6, sd = donor ? 2.1 : 0.86;		/* donors: one factor; acceptors: five */
    for (int k = 0; k < 16384; k++) RES.packed[(size_t) t * 16384 + k] = exp(mu + sd * normal());
  }
  for (int k = 0; k < 16; k++) {
    RES.packed[GDP_ME_DONOR_DI_P + k] = (k == 14) ? 1.0 : 0.004;		/* GT */
    RES.packed[GDP_ME_ACC_DI_P + k] = (k == 8) ? 1.0 : 0.004;		/* AG */
    RES.packed[GDP_ME_DONOR_DI_M + k] = (k == 4) ? 1.0 : 0.004;		/* AC */
    RES.packed[GDP_ME_ACC_DI_M + k] = (k == 13) ? 1.0 : 0.004;		/* CT */
  }
  /* the planted splice sites carry consensus contexts (benchgen_make), and those contexts score high -- like true sites
     under the real model, so that the entry point's shortcut for clean introns (genome_gap_simple, dynprog_genome.c:3005:
     both probabilities >= 0.90) fires about as often as on real data */
  {
    auto code = [](char c) { return c == 'A' ? 0ull : (c == 'C' ? 1ull : (c == 'G' ? 2ull : 3ull)); };
    auto window = [&](const char *str, int n) { uint64_t W = 0; for (int k = 0; k < n; k++) W |= code(str[k]) << (2 * k); return W; };
    double *P = RES.packed.data();
    const double hi1 = 60.0, hi5 = 2.3;				/* odds 60 in one factor / over five factors */
    for (int gc = 0; gc < 2; gc++) {				/* GT and GC donors share the table entry (the dinucleotide is skipped) */
      const uint32_t seq = (uint32_t) window(BG_DONOR9,9);
      P[GDP_ME_DONOR_P + ((seq & 0x3Fu) | ((seq >> 4) & 0x3FC0u))] = hi1;
    }
    { const uint32_t seq = (uint32_t) window(BG_ANTIDONOR9,9); P[GDP_ME_DONOR_M + ((seq & 0xFFu) | ((seq >> 4) & 0x3F00u))] = hi1; }
    {
      const uint64_t W = window(BG_ACC23,23); const uint32_t s9 = (uint32_t) (W >> 28);
      P[GDP_ME_ACC1_P + (W & 0x3FFFu)] = hi5; P[GDP_ME_ACC2_P + ((W >> 14) & 0x3FFFu)] = hi5;
      P[GDP_ME_ACC3_P + ((s9 & 0xFFu) | ((s9 >> 4) & 0x3F00u))] = hi5;
      P[GDP_ME_ACC467_P + ((W >> 8) & 0x3FFFu)] = hi5; P[GDP_ME_ACC589_P + ((W >> 22) & 0x3FFFu)] = hi5;
    }
    {
      const uint64_t W = window(BG_ANTIACC23,23); const uint32_t seq = (uint32_t) W;
      P[GDP_ME_ACC1_M + ((W >> 32) & 0x3FFFu)] = hi5; P[GDP_ME_ACC2_M + ((W >> 18) & 0x3FFFu)] = hi5;
      P[GDP_ME_ACC3_M + ((seq & 0x3Fu) | ((seq >> 4) & 0x3FC0u))] = hi5;
      P[GDP_ME_ACC467_M + ((W >> 24) & 0x3FFFu)] = hi5; P[GDP_ME_ACC589_M + ((W >> 10) & 0x3FFFu)] = hi5;
    }
  }
  const double *p = RES.packed.data();
  gmapdp_maxent_tables &T = RES.tables;
  T.donor_plus = p + GDP_ME_DONOR_P; T.acc1_plus = p + GDP_ME_ACC1_P; T.acc2_plus = p + GDP_ME_ACC2_P; T.acc3_plus = p + GDP_ME_ACC3_P;
  T.acc467_plus = p + GDP_ME_ACC467_P; T.acc589_plus = p + GDP_ME_ACC589_P; T.donor_di_plus = p + GDP_ME_DONOR_DI_P; T.acc_di_plus = p + GDP_ME_ACC_DI_P;
  T.donor_minus = p + GDP_ME_DONOR_M; T.acc1_minus = p + GDP_ME_ACC1_M; T.acc2_minus = p + GDP_ME_ACC2_M; T.acc3_minus = p + GDP_ME_ACC3_M;
  T.acc467_minus = p + GDP_ME_ACC467_M; T.acc589_minus = p + GDP_ME_ACC589_M; T.donor_di_minus = p + GDP_ME_DONOR_DI_M; T.acc_di_minus = p + GDP_ME_ACC_DI_M;
}

inline void put_nt (uint32_t *blocks, uint64_t pos, char ch) {
  const uint64_t ptr = (pos >> 5) * 3; const uint32_t j = (uint32_t) (pos & 31);
  uint32_t code = 0;
  switch (ch) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
    default: blocks[ptr + 2] |= 1u << j; break; }
  if (j < 16) blocks[ptr + 1] |= code << (2 * j); else blocks[ptr] |= code << (2 * (j - 16));
}

/* appends pad + seg + pad; returns the coordinate of seg[0] (-1: out of room) */
int64_t place (uint32_t *blocks, size_t cap_words, uint64_t &used, const char *seg, int n) {
  if (((used + (uint64_t) n + 2 * PAD) / 32 + 2) * 3 > cap_words) return -1;
  used += PAD;				/* pads are 'A' = code 0 = the zero-filled state */
  const uint64_t at = used;
  for (int k = 0; k < n; k++) put_nt(blocks,at + k,seg[k]);
  used += (uint64_t) n + PAD;
  return (int64_t) at;
}

/* MaxEnt entries of a genome-gap box whose segments lie at posL / posR (the binding's mapping on the Watson strand,
   integration/dynprog_sm100.c; dynprog_genome.c:970-1061) */
void box_coords (const benchgen_box *b, uint64_t posL, uint64_t posR, uint64_t chrhigh, gmapdp_coords *co) {
  memset(co,0,sizeof(*co));
  co->chroffset = 0; co->chrhigh = (uint32_t) chrhigh;
  co->gposL = (uint32_t) posL; co->gposR = (uint32_t) posR;
  if (b->mode == GMAPDP_GENOME) {
    co->probs = 1;
    co->probkindL = b->cdna_direction > 0 ? GDP_ME_DONOR : GDP_ME_ANTIACCEPTOR; co->probposL = (uint32_t) posL; co->probnegL = 0;
    co->probkindR = b->cdna_direction > 0 ? GDP_ME_ACCEPTOR : GDP_ME_ANTIDONOR; co->probposR = (uint32_t) (posR + b->glengthR); co->probnegR = 1;
  }
}
}

extern "C" int benchgen_resident_begin (uint64_t seed, uint64_t capacity_nt) {
  free(RES.blocks);
  RES.cap_words = (size_t) (capacity_nt / 32 + 4) * 3;
  RES.blocks = (uint32_t *) calloc(RES.cap_words,sizeof(uint32_t));
  RES.used_nt = 0;
  synth_tables(seed);
  RES.on = (RES.blocks != NULL);
  return RES.on ? 0 : -1;
}
extern "C" void benchgen_resident_tables_only (uint64_t seed) { synth_tables(seed); RES.on = true; }	/* CPU workers: benchgen_make only */
extern "C" void benchgen_resident_end (void) { free(RES.blocks); RES.blocks = NULL; RES.cap_words = 0; RES.used_nt = 0; RES.on = false; }
extern "C" const uint32_t *benchgen_resident_blocks (uint64_t *used_words, uint64_t *cap_words) {
  *used_words = (RES.used_nt / 32 + 2) * 3; *cap_words = RES.cap_words;
  return RES.blocks;
}
extern "C" const gmapdp_maxent_tables *benchgen_resident_tables (void) { return &RES.tables; }

namespace {
void decorate_genome (Rng &r, char *g, int n) {
  for (int k = 0; k < n; k++) if (r.below(1000) < 3) g[k] = 'N';
}
void decorate_query (Rng &r, char *q, char *quc, int n) {
  static const char iupac[7] = {'R','Y','W','S','M','K','N'};
  for (int k = 0; k < n; k++) {
    const int x = r.below(1000);
    if (x < 1) { q[k] = quc[k] = iupac[r.below(7)]; }
    else if (x < 3) q[k] = (char) (quc[k] | 0x20);		/* lower case in the query, upper case in its uc twin */
  }
}
}

extern "C" void benchgen_make (uint64_t seed, long i, benchgen_box *b, int small_flags) {
  const int small = small_flags & 1;
  const bool decor = (small_flags & 2) != 0;
  Rng r(seed * 1000003ull + (uint64_t) i);
  static const double errs[4] = {0.001, 0.01, 0.05, 0.15};
  memset(b,0,offsetof(benchgen_box,query));
  const int hi = small ? 150 : 2000, lo = small ? 15 : 50;
  b->mode = (int) (i % 5);
  const double e = errs[r.below(4)];
  b->defect_rate = e; b->jump_late_p = r.below(2); b->dynprogindex = 1; b->maxpeelback = 60;
  const int prefix = r.below(20);
  char tmp[2 * BG_MAXSEQ];
  rand_dna(r,b->query,prefix);
  char *q = b->query + prefix;
  int qn = 0;

  if (b->mode == GMAPDP_SINGLE) {
    const int g = r.range(lo,hi);
    rand_dna(r,b->gsegL,g);
    if (decor && r.below(20) == 0) {	/* tandem repeats: the query is the same repeat, rotated */
      const int unit = r.range(16,24), rot = r.range(7,9);
      for (int k = unit; k < g; k++) b->gsegL[k] = b->gsegL[k % unit];
      for (int k = 0; k < g; k++) tmp[k] = b->gsegL[(k + rot) % unit];
      qn = mutate(r,tmp,g,e,q,BG_MAXSEQ);
    } else
    qn = mutate(r,b->gsegL,g,e,q,BG_MAXSEQ);
    if (r.below(2)) {			/* independent rlength: wide band */
      const int want = r.range(lo,hi);
      if (want < qn) qn = want; else { rand_dna(r,q + qn,want - qn); qn = want; }
    }
    if (qn > hi) qn = hi;
    if (qn < 1) { q[0] = 'A'; qn = 1; }
    b->rlength = b->rlengthR = qn; b->glength = b->glengthR = g;
    b->roffset = prefix; b->goffset = 1000;
    b->extraband = std::max(6,abs(qn - g)); b->widebandp = 1;

  } else if (b->mode == GMAPDP_END5 || b->mode == GMAPDP_END3) {
    const int rl = r.range(lo,hi - 10), g = rl + 10;
    rand_dna(r,b->gsegL,g);
    int m = mutate(r,b->gsegL,g,e,tmp,2 * BG_MAXSEQ);
    if (r.below(3) == 0) {		/* diverging far end */
      const int k = r.below(m + 1);
      if (b->mode == GMAPDP_END3) rand_dna(r,tmp + k,m - k); else rand_dna(r,tmp,k);
    }
    while (m < rl) { rand_dna(r,tmp + m,rl - m); m = rl; }
    if (b->mode == GMAPDP_END3) memcpy(q,tmp,rl); else memcpy(q,tmp + (m - rl),rl);
    qn = rl;
    b->rlength = b->rlengthR = rl; b->glength = b->glengthR = g;
    b->extraband = 6; b->endalign = (int) ((i / 5) % 3);	/* GAP, INDELS, NOGAPS */
    if (b->mode == GMAPDP_END3) { b->roffset = prefix; b->goffset = 1000; }
    else { b->roffset = prefix + rl - 1; b->goffset = 1000 + g - 1; }

  } else if (b->mode == GMAPDP_GENOME) {
    const int rl = r.range(lo,hi - 10), a = r.range(1,rl - 1), bb = rl - a, gl = rl + 8;
    rand_dna(r,b->gsegL,gl); rand_dna(r,b->gsegR,gl);
    const int k = r.below(100);
    const char *d = k < 65 ? "GTAG" : (k < 77 ? "GCAG" : (k < 89 ? "CTAC" : NULL));
    if (d) { b->gsegL[a] = d[0]; b->gsegL[a+1] = d[1]; b->gsegR[gl-bb-2] = d[2]; b->gsegR[gl-bb-1] = d[3]; }
    if (d && RES.on) {
      /* resident form: the sites' consensus contexts around the dinucleotides (which stay as planted) */
      const int sR = gl - bb - 2;			/* position of the right site's first nt */
      if (k < 77) {					/* GT-AG / GC-AG: donor on the left, acceptor on the right */
	if (a >= 3 && a + 6 <= gl) { memcpy(b->gsegL + a - 3,BG_DONOR9,3); memcpy(b->gsegL + a + 2,BG_DONOR9 + 5,4); }
	if (sR >= 18 && sR + 5 <= gl) { memcpy(b->gsegR + sR - 18,BG_ACC23,18); memcpy(b->gsegR + sR + 2,BG_ACC23 + 20,3); }
      } else {						/* CT-AC: antiacceptor on the left, antidonor on the right */
	if (a >= 3 && a + 20 <= gl) { memcpy(b->gsegL + a - 3,BG_ANTIACC23,3); memcpy(b->gsegL + a + 2,BG_ANTIACC23 + 5,18); }
	if (sR >= 4 && sR + 5 <= gl) { memcpy(b->gsegR + sR - 4,BG_ANTIDONOR9,4); memcpy(b->gsegR + sR + 2,BG_ANTIDONOR9 + 6,3); }
      }
    }
    int m1 = mutate(r,b->gsegL,a,e,tmp,2 * BG_MAXSEQ);
    int m2 = mutate(r,b->gsegR + gl - bb,bb,e,tmp + m1,2 * BG_MAXSEQ - m1);
    int m = m1 + m2;
    if (m >= rl) memcpy(q,tmp,rl); else { memcpy(q,tmp,m); rand_dna(r,q + m,rl - m); }
    qn = rl;
    b->rlength = b->rlengthR = rl; b->glength = gl; b->glengthR = gl;
    b->roffset = prefix; b->goffset = 1000; b->rev_goffsetR = 1000 + 2 * gl + 5000;
    b->cdna_direction = k < 77 ? 1 : (k < 89 ? -1 : 0);
    b->extraband = 14; b->finalp = r.below(10) == 0; b->halfp = 0;
    probs(r,b->left_probs,gl - 1,a); probs(r,b->right_probs,gl - 1,bb);

  } else {	/* cdna */
    const int g = r.range(lo,small ? 100 : 200), k = r.range(1,g - 1), ins = r.range(10,40);
    rand_dna(r,b->gsegL,g);
    int m = mutate(r,b->gsegL,k,e,q,BG_MAXSEQ);
    rand_dna(r,q + m,ins); m += ins;
    m += mutate(r,b->gsegL + k,g - k,e,q + m,BG_MAXSEQ - m);
    while (m < g + 10) { rand_dna(r,q + m,g + 10 - m); m = g + 10; }
    qn = m;
    b->rlength = b->rlengthR = g + 8; b->glength = b->glengthR = g;
    b->roffset = prefix; b->rev_roffsetR = prefix + qn - 1; b->goffset = 1000;
    b->extraband = 14;
  }
  const int suffix = r.below(20);
  rand_dna(r,q + qn,suffix);
  b->querylength = prefix + qn + suffix;
  b->query[b->querylength] = '\0';
  for (int j = 0; j <= b->querylength; j++) b->queryuc[j] = b->query[j];	/* generated upper case already */
  if (decor) {
    decorate_query(r,b->query,b->queryuc,b->querylength);
    decorate_genome(r,b->gsegL,b->glength);
    if (b->mode == GMAPDP_GENOME) decorate_genome(r,b->gsegR,b->glengthR);
  }
  b->gsegL[b->glength] = '\0';
  if (b->mode == GMAPDP_GENOME) b->gsegR[b->glengthR] = '\0';
  if (RES.on && b->mode == GMAPDP_GENOME) {
      /* the MaxEnt values this box has in the resident genome: its two segments between pads, nothing else in reach */
      uint32_t mini[((2 * BG_MAXSEQ + 4 * PAD) / 32 + 4) * 3];
      memset(mini,0,sizeof(mini));
      uint64_t used = 0;
      const int gl = b->glength;
      const int64_t pL = place(mini,sizeof(mini) / 4,used,b->gsegL,gl), pR = place(mini,sizeof(mini) / 4,used,b->gsegR,gl);
      gmapdp_coords co;
      box_coords(b,(uint64_t) pL,(uint64_t) pR,used,&co);
      GdpGenome G; G.blocks = mini; G.nwords = sizeof(mini) / 4;
      for (int c = 0; c < gl - 1; c++) {
	b->left_probs[c] = gdp_maxent_prob(co.probkindL,G,RES.packed.data(),co.probposL + (uint32_t) c,0);
	b->right_probs[c] = gdp_maxent_prob(co.probkindR,G,RES.packed.data(),co.probposR - (uint32_t) c,0);
      }
    }
}

/* queues box i into a shim batch; returns the call id */
extern "C" int benchgen_add (gmapdp_batch *batch, const benchgen_box *b) {
  switch (b->mode) {
  case GMAPDP_SINGLE:
    return GmapDP_single_gap(batch,b->dynprogindex,b->query + b->roffset,b->queryuc + b->roffset,b->rlength,b->glength,
			     b->roffset,b->goffset,b->gsegL,b->gsegL,b->jump_late_p,b->extraband,b->widebandp,b->defect_rate);
  case GMAPDP_END5:
    return GmapDP_end5_gap(batch,b->dynprogindex,b->query + b->roffset,b->queryuc + b->roffset,b->rlength,b->glength,
			   b->roffset,b->goffset,b->gsegL,b->gsegL,b->jump_late_p,b->extraband,b->defect_rate,b->endalign,
			   b->require_pos_score_p);
  case GMAPDP_END3:
    return GmapDP_end3_gap(batch,b->dynprogindex,b->query + b->roffset,b->queryuc + b->roffset,b->rlength,b->glength,
			   b->roffset,b->goffset,b->gsegL,b->gsegL,b->jump_late_p,b->extraband,b->defect_rate,b->endalign,
			   b->require_pos_score_p);
  case GMAPDP_GENOME:
    return GmapDP_genome_gap(batch,b->dynprogindex,b->query + b->roffset,b->queryuc + b->roffset,b->rlength,b->glength,
			     b->glengthR,b->roffset,b->goffset,b->rev_goffsetR,b->gsegL,b->gsegL,b->gsegR,b->gsegR,
			     b->left_probs,b->right_probs,b->cdna_direction,b->jump_late_p,b->extraband,b->defect_rate,
			     b->maxpeelback,b->halfp,b->finalp);
  default:
    return GmapDP_cdna_gap(batch,b->dynprogindex,b->query + b->roffset,b->queryuc + b->roffset,
			   b->query + b->rev_roffsetR,b->queryuc + b->rev_roffsetR,b->rlength,b->rlengthR,b->glength,
			   b->roffset,b->rev_roffsetR,b->goffset,b->gsegL,b->gsegL,b->gsegL,b->gsegL,
			   b->jump_late_p,b->extraband,b->defect_rate);
  }
}

/* boxes [i0, i0+n) with stride `stride` (rank sharding: i = i0 + k*stride) */
extern "C" long benchgen_fill_batch (gmapdp_batch *batch, uint64_t seed, long i0, long n, long stride, int small, int modemask) {
  benchgen_box *b = (benchgen_box *) malloc(sizeof(benchgen_box));
  long added = 0;
  for (long k = 0; k < n; k++) {
    const long i = i0 + k * stride;
    if (!((modemask >> (int) (i % 5)) & 1)) continue;	/* diagnostic runs on a subset of the modes */
    benchgen_make(seed,i,b,small);
    if (RES.on && RES.blocks) {
      const int64_t pL = place(RES.blocks,RES.cap_words,RES.used_nt,b->gsegL,b->glength);
      const int64_t pR = (b->mode == GMAPDP_GENOME) ? place(RES.blocks,RES.cap_words,RES.used_nt,b->gsegR,b->glengthR) : pL;
      if (pL < 0 || pR < 0) { free(b); return -1; }
      gmapdp_coords co;
      box_coords(b,(uint64_t) pL,(uint64_t) pR,(uint64_t) (RES.cap_words / 3) * 32,&co);
      GmapDP_batch_next_coords(batch,&co);
    }
    benchgen_add(batch,b); added++;
  }
  free(b);
  return added;
}

extern "C" int benchgen_box_size (void) { return (int) sizeof(benchgen_box); }
