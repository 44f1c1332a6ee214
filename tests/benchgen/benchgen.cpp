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
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "../../include/gmapdp_shim.h"

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

extern "C" void benchgen_make (uint64_t seed, long i, benchgen_box *b, int small) {
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
  b->gsegL[b->glength] = '\0';
  if (b->mode == GMAPDP_GENOME) b->gsegR[b->glengthR] = '\0';
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
    benchgen_make(seed,i,b,small); benchgen_add(batch,b); added++;
  }
  free(b);
  return added;
}

extern "C" int benchgen_box_size (void) { return (int) sizeof(benchgen_box); }
