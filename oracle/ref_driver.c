/* TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * ref_driver.so: a thin ctypes-friendly front end over the UNMODIFIED reference objects
 * (oracle/_ref/libgmapref.so, compiled from /root/reference/src by oracle/Makefile).  It lets the
 * tests and the golden-vector generator (tests/golden/make_golden.py) call
 *   - the six SIMD fills      Dynprog_simd_{8,16}{,_upper,_lower}   dynprog_simd.h:15-133
 *   - the five entry points   Dynprog_single_gap   dynprog_single.h:23
 *                             Dynprog_genome_gap   dynprog_genome.h:23
 *                             Dynprog_cdna_gap     dynprog_cdna.h:17
 *                             Dynprog_end5_gap/end3_gap dynprog_end.h:24,46
 * on an in-memory genome (Sequence_genomic_new sequence.c:1063 + Genome_from_sequence genome.c:307)
 * and read back every out-parameter and the returned List_T of Pair_T (pairdef.h:12-51).
 * Only reference headers are included; nothing from the reference is copied.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bool.h"
#include "types.h"
#include "mode.h"
#include "list.h"
#include "pairdef.h"
#include "pairpool.h"
#include "sequence.h"
#include "genome.h"
#include "maxent_hr.h"
#include "dynprog.h"
#include "dynprog_simd.h"
#include "dynprog_single.h"
#include "dynprog_genome.h"
#include "dynprog_cdna.h"
#include "dynprog_end.h"

typedef struct {
  int querypos, genomepos, queryjump, genomejump, dynprogindex, introntype, gapp;
  char cdna, comp, genome, genomealt;
  double donor_prob, acceptor_prob;
} refdrv_pair;

typedef struct {
  Genome_T genome;
  Sequence_T seq;
  int length;
} refdrv_genome;

static Dynprog_T dynL, dynR, dynM;
static Pairpool_T pool;
static int inited = 0;

/* --indel-open / --indel-extend: the three set-up calls of gmap.c:6547-6553 with the user's penalties */
void refdrv_set_user_dynprog (int open, int extend, int enabled) {
  Dynprog_single_setup(open,extend,enabled ? true : false,false);
  Dynprog_genome_setup(true,NULL,NULL,-1,-1,open,extend,enabled ? true : false);
  Dynprog_end_setup(NULL,NULL,NULL,0,NULL,NULL,NULL,NULL,open,extend,enabled ? true : false);
}

int refdrv_init (int maxlookback, int extraquerygap, int maxpeelback, int extramaterial_end, int extramaterial_paired) {
  if (inited) return 0;
  Dynprog_init(STANDARD);
  Dynprog_single_setup(0,0,false,false);
  Dynprog_genome_setup(true,NULL,NULL,-1,-1,0,0,false);
  Dynprog_end_setup(NULL,NULL,NULL,0,NULL,NULL,NULL,NULL,0,0,false);
  dynL = Dynprog_new(maxlookback,extraquerygap,maxpeelback,extramaterial_end,extramaterial_paired,true);
  dynR = Dynprog_new(maxlookback,extraquerygap,maxpeelback,extramaterial_end,extramaterial_paired,true);
  dynM = Dynprog_new(maxlookback,extraquerygap,maxpeelback,extramaterial_end,extramaterial_paired,false);
  pool = Pairpool_new();
  inited = 1;
  return 0;
}

int refdrv_max_rlength (void) { return dynL->max_rlength; }
int refdrv_max_glength (void) { return dynL->max_glength; }

/* tables, for pinning the restatement's own tables */
int refdrv_pairdistance (int mismatchtype, int a, int b) { return pairdistance_array[mismatchtype][a][b]; }
int refdrv_consistent (int a, int b) { return consistent_array[0][a][b]; }
int refdrv_use8p_size (int mismatchtype) { return use8p_size[mismatchtype]; }

void *refdrv_genome_new (const char *chars, int len) {
  refdrv_genome *g = (refdrv_genome *) malloc(sizeof(*g));
  g->seq = Sequence_genomic_new((char *) chars,len,/*copyp*/true);
  g->genome = Genome_from_sequence(g->seq);
  g->length = len;
  return g;
}

void refdrv_get_segment (void *gh, int leftp, unsigned int coord, int length, unsigned int bound, int revcomp,
			 char *seg, char *segalt) {
  refdrv_genome *g = (refdrv_genome *) gh;
  if (leftp) {
    Genome_get_segment_left(seg,segalt,g->genome,g->genome,coord,length,bound,revcomp ? true : false);
  } else {
    Genome_get_segment_right(seg,segalt,g->genome,g->genome,coord,length,bound,revcomp ? true : false);
  }
}

/* the compressed genome as the reference holds it (Genomecomp_T: high, low, flags per 32 nt) */
const unsigned int *refdrv_genome_blocks (void *gh, unsigned long *nwords) {
  refdrv_genome *g = (refdrv_genome *) gh;
  *nwords = ((unsigned long) g->length / 32UL + 1UL) * 3UL;
  return (const unsigned int *) Genome_blocks(g->genome);
}

/* which: 0 donor, 1 acceptor, 2 antidonor, 3 antiacceptor */
double refdrv_maxent (void *gh, int which, unsigned int pos, unsigned int chroffset) {
  refdrv_genome *g = (refdrv_genome *) gh;
  switch (which) {
  case 0: return Maxent_hr_donor_prob(g->genome,g->genome,pos,chroffset);
  case 1: return Maxent_hr_acceptor_prob(g->genome,g->genome,pos,chroffset);
  case 2: return Maxent_hr_antidonor_prob(g->genome,g->genome,pos,chroffset);
  default: return Maxent_hr_antiacceptor_prob(g->genome,g->genome,pos,chroffset);
  }
}

/* ---- fill level ------------------------------------------------------------------------------
 * kind: 0 full, 1 upper, 2 lower.  rseq/gseq/galt are forward arrays; with revp the reference is
 * handed pointers to their LAST chars, as the mode files do (dynprog_genome.c:3441,3533).
 * Outputs are dense (rlength+1) x (glength+1) row-major planes; only in-band cells are written,
 * the rest keep the caller's fill value. */
int refdrv_fill (int kind, int bits, const char *rseq, const char *gseq, const char *galt,
		 int rlength, int glength, int mismatchtype, int open, int extend,
		 int lband, int uband, int jump_late_p, int revp,
		 short *H, signed char *dN, signed char *dE, signed char *dF) {
  char *rp = (char *) (revp ? rseq + rlength - 1 : rseq);
  char *gp = (char *) (revp ? gseq + glength - 1 : gseq);
  char *ap = (char *) (revp ? galt + glength - 1 : galt);
  int r, c, G1 = glength + 1;
  bool late = jump_late_p ? true : false, rv = revp ? true : false;

  if (kind == 0) {
    if (bits == 8) {
      Score8_T **m; Direction8_T **n, **e, **f;
      m = Dynprog_simd_8(&n,&e,&f,dynM,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,lband,uband,late,rv);
      for (c = 0; c <= glength; c++) for (r = 0; r <= rlength; r++)
	if (r >= c - uband && r <= c + lband) { H[r*G1+c] = m[c][r]; dN[r*G1+c] = n[c][r]; dE[r*G1+c] = e[c][r]; dF[r*G1+c] = f[c][r]; }
    } else {
      Score16_T **m; Direction16_T **n, **e, **f;
      m = Dynprog_simd_16(&n,&e,&f,dynM,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,lband,uband,late,rv);
      for (c = 0; c <= glength; c++) for (r = 0; r <= rlength; r++)
	if (r >= c - uband && r <= c + lband) { H[r*G1+c] = m[c][r]; dN[r*G1+c] = (signed char) n[c][r]; dE[r*G1+c] = (signed char) e[c][r]; dF[r*G1+c] = (signed char) f[c][r]; }
    }
  } else if (kind == 1) {
    if (bits == 8) {
      Score8_T **m; Direction8_T **n, **e;
      m = Dynprog_simd_8_upper(&n,&e,dynL,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,uband,late,rv);
      for (r = 0; r <= rlength; r++) for (c = r; c <= glength && c <= r + uband; c++) {
	H[r*G1+c] = m[c][r]; dN[r*G1+c] = n[c][r]; dE[r*G1+c] = e[c][r]; }
    } else {
      Score16_T **m; Direction16_T **n, **e;
      m = Dynprog_simd_16_upper(&n,&e,dynL,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,uband,late,rv);
      for (r = 0; r <= rlength; r++) for (c = r; c <= glength && c <= r + uband; c++) {
	H[r*G1+c] = m[c][r]; dN[r*G1+c] = (signed char) n[c][r]; dE[r*G1+c] = (signed char) e[c][r]; }
    }
  } else {
    if (bits == 8) {
      Score8_T **m; Direction8_T **n, **e;
      m = Dynprog_simd_8_lower(&n,&e,dynL,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,lband,late,rv);
      for (c = 0; c <= glength; c++) for (r = c; r <= rlength && r <= c + lband; r++) {
	H[r*G1+c] = m[r][c]; dN[r*G1+c] = n[r][c]; dE[r*G1+c] = e[r][c]; }
    } else {
      Score16_T **m; Direction16_T **n, **e;
      m = Dynprog_simd_16_lower(&n,&e,dynL,rp,gp,ap,rlength,glength,(Mismatchtype_T) mismatchtype,open,extend,lband,late,rv);
      for (c = 0; c <= glength; c++) for (r = c; r <= rlength && r <= c + lband; r++) {
	H[r*G1+c] = m[r][c]; dN[r*G1+c] = (signed char) n[r][c]; dE[r*G1+c] = (signed char) e[r][c]; }
    }
  }
  return 0;
}

/* ---- entry-point level ------------------------------------------------------------------------ */
static int dump_pairs (List_T pairs, refdrv_pair *out, int maxpairs) {
  int n = 0;
  List_T p;
  if (pairs == NULL) return -1;
  for (p = pairs; p != NULL; p = p->rest) {
    Pair_T q = (Pair_T) p->first;
    if (n < maxpairs) {
      out[n].querypos = q->querypos; out[n].genomepos = (int) q->genomepos;
      out[n].queryjump = q->queryjump; out[n].genomejump = q->genomejump;
      out[n].dynprogindex = q->dynprogindex; out[n].introntype = q->introntype; out[n].gapp = q->gapp;
      out[n].cdna = q->cdna; out[n].comp = q->comp; out[n].genome = q->genome; out[n].genomealt = q->genomealt;
      out[n].donor_prob = q->donor_prob; out[n].acceptor_prob = q->acceptor_prob;
    }
    n++;
  }
  return n;
}

/* iout: dynprogindex(in/out), finalscore, nmatches, nmismatches, nopens, nindels */
int refdrv_single_gap (void *gh, int *iout, const char *queryseq, const char *queryuc,
		       int rlength, int glength, int roffset, int goffset,
		       unsigned int chroffset, unsigned int chrhigh, int watsonp, int jump_late_p,
		       int extraband_single, int widebandp, double defect_rate,
		       refdrv_pair *pairs, int maxpairs) {
  refdrv_genome *g = (refdrv_genome *) gh;
  List_T l;
  Pairpool_reset(pool);
  l = Dynprog_single_gap(&iout[0],&iout[1],&iout[2],&iout[3],&iout[4],&iout[5],dynM,
			 (char *) &queryseq[roffset],(char *) &queryuc[roffset],rlength,glength,roffset,goffset,
			 chroffset,chrhigh,watsonp ? true : false,/*genestrand*/0,jump_late_p ? true : false,
			 g->genome,g->genome,pool,extraband_single,widebandp ? true : false,defect_rate);
  return dump_pairs(l,pairs,maxpairs);
}

/* end5p: 1 for Dynprog_end5_gap (roffset/goffset are then rev_roffset/rev_goffset) */
int refdrv_end_gap (void *gh, int end5p, int *iout, const char *queryseq, const char *queryuc,
		    int rlength, int glength, int roffset, int goffset,
		    unsigned int chroffset, unsigned int chrhigh, int watsonp, int jump_late_p,
		    int extraband_end, double defect_rate, int endalign, int require_pos_score_p,
		    refdrv_pair *pairs, int maxpairs) {
  refdrv_genome *g = (refdrv_genome *) gh;
  List_T l;
  Pairpool_reset(pool);
  if (end5p) {
    l = Dynprog_end5_gap(&iout[0],&iout[1],&iout[2],&iout[3],&iout[4],&iout[5],dynL,
			 (char *) &queryseq[roffset],(char *) &queryuc[roffset],rlength,glength,roffset,goffset,
			 chroffset,chrhigh,watsonp ? true : false,0,jump_late_p ? true : false,
			 g->genome,g->genome,pool,extraband_end,defect_rate,(Endalign_T) endalign,
			 require_pos_score_p ? true : false);
  } else {
    l = Dynprog_end3_gap(&iout[0],&iout[1],&iout[2],&iout[3],&iout[4],&iout[5],dynL,
			 (char *) &queryseq[roffset],(char *) &queryuc[roffset],rlength,glength,roffset,goffset,
			 chroffset,chrhigh,watsonp ? true : false,0,jump_late_p ? true : false,
			 g->genome,g->genome,pool,extraband_end,defect_rate,(Endalign_T) endalign,
			 require_pos_score_p ? true : false);
  }
  return dump_pairs(l,pairs,maxpairs);
}

/* iout: dynprogindex(in/out), new_leftgenomepos, new_rightgenomepos, traceback_score, nmatches, nmismatches,
         nopens, nindels, exonhead, introntype ; dout: left_prob, right_prob */
int refdrv_genome_gap (void *gh, int *iout, double *dout, const char *queryseq, const char *queryuc,
		       int rlength, int glengthL, int glengthR, int roffset, int goffsetL, int rev_goffsetR,
		       unsigned int chroffset, unsigned int chrhigh, int cdna_direction, int watsonp, int jump_late_p,
		       int extraband_paired, double defect_rate, int maxpeelback, int halfp, int finalp,
		       refdrv_pair *pairs, int maxpairs) {
  refdrv_genome *g = (refdrv_genome *) gh;
  List_T l;
  Pairpool_reset(pool);
  l = Dynprog_genome_gap(&iout[0],&iout[1],&iout[2],&dout[0],&dout[1],&iout[3],&iout[4],&iout[5],&iout[6],&iout[7],
			 &iout[8],&iout[9],dynL,dynR,(char *) &queryseq[roffset],(char *) &queryuc[roffset],
			 rlength,glengthL,glengthR,roffset,goffsetL,rev_goffsetR,/*chrnum*/1,chroffset,chrhigh,
			 cdna_direction,watsonp ? true : false,0,jump_late_p ? true : false,g->genome,g->genome,pool,
			 extraband_paired,defect_rate,maxpeelback,halfp ? true : false,finalp ? true : false);
  return dump_pairs(l,pairs,maxpairs);
}

/* iout: dynprogindex(in/out), finalscore(in/out: the reference leaves it unwritten on some exits), incompletep */
int refdrv_cdna_gap (void *gh, int *iout, const char *queryseq, const char *queryuc,
		     int rlengthL, int rlengthR, int glength, int roffsetL, int rev_roffsetR, int goffset,
		     unsigned int chroffset, unsigned int chrhigh, int watsonp, int jump_late_p,
		     int extraband_paired, double defect_rate,
		     refdrv_pair *pairs, int maxpairs) {
  refdrv_genome *g = (refdrv_genome *) gh;
  List_T l;
  bool incompletep = false;
  Pairpool_reset(pool);
  l = Dynprog_cdna_gap(&iout[0],&iout[1],&incompletep,dynL,dynR,
		       (char *) &queryseq[roffsetL],(char *) &queryuc[roffsetL],
		       (char *) &queryseq[rev_roffsetR],(char *) &queryuc[rev_roffsetR],
		       rlengthL,rlengthR,glength,roffsetL,rev_roffsetR,goffset,
		       chroffset,chrhigh,watsonp ? true : false,0,jump_late_p ? true : false,
		       g->genome,g->genome,pool,extraband_paired,defect_rate);
  iout[2] = incompletep;
  return dump_pairs(l,pairs,maxpairs);
}
