/* Workspace geometry shared by the host (sizing) and the device (addressing).
 *
 * Every warp of the persistent grid owns one workspace slab in HBM.  Direction planes (and, for the
 * bridged modes, score planes) are laid out STRIPE-MAJOR so that each wavefront step of a warp
 * is one coalesced store:
 *
 *   E-only fills (upper / lower triangle): diagonal lane mapping, see TriPacking below.
 *        dirs  : 2 bits per cell  (bit0 directions_nogap != DIAG, bit1 directions_Egap != DIAG)
 *        scores: int16 per cell
 *   Full fill: lanes = 32 rows, skewed by one column per lane (systolic), step tt = (c - c0) + lane.
 *        dirs  : ceil(T/8) x 32 words, 4 bits per cell (bits0-1 nogap: 0 DIAG 1 HORIZ 2 VERT, bit2 Egap, bit3 Fgap)
 *     with T = min(glength+1, 32 + lband + uband) + 31.
 */
#ifndef GMAPDP_LAYOUT_H
#define GMAPDP_LAYOUT_H

#include "../../include/gmapdp_b200.h"

#if defined(__CUDACC__)
#define GDP_HD __host__ __device__ __forceinline__
#else
#define GDP_HD inline
#endif

/* E-only fills (upper / lower triangle) run on a DIAGONAL lane mapping: lane = band diagonal
 * d = j - i (i = lane-axis index: row for upper, column for lower; j = step-axis index), all lanes of a
 * fill sit on the same step-axis index j = t at time t, lane d on lane-axis index i = t - d.
 * H(i-1,j-1) is the lane's own previous value, (H,E)(i,j-1) arrives from lane d-1 by one shuffle.
 * Several fills of one box share a pass when their widths (band+1) fit into 32 lanes; a fill wider
 * than 32 diagonals takes ceil((band+1)/32) passes of its own, chained through an edge array.
 * One pass owns one plane pair:  dirs [t>>4][32 lanes] (2 bits per cell), scores [t>>1][32 lanes] (int16). */
#define GDP_MAXFILLS 4

struct TriPacking {
  int nf;
  int band[GDP_MAXFILLS], nA[GDP_MAXFILLS], nB[GDP_MAXFILLS];
  int pass0[GDP_MAXFILLS], npass[GDP_MAXFILLS], lane0[GDP_MAXFILLS];
  int npasses, Tmax, maxA;
  int dirPW, scPW;		/* words per pass plane */
};

/* at most GDP_PASS_FILLS fills share a pass: the interior steps stage each fill's operands in its own shared-memory slot
   (bulk async copies, tri_fast in gmapdp_kernels.cu) and a warp has two slots per buffer */
#define GDP_PASS_FILLS 2
GDP_HD void tri_pack (TriPacking &tp) {
  int used = 0, pass = -1, inpass = 0;
  tp.Tmax = 0; tp.maxA = 0;
  /* fixed trip count and guards instead of `f < tp.nf' in the loop header: on the device the arrays then
     stay in registers (every index is a compile-time constant after unrolling) */
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int f = 0; f < GDP_MAXFILLS; f++) {
    if (f < tp.nf) {
      const int w = tp.band[f] + 1;
      if (tp.nB[f] > tp.Tmax) tp.Tmax = tp.nB[f];
      if (tp.nA[f] > tp.maxA) tp.maxA = tp.nA[f];
      if (w > 32) {
	tp.pass0[f] = pass + 1; tp.npass[f] = (w + 31) / 32; tp.lane0[f] = 0;
	pass += tp.npass[f]; used = 32; inpass = GDP_PASS_FILLS;
      } else if (pass >= 0 && used + w <= 32 && inpass < GDP_PASS_FILLS) {
	tp.pass0[f] = pass; tp.npass[f] = 1; tp.lane0[f] = used; used += w; inpass++;
      } else {
	pass++; tp.pass0[f] = pass; tp.npass[f] = 1; tp.lane0[f] = 0; used = w; inpass = 1;
      }
    }
  }
  tp.npasses = pass + 1;
  tp.dirPW = ((tp.Tmax >> 4) + 1) * 32;
  tp.scPW = ((tp.Tmax >> 1) + 1) * 32;
}

/* the fills of a box, in packing order: lower fills first (they are the narrow ones in production) */
GDP_HD void tri_fills_of (const gmapdp_box &b, TriPacking &tp) {
  const bool two = (b.mode == GMAPDP_GENOME || b.mode == GMAPDP_CDNA);
  /* constant indices (see tri_pack): two-sided boxes L lower, R lower, L upper, R upper; else L lower, L upper */
  tp.band[0] = b.lbandL; tp.nA[0] = b.glenL; tp.nB[0] = b.rlenL;
  if (two) {
    tp.band[1] = b.lbandR; tp.nA[1] = b.glenR; tp.nB[1] = b.rlenR;
    tp.band[2] = b.ubandL; tp.nA[2] = b.rlenL; tp.nB[2] = b.glenL;
    tp.band[3] = b.ubandR; tp.nA[3] = b.rlenR; tp.nB[3] = b.glenR;
    tp.nf = 4;
  } else {
    tp.band[1] = b.ubandL; tp.nA[1] = b.rlenL; tp.nB[1] = b.glenL;
    tp.band[2] = 0; tp.nA[2] = 0; tp.nB[2] = 0; tp.band[3] = 0; tp.nA[3] = 0; tp.nB[3] = 0;
    tp.pass0[2] = tp.pass0[3] = 0; tp.npass[2] = tp.npass[3] = 0; tp.lane0[2] = tp.lane0[3] = 0;
    tp.nf = 2;
  }
  tri_pack(tp);
}

struct FGeom {		/* one full fill */
  int rlen, glen, lband, uband;
  int nstripes, T, dirW;
  /* packed layout (16-bit boxes on the s16x2 path, fill_full_pk): stripe PAIRS of 64 rows, lane l holds rows l and l + 32
     of the pair, step t = (c - c0) + l + 32 * half; per pair ceil(T2/8) x 2 x 32 words */
  int pk, npairs, T2, dirW2;
};

/* padding of the full fill's per-warp column entries in shared memory: the packed path indexes them with columns
   -64 .. glen + 64 + 7 (lanes ahead of / behind the band read harmless entries instead of testing their column) */
#define GDP_PK_PAD 64
#define GDP_PK_EXTRA (GDP_PK_PAD + 64 + 16)
/* sides up to which the packed path's floor argument holds (fill_full_pk) */
#define GDP_PK_MAXSIDE 2500

GDP_HD FGeom fgeom (int rlen, int glen, int lband, int uband) {
  FGeom g;
  int w = 32 + lband + uband;
  g.rlen = rlen; g.glen = glen; g.lband = lband; g.uband = uband;
  g.nstripes = (rlen + 32) / 32;
  if (w > glen + 1) w = glen + 1;
  g.T = w + 31;
  g.dirW = ((g.T + 7) / 8) * 32;
  int w2 = 64 + lband + uband;
  if (w2 > glen + 1) w2 = glen + 1;
  g.npairs = (rlen + 64) / 64;
  g.T2 = w2 + 63;
  g.dirW2 = ((g.T2 + 7) / 8) * 64;
  g.pk = 0;
  return g;
}

/* which single-gap boxes take the packed path: 16-bit arithmetic, no alternate genome, sides within the floor argument */
GDP_HD bool gdp_full_packed (const gmapdp_box &b) {
  return (b.flags & GMAPDP_F_USE8) == 0 && b.gLalt_off == b.gL_off && b.rlenL <= GDP_PK_MAXSIDE && b.glenL <= GDP_PK_MAXSIDE;
}

#define GDP_ENT_PAD 128
GDP_HD int gdp_ent_len (const gmapdp_box &b) {
  int n = b.rlenL;
  if (b.rlenR > n) n = b.rlenR;
  if (b.glenL > n) n = b.glenL;
  if (b.glenR > n) n = b.glenR;
  return n + 4;
}

GDP_HD size_t gdp_align4 (size_t x) { return (x + 3) & ~(size_t) 3; }
GDP_HD size_t gdp_align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

/* words of per-warp workspace one box needs */
GDP_HD size_t gdp_ws_words (const gmapdp_box &b) {
  size_t w = 0;
  /* class-code arrays (bytes): step-axis codes for both sides, query and genome; dinucleotide arrays */
  size_t bytes = gdp_align16(b.rlenL + 2) + gdp_align16(b.rlenR + 2) + 2 * gdp_align16(b.glenL + 2) + 2 * gdp_align16(b.glenR + 2);
  /* ready-made PRMT selectors (uint16) of the same positions, read 16 at a time by the interior steps of the E-only fills */
  bytes += gdp_align16(2 * (size_t) (b.rlenL + 2)) + gdp_align16(2 * (size_t) (b.rlenR + 2)) + gdp_align16(2 * (size_t) (b.glenL + 2)) + gdp_align16(2 * (size_t) (b.glenR + 2));
  /* segments decoded from the resident genome */
  if (b.gflags & GMAPDP_G_SEG_L) bytes += gdp_align16(b.glenL + 2);
  if (b.gflags & GMAPDP_G_SEG_R) bytes += gdp_align16(b.glenR + 2);
  /* genome gaps: main-diagonal scores of the two upper fills (int16) */
  bytes += gdp_align16(2 * (size_t) (b.rlenL + 2)) + gdp_align16(2 * (size_t) (b.rlenR + 2));
  /* genome gaps: the same scores and the dinucleotide classes as one word per position (what the staged interior steps
     read), GDP_ENT_PAD entries in front of each array: a staged window may begin before entry 0 */
  if (b.mode == GMAPDP_GENOME) bytes += 2 * gdp_align16(4 * (size_t) (GDP_ENT_PAD + gdp_ent_len(b)));
  w += bytes / 4;
  /* script staging */
  w += (size_t) (b.rlenL + b.glenL + b.rlenR + b.glenR + 16);
  if (b.mode == GMAPDP_SINGLE) {
    FGeom f = fgeom(b.rlenL,b.glenL,b.lbandL,b.ubandL);
    const size_t d1 = (size_t) f.nstripes * f.dirW, d2 = (size_t) f.npairs * f.dirW2;
    w += d1 > d2 ? d1 : d2;
    w += 2 * (size_t) (b.glenL + 2 + GDP_PK_EXTRA) + 2;			/* stripe boundary row (8 bytes per column) */
  } else {
    const bool scores = (b.mode == GMAPDP_CDNA);	/* genome gaps evaluate their bridge inside the fills */
    TriPacking tp;
    tri_fills_of(b,tp);
    for (int f = 0; f < tp.nf; f++) w += 2 * (size_t) (tp.nA[f] + 2 + 32) + 2;	/* profile tables (+ look-ahead padding), 16-byte aligned */
    w += (size_t) tp.npasses * (tp.dirPW + (scores ? tp.scPW : 0));
    w += (size_t) tp.maxA + 2;							/* edge array of wide fills */
    if (b.mode == GMAPDP_GENOME) w += 8 * 32;					/* tie lists of the bridge (GEN_TIECAP keys per lane) */
    if (b.mode == GMAPDP_GENOME && (b.gflags & GMAPDP_G_PROBS)) w += 2 * (size_t) (b.glenL + b.glenR + 2) + 2;	/* MaxEnt arrays */
    if (b.mode == GMAPDP_CDNA) w += (size_t) (b.glenL + 1 + b.rlenL + 1) * (b.lbandR + b.ubandR + 1);	/* prefix-best tables of the cDNA bridge: per column, per rL */
  }
  return w + 64;
}

/* algorithmic in-band cells */
GDP_HD long gdp_cells_full (int r, int g, int lband, int uband) {
  long n = 0;
  for (int c = 0; c <= g; c++) {
    int lo = c - uband; if (lo < 0) lo = 0;
    int hi = c + lband; if (hi > r) hi = r;
    if (hi >= lo) n += hi - lo + 1;
  }
  return n;
}
GDP_HD long gdp_cells_tri (int nA, int nB, int band) {	/* upper: (r,g,uband); lower: (g,r,lband) */
  long n = 0;
  for (int i = 0; i <= nA; i++) {
    int hi = i + band; if (hi > nB) hi = nB;
    if (hi >= i) n += hi - i + 1;
  }
  return n;
}

#endif
