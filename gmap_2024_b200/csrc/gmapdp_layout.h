/* Workspace geometry shared by the host (sizing) and the device (addressing).
 *
 * Every warp of the persistent grid owns one workspace slab in HBM.  Direction planes (and, for the
 * bridged modes, score planes) are laid out STRIPE-MAJOR so that each wavefront step of a warp
 * is one coalesced store:
 *
 *   E-only fill (upper / lower triangle): lanes = 32 consecutive positions of the lane axis
 *     (rows for upper, columns for lower), step tt = j - 32*s.  Per stripe
 *        dirs  : ceil(T/16) x 32 words, 2 bits per cell  (bit0 directions_nogap != DIAG, bit1 directions_Egap != DIAG)
 *        scores: ceil(T/2)  x 32 words, int16 per cell
 *     with T = 32 + band.
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

struct EGeom {		/* one E-only fill */
  int nA, nB, band;	/* lane-axis length, step-axis length, band width on the step axis */
  int nstripes, T, dirW, scW;
};

GDP_HD EGeom egeom (int nA, int nB, int band) {
  EGeom g;
  g.nA = nA; g.nB = nB; g.band = band;
  g.nstripes = (nA + 32) / 32;
  g.T = 32 + band;
  g.dirW = ((g.T + 15) / 16) * 32;
  g.scW = ((g.T + 1) / 2) * 32;
  return g;
}

struct FGeom {		/* one full fill */
  int rlen, glen, lband, uband;
  int nstripes, T, dirW;
};

GDP_HD FGeom fgeom (int rlen, int glen, int lband, int uband) {
  FGeom g;
  int w = 32 + lband + uband;
  g.rlen = rlen; g.glen = glen; g.lband = lband; g.uband = uband;
  g.nstripes = (rlen + 32) / 32;
  if (w > glen + 1) w = glen + 1;
  g.T = w + 31;
  g.dirW = ((g.T + 7) / 8) * 32;
  return g;
}

GDP_HD size_t gdp_align4 (size_t x) { return (x + 3) & ~(size_t) 3; }

/* words of per-warp workspace one box needs */
GDP_HD size_t gdp_ws_words (const gmapdp_box &b) {
  size_t w = 0;
  /* class-code arrays (bytes): step-axis codes for both sides, query and genome; dinucleotide arrays */
  size_t bytes = gdp_align4(b.rlenL + 2) + gdp_align4(b.rlenR + 2) + 2 * gdp_align4(b.glenL + 2) + 2 * gdp_align4(b.glenR + 2);
  w += bytes / 4;
  /* script staging */
  w += (size_t) (b.rlenL + b.glenL + b.rlenR + b.glenR + 16);
  if (b.mode == GMAPDP_SINGLE) {
    FGeom f = fgeom(b.rlenL,b.glenL,b.lbandL,b.ubandL);
    w += (size_t) f.nstripes * f.dirW;
  } else {
    bool scores = (b.mode == GMAPDP_GENOME || b.mode == GMAPDP_CDNA);
    EGeom lu = egeom(b.rlenL,b.glenL,b.ubandL), ll = egeom(b.glenL,b.rlenL,b.lbandL);
    w += (size_t) lu.nstripes * (lu.dirW + (scores ? lu.scW : 0));
    w += (size_t) ll.nstripes * (ll.dirW + (scores ? ll.scW : 0));
    if (scores) {
      EGeom ru = egeom(b.rlenR,b.glenR,b.ubandR), rl = egeom(b.glenR,b.rlenR,b.lbandR);
      w += (size_t) ru.nstripes * (ru.dirW + ru.scW);
      w += (size_t) rl.nstripes * (rl.dirW + rl.scW);
      if (b.mode == GMAPDP_CDNA) w += (size_t) (b.glenL + 1) * (b.lbandR + b.ubandR + 1);	/* prefix-best table of the cDNA bridge */
    }
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
