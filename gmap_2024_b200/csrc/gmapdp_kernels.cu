/* gmapdp_kernels.cu -- sm_100a kernels and the C ABI of the GMAP alignment-DP engine.
 *
 * One persistent grid; every warp pulls DP boxes (largest first) from a global queue and runs the
 * whole device part of one reference entry point on it:
 *
 *   fill_full    Dynprog_simd_8 / Dynprog_simd_16            (/root/reference/src/dynprog_simd.c:2987 / :6562)
 *                32-row stripes, lane = row, lanes skewed by one column (systolic wavefront): the
 *                horizontal gap E stays in the lane's registers, the diagonal and the vertical gap
 *                (the reference's scalar F loop, :3391-3471) arrive by one warp shuffle each per step.
 *                Out-of-band lanes are computed exactly as the AVX2 code does (32 int8 lanes = one
 *                warp), which is what makes the 8-bit fill bit-exact.
 *   fill_tri<>   Dynprog_simd_{8,16}_upper / _lower          (:4304,:7714 / :5340,:8586)
 *                lane = row (upper) or column (lower), all lanes on the same step; one shuffle per step.
 *   best_end     find_best_endpoint_* (dynprog_end.c:143-560) fused into the fills.
 *   bridge_*     bridge_intron_gap_*_site_level (dynprog_genome.c:866), bridge_cdna_gap_*_ud (dynprog_cdna.c:123)
 *   tb_*         Dynprog_traceback_* (dynprog_simd.c:9154-9946) -> run-length edit script.
 *
 * Direction planes are written to HBM as packed 2-bit / 4-bit cells, 128 B per warp store
 * (layout in gmapdp_layout.h).  No tensor cores: the recurrence is a max-plus wavefront, not a
 * contraction.  Integer work is int32 VIADDMNMX / VIMNMX with explicit -128 / -32768 floors
 * (sm_100a has no hardware s8x4 saturating path).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "gmapdp_layout.h"
#include "gmapdp_tables.h"

#define WARPS_PER_BLOCK 4
#define BLOCK_THREADS (WARPS_PER_BLOCK * 32)
#define FULLMASK 0xffffffffu
#define NEG32 (-32768)

/* ------------------------------------------------------------------------------------------------ */
struct SideSeq {		/* one side of a box in DP coordinates */
  const uint8_t *Q, *G, *Ga;	/* forward arrays in the byte pool */
  int rlen, glen;
  bool rev;
  __device__ __forceinline__ int q (int r) const { return rev ? Q[rlen - r] : Q[r - 1]; }
  __device__ __forceinline__ int g (int c) const { return rev ? G[glen - c] : G[c - 1]; }
  __device__ __forceinline__ int ga (int c) const { return rev ? Ga[glen - c] : Ga[c - 1]; }
};

__device__ __forceinline__ int nt_class (int ch) {
  int u = ch & 0xDF;
  return u == 'A' ? 0 : (u == 'C' ? 1 : (u == 'G' ? 2 : (u == 'T' ? 3 : 4)));
}

__device__ __forceinline__ int prof_pick (uint32_t plo, uint32_t p4, int k) {
  /* sign-extended byte k (0..4) of the 8-byte profile: one PRMT.  Raw prmt.b32: the msb of a selector
     nibble replicates the selected byte's sign (the __byte_perm intrinsic masks that bit off). */
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(plo), "r"(p4), "r"((uint32_t) k * 0x1111u + 0x8880u));
  return (int) d;
}

__device__ __forceinline__ int clampi (int v, int lo, int hi) { return min(max(v,lo),hi); }

struct BestTrack { int bs, bk; };

struct TriPlanes {		/* output of one E-only fill */
  uint32_t *dirs, *sc;		/* sc may be NULL */
  EGeom g;
};

/* E-only fill.  LOWER=false: upper triangle, lane axis = query rows, step axis = genome columns.
 *               LOWER=true : lower triangle, lane axis = genome columns, step axis = query rows. */
template <bool LOWER>
__device__ void fill_tri (const SideSeq &sd, const uint8_t *stepcode, int band, int mt, int open, int extend,
			  bool late, int NEG, int POS, bool bits8, const TriPlanes &pl,
			  BestTrack *bt, bool lastrow, short *brow, const GdpTables *tb) {
  const int lane = threadIdx.x & 31;
  const int nA = pl.g.nA, nB = pl.g.nB;
  int s = 0;
  for (int i0 = 0; i0 <= nA; i0 += 32, s++) {
    const int i = i0 + lane;
    const bool rowact = (i <= nA);
    const int jlo = i0;
    int jhi = i0 + 31 + band; if (jhi > nB) jhi = nB;
    if (jlo > jhi) break;
    int jend = i + band; if (jend > nB) jend = nB;

    uint32_t plo = 0, p4 = 0;
    if (rowact) {
      if (!LOWER) {
	int q = (i == 0) ? 'N' : sd.q(i);
	const uint2 p = *reinterpret_cast<const uint2 *>(&tb->U[mt][q & 127][0]);
	plo = p.x; p4 = p.y;
      } else if (i == 0) {
	const uint2 p = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][bits8 ? 4 : 'N'][0]);
	plo = p.x; p4 = p.y;
      } else {
	const uint2 p = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][sd.g(i) & 127][0]);
	const uint2 pa = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][sd.ga(i) & 127][0]);
	plo = __vmaxs4(p.x,pa.x); p4 = __vmaxs4(p.y,pa.y);
      }
    }

    int E = NEG, Hcur = NEG, H = NEG;
    int bprev = (lane == 0 && i0 > 0) ? (int) brow[i0 - 1] : 0, bnext = 0;
    uint32_t dacc = 0, sacc = 0;
    uint32_t *dst = pl.dirs + (size_t) s * pl.g.dirW + lane;
    uint32_t *sst = pl.sc ? pl.sc + (size_t) s * pl.g.scW + lane : NULL;

    for (int j = jlo; j <= jhi; j++) {
      const int tt = j - jlo;
      int up = __shfl_up_sync(FULLMASK,Hcur,1);
      if (lane == 0) { up = bprev; if (i0 > 0) bnext = brow[j]; }
      const int code = stepcode[j];
      const bool act = rowact && j >= i && j <= jend;
      uint32_t bits = 0;
      if (act) {
	const int diag = (j == 0) ? 0 : (i == 0 ? NEG : up);
	const int sc = max(prof_pick(plo,p4,code & 15),prof_pick(plo,p4,code >> 4));
	const int Hd = clampi(diag + sc,NEG,POS);
	if (j == i) {
	  E = NEG; H = Hd;
	} else {
	  const int T1 = max(Hcur + open,NEG);
	  const bool dE = late ? (E >= T1) : (E > T1);
	  E = max(max(E,T1) + extend,NEG);
	  const bool dN = late ? (E >= Hd) : (E > Hd);
	  H = max(Hd,E);
	  bits = (dN ? 1u : 0u) | (dE ? 2u : 0u);
	}
	Hcur = H;
	if (bt) {
	  const int r = LOWER ? j : i, c = LOWER ? i : j;
	  if (r >= 1 && c >= 1 && (!LOWER || j > i) && (!lastrow || r == sd.rlen)) {
	    const int key = (r << 16) | c;
	    if (H > bt->bs || (H == bt->bs && (late ? key > bt->bk : key < bt->bk))) { bt->bs = H; bt->bk = key; }
	  }
	}
      }
      __syncwarp();
      if (lane == 31 && act) brow[j] = (short) H;
      if (lane == 0) bprev = bnext;
      dacc |= bits << (2 * (tt & 15));
      if ((tt & 15) == 15) { dst[(tt >> 4) * 32] = dacc; dacc = 0; }
      if (sst) {
	sacc |= ((uint32_t) H & 0xffffu) << (16 * (tt & 1));
	if (tt & 1) { sst[(tt >> 1) * 32] = sacc; sacc = 0; }
      }
    }
    const int nsteps = jhi - jlo + 1;
    if (nsteps & 15) dst[(nsteps >> 4) * 32] = dacc;
    if (sst && (nsteps & 1)) sst[(nsteps >> 1) * 32] = sacc;
    __syncwarp();
  }
}

/* 2-bit direction cell of an E-only fill; (i,j) = (lane-axis index, step-axis index).  Out of band -> 0. */
__device__ __forceinline__ uint32_t tri_dir (const TriPlanes &pl, int i, int j) {
  if (i < 0 || i > pl.g.nA || j < i || j > i + pl.g.band || j > pl.g.nB) return 0;
  const int s = i >> 5, l = i & 31, tt = j - (s << 5);
  const uint32_t w = pl.dirs[(size_t) s * pl.g.dirW + (tt >> 4) * 32 + l];
  return (w >> (2 * (tt & 15))) & 3u;
}
__device__ __forceinline__ int tri_score (const TriPlanes &pl, int i, int j) {
  const int s = i >> 5, l = i & 31, tt = j - (s << 5);
  const uint32_t w = pl.sc[(size_t) s * pl.g.scW + (tt >> 1) * 32 + l];
  return (int) (short) (w >> (16 * (tt & 1)));
}

/* Full fill: Dynprog_simd_8 / _16, stripe-faithful. */
__device__ void fill_full (const SideSeq &sd, const uint8_t *gcode, int lband, int uband, int mt, int open, int extend,
			   bool late, int NEG, int POS, uint32_t *dirs, const FGeom &fg,
			   short *Hrow, int *FF, int *corner, const GdpTables *tb) {
  const int lane = threadIdx.x & 31;
  const int rlen = sd.rlen, glen = sd.glen;
  int s = 0;
  for (int rlo = 0; rlo <= rlen; rlo += 32, s++) {
    const int rhigh = min(rlo + 31,rlen);
    const int r = rlo + lane;
    const bool rowact = (r <= rlen);
    const int c0 = max(0,rlo - lband);
    const int chigh = min(rhigh + uband,glen);
    if (c0 > chigh) continue;
    const int nsteps = (chigh - c0 + 1) + 31;

    uint32_t plo = 0, p4 = 0;
    if (rowact) {
      int q = (r == 0) ? 'N' : sd.q(r);
      const uint2 p = *reinterpret_cast<const uint2 *>(&tb->U[mt][q & 127][0]);
      plo = p.x; p4 = p.y;
    }
    int E = late ? NEG : NEG + 1;
    int Hl = NEG - open;
    int diag = NEG - open;
    int cg_out = NEG32, last_out = NEG32;
    uint32_t acc = 0;
    uint32_t *dst = dirs + (size_t) s * fg.dirW + lane;

    for (int tt = 0; tt < nsteps; tt++) {
      const int c = c0 + tt - lane;
      int cg_in = __shfl_up_sync(FULLMASK,cg_out,1);
      int last_in = __shfl_up_sync(FULLMASK,last_out,1);
      const bool act = rowact && c >= c0 && c <= chigh;
      uint32_t nib = 0;
      if (act) {
	int Hs;
	if (lane == 0) {
	  if (c == 0) Hs = (rlo == 0) ? 0 : NEG;
	  else Hs = (rlo == 0) ? NEG : (int) Hrow[c - 1];
	  if (rlo == 0 || c >= rlo + uband) { cg_in = NEG32; last_in = NEG32; }
	  else { cg_in = FF[c]; last_in = (int) Hrow[c]; }
	} else {
	  Hs = diag;
	}
	/* E (horizontal gap), dynprog_simd.c:3318-3334 */
	const int T1 = max(Hl + open,NEG);
	bool dE = late ? (E >= T1) : (E > T1);
	E = max(max(E,T1) + extend,NEG);
	/* H, :3350-3389 */
	int sc;
	if (c == 0) sc = (r == 0) ? 0 : NEG;
	else { const int code = gcode[c]; sc = max(prof_pick(plo,p4,code & 15),prof_pick(plo,p4,code >> 4)); }
	const int Hd = clampi(Hs + sc,NEG,POS);
	uint32_t dN = (late ? (E >= Hd) : (E > Hd)) ? 1u : 0u;
	int H = max(Hd,E);
	uint32_t dF = 0;
	const bool inband = (r >= c - uband) && (r <= c + lband);
	if (inband) {
	  if (r == c + lband && c > 0) { H = Hd; dE = false; dN = 0; }	/* bottom of band forced DIAG, :3395-3417 */
	  int cg, last;
	  if (r == c - uband) {						/* top of band, :3434-3446 */
	    cg = NEG32 + open + extend; last = H;
	  } else {
	    const int score = last_in + open;
	    if (late ? (cg_in >= score) : (cg_in > score)) { cg = cg_in + extend; dF = 1; }
	    else cg = score + extend;
	    last = H;
	    if (late ? (cg >= last) : (cg > last)) { last = cg; H = max(cg,NEG); dN = 2; }
	  }
	  cg_out = cg; last_out = last;
	} else {
	  last_out = H;
	}
	diag = last_in;
	Hl = H;
	if (lane == 31) { Hrow[c] = (short) H; FF[c] = cg_out; }
	if (r == rlen && c == glen) *corner = H;
	nib = dN | (dE ? 4u : 0u) | (dF ? 8u : 0u);
      }
      acc |= nib << (4 * (tt & 7));
      if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; }
    }
    if (nsteps & 7) dst[(nsteps >> 3) * 32] = acc;
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t full_dir (const uint32_t *dirs, const FGeom &fg, int r, int c) {
  if (r < c - fg.uband || r > c + fg.lband) return 0;
  const int s = r >> 5, l = r & 31;
  const int c0 = max(0,(s << 5) - fg.lband);
  const int tt = c - c0 + l;
  const uint32_t w = dirs[(size_t) s * fg.dirW + (tt >> 3) * 32 + l];
  return (w >> (4 * (tt & 7))) & 15u;
}

/* ------------------------------------------------------------------------------------------------
 * Tracebacks (lane 0).  Literal replays of the reference's loops, including the post-decrement
 * conditions (dynprog_simd.c:9175,9342,9459).
 * ---------------------------------------------------------------------------------------------- */
struct TbAcc {
  int score, nmatches, nmismatches, nopens, nindels;
  uint32_t *ops; int nops; int pend;
  __device__ __forceinline__ void flush () { if (pend) { ops[nops++] = ((uint32_t) pend << 2); pend = 0; } }
  __device__ __forceinline__ void gap (int kind, int dist, bool scored) {
    flush(); ops[nops++] = ((uint32_t) dist << 2) | (uint32_t) kind;
    if (scored) { score += -3 - dist; nopens += 1; nindels += dist; }
  }
};

__device__ __forceinline__ void tb_diag (TbAcc &a, const SideSeq &sd, int r, int c, const GdpTables *tb) {
  const int c1 = sd.q(r) & 127, c2 = sd.g(c) & 127, c2a = sd.ga(c) & 127;
  a.pend++;
  if (c2 == '*') return;
  if (c1 == c2 || c1 == c2a || ((tb->cons[c1][c2 >> 5] >> (c2 & 31)) & 1u) || ((tb->cons[c1][c2a >> 5] >> (c2a & 31)) & 1u)) {
    a.score += 1; a.nmatches += 1;
  } else {
    a.score += -3; a.nmismatches += 1;
  }
}

__device__ void tb_upper (TbAcc &a, const SideSeq &sd, const TriPlanes &pl, int r, int c, const GdpTables *tb) {
  while (r > 0 && c > 0) {
    if (tri_dir(pl,r,c) & 1u) {
      int dist = 1;
      for (;;) { const uint32_t e = tri_dir(pl,r,c) & 2u; c--; if (!e || c < 0) break; dist++; }
      a.gap(1,dist,dist < 9);
    } else { tb_diag(a,sd,r,c,tb); r--; c--; }
  }
  a.flush();
  if (c > 0) { a.score += (c < 9) ? (-3 - c) : 0; if (c < 9) { a.nopens += 1; a.nindels += c; } }
}

__device__ void tb_lower (TbAcc &a, const SideSeq &sd, const TriPlanes &pl, int r, int c, const GdpTables *tb) {
  while (r > 0 && c > 0) {
    if (tri_dir(pl,c,r) & 1u) {
      int dist = 1;
      for (;;) { const uint32_t e = tri_dir(pl,c,r) & 2u; r--; if (!e || r < 0) break; dist++; }
      a.gap(2,dist,true);
    } else { tb_diag(a,sd,r,c,tb); r--; c--; }
  }
  a.flush();
  if (r > 0) { a.score += -3 - r; a.nopens += 1; a.nindels += r; }
}

__device__ void tb_full (TbAcc &a, const SideSeq &sd, const uint32_t *dirs, const FGeom &fg, int r, int c, const GdpTables *tb) {
  while (r > 0 && c > 0) {
    const uint32_t nib = full_dir(dirs,fg,r,c);
    const uint32_t dir = nib & 3u;
    if (dir == 1u) {
      int dist = 1;
      for (;;) { if (!(c > 0)) break; const uint32_t e = full_dir(dirs,fg,r,c) & 4u; c--; if (!e) break; dist++; }
      a.gap(1,dist,dist < 9);
    } else if (dir == 2u) {
      int dist = 1;
      for (;;) { if (!(r > 0)) break; const uint32_t f = full_dir(dirs,fg,r,c) & 8u; r--; if (!f) break; dist++; }
      a.gap(2,dist,true);
    } else { tb_diag(a,sd,r,c,tb); r--; c--; }
  }
  a.flush();
  if (r == 0 && c == 0) {
  } else if (c == 0) { a.score += -3 - r; a.nopens += 1; a.nindels += r; }
  else if (c < 9) { a.score += -3 - c; a.nopens += 1; a.nindels += c; }
}

/* ------------------------------------------------------------------------------------------------
 * Bridges
 * ---------------------------------------------------------------------------------------------- */
struct Cand { int s; double p; unsigned long long ord; int rL, rR, cL, cR; };

__device__ __forceinline__ bool cand_better (const Cand &a, const Cand &b) {	/* a beats b */
  if (a.s != b.s) return a.s > b.s;
  if (a.p != b.p) return a.p > b.p;
  return a.ord < b.ord;
}

__device__ __forceinline__ Cand cand_shfl (const Cand &a, int srclane) {
  Cand o;
  o.s = __shfl_sync(FULLMASK,a.s,srclane); o.p = __shfl_sync(FULLMASK,a.p,srclane);
  o.ord = __shfl_sync(FULLMASK,a.ord,srclane);
  o.rL = __shfl_sync(FULLMASK,a.rL,srclane); o.rR = __shfl_sync(FULLMASK,a.rR,srclane);
  o.cL = __shfl_sync(FULLMASK,a.cL,srclane); o.cR = __shfl_sync(FULLMASK,a.cR,srclane);
  return o;
}

__device__ __forceinline__ int intron_points (const int *isc, int ldi, int rdi) {
  const int t = ldi & rdi;
  return t ? isc[31 - __clz(t)] : 0;
}

/* bridge_intron_gap_{8,16}_site_level, dynprog_genome.c:866-1386.  Returns finalscore. */
__device__ int bridge_genome (const gmapdp_box &b, const TriPlanes &LU, const TriPlanes &LL, const TriPlanes &RU,
			      const TriPlanes &RL, const uint8_t *ldi, const uint8_t *rdi, const double *lp,
			      const double *rp, int NEG, const int *isc, int *bestrL, int *bestrR, int *bestcL, int *bestcR) {
  const int lane = threadIdx.x & 31;
  const int rlength = b.rlenL, glengthL = b.glenL, glengthR = b.glenR;
  const int lbandL = b.lbandL, ubandL = b.ubandL, lbandR = b.lbandR, ubandR = b.ubandR;
  const int lim = b.offdiff;		/* rightoffset - leftoffset */
  Cand best, dn;			/* dn: best "with dinucleotide" candidate: p = prob sum, s = its score */
  best.s = NEG; best.p = 0.0; best.ord = 0; best.rL = best.rR = best.cL = best.cR = 0;
  dn = best; dn.p = 0.0;

  for (int rL = 1 + lane; rL < rlength; rL += 32) {
    const int rR = rlength - rL;
    unsigned long long ord = ((unsigned long long) rL << 32) + 1;
    int cloL = max(rL - lbandL,1), chighL = min(rL + ubandL,glengthL - 1);
    int cloR = max(rR - lbandR,1), chighR = min(rR + ubandR,glengthR - 1);
    int cL, cR, score, scoreL, scoreR, scoreI;
    double probL, probR;
#define CONSIDER() do { score = scoreL + scoreI + scoreR; ord++; \
      if (score > best.s || (score == best.s && probL + probR > best.p)) { \
	best.s = score; best.p = probL + probR; best.ord = ord; best.rL = rL; best.rR = rR; best.cL = cL; best.cR = cR; } } while (0)

    cL = rL; probL = lp[cL]; scoreL = tri_score(LU,rL,cL);
    cR = rR; probR = rp[cR]; scoreR = tri_score(RU,rR,cR);
    scoreI = intron_points(isc,ldi[cL],rdi[cR]);
    CONSIDER();
    if (scoreI > 0 && probL + probR > dn.p) {
      dn.s = scoreL + scoreI + scoreR; dn.p = probL + probR; dn.ord = ord; dn.rL = rL; dn.rR = rR; dn.cL = cL; dn.cR = cR;
    }
    const int ldiL = ldi[cL];
    const int scoreLdiag = scoreL, scoreRdiag = scoreR;
    const double probLdiag = probL, probRdiag = probR;

    /* indel on right */
    for (cR = cloR; cR < rR && cR < lim - cL; cR++) {
      probR = rp[cR]; scoreR = tri_score(RL,cR,rR); scoreI = intron_points(isc,ldiL,rdi[cR]); CONSIDER();
    }
    for (cR++; cR < chighR && cR < lim - cL; cR++) {
      probR = rp[cR]; scoreR = tri_score(RU,rR,cR); scoreI = intron_points(isc,ldiL,rdi[cR]); CONSIDER();
    }
    /* indel on left */
    cR = rR; probR = probRdiag; scoreR = scoreRdiag;
    const int rdiR = rdi[cR];
    for (cL = cloL; cL < rL && cL < lim - cR; cL++) {
      probL = lp[cL]; scoreL = tri_score(LL,cL,rL); scoreI = intron_points(isc,ldi[cL],rdiR); CONSIDER();
    }
    for (cL++; cL < chighL && cL < lim - cR; cL++) {
      probL = lp[cL]; scoreL = tri_score(LU,rL,cL); scoreI = intron_points(isc,ldi[cL],rdiR); CONSIDER();
    }
    (void) scoreLdiag; (void) probLdiag;
#undef CONSIDER
  }

  /* combine: max score, then max probability, then earliest in the reference's scan order */
  for (int off = 16; off > 0; off >>= 1) {
    Cand o = cand_shfl(best,lane ^ off);
    if (cand_better(o,best)) best = o;
    Cand d = cand_shfl(dn,lane ^ off);
    if (d.p > dn.p || (d.p == dn.p && d.p > 0.0 && d.ord < dn.ord)) dn = d;
  }

  int bestscore = best.s;
  bool use_dinucl;
  if (best.p > 2 * 0.85) use_dinucl = false;
  else if (dn.p == 0.0) use_dinucl = false;
  else if (dn.s < 0 || dn.s < bestscore - 9) use_dinucl = false;
  else use_dinucl = true;
  if (use_dinucl) { best = dn; bestscore = dn.s; }
  *bestrL = best.rL; *bestrR = best.rR; *bestcL = best.cL; *bestcR = best.cR;
  if (bestscore < 0) return bestscore;
  if (b.flags & GMAPDP_F_HALFP) return bestscore - intron_points(isc,ldi[best.cL],rdi[best.cR]) / 2;
  return bestscore;
}

/* bridge_cdna_gap_{8,16}_ud, dynprog_cdna.c:123-375.
 *
 * The reference scans (cL asc, cR desc, rL asc, rR asc) keeping the best score with >= (jump late:
 * the LAST maximum in scan order wins) or > (the FIRST wins), subject to rR < lim - rL.  That is an
 * argmax with a positional tie-break, so it is evaluated here as
 *   phase 1: for every cR a prefix-best table over rR (best score and its rR for rR <= k), which
 *            answers "best rR under the constraint" in O(1);
 *   phase 2: for every (cL, rL, cR) one table lookup; candidates carry their scan position as a key.
 * O(g^2 * band) instead of O(g^2 * band^2). */
__device__ int bridge_cdna (const gmapdp_box &b, const TriPlanes &LU, const TriPlanes &LL, const TriPlanes &RU,
			    const TriPlanes &RL, int NEG, uint32_t *pb, int *bestcL, int *bestcR, int *bestrL, int *bestrR) {
  const int lane = threadIdx.x & 31;
  const int glength = b.glenL, rlengthL = b.rlenL, rlengthR = b.rlenR;
  const int lbandL = b.lbandL, ubandL = b.ubandL, lbandR = b.lbandR, ubandR = b.ubandR;
  const int open = b.open, lim = b.offdiff;
  const bool late = (b.flags & GMAPDP_F_BRIDGE_LATE) != 0;
  const int PBW = lbandR + ubandR + 1;

  for (int cR = lane; cR <= glength; cR += 32) {
    const int lo = max(cR - ubandR,1), hi = min(cR + lbandR,rlengthR - 1);
    uint32_t *row = pb + (size_t) cR * PBW - (cR - ubandR);
    int bs = 0, br = 0; bool have = false;
    for (int rR = lo; rR <= hi; rR++) {
      const int sc = (rR < cR) ? tri_score(RU,rR,cR) : tri_score(RL,cR,rR);
      if (!have || (late ? sc >= bs : sc > bs)) { bs = sc; br = rR; have = true; }
      row[rR] = ((uint32_t) (bs + 32768) << 16) | (uint32_t) br;
    }
  }
  __syncwarp();

  int bs = NEG; unsigned long long bk = 0; bool have = false;
  int bcL = 0, bcR = 0, brL = 0, brR = 0;
  for (int cL = 1 + lane; cL < glength; cL += 32) {
    const int rloL = max(cL - ubandL,1), rhighL = min(cL + lbandL,rlengthL - 1);
    for (int rL = rloL; rL <= rhighL; rL++) {
      const int scoreL = (rL < cL) ? tri_score(LU,rL,cL) : tri_score(LL,cL,rL);
      const int rcap = lim - rL - 1;
      for (int cR = glength - cL; cR >= 0; cR--) {
	const int lo = max(cR - ubandR,1);
	const int k = min(min(cR + lbandR,rlengthR - 1),rcap);
	if (k < lo) continue;
	const uint32_t e = pb[(size_t) cR * PBW + (k - (cR - ubandR))];
	const int score = scoreL + ((int) (e >> 16) - 32768) + ((cR == glength - cL) ? 0 : open);
	const int rR = (int) (e & 0xffffu);
	const unsigned long long key = ((unsigned long long) (glength - cR) << 32) | ((unsigned long long) rL << 16) | (unsigned long long) rR;
	bool take;
	if (!have) take = late ? (score >= bs) : (score > bs);
	else if (score != bs) take = score > bs;
	else take = late ? (key > bk) : (key < bk);
	if (take) { bs = score; bk = key; bcL = cL; bcR = cR; brL = rL; brR = rR; have = true; }
      }
    }
  }
  /* lanes partition cL, which leads the scan order: late keeps the largest cL on ties, early the smallest */
  for (int off = 16; off > 0; off >>= 1) {
    const int os = __shfl_xor_sync(FULLMASK,bs,off), ocL = __shfl_xor_sync(FULLMASK,bcL,off), ocR = __shfl_xor_sync(FULLMASK,bcR,off);
    const int orL = __shfl_xor_sync(FULLMASK,brL,off), orR = __shfl_xor_sync(FULLMASK,brR,off);
    const int oh = __shfl_xor_sync(FULLMASK,(int) have,off);
    bool take = false;
    if (oh) {
      if (!have) take = true;
      else if (os != bs) take = os > bs;
      else take = late ? (ocL > bcL) : (ocL < bcL);
    }
    if (take) { bs = os; bcL = ocL; bcR = ocR; brL = orL; brR = orR; have = true; }
  }
  *bestcL = bcL; *bestcR = bcR; *bestrL = brL; *bestrR = brR;
  return bs;
}

/* ------------------------------------------------------------------------------------------------
 * The kernel
 * ---------------------------------------------------------------------------------------------- */
struct KernelArgs {
  const gmapdp_box *boxes;
  const int *order;		/* boxes sorted by decreasing work */
  int nboxes;
  const uint8_t *seq;
  const double *probs;
  gmapdp_result *results;
  uint32_t *script; unsigned long long script_cap;
  unsigned long long *script_cursor;
  int *queue;
  uint32_t *ws; unsigned long long ws_words;	/* per warp */
  int smem_cols;		/* boundary-row capacity per warp */
  const GdpTables *tables;
};

__device__ void process_box (const KernelArgs &ka, int bi, uint32_t *ws, short *Hrow, int *FF, const GdpTables *tb) {
  const int lane = threadIdx.x & 31;
  const gmapdp_box b = ka.boxes[bi];
  const bool use8 = (b.flags & GMAPDP_F_USE8) != 0;
  const int NEG = use8 ? -128 : -32768, POS = use8 ? 127 : 32767;
  const int mt = b.mismatchtype, open = b.open, extend = b.extend;
  const bool lateL = (b.flags & GMAPDP_F_LATE_L) != 0, lateR = (b.flags & GMAPDP_F_LATE_R) != 0;

  SideSeq L, R;
  L.Q = ka.seq + b.qL_off; L.G = ka.seq + b.gL_off; L.Ga = ka.seq + b.gLalt_off; L.rlen = b.rlenL; L.glen = b.glenL; L.rev = (b.revmask & 1) != 0;
  R.Q = ka.seq + b.qR_off; R.G = ka.seq + b.gR_off; R.Ga = ka.seq + b.gRalt_off; R.rlen = b.rlenR; R.glen = b.glenR; R.rev = (b.revmask & 2) != 0;

  /* carve the workspace */
  uint8_t *bytes = reinterpret_cast<uint8_t *>(ws);
  uint8_t *qcodeL = bytes; bytes += gdp_align4(b.rlenL + 2);
  uint8_t *qcodeR = bytes; bytes += gdp_align4(b.rlenR + 2);
  uint8_t *gcodeL = bytes; bytes += gdp_align4(b.glenL + 2);
  uint8_t *ldi = bytes; bytes += gdp_align4(b.glenL + 2);
  uint8_t *gcodeR = bytes; bytes += gdp_align4(b.glenR + 2);
  uint8_t *rdi = bytes; bytes += gdp_align4(b.glenR + 2);
  uint32_t *wp = reinterpret_cast<uint32_t *>(bytes);
  uint32_t *stage = wp; wp += b.rlenL + b.glenL + b.rlenR + b.glenR + 16;

  const bool twosided = (b.mode == GMAPDP_GENOME || b.mode == GMAPDP_CDNA);

  /* pre-pass: class codes in DP coordinates */
  for (int c = lane; c <= b.glenL; c += 32) gcodeL[c] = (c == 0) ? 0x44 : (uint8_t) (nt_class(L.g(c)) | (nt_class(L.ga(c)) << 4));
  if (b.mode != GMAPDP_SINGLE)
    for (int r = lane; r <= b.rlenL; r += 32) { const int k = (r == 0) ? 4 : nt_class(L.q(r)); qcodeL[r] = (uint8_t) (k | (k << 4)); }
  if (twosided) {
    for (int c = lane; c <= b.glenR; c += 32) gcodeR[c] = (c == 0) ? 0x44 : (uint8_t) (nt_class(R.g(c)) | (nt_class(R.ga(c)) << 4));
    for (int r = lane; r <= b.rlenR; r += 32) { const int k = (r == 0) ? 4 : nt_class(R.q(r)); qcodeR[r] = (uint8_t) (k | (k << 4)); }
  }
  if (b.mode == GMAPDP_GENOME) {
    /* dinucleotide classes, dynprog_genome.c:919-967 (forward arrays: rev_gsequenceR[-cR] = GR[glenR-1-cR]) */
    for (int cL = lane; cL <= b.glenL; cL += 32) {
      int v = 0;
      if (cL < b.glenL - 1) {
	const int a = L.G[cL], aa = L.Ga[cL], d = L.G[cL+1], da = L.Ga[cL+1];
	if ((a == 'G' || aa == 'G') && (d == 'T' || da == 'T')) v = 0x21;
	else if ((a == 'G' || aa == 'G') && (d == 'C' || da == 'C')) v = 0x10;
	else if ((a == 'A' || aa == 'A') && (d == 'T' || da == 'T')) v = 0x08;
	else if ((a == 'C' || aa == 'C') && (d == 'T' || da == 'T')) v = 0x06;
      }
      ldi[cL] = (uint8_t) v;
    }
    for (int cR = lane; cR <= b.glenR; cR += 32) {
      int v = 0;
      if (cR < b.glenR - 1) {
	const int r2 = R.G[b.glenR-2-cR], r2a = R.Ga[b.glenR-2-cR], r1 = R.G[b.glenR-1-cR], r1a = R.Ga[b.glenR-1-cR];
	if ((r2 == 'A' || r2a == 'A') && (r1 == 'G' || r1a == 'G')) v = 0x30;
	else if ((r2 == 'A' || r2a == 'A') && (r1 == 'C' || r1a == 'C')) v = 0x0C;
	else if ((r2 == 'G' || r2a == 'G') && (r1 == 'C' || r1a == 'C')) v = 0x02;
	else if ((r2 == 'A' || r2a == 'A') && (r1 == 'T' || r1a == 'T')) v = 0x01;
      }
      rdi[cR] = (uint8_t) v;
    }
  }
  __syncwarp();

  gmapdp_result res;
  memset(&res,0,sizeof(res));
  TbAcc acc;
  acc.score = acc.nmatches = acc.nmismatches = acc.nopens = acc.nindels = 0;
  acc.ops = stage; acc.nops = 0; acc.pend = 0;
  int lenA = 0, lenB = 0;

  if (b.mode == GMAPDP_SINGLE) {
    FGeom fg = fgeom(b.rlenL,b.glenL,b.lbandL,b.ubandL);
    uint32_t *dirs = wp;
    int corner = NEG;
    fill_full(L,gcodeL,b.lbandL,b.ubandL,mt,open,extend,lateL,NEG,POS,dirs,fg,Hrow,FF,&corner,tb);
    /* the corner cell's owner broadcasts it */
    {
      const int owner = b.rlenL & 31;
      res.finalscore = __shfl_sync(FULLMASK,corner,owner);
    }
    __syncwarp();
    if (lane == 0) { tb_full(acc,L,dirs,fg,b.rlenL,b.glenL,tb); lenA = acc.nops; }

  } else if (b.mode == GMAPDP_END5 || b.mode == GMAPDP_END3) {
    TriPlanes U, Lo;
    U.g = egeom(b.rlenL,b.glenL,b.ubandL); U.dirs = wp; U.sc = NULL; wp += (size_t) U.g.nstripes * U.g.dirW;
    Lo.g = egeom(b.glenL,b.rlenL,b.lbandL); Lo.dirs = wp; Lo.sc = NULL; wp += (size_t) Lo.g.nstripes * Lo.g.dirW;
    const bool lastrow = (b.flags & GMAPDP_F_LASTROW) != 0;
    BestTrack bt;
    if (lastrow) { bt.bs = NEG; bt.bk = (b.rlenL << 16); } else { bt.bs = 0; bt.bk = 0; }
    fill_tri<false>(L,gcodeL,b.ubandL,mt,open,extend,lateL,NEG,POS,use8,U,&bt,lastrow,Hrow,tb);
    fill_tri<true>(L,qcodeL,b.lbandL,mt,open,extend,lateL,NEG,POS,use8,Lo,&bt,lastrow,Hrow,tb);
    for (int off = 16; off > 0; off >>= 1) {
      const int os = __shfl_xor_sync(FULLMASK,bt.bs,off), ok = __shfl_xor_sync(FULLMASK,bt.bk,off);
      if (os > bt.bs || (os == bt.bs && (lateL ? ok > bt.bk : ok < bt.bk))) { bt.bs = os; bt.bk = ok; }
    }
    res.finalscore = bt.bs; res.bestrL = bt.bk >> 16; res.bestcL = bt.bk & 0xffff;
    __syncwarp();
    if (lane == 0 && !(b.flags & GMAPDP_F_NOTRACE)) {
      if (res.bestcL >= res.bestrL) tb_upper(acc,L,U,res.bestrL,res.bestcL,tb);
      else tb_lower(acc,L,Lo,res.bestrL,res.bestcL,tb);
      lenA = acc.nops;
    }

  } else {
    TriPlanes LU, LL, RU, RL;
    LU.g = egeom(b.rlenL,b.glenL,b.ubandL); LU.dirs = wp; wp += (size_t) LU.g.nstripes * LU.g.dirW; LU.sc = wp; wp += (size_t) LU.g.nstripes * LU.g.scW;
    LL.g = egeom(b.glenL,b.rlenL,b.lbandL); LL.dirs = wp; wp += (size_t) LL.g.nstripes * LL.g.dirW; LL.sc = wp; wp += (size_t) LL.g.nstripes * LL.g.scW;
    RU.g = egeom(b.rlenR,b.glenR,b.ubandR); RU.dirs = wp; wp += (size_t) RU.g.nstripes * RU.g.dirW; RU.sc = wp; wp += (size_t) RU.g.nstripes * RU.g.scW;
    RL.g = egeom(b.glenR,b.rlenR,b.lbandR); RL.dirs = wp; wp += (size_t) RL.g.nstripes * RL.g.dirW; RL.sc = wp; wp += (size_t) RL.g.nstripes * RL.g.scW;
    fill_tri<false>(L,gcodeL,b.ubandL,mt,open,extend,lateL,NEG,POS,use8,LU,NULL,false,Hrow,tb);
    fill_tri<true>(L,qcodeL,b.lbandL,mt,open,extend,lateL,NEG,POS,use8,LL,NULL,false,Hrow,tb);
    fill_tri<false>(R,gcodeR,b.ubandR,mt,open,extend,lateR,NEG,POS,use8,RU,NULL,false,Hrow,tb);
    fill_tri<true>(R,qcodeR,b.lbandR,mt,open,extend,lateR,NEG,POS,use8,RL,NULL,false,Hrow,tb);
    __syncwarp();
    int brL, brR, bcL, bcR, fs;
    if (b.mode == GMAPDP_GENOME) {
      const int di = b.cdna_direction > 0 ? 0 : (b.cdna_direction < 0 ? 1 : 2);
      fs = bridge_genome(b,LU,LL,RU,RL,ldi,rdi,ka.probs + b.probL_off,ka.probs + b.probR_off,NEG,
			 tb->isc[di][(b.flags & GMAPDP_F_FINALP) ? 1 : 0],&brL,&brR,&bcL,&bcR);
      if (fs < 0) res.status = 1;
    } else {
      fs = bridge_cdna(b,LU,LL,RU,RL,NEG,wp,&bcL,&bcR,&brL,&brR);
    }
    res.finalscore = fs; res.bestrL = brL; res.bestcL = bcL; res.bestrR = brR; res.bestcR = bcR;
    if (lane == 0 && res.status == 0) {
      if (bcR >= brR) tb_upper(acc,R,RU,brR,bcR,tb); else tb_lower(acc,R,RL,brR,bcR,tb);
      lenA = acc.nops;
      if (bcL >= brL) tb_upper(acc,L,LU,brL,bcL,tb); else tb_lower(acc,L,LL,brL,bcL,tb);
      lenB = acc.nops - lenA;
    }
  }

  /* publish: allocate script space, copy, write the result */
  const int ntot = __shfl_sync(FULLMASK,lenA + lenB,0);
  unsigned long long off = 0;
  if (lane == 0) off = atomicAdd(ka.script_cursor,(unsigned long long) ntot);
  off = __shfl_sync(FULLMASK,off,0);
  __syncwarp();
  if (off + ntot <= ka.script_cap)
    for (int k = lane; k < ntot; k += 32) ka.script[off + k] = stage[k];
  if (lane == 0) {
    res.tb_score = acc.score; res.nmatches = acc.nmatches; res.nmismatches = acc.nmismatches;
    res.nopens = acc.nopens; res.nindels = acc.nindels;
    res.script_off = (int32_t) off; res.script_lenA = lenA; res.script_lenB = lenB;
    ka.results[bi] = res;
  }
  __syncwarp();
}

extern __shared__ __align__(16) unsigned char dyn_smem[];

__global__ void __launch_bounds__(BLOCK_THREADS)
gmapdp_dp_kernel (KernelArgs ka) {
  /* shared: tables, then per-warp boundary rows */
  GdpTables *tb = reinterpret_cast<GdpTables *>(dyn_smem);
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ka.tables);
    uint32_t *dst = reinterpret_cast<uint32_t *>(tb);
    for (int k = threadIdx.x; k < (int) (sizeof(GdpTables) / 4); k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char *wbase = dyn_smem + ((sizeof(GdpTables) + 15) & ~15) + (size_t) warp * ((size_t) ka.smem_cols * 6);
  int *FF = reinterpret_cast<int *>(wbase);
  short *Hrow = reinterpret_cast<short *>(wbase + (size_t) ka.smem_cols * 4);
  const int gwarp = blockIdx.x * WARPS_PER_BLOCK + warp;
  uint32_t *ws = ka.ws + (size_t) gwarp * ka.ws_words;

  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(ka.queue,1);
    idx = __shfl_sync(FULLMASK,idx,0);
    if (idx >= ka.nboxes) break;
    process_box(ka,ka.order[idx],ws,Hrow,FF,tb);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Host side: context, memory, launches (the C ABI of include/gmapdp_b200.h)
 * ---------------------------------------------------------------------------------------------- */
struct gmapdp_ctx {
  int device, sm_count, grid, max_smem;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  std::string err;
  GdpTables *d_tables;
  /* device buffers (grown on demand) */
  gmapdp_box *d_boxes; size_t cap_boxes;
  int *d_order; size_t cap_order;
  uint8_t *d_seq; size_t cap_seq;
  double *d_probs; size_t cap_probs;
  gmapdp_result *d_results; size_t cap_results;
  uint32_t *d_script; size_t cap_script;
  unsigned long long *d_cursor;
  int *d_queue;
  uint32_t *d_ws; size_t cap_ws;
  /* pinned staging */
  void *h_pin; size_t cap_pin;
  /* resident batch */
  int nboxes; size_t ws_words; int smem_cols; size_t script_need;
  long launches;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return GMAPDP_ERR_CUDA; } } while (0)

template <typename T>
static int grow (gmapdp_ctx *ctx, T **p, size_t *cap, size_t need) {
  if (need <= *cap && *p) return GMAPDP_OK;
  if (*p) CK(cudaFree(*p));
  *p = NULL;
  size_t n = need + need / 4 + 64;
  CK(cudaMalloc((void **) p,n * sizeof(T)));
  *cap = n;
  return GMAPDP_OK;
}

extern "C" int gmapdp_create (gmapdp_ctx **out, int device) {
  gmapdp_ctx *ctx = new gmapdp_ctx();
  *out = ctx;
  ctx->device = device; ctx->launches = 0; ctx->nboxes = 0;
  ctx->d_tables = NULL; ctx->d_boxes = NULL; ctx->cap_boxes = 0; ctx->d_order = NULL; ctx->cap_order = 0; ctx->cap_results = 0; ctx->d_seq = NULL; ctx->cap_seq = 0;
  ctx->d_probs = NULL; ctx->cap_probs = 0; ctx->d_results = NULL; ctx->d_script = NULL; ctx->cap_script = 0;
  ctx->d_cursor = NULL; ctx->d_queue = NULL; ctx->d_ws = NULL; ctx->cap_ws = 0; ctx->h_pin = NULL; ctx->cap_pin = 0;
  ctx->stream = 0; ctx->ev0 = ctx->ev1 = 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    ctx->err = "no CUDA device: the gmapdp engine has no CPU fallback";
    return GMAPDP_ERR_CUDA;
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop,device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem = (int) prop.sharedMemPerBlockOptin;
  cudaFuncAttributes fa;
  cudaError_t fe = cudaFuncGetAttributes(&fa,gmapdp_dp_kernel);
  if (fe != cudaSuccess) {
    ctx->err = std::string("no sm_100a kernel image for this device (") + prop.name + "): " + cudaGetErrorString(fe);
    return GMAPDP_ERR_CUDA;
  }
  CK(cudaStreamCreateWithFlags(&ctx->stream,cudaStreamNonBlocking));
  CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1));
  GdpHostTables ht; GdpTables t; ht.device_tables(&t);
  CK(cudaMalloc((void **) &ctx->d_tables,sizeof(GdpTables)));
  CK(cudaMemcpy(ctx->d_tables,&t,sizeof(GdpTables),cudaMemcpyHostToDevice));
  CK(cudaMalloc((void **) &ctx->d_cursor,sizeof(unsigned long long)));
  CK(cudaMalloc((void **) &ctx->d_queue,sizeof(int)));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  ctx->grid = 0;
  return GMAPDP_OK;
}

extern "C" void gmapdp_destroy (gmapdp_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_tables); cudaFree(ctx->d_boxes); cudaFree(ctx->d_order); cudaFree(ctx->d_seq); cudaFree(ctx->d_probs);
  cudaFree(ctx->d_results); cudaFree(ctx->d_script); cudaFree(ctx->d_cursor); cudaFree(ctx->d_queue); cudaFree(ctx->d_ws);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char *gmapdp_last_error (const gmapdp_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" long gmapdp_launch_count (const gmapdp_ctx *ctx) { return ctx->launches; }

extern "C" int gmapdp_device_info (const gmapdp_ctx *ctx, int *sm_count, int *grid_blocks, int *block_threads) {
  if (sm_count) *sm_count = ctx->sm_count;
  if (grid_blocks) *grid_blocks = ctx->grid;
  if (block_threads) *block_threads = BLOCK_THREADS;
  return GMAPDP_OK;
}

static double box_work (const gmapdp_box &b) {
  if (b.mode == GMAPDP_SINGLE) return (double) (b.rlenL + 32) * (double) (std::min((int) b.glenL + 1,32 + b.lbandL + b.ubandL) + 31);
  double w = (double) (b.rlenL + 32) * (64 + b.ubandL) + (double) (b.glenL + 32) * (64 + b.lbandL);
  if (b.mode == GMAPDP_GENOME) w = 2 * w + (double) b.rlenL * 3 * (b.lbandL + b.ubandL + b.lbandR + b.ubandR);
  if (b.mode == GMAPDP_CDNA) w = 2 * w + 0.5 * (double) b.glenL * b.glenL * (b.lbandL + b.ubandL) * (b.lbandR + b.ubandR) * 0.25;
  return w;
}

extern "C" int gmapdp_upload (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
			      const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs) {
  if (!ctx || nboxes < 0 || (nboxes > 0 && (!boxes || !seqpool))) { if (ctx) ctx->err = "bad argument"; return GMAPDP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  ctx->nboxes = nboxes;
  if (nboxes == 0) return GMAPDP_OK;

  /* geometry of the batch */
  size_t ws_words = 0, script_need = 0; int maxcols = 0;
  std::vector<std::pair<double,int> > work(nboxes);
  for (int i = 0; i < nboxes; i++) {
    const gmapdp_box &b = boxes[i];
    if (b.rlenL < 0 || b.glenL < 0 || b.rlenR < 0 || b.glenR < 0 || b.open >= 0 || b.extend >= 0 || b.mode < 0 || b.mode > 4 ||
	(unsigned) b.mismatchtype > 3u) { ctx->err = "bad box"; return GMAPDP_ERR_ARG; }
    ws_words = std::max(ws_words,gdp_ws_words(b));
    script_need += (size_t) b.rlenL + b.glenL + 4;
    if (b.mode == GMAPDP_GENOME || b.mode == GMAPDP_CDNA) script_need += (size_t) b.rlenR + b.glenR + 4;
    maxcols = std::max(maxcols,std::max(std::max((int) b.glenL,(int) b.glenR),std::max((int) b.rlenL,(int) b.rlenR)) + 2);
    work[i] = std::make_pair(-box_work(b),i);
  }
  std::sort(work.begin(),work.end());
  std::vector<int> order(nboxes);
  for (int i = 0; i < nboxes; i++) order[i] = work[i].second;
  ctx->ws_words = (ws_words + 31) & ~(size_t) 31;
  ctx->smem_cols = (maxcols + 7) & ~7;
  ctx->script_need = script_need;

  /* persistent grid: as many blocks per SM as shared memory allows (<= 4), on every SM */
  size_t smem = ((sizeof(GdpTables) + 15) & ~15) + (size_t) WARPS_PER_BLOCK * ctx->smem_cols * 6;
  if ((int) smem > ctx->max_smem) { ctx->err = "box too long for the shared-memory boundary rows"; return GMAPDP_ERR_ARG; }
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel,BLOCK_THREADS,smem));
  if (occ < 1) occ = 1;
  if (occ > 6) occ = 6;
  int grid = ctx->sm_count * occ;
  int needed_blocks = (nboxes + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  if (grid > needed_blocks) grid = needed_blocks;
  ctx->grid = grid;

  if (grow(ctx,&ctx->d_boxes,&ctx->cap_boxes,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_order,&ctx->cap_order,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_results,&ctx->cap_results,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_seq,&ctx->cap_seq,seqbytes + 16)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_probs,&ctx->cap_probs,nprobs + 2)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_script,&ctx->cap_script,script_need + 64)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_ws,&ctx->cap_ws,(size_t) grid * WARPS_PER_BLOCK * ctx->ws_words)) return GMAPDP_ERR_CUDA;

  CK(cudaMemcpyAsync(ctx->d_boxes,boxes,(size_t) nboxes * sizeof(gmapdp_box),cudaMemcpyHostToDevice,ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_order,order.data(),(size_t) nboxes * sizeof(int),cudaMemcpyHostToDevice,ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_seq,seqpool,seqbytes,cudaMemcpyHostToDevice,ctx->stream));
  if (nprobs) CK(cudaMemcpyAsync(ctx->d_probs,probpool,nprobs * sizeof(double),cudaMemcpyHostToDevice,ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GMAPDP_OK;
}

extern "C" int gmapdp_run_resident (gmapdp_ctx *ctx, float *kernel_ms) {
  if (!ctx) return GMAPDP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (kernel_ms) *kernel_ms = 0.f;
  if (ctx->nboxes == 0) return GMAPDP_OK;
  KernelArgs ka;
  ka.boxes = ctx->d_boxes; ka.order = ctx->d_order; ka.nboxes = ctx->nboxes; ka.seq = ctx->d_seq; ka.probs = ctx->d_probs;
  ka.results = ctx->d_results; ka.script = ctx->d_script; ka.script_cap = ctx->cap_script; ka.script_cursor = ctx->d_cursor;
  ka.queue = ctx->d_queue; ka.ws = ctx->d_ws; ka.ws_words = ctx->ws_words; ka.smem_cols = ctx->smem_cols; ka.tables = ctx->d_tables;
  size_t smem = ((sizeof(GdpTables) + 15) & ~15) + (size_t) WARPS_PER_BLOCK * ctx->smem_cols * 6;
  CK(cudaMemsetAsync(ctx->d_cursor,0,sizeof(unsigned long long),ctx->stream));
  CK(cudaMemsetAsync(ctx->d_queue,0,sizeof(int),ctx->stream));
  CK(cudaEventRecord(ctx->ev0,ctx->stream));
  gmapdp_dp_kernel<<<ctx->grid,BLOCK_THREADS,smem,ctx->stream>>>(ka);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev1,ctx->stream));
  ctx->launches++;
  CK(cudaStreamSynchronize(ctx->stream));
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms,ctx->ev0,ctx->ev1));
  return GMAPDP_OK;
}

extern "C" int gmapdp_download (gmapdp_ctx *ctx, gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used) {
  if (!ctx) return GMAPDP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (script_used) *script_used = 0;
  if (ctx->nboxes == 0) return GMAPDP_OK;
  unsigned long long used = 0;
  CK(cudaMemcpyAsync(&used,ctx->d_cursor,sizeof(used),cudaMemcpyDeviceToHost,ctx->stream));
  CK(cudaMemcpyAsync(results,ctx->d_results,(size_t) ctx->nboxes * sizeof(gmapdp_result),cudaMemcpyDeviceToHost,ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (script_used) *script_used = (size_t) used;
  if (used > ctx->cap_script) { ctx->err = "device script pool overflow"; return GMAPDP_ERR_CAPACITY; }
  if (used > script_cap) { ctx->err = "script buffer too small"; return GMAPDP_ERR_CAPACITY; }
  if (used) CK(cudaMemcpyAsync(script,ctx->d_script,(size_t) used * sizeof(uint32_t),cudaMemcpyDeviceToHost,ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return GMAPDP_OK;
}

extern "C" int gmapdp_run_batch (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
				 const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs,
				 gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used) {
  int rc = gmapdp_upload(ctx,boxes,nboxes,seqpool,seqbytes,probpool,nprobs);
  if (rc) return rc;
  rc = gmapdp_run_resident(ctx,NULL);
  if (rc) return rc;
  return gmapdp_download(ctx,results,script,script_cap,script_used);
}

/* pinned host memory helpers (so that the shim's pools are DMA-able without a staging copy) */
extern "C" void *gmapdp_host_alloc (size_t bytes) {
  void *p = NULL;
  if (cudaHostAlloc(&p,bytes ? bytes : 1,cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return NULL; }
  return p;
}
extern "C" void gmapdp_host_free (void *p) { if (p) cudaFreeHost(p); }
extern "C" int gmapdp_host_register (void *p, size_t bytes) {
  if (!p || !bytes) return GMAPDP_OK;
  if (cudaHostRegister(p,bytes,cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return GMAPDP_ERR_CUDA; }
  return GMAPDP_OK;
}
extern "C" int gmapdp_host_unregister (void *p) {
  if (!p) return GMAPDP_OK;
  if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return GMAPDP_ERR_CUDA; }
  return GMAPDP_OK;
}
