/* gmapdp_kernels.cu -- sm_100a kernels and the C ABI of the GMAP alignment-DP engine.
 *
 * Four persistent grids, one per kind of box (single / end / genome / cdna gaps); every warp pulls DP boxes
 * (largest first) from its kind's queue and runs the whole device part of one reference entry point on it:
 *
 *   fill_full    Dynprog_simd_8 / Dynprog_simd_16            (/root/reference/src/dynprog_simd.c:2987 / :6562)
 *                32-row stripes, lane = row, lanes skewed by one column (systolic wavefront): the
 *                horizontal gap E stays in the lane's registers, the diagonal and the vertical gap
 *                (the reference's scalar F loop, :3391-3471) arrive by one warp shuffle each per step.
 *                Out-of-band lanes are computed exactly as the AVX2 code does (32 int8 lanes = one
 *                warp), which is what makes the 8-bit fill bit-exact.
 *   tri_pass<>   Dynprog_simd_{8,16}_upper / _lower          (:4304,:7714 / :5340,:8586)
 *                lane = band diagonal, all lanes on the same step; one shuffle per step; a branch-free
 *                interior loop (tri_fast) and a general step for heads, tails and wide fills.
 *   best end     find_best_endpoint_* (dynprog_end.c:143-560), fused into the fills (end gaps).
 *   bridges      bridge_intron_gap_*_site_level (dynprog_genome.c:866) evaluated inside the fills (genome gaps),
 *                bridge_cdna_gap_*_ud (dynprog_cdna.c:123) from prefix-best tables (cdna gaps).
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
#include <mutex>
#include <thread>
#include <chrono>
#include <algorithm>

#include "gmapdp_layout.h"
#include "gmapdp_genome.h"
#include "gmapdp_tables.h"

#ifndef WARPS_PER_BLOCK
#define WARPS_PER_BLOCK 4
#endif
#ifndef GMAPDP_BND_GLOBAL
#define GMAPDP_BND_GLOBAL 0	/* 1: single-gap kernel keeps its stripe boundary rows in the HBM workspace (L1) instead of shared
				   memory, which lifts its occupancy from 3 to 4-6 blocks/SM.  Measured (500 k single-gap boxes):
				   shared 69.3 ms; workspace 75.0 (4 blocks, 127 registers), 69.4 (5 blocks, 96), 67.6 (6 blocks, 80):
				   the kernel is bound by its instruction count, not by latency -- left off. */
#endif
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

/* One E-only fill (upper or lower triangle) and where its planes live; see TriPacking (gmapdp_layout.h). */
struct TriFill {
  int nA, nB, band, lane0, pass0, npass;
  int lateadd;			/* 1: ties go to the gap (jump late, >=), 0: > */
  bool lower, right;
  const uint2 *prof;		/* [0..nA] 8-byte score profiles of the lane-axis positions */
  const uint8_t *code;		/* [0..nB] class codes of the step-axis positions (class | alt class << 4) */
  const uint16_t *sel;		/* [0..nB] the same as ready-made PRMT selectors (no alt genome), 16-byte aligned */
  uint32_t *dirs, *sc;		/* planes of pass `pass0'; the fill's further passes follow at dirPW / scPW */
  int dirPW, scPW;
};

/* 2-bit direction cell; (i,j) = (lane-axis index, step-axis index).  Out of band -> 0. */
__device__ __forceinline__ uint32_t tri_dir (const TriFill &f, int i, int j) {
  const int d = j - i;
  if (i < 0 || i > f.nA || d < 0 || d > f.band || j > f.nB) return 0;
  const uint32_t w = f.dirs[(size_t) (d >> 5) * f.dirPW + (j >> 4) * 32 + f.lane0 + (d & 31)];
  return (w >> (2 * (j & 15))) & 3u;
}
__device__ __forceinline__ int tri_score (const TriFill &f, int i, int j) {
  const int d = j - i;
  const uint32_t w = f.sc[(size_t) (d >> 5) * f.scPW + (j >> 1) * 32 + f.lane0 + (d & 31)];
  return (int) (short) (w >> (16 * (j & 1)));
}

/* score profiles of the lane axis: upper = raw query char x genome class (dynprog_simd.c:4424-4449),
   lower = query class x raw genome char with the alt genome folded in (:5459-5524, :8690) */
__device__ void tri_profiles (const TriFill &f, const SideSeq &sd, int mt, bool bits8, uint2 *prof, const GdpTables *tb) {
  const int lane = threadIdx.x & 31;
  for (int i = lane; i <= f.nA; i += 32) {
    uint2 p;
    if (!f.lower) {
      const int q = (i == 0) ? 'N' : sd.q(i);
      p = *reinterpret_cast<const uint2 *>(&tb->U[mt][q & 127][0]);
    } else if (i == 0) {
      p = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][bits8 ? 4 : 'N'][0]);
    } else {
      const uint2 a = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][sd.g(i) & 127][0]);
      const uint2 c = *reinterpret_cast<const uint2 *>(&tb->Lw[mt][sd.ga(i) & 127][0]);
      p.x = __vmaxs4(a.x,c.x); p.y = __vmaxs4(a.y,c.y);
    }
    prof[i] = p;
  }
}

/* One pass: every lane owns one band diagonal of one of the (up to GDP_MAXFILLS) fills packed into it.
 *
 * Steps come in two kinds.  The general step (head and tail of a pass, wide fills, boxes with an alt genome)
 * checks per lane whether its diagonal has started / ended; its profile and class-code loads run one step
 * ahead of their use (the profile tables are padded by 32 entries in front so that the look-ahead of a
 * not-yet-active lane stays inside the allocation).  The interior step -- every lane of the pass is inside
 * its diagonal, which is > 90 % of the steps of a long box -- needs no such checks: 16 steps are unrolled
 * around one direction word, the column's score is ONE PRMT with a ready-made selector (16 selectors arrive
 * in two 16-byte loads), loads use immediate offsets, the tie rule is a template parameter when the fills of
 * the pass agree on it, and the best-endpoint search keeps (score, step) instead of (score, key). */
/* Genome gaps: the splice-site bridge (bridge_intron_gap_{8,16}_site_level, dynprog_genome.c:866-1386) is evaluated
 * INSIDE the fills.  Every candidate of the reference's scan pairs one cell (row, col) of one of the four matrices
 * with the main-diagonal cell of the opposite side on the partner row rP = rlength - row:
 *   score = H(row,col) + intron score(dinucleotide at col on this side, at rP on the other) + diagonal score of the other side at rP
 * The two diagonals are computed beforehand (a sequential clamp-add scan, one lane per side), so a cell is
 * scored the moment the fill produces it and no score plane is written or read back.  "Higher score, then
 * higher probability sum, then first in scan order" is an argmax: every lane keeps its own best under that
 * order and the lanes are combined at the end.  key = rL << 15 | segment << 12 | col orders the candidates as
 * the reference visits them (segment: 0 diagonal, 1 / 2 right lower / upper, 3 / 4 left lower / upper).
 * The probability sum (two doubles) matters only between candidates that tie on the score: inside the fills a lane
 * just appends the key of a candidate that ties with its current best to a small per-lane list (a predicated
 * store: the steps stay branch-free) and restarts the list on a strict improvement; the list is resolved once,
 * after the fills, and only for the lanes that reached the warp's best score (flat stretches far from the splice
 * site tie all the time, but cannot win).  If the list of such a lane overflowed (long runs of equal best scores:
 * low-complexity sequence) the warp repeats the fills with every tie resolved on the spot (gen_tie). */
#ifndef GEN_TIECAP
#define GEN_TIECAP 8		/* at most 8: the workspace reserves 8 x 32 words */
#endif
struct GenBest { int s, key, cnt; double p; uint32_t *ties; };	/* p < 0: not fetched yet; ties[j * 32]: j-th tied key of this lane */
struct GenCtx {
  const uint8_t *ldi, *rdi;		/* dinucleotide classes per genomic position */
  const short *dgL, *dgR;		/* main-diagonal scores of the upper fills */
  const uint32_t *entL, *entR;		/* dg << 16 | di per position: what the staged interior steps read */
  const double *lp, *rp;
  int rlength, lim;
  /* MaxEnt from the resident genome (GMAPDP_G_PROBS): entry c of the left / right array is the probability of kind
     lkind / rkind at coordinate lpos0 + lstep c / rpos0 + rstep c for c < glength - 1, 0 beyond (the arrays of
     dynprog_genome.c:970-1061); a pre-pass of the box evaluates them into the workspace (gen_maxent), lp / rp point there.
     (Evaluating only the few entries the bridge's last step reads was measured slower: 114 vs 98 ms for the genome
     kernel of the benchmark launch -- single lanes walking dependent L2 loads -- against 57 ms with uploaded arrays.) */
  GdpGenome genome; const double *me;
  uint32_t lpos0, rpos0, chroffset;
  int lstep, rstep, lkind, rkind, nlp, nrp;
  uint32_t it0, it1;			/* intron scores as a byte table: byte k+1 = score of bit k, byte 0 = 0 */
};

__device__ __noinline__ double gen_maxent (const GenCtx &g, int right, int c) {
  if (right) return (c >= 0 && c < g.nrp) ? gdp_maxent_prob(g.rkind,g.genome,g.me,g.rpos0 + (uint32_t) (g.rstep * c),g.chroffset) : 0.0;
  return (c >= 0 && c < g.nlp) ? gdp_maxent_prob(g.lkind,g.genome,g.me,g.lpos0 + (uint32_t) (g.lstep * c),g.chroffset) : 0.0;
}

__device__ __forceinline__ int gen_points (const GenCtx &g, int di) {
  uint32_t pos, d;
  asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(di));		/* 0xffffffff if di == 0 */
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(g.it0), "r"(g.it1), "r"((pos + 1u) & 7u));
  return (int) d;
}

/* probability sum of the candidate behind a key */
__device__ __forceinline__ double gen_key_prob (const GenCtx &g, int key) {
  const int rL = key >> 15, seg = (key >> 12) & 7, col = key & 4095;
  if (seg == 0) return g.lp[rL] + g.rp[g.rlength - rL];
  if (seg <= 2) return g.lp[rL] + g.rp[col];
  return g.lp[col] + g.rp[g.rlength - rL];
}

__device__ __forceinline__ void gen_tie (const GenCtx &g, GenBest &b, int key) {	/* same score as the lane's best */
  if (b.key < 0) {				/* nothing yet: the reference takes a candidate only if its probability sum is > 0 */
    const double pc = gen_key_prob(g,key);
    if (pc > 0.0) { b.key = key; b.p = pc; }
    return;
  }
  const double pc = gen_key_prob(g,key);
  if (b.p < 0.0) b.p = gen_key_prob(g,b.key);
  if (pc > b.p || (pc == b.p && key < b.key)) { b.key = key; b.p = pc; }
}

/* ---- Shared-memory staging of the interior steps' operands --------------------------------------------------------
 * The interior steps of an E-only fill (tri_fast) read, per step and lane, an 8-byte score profile, a PRMT selector and
 * -- genome gaps -- the dinucleotide class / diagonal score of the lane's own column and of its partner row.  These are
 * per-box arrays in the warp's HBM workspace; read with LDG in the steps' dependent chain they cost the kernels their
 * issue slots (long-scoreboard stalls, ncu).  They are staged instead: per pass, the steps run in blocks of STG_B; the
 * windows of the arrays that a block's steps touch are copied into the warp's shared-memory slots by bulk async copies
 * (cp.async.bulk global -> shared, completion counted by an mbarrier), one block ahead of their use (two buffers), and
 * the steps read them with LDS.  A pass holds at most two fills (GDP_PASS_FILLS): two slots per buffer.
 *
 * Window of block [t0, t0 + STG_B) (t0 a multiple of 16; lane diagonals d <= 31):
 *   profiles   prof[t0 - 32 .. t0 + STG_B)                    read at prof[t - d]
 *   selectors  sel[t0 .. t0 + STG_B)                          read at sel[t]
 *   own        ent_own[t0 - 32 .. t0 + STG_B)                 read at col = lower ? t - d : t
 *   partner    ent_oth[w0 .. w0 + STG_B + 36), w0 = ((rlength + 1) & ~3) - t0 - STG_B      read at rP = rlength - row
 * ent = diagonal score << 16 | dinucleotide class, one word per position (process_box).  All sources are 16-byte aligned
 * and all sizes multiples of 16, as the bulk copy requires.
 * Per-warp region: two mbarriers (16 bytes), the barriers' phase bits (a word), padding to 32 bytes, two buffers. */
/* Measured (B200, the 1 M-box benchmark launch; results and scripts identical): with staging the end / genome / cdna
   kernels take 20.9 / 86.3 / 23.2 ms, without 19.5 / 79.1 / 21.9 ms, the production-size stratum 35.0 vs 31.6 ms: the copies
   cost more issue slots (one UBLKCP per lane and array, serialised through uniform registers, plus the barrier traffic
   per 64 steps) than the LDS reads save over L1-hit LDGs, and the 116 - 164 KB of shared memory per SM come out of the
   L1 that the general steps, the bridges and the tracebacks live on.  So it is a build option, off by default
   (-DGMAPDP_STAGE=1; the parity suite passes in both builds). */
#ifndef GMAPDP_STAGE
#define GMAPDP_STAGE 0
#endif
#define STG_B 64
#define STG_PROF_BYTES ((STG_B + 32) * 8)
#define STG_SEL_BYTES (STG_B * 2)
#define STG_OWN_BYTES ((STG_B + 32) * 4)
#define STG_PAR_BYTES ((STG_B + 36) * 4)
template <bool EVAL> struct StgGeom {
  static constexpr int NARR = EVAL ? 4 : 2;
  static constexpr uint32_t SLOT = STG_PROF_BYTES + STG_SEL_BYTES + (EVAL ? STG_OWN_BYTES + STG_PAR_BYTES : 0);
  static constexpr uint32_t BUF = GDP_PASS_FILLS * SLOT;
  static constexpr uint32_t WARP_BYTES = 32 + 2 * BUF;
};
#define STG_WARP_BYTES_MAX (StgGeom<true>::WARP_BYTES)

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init (uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx (uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
	       :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s (uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	       :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
/* a pointer the compiler knows to be shared memory (LDS with immediate offsets), from a shared-window address */
template <typename T> __device__ __forceinline__ const T *smem_ptr (uint32_t a) {
  return reinterpret_cast<const T *>(__cvta_shared_to_generic((size_t) a));
}
/* once per kernel and warp: the staging region's barriers */
__device__ __forceinline__ void stage_init (uint32_t stg) {
  if ((threadIdx.x & 31) == 0) {
    mbar_init(stg,1); mbar_init(stg + 8,1);
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(stg + 16), "r"(0u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

struct TriLane {
  /* staging (GMAPDP_STAGE): the warp's region, this lane's slot, and -- lanes 0 .. ncopies-1 -- the copy this lane issues per block */
  uint32_t stg; int slot, ncopies;
  const char *cp_src; int cp_stride; uint32_t cp_dst, cp_bytes, cp_total;
  const uint32_t *ent_own, *ent_oth;
  /* genome gaps: this lane's side of the bridge */
  const uint8_t *own_di, *oth_di; const short *oth_dg;
  int ev_lo, ev_hi, key0, kstep; bool ev_ok;
  int d, tstart, tend, lateadd;
  bool isd0, lower, edge_in, edge_out, trk;
  const uint2 *pp;			/* pp[t] = prof[t - d] */
  const uint8_t *code;
  const uint16_t *sel;
  int Hprev, H;
  uint32_t pk_out, dacc, sacc;
  uint32_t *dplane, *splane;
};

__device__ __forceinline__ uint32_t pack_lo16 (int lo, int hi) {	/* (lo & 0xffff) | (hi << 16): one PRMT */
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ int sext_lo16 (uint32_t w) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(d) : "r"(w));
  return (int) d;
}

/* LM: tie rule of the pass's fills -- 0 all `>' (jump early), 1 all `>=' (jump late), 2 per lane */
/* MODE: 0 end gaps (endpoint search), 1 cdna gaps (score planes for the bridge), 2 genome gaps (bridge inside the fill) */
template <int MODE, int LM>
__device__ __forceinline__ void tri_fast (TriLane &s, int &t, const int tstop, const int open, const int extend, const int NEG,
					  const int POS, const uint32_t negpair, BestTrack *bt, const GenCtx *gc, GenBest *gb) {
  constexpr bool SCORES = (MODE == 1), TRACK = (MODE == 0), EVAL = (MODE == 2);
#ifndef GMAPDP_END_UNROLL
#define GMAPDP_END_UNROLL 1
#endif
#ifndef GMAPDP_GENOME_UNROLL
#define GMAPDP_GENOME_UNROLL 1
#endif
#ifndef GMAPDP_GENOME_UNR
#define GMAPDP_GENOME_UNR 4
#endif
#ifndef GMAPDP_CDNA_UNR
#define GMAPDP_CDNA_UNR 4
#endif
#ifndef GMAPDP_END_UNR
#define GMAPDP_END_UNR 8
#endif
  /* steps per unrolled body (8, 4 or 2; 16 / UNR bodies per direction word) */
  constexpr int UNR = (MODE == 2) ? GMAPDP_GENOME_UNR : ((MODE == 1) ? GMAPDP_CDNA_UNR : GMAPDP_END_UNR);
  constexpr int HALVES_UNROLLED = (MODE == 0) ? GMAPDP_END_UNROLL : ((MODE == 2) ? GMAPDP_GENOME_UNROLL : 1);
  const int la = s.lateadd;
  const bool isd0 = s.isd0;
  int H = s.Hprev;
  uint32_t pk_out = s.pk_out;
  uint32_t *dplane = s.dplane, *splane = s.splane;
  int bs = 0, btt = -1;
  if (TRACK) bs = bt->bs;
  int key = 0;
  GenBest g; g.s = 0; g.key = -1; g.cnt = 0; g.p = 0.0; g.ties = NULL;
  GenCtx gcl; gcl.it0 = 0; gcl.it1 = 0;		/* the intron score table in registers (a store sits between its uses) */
#if GMAPDP_STAGE
  /* operands from the warp's staging region (see StgGeom): block 0 is copied now, block k + 1 while block k computes */
  using SG = StgGeom<EVAL>;
  const int lane = threadIdx.x & 31;
  const uint32_t stg = s.stg;
  const int dl = max(s.d,0);				/* idle lanes read slot data of diagonal 0 */
  const uint32_t slotbase = stg + 32 + (uint32_t) s.slot * SG::SLOT;
  const char *cpsrc = s.cp_src;
  uint32_t phases;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(phases) : "r"(stg + 16) : "memory");
  __syncwarp();
  if (lane == 0) mbar_expect_tx(stg,s.cp_total);
  if (lane < s.ncopies) { bulk_g2s(stg + 32 + s.cp_dst,cpsrc,s.cp_bytes,stg); cpsrc += s.cp_stride; }
  const uint2 *pq = smem_ptr<uint2>(slotbase + (uint32_t) (32 - dl) * 8u);
  const uint32_t *sq = smem_ptr<uint32_t>(slotbase + STG_PROF_BYTES);
  const uint32_t *oent = NULL, *pent = NULL;
  if (EVAL) {
    gcl.it0 = gc->it0; gcl.it1 = gc->it1;
    g = *gb;
    /* col = lower ? t - d : t;  partner row rP = rlength - (lower ? t : t - d), window start w0 = ((rlength + 1) & ~3) - t - STG_B */
    oent = smem_ptr<uint32_t>(slotbase + STG_PROF_BYTES + STG_SEL_BYTES + (uint32_t) (s.lower ? 32 - dl : 32) * 4u);
    pent = smem_ptr<uint32_t>(slotbase + STG_PROF_BYTES + STG_SEL_BYTES + STG_OWN_BYTES
			      + (uint32_t) (gc->rlength - ((gc->rlength + 1) & ~3) + STG_B + (s.lower ? 0 : dl)) * 4u);
    key = s.key0 + t * s.kstep;
  }
  int flip = (int) SG::BUF;				/* bytes from this block's buffer to the next block's */
#else
  const uint2 *pq = s.pp + t;
  const uint32_t *sq = reinterpret_cast<const uint32_t *>(s.sel + t);
  /* genome gaps: col advances with t, the partner row goes down */
  const uint8_t *odi = NULL, *pdi = NULL; const short *pdg = NULL;
  if (EVAL) {
    gcl.it0 = gc->it0; gcl.it1 = gc->it1;
    g = *gb;
    const int col = s.lower ? t - s.d : t, rP = gc->rlength - (s.lower ? t : t - s.d);
    odi = s.own_di + col; pdi = s.oth_di + rP; pdg = s.oth_dg + rP;
    key = s.key0 + t * s.kstep;
  }
#endif
  uint32_t tieoff = (uint32_t) g.cnt * 128u;		/* byte offset of the lane's next tie slot */
  int tb = t;
#if GMAPDP_STAGE
  for (int blk = 0; t + 16 <= tstop; blk++) {
    const int tblk = min(t + STG_B,tstop);
    __syncwarp();					/* every lane is done with the buffer the next copies overwrite */
    if (tblk + 16 <= tstop) {
      const uint32_t nb = (uint32_t) (~blk & 1);
      if (lane == 0) mbar_expect_tx(stg + 8u * nb,s.cp_total);
      if (lane < s.ncopies) { bulk_g2s(stg + 32 + nb * SG::BUF + s.cp_dst,cpsrc,s.cp_bytes,stg + 8u * nb); cpsrc += s.cp_stride; }
    }
    {
      const uint32_t cb = (uint32_t) (blk & 1);
      mbar_wait(stg + 8u * cb,(phases >> cb) & 1u);
      phases ^= 1u << cb;
    }
  for (; t + 16 <= tblk; t += 16) {
#else
  {
  for (; t + 16 <= tstop; t += 16) {
#endif
    uint32_t dacc = 0;
    /* two halves of 8 unrolled steps; the big two-sided kernels keep ONE copy of the half (instruction-cache
       footprint), the small end-gap kernel unrolls both */
#pragma unroll HALVES_UNROLLED
    for (int h = 0; h < 16 / UNR; h++) {
      uint32_t sw[4];
      if (UNR == 8) { const uint4 s4 = *reinterpret_cast<const uint4 *>(sq); sw[0] = s4.x; sw[1] = s4.y; sw[2] = s4.z; sw[3] = s4.w; }
      else if (UNR == 4) { const uint2 s2 = *reinterpret_cast<const uint2 *>(sq); sw[0] = s2.x; sw[1] = s2.y; sw[2] = 0; sw[3] = 0; }
      else { sw[0] = *sq; sw[1] = 0; sw[2] = 0; sw[3] = 0; }
      sq += UNR / 2;
      uint32_t d8 = 0;
      int Heven = 0;
#pragma unroll
      for (int u = 0; u < UNR; u++) {
	uint32_t pk_in = __shfl_up_sync(FULLMASK,pk_out,1);
	if (isd0) pk_in = negpair;
	const uint2 p = pq[u];
	const uint32_t selw = (u & 1) ? (sw[u >> 1] >> 16) : sw[u >> 1];
	uint32_t scu;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(scu) : "r"(p.x), "r"(p.y), "r"(selw));
	const int Hd = clampi(H + (int) scu,NEG,POS);
	const int Hl = sext_lo16(pk_in), El = ((int) pk_in) >> 16;
	const int T1 = max(Hl + open,NEG);
	const bool dE = (LM == 0) ? (El > T1) : ((LM == 1) ? (El >= T1) : (El + la > T1));
	const int E = max(max(El,T1) + extend,NEG);
	const bool dN = (LM == 0) ? (E > Hd) : ((LM == 1) ? (E >= Hd) : (E + la > Hd));
	const int Hn = max(Hd,E);
	pk_out = pack_lo16(Hn,E);
	d8 |= (dN ? (1u << (2 * u)) : 0u) | (dE ? (2u << (2 * u)) : 0u);
	if (SCORES) {
	  if (u & 1) { *splane = pack_lo16(Heven,Hn); splane += 32; }
	  else Heven = Hn;
	}
	if (TRACK) {
	  const bool up = s.trk && ((LM == 0) ? (Hn > bs) : ((LM == 1) ? (Hn >= bs) : (Hn + la > bs)));
	  bs = up ? Hn : bs; btt = up ? tb + u : btt;
	}
	if (EVAL) {
#if GMAPDP_STAGE
	  /* ent = diagonal score << 16 | dinucleotide class: the own class is a byte, so the AND needs no mask */
	  const uint32_t pe = pent[-u];
	  const int tot = Hn + gen_points(gcl,(int) (reinterpret_cast<const uint8_t *>(oent + u)[0] & pe)) + (((int) pe) >> 16);
#else
	  const int tot = Hn + gen_points(gcl,(int) odi[u] & (int) pdi[-u]) + (int) pdg[-u];
#endif
	  const int ku = key + u * s.kstep;
	  const bool gt = s.ev_ok && tot > g.s, eq = s.ev_ok && tot == g.s;
	  if (eq && tieoff < GEN_TIECAP * 128u) *reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(g.ties) + tieoff) = (uint32_t) ku;
	  tieoff = gt ? 0u : tieoff + (eq ? 128u : 0u);
	  g.key = gt ? ku : g.key; g.s = gt ? tot : g.s;
	}
	H = Hn;
      }
      pq += UNR; tb += UNR;
#if GMAPDP_STAGE
      if (EVAL) { oent += UNR; pent -= UNR; key += UNR * s.kstep; }
#else
      if (EVAL) { odi += UNR; pdi -= UNR; pdg -= UNR; key += UNR * s.kstep; }
#endif
      dacc |= d8 << (2 * UNR * h);
    }
    *dplane = isd0 ? 0u : dacc; dplane += 32;
  }
#if GMAPDP_STAGE
    /* on to the other buffer: within a block the addresses moved by STG_B entries */
    pq = reinterpret_cast<const uint2 *>(reinterpret_cast<const char *>(pq) + (flip - STG_B * 8));
    sq = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(sq) + (flip - STG_B * 2));
    if (EVAL) {
      oent = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(oent) + (flip - STG_B * 4));
      pent = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(pent) + (flip + STG_B * 4));
    }
    flip = -flip;
#endif
  }
#if GMAPDP_STAGE
  __syncwarp();
  if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" :: "r"(stg + 16), "r"(phases) : "memory");
#endif
  if (EVAL) g.cnt = (int) (tieoff >> 7);
  s.Hprev = H; s.H = H; s.pk_out = pk_out; s.dplane = dplane; s.splane = splane;
  if (EVAL) *gb = g;
  if (TRACK) {
    if (btt >= 0) {
      const int i = btt - s.d;
      const int r = s.lower ? btt : i, c = s.lower ? i : btt;
      bt->bs = bs; bt->bk = (r << 16) | c;
    }
  }
}

template <int MODE, bool WIDE>
__device__ __forceinline__ void tri_pass_body (const TriFill (&F)[GDP_MAXFILLS], int nf, int pass, int open, int extend, int NEG, int POS,
			  BestTrack *bt, int track_rlen, bool lastrow, uint32_t *edge, bool fastok, const GenCtx *gc, GenBest *gb,
			  bool ev_now, uint32_t stg) {
  constexpr bool SCORES = (MODE == 1), TRACK = (MODE == 0), EVAL = (MODE == 2);
  const int lane = threadIdx.x & 31;
  /* lane configuration */
  TriLane s;
  s.stg = stg; s.slot = 0; s.ncopies = 0; s.cp_src = NULL; s.cp_stride = 0; s.cp_dst = 0; s.cp_bytes = 0; s.cp_total = 0;
  s.ent_own = NULL; s.ent_oth = NULL;
  int nslots = 0, cp_arr = -1;				/* fills of this pass so far; which array this lane copies (staging) */
  (void) nslots; (void) cp_arr;
  int nA = -1, nB = -1, thi = -1, tlo = 0x7fffffff;
  int maxstart = 0, minend = 0x7fffffff, lm = -1;
  s.d = -1; s.lateadd = 0; s.lower = false; s.edge_in = false; s.edge_out = false;
  s.ev_ok = false; s.ev_lo = 0x7fffffff; s.ev_hi = -1; s.key0 = 0; s.kstep = 0; s.own_di = NULL; s.oth_di = NULL; s.oth_dg = NULL;
  const uint2 *prof = NULL; s.code = NULL; s.sel = NULL;
  const uint2 *prof0 = NULL; const uint16_t *sel0 = NULL;
  s.dplane = NULL; s.splane = NULL;
#pragma unroll
  for (int f = 0; f < GDP_MAXFILLS; f++) {
    if (f < nf && pass >= F[f].pass0 && pass < F[f].pass0 + F[f].npass) {
      const int p = pass - F[f].pass0;
      const int dd = (lane - F[f].lane0) + 32 * p;
      thi = max(thi,F[f].nB);
      tlo = min(tlo,32 * p);
      maxstart = max(maxstart,F[f].band);
      minend = min(minend,min(F[f].nA,F[f].nB));
      lm = (lm < 0 || lm == F[f].lateadd) ? F[f].lateadd : 2;
      prof0 = F[f].prof; sel0 = F[f].sel;
      s.dplane = F[f].dirs + (size_t) p * F[f].dirPW;		/* the pass's planes are shared by its fills */
      s.splane = SCORES ? F[f].sc + (size_t) p * F[f].scPW : NULL;
#if GMAPDP_STAGE
      if (!WIDE) {
	/* staging: this fill takes the pass's next slot; lanes slot * NARR .. slot * NARR + NARR - 1 copy its arrays (bases
	   here, the block offsets below).  Idle lanes read the last slot. */
	using SG = StgGeom<EVAL>;
	const int sl = min(nslots,GDP_PASS_FILLS - 1);
	nslots++;
	if (s.d < 0) s.slot = sl;
	const int ca = lane - sl * SG::NARR;
	if (ca >= 0 && ca < SG::NARR) {
	  cp_arr = ca;
	  s.cp_dst = (uint32_t) sl * SG::SLOT;
	  if (ca == 0) { s.cp_src = reinterpret_cast<const char *>(F[f].prof); s.cp_bytes = STG_PROF_BYTES; s.cp_stride = STG_B * 8; }
	  else if (ca == 1) { s.cp_src = reinterpret_cast<const char *>(F[f].sel); s.cp_bytes = STG_SEL_BYTES; s.cp_stride = STG_B * 2; s.cp_dst += STG_PROF_BYTES; }
	  else if (EVAL && ca == 2) {
	    s.cp_src = reinterpret_cast<const char *>(F[f].right ? gc->entR : gc->entL); s.cp_bytes = STG_OWN_BYTES; s.cp_stride = STG_B * 4;
	    s.cp_dst += STG_PROF_BYTES + STG_SEL_BYTES;
	  } else if (EVAL) {
	    s.cp_src = reinterpret_cast<const char *>(F[f].right ? gc->entL : gc->entR); s.cp_bytes = STG_PAR_BYTES; s.cp_stride = -STG_B * 4;
	    s.cp_dst += STG_PROF_BYTES + STG_SEL_BYTES + STG_OWN_BYTES;
	  }
	}
      }
#endif
      if (lane >= F[f].lane0 && dd <= F[f].band && (F[f].npass > 1 || lane - F[f].lane0 <= F[f].band)) {
	s.d = dd; nA = F[f].nA; nB = F[f].nB; s.lateadd = F[f].lateadd; s.lower = F[f].lower;
#if GMAPDP_STAGE
	s.slot = min(nslots - 1,GDP_PASS_FILLS - 1);
#endif
	prof = F[f].prof; s.code = F[f].code; s.sel = F[f].sel;
	s.edge_in = WIDE && (p > 0 && lane == 0);
	s.edge_out = WIDE && (p + 1 < F[f].npass && lane == 31);
	if (EVAL) {
	  const bool right = F[f].right, lower = F[f].lower;
	  const int rlength = gc->rlength, glen = lower ? F[f].nA : F[f].nB;
	  s.own_di = right ? gc->rdi : gc->ldi; s.oth_di = right ? gc->ldi : gc->rdi; s.oth_dg = right ? gc->dgL : gc->dgR;
	  /* the reference's scan ranges: not the main diagonal, not the top edge of the band, inside the offset
	     limit (col + partner col < lim: constant along a diagonal), rows 1 .. rlength-1, columns 1 .. glength-2 */
	  s.ev_ok = dd >= 1 && (lower || dd <= F[f].band - 1) && ((lower ? rlength - dd : rlength + dd) < gc->lim);
	  s.ev_lo = dd + 1; s.ev_hi = lower ? rlength - 1 : min(rlength - 1 + dd,glen - 2);
	  const int seg = right ? (lower ? 1 : 2) : (lower ? 3 : 4);
	  /* key(t) = rL * 32768 + seg * 4096 + col, with row = lower ? t : t - dd, col = lower ? t - dd : t */
	  const int row0 = lower ? 0 : -dd, col0 = lower ? -dd : 0;
	  s.key0 = (right ? rlength - row0 : row0) * 32768 + seg * 4096 + col0;
	  s.kstep = right ? 1 - 32768 : 1 + 32768;
	}
      }
    }
  }
  if (thi < 0) return;
  s.dplane += lane + (size_t) (tlo >> 4) * 32;
  if (SCORES) s.splane += lane + (size_t) (tlo >> 1) * 32;
  s.tstart = (s.d >= 0) ? s.d : 0x7fffffff; s.tend = (s.d >= 0) ? min(nA + s.d,nB) : -1;
  s.isd0 = (s.d <= 0);
  s.pp = prof - s.d;
  if (s.d < 0) {					/* idle lanes of the interior steps read (and discard) a valid diagonal */
    s.pp = prof0; s.sel = sel0;
    if (EVAL) { s.own_di = gc->ldi; s.oth_di = gc->rdi; s.oth_dg = gc->dgR; s.lower = true; s.d = -1; }
  }
  s.trk = TRACK && s.d >= 0 && (!s.lower || s.d > 0) && !lastrow;
  const uint32_t negpair = ((uint32_t) NEG & 0xffffu) | ((uint32_t) NEG << 16);

  s.Hprev = (s.d == 0) ? 0 : NEG; s.H = NEG;
  s.pk_out = 0; s.dacc = 0; s.sacc = 0;

  /* interior steps [F0, F1): F0 a multiple of 16 past every lane's start, F1 before any lane's last step */
  int F0 = (maxstart + 1 + 15) & ~15, F1 = F0;
  if (EVAL) minend -= 1;				/* keeps the interior steps inside columns <= glength - 2 */
  if (!WIDE && fastok) while (F1 + 16 <= minend) F1 += 16;
  const bool hasfast = (F1 > F0);
#if GMAPDP_STAGE
  if (!WIDE && hasfast) {
    /* the copies of block 0 = [F0, F0 + STG_B): see the windows at StgGeom */
    using SG = StgGeom<EVAL>;
    s.ncopies = min(nslots,GDP_PASS_FILLS) * SG::NARR;
    s.cp_total = (uint32_t) min(nslots,GDP_PASS_FILLS) * SG::SLOT;
    if (cp_arr == 0) s.cp_src += (ptrdiff_t) (F0 - 32) * 8;
    else if (cp_arr == 1) s.cp_src += (ptrdiff_t) F0 * 2;
    else if (cp_arr == 2) s.cp_src += (ptrdiff_t) (F0 - 32) * 4;
    else if (cp_arr == 3) s.cp_src += (ptrdiff_t) (((gc->rlength + 1) & ~3) - F0 - STG_B) * 4;
  }
#endif

  int t = tlo;
  for (int phase = 0; phase < 2; phase++) {
    const int tlast = (phase == 0 && hasfast) ? F0 - 1 : thi;
    bool act_n = (t >= s.tstart && t <= s.tend);
    uint2 p_n = make_uint2(0u,0u); int cd_n = 0;
    if (act_n) { p_n = s.pp[t]; cd_n = s.code[t]; }
    for (; t <= tlast; t += 2) {
#pragma unroll
      for (int u = 0; u < 2; u++) {
	const int tc = t + u;
	uint32_t pk_in = __shfl_up_sync(FULLMASK,s.pk_out,1);
	const bool act = act_n;
	const uint2 p = p_n; const int cd = cd_n;
	act_n = (tc + 1 >= s.tstart) && (tc + 1 <= s.tend);
	if (act_n) { p_n = s.pp[tc + 1]; cd_n = s.code[tc + 1]; }
	uint32_t bits = 0;
	if (act) {
	  if (WIDE && s.edge_in) pk_in = edge[tc - s.d];
	  if (s.d == 0) pk_in = negpair;
	  const int sc = max(prof_pick(p.x,p.y,cd & 15),prof_pick(p.x,p.y,cd >> 4));
	  const int Hd = clampi(s.Hprev + sc,NEG,POS);
	  const int Hl = (int) (short) (pk_in & 0xffffu), El = ((int) pk_in) >> 16;
	  const int T1 = max(Hl + open,NEG);
	  const bool dE = (El + s.lateadd > T1);
	  const int E = max(max(El,T1) + extend,NEG);
	  const bool dN = (E + s.lateadd > Hd);
	  s.H = max(Hd,E);
	  bits = (s.d == 0) ? 0u : ((dN ? 1u : 0u) | (dE ? 2u : 0u));
	  s.Hprev = s.H;
	  s.pk_out = ((uint32_t) s.H & 0xffffu) | ((uint32_t) E << 16);
	  if (WIDE && s.edge_out) edge[tc - s.d] = s.pk_out;
	  if (TRACK) {
	    const int i = tc - s.d;
	    const int r = s.lower ? tc : i, c = s.lower ? i : tc;
	    if (r >= 1 && c >= 1 && (!s.lower || s.d > 0) && (!lastrow || r == track_rlen)) {
	      const int key = (r << 16) | c;
	      if (s.H > bt->bs || (s.H == bt->bs && (s.lateadd ? key > bt->bk : key < bt->bk))) { bt->bs = s.H; bt->bk = key; }
	    }
	  }
	  if (EVAL) {
	    if (s.ev_ok && tc >= s.ev_lo && tc <= s.ev_hi) {
	      const int col = s.lower ? tc - s.d : tc, rP = gc->rlength - (s.lower ? tc : tc - s.d);
	      const int tot = s.H + gen_points(*gc,(int) s.own_di[col] & (int) s.oth_di[rP]) + (int) s.oth_dg[rP];
	      const int kt = s.key0 + tc * s.kstep;
	      if (tot > gb->s) { gb->s = tot; gb->key = kt; gb->p = -1.0; gb->cnt = 0; }
	      else if (tot == gb->s) {
		if (ev_now) gen_tie(*gc,*gb,kt);
		else { if (gb->cnt < GEN_TIECAP) gb->ties[gb->cnt * 32] = (uint32_t) kt; gb->cnt++; }
	      }
	    }
	  }
	}
	s.dacc |= bits << (2 * (tc & 15));
	if (SCORES) s.sacc |= ((uint32_t) s.H & 0xffffu) << (16 * u);
      }
      if (SCORES) { *s.splane = s.sacc; s.splane += 32; s.sacc = 0; }
      if ((t & 15) == 14) { *s.dplane = s.dacc; s.dplane += 32; s.dacc = 0; }
    }
    if (!WIDE && phase == 0 && hasfast) {
      /* end gaps have one tie rule per box: two lean copies; the two-sided modes mix the rules: one copy, rule per lane */
      if (MODE == 0 && lm == 0) tri_fast<MODE,0>(s,t,F1,open,extend,NEG,POS,negpair,bt,gc,gb);
      else if (MODE == 0 && lm == 1) tri_fast<MODE,1>(s,t,F1,open,extend,NEG,POS,negpair,bt,gc,gb);
      else tri_fast<MODE,2>(s,t,F1,open,extend,NEG,POS,negpair,bt,gc,gb);
    } else break;
  }
  if (((thi & ~1) & 15) != 14) *s.dplane = s.dacc;
  __syncwarp();
}

/* out of line for the two-sided kernels (called from several places: one copy of the code) */
template <int MODE, bool WIDE>
__device__ __noinline__ void tri_pass_call (const TriFill (&F)[GDP_MAXFILLS], int nf, int pass, int open, int extend, int NEG, int POS,
			  BestTrack *bt, int track_rlen, bool lastrow, uint32_t *edge, bool fastok, const GenCtx *gc, GenBest *gb,
			  bool ev_now, uint32_t stg) {
  tri_pass_body<MODE,WIDE>(F,nf,pass,open,extend,NEG,POS,bt,track_rlen,lastrow,edge,fastok,gc,gb,ev_now,stg);
}
#ifndef GMAPDP_TRI_INLINE
#define GMAPDP_TRI_INLINE 3	/* bit MODE set: that mode's passes are inlined into its kernel.  Measured (500 k boxes): inlining
				   helps the end and cdna kernels (8.7 -> 8.2 ms, 12.2 -> 11.5 ms) and hurts the genome kernel (30.0 ->
				   33-35 ms at 4-6 blocks/SM: its pass is the biggest and spills once inlined) */
#endif
template <int MODE, bool WIDE>
__device__ __forceinline__ void tri_pass (const TriFill (&F)[GDP_MAXFILLS], int nf, int pass, int open, int extend, int NEG, int POS,
			  BestTrack *bt, int track_rlen, bool lastrow, uint32_t *edge, bool fastok, const GenCtx *gc, GenBest *gb,
			  bool ev_now, uint32_t stg) {
  if ((GMAPDP_TRI_INLINE >> MODE) & 1) tri_pass_body<MODE,WIDE>(F,nf,pass,open,extend,NEG,POS,bt,track_rlen,lastrow,edge,fastok,gc,gb,ev_now,stg);
  else tri_pass_call<MODE,WIDE>(F,nf,pass,open,extend,NEG,POS,bt,track_rlen,lastrow,edge,fastok,gc,gb,ev_now,stg);
}

/* all passes of a box's E-only fills */
#ifndef GMAPDP_PASS_SYNC
#define GMAPDP_PASS_SYNC 0	/* 1: boxes in step (bsync) also meet in front of each of their first GDP_MAXFILLS passes */
#endif
template <int MODE>
__device__ void tri_fill_all (const TriFill (&F)[GDP_MAXFILLS], int nf, int npasses, int open, int extend, int NEG, int POS,
			      BestTrack *bt, int track_rlen, bool lastrow, uint32_t *edge, bool fastok, const GenCtx *gc, GenBest *gb,
			      uint32_t stg, bool ev_now = false, bool bsync = false) {
  bool wide = false;
#pragma unroll
  for (int f = 0; f < GDP_MAXFILLS; f++) if (f < nf && F[f].npass > 1) wide = true;
#if GMAPDP_PASS_SYNC
  if (bsync) for (int p = npasses; p < GDP_MAXFILLS; p++) __syncthreads();	/* every box: exactly GDP_MAXFILLS pass barriers */
#endif
  for (int pass = 0; pass < npasses; pass++) {
#if GMAPDP_PASS_SYNC
    if (bsync && pass > 0 && pass < GDP_MAXFILLS) __syncthreads();
#endif
    if (wide) tri_pass<MODE,true>(F,nf,pass,open,extend,NEG,POS,bt,track_rlen,lastrow,edge,false,gc,gb,ev_now,stg);
    else tri_pass<MODE,false>(F,nf,pass,open,extend,NEG,POS,bt,track_rlen,lastrow,edge,fastok && !ev_now,gc,gb,ev_now,stg);
  }
}

/* Full fill: Dynprog_simd_8 / _16, stripe-faithful.
 *
 * Per-warp shared memory holds one 8-byte boundary entry per genome column c:
 *   .x = (H of the previous stripe's last row at column c, 16 bits) | (class code of column c << 16),  .y = FF[c]
 * (FF = the reference's int carry of the vertical gap between stripes, dynprog_simd.c:3422-3432,3471).
 * Lane 0 reads its column's entry, lane 31 rewrites it for the next stripe; the class code travels down
 * the lanes packed into the same shuffle as H, so only lanes 0 and 31 touch memory in the inner loop.
 *
 * Each stripe's steps are split into edge steps (some lane is outside the band, on its edge, on column 0
 * or past the last row: the literal restatement of :3257-3478) and interior steps (every lane strictly
 * inside the band), which need none of the edge logic. */
struct FullLane { int E, Hl, diag, cg_out; uint32_t pk_out; int bprevH; };

/* The word that travels down the lanes (and sits in the boundary entries): H in the HIGH half, the column's
   class field in the low half.  Packing is one PRMT (of H and the incoming word: the field is never extracted --
   the score's PRMT reads only the low 16 bits of its selector) and H comes back with one arithmetic shift. */
__device__ __forceinline__ uint32_t full_pack (int H, uint32_t fieldword) {	/* (H << 16) | (fieldword & 0xffff) */
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(fieldword), "r"(H));
  return d;
}
__device__ __forceinline__ int full_H (uint32_t pk) { return ((int) pk) >> 16; }

/* The 16-bit "field" that travels with H: without an alt genome it is the ready-made PRMT selector of
   the column's class (k*0x1111+0x8880), with one it is the class pair k | kalt<<4. */
template <bool ALT>
__device__ __forceinline__ int field_score (uint32_t plo, uint32_t p4, uint32_t field) {
  if (ALT) return max(prof_pick(plo,p4,(int) (field & 15u)),prof_pick(plo,p4,(int) ((field >> 4) & 15u)));
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(plo), "r"(p4), "r"(field));
  return (int) d;
}

/* edge step: the literal restatement, any lane may be inactive / on a band edge / on column 0 */
template <bool LATE, bool ALT>
__device__ __forceinline__ uint32_t full_edge (FullLane &st, const int lane, const int c, const int r, const bool rowact,
					       const int rlo, const int c0, const int chigh, const int lband, const int uband,
					       const int open, const int extend, const int NEG, const int POS,
					       const uint32_t plo, const uint32_t p4, uint2 *bnd) {
  int cg_in = __shfl_up_sync(FULLMASK,st.cg_out,1);
  const uint32_t pk_in = __shfl_up_sync(FULLMASK,st.pk_out,1);
  const bool act = rowact && c >= c0 && c <= chigh;
  uint32_t nib = 0;
  if (act) {
    int last_in = full_H(pk_in);			/* final H(r-1,c) = the F pass's last_nogap */
    uint32_t field = pk_in & 0xffffu;
    int Hs = st.diag;
    if (lane == 0) {
      const uint2 e = bnd[c];
      const int bH = full_H(e.x);
      field = e.x & 0xffffu;
      if (rlo == 0) {
	Hs = (c == 0) ? 0 : NEG;
	cg_in = NEG32; last_in = NEG32;
      } else {
	Hs = (c == 0) ? NEG : st.bprevH;
	if (c < rlo + uband) { cg_in = (int) e.y; last_in = bH; }
	else { cg_in = NEG32; last_in = NEG32; }
      }
      st.bprevH = bH;
    }
    /* E (horizontal gap), dynprog_simd.c:3318-3334 */
    const int T1 = max(st.Hl + open,NEG);
    bool dE = LATE ? (st.E >= T1) : (st.E > T1);
    st.E = max(max(st.E,T1) + extend,NEG);
    /* H, :3350-3389 */
    int sc;
    if (c == 0) sc = (r == 0) ? 0 : NEG;
    else sc = field_score<ALT>(plo,p4,field);
    const int Hd = clampi(Hs + sc,NEG,POS);
    uint32_t dN = (LATE ? (st.E >= Hd) : (st.E > Hd)) ? 1u : 0u;
    int H = max(Hd,st.E);
    bool dF = false;
    if (r >= c - uband && r <= c + lband) {
      if (r == c + lband && c > 0) { H = Hd; dE = false; dN = 0; }	/* bottom of band forced DIAG, :3395-3417 */
      if (r == c - uband) {						/* top of band, :3434-3446 */
	st.cg_out = NEG32 + open + extend;
      } else {							/* F (vertical gap), :3449-3469 */
	const int score = last_in + open;
	dF = LATE ? (cg_in >= score) : (cg_in > score);
	const int cg = (dF ? cg_in : score) + extend;
	const bool tV = LATE ? (cg >= H) : (cg > H);
	H = tV ? max(cg,NEG) : H; dN = tV ? 2u : dN;
	st.cg_out = cg;
      }
    }
    st.diag = last_in;
    st.Hl = H;
    st.pk_out = full_pack(H,field);
    if (lane == 31) bnd[c] = make_uint2(st.pk_out,(uint32_t) st.cg_out);
    nib = dN | (dE ? 4u : 0u) | (dF ? 8u : 0u);
  }
  return nib;
}

/* acc += (a >= b or a > b) ? c : 0 as a compare (ALU pipe) and a predicated multiply-add (FMA pipe): the
   full-fill kernel saturates the ALU pipe, so the direction bits are accumulated on the other one.  The bits of
   one word are distinct, so adding is or-ing. */
/* `one' is the value 1 read from the kernel arguments: a multiplier the assembler cannot fold, which keeps the
   operation a real IMAD (with a literal 1 it is turned back into an ALU add). */
template <bool GE, uint32_t C>
__device__ __forceinline__ void add_if (uint32_t &acc, int a, int b, uint32_t one) {
  if (GE) asm("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %1, %2;\n\t@q mad.lo.u32 %0, %3, %4, %0;\n\t}" : "+r"(acc) : "r"(a), "r"(b), "r"(one), "n"(C));
  else asm("{\n\t.reg .pred q;\n\tsetp.gt.s32 q, %1, %2;\n\t@q mad.lo.u32 %0, %3, %4, %0;\n\t}" : "+r"(acc) : "r"(a), "r"(b), "r"(one), "n"(C));
}

/* interior step: every lane is active and strictly inside the band, c >= 1.  Branch-free: all lanes
   read their column's boundary entry, lane 0 selects it; FIRST = the stripe starts at row 0. */
/* Per-lane multipliers that turn selects and the repacking of the travelling word into multiply-adds (FMA pipe):
   m0 = 1 on lane 0 else 0, m1 = 1 - m0, k64 = 65536, nk64 = -65536 -- all derived from the kernel argument `one',
   so the assembler has to keep them as IMADs (each ALU-pipe instruction moved over shortens the critical pipe). */
struct FullMul { uint32_t m0, m1, k64, nk64; };

template <bool LATE, bool ALT, bool FIRST, int U>
__device__ __forceinline__ void full_fast (FullLane &st, uint32_t &a8, const uint32_t one, const FullMul &fm, const bool isl0, const bool isl31, const int open, const int extend,
					       const int NEG, const int POS, const uint32_t plo, const uint32_t p4, uint2 *bp) {
  const int cg_sh = __shfl_up_sync(FULLMASK,st.cg_out,1);
  const uint32_t pk_sh = __shfl_up_sync(FULLMASK,st.pk_out,1);
  int cg_in, last_in, Hs;
  uint32_t field;
  if (FIRST) {
    const uint32_t ex = bp->x;
    field = isl0 ? ex : pk_sh;				/* low half; the high half is ignored downstream */
    Hs = isl0 ? NEG : st.diag;
    cg_in = isl0 ? NEG32 : cg_sh;
    last_in = isl0 ? NEG32 : full_H(pk_sh);
  } else {
    const uint2 e = *bp;
    const uint32_t src = e.x * fm.m0 + pk_sh * fm.m1;	/* lane 0 takes the previous stripe's last row */
    field = src;					/* low half; the high half is ignored downstream */
    Hs = st.diag;					/* lane 0: the boundary H of the previous column = its previous last_in */
    cg_in = (int) (e.y * fm.m0 + (uint32_t) cg_sh * fm.m1);
    last_in = full_H(src);
  }
  /* direction bits of this step's nibble: bit 0 E beats the diagonal, bit 1 F beats both (a set bit 1 means VERT
     whatever bit 0 says: tb_full), bit 2 Egap, bit 3 Fgap */
  const int T1 = max(st.Hl + open,NEG);
  add_if<LATE,(4u << (4 * U))>(a8,st.E,T1,one);
  st.E = max(max(st.E,T1) + extend,NEG);
  const int Hd = clampi(Hs + field_score<ALT>(plo,p4,field),NEG,POS);
  add_if<LATE,(1u << (4 * U))>(a8,st.E,Hd,one);
  int H = max(Hd,st.E);
  const int score = last_in + open;
  add_if<LATE,(8u << (4 * U))>(a8,cg_in,score,one);
  const int cg = max(cg_in,score) + extend;
  add_if<LATE,(2u << (4 * U))>(a8,cg,H,one);
  H = max(H,cg);
  st.cg_out = cg;
  st.diag = last_in;
  st.Hl = H;
  /* same field, H instead of last_in in the high half: src + (H - last_in) * 65536 */
  if (FIRST) st.pk_out = full_pack(H,field);
  else st.pk_out = (uint32_t) H * fm.k64 + ((uint32_t) last_in * fm.nk64 + field);
  if (isl31) *bp = make_uint2(st.pk_out,(uint32_t) cg);
}

/* general step, branch-free: any lane may be inactive (before its first / after its last column, past the
   last row), outside the band or on one of its edges -- everything full_edge handles except column 0.  Same
   arithmetic as full_edge, with the cases turned into predicates: dd = c - r locates the lane in the band
   (bottom edge dd == -lband, top edge dd == uband), `first' = the stripe starts at row 0, cF = rlo + uband =
   the first column whose vertical-gap carry from the previous stripe is out of reach. */
template <bool LATE, bool ALT>
__device__ __forceinline__ uint32_t full_gen (FullLane &st, const bool isl0, const bool isl31, const int c, const int dd, const bool act,
					      const bool first, const int cF, const int lband, const int uband,
					      const int open, const int extend, const int NEG, const int POS,
					      const uint32_t plo, const uint32_t p4, uint2 *bnd) {
  const int cg_sh = __shfl_up_sync(FULLMASK,st.cg_out,1);
  const uint32_t pk_sh = __shfl_up_sync(FULLMASK,st.pk_out,1);
  uint2 e = make_uint2(0u,0u);
  if (isl0 && act) e = bnd[c];
  const int bH = full_H(e.x);
  const bool carry = !first && (c < cF);
  const uint32_t field = (isl0 ? e.x : pk_sh) & 0xffffu;
  const int Hs = isl0 ? (first ? NEG : st.bprevH) : st.diag;
  const int cg_in = isl0 ? (carry ? (int) e.y : NEG32) : cg_sh;
  const int last_in = isl0 ? (carry ? bH : NEG32) : full_H(pk_sh);
  /* E (horizontal gap), H */
  const int T1 = max(st.Hl + open,NEG);
  const bool dEr = LATE ? (st.E >= T1) : (st.E > T1);
  const int En = max(max(st.E,T1) + extend,NEG);
  const int Hd = clampi(Hs + field_score<ALT>(plo,p4,field),NEG,POS);
  const bool dNr = LATE ? (En >= Hd) : (En > Hd);
  const bool bottom = (dd == -lband), top = (dd == uband);
  const bool doF = (dd < uband) && (dd >= -lband);		/* inside the band and not on its top edge */
  int H = bottom ? Hd : max(Hd,En);
  const bool dE = dEr && !bottom, dNh = dNr && !bottom;
  /* F (vertical gap) */
  const int score = last_in + open;
  const bool dFr = LATE ? (cg_in >= score) : (cg_in > score);
  const int cg = (dFr ? cg_in : score) + extend;
  const bool tV = doF && (LATE ? (cg >= H) : (cg > H));
  const bool dF = doF && dFr;
  H = tV ? max(cg,NEG) : H;
  const int cgo = top ? (NEG32 + open + extend) : (doF ? cg : st.cg_out);
  const uint32_t pk = full_pack(H,field);
  if (act) { st.E = En; st.cg_out = cgo; st.diag = last_in; st.Hl = H; st.pk_out = pk; }
  if (isl0 && act) st.bprevH = bH;
  if (isl31 && act) bnd[c] = make_uint2(pk,(uint32_t) cgo);
  const uint32_t nib = (tV ? 2u : (dNh ? 1u : 0u)) | (dE ? 4u : 0u) | (dF ? 8u : 0u);
  return act ? nib : 0u;
}

template <bool LATE, bool ALT>
__device__ void fill_full (const SideSeq &sd, int lband, int uband, int mt, int open, int extend,
			   int NEG, int POS, uint32_t *dirs, const FGeom &fg, uint2 *bnd, const GdpTables *tb, const uint32_t one) {
  const int lane = threadIdx.x & 31;
  const bool isl0 = (lane == 0), isl31 = (lane == 31);
  const int rlen = sd.rlen, glen = sd.glen;
  FullMul fm;
  fm.m0 = isl0 ? one : 0u; fm.m1 = one - fm.m0; fm.k64 = one << 16; fm.nk64 = 0u - fm.k64;
  /* the columns' class fields into the boundary entries */
  for (int c = lane; c <= glen; c += 32) {
    uint32_t field;
    if (c == 0) field = ALT ? 0x44u : (4u * 0x1111u + 0x8880u);
    else if (ALT) field = (uint32_t) (nt_class(sd.g(c)) | (nt_class(sd.ga(c)) << 4));
    else field = (uint32_t) nt_class(sd.g(c)) * 0x1111u + 0x8880u;
    bnd[c] = make_uint2(field,0u);
  }
  __syncwarp();
  int s = 0;
  for (int rlo = 0; rlo <= rlen; rlo += 32, s++) {
    const int rhigh = min(rlo + 31,rlen);
    const int r = rlo + lane;
    const bool rowact = (r <= rlen);
    const int c0 = max(0,rlo - lband);
    const int chigh = min(rhigh + uband,glen);
    if (c0 > chigh) continue;
    const int nsteps = (chigh - c0 + 1) + 31;
    /* interior steps: every lane has 1 <= c, c0 <= c <= chigh and r - lband < c < r + uband */
    int fs = nsteps, fe = -1;
    if (rhigh == rlo + 31) {
      fs = max(max(c0,1),rlo + 31 - lband + 1) - c0 + 31;
      fe = min(chigh,rlo + uband - 1) - c0;
      if (fs > fe) { fs = nsteps; fe = -1; }
    }

    uint32_t plo = 0, p4 = 0;
    if (rowact) {
      const int q = (r == 0) ? 'N' : sd.q(r);
      const uint2 p = *reinterpret_cast<const uint2 *>(&tb->U[mt][q & 127][0]);
      plo = p.x; p4 = p.y;
    }
    FullLane st;
    st.E = LATE ? NEG : NEG + 1;
    st.Hl = NEG - open; st.diag = NEG - open;
    st.cg_out = NEG32; st.pk_out = 0;
    st.bprevH = (lane == 0 && rlo > 0 && c0 > 0) ? full_H(bnd[c0 - 1].x) : NEG;
    uint32_t acc = 0;
    uint32_t *dst = dirs + (size_t) s * fg.dirW + lane;
    int tt = 0, c = c0 - lane;
#define EDGE_STEP() do { \
      const uint32_t nib = full_edge<LATE,ALT>(st,lane,c,r,rowact,rlo,c0,chigh,lband,uband,open,extend,NEG,POS,plo,p4,bnd); \
      acc |= nib << (4 * (tt & 7)); \
      if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; } tt++; c++; } while (0)
#define FAST_STEP(FIRSTFLAG) do { \
      uint32_t nib = 0; full_fast<LATE,ALT,FIRSTFLAG,0>(st,nib,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bnd + c); \
      acc |= nib << (4 * (tt & 7)); \
      if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; } tt++; c++; } while (0)
#define FAST_CHUNKS(FIRSTFLAG) do { \
      while (tt <= fe && (tt & 7)) FAST_STEP(FIRSTFLAG); \
      uint2 *bp = bnd + c; uint32_t *dp = dst + (tt >> 3) * 32; \
      while (tt + 7 <= fe) { \
	uint32_t a8 = 0; \
	full_fast<LATE,ALT,FIRSTFLAG,0>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 0); \
	full_fast<LATE,ALT,FIRSTFLAG,1>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 1); \
	full_fast<LATE,ALT,FIRSTFLAG,2>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 2); \
	full_fast<LATE,ALT,FIRSTFLAG,3>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 3); \
	full_fast<LATE,ALT,FIRSTFLAG,4>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 4); \
	full_fast<LATE,ALT,FIRSTFLAG,5>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 5); \
	full_fast<LATE,ALT,FIRSTFLAG,6>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 6); \
	full_fast<LATE,ALT,FIRSTFLAG,7>(st,a8,one,fm,isl0,isl31,open,extend,NEG,POS,plo,p4,bp + 7); \
	*dp = a8; dp += 32; bp += 8; tt += 8; c += 8; \
      } \
      while (tt <= fe) FAST_STEP(FIRSTFLAG); } while (0)
    /* a lane is active on the steps tt = lane .. lane + (chigh - c0) of its stripe (and never past the last row):
       one subtraction and one unsigned compare per step */
    const int act_first = rowact ? lane : 0x40000000;
    const unsigned act_span = (unsigned) (chigh - c0);
#define GEN_ONE(U) full_gen<LATE,ALT>(st,isl0,isl31,c + (U),c + (U) - r,(unsigned) (tt + (U) - act_first) <= act_span,rlo == 0,rlo + uband, \
				     lband,uband,open,extend,NEG,POS,plo,p4,bnd)
#define GEN_STEPS(END) do { \
      while (tt < (END) && (tt & 7)) { acc |= GEN_ONE(0) << (4 * (tt & 7)); if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; } tt++; c++; } \
      while (tt + 8 <= (END)) { \
	uint32_t a8 = 0; \
	_Pragma("unroll") for (int u = 0; u < 8; u++) a8 |= GEN_ONE(u) << (4 * u); \
	dst[(tt >> 3) * 32] = a8; tt += 8; c += 8; \
      } \
      while (tt < (END)) { acc |= GEN_ONE(0) << (4 * (tt & 7)); if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; } tt++; c++; } } while (0)
    /* column 0 needs the literal step: it is reached only by the first 32 steps of a stripe with c0 == 0 */
#ifndef GMAPDP_GEN_MODE
#define GMAPDP_GEN_MODE 1
#endif
#if GMAPDP_GEN_MODE == 0
    const int e1 = min(fs,nsteps);
    while (tt < e1) EDGE_STEP();
    { const int tt0_ = tt; if (rlo == 0) FAST_CHUNKS(true); else FAST_CHUNKS(false); if (tt > tt0_) st.bprevH = st.diag;	/* the interior steps keep lane 0's previous boundary H in diag */ }
    while (tt < nsteps) EDGE_STEP();
#elif GMAPDP_GEN_MODE == 1
    const int e0 = min((c0 >= 1) ? 0 : 32,nsteps);
    while (tt < e0) EDGE_STEP();
    const int e1 = min(fs,nsteps);
    for (int seg = 0; seg < 2; seg++) {		/* one copy of the general step: before and after the interior steps */
      const int gend = seg ? nsteps : e1;
      _Pragma("unroll 1")
      while (tt < gend) { acc |= GEN_ONE(0) << (4 * (tt & 7)); if ((tt & 7) == 7) { dst[(tt >> 3) * 32] = acc; acc = 0; } tt++; c++; }
      if (seg == 0) { { const int tt0_ = tt; if (rlo == 0) FAST_CHUNKS(true); else FAST_CHUNKS(false); if (tt > tt0_) st.bprevH = st.diag;	/* the interior steps keep lane 0's previous boundary H in diag */ } }
    }
#else
    const int e0 = min((c0 >= 1) ? 0 : 32,nsteps);
    while (tt < e0) EDGE_STEP();
    const int e1 = min(fs,nsteps);
    GEN_STEPS(e1);
    { const int tt0_ = tt; if (rlo == 0) FAST_CHUNKS(true); else FAST_CHUNKS(false); if (tt > tt0_) st.bprevH = st.diag;	/* the interior steps keep lane 0's previous boundary H in diag */ }
    GEN_STEPS(nsteps);
#endif
#undef GEN_ONE
#undef GEN_STEPS
#undef EDGE_STEP
#undef FAST_STEP
#undef FAST_CHUNKS
    if (nsteps & 7) dst[(nsteps >> 3) * 32] = acc;
    __syncwarp();
  }
}

/* ------------------------------------------------------------------------------------------------
 * fill_full_pk -- Dynprog_simd_16 (dynprog_simd.c:6562) on packed s16x2 arithmetic: TWO 32-row stripes per warp.
 *
 * Lane l holds rows rlo + l (low halfword, "A") and rlo + 32 + l (high halfword, "B") of a 64-row stripe pair; the
 * pair is one 64-stage systolic array: at step t half A of lane l is on column c0 + t - l, half B on the column 32 to
 * the left.  The travelling words (final H and the vertical-gap carry) ROTATE down the lanes: lane 0's B half takes
 * what lane 31's A half produced one step earlier, its A half the entry the previous pair's last row left in shared
 * memory.  One VIMNMX.S16x2 gives the maximum of both halves AND both comparison predicates (the direction bits:
 * predicated IMADs on the FMA pipe, as in full_fast), VIADDMNMX.S16x2 adds and floors both halves: 16 ALU-pipe
 * instructions per 64 cells where full_fast needs 19 per 32.
 *
 * Why the halves may floor at PK_FLOOR instead of replaying the reference's saturation at -32768.  The 16-bit fills
 * are exact per cell, and a traceback only ever visits cells whose score derives from H(0,0):
 * such a score is at least -(5 min(r,c) + |open| + 3 |r - c|) >= -20 012 for sides up to GDP_PK_MAXSIDE (the path
 * along the main diagonal and then along the row or column stays inside the band), whereas a value that derives from
 * a minus-infinity start (cells outside the band, row -1, column -1) is at most floor + 3 min(r,c): -32 768 + 7 500
 * in the reference, PK_FLOOR + 7 500 here.  Both stay below every real score, so every comparison a traceback can
 * read has the same outcome; no halfword ever wraps because every sum is floored before it is used again.  The band
 * edges then need no forced directions: left of its band a row is held at the floor (its first cell sees E = floor,
 * i.e. the reference's forced DIAG at the bottom of the band, :6980-7002); on the top edge of the band the inputs
 * from the row above are floored (no vertical gap into the top cell, :7019-7031); past the band a row computes
 * values nobody reads.  Column 0 scores 0 against every row (H(0,0) = 0 comes from the initial diagonal value,
 * H(r,0) from the vertical gap).
 *
 * Shared memory, 8 bytes per column (the same array as fill_full's, GDP_PK_PAD entries before column 0):
 *   .x = H of the previous pair's last row | its vertical-gap carry << 16      (lane 31 rewrites it, 63 columns behind lane 0)
 *   .y = PRMT selector of the column's class for half A | selector of column c - 32 for half B << 16
 * ---------------------------------------------------------------------------------------------- */
#define PK_FLOOR (-30000)
#define PK_F2 0x8AD08AD0u	/* PK_FLOOR in both halves */

__device__ __forceinline__ uint32_t pk_prmt (uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

struct PkLane {
  uint32_t E, H, diag, cg;		/* s16x2: A | B << 16 */
};
struct PkConst {
  uint32_t open2, ext2, selH, selC, ploA, p4A, ploB, p4B, one;
  int tA0, tB0, tAtop, tBtop;
  bool isl31;
};

/* max of both halves; the bits C_LO / C_HI are added to accA / accB where (LATE ? a >= b : a > b) holds in the low / high
   half.  One VIMNMX.S16x2 with both predicate outputs (the PTX below is the pattern ptxas fuses into it, as in
   __vibmax_s16x2: the predicate is "the maximum equals the first operand") and two predicated IMADs on the FMA pipe
   (`one' is the kernel argument 1, see add_if). */
template <bool LATE, uint32_t C_LO, uint32_t C_HI>
__device__ __forceinline__ uint32_t pk_max_bits (uint32_t a, uint32_t b, uint32_t &accA, uint32_t &accB, const uint32_t one) {
  uint32_t m;
  if (LATE) {
    asm("{\n\t.reg .pred pu, pv;\n\t.reg .s16 rs0, rs1, rs2, rs3;\n\t"
	"max.s16x2 %0, %3, %4;\n\t"
	"mov.b32 {rs0, rs1}, %0;\n\t"
	"mov.b32 {rs2, rs3}, %3;\n\t"
	"setp.eq.s16 pv, rs0, rs2;\n\t"
	"setp.eq.s16 pu, rs1, rs3;\n\t"
	"@pv mad.lo.u32 %1, %5, %6, %1;\n\t"
	"@pu mad.lo.u32 %2, %5, %7, %2;\n\t}"
	: "=r"(m), "+r"(accA), "+r"(accB) : "r"(a), "r"(b), "r"(one), "n"(C_LO), "n"(C_HI));
  } else {
    /* a > b  <=>  not (b >= a) */
    asm("{\n\t.reg .pred pu, pv;\n\t.reg .s16 rs0, rs1, rs2, rs3;\n\t"
	"max.s16x2 %0, %4, %3;\n\t"
	"mov.b32 {rs0, rs1}, %0;\n\t"
	"mov.b32 {rs2, rs3}, %4;\n\t"
	"setp.eq.s16 pv, rs0, rs2;\n\t"
	"setp.eq.s16 pu, rs1, rs3;\n\t"
	"@!pv mad.lo.u32 %1, %5, %6, %1;\n\t"
	"@!pu mad.lo.u32 %2, %5, %7, %2;\n\t}"
	: "=r"(m), "+r"(accA), "+r"(accB) : "r"(a), "r"(b), "r"(one), "n"(C_LO), "n"(C_HI));
  }
  return m;
}

/* one step of the pair; GEN: lanes may be left of their band (held at the floor) or on its top edge */
template <bool LATE, bool GEN, int U>
__device__ __forceinline__ void pk_step (PkLane &st, uint32_t &accA, uint32_t &accB, const PkConst &k, const int src, const int t, uint2 *sp) {
  const uint32_t hS = __shfl_sync(FULLMASK,st.H,src);
  const uint32_t cgS = __shfl_sync(FULLMASK,st.cg,src);
  const uint2 e = sp[U];
  uint32_t hIn = pk_prmt(e.x,hS,k.selH);		/* lane 0: boundary H | lane 31's A half << 16; others: the lane above */
  uint32_t cgIn = pk_prmt(e.x,cgS,k.selC);
  if (GEN) {
    const uint32_t mTop = ((t + U == k.tAtop) ? 0xffffu : 0u) | ((t + U == k.tBtop) ? 0xffff0000u : 0u);
    hIn = (hIn & ~mTop) | (PK_F2 & mTop);
    cgIn = (cgIn & ~mTop) | (PK_F2 & mTop);
  }
  const uint32_t sc = pk_prmt(k.ploA,k.p4A,e.y) | pk_prmt(k.ploB,k.p4B,e.y >> 16);
  constexpr uint32_t S = 4u * U;
  /* E (horizontal gap) */
  const uint32_t T1 = __viaddmax_s16x2(st.H,k.open2,PK_F2);
  const uint32_t mE = pk_max_bits<LATE,(4u << S),(4u << S)>(st.E,T1,accA,accB,k.one);
  uint32_t E = __viaddmax_s16x2(mE,k.ext2,PK_F2);
  /* diagonal */
  const uint32_t Hd = __viaddmax_s16x2(st.diag,sc,PK_F2);
  const uint32_t H1 = pk_max_bits<LATE,(1u << S),(1u << S)>(E,Hd,accA,accB,k.one);
  /* F (vertical gap) */
  const uint32_t score = __viaddmax_s16x2(hIn,k.open2,PK_F2);
  const uint32_t cgm = pk_max_bits<LATE,(8u << S),(8u << S)>(cgIn,score,accA,accB,k.one);
  uint32_t cg = __viaddmax_s16x2(cgm,k.ext2,PK_F2);
  uint32_t H = pk_max_bits<LATE,(2u << S),(2u << S)>(cg,H1,accA,accB,k.one);
  if (GEN) {
    const uint32_t mPre = ((t + U < k.tA0) ? 0xffffu : 0u) | ((t + U < k.tB0) ? 0xffff0000u : 0u);
    E = (E & ~mPre) | (PK_F2 & mPre);
    H = (H & ~mPre) | (PK_F2 & mPre);
    cg = (cg & ~mPre) | (PK_F2 & mPre);
  }
  st.diag = hIn; st.E = E; st.H = H; st.cg = cg;
  if (k.isl31) sp[U - 32].x = pk_prmt(H,cg,0x7632u);		/* the B halves: H | carry << 16 */
}

template <bool LATE, bool GEN>
__device__ __forceinline__ void pk_chunk (PkLane &st, const PkConst &k, const int src, const int t, uint2 *sp, uint32_t *dp) {
  uint32_t accA = 0, accB = 0;
  pk_step<LATE,GEN,0>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,1>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,2>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,3>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,4>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,5>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,6>(st,accA,accB,k,src,t,sp);
  pk_step<LATE,GEN,7>(st,accA,accB,k,src,t,sp);
  dp[0] = accA; dp[32] = accB;
}

template <bool LATE>
__device__ void fill_full_pk (const SideSeq &sd, const int lband, const int uband, const int mt, const int open, const int extend,
			      uint32_t *dirs, const FGeom &fg, uint2 *bnd, const GdpTables *tb, const uint32_t one) {
  const int lane = threadIdx.x & 31;
  const int rlen = sd.rlen, glen = sd.glen;
  uint2 *col = bnd + GDP_PK_PAD;			/* col[c], c = -GDP_PK_PAD .. glen + 64 + 15 */
  for (int c = lane - GDP_PK_PAD; c < glen + 64 + 16; c += 32) {
    const int cb = c - 32;
    const uint32_t ka = (c >= 1 && c <= glen) ? (uint32_t) nt_class(sd.g(c)) : 5u;		/* 5: a zero byte of the profile */
    const uint32_t kb = (cb >= 1 && cb <= glen) ? (uint32_t) nt_class(sd.g(cb)) : 5u;
    const uint32_t selA = (ka == 5u) ? 0x5555u : (0x5500u | ((ka | 8u) << 4) | ka);
    const uint32_t selB = (kb == 5u) ? 0x5555u : (0x0055u | (kb << 8) | ((kb | 8u) << 12));
    col[c] = make_uint2(PK_F2,selA | (selB << 16));
  }
  __syncwarp();
  PkConst k;
  k.one = one;
  k.open2 = ((uint32_t) open & 0xffffu) * 0x10001u; k.ext2 = ((uint32_t) extend & 0xffffu) * 0x10001u;
  k.selH = (lane == 0) ? 0x5410u : 0x7654u;		/* lane 0: e.x low half | shuffled low half << 16 */
  k.selC = (lane == 0) ? 0x5432u : 0x7654u;		/* lane 0: e.x high half | shuffled low half << 16 */
  k.isl31 = (lane == 31);
  const int src = (lane + 31) & 31;
  for (int p = 0, rlo = 0; rlo <= rlen; rlo += 64, p++) {
    const int v = min(63,rlen - rlo);			/* last row of the pair that exists */
    const int c0 = max(0,rlo - lband);
    const int chigh = min(rlo + v + uband,glen);
    if (c0 > chigh) continue;
    const int nsteps = chigh - c0 + v + 1;
    const int rA = rlo + lane, rB = rA + 32;
    {
      const int qa = (rA == 0) ? 'N' : ((rA <= rlen) ? sd.q(rA) : 0);
      const int qb = (rB <= rlen) ? sd.q(rB) : 0;
      const uint2 pa = *reinterpret_cast<const uint2 *>(&tb->U[mt][qa & 127][0]);
      const uint2 pb = *reinterpret_cast<const uint2 *>(&tb->U[mt][qb & 127][0]);
      k.ploA = (rA <= rlen) ? pa.x : 0u; k.p4A = (rA <= rlen) ? (pa.y & 0xffu) : 0u;
      k.ploB = (rB <= rlen) ? pb.x : 0u; k.p4B = (rB <= rlen) ? (pb.y & 0xffu) : 0u;
    }
    /* first step on which a half is inside its band, and its step on the top edge of the band */
    k.tA0 = lane + max(0,rA - lband) - c0;
    k.tB0 = lane + 32 + max(0,rB - lband) - c0;
    k.tAtop = rA + uband - c0 + lane;
    k.tBtop = rB + uband - c0 + lane + 32;
    PkLane st;
    st.E = PK_F2; st.H = PK_F2; st.cg = PK_F2; st.diag = PK_F2;
    if (lane == 0) {
      /* H(rlo - 1, c0 - 1): 0 stands in for it at the origin, so that H(0,0) = 0 */
      const uint32_t d = (p == 0) ? 0u : ((c0 > 0) ? (col[c0 - 1].x & 0xffffu) : (PK_F2 & 0xffffu));
      st.diag = (PK_F2 & 0xffff0000u) | d;
    }
    /* interior steps: every half inside its band and below its top edge */
    const int fs = (63 + max(0,rlo + 63 - lband) - c0 + 7) & ~7;
    const int fe = (rlo + uband - c0) & ~7;					/* exclusive */
    const int nsteps8 = (nsteps + 7) & ~7;
    uint2 *sp = col + (c0 - lane);
    uint32_t *dp = dirs + (size_t) p * fg.dirW2 + lane;
    int t = 0;
    const int g1 = (fs < fe) ? min(fs,nsteps8) : nsteps8;
#pragma unroll 1
    for (int seg = 0; seg < 2; seg++) {		/* one copy of the general chunk: before and after the interior chunks */
      const int gend = seg ? nsteps8 : g1;
#pragma unroll 1
      for (; t < gend; t += 8, sp += 8, dp += 64) pk_chunk<LATE,true>(st,k,src,t,sp,dp);
      if (seg == 0) {
	const int fend = min(fe,nsteps8);
#pragma unroll 1
	for (; t < fend; t += 8, sp += 8, dp += 64) pk_chunk<LATE,false>(st,k,src,t,sp,dp);
      }
    }
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t full_dir (const uint32_t *dirs, const FGeom &fg, int r, int c) {
  if (r < c - fg.uband || r > c + fg.lband) return 0;
  if (fg.pk) {
    const int p = r >> 6, hb = (r >> 5) & 1, l = r & 31;
    const int c0 = max(0,(p << 6) - fg.lband);
    const int t = c - c0 + l + 32 * hb;
    const uint32_t w = dirs[(size_t) p * fg.dirW2 + (t >> 3) * 64 + hb * 32 + l];
    return (w >> (4 * (t & 7))) & 15u;
  }
  const int s = r >> 5, l = r & 31;
  const int c0 = max(0,(s << 5) - fg.lband);
  const int tt = c - c0 + l;
  const uint32_t w = dirs[(size_t) s * fg.dirW + (tt >> 3) * 32 + l];
  return (w >> (4 * (tt & 7))) & 15u;
}

/* ------------------------------------------------------------------------------------------------
 * Tracebacks, warp-cooperative.  The reference walks one cell at a time (dynprog_simd.c:9154-9946);
 * here the 32 lanes look at 32 consecutive cells of the current diagonal at once: a ballot finds
 * the first non-DIAG cell, the diagonal run before it is emitted and scored in one go (matches and
 * mismatches by popcount), then the gap run is measured with a second ballot along the row / column,
 * replaying the reference's post-decrement loop conditions (:9175,9342,9459) exactly.
 * All lanes hold the same (r, c) and counters; lane 0 writes the edit script.
 * ---------------------------------------------------------------------------------------------- */
struct TbAcc {
  int score, nmatches, nmismatches, nopens, nindels;
  uint32_t *ops; int nops; int pend;
  __device__ __forceinline__ void flush () {
    if (pend) { if ((threadIdx.x & 31) == 0) ops[nops] = ((uint32_t) pend << 2); nops++; pend = 0; }
  }
  __device__ __forceinline__ void gap (int kind, int dist, bool scored) {
    flush();
    if ((threadIdx.x & 31) == 0) ops[nops] = ((uint32_t) dist << 2) | (uint32_t) kind;
    nops++;
    if (scored) { score += -3 - dist; nopens += 1; nindels += dist; }
  }
};

/* scores the n leading cells (r-k, c-k), k < n, of the diagonal; lane k looks at cell k */
__device__ __forceinline__ void tb_diag_run (TbAcc &a, const SideSeq &sd, int r, int c, int n, const GdpTables *tb) {
  const int k = threadIdx.x & 31;
  bool isM = false, isX = false;
  if (k < n) {
    const int c1 = sd.q(r - k) & 127, c2 = sd.g(c - k) & 127, c2a = sd.ga(c - k) & 127;
    if (c2 != '*') {
      isM = (c1 == c2 || c1 == c2a || ((tb->cons[c1][c2 >> 5] >> (c2 & 31)) & 1u) || ((tb->cons[c1][c2a >> 5] >> (c2a & 31)) & 1u));
      isX = !isM;
    }
  }
  const int nm = __popc(__ballot_sync(FULLMASK,isM)), nx = __popc(__ballot_sync(FULLMASK,isX));
  a.score += nm - 3 * nx; a.nmatches += nm; a.nmismatches += nx;
  a.pend += n;
}

__device__ __forceinline__ int first_set (uint32_t m) { return m ? (__ffs(m) - 1) : 32; }

__device__ __noinline__ void tb_upper (TbAcc &a, const SideSeq &sd, const TriFill &pl, int r, int c, const GdpTables *tb) {
  const int k = threadIdx.x & 31;
  while (r > 0 && c > 0) {
    const bool valid = (r - k > 0) && (c - k > 0);
    const bool stop = !valid || (tri_dir(pl,r - k,c - k) & 1u);
    const uint32_t m = __ballot_sync(FULLMASK,stop);
    const int n = first_set(m);
    if (n > 0) { tb_diag_run(a,sd,r,c,n,tb); r -= n; c -= n; }
    if (n < 32 && r > 0 && c > 0) {
      /* horizontal gap: dist = 1; while (dirE[c--][r] != DIAG) dist++; */
      int dist = 1;
      for (;;) {
	const bool z = (c - k < 0) || !(tri_dir(pl,r,c - k) & 2u);
	const int kz = first_set(__ballot_sync(FULLMASK,z));
	if (kz < 32) { dist += kz; c -= kz + 1; break; }
	dist += 32; c -= 32;
      }
      a.gap(1,dist,dist < 9);
    }
  }
  a.flush();
  if (c > 0 && c < 9) { a.score += -3 - c; a.nopens += 1; a.nindels += c; }
}

__device__ __noinline__ void tb_lower (TbAcc &a, const SideSeq &sd, const TriFill &pl, int r, int c, const GdpTables *tb) {
  const int k = threadIdx.x & 31;
  while (r > 0 && c > 0) {
    const bool valid = (r - k > 0) && (c - k > 0);
    const bool stop = !valid || (tri_dir(pl,c - k,r - k) & 1u);
    const uint32_t m = __ballot_sync(FULLMASK,stop);
    const int n = first_set(m);
    if (n > 0) { tb_diag_run(a,sd,r,c,n,tb); r -= n; c -= n; }
    if (n < 32 && r > 0 && c > 0) {
      int dist = 1;
      for (;;) {
	const bool z = (r - k < 0) || !(tri_dir(pl,c,r - k) & 2u);
	const int kz = first_set(__ballot_sync(FULLMASK,z));
	if (kz < 32) { dist += kz; r -= kz + 1; break; }
	dist += 32; r -= 32;
      }
      a.gap(2,dist,true);
    }
  }
  a.flush();
  if (r > 0) { a.score += -3 - r; a.nopens += 1; a.nindels += r; }
}

__device__ void tb_full (TbAcc &a, const SideSeq &sd, const uint32_t *dirs, const FGeom &fg, int r, int c, const GdpTables *tb) {
  const int k = threadIdx.x & 31;
  while (r > 0 && c > 0) {
    const bool valid = (r - k > 0) && (c - k > 0);
    const uint32_t nib = valid ? full_dir(dirs,fg,r - k,c - k) : 0u;
    const bool stop = !valid || (nib & 3u);
    const uint32_t m = __ballot_sync(FULLMASK,stop);
    const int n = first_set(m);
    if (n > 0) { tb_diag_run(a,sd,r,c,n,tb); r -= n; c -= n; }
    if (n < 32 && r > 0 && c > 0) {
      const uint32_t dir = __shfl_sync(FULLMASK,nib,n) & 3u;
      int dist = 1;
      if (dir == 1u) {
	/* dist = 1; while (c > 0 && dirE[c--][r] != DIAG) dist++; */
	for (;;) {
	  const bool cond = (c - k > 0);
	  const bool z = !cond || !(full_dir(dirs,fg,r,c - k) & 4u);
	  const uint32_t zm = __ballot_sync(FULLMASK,z);
	  const int kz = first_set(zm);
	  if (kz < 32) { const bool condz = (c - kz > 0); dist += kz; c -= condz ? kz + 1 : kz; break; }
	  dist += 32; c -= 32;
	}
	a.gap(1,dist,dist < 9);
      } else {
	for (;;) {
	  const bool cond = (r - k > 0);
	  const bool z = !cond || !(full_dir(dirs,fg,r - k,c) & 8u);
	  const uint32_t zm = __ballot_sync(FULLMASK,z);
	  const int kz = first_set(zm);
	  if (kz < 32) { const bool condz = (r - kz > 0); dist += kz; r -= condz ? kz + 1 : kz; break; }
	  dist += 32; r -= 32;
	}
	a.gap(2,dist,true);
      }
    }
  }
  a.flush();
  if (r == 0 && c == 0) {
  } else if (c == 0) { a.score += -3 - r; a.nopens += 1; a.nindels += r; }
  else if (c < 9) { a.score += -3 - c; a.nopens += 1; a.nindels += c; }
}

/* ------------------------------------------------------------------------------------------------
 * Bridges
 * ---------------------------------------------------------------------------------------------- */
__device__ __forceinline__ int intron_points (const int *isc, int ldi, int rdi) {
  const int t = ldi & rdi;
  return t ? isc[31 - __clz(t)] : 0;
}

/* Main-diagonal scores of the two upper fills (what the d == 0 lane of tri_pass computes: H(i,i) =
   clamp(H(i-1,i-1) + score(i,i)), no gap can reach the diagonal).  A clamped add x -> min(max(x + a, l), u) is
   closed under composition -- (a1,l1,u1) then (a2,l2,u2) = (a1 + a2, clamp(l1 + a2), clamp(u1 + a2)) with clamp
   to [l2,u2] -- so the recurrence is a warp prefix scan over 32 positions at a time. */
__device__ void gen_diagonals (const TriFill &LU, const TriFill &RU, short *dgL, short *dgR, int NEG, int POS) {
  const int lane = threadIdx.x & 31;
  for (int side = 0; side < 2; side++) {
    const TriFill &f = side ? RU : LU;
    short *dg = side ? dgR : dgL;
    const int n = min(f.nA,f.nB);
    int carry = 0;
    for (int base = 0; base <= n; base += 32) {
      const int i = base + lane;
      int a = 0, lo = -(1 << 30), hi = (1 << 30);		/* identity past the end */
      if (i <= n) {
	const uint2 p = f.prof[i]; const int cd = f.code[i];
	a = max(prof_pick(p.x,p.y,cd & 15),prof_pick(p.x,p.y,cd >> 4)); lo = NEG; hi = POS;
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
	const int a1 = __shfl_up_sync(FULLMASK,a,off), l1 = __shfl_up_sync(FULLMASK,lo,off), u1 = __shfl_up_sync(FULLMASK,hi,off);
	if (lane >= off) {
	  const int nl = min(max(l1 + a,lo),hi), nu = min(max(u1 + a,lo),hi);
	  a += a1; lo = nl; hi = nu;
	}
      }
      const int H = min(max(carry + a,lo),hi);
      if (i <= n) dg[i] = (short) H;
      carry = __shfl_sync(FULLMASK,H,31);
    }
  }
  __syncwarp();
}

/* bridge_intron_gap_{8,16}_site_level, dynprog_genome.c:866-1386, after the fills: the main-diagonal candidates
   (cL = rL, cR = rR) and the best "with dinucleotide" candidate, the combination of the lanes' bests, and the
   reference's choice between the two.  Returns finalscore. */
__device__ int bridge_genome_finish (const gmapdp_box &b, const GenCtx &g, GenBest gb, int NEG, const int *isc,
				     int *bestrL, int *bestrR, int *bestcL, int *bestcR) {
  const int lane = threadIdx.x & 31;
  const int rlength = g.rlength;
  /* the keys that tied with this lane's best inside the fills */
  {
    const int n = min(gb.cnt,GEN_TIECAP);
    for (int j = 0; j < n; j++) gen_tie(g,gb,(int) gb.ties[j * 32]);
    gb.cnt = 0;
  }
  int dns = NEG, dnkey = -1; double dnp = 0.0;		/* best diagonal candidate with a dinucleotide pair: by probability sum */
  for (int rL = 1 + lane; rL < rlength; rL += 32) {
    const int rR = rlength - rL;
    const int scoreI = gen_points(g,(int) g.ldi[rL] & (int) g.rdi[rR]);
    const int score = (int) g.dgL[rL] + scoreI + (int) g.dgR[rR];
    const int key = rL << 15;
    if (score > gb.s) { gb.s = score; gb.key = key; gb.p = -1.0; }
    else if (score == gb.s) gen_tie(g,gb,key);
    if (scoreI > 0) {
      const double ps = g.lp[rL] + g.rp[rR];
      if (ps > dnp) { dns = score; dnp = ps; dnkey = key; }
    }
  }
  if (gb.key >= 0 && gb.p < 0.0) gb.p = gen_key_prob(g,gb.key);
  if (gb.key < 0) gb.p = 0.0;
  /* combine: max score, then max probability, then earliest in the reference's scan order */
  for (int off = 16; off > 0; off >>= 1) {
    const int os = __shfl_xor_sync(FULLMASK,gb.s,off), ok = __shfl_xor_sync(FULLMASK,gb.key,off);
    const double op = __shfl_xor_sync(FULLMASK,gb.p,off);
    if (os > gb.s || (os == gb.s && (op > gb.p || (op == gb.p && ok < gb.key)))) { gb.s = os; gb.key = ok; gb.p = op; }
    const int ds = __shfl_xor_sync(FULLMASK,dns,off), dk = __shfl_xor_sync(FULLMASK,dnkey,off);
    const double dp = __shfl_xor_sync(FULLMASK,dnp,off);
    if (dp > dnp || (dp == dnp && dp > 0.0 && dk < dnkey)) { dns = ds; dnkey = dk; dnp = dp; }
  }

  int bestscore = gb.s, key = gb.key;
  bool use_dinucl;
  if (gb.p > 2 * 0.85) use_dinucl = false;
  else if (dnp == 0.0) use_dinucl = false;
  else if (dns < 0 || dns < bestscore - 9) use_dinucl = false;
  else use_dinucl = true;
  if (use_dinucl) { key = dnkey; bestscore = dns; }
  int rL = 0, rR = 0, cL = 0, cR = 0;
  if (key >= 0) {
    const int seg = (key >> 12) & 7, col = key & 4095;
    rL = key >> 15; rR = rlength - rL;
    cL = (seg >= 3) ? col : rL; cR = (seg == 1 || seg == 2) ? col : rR;
  }
  *bestrL = rL; *bestrR = rR; *bestcL = cL; *bestcR = cR;
  if (bestscore < 0) return bestscore;
  if (b.flags & GMAPDP_F_HALFP) return bestscore - intron_points(isc,g.ldi[cL],g.rdi[cR]) / 2;
  return bestscore;
}

/* bridge_cdna_gap_{8,16}_ud, dynprog_cdna.c:123-375.
 *
 * The reference scans (cL asc, cR desc, rL asc, rR asc) keeping the best score with >= (jump late:
 * the LAST maximum in scan order wins) or > (the FIRST wins), subject to rR < lim - rL.  The score is
 * scoreL(cL,rL) + scoreR(cR,rR) + pen(cL,cR) with pen = 0 for cR = glength-cL and `open' otherwise, so
 * the scan is an argmax with a positional tie-break and is evaluated from three tables:
 *   pb[cR][k]  best scoreR over rR <= k (and its rR)                       -- handles the rR constraint
 *   M[cR]      = pb[cR][all rR]                                            -- the unconstrained column best
 *   Q[j]       best of M[cR] over cR <= j (and its cR), tie rule on cR     -- the unconstrained cR best
 *   V[rL][j]   best of pb[cR][rcap(rL)] over the first j+1 columns whose rR range is cut by the constraint
 *              rR <= rcap = lim - rL - 1 (they depend on rL only: cR in (rcap - lbandR, rcap + ubandR])
 * For one (cL,rL): the pen = 0 column is looked up in pb, the cut columns below cR0 = glength - cL in V, and all
 * other columns are answered by one Q lookup.  O(g * band) instead of O(g^2 * band^2). */
__device__ int bridge_cdna (const gmapdp_box &b, const TriFill &LU, const TriFill &LL, const TriFill &RU,
			    const TriFill &RL, int NEG, uint32_t *pb, uint32_t *V, uint32_t *smemMQ, int *bestcL, int *bestcR, int *bestrL, int *bestrR) {
  const int lane = threadIdx.x & 31;
  const int glength = b.glenL, rlengthL = b.rlenL, rlengthR = b.rlenR;
  const int lbandL = b.lbandL, ubandL = b.ubandL, lbandR = b.lbandR, ubandR = b.ubandR;
  const int open = b.open, lim = b.offdiff;
  const bool late = (b.flags & GMAPDP_F_BRIDGE_LATE) != 0;
  const int PBW = lbandR + ubandR + 1;
  uint32_t *M = smemMQ, *Q = smemMQ + (glength + 1);	/* packed (score+32768)<<16 | rR  and  (score+32768)<<16 | cR */
  const uint32_t NONE = 0u;				/* score field 0 = "no rR in range" (a real score is >= -32768+... > field 0 only if > -32768) */

  for (int cR = lane; cR <= glength; cR += 32) {
    const int lo = max(cR - ubandR,1), hi = min(cR + lbandR,rlengthR - 1);
    uint32_t *row = pb + (size_t) cR * PBW - (cR - ubandR);
    int bs = 0, br = 0; bool have = false;
    /* four score loads in flight per round (they do not depend on the running best) */
    for (int rR0 = lo; rR0 <= hi; rR0 += 4) {
      int sc4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int rR = min(rR0 + u,hi); sc4[u] = (rR < cR) ? tri_score(RU,rR,cR) : tri_score(RL,cR,rR); }
#pragma unroll
      for (int u = 0; u < 4; u++) {
	const int rR = rR0 + u, sc = sc4[u];
	if (rR <= hi) {
	  if (!have || (late ? sc >= bs : sc > bs)) { bs = sc; br = rR; have = true; }
	  row[rR] = ((uint32_t) (bs + 32768) << 16) | (uint32_t) br;
	}
      }
    }
    M[cR] = have ? (((uint32_t) (bs + 32768) << 16) | (uint32_t) br) : NONE;
  }
  __syncwarp();
  /* Q: prefix best over cR.  The scan visits cR in DESCENDING order, so "late" (last wins) keeps the
     smallest cR among equal scores and "early" the largest. */
  if (lane == 0) {
    int bs = 0, bc = -1;
    for (int cR = 0; cR <= glength; cR++) {
      const uint32_t m = M[cR];
      const bool valid = (m != NONE) || (min(cR + lbandR,rlengthR - 1) >= max(cR - ubandR,1));
      if (valid) {
	const int sc = (int) (m >> 16) - 32768;
	if (bc < 0 || (late ? sc > bs : sc >= bs)) { bs = sc; bc = cR; }
      }
      Q[cR] = (bc < 0) ? 0xffffffffu : (((uint32_t) (bs + 32768) << 16) | (uint32_t) bc);
    }
  }
  __syncwarp();

  /* V: per rL, prefix best over its cut columns (same tie rule on cR as Q) */
  for (int rL = 1 + lane; rL < rlengthL; rL += 32) {
    const int rcap = lim - rL - 1;
    if (rcap >= rlengthR - 1 || rcap < 1) continue;
    const int s0 = max(rcap - lbandR + 1,0), e0 = min(rcap + ubandR,glength);
    uint32_t *row = V + (size_t) rL * PBW;
    int vs = 0, vc = -1;
    /* consecutive columns sit PBW - 1 words apart: four loads in flight per round */
    const uint32_t *src = pb + (size_t) s0 * PBW + (rcap - (s0 - ubandR));
    for (int cR = s0; cR <= e0; cR += 4) {
      uint32_t e4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) e4[u] = src[(size_t) min(u,e0 - cR) * (PBW - 1)];
      src += 4 * (size_t) (PBW - 1);
#pragma unroll
      for (int u = 0; u < 4; u++) if (cR + u <= e0) {
	const int sc = (int) (e4[u] >> 16) - 32768;
	if (vc < 0 || (late ? sc > vs : sc >= vs)) { vs = sc; vc = cR + u; }
	row[cR + u - s0] = ((uint32_t) (vs + 32768) << 16) | (uint32_t) vc;
      }
    }
  }
  __syncwarp();

  /* rR of a candidate never decides a comparison (one candidate per (cL, cR, rL)): the winner's rR is looked up
     at the end from the table its candidate came from (kind 0: pen = 0 column, 1: Q, 2: V) */
  int bs = NEG; unsigned long long bk = 0; bool have = false;
  int bcL = 0, bcR = 0, brL = 0, bkind = 0;
#define CDNA_CAND(OK,CR,SCORE_R,KIND,PEN) do { \
    const int score_ = scoreL + (SCORE_R) + (PEN); \
    const unsigned long long key_ = ((unsigned long long) cL << 48) | ((unsigned long long) (glength - (CR)) << 32) | ((unsigned long long) rL << 16); \
    bool take_; \
    if (!have) take_ = late ? (score_ >= bs) : (score_ > bs); \
    else if (score_ != bs) take_ = score_ > bs; \
    else take_ = late ? (key_ > bk) : (key_ < bk); \
    if ((OK) && take_) { bs = score_; bk = key_; bcL = cL; bcR = (CR); brL = rL; bkind = (KIND); have = true; } } while (0)

  for (int cL = 1 + lane; cL < glength; cL += 32) {
    const int rloL = max(cL - ubandL,1), rhighL = min(cL + lbandL,rlengthL - 1);
    const int cR0 = glength - cL;
    const int lo0 = max(cR0 - ubandR,1), hi0 = min(cR0 + lbandR,rlengthR - 1);
#pragma unroll 2
    for (int rL = rloL; rL <= rhighL; rL++) {
      const int rcap = lim - rL - 1;			/* rR <= rcap */
      /* every lookup of the iteration is issued before any is used (indices clamped into the tables) */
      const int k0 = min(hi0,rcap);
      const bool ok0 = (k0 >= lo0);						/* the pen = 0 column */
      const int junc = min(cR0 - 1,(rcap >= rlengthR - 1) ? cR0 - 1 : rcap - lbandR);	/* columns <= junc are unconstrained */
      const int s0 = max(rcap - lbandR + 1,0), jhi = min(cR0 - 1,rcap + ubandR) - s0;
      const bool okV = (rcap < rlengthR - 1 && rcap >= 1 && jhi >= 0);		/* columns cut by the constraint */
      const int scoreL = (rL < cL) ? tri_score(LU,rL,cL) : tri_score(LL,cL,rL);
      const uint32_t e0 = pb[(size_t) cR0 * PBW + (max(k0,lo0) - (cR0 - ubandR))];
      const uint32_t qe = Q[max(junc,0)];
      const uint32_t v = V[(size_t) rL * PBW + max(jhi,0)];
      CDNA_CAND(ok0,cR0,(int) (e0 >> 16) - 32768,0,0);
      CDNA_CAND(junc >= 0 && qe != 0xffffffffu,(int) (qe & 0xffffu),(int) (qe >> 16) - 32768,1,open);
      CDNA_CAND(okV,(int) (v & 0xffffu),(int) (v >> 16) - 32768,2,open);
    }
  }
#undef CDNA_CAND
  /* lanes partition cL, which leads the scan order: late keeps the largest cL on ties, early the smallest */
  for (int off = 16; off > 0; off >>= 1) {
    const int os = __shfl_xor_sync(FULLMASK,bs,off), ocL = __shfl_xor_sync(FULLMASK,bcL,off), ocR = __shfl_xor_sync(FULLMASK,bcR,off);
    const int orL = __shfl_xor_sync(FULLMASK,brL,off), ok = __shfl_xor_sync(FULLMASK,bkind,off);
    const int oh = __shfl_xor_sync(FULLMASK,(int) have,off);
    bool take = false;
    if (oh) {
      if (!have) take = true;
      else if (os != bs) take = os > bs;
      else take = late ? (ocL > bcL) : (ocL < bcL);
    }
    if (take) { bs = os; bcL = ocL; bcR = ocR; brL = orL; bkind = ok; have = true; }
  }
  int brR = 0;
  if (have) {
    const int rcap = lim - brL - 1;
    if (bkind == 1) brR = (int) (M[bcR] & 0xffffu);
    else {
      const int k = (bkind == 0) ? min(min(bcR + lbandR,rlengthR - 1),rcap) : rcap;
      brR = (int) (pb[(size_t) bcR * PBW + (k - (bcR - ubandR))] & 0xffffu);
    }
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
  uint32_t one;			/* the value 1 (see add_if) */
  const GdpTables *tables;
  GdpGenome genome;		/* resident genome (blocks == NULL: none attached) */
  const double *maxent;		/* packed MaxEnt tables (gmapdp_genome.h) */
  double *devprobs;		/* device-side pool of the MaxEnt arrays of GMAPDP_G_PROBS boxes (batch path) */
};

/* one segment of a resident-genome box into the workspace, lanes over positions (inlined: out of line it cost the
   single-gap and end kernels 1-2 ms per step) */
__device__ __forceinline__ void decode_segment (GdpGenome g, uint32_t p0, int n, bool neg, bool left, uint32_t chroffset, uint32_t chrhigh, uint8_t *seg) {
  const int dir = neg ? -1 : +1;
  const uint32_t lo = left ? chroffset : 0u, hi = left ? 0xffffffffu : chrhigh;
  for (int i = (int) (threadIdx.x & 31); i < n; i += 32) seg[i] = (uint8_t) gdp_segment_char(g,p0,dir,i,lo,hi);
}

/* KIND: 0 single gaps (full fill), 1 end5/end3 (E-only fills + endpoint search), 2 genome gaps, 3 cdna gaps */
template <int KIND, bool INK>
__device__ void process_box (const KernelArgs &ka, int bi, uint32_t *ws, uint2 *bnd, const GdpTables *tb, uint32_t stg, const bool bsync) {
  /* bsync: the warps of the block walk through their boxes in step (BOX_PHASE: a block barrier in front of the fills,
     behind them and in front of the publication) -- see gmapdp_dp_kernel */
#define BOX_PHASE() do { if (bsync) __syncthreads(); } while (0)
  constexpr bool FULLK = (KIND == 0);
  constexpr bool twosided = (KIND >= 2);

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
  uint8_t *qcodeL = bytes; bytes += gdp_align16(b.rlenL + 2);
  uint8_t *qcodeR = bytes; bytes += gdp_align16(b.rlenR + 2);
  uint8_t *gcodeL = bytes; bytes += gdp_align16(b.glenL + 2);
  uint8_t *ldi = bytes; bytes += gdp_align16(b.glenL + 2);
  uint8_t *gcodeR = bytes; bytes += gdp_align16(b.glenR + 2);
  uint8_t *rdi = bytes; bytes += gdp_align16(b.glenR + 2);
  uint16_t *qselL = reinterpret_cast<uint16_t *>(bytes); bytes += gdp_align16(2 * (size_t) (b.rlenL + 2));
  uint16_t *qselR = reinterpret_cast<uint16_t *>(bytes); bytes += gdp_align16(2 * (size_t) (b.rlenR + 2));
  uint16_t *gselL = reinterpret_cast<uint16_t *>(bytes); bytes += gdp_align16(2 * (size_t) (b.glenL + 2));
  uint16_t *gselR = reinterpret_cast<uint16_t *>(bytes); bytes += gdp_align16(2 * (size_t) (b.glenR + 2));
  short *dgL = reinterpret_cast<short *>(bytes); bytes += gdp_align16(2 * (size_t) (b.rlenL + 2));
  short *dgR = reinterpret_cast<short *>(bytes); bytes += gdp_align16(2 * (size_t) (b.rlenR + 2));
  uint32_t *entL = NULL, *entR = NULL;			/* genome gaps: dg << 16 | di per position (staged interior steps) */
  if (KIND == 2) {
    const size_t eb = gdp_align16(4 * (size_t) (GDP_ENT_PAD + gdp_ent_len(b)));
    entL = reinterpret_cast<uint32_t *>(bytes) + GDP_ENT_PAD; bytes += eb;
    entR = reinterpret_cast<uint32_t *>(bytes) + GDP_ENT_PAD; bytes += eb;
  }
  /* segments of resident-genome boxes: decoded once into the workspace (Genome_get_segment_right / _left, genome.c:11023,
     :11079), everything downstream reads them like uploaded ones; genomealt == genome */
  if (b.gflags & GMAPDP_G_SEG_L) {
    uint8_t *seg = bytes; bytes += gdp_align16(b.glenL + 2);
    decode_segment(ka.genome,b.gL_off,b.glenL,(b.gflags & GMAPDP_G_NEG_L) != 0,(b.gflags & GMAPDP_G_LEFT_L) != 0,b.chroffset,b.chrhigh,seg);
    L.G = L.Ga = seg;
    if (!twosided || !(b.gflags & GMAPDP_G_SEG_R)) { R.G = R.Ga = seg; }		/* single / end / cdna: one segment */
  }
  if (twosided && (b.gflags & GMAPDP_G_SEG_R)) {
    uint8_t *seg = bytes; bytes += gdp_align16(b.glenR + 2);
    decode_segment(ka.genome,b.gR_off,b.glenR,(b.gflags & GMAPDP_G_NEG_R) != 0,(b.gflags & GMAPDP_G_LEFT_R) != 0,b.chroffset,b.chrhigh,seg);
    R.G = R.Ga = seg;
  }
  if (b.gflags & (GMAPDP_G_SEG_L | GMAPDP_G_SEG_R)) __syncwarp();
  uint32_t *wp = reinterpret_cast<uint32_t *>(bytes);
  uint32_t *stage = wp; wp += b.rlenL + b.glenL + b.rlenR + b.glenR + 16;


  /* pre-pass: class codes in DP coordinates */
  if (!FULLK) {
    for (int c = lane; c <= b.glenL; c += 32) {
      const int k = (c == 0) ? 4 : nt_class(L.g(c)), ka = (c == 0) ? 4 : nt_class(L.ga(c));
      gcodeL[c] = (uint8_t) (k | (ka << 4)); gselL[c] = (uint16_t) (k * 0x1111 + 0x8880);
    }
    for (int r = lane; r <= b.rlenL; r += 32) {
      const int k = (r == 0) ? 4 : nt_class(L.q(r));
      qcodeL[r] = (uint8_t) (k | (k << 4)); qselL[r] = (uint16_t) (k * 0x1111 + 0x8880);
    }
  }
  if (!FULLK && twosided) {
    for (int c = lane; c <= b.glenR; c += 32) {
      const int k = (c == 0) ? 4 : nt_class(R.g(c)), ka = (c == 0) ? 4 : nt_class(R.ga(c));
      gcodeR[c] = (uint8_t) (k | (ka << 4)); gselR[c] = (uint16_t) (k * 0x1111 + 0x8880);
    }
    for (int r = lane; r <= b.rlenR; r += 32) {
      const int k = (r == 0) ? 4 : nt_class(R.q(r));
      qcodeR[r] = (uint8_t) (k | (k << 4)); qselR[r] = (uint16_t) (k * 0x1111 + 0x8880);
    }
  }
  if (KIND == 2) {
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

  if (FULLK) {
    FGeom fg = fgeom(b.rlenL,b.glenL,b.lbandL,b.ubandL);
#if GMAPDP_BND_GLOBAL
    /* the stripe boundary row lives in the warp's workspace (L1-resident: 8 B per column, read once and written
       once per stripe, addresses known in advance), not in shared memory: the kernel's occupancy is then set by
       its registers alone (4 blocks/SM instead of the 3 that 65 KB of boundary rows per block allowed) */
    if ((reinterpret_cast<uintptr_t>(wp) & 7) != 0) wp++;
    bnd = reinterpret_cast<uint2 *>(wp); wp += 2 * (size_t) (b.glenL + 2 + GDP_PK_EXTRA);
#endif
    uint32_t *dirs = wp;
    const bool alt = (b.gLalt_off != b.gL_off);
    fg.pk = gdp_full_packed(b) ? 1 : 0;
    BOX_PHASE();
    if (fg.pk) {
      if (lateL) fill_full_pk<true>(L,b.lbandL,b.ubandL,mt,open,extend,dirs,fg,bnd,tb,ka.one);
      else fill_full_pk<false>(L,b.lbandL,b.ubandL,mt,open,extend,dirs,fg,bnd,tb,ka.one);
    } else if (lateL) { if (alt) fill_full<true,true>(L,b.lbandL,b.ubandL,mt,open,extend,NEG,POS,dirs,fg,bnd,tb,ka.one);
		 else fill_full<true,false>(L,b.lbandL,b.ubandL,mt,open,extend,NEG,POS,dirs,fg,bnd,tb,ka.one); }
    else { if (alt) fill_full<false,true>(L,b.lbandL,b.ubandL,mt,open,extend,NEG,POS,dirs,fg,bnd,tb,ka.one);
	   else fill_full<false,false>(L,b.lbandL,b.ubandL,mt,open,extend,NEG,POS,dirs,fg,bnd,tb,ka.one); }
    __syncwarp();
    BOX_PHASE();
    tb_full(acc,L,dirs,fg,b.rlenL,b.glenL,tb); lenA = acc.nops;

  } else {
    /* E-only fills on the diagonal lane mapping: L lower, [R lower], L upper, [R upper] */
    TriPacking tp;
    tri_fills_of(b,tp);
    TriFill F[GDP_MAXFILLS];
    const bool noalt = (b.gLalt_off == b.gL_off) && (!twosided || b.gRalt_off == b.gR_off);	/* selectors carry one class */
    while ((reinterpret_cast<uintptr_t>(wp) & 15) != 0) wp++;		/* the profile tables are sources of 16-byte bulk copies */
    constexpr int NF = twosided ? 4 : 2;		/* = tp.nf; constant indices keep tp in registers */
#pragma unroll
    for (int f = 0; f < GDP_MAXFILLS; f++) {
      if (f >= NF) { F[f] = F[0]; continue; }
      const bool lower = twosided ? (f < 2) : (f == 0);
      const bool right = twosided && (f & 1);
      F[f].nA = tp.nA[f]; F[f].nB = tp.nB[f]; F[f].band = tp.band[f];
      F[f].lane0 = tp.lane0[f]; F[f].pass0 = tp.pass0[f]; F[f].npass = tp.npass[f];
      F[f].lateadd = (right ? lateR : lateL) ? 1 : 0;
      F[f].lower = lower; F[f].right = right;
      F[f].code = lower ? (right ? qcodeR : qcodeL) : (right ? gcodeR : gcodeL);
      F[f].sel = lower ? (right ? qselR : qselL) : (right ? gselR : gselL);
      F[f].prof = reinterpret_cast<const uint2 *>(wp) + 32;	/* 32 entries of look-ahead padding in front */
      F[f].dirPW = tp.dirPW; F[f].scPW = tp.scPW;
      tri_profiles(F[f],right ? R : L,mt,use8,reinterpret_cast<uint2 *>(wp) + 32,tb);
      wp += 2 * (size_t) ((tp.nA[f] + 2 + 32 + 1) & ~1);
    }
    uint32_t *dbase = wp; wp += (size_t) tp.npasses * tp.dirPW;
    uint32_t *sbase = NULL;
    if (KIND == 3) { sbase = wp; wp += (size_t) tp.npasses * tp.scPW; }
    uint32_t *edge = wp; wp += tp.maxA + 2;
#pragma unroll
    for (int f = 0; f < NF; f++) {
      F[f].dirs = dbase + (size_t) tp.pass0[f] * tp.dirPW;
      F[f].sc = sbase ? sbase + (size_t) tp.pass0[f] * tp.scPW : NULL;
    }
    __syncwarp();

    if (!twosided) {
      const TriFill &Lo = F[0], &U = F[1];
      const bool lastrow = (b.flags & GMAPDP_F_LASTROW) != 0;
      BestTrack bt;
      if (lastrow) { bt.bs = NEG; bt.bk = (b.rlenL << 16); } else { bt.bs = 0; bt.bk = 0; }
      BOX_PHASE();
      tri_fill_all<0>(F,tp.nf,tp.npasses,open,extend,NEG,POS,&bt,b.rlenL,lastrow,edge,noalt,NULL,NULL,stg,false,bsync);
      BOX_PHASE();
      for (int off = 16; off > 0; off >>= 1) {
	const int os = __shfl_xor_sync(FULLMASK,bt.bs,off), ok = __shfl_xor_sync(FULLMASK,bt.bk,off);
	if (os > bt.bs || (os == bt.bs && (lateL ? ok > bt.bk : ok < bt.bk))) { bt.bs = os; bt.bk = ok; }
      }
      res.finalscore = bt.bs; res.bestrL = bt.bk >> 16; res.bestcL = bt.bk & 0xffff;
      __syncwarp();
      if (!(b.flags & GMAPDP_F_NOTRACE)) {
	if (res.bestcL >= res.bestrL) tb_upper(acc,L,U,res.bestrL,res.bestcL,tb);
	else tb_lower(acc,L,Lo,res.bestrL,res.bestcL,tb);
	lenA = acc.nops;
      }
    } else {
      const TriFill &LL = F[0], &RL = F[1], &LU = F[2], &RU = F[3];
      int brL, brR, bcL, bcR, fs;
      if (KIND == 2) {
	const int di = b.cdna_direction > 0 ? 0 : (b.cdna_direction < 0 ? 1 : 2);
	const int *isc = tb->isc[di][(b.flags & GMAPDP_F_FINALP) ? 1 : 0];
	GenCtx gc;
	gc.ldi = ldi; gc.rdi = rdi; gc.dgL = dgL; gc.dgR = dgR;
	gc.lp = ka.probs + b.probL_off; gc.rp = ka.probs + b.probR_off;
	gc.rlength = b.rlenL; gc.lim = b.offdiff;
	if (b.gflags & GMAPDP_G_PROBS) {
	  if (INK) {
	    /* flights: the two arrays are evaluated here, into the workspace */
	    gc.genome = ka.genome; gc.me = ka.maxent; gc.chroffset = b.chroffset;
	    gc.lpos0 = b.probposL; gc.rpos0 = b.probposR;
	    gc.lstep = (b.gflags & GMAPDP_G_PSTEP_NEG_L) ? -1 : +1; gc.rstep = (b.gflags & GMAPDP_G_PSTEP_NEG_R) ? -1 : +1;
	    gc.lkind = b.probkindL; gc.rkind = b.probkindR; gc.nlp = b.glenL - 1; gc.nrp = b.glenR - 1;
	    if ((reinterpret_cast<uintptr_t>(wp) & 7) != 0) wp++;
	    double *lpb = reinterpret_cast<double *>(wp); wp += 2 * (size_t) (b.glenL + 1);
	    double *rpb = reinterpret_cast<double *>(wp); wp += 2 * (size_t) (b.glenR + 1);
	    for (int c = lane; c <= b.glenL; c += 32) lpb[c] = gen_maxent(gc,0,c);
	    for (int c = lane; c <= b.glenR; c += 32) rpb[c] = gen_maxent(gc,1,c);
	    gc.lp = lpb; gc.rp = rpb;
	    __syncwarp();
	  } else {
	    /* batches: gmapdp_maxent_kernel_batch filled them in the context's device-side pool before this kernel started */
	    gc.lp = ka.devprobs + b.probL_off; gc.rp = ka.devprobs + b.probR_off;
	  }
	}
	gc.it0 = ((uint32_t) isc[0] << 8) | ((uint32_t) isc[1] << 16) | ((uint32_t) isc[2] << 24);
	gc.it1 = (uint32_t) isc[3] | ((uint32_t) isc[4] << 8) | ((uint32_t) isc[5] << 16);
	GenBest gb; gb.s = NEG; gb.key = -1; gb.cnt = 0; gb.p = 0.0; gb.ties = wp + lane; wp += GEN_TIECAP * 32;
	gen_diagonals(LU,RU,dgL,dgR,NEG,POS);
	{
	  /* one word per position for the staged interior steps: the main-diagonal score of the side's upper fill at row i
	     (meaningful up to the shorter side) and the dinucleotide class at column i */
	  const int ne = gdp_ent_len(b), nL = min(b.rlenL,b.glenL), nR = min(b.rlenR,b.glenR);
	  for (int i = lane; i < ne; i += 32) {
	    const uint32_t dl_ = (i <= nL) ? (uint32_t) (uint16_t) dgL[i] : 0u, dr_ = (i <= nR) ? (uint32_t) (uint16_t) dgR[i] : 0u;
	    entL[i] = (dl_ << 16) | ((i <= b.glenL) ? (uint32_t) ldi[i] : 0u);
	    entR[i] = (dr_ << 16) | ((i <= b.glenR) ? (uint32_t) rdi[i] : 0u);
	  }
	  gc.entL = entL; gc.entR = entR;
	  __syncwarp();
	}
	BOX_PHASE();
	/* one call site (the fills are inlined): the second round runs only if a tie list that counts overflowed */
	for (int round = 0; round < 2; round++) {
	  tri_fill_all<2>(F,tp.nf,tp.npasses,open,extend,NEG,POS,NULL,0,false,edge,noalt,&gc,&gb,stg,round == 1,bsync && round == 0);
	  __syncwarp();
	  if (round == 1) break;
	  if (gb.key >= 0) gb.p = -1.0;				/* no probability has been fetched inside the fills */
	  /* only the lanes that reached the warp's best score can win: their tie lists are the ones that count */
	  int smax = gb.s;
	  for (int off = 16; off > 0; off >>= 1) smax = max(smax,__shfl_xor_sync(FULLMASK,smax,off));
	  if (gb.s < smax) gb.cnt = 0;
	  if (!__any_sync(FULLMASK,gb.cnt > GEN_TIECAP)) break;
	  gb.s = NEG; gb.key = -1; gb.cnt = 0; gb.p = 0.0;	/* such a list overflowed: once more, ties resolved on the spot */
	}
	BOX_PHASE();
	fs = bridge_genome_finish(b,gc,gb,NEG,isc,&brL,&brR,&bcL,&bcR);
	if (fs < 0) res.status = 1;
      } else {
	BOX_PHASE();
	tri_fill_all<1>(F,tp.nf,tp.npasses,open,extend,NEG,POS,NULL,0,false,edge,noalt,NULL,NULL,stg,false,bsync);
	__syncwarp();
	BOX_PHASE();
	fs = bridge_cdna(b,LU,LL,RU,RL,NEG,wp,wp + (size_t) (b.glenL + 1) * (b.lbandR + b.ubandR + 1),reinterpret_cast<uint32_t *>(bnd),&bcL,&bcR,&brL,&brR);
      }
      res.finalscore = fs; res.bestrL = brL; res.bestcL = bcL; res.bestrR = brR; res.bestcR = bcR;
      if (res.status == 0) {
	if (bcR >= brR) tb_upper(acc,R,RU,brR,bcR,tb); else tb_lower(acc,R,RL,brR,bcR,tb);
	lenA = acc.nops;
	if (bcL >= brL) tb_upper(acc,L,LU,brL,bcL,tb); else tb_lower(acc,L,LL,brL,bcL,tb);
	lenB = acc.nops - lenA;
      }
    }
  }

  /* publish: allocate script space, copy, write the result */
  __syncwarp();
  BOX_PHASE();
#undef BOX_PHASE
  const int ntot = lenA + lenB;
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

/* Four specialisations of one persistent kernel, one per kind of box: single gaps (full fill: needs the big
   shared-memory boundary rows, 3 blocks/SM), end gaps, genome gaps, cdna gaps (E-only fills, bridges: almost
   no shared memory).  A specialisation carries only its own mode's code, so the warps of an SM share their
   instruction-cache footprint; the four are queued back to back on one stream (launch_chunk). */
#ifndef GMAPDP_BLOCKSYNC
#define GMAPDP_BLOCKSYNC 1
#endif
#ifndef GMAPDP_SYNC_FULL_WORK
#define GMAPDP_SYNC_FULL_WORK 40000	/* single gaps: boxes below this work estimate (box_work: rows x steps per stripe) are processed in step */
#endif
#ifndef GMAPDP_FULL_MINB
#define GMAPDP_FULL_MINB (GMAPDP_BND_GLOBAL ? 4 : 3)
#endif
#ifndef GMAPDP_TRI_MINB
#define GMAPDP_TRI_MINB 5
#endif
#ifndef GMAPDP_END_MINB
#define GMAPDP_END_MINB 8		/* measured: 8 blocks/SM (64 registers) beat 5 for the end-gap kernel */
#endif
#ifndef GMAPDP_GENOME_MINB
#define GMAPDP_GENOME_MINB 8		/* with the boxes of a block in step: 8 blocks/SM (64 registers) 29.7 ms per 100 k genome gaps, 6 blocks 30.9,
					   5 blocks 33.4 (before, on per-warp queues, 6 beat 8); an 8-step interior body (GMAPDP_GENOME_UNR=8) 32.5,
					   passes inlined (GMAPDP_TRI_INLINE=7) 30.9; 256-thread blocks for all kinds 32.3 */
#endif
#ifndef GMAPDP_CDNA_MINB
#define GMAPDP_CDNA_MINB GMAPDP_TRI_MINB
#endif
template <int KIND>
__global__ void __launch_bounds__(BLOCK_THREADS,KIND == 0 ? GMAPDP_FULL_MINB : (KIND == 1 ? GMAPDP_END_MINB : (KIND == 2 ? GMAPDP_GENOME_MINB : GMAPDP_CDNA_MINB)))
gmapdp_dp_kernel (KernelArgs ka) {
  /* shared: one 8-byte boundary entry per column and warp */
  const GdpTables *tb = ka.tables;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2 *bnd = reinterpret_cast<uint2 *>(dyn_smem) + (size_t) warp * ka.smem_cols;
  const int gwarp = blockIdx.x * WARPS_PER_BLOCK + warp;
  uint32_t *ws = ka.ws + (size_t) gwarp * ka.ws_words;
  /* E-only kinds: the warp's staging region behind the boundary rows (StgGeom) */
  uint32_t stg = 0;
  if (KIND != 0 && GMAPDP_STAGE) {
    stg = smem_u32(dyn_smem + (size_t) WARPS_PER_BLOCK * ka.smem_cols * 8) + (uint32_t) warp * StgGeom<KIND == 2>::WARP_BYTES;
    stage_init(stg);
  }

#if GMAPDP_BLOCKSYNC
  /* The block takes WARPS_PER_BLOCK consecutive boxes of the work-sorted queue -- near-identical work -- and its warps
     walk through them in step (process_box: BOX_PHASE).  The kernels are 5 - 13 k instructions; a warp alone runs most
     of them once per box, and the warps of an SM sitting in different phases evict each other's code: on production-size
     boxes instruction fetch was the first stall reason (ncu: no_instruction 5 - 7 per issue).  In step, the four warps of
     a block share what they fetch. */
  /* Large single gaps first, every warp on its own (measured on the benchmark launch: 70.5 ms in step -- a full fill's
     duration depends on more than the work the queue is sorted by, and the block waits for its slowest warp -- against
     66.6 ms), until the warp meets a box below GMAPDP_SYNC_FULL_WORK: the queue is sorted by decreasing work, so from
     there on the boxes are small and the block goes on in step (production-size single gaps: 5.8 -> 4.2 ms per million
     calls).  One loop and ONE call site of process_box for both regimes (two call sites made it a real call, with the
     kernel arguments on the stack). */
  bool solo = (KIND == 0);
  __shared__ int s_idx;
  for (;;) {
    int idx;
    bool step = false, last_solo = false;
    if (solo) {
      idx = 0;
      if (lane == 0) idx = atomicAdd(ka.queue,1);
      idx = __shfl_sync(FULLMASK,idx,0);
      if (idx >= ka.nboxes) { solo = false; continue; }		/* joins the block (which then finds the queue empty) */
      /* the work estimate the queue is sorted by (box_work): monotone along the queue */
      const gmapdp_box &bq = ka.boxes[ka.order[idx]];
      last_solo = ((long long) (bq.rlenL + 32) * (min((int) bq.glenL + 1,32 + bq.lbandL + bq.ubandL) + 31) < GMAPDP_SYNC_FULL_WORK);
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_idx = atomicAdd(ka.queue,WARPS_PER_BLOCK);
      __syncthreads();
      const int idx0 = s_idx;
      if (idx0 >= ka.nboxes) break;
      step = (idx0 + WARPS_PER_BLOCK <= ka.nboxes);		/* a block without a box for every warp runs unsynchronised */
      idx = idx0 + warp;
      if (idx >= ka.nboxes) continue;
    }
    process_box<KIND,false>(ka,ka.order[idx],ws,bnd,tb,stg,step);
    if (last_solo) solo = false;
  }
#else
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(ka.queue,1);
    idx = __shfl_sync(FULLMASK,idx,0);
    if (idx >= ka.nboxes) break;
    process_box<KIND,false>(ka,ka.order[idx],ws,bnd,tb,stg,false);
  }
#endif
}

/* One kernel for boxes of every kind: the streaming runtime's flights (gmapdp_stream.cpp) hold tens to hundreds of
   small boxes, for which the four specialisations above would cost four launches plus the stream joins -- more than the
   boxes themselves.  Same per-box code, chosen per box; the instruction-cache argument for separate kernels does not
   apply to a grid that covers a fraction of the SMs for a hundred microseconds. */
__global__ void __launch_bounds__(BLOCK_THREADS,3)
gmapdp_dp_kernel_any (KernelArgs ka) {
  const GdpTables *tb = ka.tables;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint2 *bnd = reinterpret_cast<uint2 *>(dyn_smem) + (size_t) warp * ka.smem_cols;
  const int gwarp = blockIdx.x * WARPS_PER_BLOCK + warp;
  uint32_t *ws = ka.ws + (size_t) gwarp * ka.ws_words;
  uint32_t stg = 0;
  if (GMAPDP_STAGE) {
    stg = smem_u32(dyn_smem + (size_t) WARPS_PER_BLOCK * ka.smem_cols * 8) + (uint32_t) warp * STG_WARP_BYTES_MAX;
    stage_init(stg);
  }

  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(ka.queue,1);
    idx = __shfl_sync(FULLMASK,idx,0);
    if (idx >= ka.nboxes) break;
    const int bi = ka.order[idx];
    const int mode = ka.boxes[bi].mode;
    if (mode == GMAPDP_SINGLE) process_box<0,true>(ka,bi,ws,bnd,tb,stg,false);
    else if (mode == GMAPDP_GENOME) process_box<2,true>(ka,bi,ws,bnd,tb,stg,false);
    else if (mode == GMAPDP_CDNA) process_box<3,true>(ka,bi,ws,bnd,tb,stg,false);
    else process_box<1,true>(ka,bi,ws,bnd,tb,stg,false);
  }
}

/* MaxEnt arrays of the genome-gap boxes of a batch (resident genome): one launch in front of the genome kernel evaluates
   every entry of both arrays of every box into the device-side pool (`probL_off / probR_off', glength + 1 doubles per
   array), which the genome kernel then reads like uploaded arrays.  One warp walks one array, 32 consecutive positions
   per step and two steps in flight: the genome words are coalesced, the 2 (donor kinds) or 6 (acceptor kinds) table
   entries of a position are independent loads from the 1.5 MB of model tables, which stay in L2 / L1 (nothing else
   competes for them in this kernel: the same look-ups inside the genome kernel were what slowed it down, and a first
   version that ran one launch per large table -- the table in shared memory, the running odds read and rewritten per
   pass -- took 13.4 ms for the benchmark launch's 8 * 10^8 entries).  Entry c of an array is the probability at coordinate
   pos0 + step * c for c < glength - 1 and 0 beyond (the calloc'ed tail of dynprog_genome.c:970-1061), gdp_maxent_prob
   (gmapdp_genome.h): the arithmetic of Maxent_hr_*_prob in the reference's order. */
#define ME_THREADS 256
__global__ void __launch_bounds__(ME_THREADS,6)
gmapdp_maxent_kernel_batch (const gmapdp_box *boxes, const int *order, int nboxes, GdpGenome g, const double *me, double *pool) {
  const int lane = threadIdx.x & 31, nwarps = ME_THREADS / 32;
  const int gw = blockIdx.x * nwarps + (threadIdx.x >> 5), stride = gridDim.x * nwarps;
  for (int a = gw; a < 2 * nboxes; a += stride) {		/* array a: side a & 1 of box order[a >> 1] */
    const gmapdp_box &b = boxes[order[a >> 1]];
    if (!(b.gflags & GMAPDP_G_PROBS)) continue;
    const bool right = (a & 1) != 0;
    const int kind = (int) (right ? b.probkindR : b.probkindL);
    const int glen = right ? b.glenR : b.glenL;
    const uint32_t pos0 = right ? b.probposR : b.probposL;
    const int step = (b.gflags & (right ? GMAPDP_G_PSTEP_NEG_R : GMAPDP_G_PSTEP_NEG_L)) ? -1 : +1;
    double * __restrict__ out = pool + (right ? b.probR_off : b.probL_off);
    const uint32_t chroffset = b.chroffset;
    for (int c0 = lane; c0 <= glen; c0 += 64) {
      double pr[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
	const int c = c0 + 32 * u;
	pr[u] = (c < glen - 1) ? gdp_maxent_prob(kind,g,me,pos0 + (uint32_t) (step * c),chroffset) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 2; u++) if (c0 + 32 * u <= glen) out[c0 + 32 * u] = pr[u];
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Host side: context, memory, launches (the C ABI of include/gmapdp_b200.h)
 * ---------------------------------------------------------------------------------------------- */
#include "gmapdp_internal.h"

#define GDP_NK GDP_NKINDS
static inline int kind_of (int mode) { return mode == GMAPDP_SINGLE ? 0 : (mode == GMAPDP_GENOME ? 2 : (mode == GMAPDP_CDNA ? 3 : 1)); }

struct gmapdp_genome {
  int device;
  uint32_t *d_blocks; size_t nwords;
  double *d_maxent;
};

struct gmapdp_ctx {
  int device, sm_count, grid, max_smem;
  int kgrid[GDP_NK], ksmem_cols[GDP_NK]; size_t kws_words[GDP_NK];	/* per kernel kind: 0 single, 1 end, 2 genome, 3 cdna */
  uint32_t *d_kws[GDP_NK]; size_t cap_kws[GDP_NK];
  cudaStream_t kstream[GDP_NK];		/* kstream[0] == stream */
  cudaEvent_t evj[GDP_NK], evk[GDP_NK][2]; float last_ms[GDP_NK];
  size_t chunk_bytes;			/* pipelining granularity of gmapdp_run_batch (GMAPDP_CHUNK_MB, default 384) */
  std::vector<int> chunk_count;		/* [chunk][kind] boxes */
  cudaStream_t stream, copy_stream;
  std::vector<cudaEvent_t> chunk_events, chunk_done;
  const gmapdp_genome *genome = NULL;		/* attached resident genome + MaxEnt tables */
  double *d_devprobs = NULL; size_t cap_devprobs = 0;	/* MaxEnt arrays evaluated on the device (GMAPDP_G_PROBS boxes) */
  int me_kinds = 0;				/* bit k: some box of the planned batch has a probability array of kind k */
  cudaStream_t d2h_stream = 0;
  unsigned long long *h_cursors = NULL; size_t cap_cursors = 0;		/* pinned: the script cursor after each chunk */
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
  std::vector<void *> grave;		/* outgrown device buffers (see grow) */
  int any_occ; bool dev_set;		/* flights: resident blocks per SM of the all-kinds kernel; device already current in the launcher thread */
  /* the chaining engine (gmapchain_kernels.cu) hangs its state here */
  void *chain; void (*chain_free) (void *);
};

GdpCtxView gmapdp_ctx_view (gmapdp_ctx *ctx) {
  GdpCtxView v;
  v.device = ctx->device; v.sm_count = ctx->sm_count; v.err = &ctx->err; v.launches = &ctx->launches;
  v.chain = &ctx->chain; v.chain_free = &ctx->chain_free; v.grave = &ctx->grave;
  return v;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return GMAPDP_ERR_CUDA; } } while (0)

template <typename T>
static int grow (gmapdp_ctx *ctx, T **p, size_t *cap, size_t need) {
  if (need <= *cap && *p) return GMAPDP_OK;
  /* the outgrown buffer is kept until the context dies: cudaFree waits for every stream of the device -- other lanes'
     flights, the chaining engine's kernels -- and holds the driver's lock while it does; buffers grow geometrically,
     so what is parked stays below the final size */
  if (*p) ctx->grave.push_back((void *) *p);
  *p = NULL;
  size_t n = need + need / 2 + 64;
  CK(cudaMalloc((void **) p,n * sizeof(T)));
  *cap = n;
  return GMAPDP_OK;
}

extern "C" int gmapdp_device_count (void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return ndev;
}

extern "C" int gmapdp_create (gmapdp_ctx **out, int device) {
  gmapdp_ctx *ctx = new gmapdp_ctx();
  *out = ctx;
  ctx->device = device; ctx->launches = 0; ctx->nboxes = 0; ctx->chain = NULL; ctx->chain_free = NULL; ctx->any_occ = 0; ctx->dev_set = false;
  ctx->d_tables = NULL; ctx->d_boxes = NULL; ctx->cap_boxes = 0; ctx->d_order = NULL; ctx->cap_order = 0; ctx->cap_results = 0; ctx->d_seq = NULL; ctx->cap_seq = 0;
  ctx->d_probs = NULL; ctx->cap_probs = 0; ctx->d_results = NULL; ctx->d_script = NULL; ctx->cap_script = 0;
  ctx->d_cursor = NULL; ctx->d_queue = NULL; ctx->d_ws = NULL; ctx->cap_ws = 0; ctx->h_pin = NULL; ctx->cap_pin = 0;
  ctx->stream = 0; ctx->copy_stream = 0; ctx->ev0 = ctx->ev1 = 0;
  for (int k = 0; k < GDP_NK; k++) { ctx->d_kws[k] = NULL; ctx->cap_kws[k] = 0; ctx->kstream[k] = 0; ctx->evj[k] = 0; ctx->evk[k][0] = ctx->evk[k][1] = 0; ctx->last_ms[k] = 0.f; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    ctx->err = "no CUDA device: the gmapdp engine has no CPU fallback";
    return GMAPDP_ERR_CUDA;
  }
  CK(cudaSetDevice(device));
  /* Host threads that wait for the device sleep instead of spinning (cudaStreamSynchronize spins by default): in the
     drop-in hundreds of worker threads share the cores, and a chaining batch keeps its leader waiting for milliseconds.
     Kernel times are taken from CUDA events and do not depend on this.  GMAPDP_SYNC=spin keeps the driver's default. */
  {
    const char *sy = getenv("GMAPDP_SYNC");
    if (!(sy && !strcmp(sy,"spin")) && cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) cudaGetLastError();
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop,device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem = (int) prop.sharedMemPerBlockOptin - 256;	/* dynamic shared memory: the kernels hold a few bytes of static shared memory (the block's queue slot) */
  cudaFuncAttributes fa;
  cudaError_t fe = cudaFuncGetAttributes(&fa,gmapdp_dp_kernel<0>);
  if (fe != cudaSuccess) {
    ctx->err = std::string("no sm_100a kernel image for this device (") + prop.name + "): " + cudaGetErrorString(fe);
    return GMAPDP_ERR_CUDA;
  }
  CK(cudaStreamCreateWithFlags(&ctx->stream,cudaStreamNonBlocking));
  ctx->kstream[0] = ctx->stream;
  /* kstream[1..3] (one stream per kernel kind, GMAPDP_STREAMS=1) are created when first asked for: every stream takes a
     hardware queue, and the drop-in runs a dozen contexts per device */
  for (int k = 0; k < GDP_NK; k++) {
    CK(cudaEventCreateWithFlags(&ctx->evj[k],cudaEventDisableTiming));
    CK(cudaEventCreate(&ctx->evk[k][0])); CK(cudaEventCreate(&ctx->evk[k][1]));
  }
  {
    const char *cm = getenv("GMAPDP_CHUNK_MB");
    long mb = cm ? atol(cm) : 384;
    ctx->chunk_bytes = (size_t) (mb > 0 ? mb : 384) << 20;
  }
  CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1));
  GdpHostTables ht; GdpTables t; ht.device_tables(&t);
  CK(cudaMalloc((void **) &ctx->d_tables,sizeof(GdpTables)));
  CK(cudaMemcpy(ctx->d_tables,&t,sizeof(GdpTables),cudaMemcpyHostToDevice));
  CK(cudaMalloc((void **) &ctx->d_cursor,sizeof(unsigned long long)));
  CK(cudaMalloc((void **) &ctx->d_queue,GDP_NK * sizeof(int)));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel<3>,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  CK(cudaFuncSetAttribute(gmapdp_dp_kernel_any,cudaFuncAttributeMaxDynamicSharedMemorySize,ctx->max_smem));
  ctx->grid = 0;
  return GMAPDP_OK;
}

extern "C" void gmapdp_destroy (gmapdp_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->chain && ctx->chain_free) ctx->chain_free(ctx->chain);
  cudaFree(ctx->d_tables); cudaFree(ctx->d_boxes); cudaFree(ctx->d_order); cudaFree(ctx->d_seq); cudaFree(ctx->d_probs); cudaFree(ctx->d_devprobs);
  cudaFree(ctx->d_results); cudaFree(ctx->d_script); cudaFree(ctx->d_cursor); cudaFree(ctx->d_queue); cudaFree(ctx->d_ws);
  for (int k = 0; k < GDP_NK; k++) cudaFree(ctx->d_kws[k]);
  for (void *g : ctx->grave) cudaFree(g);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->chunk_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->chunk_done) cudaEventDestroy(e);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  if (ctx->h_cursors) cudaFreeHost(ctx->h_cursors);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (int k = 0; k < GDP_NK; k++) {
    if (k > 0 && ctx->kstream[k] && ctx->kstream[k] != ctx->stream) cudaStreamDestroy(ctx->kstream[k]);
    if (ctx->evj[k]) cudaEventDestroy(ctx->evj[k]);
    if (ctx->evk[k][0]) cudaEventDestroy(ctx->evk[k][0]);
    if (ctx->evk[k][1]) cudaEventDestroy(ctx->evk[k][1]);
  }
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
  if (b.mode == GMAPDP_CDNA) w = 2 * w + (double) b.glenL * (b.lbandL + b.ubandL) * (b.lbandR + b.ubandR) * 0.25;
  return w;
}

/* geometry, sorting and device allocations of a batch (no copies).  `order' receives the box ids of
   every chunk [chunk_begin[k], chunk_begin[k+1]) sorted by decreasing work. */
/* Launch order of a chunk: kind (single, end, genome, cdna), then decreasing work (LPT).  LPT does not need an
   exact order, so the work estimate is quantised to 1/16 of an octave (plan_scan stores the bucket in work[i].first)
   and the chunk is counting-sorted: O(n) instead of the 70 ms a comparison sort of 10^6 boxes costs -- the host's
   serial share in front of the launches is what bounds the end-to-end path on production-size boxes. */
#define GDP_WORK_BUCKETS 1024
static inline int work_bucket (int kind, double w) {
  int e;
  const double m = frexp(w + 1.0,&e);				/* w + 1 = m * 2^e, m in [0.5, 1) */
  int q = e * 16 + (int) ((m - 0.5) * 32.0);			/* monotone in w, 16 steps per octave */
  if (q < 0) q = 0;
  if (q > GDP_WORK_BUCKETS - 1) q = GDP_WORK_BUCKETS - 1;
  return kind * GDP_WORK_BUCKETS + (GDP_WORK_BUCKETS - 1 - q);	/* ascending bucket = kind, then decreasing work */
}
/* Boxes of one work bucket are grouped by the code variant they run (GDP_VARIANTS per bucket): the four warps of a block
   take consecutive boxes and go through them in step (gmapdp_dp_kernel), so neighbours should run the same instantiation
   -- single gaps: packed 16-bit / int32 fill, jump-late or not; end gaps: the tie rule the interior steps are templated on.
   GMAPDP_SORT_VARIANT=0 switches the sub-key off. */
#define GDP_VARIANTS 4
static inline int box_variant (const gmapdp_box &b) {
  static const bool on = !(getenv("GMAPDP_SORT_VARIANT") && atoi(getenv("GMAPDP_SORT_VARIANT")) == 0);
  if (!on) return 0;
  const int late = (b.flags & GMAPDP_F_LATE_L) ? 1 : 0;
  if (b.mode == GMAPDP_SINGLE) return (gdp_full_packed(b) ? 0 : 2) | late;
  if (b.mode == GMAPDP_END5 || b.mode == GMAPDP_END3) return late;
  return 0;
}
static inline void sort_chunk (std::vector<std::pair<double,int> > &work, int b0, int b1) {
  std::vector<int> count(GDP_NK * GDP_WORK_BUCKETS * GDP_VARIANTS + 1,0);
  for (int i = b0; i < b1; i++) count[(int) work[i].first + 1]++;
  for (size_t k = 1; k < count.size(); k++) count[k] += count[k-1];
  std::vector<std::pair<double,int> > out(b1 - b0);
  for (int i = b0; i < b1; i++) out[count[(int) work[i].first]++] = work[i];	/* stable: input order within a bucket */
  std::copy(out.begin(),out.end(),work.begin() + b0);
}

/* geometry and device allocations of a batch (no copies).  work[i] = (key, box id); within a chunk
   the sorted order lists the boxes kind by kind (key offset by -1e12 per kind) and chunk_count[k][kind] counts them. */
/* dynamic shared memory of a kernel kind: 8 bytes per column and warp (single gaps: the stripe boundary rows,
   unless they live in the workspace; cdna gaps: the bridge's M and Q tables); the other kinds use none */
static inline size_t stage_smem (int kind) {	/* staging regions of a block's warps (E-only kinds) */
  if (!GMAPDP_STAGE || kind == 0) return 0;
  return (size_t) WARPS_PER_BLOCK * (kind == 2 ? StgGeom<true>::WARP_BYTES : StgGeom<false>::WARP_BYTES);
}
static inline size_t kind_smem (const gmapdp_ctx *ctx, int kind) {
  if (kind == 0 && GMAPDP_BND_GLOBAL) return 0;
  return (size_t) WARPS_PER_BLOCK * ctx->ksmem_cols[kind] * 8 + stage_smem(kind);
}

struct PlanScan { size_t ws_words[GDP_NK], script_need, devprob_need; int maxcols[GDP_NK], me_kinds; };

/* The kernels pack rows and columns into 16-bit fields (best cells `r << 16 | c', the bridges' keys, the cDNA
   bridge's `score << 16 | rR' tables) and the box carries its bands as int16: a side longer than 32767 would
   corrupt best cells silently, so it is refused here.  Penalties are <= 0 (0 is reachable with --indel-open=0). */
#define GDP_MAX_SIDE 32767
static inline bool box_ok (const gmapdp_box &b) {
  if (b.mode < 0 || b.mode > 4 || (unsigned) b.mismatchtype > 3u || b.open > 0 || b.extend > 0) return false;
  if (b.rlenL < 0 || b.glenL < 0 || b.rlenR < 0 || b.glenR < 0) return false;
  if (b.rlenL > GDP_MAX_SIDE || b.glenL > GDP_MAX_SIDE || b.rlenR > GDP_MAX_SIDE || b.glenR > GDP_MAX_SIDE) return false;
  if (b.lbandL < 0 || b.ubandL < 0 || b.lbandR < 0 || b.ubandR < 0) return false;
  return true;
}

/* bytes a box makes the end-to-end path upload (sequences once, alt twins are shared; MaxEnt doubles) */
static inline uint32_t box_upload_bytes (const gmapdp_box &x) {
  size_t acc = (size_t) x.rlenL + x.rlenR;
  if (!(x.gflags & GMAPDP_G_SEG_L)) acc += x.glenL;
  if (!(x.gflags & GMAPDP_G_SEG_R)) acc += x.glenR;
  if (x.mode == GMAPDP_GENOME && !(x.gflags & GMAPDP_G_PROBS)) acc += 8 * ((size_t) x.glenL + x.glenR);
  return (uint32_t) acc;
}

static int plan_scan (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes, std::vector<std::pair<double,int> > &work,
		      std::vector<uint32_t> *upload_bytes, PlanScan &ps) {
  size_t ws_words[GDP_NK] = {0,0,0,0}, script_need = 0, devprob_need = 0; int maxcols[GDP_NK] = {8,8,8,8}, me_kinds = 0;
  work.resize(nboxes);
  if (upload_bytes) upload_bytes->resize(nboxes);
  {
    /* per-box geometry (workspace words, script bound, work estimate) is the host's serial cost in front of the
       first launch of the end-to-end path: spread it over a few threads for large batches */
    struct Part { size_t ws[GDP_NK], script, devprob; int cols[GDP_NK], kinds; bool bad; };
    const int nthreads = (nboxes >= 65536) ? (int) std::min<unsigned>(8u,std::max(1u,std::thread::hardware_concurrency())) : 1;
    std::vector<Part> parts(nthreads);
    auto scan = [&](int t) {
      Part &pt = parts[t];
      for (int k = 0; k < GDP_NK; k++) { pt.ws[k] = 0; pt.cols[k] = 8; }
      pt.script = 0; pt.bad = false; pt.devprob = 0; pt.kinds = 0;
      const int i0 = (int) ((long long) nboxes * t / nthreads), i1 = (int) ((long long) nboxes * (t + 1) / nthreads);
      for (int i = i0; i < i1; i++) {
	const gmapdp_box &b = boxes[i];
	if (!box_ok(b) || (b.gflags && !ctx->genome)) { pt.bad = true; return; }
	if (b.mode == GMAPDP_GENOME && (b.gflags & GMAPDP_G_PROBS)) {
	  if (b.probkindL > 3 || b.probkindR > 3) { pt.bad = true; return; }
	  pt.devprob = std::max(pt.devprob,std::max((size_t) b.probL_off + b.glenL + 1,(size_t) b.probR_off + b.glenR + 1));
	  pt.kinds |= (1 << b.probkindL) | (1 << b.probkindR);
	}
	const int kind = kind_of(b.mode);
	pt.ws[kind] = std::max(pt.ws[kind],gdp_ws_words(b));
	pt.script += (size_t) b.rlenL + b.glenL + 4;
	if (b.mode == GMAPDP_GENOME || b.mode == GMAPDP_CDNA) pt.script += (size_t) b.rlenR + b.glenR + 4;
	if (b.mode == GMAPDP_SINGLE) pt.cols[0] = std::max(pt.cols[0],(int) b.glenL + 2 + GDP_PK_EXTRA);
	else if (b.mode == GMAPDP_CDNA) pt.cols[3] = std::max(pt.cols[3],(int) b.glenL + 2);	/* M and Q tables: 2 words per column */
	work[i] = std::make_pair((double) (work_bucket(kind,box_work(b)) * GDP_VARIANTS + box_variant(b)),i);
	if (upload_bytes) (*upload_bytes)[i] = box_upload_bytes(b);
      }
    };
    if (nthreads == 1) scan(0);
    else {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; t++) th.emplace_back(scan,t);
      for (auto &x : th) x.join();
    }
    for (const Part &pt : parts) {
      if (pt.bad) { ctx->err = "bad box"; return GMAPDP_ERR_ARG; }
      for (int kind = 0; kind < GDP_NK; kind++) { ws_words[kind] = std::max(ws_words[kind],pt.ws[kind]); maxcols[kind] = std::max(maxcols[kind],pt.cols[kind]); }
      script_need += pt.script;
      devprob_need = std::max(devprob_need,pt.devprob); me_kinds |= pt.kinds;
    }
  }
  for (int k = 0; k < GDP_NK; k++) { ps.ws_words[k] = ws_words[k]; ps.maxcols[k] = maxcols[k]; }
  ps.script_need = script_need; ps.devprob_need = devprob_need; ps.me_kinds = me_kinds;
  return GMAPDP_OK;
}

static int plan_finish (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes, size_t seqbytes, size_t nprobs,
			const std::vector<int> &chunk_begin, std::vector<int> &order, std::vector<std::pair<double,int> > &work,
			bool sort_now, const PlanScan &ps) {
  const size_t *ws_words = ps.ws_words, script_need = ps.script_need;
  const int *maxcols = ps.maxcols;
  const int nchunks = (int) chunk_begin.size() - 1;
  ctx->chunk_count.assign((size_t) nchunks * GDP_NK,0);
  int largest[GDP_NK] = {0,0,0,0};
  for (int k = 0; k < nchunks; k++) {
    int *cnt = &ctx->chunk_count[(size_t) k * GDP_NK];
    for (int i = chunk_begin[k]; i < chunk_begin[k+1]; i++) cnt[kind_of(boxes[i].mode)]++;
    for (int kind = 0; kind < GDP_NK; kind++) largest[kind] = std::max(largest[kind],cnt[kind]);
  }
  order.resize(nboxes);
  if (sort_now) {
    for (int k = 0; k < nchunks; k++) sort_chunk(work,chunk_begin[k],chunk_begin[k+1]);
    for (int i = 0; i < nboxes; i++) order[i] = work[i].second;
  }
  ctx->script_need = script_need;

  /* persistent grids: as many blocks per SM as shared memory and registers allow, on every SM */
  ctx->grid = 0;
  for (int kind = 0; kind < GDP_NK; kind++) {
    ctx->kws_words[kind] = (ws_words[kind] + 31) & ~(size_t) 31;
    ctx->ksmem_cols[kind] = (maxcols[kind] + 7) & ~7;
    size_t smem = kind_smem(ctx,kind);
    if ((int) smem > ctx->max_smem) { ctx->err = "box too long for the shared-memory rows"; return GMAPDP_ERR_ARG; }
    int occ = 0;
    if (kind == 0) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel<0>,BLOCK_THREADS,smem));
    else if (kind == 1) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel<1>,BLOCK_THREADS,smem));
    else if (kind == 2) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel<2>,BLOCK_THREADS,smem));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel<3>,BLOCK_THREADS,smem));
    occ = std::min(std::max(occ,1),8);
    int grid = ctx->sm_count * occ;
    int needed_blocks = (largest[kind] + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    ctx->kgrid[kind] = std::max(std::min(grid,needed_blocks),1);
    if (largest[kind] > 0 &&
	grow(ctx,&ctx->d_kws[kind],&ctx->cap_kws[kind],(size_t) ctx->kgrid[kind] * WARPS_PER_BLOCK * ctx->kws_words[kind])) return GMAPDP_ERR_CUDA;
    if (largest[kind] > 0) ctx->grid += ctx->kgrid[kind];
  }

  if (grow(ctx,&ctx->d_boxes,&ctx->cap_boxes,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_order,&ctx->cap_order,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_results,&ctx->cap_results,(size_t) nboxes)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_seq,&ctx->cap_seq,seqbytes + 16)) return GMAPDP_ERR_CUDA;
  if (grow(ctx,&ctx->d_probs,&ctx->cap_probs,nprobs + 2)) return GMAPDP_ERR_CUDA;
  if (ps.devprob_need && grow(ctx,&ctx->d_devprobs,&ctx->cap_devprobs,ps.devprob_need + 2)) return GMAPDP_ERR_CUDA;
  ctx->me_kinds = ps.me_kinds;
  if (grow(ctx,&ctx->d_script,&ctx->cap_script,script_need + 64)) return GMAPDP_ERR_CUDA;
  return GMAPDP_OK;
}

static int plan_batch (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes, size_t seqbytes, size_t nprobs,
		       const std::vector<int> &chunk_begin, std::vector<int> &order, std::vector<std::pair<double,int> > &work,
		       bool sort_now) {
  PlanScan ps;
  int rc = plan_scan(ctx,boxes,nboxes,work,NULL,ps);
  if (rc) return rc;
  return plan_finish(ctx,boxes,nboxes,seqbytes,nprobs,chunk_begin,order,work,sort_now,ps);
}

/* launches the (up to) four kernels of one chunk (on one stream unless GMAPDP_STREAMS is set; kstream[0] == stream); all
   streams must already be ordered after the chunk's uploads.  cnt[kind] boxes of each kind, laid out kind
   by kind in `order' from `first'. */
static int launch_chunk (gmapdp_ctx *ctx, int first, const int *cnt, bool timed = false) {
  /* The four kernels are queued on ONE stream: a persistent grid fills every SM, so the kernels run one after
     the other either way, co-resident kernels of different kinds slow each other down, and on one stream each
     kernel's events bracket exactly its own execution.  GMAPDP_STREAMS=1 puts the kernels of the chunked
     host-buffer path on four streams (measured slower: 292 vs 279 ms end to end). */
  static const bool use_streams = getenv("GMAPDP_STREAMS") != NULL;
  const bool serial = timed || !use_streams;
  static const char *korder = getenv("GMAPDP_ORDER") ? getenv("GMAPDP_ORDER") : "2130";	/* launch order of the kinds: the E-only kernels first, the single-gap kernel fills in as they drain (co-resident kernels of different kinds slow each other down) */
  int start[GDP_NK], acc = first;
  for (int kind = 0; kind < GDP_NK; kind++) { start[kind] = acc; acc += cnt[kind]; if (timed) ctx->last_ms[kind] = 0.f; }
  for (int q = 0; q < GDP_NK; q++) {
    const int kind = (korder[q] - '0') & 3;
    const int count = cnt[kind];
    if (count == 0) continue;
    cudaStream_t st = serial ? ctx->stream : ctx->kstream[kind];
    KernelArgs ka;
    ka.boxes = ctx->d_boxes; ka.order = ctx->d_order + start[kind]; ka.nboxes = count;
    ka.seq = ctx->d_seq; ka.probs = ctx->d_probs; ka.results = ctx->d_results;
    ka.script = ctx->d_script; ka.script_cap = ctx->cap_script; ka.script_cursor = ctx->d_cursor;
    ka.queue = ctx->d_queue + kind; ka.ws = ctx->d_kws[kind]; ka.ws_words = ctx->kws_words[kind];
    ka.smem_cols = ctx->ksmem_cols[kind]; ka.tables = ctx->d_tables; ka.one = 1u;
    ka.genome.blocks = ctx->genome ? ctx->genome->d_blocks : NULL; ka.genome.nwords = ctx->genome ? ctx->genome->nwords : 0; ka.maxent = ctx->genome ? ctx->genome->d_maxent : NULL; ka.devprobs = ctx->d_devprobs;
    const size_t smem = kind_smem(ctx,kind);
    const int grid = std::max(1,std::min(ctx->kgrid[kind],(count + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK));
    CK(cudaMemsetAsync(ctx->d_queue + kind,0,sizeof(int),st));
    if (timed) CK(cudaEventRecord(ctx->evk[kind][0],st));
    if (kind == 2 && ctx->me_kinds && ctx->genome) {
      /* the MaxEnt arrays of this chunk's genome gaps (counted with the genome kernel's time) */
      const int mgrid = std::max(1,std::min(ctx->sm_count * 6,(2 * count + ME_THREADS / 32 - 1) / (ME_THREADS / 32)));
      gmapdp_maxent_kernel_batch<<<mgrid,ME_THREADS,0,st>>>(ctx->d_boxes,ka.order,count,ka.genome,ka.maxent,ctx->d_devprobs);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    if (kind == 0) gmapdp_dp_kernel<0><<<grid,BLOCK_THREADS,smem,st>>>(ka);
    else if (kind == 1) gmapdp_dp_kernel<1><<<grid,BLOCK_THREADS,smem,st>>>(ka);
    else if (kind == 2) gmapdp_dp_kernel<2><<<grid,BLOCK_THREADS,smem,st>>>(ka);
    else gmapdp_dp_kernel<3><<<grid,BLOCK_THREADS,smem,st>>>(ka);
    CK(cudaGetLastError());
    if (timed) CK(cudaEventRecord(ctx->evk[kind][1],st));
    ctx->launches++;
  }
  return GMAPDP_OK;
}

/* fork: the other kernel streams start after an event on `stream' */
static int fork_streams (gmapdp_ctx *ctx, cudaEvent_t ev) {
  static const bool use_streams = getenv("GMAPDP_STREAMS") != NULL;
  for (int k = 1; k < GDP_NK; k++) {
    if (!use_streams) { ctx->kstream[k] = ctx->stream; continue; }		/* everything on the one stream */
    if (ctx->kstream[k] == 0 || ctx->kstream[k] == ctx->stream) CK(cudaStreamCreateWithFlags(&ctx->kstream[k],cudaStreamNonBlocking));
    CK(cudaStreamWaitEvent(ctx->kstream[k],ev,0));
  }
  return GMAPDP_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Resident genome + MaxEnt tables (include/gmapdp_b200.h, csrc/gmapdp_genome.h)
 * ---------------------------------------------------------------------------------------------- */
static void pack_maxent (const gmapdp_maxent_tables *t, std::vector<double> &v) {
  v.assign(GDP_ME_NDOUBLES,0.0);
  const double *big[12] = {t->donor_plus,t->acc1_plus,t->acc2_plus,t->acc3_plus,t->acc467_plus,t->acc589_plus,
			   t->donor_minus,t->acc1_minus,t->acc2_minus,t->acc3_minus,t->acc467_minus,t->acc589_minus};
  for (int k = 0; k < 12; k++) memcpy(&v[(size_t) k * 16384],big[k],16384 * sizeof(double));
  memcpy(&v[GDP_ME_DONOR_DI_P],t->donor_di_plus,16 * sizeof(double)); memcpy(&v[GDP_ME_ACC_DI_P],t->acc_di_plus,16 * sizeof(double));
  memcpy(&v[GDP_ME_DONOR_DI_M],t->donor_di_minus,16 * sizeof(double)); memcpy(&v[GDP_ME_ACC_DI_M],t->acc_di_minus,16 * sizeof(double));
}

extern "C" int gmapdp_genome_create (gmapdp_genome **out, int device, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *maxent) {
  *out = NULL;
  if (!blocks || nwords == 0 || !maxent) return GMAPDP_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return GMAPDP_ERR_CUDA;
  gmapdp_genome *g = new gmapdp_genome();
  g->device = device; g->d_blocks = NULL; g->d_maxent = NULL; g->nwords = nwords;
  std::vector<double> packed;
  pack_maxent(maxent,packed);
  if (cudaMalloc((void **) &g->d_blocks,nwords * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc((void **) &g->d_maxent,packed.size() * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(g->d_blocks,blocks,nwords * sizeof(uint32_t),cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(g->d_maxent,packed.data(),packed.size() * sizeof(double),cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(g->d_blocks); cudaFree(g->d_maxent); delete g;
    return GMAPDP_ERR_CUDA;
  }
  *out = g;
  return GMAPDP_OK;
}

extern "C" void gmapdp_genome_destroy (gmapdp_genome *g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaFree(g->d_blocks); cudaFree(g->d_maxent);
  delete g;
}

extern "C" int gmapdp_genome_attach (gmapdp_ctx *ctx, const gmapdp_genome *g) {
  if (!ctx) return GMAPDP_ERR_ARG;
  if (g && g->device != ctx->device) { ctx->err = "genome lives on another device"; return GMAPDP_ERR_ARG; }
  ctx->genome = g;
  return GMAPDP_OK;
}

extern "C" int gmapdp_genome_host_char (const uint32_t *blocks, size_t nwords, uint32_t pos) {
  GdpGenome g; g.blocks = blocks; g.nwords = nwords;
  return gdp_genome_char(g,pos);
}

extern "C" double gmapdp_maxent_host_prob (const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *t, int kind, uint32_t splice_pos, uint32_t chroffset) {
  /* straight from the caller's tables (no cached copy: a caller may refill its tables in place), same steps as the device */
  const double *big[4][5] = {{t->donor_plus,NULL,NULL,NULL,NULL}, {t->acc1_plus,t->acc2_plus,t->acc3_plus,t->acc467_plus,t->acc589_plus},
			     {t->donor_minus,NULL,NULL,NULL,NULL}, {t->acc1_minus,t->acc2_minus,t->acc3_minus,t->acc467_minus,t->acc589_minus}};
  if (kind < 0 || kind > 3) return 0.0;
  const uint32_t margin = (uint32_t) gdp_maxent_margin(kind);
  if (splice_pos < chroffset + margin) return 0.0;
  double di[64];
  memcpy(di + GDP_DI_DONOR_P,t->donor_di_plus,16 * sizeof(double)); memcpy(di + GDP_DI_ACC_P,t->acc_di_plus,16 * sizeof(double));
  memcpy(di + GDP_DI_DONOR_M,t->donor_di_minus,16 * sizeof(double)); memcpy(di + GDP_DI_ACC_M,t->acc_di_minus,16 * sizeof(double));
  GdpGenome g; g.blocks = blocks; g.nwords = nwords;
  const uint64_t W = gdp_genome_window(g,splice_pos - margin);
  double odds = 0.0;
  for (int p = 0; p < gdp_maxent_npasses(kind); p++) odds = gdp_maxent_step(kind,p,W,big[kind][p],di,odds);
  return odds / (1 + odds);
}

__global__ void gmapdp_maxent_kernel (GdpGenome g, const double *me, const int *kind, const uint32_t *pos, uint32_t chroffset, int n, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = gdp_maxent_prob(kind[i],g,me,pos[i],chroffset);
}

extern "C" int gmapdp_maxent_eval (gmapdp_ctx *ctx, const int *kind, const uint32_t *pos, uint32_t chroffset, int n, double *probs) {
  if (!ctx || !ctx->genome || n < 0) { if (ctx) ctx->err = "no genome attached"; return GMAPDP_ERR_ARG; }
  if (n == 0) return GMAPDP_OK;
  CK(cudaSetDevice(ctx->device));
  int *dk = NULL; uint32_t *dp = NULL; double *dout = NULL;
  CK(cudaMalloc((void **) &dk,(size_t) n * sizeof(int))); CK(cudaMalloc((void **) &dp,(size_t) n * sizeof(uint32_t))); CK(cudaMalloc((void **) &dout,(size_t) n * sizeof(double)));
  CK(cudaMemcpy(dk,kind,(size_t) n * sizeof(int),cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp,pos,(size_t) n * sizeof(uint32_t),cudaMemcpyHostToDevice));
  GdpGenome g; g.blocks = ctx->genome->d_blocks; g.nwords = ctx->genome->nwords;
  gmapdp_maxent_kernel<<<(n + 127) / 128,128,0,ctx->stream>>>(g,ctx->genome->d_maxent,dk,dp,chroffset,n,dout);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(probs,dout,(size_t) n * sizeof(double),cudaMemcpyDeviceToHost,ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dk); cudaFree(dp); cudaFree(dout);
  return GMAPDP_OK;
}

extern "C" int gmapdp_upload (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
			      const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs) {
  if (!ctx || nboxes < 0 || (nboxes > 0 && (!boxes || !seqpool))) { if (ctx) ctx->err = "bad argument"; return GMAPDP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  ctx->nboxes = nboxes;
  if (nboxes == 0) return GMAPDP_OK;
  std::vector<int> chunk_begin = {0, nboxes}, order;
  std::vector<std::pair<double,int> > work;
  int rc = plan_batch(ctx,boxes,nboxes,seqbytes,nprobs,chunk_begin,order,work,true);
  if (rc) return rc;
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
  if (ctx->chunk_count.size() != GDP_NK) {
    /* the plan on the device is the chunked one of gmapdp_run_batch (d_order lists its chunks one after the other):
       launching "chunk 0" would recompute only a part of the boxes */
    ctx->err = "gmapdp_run_resident needs a batch uploaded by gmapdp_upload (the last gmapdp_run_batch was pipelined in several chunks)";
    return GMAPDP_ERR_ARG;
  }
  const int *cnt = &ctx->chunk_count[0];
  CK(cudaMemsetAsync(ctx->d_cursor,0,sizeof(unsigned long long),ctx->stream));
  CK(cudaEventRecord(ctx->ev0,ctx->stream));
  int rc = launch_chunk(ctx,0,cnt,true);		/* all on `stream', in launch order */
  if (rc) return rc;
  CK(cudaEventRecord(ctx->ev1,ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms,ctx->ev0,ctx->ev1));
  for (int kind = 0; kind < GDP_NK; kind++)
    if (cnt[kind] > 0) CK(cudaEventElapsedTime(&ctx->last_ms[kind],ctx->evk[kind][0],ctx->evk[kind][1]));
  return GMAPDP_OK;
}

/* CUDA-event durations of the kernels of the last gmapdp_run_resident (one stream, back to back):
   full_ms = the single-gap kernel, tri_ms = the three E-only kernels together */
extern "C" int gmapdp_last_kernel_ms (const gmapdp_ctx *ctx, float *full_ms, float *tri_ms) {
  if (full_ms) *full_ms = ctx->last_ms[0];
  if (tri_ms) *tri_ms = ctx->last_ms[1] + ctx->last_ms[2] + ctx->last_ms[3];
  return GMAPDP_OK;
}

/* the same per kind: ms[0..3] = single, end, genome, cdna */
extern "C" int gmapdp_last_kernel_ms4 (const gmapdp_ctx *ctx, float *ms) {
  for (int kind = 0; kind < GDP_NK; kind++) ms[kind] = ctx->last_ms[kind];
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

/* Host-buffer path.  Large batches are cut into chunks of consecutive boxes: the H2D copy of chunk
   k+1 (copy stream) overlaps the kernels of chunk k (compute streams); each chunk uploads only the span
   of the pools its boxes reference.  Small batches are one chunk. */
extern "C" int gmapdp_run_batch (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
				 const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs,
				 gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used) {
  return gmapdp_run_batch_chunks(ctx,boxes,nboxes,seqpool,seqbytes,probpool,nprobs,results,script,script_cap,script_used,NULL,NULL);
}

/* The same with a completion callback per chunk: as soon as the results and edit scripts of boxes [first, first + n) are
   in the caller's buffers, on_chunk(user,first,n) is called (on the calling thread, chunks in order) while the device
   works on the following chunks -- the caller's post-processing (the shim's replay into pair lists) then overlaps the
   kernels instead of following them. */
extern "C" int gmapdp_run_batch_chunks (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
					const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs,
					gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used,
					gmapdp_chunk_fn on_chunk, void *user) {
  if (!ctx || nboxes < 0 || (nboxes > 0 && (!boxes || !seqpool))) { if (ctx) ctx->err = "bad argument"; return GMAPDP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  ctx->nboxes = nboxes;
  if (script_used) *script_used = 0;
  if (nboxes == 0) return GMAPDP_OK;

  /* GMAPDP_TRACE=1: host-side timeline of this call on stderr */
  static const bool trace = getenv("GMAPDP_TRACE") != NULL;
  const auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (trace) fprintf(stderr,"gmapdp_run_batch: %-22s %8.2f ms\n",what,
		       std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now() - T0).count());
  };
  /* per-box geometry in parallel, then chunk boundaries by uploaded bytes (a short first chunk lets the device start
     early); the pool extents of a chunk are found just before its upload, under the previous chunk's kernels */
  const size_t CHUNK_BYTES = ctx->chunk_bytes;
  std::vector<int> order;
  std::vector<std::pair<double,int> > work;
  std::vector<uint32_t> upbytes;
  PlanScan ps;
  int rc = plan_scan(ctx,boxes,nboxes,work,&upbytes,ps);
  if (rc) return rc;
  lap("box scan");
  /* Chunks grow geometrically from CHUNK_BYTES / 4: a short first chunk lets the device start early, and since a
     chunk uploads faster than it computes, the next one may be larger and still arrive in time -- fewer chunk
     boundaries, where the persistent grids drain and refill.  GMAPDP_EQUAL_CHUNKS=1 keeps them small and equal
     (CHUNK_BYTES / 8) when a per-chunk consumer is attached, so that less of its work is left for the end; measured with
     the shim's replay it is SLOWER (200 k benchmark boxes: 113 vs 85 ms end to end, 1 M boxes: 540 vs 385 ms; not a matter
     of the handing-over thread lacking a core: with 15 or 14 replay threads on the 16 cores 116 / 117 ms) -- every boundary
     drains four persistent grids whose largest boxes run for milliseconds. */
  std::vector<int> chunk_begin(1,0);
  {
    static const bool equal_on = (getenv("GMAPDP_EQUAL_CHUNKS") && atoi(getenv("GMAPDP_EQUAL_CHUNKS")) != 0);
    const bool equal_chunks = (on_chunk != NULL) && equal_on;
    size_t acc = 0, limit = equal_chunks ? CHUNK_BYTES / 8 : CHUNK_BYTES / 4;
    for (int i = 0; i < nboxes; i++) {
      acc += upbytes[i];
      if (acc >= limit || i == nboxes - 1) { chunk_begin.push_back(i + 1); acc = 0; if (!equal_chunks) limit = std::min(limit * 2,CHUNK_BYTES * 8); }
    }
  }
  const int nchunks = (int) chunk_begin.size() - 1;
  rc = plan_finish(ctx,boxes,nboxes,seqbytes,nprobs,chunk_begin,order,work,false,ps);	/* chunks are sorted just in time, below */
  if (rc) return rc;
  lap("plan");
  if (ctx->copy_stream == 0) CK(cudaStreamCreateWithFlags(&ctx->copy_stream,cudaStreamNonBlocking));
  while ((int) ctx->chunk_events.size() < nchunks) {
    cudaEvent_t e; CK(cudaEventCreateWithFlags(&e,cudaEventDisableTiming)); ctx->chunk_events.push_back(e);
  }
  /* per-chunk completion needs the kernels of a chunk to be over before the next chunk's begin: the one-stream order */
  static const bool multi_streams = getenv("GMAPDP_STREAMS") != NULL;
  const bool per_chunk = (on_chunk != NULL) && !multi_streams && nchunks > 1;
  if (per_chunk) {
    if (ctx->d2h_stream == 0) CK(cudaStreamCreateWithFlags(&ctx->d2h_stream,cudaStreamNonBlocking));
    while ((int) ctx->chunk_done.size() < nchunks) {
      cudaEvent_t e; CK(cudaEventCreateWithFlags(&e,cudaEventDisableTiming | cudaEventBlockingSync)); ctx->chunk_done.push_back(e);
    }
    if (ctx->cap_cursors < (size_t) nchunks) {
      if (ctx->h_cursors) cudaFreeHost(ctx->h_cursors);
      ctx->h_cursors = NULL; ctx->cap_cursors = 0;
      CK(cudaHostAlloc((void **) &ctx->h_cursors,(size_t) (nchunks + 16) * sizeof(unsigned long long),cudaHostAllocDefault));
      ctx->cap_cursors = (size_t) nchunks + 16;
    }
  }
  CK(cudaMemsetAsync(ctx->d_cursor,0,sizeof(unsigned long long),ctx->stream));
  CK(cudaEventRecord(ctx->evj[0],ctx->stream));
  rc = fork_streams(ctx,ctx->evj[0]);
  if (rc) return rc;
  for (int k = 0; k < nchunks; k++) {
    const int b0 = chunk_begin[k], n = chunk_begin[k+1] - b0;
    cudaStream_t cs = ctx->copy_stream;
    size_t slo = (size_t) -1, shi = 0, plo = (size_t) -1, phi = 0;	/* extents of this chunk in the two pools */
    for (int i = b0; i < b0 + n; i++) {
      const gmapdp_box &x = boxes[i];
      const size_t offs[6] = {x.qL_off, x.qR_off, x.gL_off, x.gLalt_off, x.gR_off, x.gRalt_off};
      const size_t lens[6] = {(size_t) x.rlenL, (size_t) x.rlenR, (size_t) x.glenL, (size_t) x.glenL, (size_t) x.glenR, (size_t) x.glenR};
      /* segments that come from the resident genome are coordinates, not pool offsets */
      const bool inpool[6] = {true, true, !(x.gflags & GMAPDP_G_SEG_L), !(x.gflags & GMAPDP_G_SEG_L), !(x.gflags & GMAPDP_G_SEG_R), !(x.gflags & GMAPDP_G_SEG_R)};
      for (int j = 0; j < 6; j++) if (inpool[j]) { slo = std::min(slo,offs[j]); shi = std::max(shi,offs[j] + lens[j]); }
      if (x.mode == GMAPDP_GENOME && !(x.gflags & GMAPDP_G_PROBS)) {
	plo = std::min(plo,std::min((size_t) x.probL_off,(size_t) x.probR_off));
	phi = std::max(phi,std::max((size_t) x.probL_off + x.glenL,(size_t) x.probR_off + x.glenR));
      }
    }
    shi = std::min(shi,seqbytes); phi = std::min(phi,nprobs);
    sort_chunk(work,b0,b0 + n);				/* overlaps the previous chunk's kernels */
    for (int i = b0; i < b0 + n; i++) order[i] = work[i].second;
    CK(cudaMemcpyAsync(ctx->d_boxes + b0,boxes + b0,(size_t) n * sizeof(gmapdp_box),cudaMemcpyHostToDevice,cs));
    CK(cudaMemcpyAsync(ctx->d_order + b0,order.data() + b0,(size_t) n * sizeof(int),cudaMemcpyHostToDevice,cs));
    if (shi > slo) CK(cudaMemcpyAsync(ctx->d_seq + slo,seqpool + slo,shi - slo,cudaMemcpyHostToDevice,cs));
    if (phi > plo && plo != (size_t) -1)
      CK(cudaMemcpyAsync(ctx->d_probs + plo,probpool + plo,(phi - plo) * sizeof(double),cudaMemcpyHostToDevice,cs));
    CK(cudaEventRecord(ctx->chunk_events[k],cs));
    for (int kk = 0; kk < GDP_NK; kk++) CK(cudaStreamWaitEvent(ctx->kstream[kk],ctx->chunk_events[k],0));
    rc = launch_chunk(ctx,b0,&ctx->chunk_count[(size_t) k * GDP_NK]);
    if (rc) return rc;
    if (per_chunk) {
      /* the cursor as this chunk left it (its scripts end there), then its results on the download stream */
      CK(cudaMemcpyAsync(&ctx->h_cursors[k],ctx->d_cursor,sizeof(unsigned long long),cudaMemcpyDeviceToHost,ctx->stream));
      CK(cudaEventRecord(ctx->chunk_done[k],ctx->stream));
      CK(cudaStreamWaitEvent(ctx->d2h_stream,ctx->chunk_done[k],0));
      CK(cudaMemcpyAsync(results + b0,ctx->d_results + b0,(size_t) n * sizeof(gmapdp_result),cudaMemcpyDeviceToHost,ctx->d2h_stream));
    }
    if (k == 0) lap("first chunk launched");
  }
  lap("all chunks launched");
  if (per_chunk) {
    unsigned long long done_words = 0;
    for (int k = 0; k < nchunks; k++) {
      CK(cudaEventSynchronize(ctx->chunk_done[k]));
      const unsigned long long used = ctx->h_cursors[k];
      if (script_used) *script_used = (size_t) used;
      if (used > ctx->cap_script) { ctx->err = "device script pool overflow"; cudaDeviceSynchronize(); return GMAPDP_ERR_CAPACITY; }
      if (used > script_cap) { ctx->err = "script buffer too small"; cudaDeviceSynchronize(); return GMAPDP_ERR_CAPACITY; }
      if (used > done_words)
	CK(cudaMemcpyAsync(script + done_words,ctx->d_script + done_words,(size_t) (used - done_words) * sizeof(uint32_t),cudaMemcpyDeviceToHost,ctx->d2h_stream));
      done_words = used;
      CK(cudaStreamSynchronize(ctx->d2h_stream));
      on_chunk(user,chunk_begin[k],chunk_begin[k+1] - chunk_begin[k]);
    }
    CK(cudaStreamSynchronize(ctx->copy_stream));
    lap("chunks completed");
    return GMAPDP_OK;
  }
  for (int kk = 0; kk < GDP_NK; kk++) CK(cudaStreamSynchronize(ctx->kstream[kk]));
  CK(cudaStreamSynchronize(ctx->copy_stream));
  lap("kernels done");
  rc = gmapdp_download(ctx,results,script,script_cap,script_used);
  lap("results downloaded");
  if (rc == GMAPDP_OK && on_chunk) on_chunk(user,0,nboxes);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Flights: the device half of the streaming runtime (gmapdp_stream.cpp).  gmapdp_run_batch is built for one
 * large batch (parallel planning, chunked uploads); a drop-in that serves the dependent DP calls of many worker
 * threads needs the opposite: thousands of small batches per second with a few tens of microseconds of fixed
 * cost each.  A flight owns pinned host staging that the submitting threads fill in place, device twins of it,
 * and is run by two H2D copies, ONE all-kinds kernel (gmapdp_dp_kernel_any: a small flight does not fill the GPU, so
 * its boxes of all kinds share one grid) and ONE D2H copy of cursor + results + script.
 * Nothing is allocated, planned or synchronised per flight besides what follows.
 * ---------------------------------------------------------------------------------------------- */
int gdp_bucket_count (void) { return GDP_NK * GDP_WORK_BUCKETS; }

int gdp_box_geometry (const gmapdp_box *b, GdpBoxGeom *g) {
  if (!box_ok(*b)) return GMAPDP_ERR_ARG;
  g->kind = kind_of(b->mode);
  g->work = box_work(*b);
  g->cols = (b->mode == GMAPDP_SINGLE) ? (int) b->glenL + 2 + GDP_PK_EXTRA : (b->mode == GMAPDP_CDNA ? (int) b->glenL + 2 : 8);
  g->ws_words = gdp_ws_words(*b);
  if (b->mode == GMAPDP_SINGLE) {
    const FGeom f = fgeom(b->rlenL,b->glenL,b->lbandL,b->ubandL);
    g->steps = f.nstripes * f.T;
  } else {
    TriPacking tp;
    tri_fills_of(*b,tp);
    g->steps = tp.npasses * (tp.Tmax + 1);
    if (b->mode == GMAPDP_CDNA) g->steps += (b->glenL / 32 + 1) * (b->lbandL + b->ubandL + 1);	/* the bridge's (cL, rL) loop */
  }
  g->bucket = work_bucket(0,(double) g->steps);		/* flights: one queue for all kinds, longest boxes first */
  g->script_words = (size_t) b->rlenL + b->glenL + 4;
  if (b->mode == GMAPDP_GENOME || b->mode == GMAPDP_CDNA) g->script_words += (size_t) b->rlenR + b->glenR + 4;
  return GMAPDP_OK;
}

#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    f->ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return GMAPDP_ERR_CUDA; } } while (0)

static inline size_t flight_in_bytes (int nboxes) { return 32 + (size_t) nboxes * (sizeof(gmapdp_box) + sizeof(int)); }
static inline size_t flight_out_bytes (int nboxes, size_t script_words) { return (size_t) nboxes * sizeof(gmapdp_result) + script_words * sizeof(uint32_t); }

int gdp_flight_create (gmapdp_ctx *ctx, GdpFlight **out, int max_boxes, size_t pool_cap, size_t script_cap) {
  GdpFlight *f = new GdpFlight();
  memset(f,0,sizeof(*f));
  *out = f;
  f->ctx = ctx; f->max_boxes = max_boxes; f->pool_cap = pool_cap; f->script_cap = script_cap;
  FCK(cudaSetDevice(ctx->device));
  const size_t inb = flight_in_bytes(max_boxes), outb = flight_out_bytes(max_boxes,script_cap);
  FCK(cudaHostAlloc((void **) &f->h_in,inb,cudaHostAllocDefault));
  FCK(cudaHostAlloc((void **) &f->h_pool,pool_cap + 64,cudaHostAllocDefault));
  FCK(cudaHostAlloc((void **) &f->h_out,outb,cudaHostAllocDefault));
  memset(f->h_in,0,32);
  FCK(cudaMalloc((void **) &f->d_in,inb));
  FCK(cudaMalloc((void **) &f->d_pool,pool_cap + 64));
  FCK(cudaMalloc((void **) &f->d_out,outb));
  cudaEvent_t e;
  /* GMAPDP_STREAM_TIMING=1: the flight's events keep time stamps, so that the runtime can report the device's share of a
     flight's latency.  ev_done is polled (cudaEventQuery) by the lane's service thread: no interrupt path, no blocking flag. */
  f->timed = getenv("GMAPDP_STREAM_TIMING") != NULL;
  FCK(cudaEventCreateWithFlags(&e,f->timed ? cudaEventDefault : cudaEventDisableTiming)); f->ev_in = (void *) e;
  FCK(cudaEventCreateWithFlags(&e,f->timed ? cudaEventDefault : cudaEventDisableTiming)); f->ev_done = (void *) e;
  FCK(cudaEventCreateWithFlags(&e,f->timed ? cudaEventDefault : cudaEventDisableTiming)); f->ev_start = (void *) e;
  return GMAPDP_OK;
}

void gdp_flight_destroy (GdpFlight *f) {
  if (!f) return;
  cudaSetDevice(f->ctx->device);
  cudaFreeHost(f->h_in); cudaFreeHost(f->h_pool); cudaFreeHost(f->h_out);
  cudaFree(f->d_in); cudaFree(f->d_pool); cudaFree(f->d_out);
  if (f->ev_in) cudaEventDestroy((cudaEvent_t) f->ev_in);
  if (f->ev_done) cudaEventDestroy((cudaEvent_t) f->ev_done);
  if (f->ev_start) cudaEventDestroy((cudaEvent_t) f->ev_start);
  delete f;
}

int gdp_flight_launch (GdpFlight *f, int n, size_t poolbytes, size_t script_need, size_t ws_words, int maxcols) {
  gmapdp_ctx *ctx = f->ctx;
  if (n <= 0 || n > f->max_boxes || poolbytes > f->pool_cap || script_need > f->script_cap) {
    ctx->err = "flight over capacity"; return GMAPDP_ERR_CAPACITY;
  }
  if (!ctx->dev_set) { FCK(cudaSetDevice(ctx->device)); ctx->dev_set = true; }	/* the launcher thread serves one device */
  f->n = n; f->script_need = script_need;
  cudaStream_t s0 = ctx->stream;		/* the flights of a lane run one after the other on its stream and share its workspace */
  const size_t wsw = (ws_words + 31) & ~(size_t) 31;
  const int cols = (maxcols + 7) & ~7;
  const size_t smem = (size_t) WARPS_PER_BLOCK * cols * 8 + (GMAPDP_STAGE ? (size_t) WARPS_PER_BLOCK * STG_WARP_BYTES_MAX : 0);
  if ((int) smem > ctx->max_smem) { ctx->err = "box too long for the shared-memory rows"; return GMAPDP_ERR_ARG; }
  if (ctx->any_occ == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,gmapdp_dp_kernel_any,BLOCK_THREADS,(size_t) ctx->max_smem / 3) != cudaSuccess) { cudaGetLastError(); occ = 1; }
    ctx->any_occ = std::min(std::max(occ,1),3);
  }
  const int full = ctx->sm_count * ctx->any_occ;
  const int grid = std::max(1,std::min(full,(n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK));
  /* the workspace is sized for a full grid, so that it stops growing after the first large boxes (grown geometrically,
     the outgrown buffer parked: the previous flight may still be using it) */
  if (grow(ctx,&ctx->d_kws[0],&ctx->cap_kws[0],(size_t) full * WARPS_PER_BLOCK * wsw)) return GMAPDP_ERR_CUDA;
  if (f->timed) FCK(cudaEventRecord((cudaEvent_t) f->ev_start,s0));
  FCK(cudaMemcpyAsync(f->d_in,f->h_in,flight_in_bytes(n),cudaMemcpyHostToDevice,s0));
  if (poolbytes) FCK(cudaMemcpyAsync(f->d_pool,f->h_pool,poolbytes,cudaMemcpyHostToDevice,s0));
  if (f->timed) FCK(cudaEventRecord((cudaEvent_t) f->ev_in,s0));
  KernelArgs ka;
  ka.boxes = reinterpret_cast<const gmapdp_box *>(f->d_in + 32);
  ka.order = reinterpret_cast<const int *>(f->d_in + 32 + (size_t) n * sizeof(gmapdp_box)); ka.nboxes = n;
  ka.seq = f->d_pool; ka.probs = reinterpret_cast<const double *>(f->d_pool);
  ka.results = reinterpret_cast<gmapdp_result *>(f->d_out);
  ka.script = reinterpret_cast<uint32_t *>(f->d_out + (size_t) n * sizeof(gmapdp_result)); ka.script_cap = script_need;
  ka.queue = reinterpret_cast<int *>(f->d_in);						/* control block: queue head ... */
  ka.script_cursor = reinterpret_cast<unsigned long long *>(f->d_in + 8);		/* ... and script cursor, zeroed by the upload */
  ka.ws = ctx->d_kws[0]; ka.ws_words = wsw;
  ka.smem_cols = cols; ka.tables = ctx->d_tables; ka.one = 1u;
  ka.genome.blocks = ctx->genome ? ctx->genome->d_blocks : NULL; ka.genome.nwords = ctx->genome ? ctx->genome->nwords : 0; ka.maxent = ctx->genome ? ctx->genome->d_maxent : NULL; ka.devprobs = ctx->d_devprobs;
  gmapdp_dp_kernel_any<<<grid,BLOCK_THREADS,smem,s0>>>(ka);
  FCK(cudaGetLastError());
  ctx->launches++;
  FCK(cudaMemcpyAsync(f->h_out,f->d_out,flight_out_bytes(n,script_need),cudaMemcpyDeviceToHost,s0));
  FCK(cudaEventRecord((cudaEvent_t) f->ev_done,s0));
  return GMAPDP_OK;
}

int gdp_flight_poll (GdpFlight *f) {
  const cudaError_t e = cudaEventQuery((cudaEvent_t) f->ev_done);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  f->ctx->err = std::string("flight: ") + cudaGetErrorString(e);
  return GMAPDP_ERR_CUDA;
}

/* GMAPDP_STREAM_TIMING: has the device begun this flight?  (1 yes, 0 not yet) */
int gdp_flight_started (GdpFlight *f) {
  if (!f->timed) return 1;
  return cudaEventQuery((cudaEvent_t) f->ev_start) == cudaSuccess ? 1 : 0;
}

int gdp_flight_wait (GdpFlight *f) {
  FCK(cudaEventSynchronize((cudaEvent_t) f->ev_done));
  f->gpu_ms = f->copy_ms = 0.f;
  if (f->timed) {
    FCK(cudaEventElapsedTime(&f->gpu_ms,(cudaEvent_t) f->ev_start,(cudaEvent_t) f->ev_done));
    FCK(cudaEventElapsedTime(&f->copy_ms,(cudaEvent_t) f->ev_start,(cudaEvent_t) f->ev_in));
  }
  return GMAPDP_OK;
}

int gdp_flight_results (GdpFlight *f, const gmapdp_result **results, const uint32_t **script) {
  *results = reinterpret_cast<const gmapdp_result *>(f->h_out);
  *script = reinterpret_cast<const uint32_t *>(f->h_out + (size_t) f->n * sizeof(gmapdp_result));
  return GMAPDP_OK;
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
