/* Stage-2 chaining on the B200 (ABI in include/gmapchain_b200.h).
 *
 * What the reference computes (stage2.c:3667-4120, :1073-2020, :2956, :3437, :4140, :4402) is a chain DP over the k-mer
 * hits of one (query, genomic region): queryposes in ascending order, each hit choosing its best predecessor among
 * the active hits of the queryposes processed before it.  The recurrence is sequential in querypos AND in the hits of
 * one querypos (the frontier of score_querypos_lookback_mult is carried from hit to hit), so the parallel axis is the
 * problem: a persistent grid, ONE WARP PER PROBLEM pulled from a work-sorted queue.  Lane 0 walks the recurrence; all
 * 32 lanes do the data-parallel parts (frontier set-up, active-list revision of wide queryposes, best-score reduction,
 * candidate compaction, ranking by repeated arg-min, one traceback per lane).  Many resident warps hide the dependent
 * HBM/L2 latency of the walk.
 *
 * State in HBM, structure-of-arrays over the batch's hit pool (seven int32 per hit: consecutive, rootposition,
 * prev querypos, prev hit, trace label, score, next-active) plus two per querypos (first-active, processed stack).
 * After the fill the trace-label and next-active arrays are dead and are reused by the ranking (candidates, kept cells).
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gmapchain_b200.h"
#include "gmapdp_internal.h"

#define CH_BLOCK 128
#define CH_WARPS (CH_BLOCK / 32)
#define CH_FRONTIER 256

/* stage2.c:34-110 */
#define ENOUGH_CONSECUTIVE 32
#define GREEDY_NCONSECUTIVE 100
#define MAX_NACTIVE 100
#define MAX_SKIPPED 3
#define EXON_DEFN 30
#define SCORE_FOR_RESTRICT 10
#define EQUAL_DISTANCE_NOT_SPLICING 9
#define TEN_THOUSAND 8192
#define FINAL_SCORE_TOLERANCE 20
#define MIN_TERMINAL_NCONSECUTIVE 8

struct ChainParams { int splicingp, sufflookback, nsufflookback, maxintronlen; };

struct ChainDev {
  const gmapchain_problem *problems; const int *order; int nproblems;
  const int *npos; const uint32_t *cum, *mina, *maxa, *pos;
  int *consec, *root, *ppos, *phit, *trace, *score, *next;
  int *first, *proc;
  gmapchain_result *results; gmapchain_path *paths; int *pairs;
  unsigned long long paths_cap, pairs_cap;
  unsigned long long *counters;		/* [0] queue cursor, [1] path records used, [2] pairs used */
  ChainParams prm;
};

struct Prob {
  const int *npos; const uint32_t *cum, *mina, *maxa, *pos;
  int *consec, *root, *ppos, *phit, *trace, *score, *next, *first, *proc;
  int L, tot, qs, qe, k, localp, skiprep, favor_right, middlep, max_nal;
  int nproc, tracei;
};

struct Best { int consec, root, score, pp, ph, trace; };

/* ranges 2 and 3+4 of section D for one earlier querypos, from active hit `ph` on (stage2.c:1236-1420, :1712-1915) */
__device__ __forceinline__ void scan_prev (const Prob &P, const ChainParams &prm, Best &b, int &tracei,
					   uint32_t position, int qd, int pq, int ph) {
  const int o = (int) P.cum[pq], credit = -qd / P.k;
  uint32_t pp;
  while (ph != -1 && (pp = P.pos[o + ph]) + EQUAL_DISTANCE_NOT_SPLICING + qd < position) {
    const int diff = (int) (position - pp) - qd;
    int s = P.score[o + ph] + credit;
    s -= prm.splicingp ? (diff / TEN_THOUSAND + 1) : (diff + 1);
    if (s > b.score) {
      b.consec = 0;			/* diff > EQUAL_DISTANCE_FOR_CONSECUTIVE (0) in this range */
      b.root = P.root[o + ph]; b.score = s; b.pp = pq; b.ph = ph; b.trace = ++tracei;
    }
    ph = P.next[o + ph];
  }
  while (ph != -1 && (pp = P.pos[o + ph]) + P.k <= position) {
    const int gd = (int) (position - pp);
    const int s = P.score[o + ph] + 1;
    if (s > b.score) {
      b.consec = (gd == qd) ? P.consec[o + ph] + qd : 0;
      b.root = P.root[o + ph]; b.score = s; b.pp = pq; b.ph = ph; b.trace = P.trace[o + ph];
    }
    ph = P.next[o + ph];
  }
}

__device__ __forceinline__ void commit (Prob &P, int i, const Best &b) {	/* stage2.c:1428-1450 */
  P.consec[i] = b.consec; P.root[i] = b.root; P.ppos[i] = b.pp; P.phit[i] = b.ph;
  if (b.pp >= 0) { P.trace[i] = b.trace; P.score[i] = b.score; }
  else if (P.localp) { P.trace[i] = ++P.tracei; P.score[i] = P.k; }
  else { P.trace[i] = ++P.tracei; P.score[i] = b.score; }
}

/* lane 0: one querypos with a single hit in its window (score_querypos_lookback_one, stage2.c:1073) */
__device__ void score_one (Prob &P, const ChainParams &prm, int q, int h) {
  const int i = (int) P.cum[q] + h;
  const uint32_t position = P.pos[i];
  Best b = { P.k, (int) position, 0, -1, -1, 0 };
  int nlookback = prm.nsufflookback, lookback = prm.sufflookback;
  if (P.nproc > 0) {
    const int pq = P.proc[P.nproc - 1], o = (int) P.cum[pq], qd = q - pq;
    int ph = P.first[pq];
    uint32_t pp = position;
    while (ph != -1 && (pp = P.pos[o + ph]) + qd < position) ph = P.next[o + ph];
    if (pp + qd == position) {
      b.consec = P.consec[o + ph] + qd; b.root = P.root[o + ph]; b.score = P.score[o + ph] + qd;
      b.pp = pq; b.ph = ph; b.trace = P.trace[o + ph];
      nlookback = 1; lookback = prm.sufflookback / 2;
    }
  }
  bool donep = false;
  int nseen = 0, last_trace = -1;
  for (int j = P.nproc - 1; j >= 0 && b.consec < ENOUGH_CONSECUTIVE && !donep; j--, nseen++) {
    const int pq = P.proc[j], qd = q - pq;
    if (nseen > nlookback && qd - P.k > lookback) donep = true;
    int ph = P.first[pq];
    if (ph != -1) {
      const int o = (int) P.cum[pq];
      while (ph != -1 && P.trace[o + ph] == last_trace) ph = P.next[o + ph];
      if (ph != -1) last_trace = P.trace[o + ph];
      if (prm.splicingp) {
	while (ph != -1 && P.pos[o + ph] + prm.maxintronlen + qd <= position) ph = P.next[o + ph];
      }
      scan_prev(P,prm,b,P.tracei,position,qd,pq,ph);
    }
  }
  commit(P,i,b);
}

/* one querypos with several hits in its window (score_querypos_lookback_mult, stage2.c:1470).  All lanes enter;
   lane 0 walks, the others help with the frontier. */
__device__ void score_mult (Prob &P, const ChainParams &prm, int q, int lo, int hi, int *frontier, int lane) {
  const int base = (int) P.cum[q];
  if (P.nproc == 0) {
    if (lane == 0) {
      for (int h = lo; h < hi; h++) {
	const int i = base + h;
	P.consec[i] = P.k; P.root[i] = (int) P.pos[i]; P.ppos[i] = -1; P.phit[i] = -1;
	if (P.localp) { P.trace[i] = ++P.tracei; P.score[i] = P.k; }
	else P.score[i] = 0;
      }
    }
    return;
  }
  /* how far back to look (queryposes are processed in ascending order, so the distance grows along the stack and
     both tests of stage2.c:1626-1631 hold on a prefix): the last index of each prefix, and the frontier of that many */
  int max_adj = 0, max_nonadj = 0;
  {
    int n_adj = 0, n_non = 0;		/* number of leading stack entries that satisfy each test */
    for (int s0 = 0; s0 < P.nproc; s0 += 32) {
      const int ns = s0 + lane;
      bool a = false, b2 = false;
      if (ns < P.nproc) {
	const int pq = P.proc[P.nproc - 1 - ns], qd = q - pq;
	a = (ns <= 1 || qd - P.k <= prm.sufflookback / 2);
	b2 = (ns <= prm.nsufflookback || qd - P.k <= prm.sufflookback);
	if (b2 && ns < CH_FRONTIER) frontier[ns] = P.first[pq];
      }
      const unsigned ma = __ballot_sync(0xffffffffu,a), mb = __ballot_sync(0xffffffffu,b2);
      n_adj += __popc(ma); n_non += __popc(mb);
      if (mb != 0xffffffffu) break;	/* the wider test failed somewhere in this group: nothing further back counts */
    }
    max_adj = n_adj - 1; max_nonadj = n_non - 1;
    __syncwarp();
  }
  if (lane != 0) return;
  const int adjq = P.proc[P.nproc - 1], adjo = (int) P.cum[adjq], adjqd = q - adjq;
  int overall = 0, adjf = P.first[adjq];
  for (int h = lo; h < hi; h++) {
    const uint32_t position = P.pos[base + h];
    int ph = adjf;
    uint32_t pp = position;
    while (ph != -1 && (pp = P.pos[adjo + ph]) + adjqd < position) ph = P.next[adjo + ph];
    adjf = ph;
    if (pp + adjqd == position) overall = max(overall,P.consec[adjo + ph] + adjqd);
  }
  adjf = P.first[adjq];
  for (int h = lo; h < hi; h++) {
    const uint32_t position = P.pos[base + h];
    Best b;
    int ph = adjf, max_nseen;
    uint32_t pp = position;
    while (ph != -1 && (pp = P.pos[adjo + ph]) + adjqd < position) ph = P.next[adjo + ph];
    adjf = ph;
    if (pp + adjqd == position) {
      b.consec = P.consec[adjo + ph] + adjqd; b.root = P.root[adjo + ph]; b.pp = adjq; b.ph = ph;
      b.score = P.score[adjo + ph] + adjqd; b.trace = P.trace[adjo + ph];
      max_nseen = max_adj;
    } else {
      b.consec = P.k; b.root = (int) position; b.pp = -1; b.ph = -1; b.score = 0; b.trace = -1;
      max_nseen = max_nonadj;
    }
    if (overall < GREEDY_NCONSECUTIVE) {
      int last_trace = -1;
      for (int nseen = 0; nseen < P.nproc && b.consec < ENOUGH_CONSECUTIVE && nseen <= max_nseen; nseen++) {
	ph = frontier[nseen];
	if (ph != -1) {
	  const int pq = P.proc[P.nproc - 1 - nseen], o = (int) P.cum[pq], qd = q - pq;
	  while (ph != -1 && P.trace[o + ph] == last_trace) ph = P.next[o + ph];
	  if (ph != -1) last_trace = P.trace[o + ph];
	  while (ph != -1 && P.pos[o + ph] + prm.maxintronlen + qd <= position) ph = P.next[o + ph];
	  frontier[nseen] = ph;
	  scan_prev(P,prm,b,P.tracei,position,qd,pq,ph);
	}
      }
    }
    commit(P,base + h,b);
  }
}

/* active list of a querypos = its hits within SCORE_FOR_RESTRICT of its best (revise_active_lookback, stage2.c:2956) */
__device__ void revise_active (Prob &P, int q, int lo, int hi) {
  int *sc = P.score + P.cum[q], *nx = P.next + P.cum[q];
  if (lo >= hi) { P.first[q] = -1; return; }
  int best = sc[lo];
  for (int h = lo + 1; h < hi; h++) best = max(best,sc[h]);
  const int thr = max(best - SCORE_FOR_RESTRICT,0);
  int prev = -1, firsth = -1;
  for (int h = lo; h < hi; h++) {
    if (sc[h] > thr) {
      if (prev < 0) firsth = h; else nx[prev] = h;
      prev = h;
    }
  }
  if (prev >= 0) nx[prev] = -1;
  P.first[q] = firsth;
}

__device__ void new_start (Prob &P, int q) {	/* stage2.c:3793-3812, :3941-3964 */
  const int base = (int) P.cum[q];
  for (int h = 0; h < P.npos[q]; h++) {
    const int i = base + h;
    P.ppos[i] = P.phit[i] = -1; P.consec[i] = P.k; P.trace[i] = -1; P.score[i] = P.k;
  }
}

/* align_compute_scores_lookback (stage2.c:3667).  The loop state lives in lane 0 and is broadcast where the other
   lanes are asked to help. */
__device__ void chain_fill (Prob &P, const ChainParams &prm, int *frontier, int lane) {
  int q = 0, nskipped = 0, min_hits = 1000000, specific_q = -1, specific_lo = 0, specific_hi = 0;
  int grand_score = 0, grand_q = -1, grand_h = -1;

  if (lane == 0) {
    for (q = 0; q < P.qs && q < P.L; q++) P.first[q] = -1;
    while (q <= P.qe && P.npos[q] <= 0) { P.first[q] = -1; q++; }
    if (q <= P.qe) { new_start(P,q); revise_active(P,q,0,P.npos[q]); }
  }
  q = __shfl_sync(0xffffffffu,q,0);
  while (q <= P.qe) {
    int lo = 0, hi = 0, next_q = 0, mode = 0;	/* mode 0: skipped, 1: single hit / none, 2: several hits */
    if (lane == 0) {
      const uint32_t *m = P.pos + P.cum[q];
      const int n = P.npos[q];
      const uint32_t lo_b = P.mina[q], hi_b = P.maxa[q];
      int h = 0;
      if (n > 8) {	/* sorted ascending: the two linear scans of stage2.c:3838-3846 as binary searches */
	int a = 0, b2 = n;
	while (a < b2) { const int mid = (a + b2) >> 1; if (m[mid] < lo_b) a = mid + 1; else b2 = mid; }
	lo = a; b2 = n;
	while (a < b2) { const int mid = (a + b2) >> 1; if (m[mid] <= hi_b) a = mid + 1; else b2 = mid; }
	hi = a;
      } else {
	while (h < n && m[h] < lo_b) h++;
	lo = h;
	while (h < n && m[h] <= hi_b) h++;
	hi = h;
      }
      if (P.skiprep && hi - lo >= MAX_NACTIVE && nskipped <= MAX_SKIPPED) {
	P.first[q] = -1;
	nskipped++;
	if (hi - lo < min_hits) { min_hits = hi - lo; specific_q = q; specific_lo = lo; specific_hi = hi; }
	mode = 0; next_q = q + 1;
      } else {
	if (nskipped > MAX_SKIPPED) { next_q = q; q = specific_q; lo = specific_lo; hi = specific_hi; }
	else next_q = q + 1;
	mode = (hi - lo > 1) ? 2 : 1;
      }
    }
    mode = __shfl_sync(0xffffffffu,mode,0);
    if (mode == 2) {
      q = __shfl_sync(0xffffffffu,q,0); lo = __shfl_sync(0xffffffffu,lo,0); hi = __shfl_sync(0xffffffffu,hi,0);
      score_mult(P,prm,q,lo,hi,frontier,lane);
      __syncwarp();
    }
    if (lane == 0 && mode != 0) {
      const int base = (int) P.cum[q], nhits = hi - lo;
      int best_s = 0, best_h = -1;
      if (nhits > 0) {
	if (nhits == 1) {
	  score_one(P,prm,q,lo);
	  if (P.score[base + lo] > 0) { best_s = P.score[base + lo]; best_h = lo; }
	} else {
	  for (int h = lo; h < hi; h++) if (P.score[base + h] > best_s) { best_s = P.score[base + h]; best_h = h; }
	}
	nskipped = 0; min_hits = 1000000; specific_q = -1;
	if (!P.middlep && best_h < 0) new_start(P,q);
	if (prm.splicingp && best_h >= 0 && P.phit[base + best_h] < 0 && grand_q >= 0 && q >= grand_q + P.k) {	/* :3966-3990 */
	  if ((best_s = P.score[P.cum[grand_q] + grand_h] - (q - grand_q)) > 0) {
	    const uint32_t pp = P.pos[P.cum[grand_q] + grand_h];
	    for (int h = lo; h < hi; h++) {
	      const uint32_t position = P.pos[base + h];
	      if (position > pp + prm.maxintronlen) {
	      } else if (position >= pp + P.k) {
		const int i = base + h;
		P.consec[i] = P.k; P.ppos[i] = grand_q; P.phit[i] = grand_h; P.trace[i] = ++P.tracei; P.score[i] = best_s;
	      }
	    }
	  }
	}
	if (best_h >= 0 && best_s >= grand_score && P.consec[base + best_h] > EXON_DEFN) {
	  grand_score = best_s; grand_q = q; grand_h = best_h;
	}
      }
      revise_active(P,q,lo,hi);
      if (P.npos[q] > 0) P.proc[P.nproc++] = q;
    }
    P.nproc = __shfl_sync(0xffffffffu,P.nproc,0);
    P.tracei = __shfl_sync(0xffffffffu,P.tracei,0);
    q = __shfl_sync(0xffffffffu,next_q,0);
    __syncwarp();
  }
}

/* ---- ranking (get_cells_fwd, stage2.c:3437) without the two sorts ----------------------------------------------
 * The reference keeps, per rootposition, the cells that share its best score, orders them by score (stable, so ties
 * stay in rootposition / querypos-descending / hit order) and traces a prefix.  Only cells within
 * FINAL_SCORE_TOLERANCE of the overall best can be in that prefix, and a same-root cell that beats one of them is
 * itself within the tolerance: compact those candidates, then repeatedly take the arg-min of the total order
 * (score desc, root asc, querypos desc, hit asc) and keep it unless its root was already seen with a higher score. */
__device__ __forceinline__ int qpos_of (const Prob &P, int idx) {	/* last q with cum[q] <= idx */
  int a = 0, b = P.L;
  while (b - a > 1) { const int mid = (a + b) >> 1; if ((int) P.cum[mid] <= idx) a = mid; else b = mid; }
  return a;
}

__device__ __forceinline__ bool cell_before (const Prob &P, unsigned long long ka, int ia, unsigned long long kb, int ib) {
  if (ka != kb) return ka < kb;
  if (ia == ib) return false;
  const int qa = qpos_of(P,ia), qb = qpos_of(P,ib);
  if (qa != qb) return qa > qb;
  return P.favor_right ? (ia > ib) : (ia < ib);
}

__device__ void rank_and_trace (const ChainDev &D, Prob &P, gmapchain_result &res, int *sh, int lane) {
  const int start = (int) P.cum[P.qs < P.L ? P.qs : P.L - 1];
  const int end = (P.qe + 1 < P.L) ? (int) P.cum[P.qe + 1] : P.tot;
  int *cand = P.next, *kept = P.trace;	/* dead after the fill */
  int B = 0;
  for (int i = start + lane; i < end; i += 32) B = max(B,P.score[i]);
  for (int o = 16; o > 0; o >>= 1) B = max(B,__shfl_xor_sync(0xffffffffu,B,o));
  res.status = 0; res.npaths = 0; res.bestscore = B; res.ncandidates = 0; res.path_off = 0;
  if (B <= 0 || P.qs > P.qe) { res.bestscore = B > 0 ? B : 0; return; }
  const int T = max(B - FINAL_SCORE_TOLERANCE,0);
  /* compaction (order does not matter: the selection below uses a total order) */
  int ncand = 0;
  for (int i0 = start; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool c = (i < end) && P.score[i] > T;
    const unsigned m = __ballot_sync(0xffffffffu,c);
    if (c) cand[start + ncand + __popc(m & ((1u << lane) - 1))] = i;	/* write index <= read index: in place is safe */
    ncand += __popc(m);
  }
  __syncwarp();
  res.ncandidates = ncand;
  int nk = 0;
  for (int p = 0; p < ncand; p++) {
    unsigned long long bk = ~0ull; int bi = -1, bslot = -1;
    for (int s = p + lane; s < ncand; s += 32) {
      const int idx = cand[start + s];
      const unsigned long long key = ((unsigned long long) (unsigned) (0x7fffffff - P.score[idx]) << 32) | ((unsigned) P.root[idx] ^ 0x80000000u);
      if (bi < 0 || cell_before(P,key,idx,bk,bi)) { bk = key; bi = idx; bslot = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu,bk,o);
      const int oi = __shfl_xor_sync(0xffffffffu,bi,o), os = __shfl_xor_sync(0xffffffffu,bslot,o);
      if (oi >= 0 && (bi < 0 || cell_before(P,ok,oi,bk,bi))) { bk = ok; bi = oi; bslot = os; }
    }
    if (lane == 0) { cand[start + bslot] = cand[start + p]; cand[start + p] = bi; }
    __syncwarp();
    const int s = P.score[bi], r = P.root[bi];
    if (nk >= P.max_nal && s < B) break;
    /* was this root seen?  (first appearance carries the root's best score) */
    int seen_score = -1;
    for (int j = lane; j < nk; j += 32) { const int kj = kept[start + j]; if (P.root[kj] == r) seen_score = P.score[kj]; }
    for (int o = 16; o > 0; o >>= 1) seen_score = max(seen_score,__shfl_xor_sync(0xffffffffu,seen_score,o));
    if (seen_score >= 0 && seen_score != s) continue;		/* not the best end of its root */
    if (!((nk < P.max_nal || s == B) && s > B - FINAL_SCORE_TOLERANCE)) break;
    if (lane == 0) kept[start + nk] = bi;
    nk++;
    __syncwarp();
  }
  /* path records, then one traceback per lane (traceback_one, stage2.c:4140) */
  unsigned long long poff = 0;
  if (lane == 0) poff = atomicAdd(&D.counters[1],(unsigned long long) nk);
  poff = __shfl_sync(0xffffffffu,poff,0);
  res.npaths = nk; res.path_off = (uint32_t) poff;
  for (int j = lane; j < nk; j += 32) {
    const int idx = kept[start + j];
    const int q0 = qpos_of(P,idx), h0 = idx - (int) P.cum[q0];
    int q = q0, h = h0, n = 0;
    while (q >= 0 && P.consec[P.cum[q] + h] < MIN_TERMINAL_NCONSECUTIVE) { const int i = (int) P.cum[q] + h; q = P.ppos[i]; h = P.phit[i]; }
    const int q1 = q, h1 = h;
    while (q >= 0) { const int i = (int) P.cum[q] + h; n++; q = P.ppos[i]; h = P.phit[i]; }
    const unsigned long long off = atomicAdd(&D.counters[2],(unsigned long long) n);
    if (poff + j < D.paths_cap) {
      gmapchain_path pr;
      pr.score = P.score[idx]; pr.rootposition = P.root[idx]; pr.endposition = (int) P.pos[idx]; pr.querypos = q0; pr.hit = h0;
      pr.npairs = n; pr.pair_off = (uint32_t) off; pr.reserved = 0;
      D.paths[poff + j] = pr;
    }
    if (off + n <= D.pairs_cap) {
      int2 *out = reinterpret_cast<int2 *>(D.pairs) + off;
      q = q1; h = h1;
      for (int t = 0; q >= 0; t++) { const int i = (int) P.cum[q] + h; out[t] = make_int2(q,(int) P.pos[i]); q = P.ppos[i]; h = P.phit[i]; }
    }
  }
  (void) sh;
}

__global__ void __launch_bounds__(CH_BLOCK) gmapchain_kernel (ChainDev D) {
  __shared__ int frontier_s[CH_WARPS][CH_FRONTIER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *frontier = frontier_s[warp];
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(&D.counters[0],1ull);
    slot = __shfl_sync(0xffffffffu,slot,0);
    if (slot >= (unsigned long long) D.nproblems) break;
    const int pi = D.order[slot];
    const gmapchain_problem pb = D.problems[pi];
    Prob P;
    P.npos = D.npos + pb.q_off; P.cum = D.cum + pb.q_off; P.mina = D.mina + pb.q_off; P.maxa = D.maxa + pb.q_off;
    P.first = D.first + pb.q_off; P.proc = D.proc + pb.q_off;
    P.pos = D.pos + pb.p_off; P.consec = D.consec + pb.p_off; P.root = D.root + pb.p_off; P.ppos = D.ppos + pb.p_off;
    P.phit = D.phit + pb.p_off; P.trace = D.trace + pb.p_off; P.score = D.score + pb.p_off; P.next = D.next + pb.p_off;
    P.L = pb.querylength; P.tot = pb.totalpositions; P.qs = pb.querystart; P.qe = pb.queryend; P.k = pb.indexsize;
    P.localp = (pb.flags & GMAPCHAIN_F_LOCALP) != 0; P.skiprep = (pb.flags & GMAPCHAIN_F_SKIP_REPETITIVE) != 0;
    P.favor_right = (pb.flags & GMAPCHAIN_F_FAVOR_RIGHT) != 0; P.middlep = (pb.flags & GMAPCHAIN_F_MIDDLEP) != 0;
    P.max_nal = pb.max_nalignments; P.nproc = 0; P.tracei = 0;
    gmapchain_result res;
    res.reserved[0] = res.reserved[1] = res.reserved[2] = 0;
    if (P.L <= 0) { res.status = 0; res.npaths = 0; res.bestscore = 0; res.ncandidates = 0; res.path_off = 0; }
    else {
      chain_fill(P,D.prm,frontier,lane);
      __syncwarp();
      __threadfence_block();
      rank_and_trace(D,P,res,frontier,lane);
    }
    if (lane == 0) D.results[pi] = res;
    __syncwarp();
  }
}

/* ---- host side ------------------------------------------------------------------------------------------------- */
struct ChainState {
  ChainParams prm;
  bool setup_done;
  cudaStream_t stream; cudaEvent_t ev0, ev1;
  int grid;
  /* device buffers */
  gmapchain_problem *d_problems; size_t cap_problems;
  int *d_order; size_t cap_order;
  int *d_npos; uint32_t *d_cum, *d_mina, *d_maxa; size_t cap_q;
  int *d_first, *d_proc;
  uint32_t *d_pos; size_t cap_p;
  int *d_hit[7];			/* consec, root, ppos, phit, trace, score, next */
  gmapchain_result *d_results; size_t cap_results;
  gmapchain_path *d_paths; size_t cap_paths;
  int *d_pairs; size_t cap_pairs;
  unsigned long long *d_counters;
  /* resident batch */
  int nproblems; size_t nq, np;
  unsigned long long used[3];
};

#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    *v.err = std::string(#call) + ": " + cudaGetErrorString(e_); return GMAPDP_ERR_CUDA; } } while (0)

static void chain_state_free (void *p) {
  ChainState *s = (ChainState *) p;
  cudaFree(s->d_problems); cudaFree(s->d_order); cudaFree(s->d_npos); cudaFree(s->d_cum); cudaFree(s->d_mina); cudaFree(s->d_maxa);
  cudaFree(s->d_first); cudaFree(s->d_proc); cudaFree(s->d_pos);
  for (int a = 0; a < 7; a++) cudaFree(s->d_hit[a]);
  cudaFree(s->d_results); cudaFree(s->d_paths); cudaFree(s->d_pairs); cudaFree(s->d_counters);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

static int chain_state (gmapdp_ctx *ctx, ChainState **out) {
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if (*v.chain == NULL) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      *v.err = "no CUDA device: the gmapchain engine has no CPU fallback";
      return GMAPDP_ERR_CUDA;
    }
    CKC(cudaSetDevice(v.device));
    cudaFuncAttributes fa;
    cudaError_t fe = cudaFuncGetAttributes(&fa,gmapchain_kernel);
    if (fe != cudaSuccess) { *v.err = std::string("no sm_100a image of gmapchain_kernel: ") + cudaGetErrorString(fe); return GMAPDP_ERR_CUDA; }
    ChainState *s = new ChainState();
    memset((void *) s,0,sizeof(*s));
    s->prm.splicingp = 1; s->prm.sufflookback = 60; s->prm.nsufflookback = 5; s->prm.maxintronlen = 500000;	/* gmap.c:269,270,347 */
    *v.chain = s; *v.chain_free = chain_state_free;
    CKC(cudaStreamCreateWithFlags(&s->stream,cudaStreamNonBlocking));
    CKC(cudaEventCreate(&s->ev0)); CKC(cudaEventCreate(&s->ev1));
    CKC(cudaMalloc((void **) &s->d_counters,4 * sizeof(unsigned long long)));
    int per_sm = 0;
    CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,gmapchain_kernel,CH_BLOCK,0));
    s->grid = v.sm_count * (per_sm > 0 ? per_sm : 1);
  }
  *out = (ChainState *) *v.chain;
  return GMAPDP_OK;
}

template <typename T>
static int growc (GdpCtxView &v, T **p, size_t *cap, size_t need) {
  if (need <= *cap && *p) return GMAPDP_OK;
  if (*p) CKC(cudaFree(*p));
  *p = NULL;
  const size_t n = need + need / 4 + 64;
  CKC(cudaMalloc((void **) p,n * sizeof(T)));
  *cap = n;
  return GMAPDP_OK;
}

extern "C" int gmapchain_setup (gmapdp_ctx *ctx, int splicingp, int cross_species_p, int sufflookback, int nsufflookback, int maxintronlen) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  if (cross_species_p) { *v.err = "gmapchain: cross_species_p (use_canonical_p) is not supported by this version"; return GMAPDP_ERR_ARG; }
  if (sufflookback < 0 || nsufflookback < 0 || sufflookback + 64 + 2 > CH_FRONTIER || nsufflookback + 2 > CH_FRONTIER) {
    *v.err = "gmapchain: sufflookback / nsufflookback out of range"; return GMAPDP_ERR_ARG;
  }
  s->prm.splicingp = splicingp ? 1 : 0; s->prm.sufflookback = sufflookback; s->prm.nsufflookback = nsufflookback; s->prm.maxintronlen = maxintronlen;
  s->setup_done = true;
  return GMAPDP_OK;
}

extern "C" int gmapchain_upload (gmapdp_ctx *ctx, const gmapchain_problem *problems, int nproblems,
				 const int32_t *npositions, const uint32_t *cumpositions, const uint32_t *minactive, const uint32_t *maxactive,
				 size_t nquerypos, const uint32_t *positions, size_t npositions_total) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  if (nproblems < 0) { *v.err = "gmapchain: negative problem count"; return GMAPDP_ERR_ARG; }
  for (int i = 0; i < nproblems; i++) {
    const gmapchain_problem &p = problems[i];
    if (p.querylength < 0 || p.totalpositions < 0 || p.q_off + (uint64_t) p.querylength > nquerypos ||
	p.p_off + (uint64_t) p.totalpositions > npositions_total || p.indexsize <= 0 || p.indexsize > 64 ||
	p.queryend >= p.querylength || p.querystart < 0) {
      *v.err = "gmapchain: problem " + std::to_string(i) + " is out of range of its pools"; return GMAPDP_ERR_ARG;
    }
  }
  CKC(cudaSetDevice(v.device));
  size_t cq = s->cap_q, cp = s->cap_p;
  if ((rc = growc(v,&s->d_problems,&s->cap_problems,(size_t) nproblems + 1))) return rc;
  if ((rc = growc(v,&s->d_order,&s->cap_order,(size_t) nproblems + 1))) return rc;
  if ((rc = growc(v,&s->d_results,&s->cap_results,(size_t) nproblems + 1))) return rc;
  if (nquerypos + 1 > cq || !s->d_npos) {
    size_t c;
    c = cq; if ((rc = growc(v,&s->d_npos,&c,nquerypos + 1))) return rc;
    c = cq; if ((rc = growc(v,&s->d_cum,&c,nquerypos + 1))) return rc;
    c = cq; if ((rc = growc(v,&s->d_mina,&c,nquerypos + 1))) return rc;
    c = cq; if ((rc = growc(v,&s->d_maxa,&c,nquerypos + 1))) return rc;
    c = cq; if ((rc = growc(v,&s->d_first,&c,nquerypos + 1))) return rc;
    c = cq; if ((rc = growc(v,&s->d_proc,&c,nquerypos + 1))) return rc;
    s->cap_q = c;
  }
  if (npositions_total + 1 > cp || !s->d_pos) {
    size_t c = cp;
    if ((rc = growc(v,&s->d_pos,&c,npositions_total + 1))) return rc;
    for (int a = 0; a < 7; a++) { c = cp; if ((rc = growc(v,&s->d_hit[a],&c,npositions_total + 1))) return rc; }
    s->cap_p = c;
  }
  /* longest first (LPT): work grows with the hits and with the query length */
  std::vector<int> order(nproblems);
  for (int i = 0; i < nproblems; i++) order[i] = i;
  std::stable_sort(order.begin(),order.end(),[&](int a, int b) {
    return (long) problems[a].totalpositions + problems[a].querylength > (long) problems[b].totalpositions + problems[b].querylength; });
  CKC(cudaMemcpyAsync(s->d_problems,problems,(size_t) nproblems * sizeof(gmapchain_problem),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_order,order.data(),(size_t) nproblems * sizeof(int),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_npos,npositions,nquerypos * sizeof(int),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_cum,cumpositions,nquerypos * sizeof(uint32_t),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_mina,minactive,nquerypos * sizeof(uint32_t),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_maxa,maxactive,nquerypos * sizeof(uint32_t),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaMemcpyAsync(s->d_pos,positions,npositions_total * sizeof(uint32_t),cudaMemcpyHostToDevice,s->stream));
  CKC(cudaStreamSynchronize(s->stream));	/* `order` is a local */
  s->nproblems = nproblems; s->nq = nquerypos; s->np = npositions_total;
  /* output pools: a first guess, grown by the caller on GMAPDP_ERR_CAPACITY */
  if ((rc = growc(v,&s->d_paths,&s->cap_paths,(size_t) nproblems * 2 + 64))) return rc;
  if ((rc = growc(v,&s->d_pairs,&s->cap_pairs,2 * (nquerypos + 64)))) return rc;
  return GMAPDP_OK;
}

extern "C" int gmapchain_run_resident (gmapdp_ctx *ctx, float *kernel_ms) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  CKC(cudaSetDevice(v.device));
  if (kernel_ms) *kernel_ms = 0.f;
  s->used[1] = s->used[2] = 0;
  if (s->nproblems == 0) return GMAPDP_OK;
  CKC(cudaEventRecord(s->ev0,s->stream));
  /* the link matrix starts zeroed (CALLOC in Linkmatrix_1d_new / intmatrix_1d_new, stage2.c:403 / :3112) */
  for (int a = 0; a < 7; a++) CKC(cudaMemsetAsync(s->d_hit[a],0,(s->np + 1) * sizeof(int),s->stream));
  CKC(cudaMemsetAsync(s->d_first,0,(s->nq + 1) * sizeof(int),s->stream));
  CKC(cudaMemsetAsync(s->d_counters,0,4 * sizeof(unsigned long long),s->stream));
  ChainDev D;
  D.problems = s->d_problems; D.order = s->d_order; D.nproblems = s->nproblems;
  D.npos = s->d_npos; D.cum = s->d_cum; D.mina = s->d_mina; D.maxa = s->d_maxa; D.pos = s->d_pos;
  D.consec = s->d_hit[0]; D.root = s->d_hit[1]; D.ppos = s->d_hit[2]; D.phit = s->d_hit[3]; D.trace = s->d_hit[4];
  D.score = s->d_hit[5]; D.next = s->d_hit[6];
  D.first = s->d_first; D.proc = s->d_proc;
  D.results = s->d_results; D.paths = s->d_paths; D.pairs = s->d_pairs;
  D.paths_cap = s->cap_paths; D.pairs_cap = s->cap_pairs / 2;
  D.counters = s->d_counters; D.prm = s->prm;
  const int grid = std::max(1,std::min(s->grid,(s->nproblems + CH_WARPS - 1) / CH_WARPS));
  gmapchain_kernel<<<grid,CH_BLOCK,0,s->stream>>>(D);
  CKC(cudaGetLastError());
  (*v.launches)++;
  CKC(cudaEventRecord(s->ev1,s->stream));
  CKC(cudaMemcpyAsync(s->used,s->d_counters,3 * sizeof(unsigned long long),cudaMemcpyDeviceToHost,s->stream));
  CKC(cudaStreamSynchronize(s->stream));
  if (kernel_ms) CKC(cudaEventElapsedTime(kernel_ms,s->ev0,s->ev1));
  if (s->used[1] > s->cap_paths || s->used[2] > s->cap_pairs / 2) {
    /* an output pool was too small: grow it and run again (the run is deterministic) */
    if (s->used[1] > s->cap_paths) { if ((rc = growc(v,&s->d_paths,&s->cap_paths,(size_t) s->used[1]))) return rc; }
    if (s->used[2] > s->cap_pairs / 2) { if ((rc = growc(v,&s->d_pairs,&s->cap_pairs,2 * (size_t) s->used[2]))) return rc; }
    return gmapchain_run_resident(ctx,kernel_ms);
  }
  return GMAPDP_OK;
}

extern "C" int gmapchain_download (gmapdp_ctx *ctx, gmapchain_result *results, gmapchain_path *paths, size_t paths_cap, size_t *paths_used,
				   int32_t *pairs, size_t pairs_cap, size_t *pairs_used) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  CKC(cudaSetDevice(v.device));
  if (paths_used) *paths_used = (size_t) s->used[1];
  if (pairs_used) *pairs_used = (size_t) s->used[2];
  if (s->used[1] > paths_cap || s->used[2] > pairs_cap) { *v.err = "gmapchain: output buffers too small"; return GMAPDP_ERR_CAPACITY; }
  if (s->nproblems) CKC(cudaMemcpyAsync(results,s->d_results,(size_t) s->nproblems * sizeof(gmapchain_result),cudaMemcpyDeviceToHost,s->stream));
  if (s->used[1]) CKC(cudaMemcpyAsync(paths,s->d_paths,(size_t) s->used[1] * sizeof(gmapchain_path),cudaMemcpyDeviceToHost,s->stream));
  if (s->used[2]) CKC(cudaMemcpyAsync(pairs,s->d_pairs,(size_t) s->used[2] * 2 * sizeof(int),cudaMemcpyDeviceToHost,s->stream));
  CKC(cudaStreamSynchronize(s->stream));
  return GMAPDP_OK;
}

extern "C" int gmapchain_run_batch (gmapdp_ctx *ctx, const gmapchain_problem *problems, int nproblems,
				    const int32_t *npositions, const uint32_t *cumpositions, const uint32_t *minactive, const uint32_t *maxactive,
				    size_t nquerypos, const uint32_t *positions, size_t npositions_total,
				    gmapchain_result *results, gmapchain_path *paths, size_t paths_cap, size_t *paths_used,
				    int32_t *pairs, size_t pairs_cap, size_t *pairs_used) {
  int rc;
  if ((rc = gmapchain_upload(ctx,problems,nproblems,npositions,cumpositions,minactive,maxactive,nquerypos,positions,npositions_total))) return rc;
  if ((rc = gmapchain_run_resident(ctx,NULL))) return rc;
  return gmapchain_download(ctx,results,paths,paths_cap,paths_used,pairs,pairs_cap,pairs_used);
}

extern "C" int gmapchain_download_links (gmapdp_ctx *ctx, int32_t *links, int32_t *scores, size_t npositions_total) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  if (npositions_total != s->np) { *v.err = "gmapchain: link download size mismatch"; return GMAPDP_ERR_ARG; }
  CKC(cudaSetDevice(v.device));
  std::vector<int> tmp(s->np + 1);
  for (int a = 0; a < 4; a++) {
    CKC(cudaMemcpy(tmp.data(),s->d_hit[a],s->np * sizeof(int),cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < s->np; i++) links[5 * i + a] = tmp[i];
  }
  for (size_t i = 0; i < s->np; i++) links[5 * i + 4] = 0;
  CKC(cudaMemcpy(scores,s->d_hit[5],s->np * sizeof(int),cudaMemcpyDeviceToHost));
  return GMAPDP_OK;
}
