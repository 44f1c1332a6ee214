/* Stage-2 chaining on the B200 (ABI in include/gmapchain_b200.h).
 *
 * What the reference computes (stage2.c:3667-4120, :1073-2020, :2956, :3437, :4140, :4402) is a chain DP over the k-mer
 * hits of one (query, genomic region): queryposes in ascending order, each hit choosing its best predecessor among
 * the active hits of the queryposes processed before it.  The recurrence is sequential in querypos AND in the hits of
 * one querypos (the frontier of score_querypos_lookback_mult is carried from hit to hit), so the outer parallel axis
 * is the problem: a persistent grid, ONE WARP PER PROBLEM pulled from a work-sorted queue, thousands in flight.
 *
 * Inside a problem the warp runs the recurrence in lock step (every lane holds the same scalars; same-address loads
 * are one broadcast transaction), and fans out where the reference has independent work:
 *   section D of a hit (stage2.c:1190-1425 / :1690-1920) looks at up to ~68 earlier queryposes.  Lane j takes the
 *   j-th of them: it walks that querypos's active list through ranges 1, 2 and 3+4 and keeps its first best
 *   candidate; the trace-label skip of "range 0" (the only thing that links the lists) is resolved beforehand from
 *   the list heads, and an ordered prefix-max over the lanes reproduces the strict ">" of the sequential scan and
 *   its ENOUGH_CONSECUTIVE early exit (lanes past the exit are discarded, their frontier left untouched).
 *   ranking: best-score reduction, compaction of the cells within FINAL_SCORE_TOLERANCE, repeated arg-min;
 *   tracebacks: one path per lane.
 *
 * State in HBM, two int4 per hit: hot {position, score, next-active, trace label} -- all a list walk touches, one
 * 16-byte load per candidate -- and cold {consecutive, rootposition, prev querypos, prev hit}; per querypos the
 * first-active hit and the stack of processed queryposes.
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
#define FULL 0xffffffffu

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
  int4 *hot, *cold;			/* per hit */
  int *cand, *kept;			/* per hit: ranking scratch */
  int *first, *proc;			/* per querypos */
  gmapchain_result *results; gmapchain_path *paths; int *pairs;
  unsigned long long paths_cap, pairs_cap;
  unsigned long long *counters;		/* [0] queue cursor, [1] path records used, [2] pairs used */
  ChainParams prm;
};

struct Prob {
  const int *npos; const uint32_t *cum, *mina, *maxa, *pos;
  int4 *hot, *cold; int *cand, *kept, *first, *proc;
  int L, tot, qs, qe, k, localp, skiprep, favor_right, middlep, max_nal;
  int nproc, tracei;
};

struct Best { int consec, root, score, pp, ph, trace; };

/* the lookforward twins (stage2.c:2020, :2404, :3034, :4610) are the same code with the query axis and the genomic
   comparisons mirrored: "pp lies more than x before position", in the direction of the walk (unsigned, as the reference) */
template <bool FWD> __device__ __forceinline__ bool before_lt (uint32_t pp, uint32_t x, uint32_t position) { return FWD ? (pp > position + x) : (pp + x < position); }
template <bool FWD> __device__ __forceinline__ bool before_le (uint32_t pp, uint32_t x, uint32_t position) { return FWD ? (pp >= position + x) : (pp + x <= position); }
template <bool FWD> __device__ __forceinline__ bool at_dist (uint32_t pp, uint32_t x, uint32_t position) { return FWD ? (pp == position + x) : (pp + x == position); }
template <bool FWD> __device__ __forceinline__ int gdist (uint32_t pp, uint32_t position) { return FWD ? (int) (pp - position) : (int) (position - pp); }
template <bool FWD> __device__ __forceinline__ int qdist (int q, int pq) { return FWD ? pq - q : q - pq; }

#define H_POS(v) ((uint32_t) (v).x)
#define H_SCORE(v) ((v).y)
#define H_NEXT(v) ((v).z)
#define H_TRACE(v) ((v).w)

/* Everything below is executed by all 32 lanes with identical arguments unless it says otherwise; stores are
   issued by every lane with the same value (one transaction), so each lane always reads back its own writes. */

__device__ __forceinline__ void commit (Prob &P, int i, uint32_t position, const Best &b) {	/* stage2.c:1428-1450 */
  int trace, score;
  if (b.pp >= 0) { trace = b.trace; score = b.score; }
  else if (P.localp) { trace = ++P.tracei; score = P.k; }
  else { trace = ++P.tracei; score = b.score; }
  P.cold[i] = make_int4(b.consec,b.root,b.pp,b.ph);
  P.hot[i] = make_int4((int) position,score,0,trace);
}

/* Section D for one hit at `position` of querypos q: the entries ns = 0..nlast of the processed stack (newest first).
   frontier != NULL (the several-hits case) carries each entry's list position from hit to hit and applies range 1
   unconditionally (stage2.c:1722); frontier == NULL starts from first[] and applies it only when splicing (:1217). */
template <bool FWD> __device__ void section_d (Prob &P, const ChainParams &prm, Best &b, int q, uint32_t position, int nlast, int *frontier, int lane) {
  int last_trace = -1;
  for (int ns0 = 0; ns0 <= nlast && b.consec < ENOUGH_CONSECUTIVE; ns0 += 32) {
    const int ns = ns0 + lane;
    const bool valid = ns <= nlast;
    int pq = 0, qd = 0, o = 0, head = -1;
    int4 hv = make_int4(0,0,-1,0);
    if (valid) {
      pq = P.proc[P.nproc - 1 - ns]; qd = qdist<FWD>(q,pq); o = (int) P.cum[pq];
      head = frontier ? frontier[ns] : P.first[pq];
      if (head != -1) hv = P.hot[o + head];
    }
    /* range 0, in stack order: a list skips its leading hits whose label equals the label the previous non-empty
       list started with (stage2.c:1209-1214).  With a = label of a list's head and b = label of its second hit (a if
       there is none), a list maps the incoming label lt to (lt == a ? b : a): whatever comes in, what goes out is a or
       b, so along the stack the state is ONE BIT per list ("left on its second hit") and the lists compose as 2-state
       maps -- a 5-step prefix scan instead of a 32-step walk.  Lists whose first two hits share a label (then a third
       hit could matter) take the literal walk below. */
    int start = -1; int4 sv = hv;
    const unsigned havemask = __ballot_sync(FULL,head != -1);
    int4 h2 = make_int4(0,0,-1,0);
    const bool second = (head != -1) && (H_NEXT(hv) != -1);
    if (second) h2 = P.hot[o + H_NEXT(hv)];
    const bool odd = second && (H_TRACE(h2) == H_TRACE(hv));
    if (havemask == 0) {
    } else if (!__any_sync(FULL,odd)) {
      const int ta = H_TRACE(hv), tb = second ? H_TRACE(h2) : H_TRACE(hv);
      const unsigned below = havemask & ((1u << lane) - 1);
      const int prev = below ? 31 - __clz(below) : 0;
      const int pa = __shfl_sync(FULL,ta,prev), pb = __shfl_sync(FULL,tb,prev);
      unsigned map = 2u;				/* identity: bit s = image of state s */
      if (head != -1) {
	if (!below) map = (last_trace == ta) ? 3u : 0u;	/* first non-empty list of the chunk: the carried label decides */
	else map = (unsigned) (pa == ta) | ((unsigned) (pb == ta) << 1);
      }
      for (int d = 1; d < 32; d <<= 1) {
	const unsigned g = __shfl_up_sync(FULL,map,d);
	if (lane >= d) map = ((map >> (g & 1)) & 1) | (((map >> ((g >> 1) & 1)) & 1) << 1);
      }
      const int st = (int) (map & 1);			/* constant map from the first non-empty list on */
      if (head != -1) {
	if (!st) { start = head; sv = hv; }
	else if (second) { start = H_NEXT(hv); sv = h2; }
      }
      const int lastl = 31 - __clz(havemask);
      last_trace = __shfl_sync(FULL,st ? tb : ta,lastl);
    } else {
      for (unsigned m = havemask; m; m &= m - 1) {
	const int j = __ffs(m) - 1;
	int cur = __shfl_sync(FULL,head,j);
	const int oj = __shfl_sync(FULL,o,j);
	int4 cv;
	cv.x = __shfl_sync(FULL,hv.x,j); cv.y = __shfl_sync(FULL,hv.y,j); cv.z = __shfl_sync(FULL,hv.z,j); cv.w = __shfl_sync(FULL,hv.w,j);
	while (cur != -1 && H_TRACE(cv) == last_trace) {
	  cur = H_NEXT(cv);
	  if (cur != -1) cv = P.hot[oj + cur];
	}
	if (cur != -1) last_trace = H_TRACE(cv);
	if (lane == j) { start = cur; sv = cv; }
      }
    }
    /* ranges 1, 2, 3+4: each lane on its own list */
    int ph = start, newfront = start;
    int m_score = INT_MIN, m_ph = -1, m_kind = 0, m_same = 0;
    if (ph != -1) {
      const int credit = -qd / P.k;
      if (frontier || prm.splicingp) {
	while (ph != -1 && before_le<FWD>(H_POS(sv),(uint32_t) (prm.maxintronlen + qd),position)) { ph = H_NEXT(sv); if (ph != -1) sv = P.hot[o + ph]; }
      }
      newfront = ph;
      while (ph != -1 && before_lt<FWD>(H_POS(sv),(uint32_t) (EQUAL_DISTANCE_NOT_SPLICING + qd),position)) {	/* range 2 */
	const int diff = gdist<FWD>(H_POS(sv),position) - qd;
	int s = H_SCORE(sv) + credit;
	s -= prm.splicingp ? (diff / TEN_THOUSAND + 1) : (diff + 1);
	if (s > m_score) { m_score = s; m_ph = ph; m_kind = 2; m_same = 0; }
	ph = H_NEXT(sv); if (ph != -1) sv = P.hot[o + ph];
      }
      while (ph != -1 && before_le<FWD>(H_POS(sv),(uint32_t) P.k,position)) {				/* ranges 3+4 */
	const int s = H_SCORE(sv) + 1;
	if (s > m_score) { m_score = s; m_ph = ph; m_kind = 4; m_same = (gdist<FWD>(H_POS(sv),position) == qd); }
	ph = H_NEXT(sv); if (ph != -1) sv = P.hot[o + ph];
      }
    }
    /* consecutive count the hit would get from this lane's candidate (diff <= EQUAL_DISTANCE_FOR_CONSECUTIVE == 0) */
    int m_consec = 0;
    if (m_ph != -1 && m_same) m_consec = P.cold[o + m_ph].x + qd;
    /* ordered prefix max: highest score, earliest lane */
    long long key = (m_ph != -1) ? ((long long) m_score * 256 + (long long) (31 - lane)) : LLONG_MIN;
    for (int d = 1; d < 32; d <<= 1) {
      const long long up = __shfl_up_sync(FULL,key,d);
      if (lane >= d && up > key) key = up;
    }
    const bool beats = (key != LLONG_MIN) && ((int) (key >> 8) > b.score);
    const int wl = 31 - (int) (key & 0xff);			/* lane holding the running best, if it beats b */
    const int wc = __shfl_sync(FULL,m_consec,beats ? wl : 0);
    const int consec_after = beats ? wc : b.consec;
    const unsigned stopmask = __ballot_sync(FULL,valid && consec_after >= ENOUGH_CONSECUTIVE);
    const unsigned validmask = __ballot_sync(FULL,valid);
    const int jlast = 31 - __clz(validmask);
    const int jstop = stopmask ? (__ffs(stopmask) - 1) : jlast;
    if (frontier) {
      if (valid && lane <= jstop && head != -1) frontier[ns] = newfront;
      __syncwarp();
    }
    const int fbeats = __shfl_sync(FULL,(int) beats,jstop);
    if (fbeats) {
      const int w = __shfl_sync(FULL,wl,jstop);
      const int wo = __shfl_sync(FULL,o,w), wph = __shfl_sync(FULL,m_ph,w), wkind = __shfl_sync(FULL,m_kind,w);
      b.score = __shfl_sync(FULL,m_score,w); b.consec = __shfl_sync(FULL,m_consec,w);
      b.pp = __shfl_sync(FULL,pq,w); b.ph = wph;
      b.root = P.cold[wo + wph].y;
      b.trace = (wkind == 2) ? ++P.tracei : H_TRACE(P.hot[wo + wph]);
    }
    if (stopmask) break;
  }
}

/* how many leading entries of the processed stack satisfy (ns <= a || qd - k <= b): the distance grows along the
   stack (queryposes are processed in ascending order), so the test of stage2.c:1204 / :1626-1631 holds on a prefix */
template <bool FWD> __device__ int reach (const Prob &P, int q, int a, int b, int lane) {
  int n = 0;
  for (int s0 = 0; s0 < P.nproc; s0 += 32) {
    const int ns = s0 + lane;
    bool ok = false;
    if (ns < P.nproc) { const int qd = qdist<FWD>(q,P.proc[P.nproc - 1 - ns]); ok = (ns <= a || qd - P.k <= b); }
    const unsigned m = __ballot_sync(FULL,ok);
    n += __popc(m);
    if (m != FULL) break;
  }
  return n;
}

/* adjacent hit on the newest processed querypos: first active hit with pos + qd >= position, from `ph` on */
template <bool FWD> __device__ __forceinline__ int adjacent (const Prob &P, int o, int qd, uint32_t position, int ph, int4 &v, bool &found) {
  uint32_t pp = position;
  while (ph != -1) {
    v = P.hot[o + ph];
    pp = H_POS(v);
    if (before_lt<FWD>(pp,(uint32_t) qd,position)) ph = H_NEXT(v); else break;
  }
  found = at_dist<FWD>(pp,(uint32_t) qd,position);
  return ph;
}

/* score_querypos_lookback_one (stage2.c:1073) */
template <bool FWD> __device__ void score_one (Prob &P, const ChainParams &prm, int q, int h, int lane) {
  const int i = (int) P.cum[q] + h;
  const uint32_t position = P.pos[i];
  Best b = { P.k, (int) position, 0, -1, -1, 0 };
  int nlookback = prm.nsufflookback, lookback = prm.sufflookback;
  if (P.nproc > 0) {
    const int pq = P.proc[P.nproc - 1], o = (int) P.cum[pq], qd = qdist<FWD>(q,pq);
    int4 v; bool found;
    const int ph = adjacent<FWD>(P,o,qd,position,P.first[pq],v,found);
    if (found) {
      const int4 c = P.cold[o + ph];
      b.consec = c.x + qd; b.root = c.y; b.score = H_SCORE(v) + qd; b.pp = pq; b.ph = ph; b.trace = H_TRACE(v);
      nlookback = 1; lookback = prm.sufflookback / 2;
    }
    if (b.consec < ENOUGH_CONSECUTIVE) {
      /* the loop of :1196 runs through the first entry that fails the reach test, inclusive */
      const int n = reach<FWD>(P,q,nlookback,lookback,lane);
      section_d<FWD>(P,prm,b,q,position,min(n,P.nproc - 1),NULL,lane);
    }
  }
  commit(P,i,position,b);
}

/* score_querypos_lookback_mult (stage2.c:1470) */
template <bool FWD> __device__ void score_mult (Prob &P, const ChainParams &prm, int q, int lo, int hi, int *frontier, int lane) {
  const int base = (int) P.cum[q];
  if (P.nproc == 0) {
    for (int h = lo; h < hi; h++) {
      const int i = base + h;
      const uint32_t position = P.pos[i];
      int4 hv = P.hot[i];
      P.cold[i] = make_int4(P.k,(int) position,-1,-1);
      hv.x = (int) position;
      if (P.localp) { hv.w = ++P.tracei; hv.y = P.k; } else hv.y = 0;
      P.hot[i] = hv;
    }
    return;
  }
  const int max_adj = reach<FWD>(P,q,1,prm.sufflookback / 2,lane) - 1;
  const int max_nonadj = reach<FWD>(P,q,prm.nsufflookback,prm.sufflookback,lane) - 1;
  __syncwarp();
  for (int ns = lane; ns <= max_nonadj && ns < CH_FRONTIER; ns += 32) frontier[ns] = P.first[P.proc[P.nproc - 1 - ns]];
  __syncwarp();
  const int adjq = P.proc[P.nproc - 1], adjo = (int) P.cum[adjq], adjqd = qdist<FWD>(q,adjq);
  const int h0 = FWD ? hi - 1 : lo, hstep = FWD ? -1 : 1;	/* hits are walked away from the processed side */
  int overall = 0, adjf = P.first[adjq];
  for (int h = h0, cnt = lo; cnt < hi; h += hstep, cnt++) {
    int4 v; bool found;
    adjf = adjacent<FWD>(P,adjo,adjqd,P.pos[base + h],adjf,v,found);
    if (found) overall = max(overall,P.cold[adjo + adjf].x + adjqd);
  }
  adjf = P.first[adjq];
  for (int h = h0, cnt = lo; cnt < hi; h += hstep, cnt++) {
    const uint32_t position = P.pos[base + h];
    Best b;
    int4 v; bool found;
    int max_nseen;
    adjf = adjacent<FWD>(P,adjo,adjqd,position,adjf,v,found);
    if (found) {
      const int4 c = P.cold[adjo + adjf];
      b.consec = c.x + adjqd; b.root = c.y; b.pp = adjq; b.ph = adjf; b.score = H_SCORE(v) + adjqd; b.trace = H_TRACE(v);
      max_nseen = max_adj;
    } else {
      b.consec = P.k; b.root = (int) position; b.pp = -1; b.ph = -1; b.score = 0; b.trace = -1;
      max_nseen = max_nonadj;
    }
    if (overall < GREEDY_NCONSECUTIVE && b.consec < ENOUGH_CONSECUTIVE) section_d<FWD>(P,prm,b,q,position,max_nseen,frontier,lane);
    commit(P,base + h,position,b);
  }
}

/* revise_active_lookback (stage2.c:2956): the hits within SCORE_FOR_RESTRICT of the querypos's best, linked in order */
template <bool FWD> __device__ void revise_active (Prob &P, int q, int lo, int hi) {
  const int base = (int) P.cum[q];
  if (lo >= hi) { P.first[q] = -1; return; }
  int best = P.hot[base + lo].y;
  for (int h = lo + 1; h < hi; h++) best = max(best,P.hot[base + h].y);
  const int thr = max(best - SCORE_FOR_RESTRICT,0);
  int prev = -1, firsth = -1;
  for (int h = FWD ? hi - 1 : lo, cnt = lo; cnt < hi; h += FWD ? -1 : 1, cnt++) {
    if (P.hot[base + h].y > thr) {
      if (prev < 0) firsth = h; else P.hot[base + prev].z = h;
      prev = h;
    }
  }
  if (prev >= 0) P.hot[base + prev].z = -1;
  P.first[q] = firsth;
}

__device__ void new_start (Prob &P, int q) {	/* stage2.c:3793-3812, :3941-3964 (rootposition is left alone) */
  const int base = (int) P.cum[q];
  for (int h = 0; h < P.npos[q]; h++) {
    const int i = base + h;
    int4 c = P.cold[i], v = P.hot[i];
    c.x = P.k; c.z = -1; c.w = -1;
    v.x = (int) P.pos[i]; v.y = P.k; v.w = -1;
    P.cold[i] = c; P.hot[i] = v;
  }
}

/* align_compute_scores_lookback (stage2.c:3667) / align_compute_scores_lookforward (:4610) */
template <bool FWD> __device__ void chain_fill (Prob &P, const ChainParams &prm, int *frontier, int lane) {
  const int step = FWD ? -1 : 1;
  int q, nskipped = 0, min_hits = 1000000, specific_q = -1, specific_lo = 0, specific_hi = 0;
  int grand_score = 0, grand_q = -1, grand_h = -1;
#define INRANGE(q) (FWD ? (q) >= P.qs : (q) <= P.qe)

  if (FWD) { for (q = P.L - 1 - lane; q > P.qe; q -= 32) P.first[q] = -1; q = P.qe < P.L - 1 ? P.qe : P.L - 1; }
  else { for (q = lane; q < P.qs && q < P.L; q += 32) P.first[q] = -1; q = min(P.qs,P.L); }
  while (INRANGE(q) && P.npos[q] <= 0) { P.first[q] = -1; q += step; }
  if (INRANGE(q)) { new_start(P,q); revise_active<FWD>(P,q,0,P.npos[q]); }
  while (INRANGE(q)) {
    const uint32_t *m = P.pos + P.cum[q];
    const int n = P.npos[q];
    const uint32_t lo_b = P.mina[q], hi_b = P.maxa[q];
    int lo, hi, next_q;
    if (n > 8 && lo_b <= hi_b) {	/* sorted ascending: the linear scans of stage2.c:3838-3846 / :4781-4790 as binary searches */
      int a = 0, b2 = n;
      while (a < b2) { const int mid = (a + b2) >> 1; if (m[mid] < lo_b) a = mid + 1; else b2 = mid; }
      lo = a; b2 = n;
      while (a < b2) { const int mid = (a + b2) >> 1; if (m[mid] <= hi_b) a = mid + 1; else b2 = mid; }
      hi = a;
    } else if (FWD) {
      int h = n - 1;
      while (h >= 0 && m[h] > hi_b) h--;
      hi = h + 1;
      while (h >= 0 && m[h] >= lo_b) h--;
      lo = h + 1;
    } else {
      int h = 0;
      while (h < n && m[h] < lo_b) h++;
      lo = h;
      while (h < n && m[h] <= hi_b) h++;
      hi = h;
    }
    if (P.skiprep && hi - lo >= MAX_NACTIVE && nskipped <= MAX_SKIPPED) {
      P.first[q] = -1;
      nskipped++;
      if (hi - lo < min_hits) { min_hits = hi - lo; specific_q = q; specific_lo = lo; specific_hi = hi; }
      q += step;
      continue;
    }
    if (nskipped > MAX_SKIPPED) { next_q = q; q = specific_q; lo = specific_lo; hi = specific_hi; }
    else next_q = q + step;
    const int base = (int) P.cum[q], nhits = hi - lo;
    if (nhits > 0) {
      int best_s = 0, best_h = -1;
      if (nhits == 1) {
	score_one<FWD>(P,prm,q,lo,lane);
	const int s = P.hot[base + lo].y;
	if (s > 0) { best_s = s; best_h = lo; }
      } else {
	score_mult<FWD>(P,prm,q,lo,hi,frontier,lane);
	for (int h = FWD ? hi - 1 : lo, cnt = lo; cnt < hi; h += step, cnt++) {
	  const int s = P.hot[base + h].y;
	  if (s > best_s) { best_s = s; best_h = h; }
	}
      }
      nskipped = 0; min_hits = 1000000; specific_q = -1;
      if (!P.middlep && best_h < 0) new_start(P,q);
      if (prm.splicingp && best_h >= 0 && P.cold[base + best_h].w < 0 &&
	  (FWD ? (grand_q <= P.L - P.k && q + P.k <= grand_q) : (grand_q >= 0 && q >= grand_q + P.k))) {	/* :3966-3990 / :4905-4930 */
	const int4 gv = P.hot[P.cum[grand_q] + grand_h];
	if ((best_s = H_SCORE(gv) - qdist<FWD>(q,grand_q)) > 0) {
	  const uint32_t pp = H_POS(gv);
	  for (int h = lo; h < hi; h++) {
	    const uint32_t position = P.pos[base + h];
	    if (FWD ? (position + prm.maxintronlen < pp) : (position > pp + prm.maxintronlen)) {
	    } else if (FWD ? (position + P.k <= pp) : (position >= pp + P.k)) {
	      const int i = base + h;
	      int4 c = P.cold[i], v = P.hot[i];
	      c.x = P.k; c.z = grand_q; c.w = grand_h;
	      v.y = best_s; v.w = ++P.tracei;
	      P.cold[i] = c; P.hot[i] = v;
	    }
	  }
	}
      }
      if (best_h >= 0 && best_s >= grand_score && P.cold[base + best_h].x > EXON_DEFN) {
	grand_score = best_s; grand_q = q; grand_h = best_h;
      }
    }
    revise_active<FWD>(P,q,lo,hi);
    if (P.npos[q] > 0) P.proc[P.nproc++] = q;
    q = next_q;
  }
#undef INRANGE
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

__device__ void rank_and_trace (const ChainDev &D, Prob &P, gmapchain_result &res, int lane) {
  int *cand = P.cand, *kept = P.kept;
  res.status = 0; res.npaths = 0; res.bestscore = 0; res.ncandidates = 0; res.path_off = 0;
  if (P.qs > P.qe) return;
  const int start = (int) P.cum[P.qs];
  const int end = (P.qe + 1 < P.L) ? (int) P.cum[P.qe + 1] : P.tot;
  int B = 0;
  for (int i = start + lane; i < end; i += 32) B = max(B,P.hot[i].y);
  for (int o = 16; o > 0; o >>= 1) B = max(B,__shfl_xor_sync(FULL,B,o));
  if (B <= 0) return;
  res.bestscore = B;
  const int T = max(B - FINAL_SCORE_TOLERANCE,0);
  int ncand = 0;
  for (int i0 = start; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool c = (i < end) && P.hot[i].y > T;
    const unsigned m = __ballot_sync(FULL,c);
    if (c) cand[ncand + __popc(m & ((1u << lane) - 1))] = i;
    ncand += __popc(m);
  }
  __syncwarp();
  res.ncandidates = ncand;
  int nk = 0;
  for (int p = 0; p < ncand; p++) {
    unsigned long long bk = ~0ull; int bi = -1, bslot = -1;
    for (int s = p + lane; s < ncand; s += 32) {
      const int idx = cand[s];
      const unsigned long long key = ((unsigned long long) (unsigned) (0x7fffffff - P.hot[idx].y) << 32) | ((unsigned) P.cold[idx].y ^ 0x80000000u);
      if (bi < 0 || cell_before(P,key,idx,bk,bi)) { bk = key; bi = idx; bslot = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(FULL,bk,o);
      const int oi = __shfl_xor_sync(FULL,bi,o), os = __shfl_xor_sync(FULL,bslot,o);
      if (oi >= 0 && (bi < 0 || cell_before(P,ok,oi,bk,bi))) { bk = ok; bi = oi; bslot = os; }
    }
    if (lane == 0) { cand[bslot] = cand[p]; cand[p] = bi; }
    __syncwarp();
    const int s = P.hot[bi].y, r = P.cold[bi].y;
    if (nk >= P.max_nal && s < B) break;
    int seen_score = -1;		/* was this root seen?  (its first appearance carries the root's best score) */
    for (int j = lane; j < nk; j += 32) { const int kj = kept[j]; if (P.cold[kj].y == r) seen_score = P.hot[kj].y; }
    for (int o = 16; o > 0; o >>= 1) seen_score = max(seen_score,__shfl_xor_sync(FULL,seen_score,o));
    if (seen_score >= 0 && seen_score != s) continue;		/* not the best end of its root */
    if (!((nk < P.max_nal || s == B) && s > B - FINAL_SCORE_TOLERANCE)) break;
    if (lane == 0) kept[nk] = bi;
    nk++;
    __syncwarp();
  }
  /* path records, then one traceback per lane (traceback_one, stage2.c:4140) */
  unsigned long long poff = 0;
  if (lane == 0) poff = atomicAdd(&D.counters[1],(unsigned long long) nk);
  poff = __shfl_sync(FULL,poff,0);
  res.npaths = nk; res.path_off = (uint32_t) poff;
  for (int j = lane; j < nk; j += 32) {
    const int idx = kept[j];
    const int q0 = qpos_of(P,idx), h0 = idx - (int) P.cum[q0];
    int q = q0, h = h0, n = 0;
    while (q >= 0) {
      const int4 c = P.cold[P.cum[q] + h];
      if (c.x >= MIN_TERMINAL_NCONSECUTIVE) break;
      q = c.z; h = c.w;
    }
    const int q1 = q, h1 = h;
    while (q >= 0) { const int4 c = P.cold[P.cum[q] + h]; n++; q = c.z; h = c.w; }
    const unsigned long long off = atomicAdd(&D.counters[2],(unsigned long long) n);
    if (poff + j < D.paths_cap) {
      gmapchain_path pr;
      pr.score = P.hot[idx].y; pr.rootposition = P.cold[idx].y; pr.endposition = (int) P.pos[idx]; pr.querypos = q0; pr.hit = h0;
      pr.npairs = n; pr.pair_off = (uint32_t) off; pr.reserved = 0;
      D.paths[poff + j] = pr;
    }
    if (off + n <= D.pairs_cap) {
      int2 *out = reinterpret_cast<int2 *>(D.pairs) + off;
      q = q1; h = h1;
      for (int t = 0; q >= 0; t++) {
	const int i = (int) P.cum[q] + h;
	const int4 c = P.cold[i];
	out[t] = make_int2(q,(int) P.pos[i]);
	q = c.z; h = c.w;
      }
    }
  }
}

#ifndef CH_MINB
#define CH_MINB 8
#endif
__global__ void __launch_bounds__(CH_BLOCK,CH_MINB) gmapchain_kernel (ChainDev D) {
  __shared__ int frontier_s[CH_WARPS][CH_FRONTIER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int *frontier = frontier_s[warp];
  for (;;) {
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(&D.counters[0],1ull);
    slot = __shfl_sync(FULL,slot,0);
    if (slot >= (unsigned long long) D.nproblems) break;
    const int pi = D.order[slot];
    const gmapchain_problem pb = D.problems[pi];
    Prob P;
    P.npos = D.npos + pb.q_off; P.cum = D.cum + pb.q_off; P.mina = D.mina + pb.q_off; P.maxa = D.maxa + pb.q_off;
    P.first = D.first + pb.q_off; P.proc = D.proc + pb.q_off;
    P.pos = D.pos + pb.p_off; P.hot = D.hot + pb.p_off; P.cold = D.cold + pb.p_off; P.cand = D.cand + pb.p_off; P.kept = D.kept + pb.p_off;
    P.L = pb.querylength; P.tot = pb.totalpositions; P.qs = pb.querystart; P.qe = pb.queryend; P.k = pb.indexsize;
    P.localp = (pb.flags & GMAPCHAIN_F_LOCALP) != 0; P.skiprep = (pb.flags & GMAPCHAIN_F_SKIP_REPETITIVE) != 0;
    P.favor_right = (pb.flags & GMAPCHAIN_F_FAVOR_RIGHT) != 0; P.middlep = (pb.flags & GMAPCHAIN_F_MIDDLEP) != 0;
    P.max_nal = pb.max_nalignments; P.nproc = 0; P.tracei = 0;
    gmapchain_result res;
    res.status = 0; res.npaths = 0; res.bestscore = 0; res.ncandidates = 0; res.path_off = 0;
    res.reserved[0] = res.reserved[1] = res.reserved[2] = 0;
    if (P.L > 0) {
      if (pb.flags & GMAPCHAIN_F_LOOKFORWARD) chain_fill<true>(P,D.prm,frontier,lane);
      else chain_fill<false>(P,D.prm,frontier,lane);
      __syncwarp();
      rank_and_trace(D,P,res,lane);
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
  int4 *d_hot, *d_cold;			/* two int4 per hit */
  int *d_cand, *d_kept;
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
  cudaFree(s->d_hot); cudaFree(s->d_cold); cudaFree(s->d_cand); cudaFree(s->d_kept);
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
  if (*p) v.grave->push_back((void *) *p);	/* freed with the context: cudaFree would synchronise the whole device */
  *p = NULL;
  const size_t n = need + need / 2 + 64;
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

/* device buffers for batches of up to nproblems problems, nquerypos query positions and npositions_total hits;
   callers that run many small batches (the drop-in) reserve once, so that no buffer is re-allocated in the middle of a run */
extern "C" int gmapchain_reserve (gmapdp_ctx *ctx, int nproblems, size_t nquerypos, size_t npositions_total) {
  ChainState *s; int rc;
  GdpCtxView v = gmapdp_ctx_view(ctx);
  if ((rc = chain_state(ctx,&s)) != GMAPDP_OK) return rc;
  if (nproblems < 0) { *v.err = "gmapchain: negative problem count"; return GMAPDP_ERR_ARG; }
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
    c = cp; if ((rc = growc(v,&s->d_hot,&c,npositions_total + 1))) return rc;
    c = cp; if ((rc = growc(v,&s->d_cold,&c,npositions_total + 1))) return rc;
    c = cp; if ((rc = growc(v,&s->d_cand,&c,npositions_total + 1))) return rc;
    c = cp; if ((rc = growc(v,&s->d_kept,&c,npositions_total + 1))) return rc;
    s->cap_p = c;
  }
  /* output pools: a first guess, grown on GMAPDP_ERR_CAPACITY */
  if ((rc = growc(v,&s->d_paths,&s->cap_paths,(size_t) nproblems * 2 + 64))) return rc;
  if ((rc = growc(v,&s->d_pairs,&s->cap_pairs,2 * (nquerypos + 64)))) return rc;
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
  if ((rc = gmapchain_reserve(ctx,nproblems,nquerypos,npositions_total))) return rc;
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
  CKC(cudaMemsetAsync(s->d_hot,0,(s->np + 1) * sizeof(int4),s->stream));
  CKC(cudaMemsetAsync(s->d_cold,0,(s->np + 1) * sizeof(int4),s->stream));
  CKC(cudaMemsetAsync(s->d_counters,0,4 * sizeof(unsigned long long),s->stream));
  ChainDev D;
  D.problems = s->d_problems; D.order = s->d_order; D.nproblems = s->nproblems;
  D.npos = s->d_npos; D.cum = s->d_cum; D.mina = s->d_mina; D.maxa = s->d_maxa; D.pos = s->d_pos;
  D.hot = s->d_hot; D.cold = s->d_cold; D.cand = s->d_cand; D.kept = s->d_kept;
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
  std::vector<int4> tmp(s->np + 1);
  CKC(cudaMemcpy(tmp.data(),s->d_cold,s->np * sizeof(int4),cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < s->np; i++) {
    links[5 * i + 0] = tmp[i].x; links[5 * i + 1] = tmp[i].y; links[5 * i + 2] = tmp[i].z; links[5 * i + 3] = tmp[i].w; links[5 * i + 4] = 0;
  }
  CKC(cudaMemcpy(tmp.data(),s->d_hot,s->np * sizeof(int4),cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < s->np; i++) scores[i] = tmp[i].y;
  return GMAPDP_OK;
}
