/* TEST INFRASTRUCTURE ONLY -- never linked into the product (only tests/, smoke() and bench.py's CPU leg).
 *
 * Scalar restatement of the reference's stage-2 chaining.  The lookforward twins
 * (align_compute_scores_lookforward stage2.c:4610, score_querypos_lookforward_one :2020, _mult :2404,
 * revise_active_lookforward :3034, align_compute_lookforward :5062) are the same code with the query axis and the
 * genomic comparisons mirrored (c->fwd below); ranking and traceback are shared.  Lookback direction:
 *   align_compute_scores_lookback   /root/reference/src/stage2.c:3667
 *   score_querypos_lookback_one     stage2.c:1073
 *   score_querypos_lookback_mult    stage2.c:1470
 *   revise_active_lookback          stage2.c:2956
 *   get_cells_fwd                   stage2.c:3437  (+ Cell_rootposition_left/right_cmp :3230/:3277, Cell_score_cmp :3323)
 *   align_compute_lookback          stage2.c:4402  (path selection) and traceback_one :4140
 * for the non-PMAP build with use_canonical_p == false (gmap's default: cross_species_p false, stage2.c:141).
 * Flat arrays over the CSR of hits instead of the reference's pointer matrices and Intlist.
 * Parity PINNED: tests/test_chain_oracle_vs_ref.py compares links, scores, ranked cells and paths with the
 * compiled reference (oracle/_ref/ref_stage2.so); tests/golden/chain_golden.npz holds reference outputs. */
#include <stdlib.h>
#include <string.h>

#define ENOUGH_CONSECUTIVE 32		/* stage2.c:67 */
#define GREEDY_NCONSECUTIVE 100		/* :41 */
#define MAX_NACTIVE 100			/* :43 */
#define MAX_SKIPPED 3			/* :86 */
#define EXON_DEFN 30			/* :85 */
#define SCORE_FOR_RESTRICT 10		/* :88 */
#define EQUAL_DISTANCE_NOT_SPLICING 9	/* :79 */
#define TEN_THOUSAND 8192		/* :110 */
#define FINAL_SCORE_TOLERANCE 20	/* :64 */
#define MIN_TERMINAL_NCONSECUTIVE 8	/* :34 */

static int g_splicingp = 1, g_sufflookback = 60, g_nsufflookback = 5, g_maxintronlen = 500000;

void orcs2_setup (int splicingp, int cross_species_p, int sufflookback, int nsufflookback, int maxintronlen) {
  (void) cross_species_p;
  g_splicingp = splicingp; g_sufflookback = sufflookback; g_nsufflookback = nsufflookback; g_maxintronlen = maxintronlen;
}

typedef struct {
  const unsigned int *pos; const int *npos; const unsigned int *mina, *maxa;
  int L, tot, qs, qe, k, localp, skiprep, favor_right, middlep, fwd;
  int *off;				/* CSR offset of each querypos */
  int *consec, *root, *ppos, *phit, *trace, *score, *next;	/* per hit */
  int *first;				/* per querypos: head of its active chain */
  int *proc, nproc;			/* processed queryposes, oldest first (Intlist head = proc[nproc-1]) */
  int tracei;
} chain_t;

#define P(c,q,h) ((c)->pos[(c)->off[q] + (h)])
/* "pp lies more than x before position" and friends, in the direction of the walk (unsigned, as the reference) */
#define BEFORE_LT(c,pp,x,position) ((c)->fwd ? ((pp) > (position) + (x)) : ((pp) + (x) < (position)))
#define BEFORE_LE(c,pp,x,position) ((c)->fwd ? ((pp) >= (position) + (x)) : ((pp) + (x) <= (position)))
#define AT(c,pp,x,position) ((c)->fwd ? ((pp) == (position) + (x)) : ((pp) + (x) == (position)))
#define GDIST(c,pp,position) ((c)->fwd ? (int) ((pp) - (position)) : (int) ((position) - (pp)))
#define QDIST(c,q,pq) ((c)->fwd ? (pq) - (q) : (q) - (pq))

static void revise_active (chain_t *c, int q, int lo, int hi) {	/* stage2.c:2956 */
  int *sc = c->score + c->off[q], *nx = c->next + c->off[q], *ptr, best, thr, h;
  if (lo >= hi) { c->first[q] = -1; return; }
  best = sc[lo];
  for (h = lo + 1; h < hi; h++) if (sc[h] > best) best = sc[h];
  thr = best - SCORE_FOR_RESTRICT;
  if (thr < 0) thr = 0;
  c->first[q] = -1;
  ptr = &c->first[q];
  if (c->fwd) {
    for (h = hi - 1; h >= lo; h--) if (sc[h] > thr) { *ptr = h; ptr = &nx[h]; }
  } else {
    for (h = lo; h < hi; h++) if (sc[h] > thr) { *ptr = h; ptr = &nx[h]; }
  }
  *ptr = -1;
}

/* candidate from ranges 2 and 4 of section D; returns 1 if it became the best */
typedef struct { int consec, root, score, pp, ph, trace; } best_t;

static void scan_prev (chain_t *c, best_t *b, unsigned int position, int q, int pq, int ph, int range1) {
  const int o = c->off[pq], qd = QDIST(c,q,pq), credit = -qd / c->k;
  unsigned int pp;
  int s, diff;
  (void) range1;
  /* range 2 (stage2.c:1236-1350 / :1712-1850): intron-sized jumps */
  while (ph != -1 && (pp = c->pos[o + ph], BEFORE_LT(c,pp,EQUAL_DISTANCE_NOT_SPLICING + qd,position))) {
    diff = GDIST(c,pp,position) - qd;
    s = c->score[o + ph] + credit;
    s -= g_splicingp ? (diff / TEN_THOUSAND + 1) : (diff + 1);
    if (s > b->score) {
      b->consec = (diff <= 0) ? c->consec[o + ph] + qd : 0;
      b->root = c->root[o + ph]; b->score = s; b->pp = pq; b->ph = ph; b->trace = ++c->tracei;
    }
    ph = c->next[o + ph];
  }
  /* ranges 3+4 (:1357-1420 / :1860-1915): near the diagonal, down to one k-mer apart */
  while (ph != -1 && (pp = c->pos[o + ph], BEFORE_LE(c,pp,c->k,position))) {
    int gd = GDIST(c,pp,position);
    diff = gd > qd ? gd - qd : qd - gd;
    s = c->score[o + ph] + 1;
    if (s > b->score) {
      b->consec = (diff <= 0) ? c->consec[o + ph] + qd : 0;
      b->root = c->root[o + ph]; b->score = s; b->pp = pq; b->ph = ph; b->trace = c->trace[o + ph];
    }
    ph = c->next[o + ph];
  }
}

static void commit (chain_t *c, int q, int h, const best_t *b) {	/* stage2.c:1428-1450 / :1925-1950 */
  const int i = c->off[q] + h;
  c->consec[i] = b->consec; c->root[i] = b->root; c->ppos[i] = b->pp; c->phit[i] = b->ph;
  if (b->pp >= 0) { c->trace[i] = b->trace; c->score[i] = b->score; }
  else if (c->localp) { c->trace[i] = ++c->tracei; c->score[i] = c->k; }
  else { c->trace[i] = ++c->tracei; c->score[i] = b->score; }
}

static void score_one (chain_t *c, int q, int h) {	/* stage2.c:1073 */
  const unsigned int position = P(c,q,h);
  best_t b = { c->k, (int) position, 0, -1, -1, 0 };
  int nlookback = g_nsufflookback, lookback = g_sufflookback, i, nseen, last_trace, donep, ph, pq, qd, o;
  unsigned int pp;

  if (c->nproc > 0) {	/* A: the adjacent (last processed) querypos */
    pq = c->proc[c->nproc - 1]; o = c->off[pq]; qd = QDIST(c,q,pq);
    ph = c->first[pq];
    pp = position;
    while (ph != -1 && (pp = c->pos[o + ph], BEFORE_LT(c,pp,qd,position))) ph = c->next[o + ph];
    if (AT(c,pp,qd,position)) {
      b.consec = c->consec[o + ph] + qd; b.root = c->root[o + ph]; b.score = c->score[o + ph] + qd;
      b.pp = pq; b.ph = ph; b.trace = c->trace[o + ph];
      nlookback = 1; lookback = g_sufflookback / 2;
    }
  }
  /* D: earlier queryposes */
  donep = 0; nseen = 0; last_trace = -1;
  for (i = c->nproc - 1; i >= 0 && b.consec < ENOUGH_CONSECUTIVE && !donep; i--, nseen++) {
    pq = c->proc[i]; o = c->off[pq]; qd = QDIST(c,q,pq);
    if (nseen > nlookback && qd - c->k > lookback) donep = 1;
    if ((ph = c->first[pq]) != -1) {
      while (ph != -1 && c->trace[o + ph] == last_trace) ph = c->next[o + ph];	/* range 0 */
      if (ph != -1) last_trace = c->trace[o + ph];
      if (g_splicingp) {								/* range 1 */
        while (ph != -1 && BEFORE_LE(c,c->pos[o + ph],g_maxintronlen + qd,position)) ph = c->next[o + ph];
      }
      scan_prev(c,&b,position,q,pq,ph,0);
    }
  }
  commit(c,q,h,&b);
}

static void fresh_start (chain_t *c, int q, int h, unsigned int position) {	/* the "no adjacent, nothing processed" cells of _mult */
  const int i = c->off[q] + h;
  c->consec[i] = c->k; c->root[i] = (int) position; c->ppos[i] = -1; c->phit[i] = -1;
  if (c->localp) { c->trace[i] = ++c->tracei; c->score[i] = c->k; }
  else c->score[i] = 0;
}

static void score_mult (chain_t *c, int q, int lo, int hi) {	/* stage2.c:1470 */
  const int nhits = hi - lo;
  const unsigned int *positions = c->pos + c->off[q] + lo;
  int hiti, *frontier, nseen, i, max_adj = 0, max_nonadj = 0, overall = 0, adjq, adjo, adjqd, adjf, ph, max_nseen, last_trace;
  unsigned int position, pp;

  const int h0 = c->fwd ? nhits - 1 : 0, hstep = c->fwd ? -1 : 1;	/* hits are walked away from the processed side */
  int cnt;
  if (c->nproc == 0) {
    for (hiti = 0; hiti < nhits; hiti++) fresh_start(c,q,lo + hiti,positions[hiti]);
    return;
  }
  adjq = c->proc[c->nproc - 1]; adjo = c->off[adjq]; adjqd = QDIST(c,q,adjq);
  frontier = (int *) malloc((size_t) c->nproc * sizeof(int));
  nseen = 0;
  for (i = c->nproc - 1; i >= 0; i--) {
    int pq = c->proc[i], qd = QDIST(c,q,pq);
    if (nseen <= 1 || qd - c->k <= g_sufflookback / 2) max_adj = nseen;
    if (nseen <= g_nsufflookback || qd - c->k <= g_sufflookback) max_nonadj = nseen;
    frontier[nseen++] = c->first[pq];
  }
  /* can we be greedy? (:1640-1660) */
  adjf = c->first[adjq];
  for (hiti = h0, cnt = 0; cnt < nhits; hiti += hstep, cnt++) {
    position = positions[hiti];
    ph = adjf; pp = position;
    while (ph != -1 && (pp = c->pos[adjo + ph], BEFORE_LT(c,pp,adjqd,position))) ph = c->next[adjo + ph];
    adjf = ph;
    if (AT(c,pp,adjqd,position) && c->consec[adjo + ph] + adjqd > overall) overall = c->consec[adjo + ph] + adjqd;
  }
  adjf = c->first[adjq];
  for (hiti = h0, cnt = 0; cnt < nhits; hiti += hstep, cnt++) {
    best_t b;
    position = positions[hiti];
    ph = adjf; pp = position;
    while (ph != -1 && (pp = c->pos[adjo + ph], BEFORE_LT(c,pp,adjqd,position))) ph = c->next[adjo + ph];
    adjf = ph;
    if (AT(c,pp,adjqd,position)) {
      b.consec = c->consec[adjo + ph] + adjqd; b.root = c->root[adjo + ph]; b.pp = adjq; b.ph = ph;
      b.score = c->score[adjo + ph] + adjqd; b.trace = c->trace[adjo + ph];
      max_nseen = max_adj;
    } else {
      b.consec = c->k; b.root = (int) position; b.pp = -1; b.ph = -1; b.score = 0; b.trace = -1;
      max_nseen = max_nonadj;
    }
    if (overall < GREEDY_NCONSECUTIVE) {
      nseen = 0; last_trace = -1;
      for (i = c->nproc - 1; i >= 0 && b.consec < ENOUGH_CONSECUTIVE && nseen <= max_nseen; i--, nseen++) {
	if ((ph = frontier[nseen]) != -1) {
	  int pq = c->proc[i], o = c->off[pq], qd = QDIST(c,q,pq);
	  while (ph != -1 && c->trace[o + ph] == last_trace) ph = c->next[o + ph];		/* range 0 */
	  if (ph != -1) last_trace = c->trace[o + ph];
	  while (ph != -1 && BEFORE_LE(c,c->pos[o + ph],g_maxintronlen + qd,position)) ph = c->next[o + ph];	/* range 1, unconditional here */
	  frontier[nseen] = ph;
	  scan_prev(c,&b,position,q,pq,ph,1);
	}
      }
    }
    commit(c,q,lo + hiti,&b);
  }
  free(frontier);
}

static void new_start (chain_t *c, int q) {	/* stage2.c:3793-3812 and :3941-3964 */
  int h, i;
  for (h = 0; h < c->npos[q]; h++) {
    i = c->off[q] + h;
    c->ppos[i] = c->phit[i] = -1; c->consec[i] = c->k; c->trace[i] = -1; c->score[i] = c->k;
  }
}

static void chain_fill (chain_t *c) {	/* stage2.c:3667 (lookback) / :4610 (lookforward) */
  const int step = c->fwd ? -1 : 1;
  int q, h, lo, hi, nhits, nskipped = 0, min_hits = 1000000, specific_q = -1, specific_lo = 0, specific_hi = 0, next_q;
  int grand_score = 0, grand_q = -1, grand_h = -1, best_h, best_s;
#define INRANGE(q) (c->fwd ? (q) >= c->qs : (q) <= c->qe)

  if (c->fwd) { for (q = c->L - 1; q > c->qe; q--) c->first[q] = -1; }
  else { for (q = 0; q < c->qs && q < c->L; q++) c->first[q] = -1; }
  while (INRANGE(q) && c->npos[q] <= 0) { c->first[q] = -1; q += step; }
  if (INRANGE(q)) {
    new_start(c,q);
    revise_active(c,q,0,c->npos[q]);
  }
  while (INRANGE(q)) {
    const unsigned int *m = c->pos + c->off[q];
    best_s = 0; best_h = -1;
    if (c->fwd) {			/* :4781-4790 */
      h = c->npos[q] - 1;
      while (h >= 0 && m[h] > c->maxa[q]) h--;
      hi = h + 1;
      while (h >= 0 && m[h] >= c->mina[q]) h--;
      lo = h + 1;
    } else {				/* :3838-3846 */
      h = 0;
      while (h < c->npos[q] && m[h] < c->mina[q]) h++;
      lo = h;
      while (h < c->npos[q] && m[h] <= c->maxa[q]) h++;
      hi = h;
    }
    if (c->skiprep && hi - lo >= MAX_NACTIVE && nskipped <= MAX_SKIPPED) {
      c->first[q] = -1;
      nskipped++;
      if (hi - lo < min_hits) { min_hits = hi - lo; specific_q = q; specific_lo = lo; specific_hi = hi; }
      q += step;
      continue;
    }
    if (nskipped > MAX_SKIPPED) {
      next_q = q; q = specific_q; lo = specific_lo; hi = specific_hi;
    } else {
      next_q = q + step;
    }
    if ((nhits = hi - lo) > 0) {
      if (nhits == 1) {
	score_one(c,q,lo);
	if (c->score[c->off[q] + lo] > 0) { best_s = c->score[c->off[q] + lo]; best_h = lo; }
      } else {
	score_mult(c,q,lo,hi);
	if (c->fwd) { for (h = hi - 1; h >= lo; h--) if (c->score[c->off[q] + h] > best_s) { best_s = c->score[c->off[q] + h]; best_h = h; } }
	else { for (h = lo; h < hi; h++) if (c->score[c->off[q] + h] > best_s) { best_s = c->score[c->off[q] + h]; best_h = h; } }
      }
      nskipped = 0; min_hits = 1000000; specific_q = -1;
      if (!c->middlep && best_h < 0) new_start(c,q);
      if (g_splicingp && best_h >= 0 && c->phit[c->off[q] + best_h] < 0 &&
	  (c->fwd ? (grand_q <= c->L - c->k && q + c->k <= grand_q) : (grand_q >= 0 && q >= grand_q + c->k))) {	/* :3966-3990 / :4905-4930 */
	if ((best_s = c->score[c->off[grand_q] + grand_h] - QDIST(c,q,grand_q)) > 0) {
	  unsigned int pp = P(c,grand_q,grand_h), position;
	  for (h = lo; h < hi; h++) {
	    position = c->pos[c->off[q] + h];
	    if (c->fwd ? (position + g_maxintronlen < pp) : (position > pp + g_maxintronlen)) {
	    } else if (c->fwd ? (position + c->k <= pp) : (position >= pp + c->k)) {
	      int i = c->off[q] + h;
	      c->consec[i] = c->k; c->ppos[i] = grand_q; c->phit[i] = grand_h; c->trace[i] = ++c->tracei; c->score[i] = best_s;
	    }
	  }
	}
      }
      if (best_h >= 0 && best_s >= grand_score && c->consec[c->off[q] + best_h] > EXON_DEFN) {
	grand_score = best_s; grand_q = q; grand_h = best_h;
      }
    }
    revise_active(c,q,lo,hi);
    if (c->npos[q] > 0) c->proc[c->nproc++] = q;
    q = next_q;
  }
#undef INRANGE
}

/* ---- ranking (get_cells_fwd) ------------------------------------------------------------------------ */
typedef struct { int root, end, q, h, score, ord; } cell_t;
static int g_favor_right;
static int cmp_root (const void *a, const void *b) {	/* Cell_rootposition_left/right_cmp: a total order */
  const cell_t *x = (const cell_t *) a, *y = (const cell_t *) b;
  if (x->root != y->root) return x->root < y->root ? -1 : 1;
  if (x->score != y->score) return x->score > y->score ? -1 : 1;
  if (x->q != y->q) return x->q > y->q ? -1 : 1;
  if (x->h != y->h) return ((x->h < y->h) != (g_favor_right != 0)) ? -1 : 1;
  return 0;
}
static int cmp_score (const void *a, const void *b) {	/* Cell_score_cmp under glibc's stable merge sort */
  const cell_t *x = (const cell_t *) a, *y = (const cell_t *) b;
  if (x->score != y->score) return x->score > y->score ? -1 : 1;
  return x->ord < y->ord ? -1 : (x->ord > y->ord ? 1 : 0);
}

static cell_t *rank_cells (chain_t *c, int *nunique) {
  cell_t *cells = (cell_t *) malloc((size_t) (c->tot + 1) * sizeof(cell_t));
  int n = 0, q, h, i, k = 0, last_root = -1, best_for_root = -1;
  for (q = c->qs; q <= c->qe; q++) {
    for (h = 0; h < c->npos[q]; h++) {
      i = c->off[q] + h;
      if (c->score[i] > 0) {
	cells[n].root = c->root[i]; cells[n].end = (int) c->pos[i]; cells[n].q = q; cells[n].h = h; cells[n].score = c->score[i]; n++;
      }
    }
  }
  g_favor_right = c->favor_right;
  qsort(cells,n,sizeof(cell_t),cmp_root);
  for (i = 0; i < n; i++) {
    if (cells[i].root != last_root) { last_root = cells[i].root; best_for_root = cells[i].score; cells[k] = cells[i]; cells[k].ord = k; k++; }
    else if (cells[i].score == best_for_root) { cells[k] = cells[i]; cells[k].ord = k; k++; }
  }
  qsort(cells,k,sizeof(cell_t),cmp_score);
  *nunique = k;
  return cells;
}

static void chain_init (chain_t *c, const unsigned int *positions, const int *npositions, int querylength, int totalpositions,
			const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize,
			int localp, int skip_repetitive_p, int favor_right_p, int middlep, int fwd) {
  int q, o = 0;
  size_t n = (size_t) totalpositions + 1;
  memset(c,0,sizeof(*c));
  c->pos = positions; c->npos = npositions; c->mina = minactive; c->maxa = maxactive;
  c->L = querylength; c->tot = totalpositions; c->qs = querystart; c->qe = queryend; c->k = indexsize;
  c->localp = localp; c->skiprep = skip_repetitive_p; c->favor_right = favor_right_p; c->middlep = middlep; c->fwd = fwd;
  c->off = (int *) malloc((size_t) (querylength + 1) * sizeof(int));
  for (q = 0; q < querylength; q++) { c->off[q] = o; if (npositions[q] > 0) o += npositions[q]; }
  c->off[querylength] = o;
  c->consec = (int *) calloc(n,sizeof(int)); c->root = (int *) calloc(n,sizeof(int)); c->ppos = (int *) calloc(n,sizeof(int));
  c->phit = (int *) calloc(n,sizeof(int)); c->trace = (int *) calloc(n,sizeof(int)); c->score = (int *) calloc(n,sizeof(int));
  c->next = (int *) calloc(n,sizeof(int));
  c->first = (int *) calloc((size_t) querylength + 1,sizeof(int));
  c->proc = (int *) malloc((size_t) (querylength + 1) * sizeof(int));
}
static void chain_free (chain_t *c) {
  free(c->off); free(c->consec); free(c->root); free(c->ppos); free(c->phit); free(c->trace); free(c->score); free(c->next);
  free(c->first); free(c->proc);
}

static int scores_dir (int fwd, const unsigned int *positions, const int *npositions, int querylength, int totalpositions,
		  const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize,
		  int localp, int skip_repetitive_p, int favor_right_p, int middlep,
		  int *links_out, int *scores_out, int *cells_out, int cells_cap) {
  chain_t c;
  cell_t *cells;
  int n, i;
  chain_init(&c,positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize,
	     localp,skip_repetitive_p,favor_right_p,middlep,fwd);
  chain_fill(&c);
  cells = rank_cells(&c,&n);
  for (i = 0; i < totalpositions; i++) {
    if (links_out) {
      links_out[5*i+0] = c.consec[i]; links_out[5*i+1] = c.root[i]; links_out[5*i+2] = c.ppos[i];
      links_out[5*i+3] = c.phit[i]; links_out[5*i+4] = c.trace[i];
    }
    if (scores_out) scores_out[i] = c.score[i];
  }
  for (i = 0; i < n && i < cells_cap; i++) {
    cells_out[5*i+0] = cells[i].root; cells_out[5*i+1] = cells[i].end; cells_out[5*i+2] = cells[i].q;
    cells_out[5*i+3] = cells[i].h; cells_out[5*i+4] = cells[i].score;
  }
  free(cells);
  chain_free(&c);
  return n;
}

static int paths_dir (int fwd, const unsigned int *positions, const int *npositions, int querylength, int totalpositions,
		 const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize,
		 int localp, int skip_repetitive_p, int favor_right_p, int middlep, int max_nalignments,
		 const char *queryseq, const char *queryuc, int *path_len, int maxpaths, int *pairs_out, int pairs_cap) {
  chain_t c;
  cell_t *cells;
  int n, i, npaths = 0, npairs = 0, best;
  (void) queryseq; (void) queryuc;
  chain_init(&c,positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize,
	     localp,skip_repetitive_p,favor_right_p,middlep,fwd);
  chain_fill(&c);
  cells = rank_cells(&c,&n);
  if (n > 0) {
    best = cells[0].score;
    for (i = 0; i < n && (i < max_nalignments || cells[i].score == best) && cells[i].score > best - FINAL_SCORE_TOLERANCE; i++) {
      int q = cells[i].q, h = cells[i].h, j, k = 0, start = npairs, t;
      while (q >= 0 && c.consec[c.off[q] + h] < MIN_TERMINAL_NCONSECUTIVE) {	/* traceback_one, part 1: prune the 3' end */
	j = c.off[q] + h; q = c.ppos[j]; h = c.phit[j];
      }
      while (q >= 0) {
	j = c.off[q] + h;
	if (npairs < pairs_cap) { pairs_out[2*npairs] = q; pairs_out[2*npairs+1] = (int) c.pos[j]; }
	npairs++; k++;
	q = c.ppos[j]; h = c.phit[j];
      }
      if (npairs <= pairs_cap) {		/* list head = the far end of the walk (the last pair pushed): reverse what the walk produced */
	for (j = start, t = npairs - 1; j < t; j++, t--) {
	  int a = pairs_out[2*j], b = pairs_out[2*j+1];
	  pairs_out[2*j] = pairs_out[2*t]; pairs_out[2*j+1] = pairs_out[2*t+1];
	  pairs_out[2*t] = a; pairs_out[2*t+1] = b;
	}
      }
      if (npaths < maxpaths) path_len[npaths] = k;
      npaths++;
    }
  }
  free(cells);
  chain_free(&c);
  return (npairs > pairs_cap) ? -npairs : npaths;
}

#define SCORES_ARGS const unsigned int *positions, const int *npositions, int querylength, int totalpositions, \
    const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize, \
    int localp, int skip_repetitive_p, int favor_right_p, int middlep, int *links_out, int *scores_out, int *cells_out, int cells_cap
#define SCORES_PASS positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize, \
    localp,skip_repetitive_p,favor_right_p,middlep,links_out,scores_out,cells_out,cells_cap
#define PATHS_ARGS const unsigned int *positions, const int *npositions, int querylength, int totalpositions, \
    const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize, \
    int localp, int skip_repetitive_p, int favor_right_p, int middlep, int max_nalignments, \
    const char *queryseq, const char *queryuc, int *path_len, int maxpaths, int *pairs_out, int pairs_cap
#define PATHS_PASS positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize, \
    localp,skip_repetitive_p,favor_right_p,middlep,max_nalignments,queryseq,queryuc,path_len,maxpaths,pairs_out,pairs_cap

int orcs2_scores (SCORES_ARGS) { return scores_dir(0,SCORES_PASS); }
int orcs2_scores_fwd (SCORES_ARGS) { return scores_dir(1,SCORES_PASS); }
int orcs2_paths (PATHS_ARGS) { return paths_dir(0,PATHS_PASS); }
int orcs2_paths_fwd (PATHS_ARGS) { return paths_dir(1,PATHS_PASS); }
