/* dynprog_sm100.c -- the stage3-side binding of the B200 DP engine.
 *
 * Compiled with the REFERENCE's headers and linked with the reference's unmodified objects; it
 * provides the five DP entry points that stage3.c calls (dynprog_single.h:23, dynprog_genome.h:23,
 * dynprog_cdna.h:17, dynprog_end.h:24,46) on top of include/gmapdp_shim.h.  The reference's own
 * definitions of those five symbols are renamed to ref_* in the object files by the build recipe
 * (integration/Makefile), nothing in /root/reference is edited or copied.
 *
 * The DP calls of all worker threads meet in one shared device batch (rendezvous, below): with
 * `gmap -t N` the batches hold up to N boxes.  What the binding does on the host is exactly what
 * the reference entry points do before and after their fills: fetch the genomic segments with
 * Genome_get_segment_* and, for genome gaps, the MaxEnt splice-site probabilities with Maxent_hr_*.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "bool.h"
#include "types.h"
#include "mem.h"
#include "list.h"
#include "pairdef.h"
#include "pairpool.h"
#include "genome.h"
#include "maxent_hr.h"
#include "dynprog.h"
#include "dynprog_single.h"
#include "dynprog_genome.h"
#include "dynprog_cdna.h"
#include "dynprog_end.h"
#include "gmapdp_shim.h"

/* ---- rendezvous: DP calls of all worker threads are gathered into one device batch ---------------
 * A thread queues its call in the shared batch and waits.  The first thread of a generation is its
 * leader: it waits (bounded) until every thread that is currently inside a DP entry point has queued
 * its call, runs the batch on the GPU, hands each waiter its result, and wakes them.  With one thread
 * this degenerates to one batch per call, with `gmap -t N` (N >> cores: workers only block here) the
 * batches hold up to N boxes.  Results do not depend on batch composition. */
static gmapdp_ctx *sm100_ctx = NULL;
static gmapdp_batch *sm100_shared = NULL;
static pthread_once_t sm100_once = PTHREAD_ONCE_INIT;
static pthread_mutex_t sm100_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t sm100_cv_done = PTHREAD_COND_INITIALIZER, sm100_cv_submit = PTHREAD_COND_INITIALIZER;
static volatile int sm100_inflight = 0;		/* threads inside a DP entry point */
static long sm100_wait_us = 300;
static unsigned long sm100_nbatches = 0, sm100_ncalls = 0;
static double sm100_t_wait = 0.0, sm100_t_run = 0.0, sm100_t_result = 0.0, sm100_cells = 0.0;

static double now_s (void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

typedef struct sm100_slot {
  int id, done, n, cap, mode;
  int iout[10];
  double dout[2];
  gmapdp_pair *pairs;
} sm100_slot;

#define SM100_MAXPENDING 4096
static sm100_slot *sm100_pending[SM100_MAXPENDING];
static int sm100_npending = 0, sm100_leader = 0;
static __thread sm100_slot sm100_my = {0,0,0,0,0,{0},{0},NULL};

static void sm100_report (void) {
  if (getenv("GMAP_SM100_STATS"))
    fprintf(stderr,"gmap.sm100: %lu DP calls in %lu device batches (%.1f calls per batch)\n",sm100_ncalls,sm100_nbatches,
	    sm100_nbatches ? (double) sm100_ncalls / (double) sm100_nbatches : 0.0);
  if (getenv("GMAP_SM100_STATS"))
    fprintf(stderr,"gmap.sm100 timing: leader wait %.3f s, device batches %.3f s, result replay %.3f s, %.0f DP cells\n",
	    sm100_t_wait,sm100_t_run,sm100_t_result,sm100_cells);
}

static void sm100_init (void) {
  const char *dev = getenv("GMAP_SM100_DEVICE"), *w = getenv("GMAP_SM100_WAIT_US");
  if (w) sm100_wait_us = atol(w);
  if (gmapdp_create(&sm100_ctx,dev ? atoi(dev) : 0) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: %s\n",gmapdp_last_error(sm100_ctx));
    exit(9);
  }
  atexit(sm100_report);
}

/* the process-wide engine, also used by the stage 2 binding (stage2_sm100.c) */
gmapdp_ctx *sm100_context (void) {
  pthread_once(&sm100_once,sm100_init);
  return sm100_ctx;
}

/* called with the mutex held, before queueing: the shared batch is created from the first caller's limits
   (all Dynprog_T of a gmap run have the same max_rlength / max_glength, gmap.c:4898-4903) */
static gmapdp_batch *shared_batch (Dynprog_T dynprog) {
  if (sm100_shared == NULL) sm100_shared = GmapDP_batch_new(sm100_ctx,dynprog->max_rlength,dynprog->max_glength);
  return sm100_shared;
}

static void slot_reserve (int need_pairs) {
  if (need_pairs > sm100_my.cap) {
    free(sm100_my.pairs);
    sm100_my.cap = need_pairs + 1024;
    sm100_my.pairs = (gmapdp_pair *) malloc((size_t) sm100_my.cap * sizeof(gmapdp_pair));
  }
}

/* called with the mutex held and the call queued under `id`; returns with the result in sm100_my */
static void rendezvous (int id, int mode) {
  sm100_slot *me = &sm100_my;
  int k;
  me->id = id; me->done = 0; me->mode = mode;
  if (sm100_npending >= SM100_MAXPENDING) { fprintf(stderr,"gmap.sm100: too many waiting threads\n"); exit(9); }
  sm100_pending[sm100_npending++] = me;
  if (!sm100_leader) {
    struct timespec dl;
    double t0 = now_s(), t1, t2;
    sm100_leader = 1;
    clock_gettime(CLOCK_REALTIME,&dl);
    dl.tv_nsec += sm100_wait_us * 1000L;
    while (dl.tv_nsec >= 1000000000L) { dl.tv_nsec -= 1000000000L; dl.tv_sec++; }
    while (sm100_npending < sm100_inflight && sm100_npending < SM100_MAXPENDING) {
      if (pthread_cond_timedwait(&sm100_cv_submit,&sm100_mu,&dl) != 0) break;	/* timeout */
    }
    t1 = now_s();
    if (GmapDP_batch_run(sm100_shared) != GMAPDP_OK) {
      fprintf(stderr,"gmap.sm100: %s\n",GmapDP_batch_error(sm100_shared));
      exit(9);
    }
    t2 = now_s();
    sm100_cells += (double) GmapDP_batch_cells(sm100_shared);
    sm100_nbatches++; sm100_ncalls += sm100_npending;
    for (k = 0; k < sm100_npending; k++) {
      sm100_slot *s = sm100_pending[k];
      s->n = GmapDP_result(sm100_shared,s->id,s->iout,s->dout,s->pairs,s->cap);
      s->done = 1;
    }
    sm100_npending = 0;
    GmapDP_batch_clear(sm100_shared);
    sm100_t_wait += t1 - t0; sm100_t_run += t2 - t1; sm100_t_result += now_s() - t2;
    sm100_leader = 0;
    pthread_cond_broadcast(&sm100_cv_done);
  } else {
    pthread_cond_signal(&sm100_cv_submit);
    while (!me->done) pthread_cond_wait(&sm100_cv_done,&sm100_mu);
  }
}

#define ENTER() do { pthread_once(&sm100_once,sm100_init); __sync_fetch_and_add(&sm100_inflight,1); } while (0)
#define LEAVE() __sync_fetch_and_sub(&sm100_inflight,1)

/* records come head first: cons them from the tail */
static List_T pairs_to_list (Pairpool_T pool, const gmapdp_pair *p, int n) {
  List_T l = NULL;
  Pair_T g;
  int k;
  for (k = n - 1; k >= 0; k--) {
    if (p[k].gapp) {
      l = Pairpool_push_gapholder(l,pool,p[k].queryjump,p[k].genomejump,/*leftpair*/NULL,/*rightpair*/NULL,/*knownp*/false);
      g = (Pair_T) List_head(l);
      g->introntype = p[k].introntype; g->donor_prob = p[k].donor_prob; g->acceptor_prob = p[k].acceptor_prob;
    } else {
      l = Pairpool_push(l,pool,p[k].querypos,p[k].genomepos,p[k].cdna,p[k].comp,p[k].genome,p[k].genomealt,p[k].dynprogindex);
    }
  }
  return l;
}

#define SET(dst,v) do { if ((v) != GMAPDP_UNSET) *(dst) = (v); } while (0)

List_T
Dynprog_single_gap (int *dynprogindex, int *traceback_score, int *nmatches, int *nmismatches, int *nopens, int *nindels,
		    Dynprog_T dynprog, char *rsequence, char *rsequenceuc,
		    int rlength, int glength, int roffset, int goffset,
		    Univcoord_T chroffset, Univcoord_T chrhigh,
		    bool watsonp, int genestrand, bool jump_late_p,
		    Genome_T genome, Genome_T genomealt, Pairpool_T pairpool,
		    int extraband_single, bool widebandp, double defect_rate) {
  char *gseq, *galt, empty[1] = {'\0'};
  int id;
  const int *iout;
  List_T l;
  bool fetch = (rlength > 0 && glength > 0 && rlength <= dynprog->max_rlength && glength <= dynprog->max_glength);

  ENTER();
  slot_reserve(rlength + glength + 8);
  if (fetch) {
    gseq = (char *) malloc(glength + 1); galt = (char *) malloc(glength + 1);
    if (watsonp) Genome_get_segment_right(gseq,galt,genome,genomealt,chroffset+goffset,glength,chrhigh,/*revcomp*/false);
    else Genome_get_segment_left(gseq,galt,genome,genomealt,chrhigh-goffset+1,glength,chroffset,/*revcomp*/true);
  } else {
    gseq = galt = empty;
  }
  pthread_mutex_lock(&sm100_mu);
  id = GmapDP_single_gap(shared_batch(dynprog),*dynprogindex,rsequence,rsequenceuc,rlength,glength,roffset,goffset,gseq,galt,
			 jump_late_p,extraband_single,widebandp,defect_rate);
  rendezvous(id,GMAPDP_SINGLE);
  pthread_mutex_unlock(&sm100_mu);
  LEAVE();
  iout = sm100_my.iout;
  *dynprogindex = iout[0]; *traceback_score = iout[1]; *nmatches = iout[2]; *nmismatches = iout[3];
  *nopens = iout[4]; *nindels = iout[5];
  if (fetch) { free(galt); free(gseq); }
  l = (sm100_my.n < 0) ? (List_T) NULL : pairs_to_list(pairpool,sm100_my.pairs,sm100_my.n);
  return l;
}

static List_T
end_gap (bool end5, int *dynprogindex, int *traceback_score, int *nmatches, int *nmismatches, int *nopens, int *nindels,
	 Dynprog_T dynprog, char *rseq, char *rsequc, int rlength, int glength, int roffset, int goffset,
	 Univcoord_T chroffset, Univcoord_T chrhigh, bool watsonp, bool jump_late_p,
	 Genome_T genome, Genome_T genomealt, Pairpool_T pairpool, int extraband_end, double defect_rate,
	 Endalign_T endalign, bool require_pos_score_p) {
  char *gseq, *galt, empty[1] = {'\0'};
  int id, gl = glength;
  const int *iout;
  bool fetch;

  ENTER();
  slot_reserve(rlength + glength + 8);
  /* the reference chops before it fetches (dynprog_end.c:1357-1378 / :1986-1999) */
  if (endalign != QUERYEND_NOGAPS && gl > dynprog->max_glength) gl = dynprog->max_glength;
  fetch = (rlength > 0 && gl > 0 && !(end5 && goffset < 0));
  if (fetch) {
    gseq = (char *) malloc(gl + 1); galt = (char *) malloc(gl + 1);
    if (end5) {
      if (watsonp) Genome_get_segment_left(gseq,galt,genome,genomealt,chroffset+goffset+1,gl,chroffset,/*revcomp*/false);
      else Genome_get_segment_right(gseq,galt,genome,genomealt,chrhigh-goffset,gl,chrhigh,/*revcomp*/true);
    } else {
      if (watsonp) Genome_get_segment_right(gseq,galt,genome,genomealt,chroffset+goffset,gl,chrhigh,/*revcomp*/false);
      else Genome_get_segment_left(gseq,galt,genome,genomealt,chrhigh-goffset+1,gl,chroffset,/*revcomp*/true);
    }
  } else {
    gseq = galt = empty;
  }
  pthread_mutex_lock(&sm100_mu);
  if (end5) id = GmapDP_end5_gap(shared_batch(dynprog),*dynprogindex,rseq,rsequc,rlength,glength,roffset,goffset,gseq,galt,jump_late_p,
				 extraband_end,defect_rate,(int) endalign,require_pos_score_p);
  else id = GmapDP_end3_gap(shared_batch(dynprog),*dynprogindex,rseq,rsequc,rlength,glength,roffset,goffset,gseq,galt,jump_late_p,
			    extraband_end,defect_rate,(int) endalign,require_pos_score_p);
  rendezvous(id,end5 ? GMAPDP_END5 : GMAPDP_END3);
  pthread_mutex_unlock(&sm100_mu);
  LEAVE();
  iout = sm100_my.iout;
  *dynprogindex = iout[0]; *traceback_score = iout[1]; *nmatches = iout[2]; *nmismatches = iout[3];
  *nopens = iout[4]; *nindels = iout[5];
  if (fetch) { free(galt); free(gseq); }
  return (sm100_my.n < 0) ? (List_T) NULL : pairs_to_list(pairpool,sm100_my.pairs,sm100_my.n);
}

List_T
Dynprog_end5_gap (int *dynprogindex, int *traceback_score, int *nmatches, int *nmismatches,
		  int *nopens, int *nindels, Dynprog_T dynprog,
		  char *rev_rsequence, char *rev_rsequenceuc,
		  int rlength, int glength, int rev_roffset, int rev_goffset,
		  Univcoord_T chroffset, Univcoord_T chrhigh,
		  bool watsonp, int genestrand, bool jump_late_p,
		  Genome_T genome, Genome_T genomealt, Pairpool_T pairpool,
		  int extraband_end, double defect_rate, Endalign_T endalign,
		  bool require_pos_score_p) {
  return end_gap(true,dynprogindex,traceback_score,nmatches,nmismatches,nopens,nindels,dynprog,rev_rsequence,rev_rsequenceuc,
		 rlength,glength,rev_roffset,rev_goffset,chroffset,chrhigh,watsonp,jump_late_p,genome,genomealt,pairpool,
		 extraband_end,defect_rate,endalign,require_pos_score_p);
}

List_T
Dynprog_end3_gap (int *dynprogindex, int *traceback_score, int *nmatches, int *nmismatches,
		  int *nopens, int *nindels, Dynprog_T dynprog,
		  char *rsequence, char *rsequenceuc,
		  int rlength, int glength, int roffset, int goffset,
		  Univcoord_T chroffset, Univcoord_T chrhigh,
		  bool watsonp, int genestrand, bool jump_late_p,
		  Genome_T genome, Genome_T genomealt, Pairpool_T pairpool,
		  int extraband_end, double defect_rate, Endalign_T endalign,
		  bool require_pos_score_p) {
  return end_gap(false,dynprogindex,traceback_score,nmatches,nmismatches,nopens,nindels,dynprog,rsequence,rsequenceuc,
		 rlength,glength,roffset,goffset,chroffset,chrhigh,watsonp,jump_late_p,genome,genomealt,pairpool,
		 extraband_end,defect_rate,endalign,require_pos_score_p);
}

List_T
Dynprog_genome_gap (int *dynprogindex, int *new_leftgenomepos, int *new_rightgenomepos,
		    double *left_prob, double *right_prob,
		    int *traceback_score, int *nmatches, int *nmismatches,
		    int *nopens, int *nindels, int *exonhead, int *introntype,
		    Dynprog_T dynprogL, Dynprog_T dynprogR,
		    char *rsequence, char *rsequenceuc, int rlength, int glengthL, int glengthR,
		    int roffset, int goffsetL, int rev_goffsetR,
		    Chrnum_T chrnum, Univcoord_T chroffset, Univcoord_T chrhigh,
		    int cdna_direction, bool watsonp, int genestrand, bool jump_late_p,
		    Genome_T genome, Genome_T genomealt, Pairpool_T pairpool, int extraband_paired,
		    double defect_rate, int maxpeelback, bool halfp, bool finalp) {
  char *gL, *gLa, *gR, *gRa, empty[1] = {'\0'};
  double *lp = NULL, *rp = NULL;
  const double *dout;
  const int *iout;
  int id, c;
  Univcoord_T pos;
  bool fetch = (rlength > 1 && rlength <= dynprogL->max_rlength && glengthL <= dynprogL->max_glength &&
		rlength <= dynprogR->max_rlength && glengthR <= dynprogR->max_glength && glengthL > 0 && glengthR > 0);

  ENTER();
  slot_reserve(2 * rlength + glengthL + glengthR + 16);
  if (fetch) {
    gL = (char *) malloc(glengthL + 1); gLa = (char *) malloc(glengthL + 1);
    gR = (char *) malloc(glengthR + 1); gRa = (char *) malloc(glengthR + 1);
    lp = (double *) calloc(glengthL + 1,sizeof(double)); rp = (double *) calloc(glengthR + 1,sizeof(double));
    if (watsonp) {
      Genome_get_segment_right(gL,gLa,genome,genomealt,chroffset+goffsetL,glengthL,chrhigh,/*revcomp*/false);
      Genome_get_segment_left(gR,gRa,genome,genomealt,chroffset+rev_goffsetR+1,glengthR,chroffset,/*revcomp*/false);
    } else {
      Genome_get_segment_left(gL,gLa,genome,genomealt,chrhigh-goffsetL+1,glengthL,chroffset,/*revcomp*/true);
      Genome_get_segment_right(gR,gRa,genome,genomealt,chrhigh-rev_goffsetR,glengthR,chrhigh,/*revcomp*/true);
    }
    /* the probability arrays of dynprog_genome.c:970-1061 */
    if (watsonp) {
      for (c = 0; c < glengthL - 1; c++) {
	pos = chroffset + goffsetL + c;
	lp[c] = (cdna_direction > 0) ? Maxent_hr_donor_prob(genome,genomealt,pos,chroffset) : Maxent_hr_antiacceptor_prob(genome,genomealt,pos,chroffset);
      }
      for (c = 0; c < glengthR - 1; c++) {
	pos = chroffset + rev_goffsetR - c + 1;
	rp[c] = (cdna_direction > 0) ? Maxent_hr_acceptor_prob(genome,genomealt,pos,chroffset) : Maxent_hr_antidonor_prob(genome,genomealt,pos,chroffset);
      }
    } else {
      for (c = 0; c < glengthL - 1; c++) {
	pos = chrhigh - goffsetL - c + 1;
	lp[c] = (cdna_direction > 0) ? Maxent_hr_antidonor_prob(genome,genomealt,pos,chroffset) : Maxent_hr_acceptor_prob(genome,genomealt,pos,chroffset);
      }
      for (c = 0; c < glengthR - 1; c++) {
	pos = chrhigh - rev_goffsetR + c;
	rp[c] = (cdna_direction > 0) ? Maxent_hr_antiacceptor_prob(genome,genomealt,pos,chroffset) : Maxent_hr_donor_prob(genome,genomealt,pos,chroffset);
      }
    }
  } else {
    gL = gLa = gR = gRa = empty;
  }
  pthread_mutex_lock(&sm100_mu);
  id = GmapDP_genome_gap(shared_batch(dynprogL),*dynprogindex,rsequence,rsequenceuc,rlength,glengthL,glengthR,roffset,goffsetL,rev_goffsetR,
			 gL,gLa,gR,gRa,lp,rp,cdna_direction,jump_late_p,extraband_paired,defect_rate,maxpeelback,halfp,finalp);
  rendezvous(id,GMAPDP_GENOME);
  pthread_mutex_unlock(&sm100_mu);
  LEAVE();
  iout = sm100_my.iout; dout = sm100_my.dout;
  *dynprogindex = iout[0];
  SET(new_leftgenomepos,iout[1]); SET(new_rightgenomepos,iout[2]); SET(traceback_score,iout[3]);
  *nmatches = iout[4]; *nmismatches = iout[5]; *nopens = iout[6]; *nindels = iout[7];
  SET(exonhead,iout[8]); *introntype = iout[9];
  *left_prob = dout[0]; *right_prob = dout[1];
  if (fetch) { free(rp); free(lp); free(gRa); free(gR); free(gLa); free(gL); }
  return (sm100_my.n < 0) ? (List_T) NULL : pairs_to_list(pairpool,sm100_my.pairs,sm100_my.n);
}

List_T
Dynprog_cdna_gap (int *dynprogindex, int *traceback_score, bool *incompletep,
		  Dynprog_T dynprogL, Dynprog_T dynprogR, char *rsequenceL, char *rsequence_ucL,
		  char *rev_rsequenceR, char *rev_rsequence_ucR,
		  int rlengthL, int rlengthR, int glength,
		  int roffsetL, int rev_roffsetR, int goffset,
		  Univcoord_T chroffset, Univcoord_T chrhigh,
		  bool watsonp, int genestrand, bool jump_late_p,
		  Genome_T genome, Genome_T genomealt, Pairpool_T pairpool,
		  int extraband_paired, double defect_rate) {
  char *g, *ga, *rg, *rga, empty[1] = {'\0'};
  const int *iout;
  int id, rev_goffset = goffset + glength - 1;
  bool fetch = (glength > 1 && glength <= dynprogR->max_glength && rlengthR <= dynprogR->max_rlength &&
		glength <= dynprogL->max_glength && rlengthL <= dynprogL->max_rlength);

  ENTER();
  slot_reserve(rlengthL + rlengthR + 2 * glength + 32);
  if (fetch) {
    g = (char *) malloc(glength + 1); ga = (char *) malloc(glength + 1);
    rg = (char *) malloc(glength + 1); rga = (char *) malloc(glength + 1);
    if (watsonp) {
      Genome_get_segment_left(rg,rga,genome,genomealt,chroffset+rev_goffset+1,glength,chroffset,/*revcomp*/false);
      Genome_get_segment_right(g,ga,genome,genomealt,chroffset+goffset,glength,chrhigh,/*revcomp*/false);
    } else {
      Genome_get_segment_right(rg,rga,genome,genomealt,chrhigh-rev_goffset,glength,chrhigh,/*revcomp*/true);
      Genome_get_segment_left(g,ga,genome,genomealt,chrhigh-goffset+1,glength,chroffset,/*revcomp*/true);
    }
  } else {
    g = ga = rg = rga = empty;
  }
  pthread_mutex_lock(&sm100_mu);
  id = GmapDP_cdna_gap(shared_batch(dynprogL),*dynprogindex,rsequenceL,rsequence_ucL,rev_rsequenceR,rev_rsequence_ucR,rlengthL,rlengthR,glength,
		       roffsetL,rev_roffsetR,goffset,g,ga,rg,rga,jump_late_p,extraband_paired,defect_rate);
  rendezvous(id,GMAPDP_CDNA);
  pthread_mutex_unlock(&sm100_mu);
  LEAVE();
  iout = sm100_my.iout;
  *dynprogindex = iout[0];
  SET(traceback_score,iout[1]);
  if (iout[2]) *incompletep = true;
  if (fetch) { free(rga); free(rg); free(ga); free(g); }
  return (sm100_my.n < 0) ? (List_T) NULL : pairs_to_list(pairpool,sm100_my.pairs,sm100_my.n);
}
