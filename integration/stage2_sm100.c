/* stage2_sm100.c -- gmap.sm100's stage 2: the reference's stage2.c, compiled IN PLACE and unmodified, with its two static
 * chaining drivers served by the B200 chaining engine (include/gmapchain_b200.h):
 *
 *     align_compute_lookback     stage2.c:4402   (Stage2_compute :6548, Stage2_compute_one :6859, Stage2_compute_ends :7194)
 *     align_compute_lookforward  stage2.c:5062   (Stage2_compute_starts :7020)
 *
 * The functions are `static`, so the seam is inside the translation unit: the two names are function-like macros while
 * stage2.c is being read.  They paste a prefix onto the FIRST token of the argument list, which is `Chrpos_T` at the
 * definition (then the reference's own body is compiled under the name ref_align_compute_*) and `mappings` at every call
 * site (then the call goes to the sm100_* functions at the bottom of this file).  Everything else of stage 2 --
 * Oligoindex mappings, Diag bounds, convert_to_nucleotides, the filters -- is the reference's code.
 * What is not on the device yet (use_canonical_p, SNP-tolerant tracebacks) is handed to the reference's own body.
 */
#include <pthread.h>
#include "bool.h"
#include "types.h"
#include "genomicpos.h"
#include "list.h"
#include "genome.h"
#include "pairpool.h"
#include "cellpool.h"
#include "gmapchain_b200.h"

#define CHAIN_PARAMS Chrpos_T **mappings, int *npositions, int totalpositions, bool oned_matrix_p, \
    Chrpos_T *minactive, Chrpos_T *maxactive, int *firstactive, int *nactive, Cellpool_T cellpool, \
    char *queryseq_ptr, char *queryuc_ptr, int querylength, int querystart, int queryend, \
    Genome_T genome, Genome_T genomealt, Univcoord_T chroffset, Univcoord_T chrhigh, bool plusp, \
    int indexsize, Pairpool_T pairpool, bool localp, bool skip_repetitive_p, bool use_canonical_p, \
    int non_canonical_penalty, bool favor_right_p, bool middlep, int max_nalignments
#define CHAIN_ARGS mappings,npositions,totalpositions,oned_matrix_p,minactive,maxactive,firstactive,nactive,cellpool, \
    queryseq_ptr,queryuc_ptr,querylength,querystart,queryend,genome,genomealt,chroffset,chrhigh,plusp, \
    indexsize,pairpool,localp,skip_repetitive_p,use_canonical_p,non_canonical_penalty,favor_right_p,middlep,max_nalignments

static List_T ref_align_compute_lookback (CHAIN_PARAMS);
static List_T ref_align_compute_lookforward (CHAIN_PARAMS);
static List_T sm100_lookback (CHAIN_PARAMS);
static List_T sm100_lookforward (CHAIN_PARAMS);

#define align_compute_lookback(first, ...) ACL_##first, __VA_ARGS__)
#define ACL_Chrpos_T ref_align_compute_lookback (Chrpos_T
#define ACL_mappings sm100_lookback (mappings
#define align_compute_lookforward(first, ...) ACF_##first, __VA_ARGS__)
#define ACF_Chrpos_T ref_align_compute_lookforward (Chrpos_T
#define ACF_mappings sm100_lookforward (mappings

#include "stage2.c"

#undef align_compute_lookback
#undef align_compute_lookforward

/* ---- the device side ------------------------------------------------------------------------------------------ */
extern gmapdp_ctx *sm100_context (void);		/* dynprog_sm100.c: the process-wide engine */

/* Rendezvous, as for the DP calls (dynprog_sm100.c): the chaining calls of all worker threads are queued in one
   shared batch; the first thread of a generation is its leader, waits (bounded) for the threads that are inside a
   chaining call to queue theirs, runs the batch, and every thread then reads its own paths.  The batch is cleared by
   the last reader; calls that arrive meanwhile wait for that. */
static pthread_mutex_t chain_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t chain_cv_done = PTHREAD_COND_INITIALIZER, chain_cv_submit = PTHREAD_COND_INITIALIZER,
  chain_cv_idle = PTHREAD_COND_INITIALIZER;
static gmapchain_batch *chain_batch = NULL;
static bool chain_setup_done = false, chain_busy = false, chain_leader = false;
static int chain_npending = 0, chain_readers = 0;
static volatile int chain_inflight = 0;
static unsigned long chain_gen = 0, chain_ncalls = 0, chain_nbatches = 0, chain_nref = 0;
static long chain_wait_us = 300;

static void chain_report (void) {
  if (getenv("GMAP_SM100_STATS"))
    fprintf(stderr,"gmap.sm100 stage 2: %lu chaining calls on the device in %lu batches, %lu handed to the reference body\n",
	    chain_ncalls,chain_nbatches,chain_nref);
}

typedef struct { int n; int *qpos; uint32_t *gpos; } chain_path_t;

static List_T
chain_on_device (bool forwardp, CHAIN_PARAMS) {
  List_T all_paths = NULL, path;
  gmapdp_ctx *ctx = sm100_context();
  int id, npaths = 0, k, n, t, cap, cell[5];
  unsigned long mygen;
  chain_path_t *got = NULL;

  (void) firstactive; (void) nactive; (void) cellpool; (void) genome; (void) genomealt; (void) chroffset; (void) chrhigh; (void) plusp;
  cap = querylength + 16;
  __sync_fetch_and_add(&chain_inflight,1);

  pthread_mutex_lock(&chain_mu);
  if (chain_setup_done == false) {
    const char *w = getenv("GMAP_SM100_WAIT_US");
    if (w) chain_wait_us = atol(w);
    /* the values Stage2_setup stored (stage2.c:129-160, gmap.c:6544) */
    if (gmapchain_setup(ctx,splicingp,/*cross_species_p*/0,sufflookback,nsufflookback,maxintronlen) != GMAPDP_OK) {
      fprintf(stderr,"gmap.sm100: %s\n",gmapdp_last_error(ctx));
      exit(9);
    }
    chain_batch = GmapChain_batch_new(ctx);
    chain_setup_done = true;
    atexit(chain_report);
  }
  while (chain_busy) pthread_cond_wait(&chain_cv_idle,&chain_mu);	/* the previous generation is still being read */
  if (forwardp) {
    id = GmapChain_lookforward(chain_batch,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			       querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			       non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  } else {
    id = GmapChain_lookback(chain_batch,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			    querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			    non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  }
  if (id < 0) {
    fprintf(stderr,"gmap.sm100: %s\n",GmapChain_batch_error(chain_batch));
    exit(9);
  }
  chain_npending++;
  mygen = chain_gen;
  if (!chain_leader) {
    struct timespec dl;
    chain_leader = true;
    clock_gettime(CLOCK_REALTIME,&dl);
    dl.tv_nsec += chain_wait_us * 1000L;
    while (dl.tv_nsec >= 1000000000L) { dl.tv_nsec -= 1000000000L; dl.tv_sec++; }
    while (chain_npending < chain_inflight) {
      if (pthread_cond_timedwait(&chain_cv_submit,&chain_mu,&dl) != 0) break;
    }
    chain_busy = true;
    if (GmapChain_batch_run(chain_batch) != GMAPDP_OK) {
      fprintf(stderr,"gmap.sm100: %s\n",GmapChain_batch_error(chain_batch));
      exit(9);
    }
    chain_ncalls += chain_npending; chain_nbatches++;
    chain_readers = chain_npending; chain_npending = 0;
    chain_leader = false; chain_gen++;
    pthread_cond_broadcast(&chain_cv_done);
  } else {
    pthread_cond_signal(&chain_cv_submit);
    while (chain_gen == mygen) pthread_cond_wait(&chain_cv_done,&chain_mu);
  }
  /* copy this call's paths out of the shared batch */
  npaths = GmapChain_npaths(chain_batch,id);
  got = (chain_path_t *) MALLOC((npaths + 1) * sizeof(chain_path_t));
  for (k = 0; k < npaths; k++) {
    got[k].qpos = (int *) MALLOC(cap * sizeof(int));
    got[k].gpos = (uint32_t *) MALLOC(cap * sizeof(uint32_t));
    if ((got[k].n = GmapChain_path(chain_batch,id,k,cell,got[k].qpos,got[k].gpos,cap)) < 0) {
      fprintf(stderr,"gmap.sm100: a stage 2 path is longer than the query\n");
      exit(9);
    }
  }
  if (--chain_readers == 0) {
    GmapChain_batch_clear(chain_batch);
    chain_busy = false;
    pthread_cond_broadcast(&chain_cv_idle);
  }
  pthread_mutex_unlock(&chain_mu);
  __sync_fetch_and_sub(&chain_inflight,1);

  for (k = 0; k < npaths; k++) {
    /* traceback_one (stage2.c:4265-4272) conses pairs while it walks away from the cell, so the list head is the far
       end of the walk; GmapChain_path hands the pairs in list order, head first: push from the tail */
    n = got[k].n;
    path = (List_T) NULL;
    for (t = n - 1; t >= 0; t--) {
      path = Pairpool_push(path,pairpool,got[k].qpos[t],got[k].gpos[t],queryseq_ptr[got[k].qpos[t]],MATCH_COMP,
			   queryuc_ptr[got[k].qpos[t]],/*genomealt*/queryuc_ptr[got[k].qpos[t]],/*dynprogindex*/0);
    }
    all_paths = List_push(all_paths,(void *) path);	/* rank order, as stage2.c:4487 */
    FREE(got[k].gpos);
    FREE(got[k].qpos);
  }
  FREE(got);
  return all_paths;
}

/* GMAP_SM100_STAGE2=0 keeps stage 2 on the host (the reference's own body) */
static bool chain_enabled (void) {
  static int enabled = -1;
  if (enabled < 0) { const char *e = getenv("GMAP_SM100_STAGE2"); enabled = (e && e[0] == '0') ? 0 : 1; }
  return enabled == 1;
}

static List_T
sm100_lookback (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false || totalpositions <= 0 || !chain_enabled()) {
    __sync_fetch_and_add(&chain_nref,1);
    return ref_align_compute_lookback(CHAIN_ARGS);
  }
  return chain_on_device(/*forwardp*/false,CHAIN_ARGS);
}

static List_T
sm100_lookforward (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false || totalpositions <= 0 || !chain_enabled()) {
    __sync_fetch_and_add(&chain_nref,1);
    return ref_align_compute_lookforward(CHAIN_ARGS);
  }
  return chain_on_device(/*forwardp*/true,CHAIN_ARGS);
}
