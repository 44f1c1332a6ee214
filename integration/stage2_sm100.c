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
 * No call is ever handed to the reference's own body (it is compiled under the name ref_align_compute_* only because
 * the definition is part of stage2.c): what the device does not implement -- the cross-species canonical test
 * (use_canonical_p), SNP-tolerant mode, the 2-D link matrix that stage2.c itself asserts away (:3734) -- stops the
 * program with a message instead of silently running on the CPU.
 */
#include <pthread.h>
#include <time.h>
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
extern int sm100_device_of_group (int group);		/* dynprog_sm100.c: the devices of the process */

/* Chaining calls of the worker threads meet in shared device batches, like the DP calls (dynprog_sm100.c), but a
   chaining problem is a long sequential walk (milliseconds per problem on one warp), so what matters is that many
   batches are in flight at once.  The threads are spread over GMAP_SM100_CHAIN_GROUPS groups (default 8), each with
   its own engine context (own stream and device buffers; group g on device g mod #devices) and two batches: one is
   filled while the other runs or is being read.  The first thread that finds its batch not yet run and the group
   idle leads it: it runs the batch WITHOUT holding the group's lock; everybody then reads his own paths.  No
   timeouts; a call never waits for anything but the batch in front of its own. */
typedef struct chain_group {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  gmapdp_ctx *ctx;
  gmapchain_batch *batch[2];
  int fill;			/* the batch that accepts calls */
  int npending[2];		/* calls queued in batch k and not yet run */
  int readers[2];		/* calls of batch k that have not read their paths yet */
  unsigned long seq[2];		/* completed runs of batch k */
  bool running, inited;
  double t_run;			/* seconds spent in GmapChain_batch_run */
} chain_group;

#define CHAIN_MAXGROUPS 64
static chain_group chain_groups[CHAIN_MAXGROUPS];
static int chain_ngroups = 0;
static pthread_once_t chain_once = PTHREAD_ONCE_INIT;
static unsigned long chain_ncalls = 0, chain_nbatches = 0, chain_next_group = 0;
static __thread int chain_my_group = -1;

static double chain_t_init = 0.0;
static unsigned long chain_init_us = 0;
static double chain_now (void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

static void chain_report (void) {
  double t = 0.0;
  int g;
  if (getenv("GMAP_SM100_STATS") == NULL) return;
  for (g = 0; g < chain_ngroups; g++) if (chain_groups[g].inited) t += chain_groups[g].t_run;
  fprintf(stderr,"gmap.sm100 stage 2: %lu chaining calls on the device in %lu batches, %lu handed to the reference body\n",
	  chain_ncalls,chain_nbatches,0UL);
  fprintf(stderr,"gmap.sm100 stage 2 runtime: %d groups, mean batch latency %.1f us, start-up %.3f s\n",chain_ngroups,
	  chain_nbatches ? 1e6 * t / (double) chain_nbatches : 0.0,chain_t_init + 1e-6 * (double) chain_init_us);
}

static void chain_init (void) {
  const char *e = getenv("GMAP_SM100_CHAIN_GROUPS");
  const double t_begin = chain_now();
  int g;
  chain_ngroups = e ? atoi(e) : 8;
  if (chain_ngroups < 1) chain_ngroups = 1;
  if (chain_ngroups > CHAIN_MAXGROUPS) chain_ngroups = CHAIN_MAXGROUPS;
  for (g = 0; g < chain_ngroups; g++) {
    chain_group *G = &chain_groups[g];
    pthread_mutex_init(&G->mu,NULL);
    pthread_cond_init(&G->cv,NULL);
    G->inited = false;
  }
  chain_t_init = chain_now() - t_begin;
  atexit(chain_report);
}

/* a group's engine is created by the first call that lands on it (group mutex held): a run with few threads never
   pays for contexts it does not use */
static void chain_group_init (chain_group *G, int g) {
  const double t_begin = chain_now();
  /* the values Stage2_setup stored (stage2.c:129-160, gmap.c:6544) */
  if (gmapdp_create(&G->ctx,sm100_device_of_group(g)) != GMAPDP_OK ||
      gmapchain_setup(G->ctx,splicingp,/*cross_species_p*/0,sufflookback,nsufflookback,maxintronlen) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: %s\n",gmapdp_last_error(G->ctx));
    exit(9);
  }
  /* room for a few dozen queries of ordinary size per batch: device buffers are then not re-allocated mid-run */
  if (gmapchain_reserve(G->ctx,64,(size_t) 1 << 16,(size_t) 1 << 19) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: %s\n",gmapdp_last_error(G->ctx));
    exit(9);
  }
  G->batch[0] = GmapChain_batch_new(G->ctx); G->batch[1] = GmapChain_batch_new(G->ctx);
  G->fill = 0; G->npending[0] = G->npending[1] = 0; G->readers[0] = G->readers[1] = 0; G->seq[0] = G->seq[1] = 0;
  G->running = false; G->t_run = 0.0;
  G->inited = true;
  __sync_fetch_and_add(&chain_init_us,(unsigned long) (1e6 * (chain_now() - t_begin)));
}

typedef struct { int n; int *qpos; uint32_t *gpos; } chain_path_t;

static List_T
chain_on_device (bool forwardp, CHAIN_PARAMS) {
  List_T all_paths = NULL, path;
  chain_group *G;
  gmapchain_batch *B;
  int id, npaths = 0, k, n, t, cap, cell[5], kb;
  unsigned long myseq;
  chain_path_t *got = NULL;

  (void) firstactive; (void) nactive; (void) cellpool; (void) genome; (void) genomealt; (void) chroffset; (void) chrhigh; (void) plusp;
  cap = querylength + 16;
  pthread_once(&chain_once,chain_init);
  if (chain_my_group < 0) chain_my_group = (int) (__sync_fetch_and_add(&chain_next_group,1) % (unsigned long) chain_ngroups);
  G = &chain_groups[chain_my_group];

  pthread_mutex_lock(&G->mu);
  if (G->inited == false) chain_group_init(G,chain_my_group);
  while (G->readers[G->fill] > 0) pthread_cond_wait(&G->cv,&G->mu);	/* its previous run is still being read */
  kb = G->fill; B = G->batch[kb];
  if (forwardp) {
    id = GmapChain_lookforward(B,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			       querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			       non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  } else {
    id = GmapChain_lookback(B,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			    querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			    non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  }
  if (id < 0) {
    fprintf(stderr,"gmap.sm100: chaining call refused (%d): %s\n",id,GmapChain_batch_error(B));
    exit(9);
  }
  G->npending[kb]++;
  myseq = G->seq[kb];
  while (G->seq[kb] == myseq) {
    if (G->running == false && G->fill == kb) {
      /* lead: later calls go to the other batch while this one runs */
      int nrun = G->npending[kb];
      double t0;
      int rc;
      G->running = true; G->fill = kb ^ 1; G->npending[kb] = 0;
      pthread_mutex_unlock(&G->mu);
      t0 = chain_now();
      if ((rc = GmapChain_batch_run(B)) != GMAPDP_OK) {
	fprintf(stderr,"gmap.sm100: chaining batch failed (%d): %s\n",rc,GmapChain_batch_error(B));
	exit(9);
      }
      t0 = chain_now() - t0;
      pthread_mutex_lock(&G->mu);
      G->t_run += t0;
      __sync_fetch_and_add(&chain_ncalls,(unsigned long) nrun); __sync_fetch_and_add(&chain_nbatches,1UL);
      G->readers[kb] = nrun; G->seq[kb]++; G->running = false;
      pthread_cond_broadcast(&G->cv);
    } else {
      pthread_cond_wait(&G->cv,&G->mu);
    }
  }
  pthread_mutex_unlock(&G->mu);

  /* copy this call's paths out of the shared batch (read-only until its last reader clears it) */
  npaths = GmapChain_npaths(B,id);
  got = (chain_path_t *) MALLOC((npaths + 1) * sizeof(chain_path_t));
  for (k = 0; k < npaths; k++) {
    got[k].qpos = (int *) MALLOC(cap * sizeof(int));
    got[k].gpos = (uint32_t *) MALLOC(cap * sizeof(uint32_t));
    if ((got[k].n = GmapChain_path(B,id,k,cell,got[k].qpos,got[k].gpos,cap)) < 0) {
      fprintf(stderr,"gmap.sm100: a stage 2 path is longer than the query\n");
      exit(9);
    }
  }
  pthread_mutex_lock(&G->mu);
  if (--G->readers[kb] == 0) {
    GmapChain_batch_clear(B);
    pthread_cond_broadcast(&G->cv);
  }
  pthread_mutex_unlock(&G->mu);

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

/* a configuration the chaining engine does not implement: refuse loudly (there is no CPU fallback) */
static void chain_refuse (bool use_canonical_p, bool oned_matrix_p) {
  if (use_canonical_p == true)
    fprintf(stderr,"gmap.sm100: --cross-species (use_canonical_p: stage2.c:1252,1816,2195,2750) is not served by the device chaining engine\n");
  else if (snps_p == true)
    fprintf(stderr,"gmap.sm100: SNP-tolerant alignment (-v) is not served by the device chaining engine\n");
  else if (oned_matrix_p == false)
    fprintf(stderr,"gmap.sm100: the 2-D link matrix (oned_matrix_p == false) is not served by the device chaining engine\n");
  exit(9);
}

static List_T
sm100_lookback (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false) chain_refuse(use_canonical_p,oned_matrix_p);
  if (totalpositions <= 0) return (List_T) NULL;	/* no hit, no cell: stage2.c:4464 (ncells == 0) */
  return chain_on_device(/*forwardp*/false,CHAIN_ARGS);
}

static List_T
sm100_lookforward (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false) chain_refuse(use_canonical_p,oned_matrix_p);
  if (totalpositions <= 0) return (List_T) NULL;	/* stage2.c:5124 */
  return chain_on_device(/*forwardp*/true,CHAIN_ARGS);
}

/* keeps the (never called) reference bodies referenced, so that -Wunused does not hide a renamed seam */
void *sm100_stage2_reference_bodies[2] = { (void *) ref_align_compute_lookback, (void *) ref_align_compute_lookforward };
