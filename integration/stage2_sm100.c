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

static pthread_mutex_t chain_mu = PTHREAD_MUTEX_INITIALIZER;	/* one chaining batch on the device at a time */
static gmapchain_batch *chain_batch = NULL;
static bool chain_setup_done = false;
static unsigned long chain_ncalls = 0, chain_nref = 0;

static void chain_report (void) {
  if (getenv("GMAP_SM100_STATS"))
    fprintf(stderr,"gmap.sm100 stage 2: %lu chaining calls on the device, %lu handed to the reference body\n",chain_ncalls,chain_nref);
}

static List_T
chain_on_device (bool forwardp, CHAIN_PARAMS) {
  List_T all_paths = NULL, path;
  gmapdp_ctx *ctx = sm100_context();
  int id, npaths, k, n, t, cap, cell[5];
  int *qpos;
  uint32_t *gpos;

  (void) firstactive; (void) nactive; (void) cellpool; (void) genome; (void) genomealt; (void) chroffset; (void) chrhigh; (void) plusp;
  cap = querylength + 16;
  qpos = (int *) MALLOC(cap * sizeof(int));
  gpos = (uint32_t *) MALLOC(cap * sizeof(uint32_t));

  pthread_mutex_lock(&chain_mu);
  if (chain_setup_done == false) {
    /* the values Stage2_setup stored (stage2.c:129-160, gmap.c:6544) */
    if (gmapchain_setup(ctx,splicingp,/*cross_species_p*/0,sufflookback,nsufflookback,maxintronlen) != GMAPDP_OK) {
      fprintf(stderr,"gmap.sm100: %s\n",gmapdp_last_error(ctx));
      exit(9);
    }
    chain_batch = GmapChain_batch_new(ctx);
    chain_setup_done = true;
    atexit(chain_report);
  }
  GmapChain_batch_clear(chain_batch);
  if (forwardp) {
    id = GmapChain_lookforward(chain_batch,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			       querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			       non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  } else {
    id = GmapChain_lookback(chain_batch,(uint32_t *const *) mappings,npositions,totalpositions,minactive,maxactive,
			    querylength,querystart,queryend,indexsize,localp,skip_repetitive_p,use_canonical_p,
			    non_canonical_penalty,favor_right_p,middlep,max_nalignments);
  }
  if (id < 0 || GmapChain_batch_run(chain_batch) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: %s\n",GmapChain_batch_error(chain_batch));
    exit(9);
  }
  chain_ncalls++;
  npaths = GmapChain_npaths(chain_batch,id);
  for (k = 0; k < npaths; k++) {
    if ((n = GmapChain_path(chain_batch,id,k,cell,qpos,gpos,cap)) < 0) {
      fprintf(stderr,"gmap.sm100: a stage 2 path is longer than the query\n");
      exit(9);
    }
    /* traceback_one (stage2.c:4265-4272) conses pairs while it walks away from the cell, so the list head is the far
       end of the walk; GmapChain_path hands the pairs in list order, head first: push from the tail */
    path = (List_T) NULL;
    for (t = n - 1; t >= 0; t--) {
      path = Pairpool_push(path,pairpool,qpos[t],gpos[t],queryseq_ptr[qpos[t]],MATCH_COMP,
			   queryuc_ptr[qpos[t]],/*genomealt*/queryuc_ptr[qpos[t]],/*dynprogindex*/0);
    }
    all_paths = List_push(all_paths,(void *) path);	/* rank order, as stage2.c:4487 */
  }
  pthread_mutex_unlock(&chain_mu);

  FREE(gpos);
  FREE(qpos);
  return all_paths;
}

static List_T
sm100_lookback (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false || totalpositions <= 0) {
    __sync_fetch_and_add(&chain_nref,1);
    return ref_align_compute_lookback(CHAIN_ARGS);
  }
  return chain_on_device(/*forwardp*/false,CHAIN_ARGS);
}

static List_T
sm100_lookforward (CHAIN_PARAMS) {
  if (use_canonical_p == true || snps_p == true || oned_matrix_p == false || totalpositions <= 0) {
    __sync_fetch_and_add(&chain_nref,1);
    return ref_align_compute_lookforward(CHAIN_ARGS);
  }
  return chain_on_device(/*forwardp*/true,CHAIN_ARGS);
}
