/* dynprog_sm100.c -- the stage3-side binding of the B200 DP engine.
 *
 * Compiled with the REFERENCE's headers and linked with the reference's unmodified objects; it
 * provides the five DP entry points that stage3.c calls (dynprog_single.h:23, dynprog_genome.h:23,
 * dynprog_cdna.h:17, dynprog_end.h:24,46) on top of include/gmapdp_shim.h.  The reference's own
 * definitions of those five symbols are renamed to ref_* in the object files by the build recipe
 * (integration/Makefile), nothing in /root/reference is edited or copied.
 *
 * The DP calls of all worker threads meet in shared device batches (the streaming runtime of
 * include/gmapdp_stream.h).  What the binding does on the host is exactly what
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
#include <time.h>
#include "iit-read.h"
#include "splicetrie_build.h"
#include "mode.h"
#include "gmapdp_shim.h"
#include "gmapdp_stream.h"

/* ---- the engine: one streaming runtime for the whole process (include/gmapdp_stream.h) ------------------
 * Every worker thread prepares its call in a private one-call batch (argument checks, the *_simple shortcuts,
 * penalties, bands: gmapdp_shim.h), submits the device box -- if there is one -- to the runtime, sleeps until its
 * flight is back, and replays its own edit script into pairs.  Calls of different threads share flights; nothing
 * but the submission slot is serialised.  GMAP_SM100_DEVICES="0,1,2,3" (or "all") gives one lane per GPU: a worker
 * thread stays on one lane, so a query's dependent chain of calls stays on one device (gmap.c:4895-4907 keeps
 * all per-worker state for the life of the thread). */
static gmapdp_stream *sm100_stream = NULL;
static pthread_once_t sm100_once = PTHREAD_ONCE_INIT;
static pthread_key_t sm100_key;
static int sm100_user_open = 0, sm100_user_extend = 0, sm100_user_dynprog_p = 0;
static int sm100_devices[64], sm100_ndevices = 0;
static unsigned long sm100_ncalls = 0, sm100_nhost = 0;
static double sm100_t0 = 0.0, sm100_t_init = 0.0;

static double now_s (void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

/* GMAP_SM100_STATS: where the host time of the five entry points goes (time-stamp counter, summed over all threads):
   0 segment fetch + MaxEnt, 1 call preparation (shim), 2 submit, 3 asleep waiting for the flight, 4 script replay, 5 pairs -> List_T */
#define SM100_NPROF 6
static int sm100_prof_on = 0;
static unsigned long long sm100_prof[SM100_NPROF], sm100_tsc0 = 0;
static __thread unsigned long long sm100_prof_last;
static inline unsigned long long sm100_tsc (void) { unsigned int lo, hi; __asm__ __volatile__("rdtsc" : "=a"(lo), "=d"(hi)); return ((unsigned long long) hi << 32) | lo; }
#define PROF_BEGIN() do { if (sm100_prof_on) sm100_prof_last = sm100_tsc(); } while (0)
#define PROF_MARK(k) do { if (sm100_prof_on) { unsigned long long t_ = sm100_tsc(); __sync_fetch_and_add(&sm100_prof[k],t_ - sm100_prof_last); sm100_prof_last = t_; } } while (0)

static void sm100_die (const char *what) {
  fprintf(stderr,"gmap.sm100: %s\n",what);
  exit(9);
}

static void sm100_report_genome (void);
static void sm100_report (void) {
  double st[GMAPDP_STREAM_NSTATS];
  if (getenv("GMAP_SM100_STATS") == NULL || sm100_stream == NULL) return;
  gmapdp_stream_stats(sm100_stream,st);
  fprintf(stderr,"gmap.sm100: %lu DP calls in %.0f device batches (%.1f calls per batch)\n",sm100_ncalls,st[1],
	  st[1] > 0.0 ? (double) sm100_ncalls / st[1] : 0.0);
  fprintf(stderr,"gmap.sm100 runtime: %d device(s), %.0f boxes on the device, %lu calls resolved on the host, largest batch %.0f, "
	  "mean batch latency %.1f us, mean wait per box %.1f us, %.0f kernel launches, %.1f MB up, %.1f MB down, %.3f s since first call\n",
	  gmapdp_stream_ndevices(sm100_stream),st[0],sm100_nhost,st[2],st[1] > 0.0 ? 1e6 * st[4] / st[1] : 0.0,
	  st[0] > 0.0 ? 1e6 * st[5] / st[0] : 0.0,st[6],st[7] / 1e6,st[8] / 1e6,now_s() - sm100_t0);
  if (sm100_prof_on) {
    const double hz = (double) (sm100_tsc() - sm100_tsc0) / (now_s() - sm100_t0);
    fprintf(stderr,"gmap.sm100 host time in the entry points (thread-seconds): fetch+MaxEnt %.2f, prepare %.2f, submit %.2f, asleep %.2f, replay %.2f, list %.2f\n",
	    sm100_prof[0] / hz,sm100_prof[1] / hz,sm100_prof[2] / hz,sm100_prof[3] / hz,sm100_prof[4] / hz,sm100_prof[5] / hz);
  }
  sm100_report_genome();
  fprintf(stderr,"gmap.sm100 start-up: %.3f s to create the runtime (contexts, pinned staging, device buffers)\n",sm100_t_init);
  fprintf(stderr,"gmap.sm100 service threads: scheduling mode %d (1 FIFO, 2 short slices, 3 default)\n",gmapdp_stream_sched_mode(sm100_stream));
  {
    int lane, nl = gmapdp_stream_nlanes(sm100_stream);
    double ls[11];
    for (lane = 0; lane < nl; lane++) {
      gmapdp_stream_lane_stats(sm100_stream,lane,ls);
      if (ls[1] > 0.0)
	fprintf(stderr,"gmap.sm100 lane %d: %.0f boxes in %.0f batches (largest %.0f), batch latency %.1f us, device %.1f us, upload %.1f us; "
		"host stages: close->launch %.1f, launch call %.1f, launched->seen %.1f (device begun after %.1f), root wake-ups %.1f us\n",
		lane,ls[0],ls[1],ls[2],ls[3],ls[4],ls[5],ls[6],ls[7],ls[8],ls[10],ls[9]);
    }
  }
}


static void sm100_free_thread_fwd (void *p);
static void sm100_init (void) {
  const char *dev = getenv("GMAP_SM100_DEVICES");
  int devices[64], n = 0, rc;
  if (dev == NULL) dev = getenv("GMAP_SM100_DEVICE");
  if (dev == NULL || dev[0] == '\0') devices[n++] = 0;
  else if (!strcmp(dev,"all")) {
    int count = gmapdp_device_count();
    if (count <= 0) sm100_die("no CUDA device: the DP engine has no CPU fallback");
    for (n = 0; n < count && n < 64; n++) devices[n] = n;
  } else {
    const char *q = dev;
    while (*q && n < 64) {
      devices[n++] = atoi(q);
      while (*q && *q != ',') q++;
      if (*q == ',') q++;
    }
  }
  memcpy(sm100_devices,devices,sizeof(devices)); sm100_ndevices = n;
  sm100_t_init = now_s();
  if ((rc = gmapdp_stream_create(&sm100_stream,devices,n,0)) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: %s\n",gmapdp_stream_error(sm100_stream));
    exit(9);
  }
  pthread_key_create(&sm100_key,sm100_free_thread_fwd);
  sm100_t0 = now_s(); sm100_tsc0 = sm100_tsc();
  sm100_t_init = sm100_t0 - sm100_t_init;
  sm100_prof_on = getenv("GMAP_SM100_STATS") != NULL;
  atexit(sm100_report);
}

/* The runtime is created on a background thread started from Dynprog_init (gmap.c:6348: early in main, before the
   genome index is loaded), so that CUDA context creation and the pinned allocations overlap the program's own start-up;
   the first DP call waits for it. */
static pthread_t sm100_init_thread;
static int sm100_init_started = 0;
static void *sm100_init_main (void *arg) { (void) arg; pthread_once(&sm100_once,sm100_init); return NULL; }
static void sm100_start_init (void) {
  if (__sync_bool_compare_and_swap(&sm100_init_started,0,1)) {
    if (pthread_create(&sm100_init_thread,NULL,sm100_init_main,NULL) != 0) sm100_init_started = 2;
    else pthread_detach(sm100_init_thread);
  }
}
static void sm100_wait_ready (void) { pthread_once(&sm100_once,sm100_init); }	/* joins the background thread's once, or runs it */

/* the devices of the process, for the chaining engine's groups (stage2_sm100.c): group g lives on device g mod n */
int sm100_device_of_group (int group) {
  pthread_once(&sm100_once,sm100_init);
  return sm100_devices[group % sm100_ndevices];
}

/* ---- set-up calls: what the five entry points read from the reference's module statics ---------------------
 * gmap.c:6348 Dynprog_init(mode), :6547 Dynprog_single_setup, :6548 Dynprog_genome_setup, :6551 Dynprog_end_setup.
 * The reference's own functions still run (their statics serve the entry points that stay on the host:
 * Dynprog_microexon_int, Dynprog_end5/3_known, *_splicejunction); here the values the DEVICE path depends on are
 * recorded, and the settings it does not implement are refused loudly instead of being ignored. */
extern void ref_Dynprog_init (Mode_T mode);
extern void ref_Dynprog_single_setup (int user_open_in, int user_extend_in, bool user_dynprog_p_in, bool homopolymerp_in);
extern void ref_Dynprog_genome_setup (bool novelsplicingp_in, IIT_T splicing_iit_in, int *splicing_divint_crosstable_in,
				      int donor_typeint_in, int acceptor_typeint_in, int user_open_in, int user_extend_in, bool user_dynprog_p_in);
extern void ref_Dynprog_end_setup (Univcoord_T *splicesites_in, Splicetype_T *splicetypes_in, Chrpos_T *splicedists_in, int nsplicesites_in,
				   Trieoffset_T *trieoffsets_obs_in, Triecontent_T *triecontents_obs_in,
				   Trieoffset_T *trieoffsets_max_in, Triecontent_T *triecontents_max_in,
				   int user_open_in, int user_extend_in, bool user_dynprog_p_in);

void
Dynprog_init (Mode_T mode) {
  if (mode != STANDARD) sm100_die("--mode other than standard (cmet / atoi / ttoc) changes the DP score tables; this build serves the standard tables only");
  ref_Dynprog_init(mode);
  sm100_start_init();
}

static void sm100_user_penalties (int user_open_in, int user_extend_in, bool user_dynprog_p_in) {
  sm100_user_open = user_open_in; sm100_user_extend = user_extend_in; sm100_user_dynprog_p = user_dynprog_p_in ? 1 : 0;
}

void
Dynprog_single_setup (int user_open_in, int user_extend_in, bool user_dynprog_p_in, bool homopolymerp_in) {
  if (homopolymerp_in == true) sm100_die("--homopolymer (dynprog_single.c:535) is not served by the device DP; refusing to give different alignments");
  sm100_user_penalties(user_open_in,user_extend_in,user_dynprog_p_in);
  ref_Dynprog_single_setup(user_open_in,user_extend_in,user_dynprog_p_in,homopolymerp_in);
}

void
Dynprog_genome_setup (bool novelsplicingp_in, IIT_T splicing_iit_in, int *splicing_divint_crosstable_in,
		      int donor_typeint_in, int acceptor_typeint_in, int user_open_in, int user_extend_in, bool user_dynprog_p_in) {
  if (splicing_iit_in != NULL) sm100_die("known splice sites (-s / --use-splicing: the intron-level bridge of dynprog_genome.c:615,1489) are not served by the device DP");
  sm100_user_penalties(user_open_in,user_extend_in,user_dynprog_p_in);
  ref_Dynprog_genome_setup(novelsplicingp_in,splicing_iit_in,splicing_divint_crosstable_in,donor_typeint_in,acceptor_typeint_in,
			   user_open_in,user_extend_in,user_dynprog_p_in);
}

void
Dynprog_end_setup (Univcoord_T *splicesites_in, Splicetype_T *splicetypes_in, Chrpos_T *splicedists_in, int nsplicesites_in,
		   Trieoffset_T *trieoffsets_obs_in, Triecontent_T *triecontents_obs_in,
		   Trieoffset_T *trieoffsets_max_in, Triecontent_T *triecontents_max_in,
		   int user_open_in, int user_extend_in, bool user_dynprog_p_in) {
  sm100_user_penalties(user_open_in,user_extend_in,user_dynprog_p_in);
  ref_Dynprog_end_setup(splicesites_in,splicetypes_in,splicedists_in,nsplicesites_in,trieoffsets_obs_in,triecontents_obs_in,
			trieoffsets_max_in,triecontents_max_in,user_open_in,user_extend_in,user_dynprog_p_in);
}

/* per worker thread: a private one-call batch (all Dynprog_T of a gmap run have the same limits, gmap.c:4898-4903)
   and the mailbox its boxes' results land in */
typedef struct sm100_thread {
  gmapdp_batch *batch;
  gmapdp_mailbox mailbox;
  int has_genome;		/* the batch knows the resident genome (GmapDP_batch_genome) */
} sm100_thread;

static void sm100_free_thread_fwd (void *p);
static void sm100_free_thread (void *p) {
  sm100_thread *t = (sm100_thread *) p;
  if (t) { GmapDP_batch_free(t->batch); free(t->mailbox.ops); free(t); }
}

static void sm100_free_thread_fwd (void *p) { sm100_free_thread(p); }

static sm100_thread *my_thread (Dynprog_T dynprog) {
  sm100_thread *t;
  sm100_wait_ready();
  t = (sm100_thread *) pthread_getspecific(sm100_key);
  if (t == NULL) {
    t = (sm100_thread *) calloc(1,sizeof(sm100_thread));
    if ((t->batch = GmapDP_batch_new(NULL,dynprog->max_rlength,dynprog->max_glength)) == NULL) sm100_die("Dynprog_T limits beyond 32767");
    if (GmapDP_batch_user_dynprog(t->batch,sm100_user_open,sm100_user_extend,sm100_user_dynprog_p) != GMAPDP_OK) sm100_die(GmapDP_batch_error(t->batch));
    t->mailbox.ops_cap = 2 * ((size_t) dynprog->max_rlength + (size_t) dynprog->max_glength) + 64;
    t->mailbox.ops = (uint32_t *) malloc(t->mailbox.ops_cap * sizeof(uint32_t));
    pthread_setspecific(sm100_key,t);
  }
  GmapDP_batch_clear(t->batch);
  return t;
}
static gmapdp_batch *my_batch (Dynprog_T dynprog) { return my_thread(dynprog)->batch; }
static sm100_thread *cur_thread (void) { return (sm100_thread *) pthread_getspecific(sm100_key); }	/* after my_batch */

/* ------------------------------------------------------------------------------------------------
 * Resident genome.  The first DP call hands the process's genome -- the reference's own
 * compressed blocks, Genome_blocks(genome), read in place -- and the MaxEnt tables (maxent_sm100.c) to the runtime, which
 * copies them to every device once.  From then on a call announces genome COORDINATES (GmapDP_batch_next_coords) and
 * the device decodes the segments and evaluates Maxent_hr_*_prob itself: nothing genomic is uploaded per call, and the
 * binding no longer computes the probability arrays of dynprog_genome.c:970-1061 on the host.  The characters are
 * still fetched here: the pair lists carry them.  Only when genomealt == genome (no SNP-tolerant alignment) and the
 * blocks are in memory; GMAP_SM100_RESIDENT=0 keeps the upload form.
 * ---------------------------------------------------------------------------------------------- */
extern void sm100_maxent_tables (gmapdp_maxent_tables *t);
static pthread_mutex_t sm100_genome_mu = PTHREAD_MUTEX_INITIALIZER;
static const Genomecomp_T *sm100_gblocks = NULL;
static size_t sm100_gwords = 0;
static volatile int sm100_genome_state = 0;		/* 0 not tried yet, 1 resident, -1 not available */
static unsigned long sm100_ncoords = 0;			/* calls announced as genome coordinates */
static gmapdp_maxent_tables sm100_me;

static bool sm100_resident (sm100_thread *t, Genome_T genome, Genome_T genomealt) {
  const Genomecomp_T *blocks;
  if (sm100_genome_state < 0 || genome == NULL) return false;
  blocks = Genome_blocks(genome);
  if (blocks == NULL || (genomealt != genome && Genome_blocks(genomealt) != blocks)) return false;
  if (sm100_genome_state == 0) {
    pthread_mutex_lock(&sm100_genome_mu);
    if (sm100_genome_state == 0) {
      const char *e = getenv("GMAP_SM100_RESIDENT");
      if (e && atoi(e) == 0) sm100_genome_state = -1;
      else {
	const size_t nwords = (((size_t) Genome_genomelength(genome) + 31) / 32) * 3;
	sm100_maxent_tables(&sm100_me);
	if (gmapdp_stream_genome(sm100_stream,(const uint32_t *) blocks,nwords,&sm100_me) == GMAPDP_OK) {
	  sm100_gblocks = blocks; sm100_gwords = nwords;
	  __sync_synchronize();
	  sm100_genome_state = 1;
	} else {
	  fprintf(stderr,"gmap.sm100: genome not made resident (%s): genomic segments are uploaded per call\n",gmapdp_stream_error(sm100_stream));
	  sm100_genome_state = -1;
	}
      }
    }
    pthread_mutex_unlock(&sm100_genome_mu);
  }
  if (sm100_genome_state != 1 || blocks != sm100_gblocks) return false;
  if (!t->has_genome) { GmapDP_batch_genome(t->batch,(const uint32_t *) sm100_gblocks,sm100_gwords,&sm100_me); t->has_genome = 1; }
  __sync_fetch_and_add(&sm100_ncoords,1);
  return true;
}

static void sm100_report_genome (void) {
  fprintf(stderr,"gmap.sm100 genome: %s, %lu calls sent as coordinates (%.1f MB of blocks per device)\n",
	  sm100_genome_state == 1 ? "resident on the device" : "uploaded per call",sm100_ncoords,4e-6 * (double) sm100_gwords);
}

/* where the arrays of the fetches below lie in the genome: "forward" = the segment fetched from goffset on
   (Genome_get_segment_right on the Watson strand, _left + revcomp on the Crick strand), "backward" = the segment that
   ends at rev_goffset */
static void sm100_seg_forward (bool watsonp, Univcoord_T chroffset, Univcoord_T chrhigh, int goffset, uint32_t *gpos, int *neg, int *left) {
  if (watsonp) { *gpos = (uint32_t) (chroffset + goffset); *neg = 0; *left = 0; }
  else { *gpos = (uint32_t) (chrhigh - goffset); *neg = 1; *left = 1; }
}
static void sm100_seg_backward (bool watsonp, Univcoord_T chroffset, Univcoord_T chrhigh, int rev_goffset, int glength, uint32_t *gpos, int *neg, int *left) {
  if (watsonp) { *gpos = (uint32_t) (chroffset + rev_goffset + 1 - glength); *neg = 0; *left = 1; }
  else { *gpos = (uint32_t) (chrhigh - rev_goffset + glength - 1); *neg = 1; *left = 0; }
}

/* runs the call queued in this thread's batch; afterwards GmapDP_result_view(b,id,...) has the result */
static void run_call (gmapdp_batch *b) {
  const gmapdp_box *boxes; const uint8_t *seq; const double *probs;
  sm100_thread *t = (sm100_thread *) pthread_getspecific(sm100_key);
  size_t seqbytes, nprobs;
  int nboxes;

  __sync_fetch_and_add(&sm100_ncalls,1);
  if (GmapDP_batch_device_view(b,&boxes,&nboxes,&seq,&seqbytes,&probs,&nprobs) != GMAPDP_OK) sm100_die(GmapDP_batch_error(b));
  if (nboxes == 0) { __sync_fetch_and_add(&sm100_nhost,1); return; }	/* resolved by the entry point's own shortcuts */
  if (gmapdp_stream_submit(sm100_stream,&boxes[0],seq,seqbytes,probs,nprobs,&t->mailbox) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: DP submit: %s\n",gmapdp_stream_error(sm100_stream)); exit(9);
  }
  PROF_MARK(2);
  if (gmapdp_stream_wait(sm100_stream,&t->mailbox) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: DP wait: %s\n",gmapdp_stream_error(sm100_stream)); exit(9);
  }
  PROF_MARK(3);
  if (GmapDP_batch_complete(b,&t->mailbox.result,t->mailbox.ops) != GMAPDP_OK) {
    fprintf(stderr,"gmap.sm100: DP replay: %s\n",GmapDP_batch_error(b)); exit(9);
  }
  PROF_MARK(4);
}

/* records come head first: cons them from the tail (compact 8-byte records, gmapdp_shim.h: positions relative to the side
   of the call the pair lies on, the side in the top bit of comp; gap holders carry an index into the call's gap table,
   ordinary pairs the call's dynprogindex) */
static List_T pairs_to_list (Pairpool_T pool, const gmapdp_cpair *p, const gmapdp_gapinfo *gaps, const gmapdp_pairbase *base, int dpi, int n) {
  List_T l = NULL;
  Pair_T g;
  int k;
  for (k = n - 1; k >= 0; k--) {
    if (p[k].qrel == GMAPDP_CPAIR_GAP) {
      const gmapdp_gapinfo *gi = &gaps[p[k].grel];
      l = Pairpool_push_gapholder(l,pool,gi->queryjump,gi->genomejump,/*leftpair*/NULL,/*rightpair*/NULL,/*knownp*/false);
      g = (Pair_T) List_head(l);
      g->introntype = gi->introntype; g->donor_prob = gi->donor_prob; g->acceptor_prob = gi->acceptor_prob;
    } else {
      const int side = ((unsigned char) p[k].comp >> 7) & 1;
      l = Pairpool_push(l,pool,base->q[side] + (int) p[k].qrel,base->g[side] + (int) p[k].grel,p[k].cdna,(char) (p[k].comp & 0x7f),
			p[k].genome,p[k].genomealt,dpi);
    }
  }
  PROF_MARK(5);
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
  int id, n;
  const int *iout;
  const gmapdp_cpair *pairs;
  const gmapdp_pairbase *base;
  const gmapdp_gapinfo *gaps;
  int dpi;
  gmapdp_batch *b = my_batch(dynprog);
  List_T l;
  bool fetch = (rlength > 0 && glength > 0 && rlength <= dynprog->max_rlength && glength <= dynprog->max_glength);

  PROF_BEGIN();
  if (fetch) {
    gseq = (char *) malloc(glength + 1); galt = (char *) malloc(glength + 1);
    if (watsonp) Genome_get_segment_right(gseq,galt,genome,genomealt,chroffset+goffset,glength,chrhigh,/*revcomp*/false);
    else Genome_get_segment_left(gseq,galt,genome,genomealt,chrhigh-goffset+1,glength,chroffset,/*revcomp*/true);
  } else {
    gseq = galt = empty;
  }
  if (fetch && sm100_resident(cur_thread(),genome,genomealt)) {
    gmapdp_coords co;
    memset(&co,0,sizeof(co));
    co.chroffset = (uint32_t) chroffset; co.chrhigh = (uint32_t) chrhigh;
    sm100_seg_forward(watsonp,chroffset,chrhigh,goffset,&co.gposL,&co.negL,&co.leftL);
    GmapDP_batch_next_coords(b,&co);
  }
  PROF_MARK(0);
  id = GmapDP_single_gap(b,*dynprogindex,rsequence,rsequenceuc,rlength,glength,roffset,goffset,gseq,galt,
			 jump_late_p,extraband_single,widebandp,defect_rate);
  PROF_MARK(1);
  run_call(b);
  n = GmapDP_result_view(b,id,&iout,NULL,&pairs,&gaps,&base,&dpi);
  *dynprogindex = iout[0]; *traceback_score = iout[1]; *nmatches = iout[2]; *nmismatches = iout[3];
  *nopens = iout[4]; *nindels = iout[5];
  if (fetch) { free(galt); free(gseq); }
  l = (n < 0) ? (List_T) NULL : pairs_to_list(pairpool,pairs,gaps,base,dpi,n);
  return l;
}

static List_T
end_gap (bool end5, int *dynprogindex, int *traceback_score, int *nmatches, int *nmismatches, int *nopens, int *nindels,
	 Dynprog_T dynprog, char *rseq, char *rsequc, int rlength, int glength, int roffset, int goffset,
	 Univcoord_T chroffset, Univcoord_T chrhigh, bool watsonp, bool jump_late_p,
	 Genome_T genome, Genome_T genomealt, Pairpool_T pairpool, int extraband_end, double defect_rate,
	 Endalign_T endalign, bool require_pos_score_p) {
  char *gseq, *galt, empty[1] = {'\0'};
  int id, n, gl = glength;
  const int *iout;
  const gmapdp_cpair *pairs;
  const gmapdp_pairbase *base;
  const gmapdp_gapinfo *gaps;
  int dpi;
  gmapdp_batch *b = my_batch(dynprog);
  bool fetch;

  /* the reference chops before it fetches (dynprog_end.c:1357-1378 / :1986-1999) */
  if (endalign != QUERYEND_NOGAPS && gl > dynprog->max_glength) gl = dynprog->max_glength;
  fetch = (rlength > 0 && gl > 0 && !(end5 && goffset < 0));
  PROF_BEGIN();
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
  if (fetch && sm100_resident(cur_thread(),genome,genomealt)) {
    gmapdp_coords co;
    memset(&co,0,sizeof(co));
    co.chroffset = (uint32_t) chroffset; co.chrhigh = (uint32_t) chrhigh;
    if (end5) sm100_seg_backward(watsonp,chroffset,chrhigh,goffset,gl,&co.gposL,&co.negL,&co.leftL);
    else sm100_seg_forward(watsonp,chroffset,chrhigh,goffset,&co.gposL,&co.negL,&co.leftL);
    GmapDP_batch_next_coords(b,&co);
  }
  PROF_MARK(0);
  if (end5) id = GmapDP_end5_gap(b,*dynprogindex,rseq,rsequc,rlength,glength,roffset,goffset,gseq,galt,jump_late_p,
				 extraband_end,defect_rate,(int) endalign,require_pos_score_p);
  else id = GmapDP_end3_gap(b,*dynprogindex,rseq,rsequc,rlength,glength,roffset,goffset,gseq,galt,jump_late_p,
			    extraband_end,defect_rate,(int) endalign,require_pos_score_p);
  PROF_MARK(1);
  run_call(b);
  n = GmapDP_result_view(b,id,&iout,NULL,&pairs,&gaps,&base,&dpi);
  *dynprogindex = iout[0]; *traceback_score = iout[1]; *nmatches = iout[2]; *nmismatches = iout[3];
  *nopens = iout[4]; *nindels = iout[5];
  if (fetch) { free(galt); free(gseq); }
  return (n < 0) ? (List_T) NULL : pairs_to_list(pairpool,pairs,gaps,base,dpi,n);
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
  const gmapdp_cpair *pairs;
  const gmapdp_pairbase *base;
  const gmapdp_gapinfo *gaps;
  int dpi;
  gmapdp_batch *b = my_batch(dynprogL);
  int id, n, c;
  Univcoord_T pos;
  bool fetch = (rlength > 1 && rlength <= dynprogL->max_rlength && glengthL <= dynprogL->max_glength &&
		rlength <= dynprogR->max_rlength && glengthR <= dynprogR->max_glength && glengthL > 0 && glengthR > 0);

  PROF_BEGIN();
  if (fetch) {
    gL = (char *) malloc(glengthL + 1); gLa = (char *) malloc(glengthL + 1);
    gR = (char *) malloc(glengthR + 1); gRa = (char *) malloc(glengthR + 1);
    if (watsonp) {
      Genome_get_segment_right(gL,gLa,genome,genomealt,chroffset+goffsetL,glengthL,chrhigh,/*revcomp*/false);
      Genome_get_segment_left(gR,gRa,genome,genomealt,chroffset+rev_goffsetR+1,glengthR,chroffset,/*revcomp*/false);
    } else {
      Genome_get_segment_left(gL,gLa,genome,genomealt,chrhigh-goffsetL+1,glengthL,chroffset,/*revcomp*/true);
      Genome_get_segment_right(gR,gRa,genome,genomealt,chrhigh-rev_goffsetR,glengthR,chrhigh,/*revcomp*/true);
    }
    if (sm100_resident(cur_thread(),genome,genomealt)) {
      /* the probability arrays of dynprog_genome.c:970-1061 as coordinates: entry c of the left / right array is the
	 probability of the given kind at probpos +- c; the device evaluates them (gmapdp_genome.h) */
      gmapdp_coords co;
      memset(&co,0,sizeof(co));
      co.chroffset = (uint32_t) chroffset; co.chrhigh = (uint32_t) chrhigh;
      sm100_seg_forward(watsonp,chroffset,chrhigh,goffsetL,&co.gposL,&co.negL,&co.leftL);
      sm100_seg_backward(watsonp,chroffset,chrhigh,rev_goffsetR,glengthR,&co.gposR,&co.negR,&co.leftR);
      co.probs = 1;
      if (watsonp) {
	co.probposL = (uint32_t) (chroffset + goffsetL); co.probnegL = 0; co.probkindL = (cdna_direction > 0) ? 0 : 3;
	co.probposR = (uint32_t) (chroffset + rev_goffsetR + 1); co.probnegR = 1; co.probkindR = (cdna_direction > 0) ? 1 : 2;
      } else {
	co.probposL = (uint32_t) (chrhigh - goffsetL + 1); co.probnegL = 1; co.probkindL = (cdna_direction > 0) ? 2 : 1;
	co.probposR = (uint32_t) (chrhigh - rev_goffsetR); co.probnegR = 0; co.probkindR = (cdna_direction > 0) ? 3 : 0;
      }
      GmapDP_batch_next_coords(b,&co);
    } else {
    lp = (double *) calloc(glengthL + 1,sizeof(double)); rp = (double *) calloc(glengthR + 1,sizeof(double));
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
    }
  } else {
    gL = gLa = gR = gRa = empty;
  }
  PROF_MARK(0);
  id = GmapDP_genome_gap(b,*dynprogindex,rsequence,rsequenceuc,rlength,glengthL,glengthR,roffset,goffsetL,rev_goffsetR,
			 gL,gLa,gR,gRa,lp,rp,cdna_direction,jump_late_p,extraband_paired,defect_rate,maxpeelback,halfp,finalp);
  PROF_MARK(1);
  run_call(b);
  n = GmapDP_result_view(b,id,&iout,&dout,&pairs,&gaps,&base,&dpi);
  *dynprogindex = iout[0];
  SET(new_leftgenomepos,iout[1]); SET(new_rightgenomepos,iout[2]); SET(traceback_score,iout[3]);
  *nmatches = iout[4]; *nmismatches = iout[5]; *nopens = iout[6]; *nindels = iout[7];
  SET(exonhead,iout[8]); *introntype = iout[9];
  *left_prob = dout[0]; *right_prob = dout[1];
  if (fetch) { free(rp); free(lp); free(gRa); free(gR); free(gLa); free(gL); }
  return (n < 0) ? (List_T) NULL : pairs_to_list(pairpool,pairs,gaps,base,dpi,n);
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
  const gmapdp_cpair *pairs;
  const gmapdp_pairbase *base;
  const gmapdp_gapinfo *gaps;
  int dpi;
  gmapdp_batch *b = my_batch(dynprogL);
  int id, n, rev_goffset = goffset + glength - 1;
  bool fetch = (glength > 1 && glength <= dynprogR->max_glength && rlengthR <= dynprogR->max_rlength &&
		glength <= dynprogL->max_glength && rlengthL <= dynprogL->max_rlength);

  PROF_BEGIN();
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
  if (fetch && sm100_resident(cur_thread(),genome,genomealt)) {
    gmapdp_coords co;
    memset(&co,0,sizeof(co));
    co.chroffset = (uint32_t) chroffset; co.chrhigh = (uint32_t) chrhigh;
    sm100_seg_forward(watsonp,chroffset,chrhigh,goffset,&co.gposL,&co.negL,&co.leftL);
    GmapDP_batch_next_coords(b,&co);		/* taken only if both fetches gave the same characters (no chromosome edge) */
  }
  PROF_MARK(0);
  id = GmapDP_cdna_gap(b,*dynprogindex,rsequenceL,rsequence_ucL,rev_rsequenceR,rev_rsequence_ucR,rlengthL,rlengthR,glength,
		       roffsetL,rev_roffsetR,goffset,g,ga,rg,rga,jump_late_p,extraband_paired,defect_rate);
  PROF_MARK(1);
  run_call(b);
  n = GmapDP_result_view(b,id,&iout,NULL,&pairs,&gaps,&base,&dpi);
  *dynprogindex = iout[0];
  SET(traceback_score,iout[1]);
  if (iout[2]) *incompletep = true;
  if (fetch) { free(rga); free(rg); free(ga); free(g); }
  return (n < 0) ? (List_T) NULL : pairs_to_list(pairpool,pairs,gaps,base,dpi,n);
}
