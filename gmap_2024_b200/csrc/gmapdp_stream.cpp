/* gmapdp_stream.cpp -- the batching runtime (include/gmapdp_stream.h): host threads submit single DP boxes,
 * flights carry them to the device.  Plain C++ and pthreads; everything CUDA is behind the GdpFlight functions
 * of gmapdp_kernels.cu (gmapdp_internal.h).  No DP cell is computed here.
 *
 * Life of a flight, all of it inside its lane's service thread:
 *   open      submitters reserve a slot under the lane mutex and copy their box, sequences and probabilities into the
 *             pinned staging outside it
 *   out       closed, ordered, launched (two H2D copies, one kernel, one D2H copy, one event) on the lane's stream
 *   landed    the event has fired: every box's result and edit script is copied into its owner's MAILBOX, the owners
 *             are woken (as a tree), and the flight goes back to the free list at once
 * A flight never waits for a worker thread: with hundreds of workers on a few cores the slowest owner of a flight can
 * be a scheduling quantum late, and a flight held until then was what starved the submitters of open flights.
 */
#include "../../include/gmapdp_stream.h"
#include "gmapdp_internal.h"

#include <pthread.h>
#include <unistd.h>
#include <limits.h>
#include <time.h>
#include <sys/syscall.h>
#include <linux/futex.h>

#include <atomic>
#include <deque>
#include <string>
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <cstdint>

namespace {

double now_s () {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

inline size_t align16 (size_t x) { return (x + 15) & ~(size_t) 15; }
long env_long (const char *name, long dflt) { const char *e = getenv(name); return (e && *e) ? atol(e) : dflt; }

inline std::atomic<int> *state_of (gmapdp_mailbox *mb) { return reinterpret_cast<std::atomic<int> *>(&mb->state); }

void futex_wait (std::atomic<int> *w, int expected) {
  syscall(SYS_futex,reinterpret_cast<int *>(w),FUTEX_WAIT_PRIVATE,expected,NULL,NULL,0);
}

/* tells a mailbox's owner that its box is back (state 1) or lost (state 2) */
void wake_mailbox (gmapdp_mailbox *mb, int state) {
  state_of(mb)->store(state,std::memory_order_release);
  syscall(SYS_futex,&mb->state,FUTEX_WAKE_PRIVATE,1,NULL,NULL,0);
}

struct Lane;

struct Flight {
  GdpFlight *dev = NULL;
  Lane *lane = NULL;
  /* accumulation, under the lane mutex */
  int n = 0; bool full = false;
  size_t pool_used = 0, script_need = 0, ws_words = 0;
  int maxcols = 8;
  std::vector<int> bucket;			/* per box: launch-order key */
  std::vector<gmapdp_mailbox *> owner;		/* per box: where its result goes */
  std::atomic<int> writers{0};			/* submitters still copying into the staging */
  int rc = 0;
  double t_first = 0.0, t_close = 0.0, t_launch = 0.0, t_launched = 0.0, t_started = 0.0;
  void reset () { n = 0; full = false; pool_used = script_need = ws_words = 0; maxcols = 8; rc = 0; }
};

struct Lane {
  gmapdp_stream *owner = NULL;
  gmapdp_ctx *ctx = NULL;
  int device = 0, depth = 2, max_flights = 6;
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_open;
  Flight *open = NULL;
  std::deque<Flight *> free_list;
  std::vector<Flight *> all;
  bool stop = false, failed = false;
  pthread_t service;
  bool threads_started = false;
  long spin_us = 400;
  double ema_flight = 150e-6;			/* running estimate of launch -> landed */
  /* statistics, under mu */
  double boxes = 0, flights = 0, largest = 0, t_flight = 0, t_wait = 0, h2d = 0, d2h = 0, gpu_s = 0, copy_s = 0;
  double t_stage[5] = {0,0,0,0,0};	/* close -> launch call, launch call, launch returned -> landed, distribution + root wake-ups, launch returned -> begun */
  long launches = 0;
  std::string err;
};

}  // namespace

struct gmapdp_stream {
  std::vector<Lane *> lanes;		/* [device][class] */
  int ndevices = 0, nclasses = 1;
  int class_limit[8];			/* class c takes boxes of fewer than class_limit[c] steps (the last class: everything) */
  std::atomic<unsigned> next_lane{0};
  int max_boxes = 0;
  size_t pool_cap = 0, script_cap = 0;
  std::string err;
  pthread_key_t lane_key; bool have_key = false;
  std::vector<gmapdp_genome *> genomes;	/* one per device (gmapdp_stream_genome) */
  std::atomic<int> genome_ready{0};
};

namespace {

thread_local std::string tls_err;	/* errors of the per-call functions are per thread (hundreds of threads share one stream) */
const int WAKE_ROOTS = 2;

/* longest boxes first: counting sort over the boxes' bucket keys (stable) */
void build_order (Flight *F, std::vector<int> &count) {
  const int nb = gdp_bucket_count();
  count.assign((size_t) nb + 1,0);
  for (int i = 0; i < F->n; i++) count[(size_t) F->bucket[i] + 1]++;
  for (int k = 1; k <= nb; k++) count[k] += count[k-1];
  int *order = gdp_flight_order(F->dev,F->n);
  for (int i = 0; i < F->n; i++) order[count[F->bucket[i]]++] = i;
}

/* The service thread of a lane is latency-critical, while hundreds of worker threads compete for the cores.  Linux's
   EEVDF scheduler (6.6+) lets a thread ask for a short time slice: it is then picked -- and, from 6.12, preempts --
   ahead of threads with the default 3 ms slice.  Unprivileged; ignored by older kernels. */
struct sched_attr_v1 {
  uint32_t size, sched_policy; uint64_t sched_flags; int32_t sched_nice; uint32_t sched_priority;
  uint64_t sched_runtime, sched_deadline, sched_period; uint32_t sched_util_min, sched_util_max;
};
std::atomic<int> sched_mode{0};		/* what the service threads got: 2 short slices, 3 nothing */
void service_thread_priority () {
  sched_attr_v1 a;
  memset(&a,0,sizeof(a));
  a.size = sizeof(a);
  a.sched_policy = 0 /* SCHED_OTHER */; a.sched_runtime = 100 * 1000;	/* 100 us slices */
  sched_mode.store(syscall(SYS_sched_setattr,0,&a,0) == 0 ? 2 : 3);
}

Flight *new_flight (Lane *L) {		/* lane mutex held, or the lane not yet shared */
  gmapdp_stream *s = L->owner;
  Flight *F = new Flight();
  F->lane = L;
  if (gdp_flight_create(L->ctx,&F->dev,s->max_boxes,s->pool_cap,s->script_cap) != GMAPDP_OK) {
    if (F->dev) gdp_flight_destroy(F->dev);
    delete F;
    return NULL;
  }
  F->bucket.resize(s->max_boxes); F->owner.resize(s->max_boxes);
  F->reset();
  L->all.push_back(F);
  return F;
}

/* the flight's event has fired (or its launch failed): results into the mailboxes, owners woken, flight recycled */
void land_flight (Lane *L, Flight *F, int rc) {
  const gmapdp_result *results = NULL; const uint32_t *script = NULL;
  if (rc == GMAPDP_OK) rc = gdp_flight_wait(F->dev);		/* returns at once; reads the device timings */
  if (rc == GMAPDP_OK) rc = gdp_flight_results(F->dev,&results,&script);
  const double t1 = now_s();
  const int n = F->n;
  double waited = 0.0;
  if (rc != GMAPDP_OK) {
    pthread_mutex_lock(&L->mu);
    L->failed = true; L->err = gmapdp_last_error(L->ctx);
    pthread_mutex_unlock(&L->mu);
  }
  /* Wake-ups fan out as a tree: this thread wakes the owners of boxes 0 and 1, the owner of box i wakes those of boxes
     2 + 2i and 3 + 2i before it goes on.  Waking a sleeping thread costs a few microseconds of kernel time in the waker:
     hundreds of them from this one thread would serialise the whole runtime.  The tree hangs off the mailboxes, not
     off the flight, which is free again as soon as this function returns. */
  for (int i = 0; i < n; i++) {
    gmapdp_mailbox *mb = F->owner[i];
    waited += t1 - mb->t_submit;
    mb->rc = rc;
    if (rc == GMAPDP_OK) {
      const gmapdp_result &r = results[i];
      const size_t len = (size_t) r.script_lenA + (size_t) r.script_lenB;
      if (r.script_off < 0 || (size_t) r.script_off + len > F->script_need || len > mb->ops_cap) mb->rc = GMAPDP_ERR_CAPACITY;
      else { mb->result = r; mb->result.script_off = 0; memcpy(mb->ops,script + r.script_off,len * sizeof(uint32_t)); }
    }
    const int c = WAKE_ROOTS + 2 * i;
    mb->wake[0] = (c < n) ? F->owner[c] : NULL;
    mb->wake[1] = (c + 1 < n) ? F->owner[c + 1] : NULL;
  }
  for (int i = 0; i < WAKE_ROOTS && i < n; i++) wake_mailbox(F->owner[i],rc == GMAPDP_OK ? 1 : 2);
  const double t2 = now_s();
  const double st4 = (F->t_started > 0.0) ? F->t_started - F->t_launched : 0.0;
  pthread_mutex_lock(&L->mu);
  L->t_flight += t1 - F->t_launch; L->t_wait += waited;
  L->t_stage[0] += F->t_launch - F->t_close; L->t_stage[1] += F->t_launched - F->t_launch; L->t_stage[2] += t1 - F->t_launched;
  L->t_stage[3] += t2 - t1; L->t_stage[4] += st4;
  L->gpu_s += 1e-3 * F->dev->gpu_ms; L->copy_s += 1e-3 * F->dev->copy_ms;
  L->ema_flight = 0.9 * L->ema_flight + 0.1 * (t1 - F->t_launched);
  F->reset();
  if (L->open == NULL) { L->open = F; pthread_cond_broadcast(&L->cv_open); }
  else L->free_list.push_back(F);
  pthread_mutex_unlock(&L->mu);
}

/* One service thread per lane: launcher and completer in one loop, so that no hand-over between threads sits in a
   flight's critical path.  While a flight is out the thread polls its event (and launches the next flight when the
   rules below say so); only when the lane is idle does it sleep on the condition variable.

   The flights of a lane run one after the other on the lane's stream.  When to close the open flight and launch it:
     - nothing is out: at once, with whatever it holds (latency);
     - one flight is out: when the open one has grown to `fill_boxes' boxes, or when the flight that is out is
       expected back within a fraction of its usual duration -- the next flight is then queued behind it in time and
       carries everything that arrived meanwhile (throughput; the duration is a running average);
     - `depth' flights are out: never.  */
void *service_main (void *arg) {
  Lane *L = (Lane *) arg;
  service_thread_priority();
  std::vector<int> count;
  std::vector<Flight *> out;		/* launched and not yet landed, oldest first (this thread only) */
  const double linger_frac = 1e-2 * (double) env_long("GMAPDP_STREAM_LINGER_PCT",60);
  const int fill_boxes = (int) env_long("GMAPDP_STREAM_FILL",512);
  pthread_mutex_lock(&L->mu);
  for (;;) {
    if (L->stop && out.empty()) break;
    Flight *F = L->open;
    bool go = false;
    if (F && F->n > 0 && (int) out.size() < L->depth && !L->stop) {
      go = out.empty() || F->full || F->n >= fill_boxes || now_s() - out.back()->t_launched >= linger_frac * L->ema_flight;
    }
    if (go) {
      L->open = NULL;
      if (L->free_list.empty() && (int) L->all.size() < L->max_flights) {
	/* flights are created on demand (pinned staging is slow to allocate: a run that needs three never pays for more) */
	Flight *N = new_flight(L);
	if (N) L->free_list.push_back(N);
      }
      if (!L->free_list.empty()) { L->open = L->free_list.front(); L->free_list.pop_front(); pthread_cond_broadcast(&L->cv_open); }
      L->boxes += F->n; L->flights += 1; if (F->n > L->largest) L->largest = F->n;
      L->h2d += 32.0 + (double) F->n * (sizeof(gmapdp_box) + sizeof(int)) + (double) F->pool_used;
      L->d2h += 64.0 * F->n + 4.0 * (double) F->script_need;
      L->launches++;
      pthread_mutex_unlock(&L->mu);
      F->t_close = now_s();
      while (F->writers.load(std::memory_order_acquire) != 0) sched_yield();	/* submitters finishing their copies */
      build_order(F,count);
      F->t_launch = now_s();
      F->rc = gdp_flight_launch(F->dev,F->n,F->pool_used,F->script_need,F->ws_words,F->maxcols);
      F->t_launched = now_s(); F->t_started = 0.0;
      out.push_back(F);
      pthread_mutex_lock(&L->mu);
      continue;
    }
    if (!out.empty()) {
      /* something is out: poll the oldest (one stream: they land in order); the lock is dropped, submitters are never
	 held up by this loop */
      pthread_mutex_unlock(&L->mu);
      Flight *O = out.front();
      const double t_now = now_s();
      if (O->t_started == 0.0 && O->rc == GMAPDP_OK && gdp_flight_started(O->dev)) O->t_started = t_now;
      const int p = (O->rc == GMAPDP_OK) ? gdp_flight_poll(O->dev) : 1;
      if (p != 0) { out.erase(out.begin()); land_flight(L,O,p < 0 ? p : O->rc); }
      else if (t_now - O->t_launched > 1e-6 * (double) L->spin_us) {
	/* a long flight (a large box): stop burning the core, look again every 50 us (GMAPDP_STREAM_SPIN_US) */
	struct timespec ts = {0, 50000};
	nanosleep(&ts,NULL);
      }
      pthread_mutex_lock(&L->mu);
      continue;
    }
    if (L->stop) break;
    pthread_cond_wait(&L->cv_work,&L->mu);		/* idle lane */
  }
  pthread_mutex_unlock(&L->mu);
  return NULL;
}

void lane_destroy (Lane *L) {
  if (!L) return;
  if (L->threads_started) {
    pthread_mutex_lock(&L->mu);
    L->stop = true;
    pthread_cond_broadcast(&L->cv_work); pthread_cond_broadcast(&L->cv_open);
    pthread_mutex_unlock(&L->mu);
    pthread_join(L->service,NULL);
  }
  for (Flight *F : L->all) { if (F->dev) gdp_flight_destroy(F->dev); delete F; }
  if (L->ctx) gmapdp_destroy(L->ctx);
  pthread_mutex_destroy(&L->mu);
  pthread_cond_destroy(&L->cv_work); pthread_cond_destroy(&L->cv_open);
  delete L;
}

}  // namespace

extern "C" int gmapdp_stream_create (gmapdp_stream **out, const int *devices, int ndevices, int max_boxes) {
  gmapdp_stream *s = new gmapdp_stream();
  *out = s;
  if (ndevices <= 0 || !devices) { s->err = "gmapdp_stream_create: no device given"; return GMAPDP_ERR_ARG; }
  s->max_boxes = max_boxes > 0 ? max_boxes : (int) env_long("GMAPDP_STREAM_BOXES",2048);
  /* capacities of a flight: a box needs at most ~10 KB of sequence, ~40 KB of probabilities and ~5.4 K script words
     (660 x 2000 per side) */
  s->pool_cap = (size_t) env_long("GMAPDP_STREAM_POOL_MB",4) << 20;		/* sequences + MaxEnt probabilities */
  s->script_cap = (size_t) env_long("GMAPDP_STREAM_SCRIPT_MB",2) << 18;	/* words */
  const int nflights = (int) env_long("GMAPDP_STREAM_FLIGHTS",6), depth = (int) env_long("GMAPDP_STREAM_DEPTH",2);
  /* Size classes.  A flight is back when its SLOWEST box is: one warp per box, and a 660 x 990 end gap (eleven 32-diagonal
     passes of ~1000 steps) keeps its warp busy a hundred times longer than the 100 x 110 boxes that make up 95 % of the
     calls.  So every device has one lane per class -- own context, stream, workspace, flights -- and their kernels
     run side by side: small boxes never wait for a large one.  GMAPDP_STREAM_CLASSES="a,b" sets the step limits. */
  {
    const char *e = getenv("GMAPDP_STREAM_CLASSES");
    std::string spec = e ? e : "700,6000";
    s->nclasses = 1;
    size_t pos = 0;
    while (pos < spec.size() && s->nclasses < 8) {
      const long v = atol(spec.c_str() + pos);
      if (v > 0) s->class_limit[s->nclasses - 1] = (int) v, s->nclasses++;
      pos = spec.find(',',pos);
      if (pos == std::string::npos) break;
      pos++;
    }
    s->class_limit[s->nclasses - 1] = INT_MAX;
  }
  s->ndevices = ndevices;
  if (pthread_key_create(&s->lane_key,NULL) != 0) { s->err = "pthread_key_create failed"; return GMAPDP_ERR_ARG; }
  s->have_key = true;
  for (int dc = 0; dc < ndevices * s->nclasses; dc++) {
    const int d = dc / s->nclasses;
    Lane *L = new Lane();
    s->lanes.push_back(L);
    L->owner = s; L->device = devices[d]; L->depth = depth < 1 ? 1 : depth;
    pthread_mutex_init(&L->mu,NULL);
    pthread_cond_init(&L->cv_work,NULL); pthread_cond_init(&L->cv_open,NULL);
    int rc = gmapdp_create(&L->ctx,devices[d]);
    if (rc != GMAPDP_OK) { s->err = gmapdp_last_error(L->ctx); return rc; }
    L->max_flights = nflights < 3 ? 3 : nflights;
    for (int k = 0; k < 2; k++) {
      Flight *F = new_flight(L);
      if (!F) { s->err = gmapdp_last_error(L->ctx); return GMAPDP_ERR_CUDA; }
      if (k == 0) L->open = F; else L->free_list.push_back(F);
    }
    L->spin_us = env_long("GMAPDP_STREAM_SPIN_US",400);
    if (pthread_create(&L->service,NULL,service_main,L) != 0) {
      s->err = "pthread_create failed"; return GMAPDP_ERR_ARG;
    }
    L->threads_started = true;
  }
  return GMAPDP_OK;
}

extern "C" void gmapdp_stream_destroy (gmapdp_stream *s) {
  if (!s) return;
  for (Lane *L : s->lanes) lane_destroy(L);
  for (gmapdp_genome *g : s->genomes) gmapdp_genome_destroy(g);
  if (s->have_key) pthread_key_delete(s->lane_key);
  delete s;
}

extern "C" const char *gmapdp_stream_error (const gmapdp_stream *s) {
  if (!tls_err.empty()) return tls_err.c_str();
  return s ? s->err.c_str() : "null stream";
}
extern "C" int gmapdp_stream_ndevices (const gmapdp_stream *s) { return s->ndevices; }

/* Resident genome for every lane: one device copy per GPU (gmapdp_genome_create), attached to all of its lanes.  Call it
   once, before the first box with gflags is submitted. */
extern "C" int gmapdp_stream_genome (gmapdp_stream *s, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *maxent) {
  if (s->genome_ready.load()) { s->err = "gmapdp_stream_genome: a genome is already resident"; return GMAPDP_ERR_ARG; }
  for (int d = 0; d < s->ndevices; d++) {
    Lane *L0 = s->lanes[(size_t) d * s->nclasses];
    gmapdp_genome *g = NULL;
    int rc = gmapdp_genome_create(&g,L0->device,blocks,nwords,maxent);
    if (rc != GMAPDP_OK) { s->err = "gmapdp_stream_genome: upload failed"; return rc; }
    s->genomes.push_back(g);
    for (int c = 0; c < s->nclasses; c++) {
      Lane *L = s->lanes[(size_t) d * s->nclasses + c];
      pthread_mutex_lock(&L->mu);
      rc = gmapdp_genome_attach(L->ctx,g);
      pthread_mutex_unlock(&L->mu);
      if (rc != GMAPDP_OK) { s->err = gmapdp_last_error(L->ctx); return rc; }
    }
  }
  s->genome_ready.store(1);
  return GMAPDP_OK;
}

extern "C" int gmapdp_stream_submit (gmapdp_stream *s, const gmapdp_box *box, const uint8_t *seq, size_t seqbytes,
				     const double *probs, size_t nprobs, gmapdp_mailbox *mb) {
  GdpBoxGeom g;
  if (gdp_box_geometry(box,&g) != GMAPDP_OK) { tls_err = "bad box"; return GMAPDP_ERR_ARG; }
  /* the box's share of the flight's pool: its sequences, then (8-aligned) its probabilities */
  const size_t sb = align16(seqbytes), need = sb + nprobs * sizeof(double);
  if (need > s->pool_cap || g.script_words > s->script_cap) { tls_err = "box larger than a flight"; return GMAPDP_ERR_CAPACITY; }
  if (!mb || !mb->ops || mb->ops_cap < g.script_words) { tls_err = "mailbox too small for the box's edit script"; return GMAPDP_ERR_ARG; }
  if (box->gflags && !s->genome_ready.load(std::memory_order_acquire)) { tls_err = "box refers to a resident genome, none was given (gmapdp_stream_genome)"; return GMAPDP_ERR_ARG; }
  /* a thread stays on one device for its lifetime; the lane there follows the box's size class */
  int di = (int) (intptr_t) pthread_getspecific(s->lane_key) - 1;
  if (di < 0) {
    di = (int) (s->next_lane.fetch_add(1) % (unsigned) s->ndevices);
    pthread_setspecific(s->lane_key,(void *) (intptr_t) (di + 1));
  }
  int cls = 0;
  while (g.steps >= s->class_limit[cls]) cls++;
  Lane *L = s->lanes[(size_t) di * s->nclasses + cls];
  state_of(mb)->store(0,std::memory_order_relaxed);
  mb->rc = 0; mb->wake[0] = mb->wake[1] = NULL;

  pthread_mutex_lock(&L->mu);
  Flight *F;
  for (;;) {
    if (L->failed || L->stop) { tls_err = L->failed ? L->err : "stream is shutting down"; pthread_mutex_unlock(&L->mu); return GMAPDP_ERR_CUDA; }
    F = L->open;
    if (F && !F->full && F->n < s->max_boxes && F->pool_used + need <= s->pool_cap && F->script_need + g.script_words <= s->script_cap) break;
    if (F && F->n > 0) { F->full = true; pthread_cond_signal(&L->cv_work); }
    pthread_cond_wait(&L->cv_open,&L->mu);
  }
  const int idx = F->n++;
  const size_t sbase = F->pool_used, pbase = (sbase + sb) / sizeof(double);	/* bytes; doubles from the start of the pool */
  F->pool_used += need; F->script_need += g.script_words;
  if (g.ws_words > F->ws_words) F->ws_words = g.ws_words;
  if (g.cols > F->maxcols) F->maxcols = g.cols;
  F->bucket[idx] = g.bucket;
  F->owner[idx] = mb;
  F->writers.fetch_add(1,std::memory_order_relaxed);
  const double t_now = now_s();
  if (idx == 0) { F->t_first = t_now; pthread_cond_signal(&L->cv_work); }
  pthread_mutex_unlock(&L->mu);

  mb->t_submit = t_now;
  gmapdp_box x = *box;
  x.qL_off += (uint32_t) sbase; x.qR_off += (uint32_t) sbase;
  /* segments / probabilities that come from the resident genome are coordinates, not offsets into the flight's pool */
  if (!(x.gflags & GMAPDP_G_SEG_L)) { x.gL_off += (uint32_t) sbase; x.gLalt_off += (uint32_t) sbase; }
  if (!(x.gflags & GMAPDP_G_SEG_R)) { x.gR_off += (uint32_t) sbase; x.gRalt_off += (uint32_t) sbase; }
  if (!(x.gflags & GMAPDP_G_PROBS)) { x.probL_off += (uint32_t) pbase; x.probR_off += (uint32_t) pbase; }
  gdp_flight_boxes(F->dev)[idx] = x;
  if (seqbytes) memcpy(F->dev->h_pool + sbase,seq,seqbytes);
  if (nprobs) memcpy(F->dev->h_pool + sbase + sb,probs,nprobs * sizeof(double));
  F->writers.fetch_sub(1,std::memory_order_release);
  return GMAPDP_OK;
}

extern "C" int gmapdp_stream_wait (gmapdp_stream *s, gmapdp_mailbox *mb) {
  (void) s;
  int st;
  while ((st = state_of(mb)->load(std::memory_order_acquire)) == 0) futex_wait(state_of(mb),0);
  /* pass the news on before doing anything else (see land_flight) */
  for (int k = 0; k < 2; k++) if (mb->wake[k]) { gmapdp_mailbox *c = mb->wake[k]; mb->wake[k] = NULL; wake_mailbox(c,st); }
  if (st != 1 || mb->rc != GMAPDP_OK) {
    tls_err = (mb->rc == GMAPDP_ERR_CAPACITY) ? "device script pool overflow" : "flight failed (CUDA error on its lane)";
    return mb->rc ? mb->rc : GMAPDP_ERR_CUDA;
  }
  return GMAPDP_OK;
}

/* per lane: boxes, flights, largest flight, mean flight latency (us), mean device time (us, GMAPDP_STREAM_TIMING), mean upload time (us), stages */
extern "C" int gmapdp_stream_lane_stats (const gmapdp_stream *s, int lane, double *out) {
  if (lane < 0 || lane >= (int) s->lanes.size()) return GMAPDP_ERR_ARG;
  Lane *L = s->lanes[lane];
  pthread_mutex_lock(&L->mu);
  out[0] = L->boxes; out[1] = L->flights; out[2] = L->largest;
  out[3] = L->flights > 0 ? 1e6 * L->t_flight / L->flights : 0.0;
  out[4] = L->flights > 0 ? 1e6 * L->gpu_s / L->flights : 0.0;
  out[5] = L->flights > 0 ? 1e6 * L->copy_s / L->flights : 0.0;
  for (int k = 0; k < 5; k++) out[6 + k] = L->flights > 0 ? 1e6 * L->t_stage[k] / L->flights : 0.0;
  pthread_mutex_unlock(&L->mu);
  return GMAPDP_OK;
}
extern "C" int gmapdp_stream_nlanes (const gmapdp_stream *s) { return (int) s->lanes.size(); }
extern "C" int gmapdp_stream_sched_mode (const gmapdp_stream *s) { (void) s; return sched_mode.load(); }

extern "C" void gmapdp_stream_stats (const gmapdp_stream *s, double *out) {
  for (int k = 0; k < GMAPDP_STREAM_NSTATS; k++) out[k] = 0.0;
  for (Lane *L : s->lanes) {
    pthread_mutex_lock(&L->mu);
    out[0] += L->boxes; out[1] += L->flights; if (L->largest > out[2]) out[2] = L->largest;
    out[3] += L->gpu_s;
    out[4] += L->t_flight; out[5] += L->t_wait; out[6] += (double) L->launches;
    out[7] += L->h2d; out[8] += L->d2h;
    pthread_mutex_unlock(&L->mu);
  }
}
