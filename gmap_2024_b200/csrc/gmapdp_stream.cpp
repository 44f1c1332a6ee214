/* gmapdp_stream.cpp -- the batching runtime (include/gmapdp_stream.h): host threads submit single DP boxes,
 * flights carry them to the device.  Plain C++ and pthreads; everything CUDA is behind the GdpFlight functions
 * of gmapdp_kernels.cu (gmapdp_internal.h).  No DP cell is computed here. */
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

void futex_wait (std::atomic<int> *w, int expected) {
  syscall(SYS_futex,reinterpret_cast<int *>(w),FUTEX_WAIT_PRIVATE,expected,NULL,NULL,0);
}

struct Lane;

struct Flight {
  GdpFlight *dev = NULL;
  Lane *lane = NULL;
  /* accumulation, under the lane mutex */
  int n = 0; bool full = false;
  size_t pool_used = 0, script_need = 0, ws_words = 0;
  int maxcols = 8;
  std::vector<int> bucket;		/* per box: launch-order key */
  std::vector<double> t_submit;
  std::atomic<int> writers{0};		/* submitters still copying into the staging */
  /* completion */
  std::atomic<int> done{0};		/* 0 in progress, 1 results ready, 2 failed */
  std::atomic<int> *woken = NULL;	/* per box: futex word its owner sleeps on (0 asleep / not yet told, 1 told) */
  std::atomic<int> readers{0};		/* boxes not yet released by their owners */
  int rc = 0;
  const gmapdp_result *results = NULL; const uint32_t *script = NULL;
  double t_first = 0.0, t_close = 0.0, t_launch = 0.0, t_launched = 0.0, t_started = 0.0;
  void reset () {
    n = 0; full = false; pool_used = script_need = ws_words = 0; maxcols = 8;
    done.store(0,std::memory_order_relaxed); rc = 0; results = NULL; script = NULL;
  }
};

struct Lane {
  gmapdp_stream *owner = NULL;
  gmapdp_ctx *ctx = NULL;
  int device = 0, depth = 2;
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_open;
  Flight *open = NULL;
  std::deque<Flight *> free_list;
  std::vector<Flight *> all;
  int inflight = 0;
  bool stop = false, failed = false;
  pthread_t service;
  bool threads_started = false;
  long spin_us = 400;
  /* statistics, under mu */
  double boxes = 0, flights = 0, largest = 0, t_flight = 0, t_wait = 0, h2d = 0, d2h = 0, gpu_s = 0, copy_s = 0;
  double t_stage[5] = {0,0,0,0,0};	/* close -> launch call, launch call, launch returned -> completion seen, root wake-ups */
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
};

namespace {

thread_local std::string tls_err;	/* errors of the per-call functions are per thread (hundreds of threads share one stream) */
const int WAKE_ROOTS = 2;
void wake_box (Flight *F, int i) {
  F->woken[i].store(1,std::memory_order_release);
  syscall(SYS_futex,reinterpret_cast<int *>(&F->woken[i]),FUTEX_WAKE_PRIVATE,1,NULL,NULL,0);
}

/* kind by kind, each kind by decreasing work: counting sort over the boxes' bucket keys (stable) */
void build_order (Flight *F, std::vector<int> &count) {
  const int nb = gdp_bucket_count();
  count.assign((size_t) nb + 1,0);
  for (int i = 0; i < F->n; i++) count[(size_t) F->bucket[i] + 1]++;
  for (int k = 1; k <= nb; k++) count[k] += count[k-1];
  int *order = gdp_flight_order(F->dev,F->n);
  for (int i = 0; i < F->n; i++) order[count[F->bucket[i]]++] = i;
}

/* The launcher and the completer of a lane are latency-critical and almost always asleep, while hundreds of worker
   threads compete for the cores.  Linux's EEVDF scheduler (6.6+) lets a thread ask for a short time slice: it is then
   picked -- and, from 6.12, preempts -- ahead of threads with the default 3 ms slice.  Unprivileged; ignored by older
   kernels.  GMAPDP_STREAM_RT=1 asks for SCHED_FIFO instead (needs CAP_SYS_NICE). */
struct sched_attr_v1 {
  uint32_t size, sched_policy; uint64_t sched_flags; int32_t sched_nice; uint32_t sched_priority;
  uint64_t sched_runtime, sched_deadline, sched_period; uint32_t sched_util_min, sched_util_max;
};
std::atomic<int> sched_mode{0};		/* what the service threads got: 1 SCHED_FIFO, 2 short slices, 3 nothing */
void service_thread_priority () {
  static const bool rt = getenv("GMAPDP_STREAM_RT") != NULL;
  sched_attr_v1 a;
  memset(&a,0,sizeof(a));
  a.size = sizeof(a);
  if (rt) {
    a.sched_policy = 1 /* SCHED_FIFO */; a.sched_priority = 10;
    if (syscall(SYS_sched_setattr,0,&a,0) == 0) { sched_mode.store(1); return; }
    memset(&a,0,sizeof(a)); a.size = sizeof(a);
  }
  a.sched_policy = 0 /* SCHED_OTHER */; a.sched_runtime = 100 * 1000;	/* 100 us slices */
  sched_mode.store(syscall(SYS_sched_setattr,0,&a,0) == 0 ? 2 : 3);
}

/* One service thread per lane: launcher and completer in one loop, so that no hand-over between threads sits in a
   flight's critical path.  While a flight is out the thread polls its event (and keeps launching the next flight as
   soon as the rules below allow); only when the lane is idle does it sleep on the condition variable.

   When to close the open flight and launch it:
     - nothing is in flight: at once, with whatever it holds (latency);
     - otherwise (at most `depth' flights in flight): when it has grown to `fill_boxes' boxes, or its first box has waited
       `linger_us' -- the boxes that arrive while a flight is out would otherwise leave one by one.  */
void complete_flight (Lane *L, Flight *F, int rc) {
  if (rc == GMAPDP_OK) rc = gdp_flight_wait(F->dev);		/* the event has fired (or the poll gave up): returns at once, reads the timings */
  if (rc == GMAPDP_OK) rc = gdp_flight_results(F->dev,&F->results,&F->script);
  F->rc = rc;
  const double t1 = now_s();
  /* everything this thread still needs from the flight is read NOW: once its last owner has been told, the flight
     can be released, reopened and refilled before this thread runs again */
  const int n = F->n;
  double waited = 0.0;
  for (int i = 0; i < n; i++) waited += t1 - F->t_submit[i];
  const double st0 = F->t_launch - F->t_close, st1 = F->t_launched - F->t_launch, st2 = t1 - F->t_launched, t_launch = F->t_launch;
  const double gpu_s = 1e-3 * F->dev->gpu_ms, copy_s = 1e-3 * F->dev->copy_ms;
  const double st4 = (F->t_started > 0.0) ? F->t_started - F->t_launched : 0.0;
  if (rc != GMAPDP_OK) {
    pthread_mutex_lock(&L->mu);
    L->failed = true; L->err = gmapdp_last_error(L->ctx);
    pthread_mutex_unlock(&L->mu);
  }
  F->done.store(rc == GMAPDP_OK ? 1 : 2,std::memory_order_release);
  /* Wake-ups fan out as a tree: this thread tells the first WAKE_ROOTS boxes' owners, every owner tells two more before
     it goes on (box i -> boxes WAKE_ROOTS + 2i and WAKE_ROOTS + 2i + 1).  Waking a sleeping thread costs a few microseconds of
     kernel time in the waker; hundreds of them from this one thread would serialise the whole runtime. */
  if (n > 1) wake_box(F,1);		/* box 0's owner last: with n == 1 it is the one that can release the flight */
  wake_box(F,0);
  const double t2 = now_s();
  pthread_mutex_lock(&L->mu);
  L->inflight--;
  L->t_flight += t1 - t_launch; L->t_wait += waited;
  L->t_stage[0] += st0; L->t_stage[1] += st1; L->t_stage[2] += st2; L->t_stage[3] += t2 - t1; L->t_stage[4] += st4;
  L->gpu_s += gpu_s; L->copy_s += copy_s;
  pthread_mutex_unlock(&L->mu);
}

void *service_main (void *arg) {
  Lane *L = (Lane *) arg;
  service_thread_priority();
  std::vector<int> count;
  std::deque<Flight *> out;		/* launched, oldest first (this thread only) */
  const double linger = 1e-6 * (double) env_long("GMAPDP_STREAM_LINGER_US",40);
  const int fill_boxes = (int) env_long("GMAPDP_STREAM_FILL",64);
  pthread_mutex_lock(&L->mu);
  for (;;) {
    if (L->stop && out.empty()) break;
    Flight *F = L->open;
    bool go = false;
    if (F && F->n > 0 && (int) out.size() < L->depth && !L->stop) {
      go = out.empty() || F->full || F->n >= fill_boxes || now_s() - F->t_first >= linger;
    }
    if (go) {
      L->open = NULL;
      if (!L->free_list.empty()) { L->open = L->free_list.front(); L->free_list.pop_front(); pthread_cond_broadcast(&L->cv_open); }
      L->inflight++;
      L->boxes += F->n; L->flights += 1; if (F->n > L->largest) L->largest = F->n;
      L->h2d += 32.0 + (double) F->n * (sizeof(gmapdp_box) + sizeof(int)) + (double) F->pool_used;
      L->d2h += 64.0 * F->n + 4.0 * (double) F->script_need;
      pthread_mutex_unlock(&L->mu);
      F->t_close = now_s();
      while (F->writers.load(std::memory_order_acquire) != 0) sched_yield();	/* submitters finishing their copies */
      build_order(F,count);
      F->t_launch = now_s();
      F->readers.store(F->n,std::memory_order_relaxed);
      F->rc = gdp_flight_launch(F->dev,F->n,F->pool_used,F->script_need,F->ws_words,F->maxcols);
      F->t_launched = now_s(); F->t_started = 0.0;
      out.push_back(F);
      pthread_mutex_lock(&L->mu);
      continue;
    }
    if (!out.empty()) {
      /* something is out: poll it; the lock is dropped, submitters are never held up by this loop */
      pthread_mutex_unlock(&L->mu);
      Flight *O = out.front();
      if (O->t_started == 0.0 && O->rc == GMAPDP_OK && gdp_flight_started(O->dev)) O->t_started = now_s();
      int p = (O->rc == GMAPDP_OK) ? gdp_flight_poll(O->dev) : 1;
      if (p != 0) { out.pop_front(); complete_flight(L,O,p < 0 ? p : O->rc); }
      else if (now_s() - O->t_launched > 1e-6 * (double) L->spin_us) {
	/* a long flight (a large box): stop burning the core, look again every 50 us (GMAPDP_STREAM_SPIN_US, default 400) */
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
  for (Flight *F : L->all) { if (F->dev) gdp_flight_destroy(F->dev); delete[] F->woken; delete F; }
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
  /* capacities of a flight: a box needs at most ~10 KB of sequence and ~5.4 K script words (660 x 2000 per side) */
  s->pool_cap = (size_t) env_long("GMAPDP_STREAM_POOL_MB",4) << 20;		/* sequences + MaxEnt probabilities */
  s->script_cap = (size_t) env_long("GMAPDP_STREAM_SCRIPT_MB",2) << 18;	/* words */
  const int nflights = (int) env_long("GMAPDP_STREAM_FLIGHTS",12), depth = (int) env_long("GMAPDP_STREAM_DEPTH",2);
  /* Size classes.  A flight is back when its SLOWEST box is: one warp per box, and a 660 x 990 end gap (eleven 32-diagonal
     passes of ~1000 steps) keeps its warp busy a hundred times longer than the 100 x 110 boxes that make up 95 % of the
     calls.  So every device has one lane per class -- own context, stream, workspaces, flights -- and their kernels
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
    for (int k = 0; k < (nflights < 3 ? 3 : nflights); k++) {
      Flight *F = new Flight();
      L->all.push_back(F);
      F->lane = L;
      rc = gdp_flight_create(L->ctx,&F->dev,s->max_boxes,s->pool_cap,s->script_cap);
      if (rc != GMAPDP_OK) { s->err = gmapdp_last_error(L->ctx); return rc; }
      F->bucket.resize(s->max_boxes); F->t_submit.resize(s->max_boxes);
      F->woken = new std::atomic<int>[s->max_boxes];
      for (int i = 0; i < s->max_boxes; i++) F->woken[i].store(0,std::memory_order_relaxed);
      F->reset();
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
  if (s->have_key) pthread_key_delete(s->lane_key);
  delete s;
}

extern "C" const char *gmapdp_stream_error (const gmapdp_stream *s) {
  if (!tls_err.empty()) return tls_err.c_str();
  return s ? s->err.c_str() : "null stream";
}
extern "C" int gmapdp_stream_ndevices (const gmapdp_stream *s) { return s->ndevices; }

extern "C" int gmapdp_stream_submit (gmapdp_stream *s, const gmapdp_box *box, const uint8_t *seq, size_t seqbytes,
				     const double *probs, size_t nprobs, gmapdp_ticket *ticket) {
  GdpBoxGeom g;
  if (gdp_box_geometry(box,&g) != GMAPDP_OK) { tls_err = "bad box"; return GMAPDP_ERR_ARG; }
  /* the box's share of the flight's pool: its sequences, then (8-aligned) its probabilities */
  const size_t sb = align16(seqbytes), need = sb + nprobs * sizeof(double);
  if (need > s->pool_cap || g.script_words > s->script_cap) { tls_err = "box larger than a flight"; return GMAPDP_ERR_CAPACITY; }
  /* a thread stays on one device for its lifetime; the lane there follows the box's size class */
  int di = (int) (intptr_t) pthread_getspecific(s->lane_key) - 1;
  if (di < 0) {
    di = (int) (s->next_lane.fetch_add(1) % (unsigned) s->ndevices);
    pthread_setspecific(s->lane_key,(void *) (intptr_t) (di + 1));
  }
  int cls = 0;
  while (g.steps >= s->class_limit[cls]) cls++;
  const int li = di * s->nclasses + cls;
  Lane *L = s->lanes[li];

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
  F->woken[idx].store(0,std::memory_order_relaxed);
  F->writers.fetch_add(1,std::memory_order_relaxed);
  const double t_now = now_s();
  if (idx == 0) { F->t_first = t_now; pthread_cond_signal(&L->cv_work); }
  pthread_mutex_unlock(&L->mu);

  F->t_submit[idx] = t_now;
  gmapdp_box x = *box;
  x.qL_off += (uint32_t) sbase; x.qR_off += (uint32_t) sbase; x.gL_off += (uint32_t) sbase; x.gLalt_off += (uint32_t) sbase;
  x.gR_off += (uint32_t) sbase; x.gRalt_off += (uint32_t) sbase;
  x.probL_off += (uint32_t) pbase; x.probR_off += (uint32_t) pbase;
  gdp_flight_boxes(F->dev)[idx] = x;
  if (seqbytes) memcpy(F->dev->h_pool + sbase,seq,seqbytes);
  if (nprobs) memcpy(F->dev->h_pool + sbase + sb,probs,nprobs * sizeof(double));
  F->writers.fetch_sub(1,std::memory_order_release);
  ticket->flight = F; ticket->index = idx; ticket->lane = li;
  return GMAPDP_OK;
}

extern "C" int gmapdp_stream_wait (gmapdp_stream *s, const gmapdp_ticket *ticket, const gmapdp_result **result, const uint32_t **ops) {
  Flight *F = (Flight *) ticket->flight;
  const int i = ticket->index;
  while (F->woken[i].load(std::memory_order_acquire) == 0) futex_wait(&F->woken[i],0);
  /* pass the news on before doing anything else (see completer_main) */
  for (int c = WAKE_ROOTS + 2 * i; c <= WAKE_ROOTS + 2 * i + 1 && c < F->n; c++) wake_box(F,c);
  const int d = F->done.load(std::memory_order_acquire);
  if (d == 0) { tls_err = "internal error: a box's owner was woken before its flight completed"; return GMAPDP_ERR_ARG; }
  if (d != 1) { tls_err = "flight failed: " + F->lane->err; return F->rc ? F->rc : GMAPDP_ERR_CUDA; }
  const gmapdp_result *r = F->results + ticket->index;
  if (r->script_off < 0 || (size_t) r->script_off + (size_t) r->script_lenA + (size_t) r->script_lenB > F->script_need) {
    tls_err = "device script pool overflow"; return GMAPDP_ERR_CAPACITY;
  }
  *result = r; *ops = F->script + r->script_off;
  return GMAPDP_OK;
}

extern "C" void gmapdp_stream_release (gmapdp_stream *s, const gmapdp_ticket *ticket) {
  (void) s;
  Flight *F = (Flight *) ticket->flight;
  if (F->readers.fetch_sub(1,std::memory_order_acq_rel) == 1) {
    Lane *L = F->lane;
    pthread_mutex_lock(&L->mu);
    F->reset();
    if (L->open == NULL) { L->open = F; pthread_cond_broadcast(&L->cv_open); }
    else L->free_list.push_back(F);
    pthread_mutex_unlock(&L->mu);
  }
}

/* per lane: boxes, flights, largest flight, mean flight latency (us), mean device time (us, GMAPDP_STREAM_TIMING), mean upload time (us) */
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
    out[4] += L->t_flight; out[5] += L->t_wait; out[6] += (double) gmapdp_launch_count(L->ctx);
    out[7] += L->h2d; out[8] += L->d2h;
    pthread_mutex_unlock(&L->mu);
  }
}
