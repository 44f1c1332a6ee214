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

namespace {

double now_s () {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

inline size_t align16 (size_t x) { return (x + 15) & ~(size_t) 15; }

void futex_wait (std::atomic<int> *w, int expected) {
  syscall(SYS_futex,reinterpret_cast<int *>(w),FUTEX_WAIT_PRIVATE,expected,NULL,NULL,0);
}
void futex_wake_all (std::atomic<int> *w) {
  syscall(SYS_futex,reinterpret_cast<int *>(w),FUTEX_WAKE_PRIVATE,INT_MAX,NULL,NULL,0);
}

struct Lane;

struct Flight {
  GdpFlight *dev = NULL;
  Lane *lane = NULL;
  /* accumulation, under the lane mutex */
  int n = 0; bool full = false;
  size_t seq_used = 0, prob_used = 0, script_need = 0;
  size_t ws_words[GDP_NKINDS]; int maxcols[GDP_NKINDS], cnt[GDP_NKINDS];
  std::vector<int> bucket;		/* per box: launch-order key */
  std::vector<double> t_submit;
  std::atomic<int> writers{0};		/* submitters still copying into the staging */
  /* completion */
  std::atomic<int> done{0};		/* futex word: 0 in progress, 1 results ready, 2 failed */
  std::atomic<int> readers{0};		/* boxes not yet released by their owners */
  int rc = 0;
  const gmapdp_result *results = NULL; const uint32_t *script = NULL;
  double t_launch = 0.0;
  void reset () {
    n = 0; full = false; seq_used = prob_used = script_need = 0;
    for (int k = 0; k < GDP_NKINDS; k++) { ws_words[k] = 0; maxcols[k] = 8; cnt[k] = 0; }
    done.store(0,std::memory_order_relaxed); rc = 0; results = NULL; script = NULL;
  }
};

struct Lane {
  gmapdp_stream *owner = NULL;
  gmapdp_ctx *ctx = NULL;
  int device = 0, depth = 2;
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_open, cv_launched;
  Flight *open = NULL;
  std::deque<Flight *> free_list, launched;
  std::vector<Flight *> all;
  int inflight = 0;
  bool stop = false, failed = false;
  pthread_t launcher, completer;
  bool threads_started = false;
  /* statistics, under mu */
  double boxes = 0, flights = 0, largest = 0, t_flight = 0, t_wait = 0, h2d = 0, d2h = 0;
  std::string err;
};

}  // namespace

struct gmapdp_stream {
  std::vector<Lane *> lanes;
  std::atomic<unsigned> next_lane{0};
  int max_boxes = 0;
  size_t seq_cap = 0, prob_cap = 0, script_cap = 0;
  std::string err;
  pthread_key_t lane_key; bool have_key = false;
};

namespace {

/* kind by kind, each kind by decreasing work: counting sort over the boxes' bucket keys (stable) */
void build_order (Flight *F, std::vector<int> &count) {
  const int nb = gdp_bucket_count();
  count.assign((size_t) nb + 1,0);
  for (int i = 0; i < F->n; i++) count[(size_t) F->bucket[i] + 1]++;
  for (int k = 1; k <= nb; k++) count[k] += count[k-1];
  int *order = F->dev->h_order;
  for (int i = 0; i < F->n; i++) order[count[F->bucket[i]]++] = i;
}

void *launcher_main (void *arg) {
  Lane *L = (Lane *) arg;
  std::vector<int> count;
  pthread_mutex_lock(&L->mu);
  for (;;) {
    while (!L->stop && !(L->open && L->open->n > 0 && L->inflight < L->depth)) pthread_cond_wait(&L->cv_work,&L->mu);
    if (L->stop) break;
    Flight *F = L->open;
    L->open = NULL;
    if (!L->free_list.empty()) { L->open = L->free_list.front(); L->free_list.pop_front(); pthread_cond_broadcast(&L->cv_open); }
    L->inflight++;
    L->boxes += F->n; L->flights += 1; if (F->n > L->largest) L->largest = F->n;
    pthread_mutex_unlock(&L->mu);

    while (F->writers.load(std::memory_order_acquire) != 0) sched_yield();	/* submitters finishing their copies */
    build_order(F,count);
    F->t_launch = now_s();
    F->readers.store(F->n,std::memory_order_relaxed);
    F->rc = gdp_flight_launch(F->dev,F->n,F->seq_used,F->prob_used,F->script_need,F->ws_words,F->maxcols,F->cnt);

    pthread_mutex_lock(&L->mu);
    L->h2d += (double) F->n * (sizeof(gmapdp_box) + sizeof(int)) + (double) F->seq_used + 8.0 * (double) F->prob_used;
    L->d2h += 16.0 + 64.0 * F->n + 4.0 * (double) F->script_need;
    L->launched.push_back(F);
    pthread_cond_signal(&L->cv_launched);
  }
  pthread_mutex_unlock(&L->mu);
  return NULL;
}

void *completer_main (void *arg) {
  Lane *L = (Lane *) arg;
  pthread_mutex_lock(&L->mu);
  for (;;) {
    while (!L->stop && L->launched.empty()) pthread_cond_wait(&L->cv_launched,&L->mu);
    if (L->launched.empty()) break;		/* stop, and nothing left to complete */
    Flight *F = L->launched.front();
    L->launched.pop_front();
    pthread_mutex_unlock(&L->mu);

    int rc = F->rc;
    if (rc == GMAPDP_OK) rc = gdp_flight_wait(F->dev);
    if (rc == GMAPDP_OK) rc = gdp_flight_results(F->dev,&F->results,&F->script);
    F->rc = rc;
    const double t1 = now_s();
    double waited = 0.0;
    for (int i = 0; i < F->n; i++) waited += t1 - F->t_submit[i];
    if (rc != GMAPDP_OK) {
      pthread_mutex_lock(&L->mu);
      L->failed = true; L->err = gmapdp_last_error(L->ctx);
      pthread_mutex_unlock(&L->mu);
    }
    F->done.store(rc == GMAPDP_OK ? 1 : 2,std::memory_order_release);
    futex_wake_all(&F->done);

    pthread_mutex_lock(&L->mu);
    L->inflight--;
    L->t_flight += t1 - F->t_launch; L->t_wait += waited;
    pthread_cond_signal(&L->cv_work);
  }
  pthread_mutex_unlock(&L->mu);
  return NULL;
}

void lane_destroy (Lane *L) {
  if (!L) return;
  if (L->threads_started) {
    pthread_mutex_lock(&L->mu);
    L->stop = true;
    pthread_cond_broadcast(&L->cv_work); pthread_cond_broadcast(&L->cv_launched); pthread_cond_broadcast(&L->cv_open);
    pthread_mutex_unlock(&L->mu);
    pthread_join(L->launcher,NULL); pthread_join(L->completer,NULL);
  }
  for (Flight *F : L->all) { if (F->dev) gdp_flight_destroy(F->dev); delete F; }
  if (L->ctx) gmapdp_destroy(L->ctx);
  pthread_mutex_destroy(&L->mu);
  pthread_cond_destroy(&L->cv_work); pthread_cond_destroy(&L->cv_open); pthread_cond_destroy(&L->cv_launched);
  delete L;
}

long env_long (const char *name, long dflt) { const char *e = getenv(name); return (e && *e) ? atol(e) : dflt; }

}  // namespace

extern "C" int gmapdp_stream_create (gmapdp_stream **out, const int *devices, int ndevices, int max_boxes) {
  gmapdp_stream *s = new gmapdp_stream();
  *out = s;
  if (ndevices <= 0 || !devices) { s->err = "gmapdp_stream_create: no device given"; return GMAPDP_ERR_ARG; }
  s->max_boxes = max_boxes > 0 ? max_boxes : (int) env_long("GMAPDP_STREAM_BOXES",8192);
  /* capacities of a flight: a box needs at most ~10 KB of sequence and ~5.4 K script words (660 x 2000 per side) */
  s->seq_cap = (size_t) env_long("GMAPDP_STREAM_SEQ_MB",8) << 20;
  s->prob_cap = (size_t) env_long("GMAPDP_STREAM_PROB_MB",8) << 17;		/* doubles */
  s->script_cap = (size_t) env_long("GMAPDP_STREAM_SCRIPT_MB",8) << 18;	/* words */
  const int nflights = (int) env_long("GMAPDP_STREAM_FLIGHTS",6), depth = (int) env_long("GMAPDP_STREAM_DEPTH",2);
  if (pthread_key_create(&s->lane_key,NULL) != 0) { s->err = "pthread_key_create failed"; return GMAPDP_ERR_ARG; }
  s->have_key = true;
  for (int d = 0; d < ndevices; d++) {
    Lane *L = new Lane();
    s->lanes.push_back(L);
    L->owner = s; L->device = devices[d]; L->depth = depth < 1 ? 1 : depth;
    pthread_mutex_init(&L->mu,NULL);
    pthread_cond_init(&L->cv_work,NULL); pthread_cond_init(&L->cv_open,NULL); pthread_cond_init(&L->cv_launched,NULL);
    int rc = gmapdp_create(&L->ctx,devices[d]);
    if (rc != GMAPDP_OK) { s->err = gmapdp_last_error(L->ctx); return rc; }
    for (int k = 0; k < (nflights < 3 ? 3 : nflights); k++) {
      Flight *F = new Flight();
      L->all.push_back(F);
      F->lane = L;
      rc = gdp_flight_create(L->ctx,&F->dev,s->max_boxes,s->seq_cap,s->prob_cap,s->script_cap);
      if (rc != GMAPDP_OK) { s->err = gmapdp_last_error(L->ctx); return rc; }
      F->bucket.resize(s->max_boxes); F->t_submit.resize(s->max_boxes);
      F->reset();
      if (k == 0) L->open = F; else L->free_list.push_back(F);
    }
    if (pthread_create(&L->launcher,NULL,launcher_main,L) != 0 || pthread_create(&L->completer,NULL,completer_main,L) != 0) {
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

extern "C" const char *gmapdp_stream_error (const gmapdp_stream *s) { return s ? s->err.c_str() : "null stream"; }
extern "C" int gmapdp_stream_ndevices (const gmapdp_stream *s) { return (int) s->lanes.size(); }

extern "C" int gmapdp_stream_submit (gmapdp_stream *s, const gmapdp_box *box, const uint8_t *seq, size_t seqbytes,
				     const double *probs, size_t nprobs, gmapdp_ticket *ticket) {
  GdpBoxGeom g;
  if (gdp_box_geometry(box,&g) != GMAPDP_OK) { s->err = "bad box"; return GMAPDP_ERR_ARG; }
  const size_t sb = align16(seqbytes);
  if (sb > s->seq_cap || nprobs > s->prob_cap || g.script_words > s->script_cap) { s->err = "box larger than a flight"; return GMAPDP_ERR_CAPACITY; }
  /* a thread stays on one lane (device) for its lifetime */
  int li = (int) (intptr_t) pthread_getspecific(s->lane_key) - 1;
  if (li < 0) {
    li = (int) (s->next_lane.fetch_add(1) % (unsigned) s->lanes.size());
    pthread_setspecific(s->lane_key,(void *) (intptr_t) (li + 1));
  }
  Lane *L = s->lanes[li];

  pthread_mutex_lock(&L->mu);
  Flight *F;
  for (;;) {
    if (L->failed || L->stop) { s->err = L->failed ? L->err : "stream is shutting down"; pthread_mutex_unlock(&L->mu); return GMAPDP_ERR_CUDA; }
    F = L->open;
    if (F && !F->full && F->n < s->max_boxes && F->seq_used + sb <= s->seq_cap && F->prob_used + nprobs <= s->prob_cap &&
	F->script_need + g.script_words <= s->script_cap) break;
    if (F && F->n > 0) { F->full = true; pthread_cond_signal(&L->cv_work); }
    pthread_cond_wait(&L->cv_open,&L->mu);
  }
  const int idx = F->n++;
  const size_t sbase = F->seq_used, pbase = F->prob_used;
  F->seq_used += sb; F->prob_used += nprobs; F->script_need += g.script_words;
  if (g.ws_words > F->ws_words[g.kind]) F->ws_words[g.kind] = g.ws_words;
  if (g.cols > F->maxcols[g.kind]) F->maxcols[g.kind] = g.cols;
  F->cnt[g.kind]++;
  F->bucket[idx] = g.bucket;
  F->writers.fetch_add(1,std::memory_order_relaxed);
  if (idx == 0) pthread_cond_signal(&L->cv_work);
  pthread_mutex_unlock(&L->mu);

  F->t_submit[idx] = now_s();
  gmapdp_box x = *box;
  x.qL_off += (uint32_t) sbase; x.qR_off += (uint32_t) sbase; x.gL_off += (uint32_t) sbase; x.gLalt_off += (uint32_t) sbase;
  x.gR_off += (uint32_t) sbase; x.gRalt_off += (uint32_t) sbase;
  x.probL_off += (uint32_t) pbase; x.probR_off += (uint32_t) pbase;
  F->dev->h_boxes[idx] = x;
  if (seqbytes) memcpy(F->dev->h_seq + sbase,seq,seqbytes);
  if (nprobs) memcpy(F->dev->h_probs + pbase,probs,nprobs * sizeof(double));
  F->writers.fetch_sub(1,std::memory_order_release);
  ticket->flight = F; ticket->index = idx; ticket->lane = li;
  return GMAPDP_OK;
}

extern "C" int gmapdp_stream_wait (gmapdp_stream *s, const gmapdp_ticket *ticket, const gmapdp_result **result, const uint32_t **ops) {
  Flight *F = (Flight *) ticket->flight;
  int d;
  while ((d = F->done.load(std::memory_order_acquire)) == 0) futex_wait(&F->done,0);
  if (d != 1) { s->err = F->lane->err; return F->rc ? F->rc : GMAPDP_ERR_CUDA; }
  const gmapdp_result *r = F->results + ticket->index;
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

extern "C" void gmapdp_stream_stats (const gmapdp_stream *s, double *out) {
  for (int k = 0; k < GMAPDP_STREAM_NSTATS; k++) out[k] = 0.0;
  for (Lane *L : s->lanes) {
    pthread_mutex_lock(&L->mu);
    out[0] += L->boxes; out[1] += L->flights; if (L->largest > out[2]) out[2] = L->largest;
    out[4] += L->t_flight; out[5] += L->t_wait; out[6] += (double) gmapdp_launch_count(L->ctx);
    out[7] += L->h2d; out[8] += L->d2h;
    pthread_mutex_unlock(&L->mu);
  }
}
