/* Internal seams between the translation units of libgmapdp_b200.so (not part of the C ABI):
 *   - what the chaining engine (gmapchain_kernels.cu) needs from the context that gmapdp_kernels.cu owns;
 *   - "flights", the device half of the streaming runtime (gmapdp_stream.cpp): one flight = the pinned host
 *     staging and the device buffers of one small batch in flight. */
#ifndef GMAPDP_INTERNAL_H
#define GMAPDP_INTERNAL_H

#include <string>
#include <vector>
#include "../../include/gmapdp_b200.h"

struct GdpCtxView {
  int device, sm_count;
  std::string *err;
  long *launches;
  void **chain;
  void (**chain_free) (void *);
  std::vector<void *> *grave;	/* outgrown device buffers, freed when the context is destroyed (cudaFree synchronises the
				   whole device: never while other engines' kernels are in flight) */
};
GdpCtxView gmapdp_ctx_view (gmapdp_ctx *ctx);

#define GDP_NKINDS 4		/* kernel specialisations: 0 single, 1 end5/end3, 2 genome, 3 cdna */

/* per-box planning data, computed by the thread that submits the box */
struct GdpBoxGeom {
  int kind;			/* which kernel serves it */
  int bucket;			/* launch-order key: kind, then decreasing work (quantised) */
  double work;			/* estimate of the box's duration on its warp (arbitrary unit, monotone) */
  int steps;			/* wavefront steps its warp executes one after the other: stripes x columns, passes x length */
  int cols;			/* shared-memory columns it needs (kinds 0 and 3) */
  size_t ws_words;		/* per-warp workspace */
  size_t script_words;		/* upper bound of its edit script */
};
/* GMAPDP_ERR_ARG for a malformed box (negative or > 32767 lengths / bands, positive penalties, bad mode) */
int gdp_box_geometry (const gmapdp_box *b, GdpBoxGeom *g);
int gdp_bucket_count (void);

struct GdpFlight {
  gmapdp_ctx *ctx;
  int max_boxes; size_t pool_cap, script_cap;	/* bytes; words */
  /* pinned host staging, filled in place by the submitting threads:
       h_in   = [32 B control block, zero][boxes: n x 96 B][launch order: n x 4 B]   -- ONE H2D copy of 32 + 100 n bytes
       h_pool = sequences and MaxEnt probabilities of the boxes (probabilities 8-aligned; a box's prob offsets are in
                doubles from the start of the pool)                                    -- ONE H2D copy of the used bytes
       h_out  = [results: n x 64 B][script words]                                      -- ONE D2H copy */
  unsigned char *h_in, *h_pool, *h_out;
  unsigned char *d_in, *d_pool, *d_out;
  void *ev_start, *ev_in, *ev_done;	/* cudaEvent_t */
  int n; size_t script_need;
  bool timed; float gpu_ms, copy_ms;	/* GMAPDP_STREAM_TIMING: device time of the last flight (all of it / its uploads) */
};
static inline gmapdp_box *gdp_flight_boxes (GdpFlight *f) { return reinterpret_cast<gmapdp_box *>(f->h_in + 32); }
static inline int *gdp_flight_order (GdpFlight *f, int n) { return reinterpret_cast<int *>(f->h_in + 32 + (size_t) n * sizeof(gmapdp_box)); }
int gdp_flight_create (gmapdp_ctx *ctx, GdpFlight **f, int max_boxes, size_t pool_cap, size_t script_cap);
void gdp_flight_destroy (GdpFlight *f);
/* asynchronous: two H2D copies, ONE kernel that serves boxes of every kind (a flight is small: what counts is the
   number of driver calls, not the instruction-cache footprint), one D2H copy, one event.  gdp_flight_order(f,n) lists
   the box ids by decreasing duration estimate. */
int gdp_flight_launch (GdpFlight *f, int n, size_t poolbytes, size_t script_need, size_t ws_words, int maxcols);
int gdp_flight_poll (GdpFlight *f);		/* 1 finished, 0 still running, < 0 error */
int gdp_flight_started (GdpFlight *f);	/* GMAPDP_STREAM_TIMING: 1 once the device has begun the flight */
int gdp_flight_wait (GdpFlight *f);		/* blocks; 0 or error */
/* after completion: the results (n entries) and the script pool in the pinned output buffer */
int gdp_flight_results (GdpFlight *f, const gmapdp_result **results, const uint32_t **script);

#endif
