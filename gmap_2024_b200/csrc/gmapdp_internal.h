/* Internal seams between the translation units of libgmapdp_b200.so (not part of the C ABI):
 *   - what the chaining engine (gmapchain_kernels.cu) needs from the context that gmapdp_kernels.cu owns;
 *   - "flights", the device half of the streaming runtime (gmapdp_stream.cpp): one flight = the pinned host
 *     staging and the device buffers of one small batch in flight. */
#ifndef GMAPDP_INTERNAL_H
#define GMAPDP_INTERNAL_H

#include <string>
#include "../../include/gmapdp_b200.h"

struct GdpCtxView {
  int device, sm_count;
  std::string *err;
  long *launches;
  void **chain;
  void (**chain_free) (void *);
};
GdpCtxView gmapdp_ctx_view (gmapdp_ctx *ctx);

#define GDP_NKINDS 4		/* kernel specialisations: 0 single, 1 end5/end3, 2 genome, 3 cdna */

/* per-box planning data, computed by the thread that submits the box */
struct GdpBoxGeom {
  int kind;			/* which kernel serves it */
  int bucket;			/* launch-order key: kind, then decreasing work (quantised) */
  int cols;			/* shared-memory columns it needs (kinds 0 and 3) */
  size_t ws_words;		/* per-warp workspace */
  size_t script_words;		/* upper bound of its edit script */
};
/* GMAPDP_ERR_ARG for a malformed box (negative or > 32767 lengths / bands, positive penalties, bad mode) */
int gdp_box_geometry (const gmapdp_box *b, GdpBoxGeom *g);
int gdp_bucket_count (void);

struct GdpFlight {
  gmapdp_ctx *ctx;
  int max_boxes; size_t seq_cap, prob_cap, script_cap;	/* script_cap in words */
  /* pinned host staging: inputs are written in place by the submitting threads */
  gmapdp_box *h_boxes; int *h_order; uint8_t *h_seq; double *h_probs;
  unsigned char *h_out;		/* [16 B: script cursor][results: n x 64 B][script words] -- one D2H copy */
  /* device twins */
  gmapdp_box *d_boxes; int *d_order; uint8_t *d_seq; double *d_probs; unsigned char *d_out;
  int *d_ctl;			/* 4 queue heads */
  void *ev_in, *ev_k[GDP_NKINDS], *ev_done;	/* cudaEvent_t */
  int n; size_t script_need;
};
int gdp_flight_create (gmapdp_ctx *ctx, GdpFlight **f, int max_boxes, size_t seq_cap, size_t prob_cap, size_t script_cap);
void gdp_flight_destroy (GdpFlight *f);
/* asynchronous: H2D of the used parts of the staging, the (up to) four kernels, D2H of cursor + results + script.
   h_order lists the box ids kind by kind (cnt[kind] of each), each kind by decreasing work. */
int gdp_flight_launch (GdpFlight *f, int n, size_t seqbytes, size_t nprobs, size_t script_need,
		       const size_t *ws_words, const int *maxcols, const int *cnt);
int gdp_flight_poll (GdpFlight *f);		/* 1 finished, 0 still running, < 0 error */
int gdp_flight_wait (GdpFlight *f);		/* blocks; 0 or error */
/* after completion: the results (n entries) and the script pool in the pinned output buffer */
int gdp_flight_results (GdpFlight *f, const gmapdp_result **results, const uint32_t **script);

#endif
