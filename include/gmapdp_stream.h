/* gmapdp_stream -- the batching runtime of the drop-in: DP boxes submitted by many host threads, one at a
 * time, are gathered into device batches ("flights") and run on one or more B200s.
 *
 * Why it exists.  Behind the reference's call sites (stage3.c traverse_single_gap :8999, traverse_cdna_gap :9181,
 * traverse_genome_gap :9341, extend_ending5/3 :10281/:10519) every DP call is synchronous and the calls of one
 * query form a dependent chain, so a batch can only be made of calls of DIFFERENT worker threads
 * (gmap.c:4867 worker_thread, `gmap -t N').  The reference has no such component: its threads never meet
 * (gmap.c:4895-4907, per-worker Dynprog_T / Pairpool_T).  This runtime is the meeting point:
 *
 *   worker thread                          service thread of a lane (one lane per GPU and box size class)
 *   -------------                          ---------------------------------------------------------------
 *   gmapdp_stream_submit(box, mailbox) --> closes the open flight when the rules of gmapdp_stream.cpp say so (at once
 *     reserves a slot in the open flight,  if the lane is idle; otherwise in time to queue behind the flight that is
 *     copies its sequences into the        out), orders the boxes, issues 2 H2D copies + 1 kernel + 1 D2H copy;
 *     pinned staging                       polls the flight's event; when it has fired, copies every box's result and
 *   gmapdp_stream_wait(mailbox)            edit script into its owner's mailbox, wakes the first two owners and
 *     sleeps on the mailbox's futex;       recycles the flight
 *     once woken, wakes two more owners
 *     (tree), then reads its mailbox
 *   ... replays its own edit script (in parallel with all other workers) ...
 *
 * Flights are self-clocking: with an idle GPU a flight leaves with whatever it holds (latency), under load the
 * next one fills while the previous one runs (throughput); no timeouts, no global lock around device work, no
 * thundering herd, and a flight never waits for a worker thread.  Results never depend on how boxes were grouped.
 *
 * Multi-GPU: every device has its own lanes (context, stream, service thread, flights -- one lane per box size class); a
 * worker thread is pinned to one device for its lifetime, so the dependent chain of a query stays on one device
 * (SURVEY.md section 8e).  No collective.
 * There is no CPU fallback: creation fails without an sm_100 device.
 */
#ifndef GMAPDP_STREAM_H
#define GMAPDP_STREAM_H

#include "gmapdp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmapdp_stream gmapdp_stream;

/* Where a box's result lands.  The caller owns it (one per thread is enough) and provides the script buffer; everything
 * else is written by the runtime.  After gmapdp_stream_wait: `result' (its script_off is 0) and `ops' (script_lenA +
 * script_lenB words) stay valid until the mailbox is submitted again. */
typedef struct gmapdp_mailbox {
  gmapdp_result result;
  uint32_t *ops;		/* caller's buffer for the edit script */
  size_t ops_cap;		/* its capacity in words: at least rlenL + glenL + rlenR + glenR + 8 of the box */
  int state;			/* futex word: 0 pending, 1 landed, 2 lost */
  int rc;
  struct gmapdp_mailbox *wake[2];	/* owners this one wakes when it is woken (the runtime's wake-up tree) */
  double t_submit;
} gmapdp_mailbox;

/* devices[ndevices]: CUDA device ordinals.  max_boxes: capacity of a flight (0 = default 2048). */
int gmapdp_stream_create (gmapdp_stream **s, const int *devices, int ndevices, int max_boxes);
void gmapdp_stream_destroy (gmapdp_stream *s);
const char *gmapdp_stream_error (const gmapdp_stream *s);
int gmapdp_stream_ndevices (const gmapdp_stream *s);
/* Resident genome and MaxEnt tables for every device of the stream (gmapdp_genome_create in gmapdp_b200.h): boxes may
 * then carry genome coordinates (gflags).  Once, before the first such box. */
int gmapdp_stream_genome (gmapdp_stream *s, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *maxent);

/* Submits one box.  The box's *_off fields are offsets into `seq' / `probs' (as produced by the entry points of
 * gmapdp_shim.h on a private batch, GmapDP_batch_device_view); both arrays are copied before the call returns.
 * Blocks only while no flight of the lane can take the box.  A thread may have several mailboxes in flight. */
int gmapdp_stream_submit (gmapdp_stream *s, const gmapdp_box *box, const uint8_t *seq, size_t seqbytes,
			  const double *probs, size_t nprobs, gmapdp_mailbox *mailbox);
/* Blocks until the box's flight has landed; the mailbox then holds the result. */
int gmapdp_stream_wait (gmapdp_stream *s, gmapdp_mailbox *mailbox);

/* counters since creation, summed over lanes:
 *   [0] boxes  [1] flights  [2] largest flight  [3] seconds of device time (with GMAPDP_STREAM_TIMING=1)  [4] seconds a flight spent between launch and completion (sum)
 *   [5] seconds boxes waited between submit and wake-up (sum)  [6] kernel launches  [7] bytes host->device  [8] bytes device->host */
#define GMAPDP_STREAM_NSTATS 9
void gmapdp_stream_stats (const gmapdp_stream *s, double *out);
/* Lanes: every device has one lane per box size class (small boxes never share a flight with a large one, whose warp
 * would keep the whole flight waiting).  out[11] = boxes, flights, largest flight, mean flight latency (us), mean device
 * time per flight (us; needs GMAPDP_STREAM_TIMING=1), mean upload time per flight (us), then the host stages of a flight
 * (us): closed -> launch call, the launch call, launch returned -> completion seen, root wake-ups,
 * launch returned -> device seen to have begun (GMAPDP_STREAM_TIMING=1).
 * Lane = device index x classes + class. */
int gmapdp_stream_nlanes (const gmapdp_stream *s);
/* how the lanes' service threads are scheduled: 1 SCHED_FIFO (GMAPDP_STREAM_RT=1 and permitted), 2 short EEVDF slices, 3 default */
int gmapdp_stream_sched_mode (const gmapdp_stream *s);
int gmapdp_stream_lane_stats (const gmapdp_stream *s, int lane, double *out);

#ifdef __cplusplus
}
#endif
#endif
