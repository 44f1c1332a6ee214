/* gmapdp_shim -- host side of the boundary: the reference's five DP entry points as a batch API.
 *
 * Each GmapDP_*_gap call mirrors one reference function (same argument meaning, same exits, same
 * out-parameters) but, instead of filling matrices on the CPU, queues a device box:
 *
 *   GmapDP_single_gap  <-> Dynprog_single_gap  /root/reference/src/dynprog_single.c:428  (dynprog_single.h:23)
 *   GmapDP_genome_gap  <-> Dynprog_genome_gap  dynprog_genome.c:3287                     (dynprog_genome.h:23)
 *   GmapDP_cdna_gap    <-> Dynprog_cdna_gap    dynprog_cdna.c:786                        (dynprog_cdna.h:17)
 *   GmapDP_end5_gap    <-> Dynprog_end5_gap    dynprog_end.c:1293                        (dynprog_end.h:24)
 *   GmapDP_end3_gap    <-> Dynprog_end3_gap    dynprog_end.c:1924                        (dynprog_end.h:46)
 *
 * Differences from the reference signatures, all forced by the process boundary:
 *   - Genome_T/chroffset/chrhigh/watsonp are replaced by the genomic segment chars the reference
 *     itself would fetch with Genome_get_segment_right/left (genome.c:11023/11079); the caller
 *     (the stage3-side binding) fetches them with the reference's own code.
 *   - Dynprog_T is replaced by the two limits the entry points read from it (max_rlength,
 *     max_glength; dynprog.c:602-627), given once per batch.
 *   - for genome gaps the MaxEnt splice-site probabilities (dynprog_genome.c:970-1061, generated
 *     table code in maxent_hr.c) are passed in as two arrays.
 *   - results are not returned at once: the call returns an id; GmapDP_batch_run() executes every
 *     queued box in one device batch and GmapDP_result() hands back the out-parameters and the pair
 *     list (head first, as the reference's List_T).  Calls that the reference resolves without a
 *     fill (size checks, single_gap_simple, genome_gap_simple, QUERYEND_NOGAPS, require_pos_score_p)
 *     are resolved on the host exactly as there and never reach the device.
 *
 * iout / dout layouts (the reference's out-parameters in order):
 *   single, end5, end3: iout[6] = dynprogindex, traceback_score, nmatches, nmismatches, nopens, nindels
 *   genome: iout[10] = dynprogindex, new_leftgenomepos, new_rightgenomepos, traceback_score, nmatches,
 *           nmismatches, nopens, nindels, exonhead, introntype ; dout[2] = left_prob, right_prob
 *   cdna:   iout[3] = dynprogindex, traceback_score, incompletep
 * Out-parameters that the reference leaves unwritten on an exit path read GMAPDP_UNSET.
 */
#ifndef GMAPDP_SHIM_H
#define GMAPDP_SHIM_H

#include "gmapdp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the observable fields of struct Pair_T (pairdef.h:12-51) */
typedef struct gmapdp_pair {
  int querypos, genomepos, queryjump, genomejump, dynprogindex, introntype, gapp;
  char cdna, comp, genome, genomealt;
  double donor_prob, acceptor_prob;
} gmapdp_pair;

/* How the lists are stored and handed out without a copy (GmapDP_result_view): one 8-BYTE record per pair -- the fields
 * that differ from pair to pair -- and, for the few gap holders of a list (Pairpool_push_gapholder, pairpool.c:375), an entry
 * in a side table.  Positions are relative to the side of the call the pair lies on (a genome / cdna gap has two, up to
 * 32767 long each; every other call one):
 *     side      = (comp >> 7) & 1                    (the COMP_* characters are 7-bit)
 *     querypos  = base->q[side] + qrel,   genomepos = base->g[side] + grel,   comp = comp & 0x7f
 * qrel == GMAPDP_CPAIR_GAP marks a gap holder: querypos = genomepos = -1, blank chars, dynprogindex 0, and the jumps /
 * introntype / probabilities of gaps[grel].  The other fields of a gmapdp_pair follow: gapp = (gap holder); an ordinary
 * pair has queryjump = genomejump = introntype = 0, both probabilities 0.0 and the call's dynprogindex.
 * (A large batch writes ~10^9 pairs.  At 48 bytes each the host's memory traffic, not the GPU, set the end-to-end rate;
 * at 16 bytes -- absolute positions -- the replay of the benchmark step still took 320 ms on 16 host threads next to the
 * device's 206 ms.) */
#define GMAPDP_CPAIR_GAP 0xffffu
typedef struct gmapdp_cpair {
  uint16_t qrel, grel;
  char cdna, comp, genome, genomealt;
} gmapdp_cpair;

typedef struct gmapdp_pairbase {
  int q[2], g[2];		/* per side: what the relative positions of its pairs are relative to */
} gmapdp_pairbase;

typedef struct gmapdp_gapinfo {
  int queryjump, genomejump, introntype, pad_;
  double donor_prob, acceptor_prob;
} gmapdp_gapinfo;

#define GMAPDP_UNSET (-999)

/* Resident genome (device side: gmapdp_genome_create / _attach in gmapdp_b200.h).  The
 * entry points keep their character arguments -- the replay needs the characters for the pairs -- but a call announced
 * by GmapDP_batch_next_coords sends the device COORDINATES instead: where char 0 of each segment array lies in the
 * genome and in which direction the array runs, and for genome gaps where entry 0 of each probability array lies, its
 * direction and which of Maxent_hr_{donor,acceptor,antidonor,antiacceptor}_prob (0..3) fills it (the arrays of
 * dynprog_genome.c:970-1061; the entry point's own two lookups are then served by the same arithmetic on the host, and
 * left_probabilities / right_probabilities may be NULL).  Ignored for calls with an alternate genome. */
typedef struct gmapdp_coords {
  uint32_t chroffset, chrhigh;
  uint32_t gposL, gposR;	/* genome coordinate of element 0 of the L (or only) / R segment array */
  int negL, negR;		/* the array runs towards lower coordinates and is complemented (revcomp fetch) */
  int leftL, leftR;		/* fetched with Genome_get_segment_left ('*' below chroffset), else _right ('*' from chrhigh) */
  int probs;			/* genome gaps: probabilities from the MaxEnt model */
  uint32_t probposL, probposR;	/* splice coordinate of entry 0 of the left / right array */
  int probnegL, probnegR;	/* entry c lies at probpos - c (else + c) */
  int probkindL, probkindR;	/* 0 donor, 1 acceptor, 2 antidonor, 3 antiacceptor */
} gmapdp_coords;

/* Endalign_T, dynprog.h:27 */
enum { GMAPDP_QUERYEND_GAP = 0, GMAPDP_QUERYEND_INDELS = 1, GMAPDP_QUERYEND_NOGAPS = 2, GMAPDP_BEST_LOCAL = 3 };

typedef struct gmapdp_batch gmapdp_batch;

/* max_rlength / max_glength: what Dynprog_new(maxlookback,extraquerygap,maxpeelback,extramaterial_end,
 * extramaterial_paired) would compute (dynprog.c:602-627); GmapDP_maxlengths does that arithmetic. */
void GmapDP_maxlengths (int *max_rlength, int *max_glength, int maxlookback, int extraquerygap, int maxpeelback,
			int extramaterial_end, int extramaterial_paired);
/* NULL if a limit exceeds 32767 (rows and columns travel in 16-bit fields on the device) */
gmapdp_batch *GmapDP_batch_new (gmapdp_ctx *ctx, int max_rlength, int max_glength);
void GmapDP_batch_free (gmapdp_batch *b);
/* --indel-open / --indel-extend: what Dynprog_single_setup / _genome_setup / _end_setup (dynprog_single.c:101,
 * dynprog_genome.c:192, dynprog_end.c:120) store as user_open / user_extend / user_dynprog_p.  When set, single, genome and
 * end gaps use these penalties instead of the defect_rate table (dynprog_single.c:470, dynprog_genome.c:3367,
 * dynprog_end.c:1334,1964); cdna gaps never do (dynprog_cdna.c has no such branch).  Both in [-127, 0] (gmap.c:5425-5445). */
int GmapDP_batch_user_dynprog (gmapdp_batch *b, int user_open, int user_extend, int user_dynprog_p);
void GmapDP_batch_clear (gmapdp_batch *b);
/* the host's view of the genome the batch's context has attached (blocks and tables must stay valid) */
int GmapDP_batch_genome (gmapdp_batch *b, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *tables);
/* the next entry-point call queued on b refers to these coordinates (NULL: cancel) */
void GmapDP_batch_next_coords (gmapdp_batch *b, const gmapdp_coords *co);

int GmapDP_single_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
		       int rlength, int glength, int roffset, int goffset,
		       const char *gsequence, const char *gsequence_alt, int jump_late_p,
		       int extraband_single, int widebandp, double defect_rate);

/* rev_rsequence / rev_rsequenceuc point at the LAST char of the query part, as in the reference;
 * rev_gsequence is the fetched array (glength chars, ascending), read from its end. */
int GmapDP_end5_gap (gmapdp_batch *b, int dynprogindex, const char *rev_rsequence, const char *rev_rsequenceuc,
		     int rlength, int glength, int rev_roffset, int rev_goffset,
		     const char *rev_gsequence, const char *rev_gsequence_alt, int jump_late_p,
		     int extraband_end, double defect_rate, int endalign, int require_pos_score_p);
int GmapDP_end3_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
		     int rlength, int glength, int roffset, int goffset,
		     const char *gsequence, const char *gsequence_alt, int jump_late_p,
		     int extraband_end, double defect_rate, int endalign, int require_pos_score_p);

int GmapDP_genome_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
		       int rlength, int glengthL, int glengthR, int roffset, int goffsetL, int rev_goffsetR,
		       const char *gsequenceL, const char *gsequenceL_alt,
		       const char *rev_gsequenceR, const char *rev_gsequenceR_alt,
		       const double *left_probabilities, const double *right_probabilities,
		       int cdna_direction, int jump_late_p, int extraband_paired, double defect_rate,
		       int maxpeelback, int halfp, int finalp);

int GmapDP_cdna_gap (gmapdp_batch *b, int dynprogindex,
		     const char *rsequenceL, const char *rsequence_ucL,
		     const char *rev_rsequenceR, const char *rev_rsequence_ucR,
		     int rlengthL, int rlengthR, int glength, int roffsetL, int rev_roffsetR, int goffset,
		     const char *gsequence, const char *gsequence_alt,
		     const char *rev_gsequence, const char *rev_gsequence_alt,
		     int jump_late_p, int extraband_paired, double defect_rate);

/* Runs every queued device box (one gmapdp_run_batch) and completes all calls: the edit scripts are replayed into pair
 * lists (large batches: on several host threads, GMAPDP_REPLAY_THREADS). */
int GmapDP_batch_run (gmapdp_batch *b);
/* Puts the device calls back into their queued state, so that GmapDP_batch_run can be repeated on the same batch. */
void GmapDP_batch_rewind (gmapdp_batch *b);
/* resident variant for benchmarks: pack + upload once, then time the kernel alone */
int GmapDP_batch_upload (gmapdp_batch *b);
int GmapDP_batch_run_resident (gmapdp_batch *b, float *kernel_ms);
int GmapDP_batch_finish (gmapdp_batch *b);	/* download + replay after run_resident */
int GmapDP_batch_run_device (gmapdp_batch *b);	/* gmapdp_run_batch only: H2D + kernel + D2H, no pair replay */
int GmapDP_batch_download (gmapdp_batch *b);	/* D2H of results + scripts after run_resident */
unsigned long long GmapDP_batch_digest (const gmapdp_batch *b);	/* checksum of all device results + scripts */
long GmapDP_batch_cells_full (const gmapdp_batch *b);	/* cells of the single-gap boxes (full fills) ... */
long GmapDP_batch_cells8_full (const gmapdp_batch *b);	/* ... and their 8-bit share */
long GmapDP_batch_cells8 (const gmapdp_batch *b);	/* the share of GmapDP_batch_cells filled in 8-bit mode */

int GmapDP_batch_ncalls (const gmapdp_batch *b);
int GmapDP_batch_nboxes (const gmapdp_batch *b);	/* calls that reached the device */
long GmapDP_batch_cells (const gmapdp_batch *b);	/* algorithmic in-band cells of the queued boxes */
size_t GmapDP_batch_h2d_bytes (const gmapdp_batch *b);
size_t GmapDP_batch_d2h_bytes (const gmapdp_batch *b);
const char *GmapDP_batch_error (const gmapdp_batch *b);

/* For callers that execute the device boxes themselves (gmapdp_stream.h: batches shared by many threads): the device
 * part of the queued calls, and the completion of the calls from results produced elsewhere (results[k] = box k of the
 * view, its ops at script + results[k].script_off). */
int GmapDP_batch_device_view (const gmapdp_batch *b, const gmapdp_box **boxes, int *nboxes, const uint8_t **seqpool, size_t *seqbytes,
			      const double **probpool, size_t *nprobs);
int GmapDP_batch_complete (gmapdp_batch *b, const gmapdp_result *results, const uint32_t *script);

/* Returns the number of pairs (list head first), or -1 for the reference's NULL list.  If the list is longer than
 * maxpairs only maxpairs records are copied: compare the return value with maxpairs, or use GmapDP_result_view. */
int GmapDP_result (const gmapdp_batch *b, int id, int *iout, double *dout, gmapdp_pair *pairs, int maxpairs);
/* the same without a copy: pointers into the batch, valid until it is cleared; *dynprogindex = the index the call's
   ordinary pairs carry */
int GmapDP_result_view (const gmapdp_batch *b, int id, const int **iout, const double **dout, const gmapdp_cpair **pairs,
			const gmapdp_gapinfo **gaps, const gmapdp_pairbase **base, int *dynprogindex);
/* device-side view of one call (NULL if it never reached the device) */
const gmapdp_result *GmapDP_device_result (const gmapdp_batch *b, int id);

#ifdef __cplusplus
}
#endif
#endif
