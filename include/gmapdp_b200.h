/* gmapdp_b200 -- C ABI of the B200 (sm_100a) alignment dynamic-programming engine for GMAP.
 *
 * This is the device boundary: plain C, plain pointers and sizes, no torch / CUDA types.
 * It replaces, behind the reference's own call sites in stage3.c, what the reference implements
 * in /root/reference/src:
 *
 *   fills      Dynprog_simd_8 / _16            dynprog_simd.c:2987 / :6562   (full band, E and F)
 *              Dynprog_simd_8_upper / _16_upper dynprog_simd.c:4304 / :7714   (upper triangle, E only)
 *              Dynprog_simd_8_lower / _16_lower dynprog_simd.c:5340 / :8586   (lower triangle, E only)
 *   searches   find_best_endpoint_{8,16}                       dynprog_end.c:143 / :220
 *              find_best_endpoint_to_queryend_indels_{8,16}    dynprog_end.c:358 / :437
 *              bridge_intron_gap_{8,16}_site_level             dynprog_genome.c:866 / :1742
 *              bridge_cdna_gap_{8,16}_ud                       dynprog_cdna.c:123 / :387
 *   tracebacks Dynprog_traceback_{8,16}{,_upper,_lower}        dynprog_simd.c:9154-9946
 *
 * One *box* = the device part of one call of Dynprog_single_gap / _genome_gap / _cdna_gap /
 * _end5_gap / _end3_gap (dynprog_single.h:23, dynprog_genome.h:23, dynprog_cdna.h:17,
 * dynprog_end.h:24,46).  The host side of those calls (argument checks, penalties from
 * defect_rate, the *_simple shortcuts, Pair_T list construction) lives in gmapdp_shim.h.
 *
 * Results are bit-exact with the reference's SIMD (avx2) dynprog: same scores, same best cells,
 * same traceback path.  There is no CPU fallback: every compute entry point fails with
 * GMAPDP_ERR_CUDA if no sm_100 device / kernel image is available.
 */
#ifndef GMAPDP_B200_H
#define GMAPDP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMAPDP_OK 0
#define GMAPDP_ERR_CUDA (-1)
#define GMAPDP_ERR_ARG (-2)
#define GMAPDP_ERR_CAPACITY (-3)

/* box modes */
enum { GMAPDP_SINGLE = 0, GMAPDP_GENOME = 1, GMAPDP_CDNA = 2, GMAPDP_END5 = 3, GMAPDP_END3 = 4 };

/* flags */
#define GMAPDP_F_LATE_L      0x01	/* jump_late_p passed to the L-side (or only) fills */
#define GMAPDP_F_LATE_R      0x02	/* jump_late_p passed to the R-side fills (= !jump_late_p, dynprog_genome.c:3538) */
#define GMAPDP_F_USE8        0x04	/* 8-bit saturation (use8p), else 16-bit */
#define GMAPDP_F_FINALP      0x08
#define GMAPDP_F_HALFP       0x10
#define GMAPDP_F_LASTROW     0x20	/* end modes: QUERYEND_INDELS search (last row only, initial best -inf) */
#define GMAPDP_F_NOTRACE     0x40	/* end modes: require_pos_score_p => the reference skips the traceback */
#define GMAPDP_F_BRIDGE_LATE 0x80	/* cdna bridge tie rule: >= (jump_late_p) instead of > */

/* gmapdp_box.gflags: where the genomic data of a box come from (see gmapdp_genome_attach) */
#define GMAPDP_G_SEG_L      0x01	/* the L (or only) segment is read from the resident genome: gL_off = coordinate of its char 0 */
#define GMAPDP_G_SEG_R      0x02	/* the same for the R segment (gR_off) */
#define GMAPDP_G_NEG_L      0x04	/* L segment runs towards lower coordinates and is complemented (revcomp fetch) */
#define GMAPDP_G_NEG_R      0x08
#define GMAPDP_G_LEFT_L     0x10	/* L segment was fetched "leftwards" (Genome_get_segment_left: '*' below chroffset); else */
#define GMAPDP_G_LEFT_R     0x20	/*   "rightwards" (Genome_get_segment_right: '*' from chrhigh on) */
#define GMAPDP_G_PROBS      0x40	/* genome mode: MaxEnt probabilities are computed on the device (probposL / probposR, probkindL / R); */
					/*   probL_off / probR_off = offsets (doubles) of the two arrays (glength + 1 entries each) in the */
					/*   context's device-side probability pool (batch path; flights evaluate into the workspace) */
#define GMAPDP_G_PSTEP_NEG_L 0x80	/* entry c of the left array is the coordinate probposL - c (else + c) */
#define GMAPDP_G_PSTEP_NEG_R 0x100

/* One DP box (96 bytes).  Sequences live in one byte pool; every *_off is an unsigned byte offset into it (pool < 4 GiB).
 * All sequence arrays are stored FORWARD (ascending memory = ascending coordinate); sides that the
 * reference addresses through "rev_" pointers set the corresponding REV flag bit in `revmask` and
 * are read from their last element backwards, exactly like rev_rsequence / rev_gsequence. */
typedef struct gmapdp_box {
  int32_t mode;
  int32_t flags;
  int32_t rlenL, rlenR;		/* rlength (single/end/genome: rlenR == rlenL); cdna: rlengthL, rlengthR */
  int32_t glenL, glenR;		/* glength (single/end/cdna: glenR == glenL); genome: glengthL, glengthR */
  int8_t  mismatchtype, open, extend, cdna_direction;	/* Mismatchtype_T dynprog.h:54; penalties are negative */
  int16_t lbandL, ubandL, lbandR, ubandR;		/* Dynprog_compute_bands dynprog.c:1246 */
  uint32_t qL_off, qR_off;	/* upper-cased query chars of the L / R side (rlenL / rlenR bytes) */
  uint32_t gL_off, gLalt_off;	/* genomic segment of the L (or only) side and its alt-genome twin (== gL_off if none) */
  uint32_t gR_off, gRalt_off;	/* genomic segment of the R side */
  uint32_t probL_off, probR_off;	/* genome mode: offsets (in doubles) of left/right MaxEnt probabilities, glen-1 entries each */
  int32_t offdiff;		/* genome: rev_goffsetR - goffsetL ; cdna: rev_roffsetR - roffsetL (bridge constraint) */
  int32_t revmask;		/* bit0: L side is read reversed (end5) ; bit1: R side is read reversed (genome, cdna) */
  /* resident-genome boxes (all zero otherwise) */
  uint32_t chroffset, chrhigh;	/* bounds of the chromosome the segments lie on (Univcoord_T) */
  uint16_t gflags;		/* GMAPDP_G_* */
  uint8_t  probkindL, probkindR;	/* 0 donor, 1 acceptor, 2 antidonor, 3 antiacceptor (Maxent_hr_*_prob) */
  uint32_t probposL, probposR;	/* GMAPDP_G_PROBS: splice coordinate of entry 0 of the left / right probability array */
} gmapdp_box;

/* Edit script: one uint32 per op, in traceback order (from the best cell back to the origin).
 *   op = (length << 2) | kind ;  kind 0 = `length' DIAG steps, 1 = horizontal gap (genome skip) of
 *   `length' columns, 2 = vertical gap (query skip) of `length' rows.  The terminal "flush to the
 *   axis" indel of the reference's tracebacks is NOT in the script: the replayer derives it from
 *   the (r, c) left over, as Dynprog_traceback_* do (dynprog_simd.c:9282-9310). */
typedef struct gmapdp_result {
  int32_t status;		/* 0 = filled; 1 = bridge rejected (finalscore < 0) */
  int32_t finalscore;		/* best-endpoint score / bridge score / matrix corner (single) */
  int32_t bestrL, bestcL, bestrR, bestcR;	/* end modes: bestr/bestc in the L fields */
  int32_t tb_score, nmatches, nmismatches, nopens, nindels;	/* accumulated by the device tracebacks (scores.h:5-10) */
  int32_t script_off;		/* word offset of this box's ops in the script pool */
  int32_t script_lenA;		/* ops of the first traced segment (R side for genome/cdna; the only one otherwise) */
  int32_t script_lenB;		/* ops of the second traced segment (L side for genome/cdna) */
  int32_t cells;		/* algorithmic in-band cells filled for this box */
  int32_t reserved;
} gmapdp_result;

typedef struct gmapdp_ctx gmapdp_ctx;

/* number of CUDA devices visible to the process (0 if none / no driver) */
int gmapdp_device_count (void);
/* Creates an engine on CUDA device `device` (one context per GPU / per process rank). */
int gmapdp_create (gmapdp_ctx **ctx, int device);
void gmapdp_destroy (gmapdp_ctx *ctx);
const char *gmapdp_last_error (const gmapdp_ctx *ctx);
/* number of SMs and the persistent grid the engine launches (for reporting) */
int gmapdp_device_info (const gmapdp_ctx *ctx, int *sm_count, int *grid_blocks, int *block_threads);

/* Host-buffer path: H2D copy of the batch, fills + searches + tracebacks on the device, D2H of
 * results and scripts.  `results` has nboxes entries; `script` has script_cap words.
 * Returns GMAPDP_ERR_CAPACITY (with *script_used = words needed) if the script pool is too small. */
int gmapdp_run_batch (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
		      const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs,
		      gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used);

/* The same with a completion callback: on_chunk(user,first,n) is called on the calling thread as soon as results[first ..
 * first + n) and their edit scripts are in the caller's buffers -- chunk by chunk, in order, while the device works on the
 * following chunks -- so that the caller's post-processing overlaps the kernels.  A batch that runs as one chunk gets
 * one call at the end. */
typedef void (*gmapdp_chunk_fn) (void *user, int first_box, int nboxes);
int gmapdp_run_batch_chunks (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
			     const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs,
			     gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used,
			     gmapdp_chunk_fn on_chunk, void *user);

/* Resident genome and splice-site model.  `blocks' is the reference's compressed genome
 * exactly as genome.c reads it (three 32-bit words -- high, low, flags -- per 32 nt; uncompress_mmap, genome.c:8985): it
 * is copied to the device once and stays there.  `maxent' holds the parameters of the reference's MaxEnt model (the static
 * arrays of maxent_hr.c; 16384 doubles per score table, 16 per dinucleotide table).  One genome object serves every
 * context of its device (gmapdp_genome_attach).  Boxes with gflags then carry genome COORDINATES instead of genomic
 * characters and probability arrays: segments are decoded and Maxent_hr_{donor,acceptor,antidonor,antiacceptor}_prob
 * (maxent_hr.c:27357-27650) evaluated on the device, bit-identically (IEEE double).  Only genomealt == genome. */
typedef struct gmapdp_maxent_tables {
  const double *donor_plus, *donor_di_plus, *acc1_plus, *acc2_plus, *acc3_plus, *acc_di_plus, *acc467_plus, *acc589_plus;
  const double *donor_minus, *donor_di_minus, *acc1_minus, *acc2_minus, *acc3_minus, *acc_di_minus, *acc467_minus, *acc589_minus;
} gmapdp_maxent_tables;
typedef struct gmapdp_genome gmapdp_genome;
int gmapdp_genome_create (gmapdp_genome **g, int device, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *maxent);
void gmapdp_genome_destroy (gmapdp_genome *g);
int gmapdp_genome_attach (gmapdp_ctx *ctx, const gmapdp_genome *g);	/* g must outlive ctx's last run */
/* the same arithmetic on the host (for callers that need single values): char of one position, one probability */
int gmapdp_genome_host_char (const uint32_t *blocks, size_t nwords, uint32_t pos);
double gmapdp_maxent_host_prob (const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *maxent, int kind, uint32_t splice_pos, uint32_t chroffset);
/* device check of the same: probs[k] = probability `kind[k]' at pos[k] (testing; n values) */
int gmapdp_maxent_eval (gmapdp_ctx *ctx, const int *kind, const uint32_t *pos, uint32_t chroffset, int n, double *probs);

/* Resident path (benchmarks, pipelined callers): upload once, run any number of times with the
 * inputs already in HBM, download when wanted.  kernel_ms (may be NULL) receives the CUDA-event
 * time of the DP kernel alone, measured on the stream it was launched on. */
int gmapdp_upload (gmapdp_ctx *ctx, const gmapdp_box *boxes, int nboxes,
		   const uint8_t *seqpool, size_t seqbytes, const double *probpool, size_t nprobs);
int gmapdp_run_resident (gmapdp_ctx *ctx, float *kernel_ms);
int gmapdp_download (gmapdp_ctx *ctx, gmapdp_result *results, uint32_t *script, size_t script_cap, size_t *script_used);
/* A batch is served by four specialisations of the DP kernel: one for the single-gap boxes (full fills) and
 * one each for the end, genome and cdna boxes (E-only fills + searches / bridges).  gmapdp_run_resident queues
 * them back to back on one stream (each is a persistent grid over all SMs), so their CUDA-event durations are
 * exact: full_ms = the single-gap kernel, tri_ms = the other three together; ms[4] = single, end, genome, cdna.
 * gmapdp_run_batch queues the kernels of its chunks the same way, behind each chunk's upload. */
int gmapdp_last_kernel_ms (const gmapdp_ctx *ctx, float *full_ms, float *tri_ms);
int gmapdp_last_kernel_ms4 (const gmapdp_ctx *ctx, float *ms);
/* number of kernel launches issued by this context so far */
long gmapdp_launch_count (const gmapdp_ctx *ctx);

/* Pinned host memory for the batch buffers (optional: pageable buffers work, through a staging copy). */
void *gmapdp_host_alloc (size_t bytes);
void gmapdp_host_free (void *p);
int gmapdp_host_register (void *p, size_t bytes);
int gmapdp_host_unregister (void *p);

#ifdef __cplusplus
}
#endif
#endif
