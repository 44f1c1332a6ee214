/* gmapchain_b200 -- C ABI of the B200 (sm_100a) stage-2 chaining engine for GMAP.
 *
 * Replaces, behind the reference's static align_compute_lookback (/root/reference/src/stage2.c:4402; call sites
 * :6548 Stage2_compute, :6859 Stage2_compute_one, :7194 Stage2_compute_ends), the chaining dynamic programme over
 * k-mer hits and what follows it:
 *
 *   align_compute_scores_lookback  stage2.c:3667   (links + fwd_scores over mappings[querypos][hit])
 *   score_querypos_lookback_one    stage2.c:1073,  score_querypos_lookback_mult  stage2.c:1470
 *   revise_active_lookback         stage2.c:2956
 *   get_cells_fwd                  stage2.c:3437   (one best end cell per rootposition, ranked by score)
 *   traceback_one                  stage2.c:4140   (path of (querypos, position), 3' end pruned)
 *
 * One *problem* = one call of align_compute_lookback.  Results are identical to the reference's: the same
 * paths in the same order.  Both directions are served: the lookforward twins (align_compute_lookforward stage2.c:5062,
 * align_compute_scores_lookforward :4610, score_querypos_lookforward_one :2020, _mult :2404, revise_active_lookforward
 * :3034; call site :7020 Stage2_compute_starts) are the same engine with the GMAPCHAIN_F_LOOKFORWARD flag.
 * Scope of this version: the non-PMAP build, use_canonical_p == false
 * (gmap's default; the cross-species canonical test needs the genome on the device).
 * There is no CPU fallback: every compute entry point fails with GMAPDP_ERR_CUDA without an sm_100 device.
 *
 * Plain C, plain pointers and sizes.  The engine shares its context with the DP engine (gmapdp_b200.h).
 */
#ifndef GMAPCHAIN_B200_H
#define GMAPCHAIN_B200_H

#include "gmapdp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GMAPCHAIN_F_LOCALP          0x1
#define GMAPCHAIN_F_SKIP_REPETITIVE 0x2
#define GMAPCHAIN_F_FAVOR_RIGHT     0x4
#define GMAPCHAIN_F_MIDDLEP         0x8
#define GMAPCHAIN_F_LOOKFORWARD     0x10	/* align_compute_lookforward (stage2.c:5062): queryposes descending */

/* One chaining problem (48 bytes).  Per-querypos arrays live in four parallel pools indexed from q_off
 * (npositions, cumulative hit offset, minactive, maxactive); the hits of all queryposes, each sorted
 * ascending as Oligoindex_get_mappings leaves them, are contiguous in the position pool from p_off. */
typedef struct gmapchain_problem {
  int32_t querylength, querystart, queryend, indexsize;
  int32_t flags, max_nalignments;
  int32_t totalpositions, reserved;
  uint64_t q_off;		/* first querypos of this problem in the per-querypos pools */
  uint64_t p_off;		/* first hit of this problem in the position pool */
} gmapchain_problem;

typedef struct gmapchain_result {	/* 32 bytes */
  int32_t status;		/* 0 ok */
  int32_t npaths;		/* paths traced (stage2.c:4476: up to max_nalignments, plus ties at the best score, within 20 of it) */
  int32_t bestscore;		/* cells[0]->score, 0 if there is no positive cell */
  int32_t ncandidates;		/* cells within FINAL_SCORE_TOLERANCE of the best (diagnostic) */
  uint32_t path_off;		/* first path record of this problem */
  uint32_t reserved[3];
} gmapchain_result;

typedef struct gmapchain_path {		/* 32 bytes: the Cell_T the path starts from, and its pairs */
  int32_t score, rootposition, endposition, querypos, hit;
  int32_t npairs;
  uint32_t pair_off;		/* index into the pair pool (2 x int32 per pair: querypos, position) */
  int32_t reserved;
} gmapchain_path;

/* Stage2_setup (stage2.c:129): splicingp, sufflookback, nsufflookback, maxintronlen (gmap.c:269,270,347).
 * cross_species_p must be 0. */
int gmapchain_setup (gmapdp_ctx *ctx, int splicingp, int cross_species_p, int sufflookback, int nsufflookback, int maxintronlen);

/* Optional: allocates the device buffers for batches of up to these sizes once (they otherwise grow on demand). */
int gmapchain_reserve (gmapdp_ctx *ctx, int nproblems, size_t nquerypos, size_t npositions_total);

/* Host-buffer path: H2D of the pools, chaining + ranking + tracebacks on the device, D2H of results.
 * Pairs come in traceback order (highest querypos first).  Returns GMAPDP_ERR_CAPACITY, with *paths_used /
 * *pairs_used = what is needed, if an output pool is too small. */
int gmapchain_run_batch (gmapdp_ctx *ctx, const gmapchain_problem *problems, int nproblems,
			 const int32_t *npositions, const uint32_t *cumpositions, const uint32_t *minactive, const uint32_t *maxactive,
			 size_t nquerypos, const uint32_t *positions, size_t npositions_total,
			 gmapchain_result *results, gmapchain_path *paths, size_t paths_cap, size_t *paths_used,
			 int32_t *pairs, size_t pairs_cap, size_t *pairs_used);

/* Resident path (benchmarks): upload once, run many times, download when wanted. */
int gmapchain_upload (gmapdp_ctx *ctx, const gmapchain_problem *problems, int nproblems,
		      const int32_t *npositions, const uint32_t *cumpositions, const uint32_t *minactive, const uint32_t *maxactive,
		      size_t nquerypos, const uint32_t *positions, size_t npositions_total);
int gmapchain_run_resident (gmapdp_ctx *ctx, float *kernel_ms);
int gmapchain_download (gmapdp_ctx *ctx, gmapchain_result *results, gmapchain_path *paths, size_t paths_cap, size_t *paths_used,
			int32_t *pairs, size_t pairs_cap, size_t *pairs_used);
/* Parity/debug: the link matrix of the last run, 5 x int32 per hit (fwd_consecutive, fwd_rootposition, fwd_pos,
 * fwd_hit, 0) and fwd_scores, in position-pool order.  fwd_tracei is not kept (its storage is reused by the ranking). */
int gmapchain_download_links (gmapdp_ctx *ctx, int32_t *links, int32_t *scores, size_t npositions_total);

/* ---- host mirror of align_compute_lookback (argument meaning as stage2.c:4402) ---------------------------- */
typedef struct gmapchain_batch gmapchain_batch;
gmapchain_batch *GmapChain_batch_new (gmapdp_ctx *ctx);
void GmapChain_batch_free (gmapchain_batch *b);
void GmapChain_batch_clear (gmapchain_batch *b);
/* Queues one call; returns its id (>= 0) or a negative GMAPDP_ERR_*.  mappings[q] points at npositions[q] sorted hits
 * (ignored when npositions[q] <= 0).  use_canonical_p must be 0. */
int GmapChain_lookback (gmapchain_batch *b, uint32_t *const *mappings, const int *npositions, int totalpositions,
			const uint32_t *minactive, const uint32_t *maxactive, int querylength, int querystart, int queryend,
			int indexsize, int localp, int skip_repetitive_p, int use_canonical_p, int non_canonical_penalty,
			int favor_right_p, int middlep, int max_nalignments);
/* The same for align_compute_lookforward (stage2.c:5062; Stage2_compute_starts): the paths then come with their
 * HIGHEST querypos first, which is the order of the List_T the reference returns in that direction. */
int GmapChain_lookforward (gmapchain_batch *b, uint32_t *const *mappings, const int *npositions, int totalpositions,
			   const uint32_t *minactive, const uint32_t *maxactive, int querylength, int querystart, int queryend,
			   int indexsize, int localp, int skip_repetitive_p, int use_canonical_p, int non_canonical_penalty,
			   int favor_right_p, int middlep, int max_nalignments);
int GmapChain_batch_run (gmapchain_batch *b);		/* gmapchain_run_batch, growing the output pools as needed */
int GmapChain_batch_upload (gmapchain_batch *b);
int GmapChain_batch_run_resident (gmapchain_batch *b, float *kernel_ms);
int GmapChain_batch_download (gmapchain_batch *b);
int GmapChain_npaths (const gmapchain_batch *b, int id);
/* Path k of call id in the reference's list order (lookback: lowest querypos first); cell[5] = rootposition, endposition,
 * querypos, hit, score of the Cell_T it was traced from.  Returns the number of pairs (or -needed if cap is short). */
int GmapChain_path (const gmapchain_batch *b, int id, int k, int *cell, int *querypos, uint32_t *position, int cap);
int GmapChain_batch_ncalls (const gmapchain_batch *b);
long GmapChain_batch_nhits (const gmapchain_batch *b);		/* sum of totalpositions */
long GmapChain_batch_h2d_bytes (const gmapchain_batch *b);
long GmapChain_batch_d2h_bytes (const gmapchain_batch *b);
unsigned long long GmapChain_batch_digest (const gmapchain_batch *b);
const char *GmapChain_batch_error (const gmapchain_batch *b);

#ifdef __cplusplus
}
#endif
#endif
