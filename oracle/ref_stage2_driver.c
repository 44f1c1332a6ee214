/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * ctypes front end over the reference's stage-2 chaining.  The chaining
 * functions are `static` in the reference's stage2.c, so this translation unit compiles that file IN PLACE
 * (the #include below resolves through -I/root/reference/src; nothing is copied) and adds plain-C entry
 * points after it:
 *
 *   refs2_scores[_fwd] : align_compute_scores_lookback / _lookforward (stage2.c:3667 / :4610) -> links, fwd_scores, ranked cells
 *   refs2_paths[_fwd]  : align_compute_lookback / _lookforward        (stage2.c:4402 / :5062) -> the traced paths (querypos, position)
 */
#include "stage2.c"

static Pairpool_T s2_pairpool = NULL;
static Cellpool_T s2_cellpool = NULL;

void refs2_setup (int splicingp_in, int cross_species_p, int sufflookback_in, int nsufflookback_in, int maxintronlen_in) {
  Stage2_setup(splicingp_in ? true : false,cross_species_p ? true : false,/*suboptimal_score_start*/0,/*suboptimal_score_end*/0,
	       sufflookback_in,nsufflookback_in,maxintronlen_in,STANDARD,/*snps_p*/false);
  if (s2_pairpool == NULL) {
    s2_pairpool = Pairpool_new();
    s2_cellpool = Cellpool_new();
  }
}

static Chrpos_T **make_mappings (const unsigned int *positions, const int *npositions, int querylength) {
  Chrpos_T **m = (Chrpos_T **) malloc((size_t) querylength * sizeof(Chrpos_T *));
  size_t off = 0;
  int q;
  for (q = 0; q < querylength; q++) {
    m[q] = (Chrpos_T *) &positions[off];
    if (npositions[q] > 0) off += (size_t) npositions[q];
  }
  return m;
}

/* links_out: 5 ints per position (fwd_consecutive, fwd_rootposition, fwd_pos, fwd_hit, fwd_tracei), CSR order;
   cells_out: 5 ints per ranked cell (rootposition, endposition, querypos, hit, score).  Returns ncells. */
static int scores_dir (int forwardp, const unsigned int *positions, const int *npositions, int querylength, int totalpositions,
		  const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize,
		  int localp, int skip_repetitive_p, int favor_right_p, int middlep,
		  int *links_out, int *scores_out, int *cells_out, int cells_cap) {
  Chrpos_T **mappings = make_mappings(positions,npositions,querylength);
  struct Link_T **links = Linkmatrix_1d_new(querylength,(int *) npositions,totalpositions);
  int **fwd_scores = intmatrix_1d_new(querylength,(int *) npositions,totalpositions);
  int *firstactive = (int *) calloc(querylength,sizeof(int)), *nactive = (int *) calloc(querylength,sizeof(int));
  int ncells = 0, i;
  Cell_T *cells;

  Cellpool_reset(s2_cellpool);
  if (forwardp) {
    cells = align_compute_scores_lookforward(&ncells,links,fwd_scores,mappings,(int *) npositions,totalpositions,
					     /*oned_matrix_p*/true,(Chrpos_T *) minactive,(Chrpos_T *) maxactive,firstactive,nactive,s2_cellpool,
					     querystart,queryend,querylength,/*genome*/NULL,/*genomealt*/NULL,
					     /*chroffset*/0,/*chrhigh*/0,/*plusp*/true,indexsize,
					     localp ? true : false,skip_repetitive_p ? true : false,
					     /*use_canonical_p*/false,/*non_canonical_penalty*/NON_CANONICAL_PENALTY_MIDDLE,
					     favor_right_p ? true : false,middlep ? true : false);
  } else {
    cells = align_compute_scores_lookback(&ncells,links,fwd_scores,mappings,(int *) npositions,totalpositions,
					  /*oned_matrix_p*/true,(Chrpos_T *) minactive,(Chrpos_T *) maxactive,firstactive,nactive,s2_cellpool,
					  querystart,queryend,querylength,/*genome*/NULL,/*genomealt*/NULL,
					  /*chroffset*/0,/*chrhigh*/0,/*plusp*/true,indexsize,
					  localp ? true : false,skip_repetitive_p ? true : false,
					  /*use_canonical_p*/false,/*non_canonical_penalty*/NON_CANONICAL_PENALTY_MIDDLE,
					  favor_right_p ? true : false,middlep ? true : false);
  }
  for (i = 0; i < totalpositions; i++) {
    if (links_out) {
      links_out[5*i+0] = links[0][i].fwd_consecutive; links_out[5*i+1] = links[0][i].fwd_rootposition;
      links_out[5*i+2] = links[0][i].fwd_pos; links_out[5*i+3] = links[0][i].fwd_hit; links_out[5*i+4] = links[0][i].fwd_tracei;
    }
    if (scores_out) scores_out[i] = fwd_scores[0][i];
  }
  for (i = 0; i < ncells && i < cells_cap; i++) {
    cells_out[5*i+0] = cells[i]->rootposition; cells_out[5*i+1] = cells[i]->endposition;
    cells_out[5*i+2] = cells[i]->querypos; cells_out[5*i+3] = cells[i]->hit; cells_out[5*i+4] = cells[i]->score;
  }
  if (cells) FREE(cells);
  Linkmatrix_1d_free(&links);
  intmatrix_1d_free(&fwd_scores);
  free(nactive); free(firstactive); free(mappings);
  return ncells;
}

/* path_len[k] = pairs of path k (k in cell-rank order); pairs_out = (querypos, genomepos) per pair, paths
   concatenated, each from its lowest querypos upwards.  Returns npaths (or -needed pairs if pairs_cap is short). */
static int paths_dir (int forwardp, const unsigned int *positions, const int *npositions, int querylength, int totalpositions,
		 const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize,
		 int localp, int skip_repetitive_p, int favor_right_p, int middlep, int max_nalignments,
		 const char *queryseq, const char *queryuc, int *path_len, int maxpaths, int *pairs_out, int pairs_cap) {
  Chrpos_T **mappings = make_mappings(positions,npositions,querylength);
  int *firstactive = (int *) calloc(querylength,sizeof(int)), *nactive = (int *) calloc(querylength,sizeof(int));
  List_T all_paths, p, q;
  int npaths = 0, npairs = 0, k;
  Pair_T pair;

  Pairpool_reset(s2_pairpool);
  Cellpool_reset(s2_cellpool);
  if (forwardp) {
    all_paths = align_compute_lookforward(mappings,(int *) npositions,totalpositions,/*oned_matrix_p*/true,
					  (Chrpos_T *) minactive,(Chrpos_T *) maxactive,firstactive,nactive,s2_cellpool,
					  (char *) queryseq,(char *) queryuc,querylength,querystart,queryend,
					  /*genome*/NULL,/*genomealt*/NULL,/*chroffset*/0,/*chrhigh*/0,/*plusp*/true,indexsize,s2_pairpool,
					  localp ? true : false,skip_repetitive_p ? true : false,
					  /*use_canonical_p*/false,NON_CANONICAL_PENALTY_MIDDLE,
					  favor_right_p ? true : false,middlep ? true : false,max_nalignments);
  } else {
    all_paths = align_compute_lookback(mappings,(int *) npositions,totalpositions,/*oned_matrix_p*/true,
				       (Chrpos_T *) minactive,(Chrpos_T *) maxactive,firstactive,nactive,s2_cellpool,
				       (char *) queryseq,(char *) queryuc,querylength,querystart,queryend,
				       /*genome*/NULL,/*genomealt*/NULL,/*chroffset*/0,/*chrhigh*/0,/*plusp*/true,indexsize,s2_pairpool,
				       localp ? true : false,skip_repetitive_p ? true : false,
				       /*use_canonical_p*/false,NON_CANONICAL_PENALTY_MIDDLE,
				       favor_right_p ? true : false,middlep ? true : false,max_nalignments);
  }
  all_paths = List_reverse(all_paths);		/* pushed in rank order, so the list came out last first */
  for (p = all_paths; p != NULL; p = List_next(p)) {
    k = 0;
    for (q = (List_T) List_head(p); q != NULL; q = List_next(q)) {
      pair = (Pair_T) List_head(q);
      if (npairs < pairs_cap) { pairs_out[2*npairs] = pair->querypos; pairs_out[2*npairs+1] = (int) pair->genomepos; }
      npairs++; k++;
    }
    if (npaths < maxpaths) path_len[npaths] = k;
    npaths++;
  }
  List_free(&all_paths);
  free(nactive); free(firstactive); free(mappings);
  return (npairs > pairs_cap) ? -npairs : npaths;
}

#define SCORES_ARGS const unsigned int *positions, const int *npositions, int querylength, int totalpositions, \
    const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize, \
    int localp, int skip_repetitive_p, int favor_right_p, int middlep, int *links_out, int *scores_out, int *cells_out, int cells_cap
#define SCORES_PASS positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize, \
    localp,skip_repetitive_p,favor_right_p,middlep,links_out,scores_out,cells_out,cells_cap
#define PATHS_ARGS const unsigned int *positions, const int *npositions, int querylength, int totalpositions, \
    const unsigned int *minactive, const unsigned int *maxactive, int querystart, int queryend, int indexsize, \
    int localp, int skip_repetitive_p, int favor_right_p, int middlep, int max_nalignments, \
    const char *queryseq, const char *queryuc, int *path_len, int maxpaths, int *pairs_out, int pairs_cap
#define PATHS_PASS positions,npositions,querylength,totalpositions,minactive,maxactive,querystart,queryend,indexsize, \
    localp,skip_repetitive_p,favor_right_p,middlep,max_nalignments,queryseq,queryuc,path_len,maxpaths,pairs_out,pairs_cap

int refs2_scores (SCORES_ARGS) { return scores_dir(0,SCORES_PASS); }		/* align_compute_scores_lookback */
int refs2_scores_fwd (SCORES_ARGS) { return scores_dir(1,SCORES_PASS); }	/* align_compute_scores_lookforward, stage2.c:4610 */
int refs2_paths (PATHS_ARGS) { return paths_dir(0,PATHS_PASS); }		/* align_compute_lookback */
int refs2_paths_fwd (PATHS_ARGS) { return paths_dir(1,PATHS_PASS); }		/* align_compute_lookforward, stage2.c:5062 */
