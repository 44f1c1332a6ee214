/* Host mirror of align_compute_lookback (stage2.c:4402) over the chaining engine: calls are queued into one
 * device batch, run together, and read back per call in the reference's own order (paths by cell rank, pairs of a
 * path from its lowest querypos upwards, the order of the List_T that traceback_one returns). */
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gmapchain_b200.h"

/* page-locks a pool for the duration of its use when it is large enough for DMA speed to matter */
struct Pinned {
  void *p; size_t bytes;
  Pinned () : p(NULL), bytes(0) {}
  void release () { if (p) gmapdp_host_unregister(p); p = NULL; bytes = 0; }
  void ensure (const void *q, size_t n) {
    if (q == p && n <= bytes) return;
    release();
    if (q && n >= ((size_t) 8 << 20) && gmapdp_host_register((void *) q,n) == GMAPDP_OK) { p = (void *) q; bytes = n; }
  }
};

struct gmapchain_batch {
  gmapdp_ctx *ctx;
  Pinned pin[8];
  std::vector<gmapchain_problem> problems;
  std::vector<int32_t> npos;
  std::vector<uint32_t> cum, mina, maxa, pos;
  std::vector<gmapchain_result> results;
  std::vector<gmapchain_path> paths;
  std::vector<int32_t> pairs;
  size_t paths_used, pairs_used;
  bool have_results;
  std::string err;
};

extern "C" gmapchain_batch *GmapChain_batch_new (gmapdp_ctx *ctx) {
  gmapchain_batch *b = new gmapchain_batch();
  b->ctx = ctx; b->paths_used = b->pairs_used = 0; b->have_results = false;
  return b;
}
static void unpin_all (gmapchain_batch *b) { for (int i = 0; i < 8; i++) b->pin[i].release(); }
extern "C" void GmapChain_batch_free (gmapchain_batch *b) { unpin_all(b); delete b; }
extern "C" void GmapChain_batch_clear (gmapchain_batch *b) {
  b->problems.clear(); b->npos.clear(); b->cum.clear(); b->mina.clear(); b->maxa.clear(); b->pos.clear();
  b->results.clear(); b->paths_used = b->pairs_used = 0; b->have_results = false; b->err.clear();
}

static int queue_call (gmapchain_batch *b, int forwardp, uint32_t *const *mappings, const int *npositions, int totalpositions,
				   const uint32_t *minactive, const uint32_t *maxactive, int querylength, int querystart, int queryend,
				   int indexsize, int localp, int skip_repetitive_p, int use_canonical_p, int non_canonical_penalty,
				   int favor_right_p, int middlep, int max_nalignments) {
  (void) non_canonical_penalty;
  if (use_canonical_p) { b->err = "GmapChain: use_canonical_p is not supported by this version"; return GMAPDP_ERR_ARG; }
  if (querylength < 0 || totalpositions < 0 || querystart < 0 || queryend >= querylength || indexsize <= 0) {
    b->err = "GmapChain_lookback: bad sizes"; return GMAPDP_ERR_ARG;
  }
  for (int i = 0; i < 5; i++) b->pin[i].release();		/* the pools may move while they grow */
  gmapchain_problem p;
  memset(&p,0,sizeof(p));
  p.querylength = querylength; p.querystart = querystart; p.queryend = queryend; p.indexsize = indexsize;
  p.flags = (localp ? GMAPCHAIN_F_LOCALP : 0) | (skip_repetitive_p ? GMAPCHAIN_F_SKIP_REPETITIVE : 0) |
    (favor_right_p ? GMAPCHAIN_F_FAVOR_RIGHT : 0) | (middlep ? GMAPCHAIN_F_MIDDLEP : 0) | (forwardp ? GMAPCHAIN_F_LOOKFORWARD : 0);
  p.max_nalignments = max_nalignments;
  p.q_off = b->npos.size(); p.p_off = b->pos.size();
  uint32_t run = 0;
  for (int q = 0; q < querylength; q++) {
    const int n = npositions[q];
    b->npos.push_back(n); b->cum.push_back(run); b->mina.push_back(minactive[q]); b->maxa.push_back(maxactive[q]);
    if (n > 0) {
      b->pos.insert(b->pos.end(),mappings[q],mappings[q] + n);
      run += (uint32_t) n;
    }
  }
  if ((int) run != totalpositions) {	/* Linkmatrix_1d_new lays the rows out by npositions; totalpositions is their sum */
    b->npos.resize(p.q_off); b->cum.resize(p.q_off); b->mina.resize(p.q_off); b->maxa.resize(p.q_off); b->pos.resize(p.p_off);
    b->err = "GmapChain_lookback: totalpositions does not match npositions"; return GMAPDP_ERR_ARG;
  }
  p.totalpositions = totalpositions;
  b->problems.push_back(p);
  b->have_results = false;
  return (int) b->problems.size() - 1;
}

extern "C" int GmapChain_lookback (gmapchain_batch *b, uint32_t *const *mappings, const int *npositions, int totalpositions,
				   const uint32_t *minactive, const uint32_t *maxactive, int querylength, int querystart, int queryend,
				   int indexsize, int localp, int skip_repetitive_p, int use_canonical_p, int non_canonical_penalty,
				   int favor_right_p, int middlep, int max_nalignments) {
  return queue_call(b,0,mappings,npositions,totalpositions,minactive,maxactive,querylength,querystart,queryend,indexsize,localp,
		    skip_repetitive_p,use_canonical_p,non_canonical_penalty,favor_right_p,middlep,max_nalignments);
}
extern "C" int GmapChain_lookforward (gmapchain_batch *b, uint32_t *const *mappings, const int *npositions, int totalpositions,
				      const uint32_t *minactive, const uint32_t *maxactive, int querylength, int querystart, int queryend,
				      int indexsize, int localp, int skip_repetitive_p, int use_canonical_p, int non_canonical_penalty,
				      int favor_right_p, int middlep, int max_nalignments) {
  return queue_call(b,1,mappings,npositions,totalpositions,minactive,maxactive,querylength,querystart,queryend,indexsize,localp,
		    skip_repetitive_p,use_canonical_p,non_canonical_penalty,favor_right_p,middlep,max_nalignments);
}

static int fail (gmapchain_batch *b, int rc) { b->err = gmapdp_last_error(b->ctx); return rc; }

extern "C" int GmapChain_batch_upload (gmapchain_batch *b) {
  static const int32_t z32 = 0; static const uint32_t zu = 0;
  b->pin[0].ensure(b->npos.data(),b->npos.size() * 4); b->pin[1].ensure(b->cum.data(),b->cum.size() * 4);
  b->pin[2].ensure(b->mina.data(),b->mina.size() * 4); b->pin[3].ensure(b->maxa.data(),b->maxa.size() * 4);
  b->pin[4].ensure(b->pos.data(),b->pos.size() * 4);
  const int rc = gmapchain_upload(b->ctx,b->problems.data(),(int) b->problems.size(),
				  b->npos.empty() ? &z32 : b->npos.data(),b->cum.empty() ? &zu : b->cum.data(),
				  b->mina.empty() ? &zu : b->mina.data(),b->maxa.empty() ? &zu : b->maxa.data(),b->npos.size(),
				  b->pos.empty() ? &zu : b->pos.data(),b->pos.size());
  return rc ? fail(b,rc) : GMAPDP_OK;
}
extern "C" int GmapChain_batch_run_resident (gmapchain_batch *b, float *kernel_ms) {
  const int rc = gmapchain_run_resident(b->ctx,kernel_ms);
  return rc ? fail(b,rc) : GMAPDP_OK;
}
extern "C" int GmapChain_batch_download (gmapchain_batch *b) {
  b->results.resize(b->problems.size() + 1);
  for (int attempt = 0; attempt < 2; attempt++) {
    if (b->paths.size() < 64) b->paths.resize(64);
    if (b->pairs.size() < 128) b->pairs.resize(128);
    b->pin[5].ensure(b->paths.data(),b->paths.size() * sizeof(gmapchain_path));
    b->pin[6].ensure(b->pairs.data(),b->pairs.size() * sizeof(int32_t));
    const int rc = gmapchain_download(b->ctx,b->results.data(),b->paths.data(),b->paths.size(),&b->paths_used,
				      b->pairs.data(),b->pairs.size() / 2,&b->pairs_used);
    if (rc == GMAPDP_OK) { b->have_results = true; return GMAPDP_OK; }
    if (rc != GMAPDP_ERR_CAPACITY) return fail(b,rc);
    b->pin[5].release(); b->pin[6].release();
    b->paths.resize(b->paths_used + 64); b->pairs.resize(2 * b->pairs_used + 128);
  }
  return fail(b,GMAPDP_ERR_CAPACITY);
}
extern "C" int GmapChain_batch_run (gmapchain_batch *b) {
  static const bool trace = getenv("GMAPDP_TRACE") != NULL;	/* host-side timeline on stderr */
  const auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (trace) fprintf(stderr,"GmapChain_batch_run: %-12s %8.2f ms\n",what,
		       std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now() - T0).count());
  };
  int rc;
  if ((rc = GmapChain_batch_upload(b))) return rc;
  lap("uploaded");
  if ((rc = GmapChain_batch_run_resident(b,NULL))) return rc;
  lap("kernel done");
  rc = GmapChain_batch_download(b);
  lap("downloaded");
  return rc;
}

extern "C" int GmapChain_npaths (const gmapchain_batch *b, int id) {
  if (!b->have_results || id < 0 || id >= (int) b->problems.size()) return GMAPDP_ERR_ARG;
  return b->results[id].npaths;
}
extern "C" int GmapChain_path (const gmapchain_batch *b, int id, int k, int *cell, int *querypos, uint32_t *position, int cap) {
  if (!b->have_results || id < 0 || id >= (int) b->problems.size() || k < 0 || k >= b->results[id].npaths) return GMAPDP_ERR_ARG;
  const gmapchain_path &p = b->paths[(size_t) b->results[id].path_off + k];
  if (cell) { cell[0] = p.rootposition; cell[1] = p.endposition; cell[2] = p.querypos; cell[3] = p.hit; cell[4] = p.score; }
  if (p.npairs > cap) return -p.npairs;
  const int32_t *src = b->pairs.data() + 2 * (size_t) p.pair_off;
  for (int t = 0; t < p.npairs; t++) {		/* device order is the traceback's: reverse into list order */
    const int s = p.npairs - 1 - t;
    querypos[t] = src[2 * s]; position[t] = (uint32_t) src[2 * s + 1];
  }
  return p.npairs;
}
extern "C" int GmapChain_batch_ncalls (const gmapchain_batch *b) { return (int) b->problems.size(); }
extern "C" long GmapChain_batch_nhits (const gmapchain_batch *b) { return (long) b->pos.size(); }
extern "C" long GmapChain_batch_h2d_bytes (const gmapchain_batch *b) {
  return (long) (b->problems.size() * (sizeof(gmapchain_problem) + sizeof(int)) + b->npos.size() * 16 + b->pos.size() * 4);
}
extern "C" long GmapChain_batch_d2h_bytes (const gmapchain_batch *b) {
  return (long) (b->problems.size() * sizeof(gmapchain_result) + b->paths_used * sizeof(gmapchain_path) + b->pairs_used * 8);
}
extern "C" unsigned long long GmapChain_batch_digest (const gmapchain_batch *b) {	/* FNV-1a over everything a caller can read */
  unsigned long long h = 1469598103934665603ull;
  auto mix = [&](long long v) { for (int i = 0; i < 8; i++) { h ^= (unsigned long long) ((v >> (8 * i)) & 0xff); h *= 1099511628211ull; } };
  if (!b->have_results) return 0;
  for (size_t i = 0; i < b->problems.size(); i++) {
    const gmapchain_result &r = b->results[i];
    mix(r.npaths); mix(r.bestscore);
    for (int k = 0; k < r.npaths; k++) {
      const gmapchain_path &p = b->paths[(size_t) r.path_off + k];
      mix(p.score); mix(p.rootposition); mix(p.endposition); mix(p.querypos); mix(p.hit); mix(p.npairs);
      const int32_t *src = b->pairs.data() + 2 * (size_t) p.pair_off;
      for (int t = 0; t < 2 * p.npairs; t++) mix(src[t]);
    }
  }
  return h;
}
extern "C" const char *GmapChain_batch_error (const gmapchain_batch *b) { return b->err.c_str(); }
