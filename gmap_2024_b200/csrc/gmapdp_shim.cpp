/* gmapdp_shim.cpp -- host side of the five reference entry points (see include/gmapdp_shim.h).
 *
 * Per call: (1) the reference's argument checks and early exits, (2) its no-fill shortcuts
 * (single_gap_simple dynprog_single.c:345, genome_gap_simple dynprog_genome.c:3005,
 * QUERYEND_NOGAPS dynprog_end.c:577/649), (3) otherwise a device box; after the batch has run,
 * (4) the device's edit script is replayed into the pair list exactly as Dynprog_traceback_*
 * (dynprog_simd.c:9154-9946) and Pairpool_add_queryskip/_genomeskip/_push_gapholder
 * (pairpool.c:981/1068/375) would have built it, and (5) the entry point's own post-processing
 * (leading-indel stripping, nmatches+1<nmismatches, Pair_maxnegscore pair.c:8528, insert pairs
 * dynprog_cdna.c:1008-1041) is applied.
 *
 * This file never computes a DP cell: without a device GmapDP_batch_run fails.
 */
#include "../../include/gmapdp_shim.h"
#include "gmapdp_layout.h"
#include "gmapdp_tables.h"

#include <string>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <atomic>
#include <mutex>
#include <condition_variable>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

enum { COMP_DYNMATCH = '*', COMP_AMBIG = ':', COMP_MISMATCH = ' ', COMP_INDEL = '-', COMP_SHORTGAP = '~' };
const int NEG_INFINITY_32 = -32768;
const int MICROINTRON_LENGTH = 9, LAZY_INDEL = 1, INSERT_PAIRS = 9;

const GdpHostTables &tables () { static GdpHostTables t; return t; }
struct DevTablesHolder { GdpTables d; DevTablesHolder () { tables().device_tables(&d); } };
const GdpTables &dev_tables () { static DevTablesHolder h; return h.d; }	/* C++11 magic static: thread-safe */
double prob_at (const std::vector<double> &v, int k) { return (k >= 0 && (size_t) k < v.size()) ? v[k] : 0.0; }

struct Counts { int score = 0, nmatches = 0, nmismatches = 0, nopens = 0, nindels = 0; };

/* A pair list under construction and the finished list of a call: compact records (gmapdp_cpair, 16 bytes) in a plain
   array that is never value-initialised and keeps its memory when cleared, plus the side table of its gap holders.
   The reference conses, so a list's head is the element pushed last; the entry points hand out some lists (or parts:
   the right-hand side of a genome / cdna gap, a 5' end) in reversed order.  Those parts are written BACKWARDS from a
   middle mark instead of being reversed afterwards: every pass over the ~10^9 pairs of a large batch costs as much as
   writing them (the records are written with non-temporal stores for the same reason: no read-for-ownership). */
struct Pushed {
  gmapdp_cpair *buf = NULL;
  size_t lo = 0, hi = 0, cap = 0;		/* elements buf[lo .. hi) */
  std::vector<gmapdp_gapinfo> gaps;
  /* the records hold positions relative to the side they lie on (include/gmapdp_shim.h): the bases, and the side the
     pairs pushed next belong to */
  gmapdp_pairbase base = {{0,0},{0,0}};
  int side = 0;
  void use_side (int k, int qbase, int gbase) { side = k; base.q[k] = qbase; base.g[k] = gbase; }
  Pushed () {}
  Pushed (const Pushed &) = delete;
  Pushed &operator= (const Pushed &) = delete;
  Pushed (Pushed &&o) noexcept : buf(o.buf), lo(o.lo), hi(o.hi), cap(o.cap), gaps(std::move(o.gaps)), base(o.base), side(o.side) { o.buf = NULL; o.lo = o.hi = o.cap = 0; }
  Pushed &operator= (Pushed &&o) noexcept {
    if (this != &o) { free(buf); buf = o.buf; lo = o.lo; hi = o.hi; cap = o.cap; gaps = std::move(o.gaps); base = o.base; side = o.side; o.buf = NULL; o.lo = o.hi = o.cap = 0; }
    return *this;
  }
  ~Pushed () { free(buf); }
  /* empties the list; room for `back' elements pushed backwards and `fwd' pushed forwards */
  void start (size_t back, size_t fwd) {
    if (back + fwd > cap) {
      free(buf);
      buf = NULL;
      if (posix_memalign((void **) &buf,64,(back + fwd) * sizeof(gmapdp_cpair)) != 0) abort();
      cap = back + fwd;
    }
    lo = hi = back;
    gaps.clear();
  }
  void clear () { lo = hi = 0; gaps.clear(); }
  size_t size () const { return hi - lo; }
  bool empty () const { return hi == lo; }
  const gmapdp_cpair *data () const { return buf + lo; }
  const gmapdp_cpair &operator[] (size_t i) const { return buf[lo + i]; }
  void grow_fwd () {		/* lists of unknown length (host-resolved calls): amortised doubling */
    const size_t ncap = cap ? 2 * cap : 64;
    gmapdp_cpair *nb = NULL;
    if (posix_memalign((void **) &nb,64,ncap * sizeof(gmapdp_cpair)) != 0) abort();
    if (hi > lo) memcpy(nb + lo,buf + lo,(hi - lo) * sizeof(gmapdp_cpair));
    free(buf); buf = nb; cap = ncap;
  }
  gmapdp_cpair *slot (bool back) {
    if (back) { if (lo == 0) abort(); return &buf[--lo]; }
    if (hi == cap) grow_fwd();
    return &buf[hi++];
  }
  void drop_front (size_t k) { lo += k; }
  void drop_back (size_t k) { hi -= k; }
  void reverse () { std::reverse(buf + lo,buf + hi); }
};

#if defined(__SSE2__) && !defined(GMAPDP_NO_NT_STORES)
/* software prefetch distance of the replay's five streams, in bytes of input (0: none, the default).  It is host
   dependent: on the development container 256 took the eight-thread replay of 100 k end gaps from 141 to 20 ms, on the
   GPU box's host it made the end-to-end step slower (431-454 vs 399 ms). */
#ifndef GMAPDP_REPLAY_PREFETCH
#define GMAPDP_REPLAY_PREFETCH 0
#endif

/* ordinary or non-temporal stores of the records (GMAPDP_NT_STORES=1 selects non-temporal ones).  On the GPU box's host the
   two measure the same for a large batch (85.2 vs 84.8 ms end to end on 200 k benchmark boxes); ordinary stores leave a
   small list in the cache for whoever reads it next -- the entry point's own look at its list, and in the drop-in the
   binding that conses it into the caller's pool right away. */
const bool g_nt_stores = (getenv("GMAPDP_NT_STORES") && atoi(getenv("GMAPDP_NT_STORES")) != 0);
/* two records (16 bytes) at p; p is 8-byte aligned */
inline void store2 (gmapdp_cpair *p, __m128i v) {
  if (!g_nt_stores) _mm_storeu_si128(reinterpret_cast<__m128i *>(p),v);
  else if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) _mm_stream_si128(reinterpret_cast<__m128i *>(p),v);
  else {
    _mm_stream_si64(reinterpret_cast<long long *>(p),_mm_cvtsi128_si64(v));
    _mm_stream_si64(reinterpret_cast<long long *>(p + 1),_mm_cvtsi128_si64(_mm_unpackhi_epi64(v,v)));
  }
}
#endif

/* one record: positions relative to the list's current side, the side in the top bit of comp */
inline void store_pair (Pushed &l, gmapdp_cpair *p, int querypos, int genomepos, char cdna, char comp, char genome, char genomealt) {
  const unsigned qrel = (unsigned) (querypos - l.base.q[l.side]), grel = (unsigned) (genomepos - l.base.g[l.side]);
  if (qrel >= GMAPDP_CPAIR_GAP || grel > 0xffffu) abort();	/* a position outside its side: never with sides <= 32767 */
  const uint64_t rec = (uint64_t) qrel | ((uint64_t) grel << 16) | ((uint64_t) (uint8_t) cdna << 32)
    | ((uint64_t) (uint8_t) (comp | (l.side ? 0x80 : 0)) << 40) | ((uint64_t) (uint8_t) genome << 48) | ((uint64_t) (uint8_t) genomealt << 56);
#if defined(__SSE2__) && !defined(GMAPDP_NO_NT_STORES)
  if (g_nt_stores) { _mm_stream_si64(reinterpret_cast<long long *>(p),(long long) rec); return; }
#endif
  memcpy(p,&rec,sizeof(rec));
}

inline void push_pair (Pushed &l, bool back, int querypos, int genomepos, char cdna, char comp, char genome, char genomealt) {
  if (querypos < 0 || genomepos < 0) return;		/* Pairpool_push, pairpool.c:190 */
  store_pair(l,l.slot(back),querypos,genomepos,cdna,comp,genome,genomealt);
}

inline gmapdp_gapinfo &push_gapholder (Pushed &l, bool back, int queryjump, int genomejump) {
  gmapdp_gapinfo g;
  g.queryjump = queryjump; g.genomejump = genomejump; g.introntype = 0; g.pad_ = 0; g.donor_prob = 0.0; g.acceptor_prob = 0.0;
  l.gaps.push_back(g);
  if (l.gaps.size() > 0xffffu) abort();
  const uint64_t rec = (uint64_t) GMAPDP_CPAIR_GAP | ((uint64_t) (l.gaps.size() - 1) << 16) | 0x2020202000000000ull;
  gmapdp_cpair *p = l.slot(back);
#if defined(__SSE2__) && !defined(GMAPDP_NO_NT_STORES)
  if (g_nt_stores) { _mm_stream_si64(reinterpret_cast<long long *>(p),(long long) rec); return l.gaps.back(); }
#endif
  memcpy(p,&rec,sizeof(rec));
  return l.gaps.back();
}

inline bool cpair_is_gap (const gmapdp_cpair &p) { return p.qrel == GMAPDP_CPAIR_GAP; }
inline char cpair_comp (const gmapdp_cpair &p) { return (char) (p.comp & 0x7f); }

/* one traced side, with the reference's pointer conventions */
struct Side {
  const char *rseq, *rsequc, *gseq, *galt;
  int queryoffset, genomeoffset;
  bool revp;
  bool plain = false;		/* every character of rsequc, gseq and galt is one of A C G T: a mismatch cannot be an ambiguity match,
				   and there is no '*' (Call::plain_p) */
};

struct Replayer {
  Pushed &l; const Side &sd; Counts &n; const GdpHostTables &t; const bool back;
  /* back: the list wanted is the reverse of the push order -- written backwards */
  /* side: which of the call's (up to two) sides sd is; the records of a reversed side are relative to 32767 below its offsets */
  Replayer (Pushed &l_, const Side &sd_, Counts &n_, bool back_ = false, int side = 0) : l(l_), sd(sd_), n(n_), t(tables()), back(back_) {
    l.use_side(side,sd.queryoffset - (sd.revp ? 32767 : 0),sd.genomeoffset - (sd.revp ? 32767 : 0));
  }

  void diag (int r, int c) {
    int qc = r - 1, gc = c - 1;
    if (sd.revp) { qc = -qc; gc = -gc; }
    const char c1 = sd.rseq[qc], c1uc = sd.rsequc[qc], c2 = sd.gseq[gc], c2a = sd.galt[gc];
    if (c2 == '*') return;
    if (c1uc == c2 || c1uc == c2a) {
      n.score += 1; n.nmatches++;
      push_pair(l,back,sd.queryoffset + qc,sd.genomeoffset + gc,c1,COMP_DYNMATCH,c2,c2a);
    } else if (t.cons[c1uc & 127][c2 & 127] || t.cons[c1uc & 127][c2a & 127]) {
      n.score += 1; n.nmatches++;
      push_pair(l,back,sd.queryoffset + qc,sd.genomeoffset + gc,c1,COMP_AMBIG,c2,c2a);
    } else {
      n.score += -3; n.nmismatches++;
      push_pair(l,back,sd.queryoffset + qc,sd.genomeoffset + gc,c1,COMP_MISMATCH,c2,c2a);
    }
  }
  void hgap (int r, int c, int dist) {		/* Pairpool_add_genomeskip */
    int qc = r - 1, left = c - dist, right = c - 1, step = -1;
    if (sd.revp) { qc = -qc; int t = left; left = -right; right = -t; step = +1; }
    if (dist < MICROINTRON_LENGTH) {
      int gc = sd.revp ? left : right;
      for (int j = 0; j < dist; j++, gc += step)
	push_pair(l,back,sd.queryoffset + qc,sd.genomeoffset + gc,' ',COMP_INDEL,sd.gseq[gc],sd.galt[gc]);
      n.score += -3 - dist; n.nopens++; n.nindels += dist;
    } else {
      push_gapholder(l,back,0,dist);
    }
  }
  void vgap (int r, int c, int dist) {		/* Pairpool_add_queryskip */
    int qc = r - 1, gc = c - 1, step = -1;
    if (sd.revp) { qc = -qc; gc = -gc; step = +1; }
    for (int j = 0; j < dist; j++, qc += step)
      push_pair(l,back,sd.queryoffset + qc,sd.genomeoffset + gc,sd.rseq[qc],COMP_INDEL,' ',' ');
    n.score += -3 - dist; n.nopens++; n.nindels += dist;
  }
  /* `len' diagonal steps from (r,c): diag() in one loop over four pointers.  Only for lists whose room was reserved by
     Pushed::start (every device call: a side pushes at most rlength + glength pairs). */
  void diag_run (int r, int c, int len) {
    int qc = r - 1, gc = c - 1, step = -1;
    if (sd.revp) { qc = -qc; gc = -gc; step = +1; }
    const char *rs = sd.rseq + qc, *ru = sd.rsequc + qc, *gs = sd.gseq + gc, *ga = sd.galt + gc;
    int qpos = sd.queryoffset + qc, gpos = sd.genomeoffset + gc;
    gmapdp_cpair *out = l.buf + (back ? l.lo : l.hi);
    const int ostep = back ? -1 : +1;
    if (back) out--;
    int nm = 0, nx = 0;
    int j = 0;
#if defined(__SSE2__) && !defined(GMAPDP_NO_NT_STORES)
    /* Eight pairs at a time.  Whichever way a side is walked -- forwards in memory and written backwards (rev sides), or
       backwards in memory and written forwards -- the record of the input byte at ascending memory index m goes to the
       descending output slot base - m, with querypos q0 + m and genomepos g0 + m.  Chunks with a '*' (chromosome edge)
       or a negative position take the scalar loop below. */
    if ((step > 0) == back) {
      const __m128i star = _mm_set1_epi8('*'), dyn = _mm_set1_epi8((char) COMP_DYNMATCH), mis = _mm_set1_epi8((char) COMP_MISMATCH);
      const __m128i sidebit = _mm_set1_epi8(l.side ? (char) 0x80 : (char) 0);
      const __m128i iota16 = _mm_set_epi16(7,6,5,4,3,2,1,0);
      const int qb = l.base.q[l.side], gb = l.base.g[l.side];
      {
	/* Sixteen pairs at a time: comp is a blend of the compare mask.  A call whose sequences are plain A C G T
	   (Call::plain_p) needs nothing else; otherwise the block leaves a chromosome edge ('*') to the careful loops below
	   and looks the FEW mismatching bytes up in the consistency table (an ambiguity match needs an IUPAC code).  A run of
	   16 or more ends with one more block laid over its last sixteen pairs (the records it rewrites are the same; only
	   the new lanes are counted) instead of up to fifteen scalar steps. */
	const bool plain = sd.plain;
	const __m128i mis_s = _mm_or_si128(mis,sidebit), dynxmis = _mm_xor_si128(dyn,mis), eight = _mm_set1_epi16(8);
	for (;;) {
	  unsigned lanes = 0xffffu;			/* memory lanes of this block that are new */
	  int back16 = 0;				/* pairs of the run this block steps back over */
	  if (len - j < 16) {
	    if (j == 0 || j == len) break;		/* a run shorter than 16, or done */
	    back16 = 16 - (len - j);
	    lanes = (step > 0) ? (0xffffu << back16) & 0xffffu : 0xffffu >> back16;
	  }
	  const char *rs_ = rs - back16 * step, *ru_ = ru - back16 * step, *gs_ = gs - back16 * step, *ga_ = ga - back16 * step;
	  const int lowoff = (step > 0) ? 0 : -15;
	  const int q0 = qpos - back16 * step + lowoff, g0 = gpos - back16 * step + lowoff;
	  if ((q0 | g0) < 0) break;
	  const __m128i vru = _mm_loadu_si128(reinterpret_cast<const __m128i *>(ru_ + lowoff));
	  const __m128i vgs = _mm_loadu_si128(reinterpret_cast<const __m128i *>(gs_ + lowoff));
	  const __m128i vga = _mm_loadu_si128(reinterpret_cast<const __m128i *>(ga_ + lowoff));
	  const __m128i vrs = _mm_loadu_si128(reinterpret_cast<const __m128i *>(rs_ + lowoff));
	  if (!plain && _mm_movemask_epi8(_mm_cmpeq_epi8(vgs,star)) != 0) break;
	  const __m128i eq = _mm_or_si128(_mm_cmpeq_epi8(vru,vgs),_mm_cmpeq_epi8(vru,vga));
	  __m128i comp = _mm_xor_si128(mis_s,_mm_and_si128(eq,dynxmis));
	  unsigned okbits = (unsigned) _mm_movemask_epi8(eq);
	  if (!plain && okbits != 0xffffu) {
	    unsigned amb = 0;
	    for (unsigned w = ~okbits & 0xffffu; w; w &= w - 1) {
	      const int m = __builtin_ctz(w), a = ru_[lowoff + m] & 127;
	      if (t.cons[a][gs_[lowoff + m] & 127] || t.cons[a][ga_[lowoff + m] & 127]) amb |= 1u << m;
	    }
	    if (amb) {
	      alignas(16) char cb[16];
	      _mm_store_si128(reinterpret_cast<__m128i *>(cb),comp);
	      for (unsigned w = amb; w; w &= w - 1) cb[__builtin_ctz(w)] = (char) (COMP_AMBIG | (l.side ? 0x80 : 0));
	      comp = _mm_load_si128(reinterpret_cast<const __m128i *>(cb));
	      okbits |= amb;
	    }
	  }
	  const int nnew = 16 - back16;
	  const int good = __builtin_popcount(okbits & lanes);
	  nm += good; nx += nnew - good;
	  const __m128i rcl = _mm_unpacklo_epi8(vrs,comp), rch = _mm_unpackhi_epi8(vrs,comp);
	  const __m128i ggl = _mm_unpacklo_epi8(vgs,vga), ggh = _mm_unpackhi_epi8(vgs,vga);
	  const __m128i vq = _mm_add_epi16(_mm_set1_epi16((short) (q0 - qb)),iota16), vg = _mm_add_epi16(_mm_set1_epi16((short) (g0 - gb)),iota16);
	  const __m128i vq8 = _mm_add_epi16(vq,eight), vg8 = _mm_add_epi16(vg,eight);
	  const __m128i c0 = _mm_unpacklo_epi16(rcl,ggl), c1 = _mm_unpackhi_epi16(rcl,ggl), c2 = _mm_unpacklo_epi16(rch,ggh), c3 = _mm_unpackhi_epi16(rch,ggh);
	  const __m128i p0 = _mm_unpacklo_epi16(vq,vg), p1 = _mm_unpackhi_epi16(vq,vg), p2 = _mm_unpacklo_epi16(vq8,vg8), p3 = _mm_unpackhi_epi16(vq8,vg8);
	  gmapdp_cpair *o_ = out - back16 * ostep;
	  gmapdp_cpair *base = (step > 0) ? o_ : o_ + 15;	/* slot of memory index 0; index m goes to base - m */
	  store2(base - 1,_mm_shuffle_epi32(_mm_unpacklo_epi32(p0,c0),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 3,_mm_shuffle_epi32(_mm_unpackhi_epi32(p0,c0),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 5,_mm_shuffle_epi32(_mm_unpacklo_epi32(p1,c1),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 7,_mm_shuffle_epi32(_mm_unpackhi_epi32(p1,c1),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 9,_mm_shuffle_epi32(_mm_unpacklo_epi32(p2,c2),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 11,_mm_shuffle_epi32(_mm_unpackhi_epi32(p2,c2),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 13,_mm_shuffle_epi32(_mm_unpacklo_epi32(p3,c3),_MM_SHUFFLE(1,0,3,2)));
	  store2(base - 15,_mm_shuffle_epi32(_mm_unpackhi_epi32(p3,c3),_MM_SHUFFLE(1,0,3,2)));
	  j += nnew; rs += nnew * step; ru += nnew * step; gs += nnew * step; ga += nnew * step; qpos += nnew * step; gpos += nnew * step; out += nnew * ostep;
	}
      }
      for (; len - j >= 8; j += 8, rs += 8 * step, ru += 8 * step, gs += 8 * step, ga += 8 * step, qpos += 8 * step, gpos += 8 * step, out += 8 * ostep) {
	const int lowoff = (step > 0) ? 0 : -7;			/* ascending-memory start of the chunk */
	const int q0 = qpos + lowoff, g0 = gpos + lowoff;
	if ((q0 | g0) < 0) break;
#if GMAPDP_REPLAY_PREFETCH
	__builtin_prefetch(ru + step * GMAPDP_REPLAY_PREFETCH,0,0); __builtin_prefetch(gs + step * GMAPDP_REPLAY_PREFETCH,0,0);
	__builtin_prefetch(ga + step * GMAPDP_REPLAY_PREFETCH,0,0); __builtin_prefetch(rs + step * GMAPDP_REPLAY_PREFETCH,0,0);
	__builtin_prefetch(out + ostep * (GMAPDP_REPLAY_PREFETCH / 4),1,0);
#endif
	const __m128i vru = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(ru + lowoff));
	const __m128i vgs = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(gs + lowoff));
	const __m128i vga = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(ga + lowoff));
	const __m128i vrs = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(rs + lowoff));
	if (_mm_movemask_epi8(_mm_cmpeq_epi8(vgs,star)) & 0xff) break;
	const __m128i eq = _mm_or_si128(_mm_cmpeq_epi8(vru,vgs),_mm_cmpeq_epi8(vru,vga));
	__m128i comp = _mm_or_si128(_mm_and_si128(eq,dyn),_mm_andnot_si128(eq,mis));
	const int neq = (~_mm_movemask_epi8(eq)) & 0xff;
	int good = 8;
	if (neq) {						/* the bytes that differ: ambiguous (consistent) or mismatch */
	  alignas(16) char cb[16];
	  _mm_store_si128(reinterpret_cast<__m128i *>(cb),comp);
	  for (int m = 0; m < 8; m++) if ((neq >> m) & 1) {
	    const int a = ru[lowoff + m] & 127;
	    if (t.cons[a][gs[lowoff + m] & 127] || t.cons[a][ga[lowoff + m] & 127]) cb[m] = (char) COMP_AMBIG; else good--;
	  }
	  comp = _mm_load_si128(reinterpret_cast<const __m128i *>(cb));
	}
	nm += good; nx += 8 - good;
	comp = _mm_or_si128(comp,sidebit);
	/* chars words (cdna | comp << 8 | genome << 16 | genomealt << 24) of bytes 0-3 and 4-7 */
	const __m128i rc = _mm_unpacklo_epi8(vrs,comp), gg = _mm_unpacklo_epi8(vgs,vga);
	const __m128i c03 = _mm_unpacklo_epi16(rc,gg), c47 = _mm_unpackhi_epi16(rc,gg);
	/* position words (qrel | grel << 16) of bytes 0-3 and 4-7 */
	const __m128i vq = _mm_add_epi16(_mm_set1_epi16((short) (q0 - qb)),iota16), vg = _mm_add_epi16(_mm_set1_epi16((short) (g0 - gb)),iota16);
	const __m128i p03 = _mm_unpacklo_epi16(vq,vg), p47 = _mm_unpackhi_epi16(vq,vg);
	/* the record of memory index m goes to slot base - m: two records per 16-byte store, the higher index first */
	gmapdp_cpair *base = (step > 0) ? out : out + 7;	/* slot of memory index 0 */
	store2(base - 1,_mm_shuffle_epi32(_mm_unpacklo_epi32(p03,c03),_MM_SHUFFLE(1,0,3,2)));
	store2(base - 3,_mm_shuffle_epi32(_mm_unpackhi_epi32(p03,c03),_MM_SHUFFLE(1,0,3,2)));
	store2(base - 5,_mm_shuffle_epi32(_mm_unpacklo_epi32(p47,c47),_MM_SHUFFLE(1,0,3,2)));
	store2(base - 7,_mm_shuffle_epi32(_mm_unpackhi_epi32(p47,c47),_MM_SHUFFLE(1,0,3,2)));
      }
    }
#endif
    for (; j < len; j++, rs += step, ru += step, gs += step, ga += step, qpos += step, gpos += step) {
      const char c2 = *gs;
      if (c2 == '*') continue;
      const char c1uc = *ru, c2a = *ga;
      char comp;
      if (c1uc == c2 || c1uc == c2a) { comp = COMP_DYNMATCH; nm++; }
      else if (t.cons[c1uc & 127][c2 & 127] || t.cons[c1uc & 127][c2a & 127]) { comp = COMP_AMBIG; nm++; }
      else { comp = COMP_MISMATCH; nx++; }
      if (qpos < 0 || gpos < 0) continue;			/* Pairpool_push, pairpool.c:190 */
      store_pair(l,out,qpos,gpos,*rs,comp,c2,c2a);
      out += ostep;
    }
    if (back) l.lo = (size_t) (out + 1 - l.buf); else l.hi = (size_t) (out - l.buf);
    n.score += nm - 3 * nx; n.nmatches += nm; n.nmismatches += nx;
  }
  /* kind: 0 full, 1 upper, 2 lower */
  void run (int kind, int r, int c, const uint32_t *ops, int nops) {
    for (int k = 0; k < nops; k++) {
      const int len = (int) (ops[k] >> 2), what = (int) (ops[k] & 3u);
      if (what == 0) { diag_run(r,c,len); r -= len; c -= len; }
      else if (what == 1) { hgap(r,c,len); c -= len; }
      else { vgap(r,c,len); r -= len; }
    }
    if (kind == 0) {
      if (r == 0 && c == 0) { }
      else if (c == 0) vgap(r,0 + LAZY_INDEL,r);
      else hgap(0 + LAZY_INDEL,c,c);
    } else if (kind == 1) {
      if (c != 0) hgap(0 + LAZY_INDEL,c,c);
    } else {
      if (r != 0) vgap(r,0 + LAZY_INDEL,r);
    }
  }
};

int maxnegscore_headfirst (const Pushed &l) {	/* Pair_maxnegscore, pair.c:8528; the list's head is its last pushed element */
  int maxneg = 0, prevhigh = 0, score = 0;
  const size_t n = l.size();
  size_t i = 0;
#define HF(I) l[n - 1 - (I)]
  while (i < n) {
    if (cpair_is_gap(HF(i))) i++;
    else if (cpair_comp(HF(i)) == COMP_MISMATCH) { score += -3; maxneg = std::min(maxneg,score - prevhigh); i++; }
    else if (cpair_comp(HF(i)) == COMP_INDEL) {
      score += -3 + -1; i++;
      while (i < n && !cpair_is_gap(HF(i)) && cpair_comp(HF(i)) == COMP_INDEL) { score += -1; i++; }
      maxneg = std::min(maxneg,score - prevhigh);
    } else { score += 1; prevhigh = std::max(prevhigh,score); i++; }
  }
#undef HF
  return maxneg;
}

int left_dinucl (char a, char aa, char b, char ba) {
  if ((a == 'G' || aa == 'G') && (b == 'T' || ba == 'T')) return 0x21;
  if ((a == 'G' || aa == 'G') && (b == 'C' || ba == 'C')) return 0x10;
  if ((a == 'A' || aa == 'A') && (b == 'T' || ba == 'T')) return 0x08;
  if ((a == 'C' || aa == 'C') && (b == 'T' || ba == 'T')) return 0x06;
  return 0;
}
int right_dinucl (char r2, char r2a, char r1, char r1a) {
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'G' || r1a == 'G')) return 0x30;
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'C' || r1a == 'C')) return 0x0C;
  if ((r2 == 'G' || r2a == 'G') && (r1 == 'C' || r1a == 'C')) return 0x02;
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'T' || r1a == 'T')) return 0x01;
  return 0;
}

void bump (int &dynprogindex) { dynprogindex += (dynprogindex > 0 ? +1 : -1); }
int quality (double defect_rate) { return defect_rate < 0.003 ? 0 : (defect_rate < 0.014 ? 1 : 2); }	/* dynprog.h:57-58 */

void compute_bands (int &lband, int &uband, int rlength, int glength, int extraband, bool widebandp) {	/* dynprog.c:1246 */
  if (!widebandp) { lband = extraband; uband = extraband; }
  else if (glength >= rlength) { uband = glength - rlength + extraband; lband = extraband; }
  else { lband = rlength - glength + extraband; uband = extraband; }
}

struct Call {
  int mode = 0;
  bool done = false;		/* resolved on the host */
  int box = -1;			/* index into the device batch */
  bool isnull = true;
  int iout[10]; double dout[2];
  Pushed pairs;			/* head first */
  int dpi = 0;			/* dynprogindex of the call's ordinary pairs */
  bool devp = false; gmapdp_coords co;	/* genome gap whose probabilities come from MaxEnt over the resident genome */
  const gmapdp_batch *owner = NULL;
  /* copies of the caller's sequences: forward arrays */
  std::string q, quc, qR, qRuc, gL, gLa, gR, gRa;
  std::vector<double> lp, rp;
  int rlenL = 0, rlenR = 0, glenL = 0, glenR = 0;
  int roffset = 0, goffset = 0, roffsetR = 0, goffsetR = 0;
  int endalign = 0, introntype_in = 0;
  bool require_pos = false, end5 = false;
  int plain = -1;		/* 1: the upper-cased query and the genomic segments consist of A C G T only (plain_p); -1: not looked at yet */
  bool ucL = false, ucR = false, altL = false, altR = false;	/* with plain: the query is its own upper case / the alt segment equals the segment
								   (the replay then reads two character streams per side instead of four) */
  int iout_q[10]; double dout_q[2];	/* the out-parameters as they stood when the box was queued (GmapDP_batch_rewind) */
  void queued () { for (int i = 0; i < 10; i++) iout_q[i] = iout[i]; dout_q[0] = dout[0]; dout_q[1] = dout[1]; }
  Call () { for (int i = 0; i < 10; i++) iout[i] = 0; dout[0] = dout[1] = 0.0; }
};

}  // namespace

struct gmapdp_batch {
  gmapdp_ctx *ctx;
  int max_rlength, max_glength;
  std::vector<Call> calls;
  std::vector<gmapdp_box> boxes;
  std::vector<int> box_call;
  std::vector<uint8_t> seqpool;
  std::vector<double> probpool;
  /* resident genome (GmapDP_batch_genome) and the coordinates of the next queued call (GmapDP_batch_next_coords) */
  const uint32_t *g_blocks = NULL; size_t g_nwords = 0; const gmapdp_maxent_tables *g_tables = NULL;
  gmapdp_coords next_co, cur_co; bool has_next_co = false, has_cur_co = false;
  size_t devprob_need = 0;			/* doubles of device-evaluated probability arrays the queued boxes refer to */
  gmapdp_result *results = NULL; size_t nresults = 0;	/* pinned */
  uint32_t *script = NULL; size_t script_cap = 0;	/* pinned */
  size_t script_used = 0;
  long cells = 0, cells8 = 0, cells_full = 0, cells8_full = 0;	/* _full: the single-gap (full fill) share */
  bool pinned = false, bufs_pinned = false;
  int count_mismatch = 0;
  std::string err;
  bool uploaded = false, overflow = false;
  /* --indel-open / --indel-extend (user_dynprog_p: dynprog_single.c:470, dynprog_genome.c:3367, dynprog_end.c:1334,1964) */
  bool user_dynprog = false; int user_open = 0, user_extend = 0;
  /* registered pool extents (the vectors may move if a call is queued after pinning) */
  void *pin_boxes = NULL, *pin_seq = NULL, *pin_probs = NULL;

  uint32_t add_bytes (const char *p, int n) {
    if (seqpool.size() + (size_t) n + 4 > 0xFFFFFFF0ull) { err = "sequence pool exceeds 4 GiB: split the batch"; overflow = true; return 0; }
    uint32_t off = (uint32_t) seqpool.size();
    seqpool.insert(seqpool.end(),(const uint8_t *) p,(const uint8_t *) p + n);
    while (seqpool.size() & 3) seqpool.push_back(0);
    return off;
  }
};

extern "C" void GmapDP_maxlengths (int *max_rlength, int *max_glength, int maxlookback, int extraquerygap, int maxpeelback,
				   int extramaterial_end, int extramaterial_paired) {
  *max_rlength = std::max(maxlookback + maxpeelback,500);
  *max_glength = std::max(*max_rlength + extraquerygap + std::max(extramaterial_end,extramaterial_paired),2000);
}

/* the kernels pack rows and columns into 16-bit fields (best cells, bridge keys): one box side is at most 32767 long */
#define GDP_MAX_SIDE 32767

extern "C" gmapdp_batch *GmapDP_batch_new (gmapdp_ctx *ctx, int max_rlength, int max_glength) {
  if (max_rlength < 0 || max_glength < 0 || max_rlength > GDP_MAX_SIDE || max_glength > GDP_MAX_SIDE) return NULL;
  gmapdp_batch *b = new gmapdp_batch();
  b->ctx = ctx; b->max_rlength = max_rlength; b->max_glength = max_glength;
  return b;
}
extern "C" int GmapDP_batch_user_dynprog (gmapdp_batch *b, int user_open, int user_extend, int user_dynprog_p) {
  /* gmap.c:5425-5445 clamps both to [-127, 0] */
  if (user_dynprog_p && (user_open > 0 || user_extend > 0 || user_open < -127 || user_extend < -127)) {
    b->err = "user penalties must be in [-127, 0]"; return GMAPDP_ERR_ARG;
  }
  b->user_dynprog = user_dynprog_p != 0; b->user_open = user_open; b->user_extend = user_extend;
  return GMAPDP_OK;
}
static void host_free (gmapdp_batch *b, void *p) { if (b->bufs_pinned) gmapdp_host_free(p); else free(p); }
static void unpin (gmapdp_batch *b, bool release_buffers) {
  if (b->pinned) {
    gmapdp_host_unregister(b->pin_boxes); gmapdp_host_unregister(b->pin_seq); gmapdp_host_unregister(b->pin_probs);
    b->pin_boxes = b->pin_seq = b->pin_probs = NULL;
    b->pinned = false;
  }
  if (release_buffers) {
    host_free(b,b->results); b->results = NULL; b->nresults = 0;
    host_free(b,b->script); b->script = NULL; b->script_cap = 0;
  }
}
/* a call queued after the pools were page-locked may move them: drop the registration first */
static inline void begin_queue (gmapdp_batch *b) {
  if (b->pinned) unpin(b,false);
  b->uploaded = false;
  /* coordinates announced by GmapDP_batch_next_coords belong to this call, whatever becomes of it */
  b->has_cur_co = b->has_next_co; b->cur_co = b->next_co; b->has_next_co = false;
}

/* Resident-genome form of a box: the genomic segment(s) as coordinates (and, for genome gaps, the MaxEnt probabilities
   as coordinates of their first entries) instead of bytes in the pools.  Only when the call came with coordinates and
   has no alternate genome; returns false otherwise (the caller then uploads the characters as usual). */
static bool take_coords (gmapdp_batch *b, Call &c, gmapdp_box &x, bool twosegments) {
  if (!b->has_cur_co || !b->g_blocks) return false;
  if (c.gLa != c.gL || (twosegments && c.gRa != c.gR)) return false;
  const gmapdp_coords &co = b->cur_co;
  x.chroffset = co.chroffset; x.chrhigh = co.chrhigh;
  x.gL_off = x.gLalt_off = co.gposL;
  x.gflags |= GMAPDP_G_SEG_L | (co.negL ? GMAPDP_G_NEG_L : 0) | (co.leftL ? GMAPDP_G_LEFT_L : 0);
  if (twosegments) {
    x.gR_off = x.gRalt_off = co.gposR;
    x.gflags |= GMAPDP_G_SEG_R | (co.negR ? GMAPDP_G_NEG_R : 0) | (co.leftR ? GMAPDP_G_LEFT_R : 0);
  }
  if (c.devp) {
    x.gflags |= GMAPDP_G_PROBS | (co.probnegL ? GMAPDP_G_PSTEP_NEG_L : 0) | (co.probnegR ? GMAPDP_G_PSTEP_NEG_R : 0);
    x.probposL = co.probposL; x.probposR = co.probposR;
    x.probkindL = (uint8_t) co.probkindL; x.probkindR = (uint8_t) co.probkindR;
    /* room for the two arrays in the context's device-side pool (nothing is uploaded for them) */
    x.probL_off = (uint32_t) b->devprob_need; b->devprob_need += (size_t) x.glenL + 1;
    x.probR_off = (uint32_t) b->devprob_need; b->devprob_need += (size_t) x.glenR + 1;
    if (b->devprob_need > 0xffffffffull) b->overflow = true;
  }
  return true;
}

/* entry k of a genome gap's left / right probability array (0.0 beyond glength - 2, as in the calloc'ed arrays of
   dynprog_genome.c:970-1061) */
static double call_prob (const Call &c, bool right, int k) {
  if (!c.devp) return prob_at(right ? c.rp : c.lp,k);
  const int n = (right ? c.glenR : c.glenL) - 1;
  if (k < 0 || k >= n) return 0.0;
  const gmapdp_coords &co = c.co;
  const uint32_t pos = right ? (co.probnegR ? co.probposR - (uint32_t) k : co.probposR + (uint32_t) k)
			     : (co.probnegL ? co.probposL - (uint32_t) k : co.probposL + (uint32_t) k);
  return gmapdp_maxent_host_prob(c.owner->g_blocks,c.owner->g_nwords,c.owner->g_tables,right ? co.probkindR : co.probkindL,pos,co.chroffset);
}

extern "C" int GmapDP_batch_genome (gmapdp_batch *b, const uint32_t *blocks, size_t nwords, const gmapdp_maxent_tables *tables) {
  b->g_blocks = blocks; b->g_nwords = nwords; b->g_tables = tables;
  return GMAPDP_OK;
}
extern "C" void GmapDP_batch_next_coords (gmapdp_batch *b, const gmapdp_coords *co) {
  if (co) { b->next_co = *co; b->has_next_co = true; } else b->has_next_co = false;
}

extern "C" void GmapDP_batch_clear (gmapdp_batch *b) {
  unpin(b,/*release_buffers*/false);		/* result / script buffers are reused by the next batch */
  b->calls.clear(); b->boxes.clear(); b->box_call.clear(); b->seqpool.clear(); b->probpool.clear(); b->devprob_need = 0;
  b->script_used = 0; b->cells = 0; b->cells8 = 0; b->cells_full = 0; b->cells8_full = 0; b->uploaded = false; b->overflow = false; b->err.clear();
  b->count_mismatch = 0;
}
extern "C" void GmapDP_batch_free (gmapdp_batch *b) { if (b) { unpin(b,true); delete b; } }
extern "C" long GmapDP_batch_cells8 (const gmapdp_batch *b) { return b->cells8; }
extern "C" int GmapDP_batch_ncalls (const gmapdp_batch *b) { return (int) b->calls.size(); }
extern "C" int GmapDP_batch_nboxes (const gmapdp_batch *b) { return (int) b->boxes.size(); }
extern "C" long GmapDP_batch_cells (const gmapdp_batch *b) { return b->cells; }
extern "C" const char *GmapDP_batch_error (const gmapdp_batch *b) { return b->err.c_str(); }
extern "C" size_t GmapDP_batch_h2d_bytes (const gmapdp_batch *b) {
  return b->boxes.size() * (sizeof(gmapdp_box) + sizeof(int)) + b->seqpool.size() + b->probpool.size() * sizeof(double);
}
extern "C" size_t GmapDP_batch_d2h_bytes (const gmapdp_batch *b) {
  return b->boxes.size() * sizeof(gmapdp_result) + b->script_used * sizeof(uint32_t) + sizeof(unsigned long long);
}
static void add_cells (gmapdp_batch *b, long n, bool use8, bool full = false) {
  b->cells += n; if (use8) b->cells8 += n;
  if (full) { b->cells_full += n; if (use8) b->cells8_full += n; }
}
extern "C" long GmapDP_batch_cells_full (const gmapdp_batch *b) { return b->cells_full; }
extern "C" long GmapDP_batch_cells8_full (const gmapdp_batch *b) { return b->cells8_full; }

static gmapdp_box blank_box () { gmapdp_box x; memset(&x,0,sizeof(x)); return x; }

static void diag_only (Pushed &l, const Side &sd, Counts &n, int r, int c, int side = 0) { Replayer rp(l,sd,n,false,side); rp.diag(r,c); }

/* ------------------------------------------------------------------------------------------------ */
extern "C" int GmapDP_single_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
				  int rlength, int glength, int roffset, int goffset,
				  const char *gsequence, const char *gsequence_alt, int jump_late_p,
				  int extraband_single, int widebandp, double defect_rate) {
  static const int opens[3] = {-8,-7,-6}, extends[3] = {-3,-2,-1};		/* dynprog.h:59-68 */
  begin_queue(b);
  b->calls.emplace_back();
  Call &c = b->calls.back();
  const int id = (int) b->calls.size() - 1;
  c.mode = GMAPDP_SINGLE;
  c.iout[0] = dynprogindex; c.dpi = dynprogindex;
  const int qual = quality(defect_rate), mt = qual;
  if (rlength <= 0 || glength <= 0 || rlength > b->max_rlength || glength > b->max_glength) {
    c.iout[1] = NEG_INFINITY_32; bump(c.iout[0]); c.done = true; return id;
  }
  if (gsequence[0] == '\0') { c.iout[1] = NEG_INFINITY_32; c.done = true; return id; }
  c.q.assign(rsequence,rlength); c.quc.assign(rsequenceuc,rlength);
  c.gL.assign(gsequence,glength); c.gLa.assign(gsequence_alt,glength);
  c.rlenL = rlength; c.glenL = glength; c.roffset = roffset; c.goffset = goffset;

  if (glength == rlength) {			/* single_gap_simple */
    Pushed l; Counts n;
    Side sd = {c.q.data(),c.quc.data(),c.gL.data(),c.gLa.data(),roffset,goffset,false};
    for (int r = 1; r <= rlength; r++) diag_only(l,sd,n,r,r);
    if (n.nmismatches <= 1 && !l.empty()) {
      c.iout[1] = n.score; c.iout[2] = n.nmatches; c.iout[3] = n.nmismatches; c.iout[4] = c.iout[5] = 0;
      bump(c.iout[0]);
      l.reverse(); c.pairs = std::move(l); c.isnull = false; c.done = true;
      return id;
    }
  }

  gmapdp_box x = blank_box();
  int lband, uband;
  compute_bands(lband,uband,rlength,glength,extraband_single,widebandp != 0);
  const bool use8 = rlength < tables().use8p[mt] && glength < tables().use8p[mt];		/* && : dynprog_single.c:596 */
  x.mode = GMAPDP_SINGLE;
  x.flags = (jump_late_p ? GMAPDP_F_LATE_L : 0) | (use8 ? GMAPDP_F_USE8 : 0);
  x.rlenL = x.rlenR = rlength; x.glenL = x.glenR = glength;
  x.mismatchtype = (int8_t) mt; x.open = (int8_t) (b->user_dynprog ? b->user_open : opens[qual]); x.extend = (int8_t) (b->user_dynprog ? b->user_extend : extends[qual]);
  x.lbandL = x.lbandR = (int16_t) lband; x.ubandL = x.ubandR = (int16_t) uband;
  x.qL_off = x.qR_off = b->add_bytes(c.quc.data(),rlength);
  if (take_coords(b,c,x,false)) x.gR_off = x.gRalt_off = x.gL_off;
  else {
    x.gL_off = x.gR_off = b->add_bytes(c.gL.data(),glength);
    x.gLalt_off = x.gRalt_off = (c.gLa == c.gL) ? x.gL_off : b->add_bytes(c.gLa.data(),glength);
  }
  add_cells(b,gdp_cells_full(rlength,glength,lband,uband),use8,true);
  c.box = (int) b->boxes.size(); c.queued();
  b->boxes.push_back(x); b->box_call.push_back(id);
  return id;
}

static int end_gap (gmapdp_batch *b, bool end5, int dynprogindex, const char *rseq, const char *rsequc,
		    int rlength, int glength, int roffset, int goffset, const char *gseq, const char *galt,
		    int jump_late_p, int extraband_end, double defect_rate, int endalign, int require_pos_score_p) {
  static const int opens[3] = {-10,-8,-6}, extends[3] = {-2,-2,-2};		/* dynprog_end.c:89-95 */
  begin_queue(b);
  b->calls.emplace_back();
  Call &c = b->calls.back();
  const int id = (int) b->calls.size() - 1;
  c.mode = end5 ? GMAPDP_END5 : GMAPDP_END3;
  c.end5 = end5; c.endalign = endalign; c.require_pos = require_pos_score_p != 0;
  c.iout[0] = dynprogindex; c.dpi = dynprogindex;
  const int qual = quality(defect_rate);
  /* early exits leave traceback_score = 0 and the counts = 0, dynprogindex untouched */
  if (rlength <= 0) { c.done = true; return id; }
  if (endalign != GMAPDP_QUERYEND_NOGAPS && rlength > b->max_rlength) rlength = b->max_rlength;
  if (end5 && goffset < 0) { c.done = true; return id; }
  if (glength <= 0) { c.done = true; return id; }
  if (endalign != GMAPDP_QUERYEND_NOGAPS && glength > b->max_glength) glength = b->max_glength;
  if (gseq[0] == '\0') { c.done = true; return id; }

  /* forward copies.  5': the query is read backwards from rseq[0]; the segment array's LAST glength
     chars are the ones the reference fetches after chopping (it fetches leftwards from rev_goffset). */
  if (end5) { c.q.assign(rseq - (rlength - 1),rlength); c.quc.assign(rsequc - (rlength - 1),rlength); }
  else { c.q.assign(rseq,rlength); c.quc.assign(rsequc,rlength); }
  c.gL.assign(gseq,glength); c.gLa.assign(galt,glength);
  c.rlenL = rlength; c.glenL = glength; c.roffset = roffset; c.goffset = goffset;

  if (endalign == GMAPDP_QUERYEND_NOGAPS) {
    /* find_best_endpoint_to_queryend_nogaps + traceback_nogaps: no fill */
    int best = std::min(rlength,glength);
    Pushed l; Counts n;
    Side sd;
    if (end5) sd = Side{c.q.data() + rlength - 1,c.quc.data() + rlength - 1,c.gL.data() + glength - 1,c.gLa.data() + glength - 1,roffset,goffset,true};
    else sd = Side{c.q.data(),c.quc.data(),c.gL.data(),c.gLa.data(),roffset,goffset,false};
    for (int r = best, cc = best; r > 0 && cc > 0; r--, cc--) diag_only(l,sd,n,r,cc);
    c.iout[1] = n.score; c.iout[2] = n.nmatches; c.iout[3] = n.nmismatches;
    /* no (nmatches+1 < nmismatches) filter for NOGAPS; strip leading INDEL_COMP pairs: there are none */
    bump(c.iout[0]);
    if (!l.empty()) {
      c.isnull = false;
      if (end5) l.reverse();
      c.pairs = std::move(l);
    }
    c.done = true;
    return id;
  }
  if (require_pos_score_p) {
    /* the reference still runs the fills, then skips the traceback because *traceback_score was just
       zeroed (dynprog_end.c:1558-1574): the result does not depend on the fill, so no box is queued */
    bump(c.iout[0]); c.done = true; return id;
  }

  gmapdp_box x = blank_box();
  int lband, uband;
  const bool wide = (endalign != GMAPDP_QUERYEND_INDELS);
  compute_bands(lband,uband,rlength,glength,extraband_end,wide);
  const int mt = 3;	/* ENDQ */
  const bool use8 = rlength < tables().use8p[mt] || glength < tables().use8p[mt];		/* || : dynprog_end.c:1414 */
  const bool late = end5 ? !jump_late_p : (jump_late_p != 0);
  x.mode = c.mode;
  x.flags = (late ? GMAPDP_F_LATE_L : 0) | (use8 ? GMAPDP_F_USE8 : 0) | (wide ? 0 : GMAPDP_F_LASTROW);
  x.rlenL = x.rlenR = rlength; x.glenL = x.glenR = glength;
  x.mismatchtype = (int8_t) mt; x.open = (int8_t) (b->user_dynprog ? b->user_open : opens[qual]); x.extend = (int8_t) (b->user_dynprog ? b->user_extend : extends[qual]);
  x.lbandL = x.lbandR = (int16_t) lband; x.ubandL = x.ubandR = (int16_t) uband;
  x.qL_off = x.qR_off = b->add_bytes(c.quc.data(),rlength);
  if (take_coords(b,c,x,false)) x.gR_off = x.gRalt_off = x.gL_off;
  else {
    x.gL_off = x.gR_off = b->add_bytes(c.gL.data(),glength);
    x.gLalt_off = x.gRalt_off = (c.gLa == c.gL) ? x.gL_off : b->add_bytes(c.gLa.data(),glength);
  }
  x.revmask = end5 ? 1 : 0;
  add_cells(b,gdp_cells_tri(rlength,glength,uband) + gdp_cells_tri(glength,rlength,lband),use8);
  c.box = (int) b->boxes.size(); c.queued();
  b->boxes.push_back(x); b->box_call.push_back(id);
  return id;
}

extern "C" int GmapDP_end5_gap (gmapdp_batch *b, int dynprogindex, const char *rev_rsequence, const char *rev_rsequenceuc,
				int rlength, int glength, int rev_roffset, int rev_goffset,
				const char *rev_gsequence, const char *rev_gsequence_alt, int jump_late_p,
				int extraband_end, double defect_rate, int endalign, int require_pos_score_p) {
  return end_gap(b,true,dynprogindex,rev_rsequence,rev_rsequenceuc,rlength,glength,rev_roffset,rev_goffset,
		 rev_gsequence,rev_gsequence_alt,jump_late_p,extraband_end,defect_rate,endalign,require_pos_score_p);
}
extern "C" int GmapDP_end3_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
				int rlength, int glength, int roffset, int goffset,
				const char *gsequence, const char *gsequence_alt, int jump_late_p,
				int extraband_end, double defect_rate, int endalign, int require_pos_score_p) {
  return end_gap(b,false,dynprogindex,rsequence,rsequenceuc,rlength,glength,roffset,goffset,
		 gsequence,gsequence_alt,jump_late_p,extraband_end,defect_rate,endalign,require_pos_score_p);
}

/* ------------------------------------------------------------------------------------------------ */
extern "C" int GmapDP_genome_gap (gmapdp_batch *b, int dynprogindex, const char *rsequence, const char *rsequenceuc,
				  int rlength, int glengthL, int glengthR, int roffset, int goffsetL, int rev_goffsetR,
				  const char *gsequenceL, const char *gsequenceL_alt,
				  const char *rev_gsequenceR, const char *rev_gsequenceR_alt,
				  const double *left_probabilities, const double *right_probabilities,
				  int cdna_direction, int jump_late_p, int extraband_paired, double defect_rate,
				  int maxpeelback, int halfp, int finalp) {
  static const int opens[3] = {-8,-7,-6}, extends[3] = {-3,-2,-1};	/* PAIRED_* == SINGLE_*, dynprog.h:59-75 */
  (void) maxpeelback;
  begin_queue(b);
  b->calls.emplace_back();
  Call &c = b->calls.back();
  const int id = (int) b->calls.size() - 1;
  c.mode = GMAPDP_GENOME;
  int *dynprogindex_p = &c.iout[0], *new_left = &c.iout[1], *new_right = &c.iout[2], *tbscore = &c.iout[3];
  int *nmatches = &c.iout[4], *nmismatches = &c.iout[5], *exonhead = &c.iout[8], *introntype = &c.iout[9];
  c.dpi = dynprogindex;
  /* out-parameters the reference may leave unwritten keep the caller's sentinel */
  for (int k = 1; k < 10; k++) c.iout[k] = GMAPDP_UNSET;
  *dynprogindex_p = dynprogindex;
  c.iout[4] = c.iout[5] = c.iout[6] = c.iout[7] = 0; c.dout[0] = c.dout[1] = 0.0; *introntype = 0;
  if (rlength <= 1) { *tbscore = NEG_INFINITY_32; c.done = true; return id; }
  const int qual = quality(defect_rate), mt = qual;
  if (rlength > b->max_rlength || glengthL > b->max_glength || glengthR > b->max_glength) {
    *new_left = goffsetL - 1; *new_right = rev_goffsetR + 1; *exonhead = roffset + rlength - 1;
    bump(*dynprogindex_p); *tbscore = NEG_INFINITY_32; c.done = true; return id;
  }
  if (gsequenceL[0] == '\0' || rev_gsequenceR[0] == '\0') { *tbscore = NEG_INFINITY_32; c.done = true; return id; }

  c.q.assign(rsequence,rlength); c.quc.assign(rsequenceuc,rlength);
  c.gL.assign(gsequenceL,glengthL); c.gLa.assign(gsequenceL_alt,glengthL);
  c.gR.assign(rev_gsequenceR,glengthR); c.gRa.assign(rev_gsequenceR_alt,glengthR);
  c.owner = b;
  if (b->has_cur_co && b->cur_co.probs && b->g_blocks && b->g_tables && gsequenceL_alt && strncmp(gsequenceL,gsequenceL_alt,glengthL) == 0 && strncmp(rev_gsequenceR,rev_gsequenceR_alt,glengthR) == 0) { c.devp = true; c.co = b->cur_co; }
  else {
    if (!left_probabilities || !right_probabilities) { b->err = "GmapDP_genome_gap: no probability arrays and no resident genome"; c.done = true; *tbscore = NEG_INFINITY_32; return id; }
    c.lp.assign(left_probabilities,left_probabilities + std::max(glengthL - 1,0));
    c.rp.assign(right_probabilities,right_probabilities + std::max(glengthR - 1,0));
  }
  c.rlenL = rlength; c.glenL = glengthL; c.glenR = glengthR; c.roffset = roffset; c.goffset = goffsetL; c.goffsetR = rev_goffsetR;
  const int rev_roffset = roffset + rlength - 1;
  const GdpHostTables &T = tables();

  if (!finalp && defect_rate < 0.014) {
    /* genome_gap_simple (prelim intron scores; no known splice sites on this path) */
    const GdpTables &dt = dev_tables();
    const int *isc = dt.isc[cdna_direction > 0 ? 0 : (cdna_direction < 0 ? 1 : 2)][0];
    auto pairscore = [&](char q, char g, char ga) { return (int) std::max(T.pd[mt][q & 127][g & 127],T.pd[mt][q & 127][ga & 127]); };
    const char *ruc = c.quc.data(), *gl = c.gL.data(), *gla = c.gLa.data();
    const char *rgr = c.gR.data() + glengthR - 1, *rgra = c.gRa.data() + glengthR - 1, *rruc = ruc + rlength - 1;
    int scoreR = 0, scoreL = 0, bestscore = 0, bestscoreI = 0, bestrL = 0, bestrR = 0, itype = 0;
    for (int rR = 1; rR < rlength; rR++) scoreR += pairscore(rruc[1-rR],rgr[1-rR],rgra[1-rR]);
    for (int rL = 1, rR = rlength - 1; rL < rlength; rL++, rR--) {
      scoreL += pairscore(ruc[rL-1],gl[rL-1],gla[rL-1]);
      const int ldi = left_dinucl(gl[rL],gla[rL],gl[rL+1],gla[rL+1]);
      const int rdi = right_dinucl(rgr[-rR-1],rgra[-rR-1],rgr[-rR],rgra[-rR]);
      itype = ldi & rdi;
      const int scoreI = itype ? isc[31 - __builtin_clz((unsigned) itype)] : 0;
      int score;
      if (itype != 0 && (score = scoreL + scoreI + scoreR) >= bestscore) {
	bestscore = score; bestscoreI = scoreI; bestrL = rL; bestrR = rR; *introntype = itype;
      }
      scoreR -= pairscore(rruc[1-rR],rgr[1-rR],rgra[1-rR]);
    }
    const int finalscore = halfp ? bestscore - bestscoreI / 2 : bestscore;
    bool result = finalscore > 0;
    if (result) {
      c.dout[0] = call_prob(c,false,bestrL); c.dout[1] = call_prob(c,true,bestrR);		/* get_splicesite_probs at (bestrL, bestrR) */
      if (c.dout[0] < 0.90 || c.dout[1] < 0.90) result = false;
    }
    *tbscore = *nmatches = *nmismatches = 0;
    if (result) {
      Pushed l; Counts n;
      Side sl = {c.q.data(),c.quc.data(),gl,gla,roffset,goffsetL,false};
      for (int r = 1; r <= bestrL; r++) diag_only(l,sl,n,r,r);
      *new_left = goffsetL + (bestrL - 1);
      *new_right = *exonhead = rev_goffsetR - (bestrR - 1);
      gmapdp_gapinfo &gp = push_gapholder(l,false,0,(*new_right) - (*new_left) - 1);
      gp.introntype = itype;		/* the LAST iteration's introntype (dynprog_genome.c:3230) */
      gp.donor_prob = c.dout[0]; gp.acceptor_prob = c.dout[1];
      Side sr = {c.q.data() + rlength - 1,c.quc.data() + rlength - 1,rgr,rgra,rev_roffset,rev_goffsetR,true};
      for (int r = bestrR; r > 0; r--) diag_only(l,sr,n,r,r,1);
      *tbscore = n.score; *nmatches = n.nmatches; *nmismatches = n.nmismatches;
      bump(*dynprogindex_p);
      l.reverse(); c.pairs = std::move(l); c.isnull = false; c.done = true;
      return id;
    }
  }
  c.introntype_in = *introntype;

  gmapdp_box x = blank_box();
  int lbandL, ubandL, lbandR, ubandR;
  compute_bands(lbandL,ubandL,rlength,glengthL,extraband_paired,true);
  compute_bands(lbandR,ubandR,rlength,glengthR,extraband_paired,true);
  const bool use8 = rlength < T.use8p[mt] || (glengthL < T.use8p[mt] && glengthR < T.use8p[mt]);	/* dynprog_genome.c:3503 */
  x.mode = GMAPDP_GENOME;
  x.flags = (jump_late_p ? GMAPDP_F_LATE_L : GMAPDP_F_LATE_R) | (use8 ? GMAPDP_F_USE8 : 0) |
    (finalp ? GMAPDP_F_FINALP : 0) | (halfp ? GMAPDP_F_HALFP : 0);
  x.rlenL = x.rlenR = rlength; x.glenL = glengthL; x.glenR = glengthR;
  x.mismatchtype = (int8_t) mt; x.open = (int8_t) (b->user_dynprog ? b->user_open : opens[qual]); x.extend = (int8_t) (b->user_dynprog ? b->user_extend : extends[qual]);
  x.cdna_direction = (int8_t) cdna_direction;
  x.lbandL = (int16_t) lbandL; x.ubandL = (int16_t) ubandL; x.lbandR = (int16_t) lbandR; x.ubandR = (int16_t) ubandR;
  x.qL_off = x.qR_off = b->add_bytes(c.quc.data(),rlength);
  if (!take_coords(b,c,x,true)) {
    x.gL_off = b->add_bytes(c.gL.data(),glengthL);
    x.gLalt_off = (c.gLa == c.gL) ? x.gL_off : b->add_bytes(c.gLa.data(),glengthL);
    x.gR_off = b->add_bytes(c.gR.data(),glengthR);
    x.gRalt_off = (c.gRa == c.gR) ? x.gR_off : b->add_bytes(c.gRa.data(),glengthR);
  }
  if (!c.devp) {
    x.probL_off = (uint32_t) b->probpool.size(); b->probpool.insert(b->probpool.end(),c.lp.begin(),c.lp.end()); b->probpool.push_back(0.0);
    x.probR_off = (uint32_t) b->probpool.size(); b->probpool.insert(b->probpool.end(),c.rp.begin(),c.rp.end()); b->probpool.push_back(0.0);
  }
  x.offdiff = rev_goffsetR - goffsetL;
  x.revmask = 2;
  add_cells(b,gdp_cells_tri(rlength,glengthL,ubandL) + gdp_cells_tri(glengthL,rlength,lbandL) +
	    gdp_cells_tri(rlength,glengthR,ubandR) + gdp_cells_tri(glengthR,rlength,lbandR),use8);
  c.box = (int) b->boxes.size(); c.queued();
  b->boxes.push_back(x); b->box_call.push_back(id);
  return id;
}

/* ------------------------------------------------------------------------------------------------ */
extern "C" int GmapDP_cdna_gap (gmapdp_batch *b, int dynprogindex,
				const char *rsequenceL, const char *rsequence_ucL,
				const char *rev_rsequenceR, const char *rev_rsequence_ucR,
				int rlengthL, int rlengthR, int glength, int roffsetL, int rev_roffsetR, int goffset,
				const char *gsequence, const char *gsequence_alt,
				const char *rev_gsequence, const char *rev_gsequence_alt,
				int jump_late_p, int extraband_paired, double defect_rate) {
  begin_queue(b);
  b->calls.emplace_back();
  Call &c = b->calls.back();
  const int id = (int) b->calls.size() - 1;
  c.mode = GMAPDP_CDNA;
  c.iout[0] = dynprogindex; c.dpi = dynprogindex; c.iout[1] = GMAPDP_UNSET; c.iout[2] = 0;
  if (glength <= 1) { c.done = true; return id; }
  const int mt = quality(defect_rate);
  if (glength > b->max_glength || rlengthR > b->max_rlength || rlengthL > b->max_rlength) { bump(c.iout[0]); c.done = true; return id; }
  if (gsequence[0] == '\0' || rev_gsequence[0] == '\0') { c.done = true; return id; }
  const GdpHostTables &T = tables();
  /* rsequenceL and rev_rsequenceR address the same query: keep the whole span [roffsetL, rev_roffsetR] of the raw chars,
     which the insert-pairs loop (dynprog_cdna.c:1010-1014) indexes past rlengthL */
  c.q.assign(rsequenceL,std::max(rlengthL,rev_roffsetR - roffsetL + 1)); c.quc.assign(rsequence_ucL,rlengthL);
  c.qR.assign(rev_rsequenceR - (rlengthR - 1),rlengthR); c.qRuc.assign(rev_rsequence_ucR - (rlengthR - 1),rlengthR);
  c.gL.assign(gsequence,glength); c.gLa.assign(gsequence_alt,glength);
  c.gR.assign(rev_gsequence,glength); c.gRa.assign(rev_gsequence_alt,glength);
  c.rlenL = rlengthL; c.rlenR = rlengthR; c.glenL = c.glenR = glength;
  c.roffset = roffsetL; c.roffsetR = rev_roffsetR; c.goffset = goffset;

  gmapdp_box x = blank_box();
  int lbandL, ubandL, lbandR, ubandR;
  compute_bands(lbandL,ubandL,rlengthL,glength,extraband_paired,true);
  compute_bands(lbandR,ubandR,rlengthR,glength,extraband_paired,true);
  const bool use8 = glength < T.use8p[mt] || (rlengthL < T.use8p[mt] && rlengthR <= T.use8p[mt]);	/* stray <= : dynprog_cdna.c:909 */
  x.mode = GMAPDP_CDNA;
  x.flags = (jump_late_p ? (GMAPDP_F_LATE_L | GMAPDP_F_BRIDGE_LATE) : GMAPDP_F_LATE_R) | (use8 ? GMAPDP_F_USE8 : 0);
  x.rlenL = rlengthL; x.rlenR = rlengthR; x.glenL = x.glenR = glength;
  x.mismatchtype = (int8_t) mt; x.open = -10; x.extend = -7;			/* CDNA_OPEN / CDNA_EXTEND, dynprog_cdna.c:32-38 */
  x.lbandL = (int16_t) lbandL; x.ubandL = (int16_t) ubandL; x.lbandR = (int16_t) lbandR; x.ubandR = (int16_t) ubandR;
  x.qL_off = b->add_bytes(c.quc.data(),rlengthL);
  x.qR_off = b->add_bytes(c.qRuc.data(),rlengthR);
  if (c.gR == c.gL && take_coords(b,c,x,false)) x.gR_off = x.gRalt_off = x.gL_off;
  else {
    x.gL_off = b->add_bytes(c.gL.data(),glength);
    x.gLalt_off = (c.gLa == c.gL) ? x.gL_off : b->add_bytes(c.gLa.data(),glength);
    x.gR_off = (c.gR == c.gL) ? x.gL_off : b->add_bytes(c.gR.data(),glength);
    x.gRalt_off = (c.gRa == c.gR) ? x.gR_off : b->add_bytes(c.gRa.data(),glength);
  }
  x.offdiff = rev_roffsetR - roffsetL;
  x.revmask = 2;
  add_cells(b,gdp_cells_tri(rlengthL,glength,ubandL) + gdp_cells_tri(glength,rlengthL,lbandL) +
	    gdp_cells_tri(rlengthR,glength,ubandR) + gdp_cells_tri(glength,rlengthR,lbandR),use8);
  c.box = (int) b->boxes.size(); c.queued();
  b->boxes.push_back(x); b->box_call.push_back(id);
  return id;
}

/* ------------------------------------------------------------------------------------------------
 * Completion: replay the device results
 * ---------------------------------------------------------------------------------------------- */
/* Are the characters the replay compares -- the upper-cased query, the genomic segments and their alt twins -- all plain
   A C G T?  Then a mismatch is never an ambiguity match (the consistency table only relates IUPAC codes) and no position
   is a chromosome edge ('*'): Replayer::diag_run takes its short path.  Looked at once per call. */
static bool all_acgt (const std::string &s) {
  const char *p = s.data();
  size_t n = s.size(), i = 0;
#if defined(__SSE2__)
  const __m128i A = _mm_set1_epi8('A'), Cc = _mm_set1_epi8('C'), G = _mm_set1_epi8('G'), T = _mm_set1_epi8('T');
  for (; i + 16 <= n; i += 16) {
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p + i));
    const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(x,A),_mm_cmpeq_epi8(x,Cc)),_mm_or_si128(_mm_cmpeq_epi8(x,G),_mm_cmpeq_epi8(x,T)));
    if (_mm_movemask_epi8(ok) != 0xffff) return false;
  }
#endif
  for (; i < n; i++) if (p[i] != 'A' && p[i] != 'C' && p[i] != 'G' && p[i] != 'T') return false;
  return true;
}
static bool plain_p (Call &c) {
  if (c.plain < 0) {
    static const bool acgt_distinct = []() {	/* the premise, checked against the table itself */
      const GdpHostTables &t = tables();
      const char nt[4] = {'A','C','G','T'};
      for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) if (a != b && t.cons[(int) nt[a]][(int) nt[b]]) return false;
      return true;
    }();
    c.plain = (acgt_distinct && all_acgt(c.quc) && all_acgt(c.qRuc) && all_acgt(c.gL) && all_acgt(c.gLa) && all_acgt(c.gR) && all_acgt(c.gRa)) ? 1 : 0;
    c.ucL = (c.q.size() >= c.quc.size() && c.q.compare(0,c.quc.size(),c.quc) == 0);
    c.ucR = (c.qR.size() >= c.qRuc.size() && c.qR.compare(0,c.qRuc.size(),c.qRuc) == 0);
    c.altL = (c.gLa == c.gL); c.altR = (c.gRa == c.gR);
  }
  return c.plain == 1;
}

static void check_counts (int *count_mismatch, const gmapdp_result &r, const Counts &n) {
  if (r.tb_score != n.score || r.nmatches != n.nmatches || r.nmismatches != n.nmismatches || r.nopens != n.nopens || r.nindels != n.nindels)
    __atomic_fetch_add(count_mismatch,1,__ATOMIC_RELAXED);
}

/* thread-safe: touches only its own call (and the shared mismatch counter, atomically) */
static void finish_call (Call &c, const gmapdp_result &r, const uint32_t *ops, int *count_mismatch) {
  c.dpi = c.iout[0];
  /* the list is built in place in the call's own array (its memory survives GmapDP_batch_rewind) */
  Pushed &l = c.pairs;
  const bool plain = plain_p(c);
  /* identical streams are read once: the query's upper case where the query is upper case, the alt segment where there is none */
  const char *const qucL = c.ucL ? c.q.data() : c.quc.data(), *const qucR = c.ucR ? c.qR.data() : c.qRuc.data();
  const char *const gLa = c.altL ? c.gL.data() : c.gLa.data(), *const gRa = c.altR ? c.gR.data() : c.gRa.data();
  const size_t sideL = (size_t) c.rlenL + c.glenL + 4, sideR = (size_t) c.rlenR + c.glenR + 4;
  if (c.mode == GMAPDP_SINGLE) {
    Counts n;
    l.start(0,sideL);
    Side sd = {c.q.data(),qucL,c.gL.data(),gLa,c.roffset,c.goffset,false};
    sd.plain = plain;
    Replayer(l,sd,n).run(0,c.rlenL,c.glenL,ops,r.script_lenA);
    check_counts(count_mismatch,r,n);
    c.iout[1] = n.score; c.iout[2] = n.nmatches; c.iout[3] = n.nmismatches; c.iout[4] = n.nopens; c.iout[5] = n.nindels;
    bump(c.iout[0]);
    c.isnull = l.empty();			/* List_reverse of the consed list = push order */

  } else if (c.mode == GMAPDP_END5 || c.mode == GMAPDP_END3) {
    Counts n;
    Side sd;
    /* 5' ends hand out the reversed list: written backwards, so the pairs pushed first are its tail */
    if (c.end5) { l.start(sideL,0); sd = Side{c.q.data() + c.rlenL - 1,qucL + c.rlenL - 1,c.gL.data() + c.glenL - 1,gLa + c.glenL - 1,c.roffset,c.goffset,true}; }
    else { l.start(0,sideL); sd = Side{c.q.data(),qucL,c.gL.data(),gLa,c.roffset,c.goffset,false}; }
    sd.plain = plain;
    Replayer(l,sd,n,c.end5).run(r.bestcL >= r.bestrL ? 1 : 2,r.bestrL,r.bestcL,ops,r.script_lenA);
    check_counts(count_mismatch,r,n);
    c.iout[1] = n.score; c.iout[2] = n.nmatches; c.iout[3] = n.nmismatches; c.iout[4] = n.nopens; c.iout[5] = n.nindels;
    if ((c.endalign == GMAPDP_QUERYEND_GAP || c.endalign == GMAPDP_BEST_LOCAL) && (n.nmatches + 1) < n.nmismatches) {
      c.iout[1] = 0; l.clear();
    } else {
      /* leading indels of the list in push order */
      const size_t sz = l.size();
      size_t k = 0;
      if (c.end5) { while (k < sz && cpair_comp(l[sz - 1 - k]) == COMP_INDEL) k++; l.drop_back(k); }
      else { while (k < sz && cpair_comp(l[k]) == COMP_INDEL) k++; l.drop_front(k); }
    }
    bump(c.iout[0]);
    c.isnull = l.empty();

  } else if (c.mode == GMAPDP_GENOME) {
    if (r.status != 0) { c.iout[3] = -100; c.isnull = true; l.clear(); return; }
    const int rlength = c.rlenL, rev_roffset = c.roffset + rlength - 1;
    const int bestrL = r.bestrL, bestrR = r.bestrR, bestcL = r.bestcL, bestcR = r.bestcR;
    c.dout[0] = call_prob(c,false,bestcL); c.dout[1] = call_prob(c,true,bestcR);
    c.iout[1] = c.goffset + (bestcL - 1);
    c.iout[2] = c.goffsetR - (bestcR - 1);
    c.iout[8] = rev_roffset - (bestrR - 1);
    Counts n;
    l.start(sideR,sideL + 1);
    /* the right-hand side reversed (written backwards), the gap holder, the left-hand side */
    Side sr = {c.q.data() + rlength - 1,qucL + rlength - 1,c.gR.data() + c.glenR - 1,gRa + c.glenR - 1,rev_roffset,c.goffsetR,true};
    sr.plain = plain;
    Replayer(l,sr,n,true,1).run(bestcR >= bestrR ? 1 : 2,bestrR,bestcR,ops,r.script_lenA);
    gmapdp_gapinfo &gp = push_gapholder(l,false,(rev_roffset - bestrR) - (c.roffset + bestrL) + 1,c.iout[2] - c.iout[1] - 1);
    gp.introntype = c.introntype_in; gp.donor_prob = c.dout[0]; gp.acceptor_prob = c.dout[1];
    Side sl = {c.q.data(),qucL,c.gL.data(),gLa,c.roffset,c.goffset,false};
    sl.plain = plain;
    Replayer(l,sl,n).run(bestcL >= bestrL ? 1 : 2,bestrL,bestcL,ops + r.script_lenA,r.script_lenB);
    check_counts(count_mismatch,r,n);
    c.iout[3] = n.score; c.iout[4] = n.nmatches; c.iout[5] = n.nmismatches; c.iout[6] = n.nopens; c.iout[7] = n.nindels;
    if (l.size() == 1) l.clear();
    bump(c.iout[0]);
    if (maxnegscore_headfirst(l) < -10) { c.iout[3] = -100; c.isnull = true; l.clear(); return; }
    c.isnull = l.empty();

  } else {	/* cdna */
    const int bestrL = r.bestrL, bestrR = r.bestrR, bestcL = r.bestcL, bestcR = r.bestcR;
    const int rev_goffset = c.goffset + c.glenL - 1;
    Counts n;
    l.start(sideR,sideL + 2 * INSERT_PAIRS + 2);
    Side sr = {c.qR.data() + c.rlenR - 1,qucR + c.rlenR - 1,c.gR.data() + c.glenL - 1,gRa + c.glenL - 1,c.roffsetR,rev_goffset,true};
    sr.plain = plain;
    Replayer(l,sr,n,true,1).run(bestcR >= bestrR ? 1 : 2,bestrR,bestcR,ops,r.script_lenA);
    const int queryjump = (c.roffsetR - bestrR) - (c.roffset + bestrL) + 1;
    const int genomejump = (rev_goffset - bestcR) - (c.goffset + bestcL) + 1;
    if (queryjump == INSERT_PAIRS && genomejump == INSERT_PAIRS) {
      l.use_side(0,c.roffset,c.goffset);			/* the inserted pairs lie next to the left-hand side */
      for (int k = c.roffsetR - bestrR; k >= c.roffset + bestrL; k--)
	push_pair(l,false,k,rev_goffset - bestcR + 1,c.q[k - c.roffset],COMP_SHORTGAP,' ',' ');
      for (int k = rev_goffset - bestcR; k >= c.goffset + bestcL; k--)
	push_pair(l,false,c.roffset + bestrL,k,' ',COMP_SHORTGAP,c.gL[k - c.goffset],gLa[k - c.goffset]);
    } else {
      push_gapholder(l,false,queryjump,genomejump);
      c.iout[2] = 1;
    }
    Side sl = {c.q.data(),qucL,c.gL.data(),gLa,c.roffset,c.goffset,false};
    sl.plain = plain;
    Replayer(l,sl,n).run(bestcL >= bestrL ? 1 : 2,bestrL,bestcL,ops + r.script_lenA,r.script_lenB);
    check_counts(count_mismatch,r,n);
    c.iout[1] = n.score;
    if (l.size() == 1) l.clear();
    bump(c.iout[0]);
    c.isnull = l.empty();
  }
#if defined(__SSE2__) && !defined(GMAPDP_NO_NT_STORES)
  _mm_sfence();		/* the records were written with non-temporal stores: visible before the call is marked done */
#endif
}

static int replay_threads () {
  const char *e = getenv("GMAPDP_REPLAY_THREADS");
  int n = e ? atoi(e) : (int) std::min(32u,std::max(1u,std::thread::hardware_concurrency()));
  return n < 1 ? 1 : n;
}

/* The replay of one call touches nothing but that call, so large batches are replayed by several host threads
   (GMAPDP_REPLAY_THREADS, default: all hardware threads, at most 32). */
static int finish_all (gmapdp_batch *b, const gmapdp_result *results, const uint32_t *script) {
  const size_t n = b->calls.size();
  auto work = [&](size_t i0, size_t i1) {
    for (size_t i = i0; i < i1; i++) {
      Call &c = b->calls[i];
      if (!c.done && c.box >= 0) { const gmapdp_result &r = results[c.box]; finish_call(c,r,script + r.script_off,&b->count_mismatch); c.done = true; }
    }
  };
  const int nthreads = (b->boxes.size() >= 4096) ? replay_threads() : 1;
  if (nthreads == 1) work(0,n);
  else {
    /* dynamic blocks: boxes differ by three orders of magnitude in pair count */
    std::atomic<size_t> next(0);
    const size_t blk = 256;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&]() { for (;;) { const size_t i0 = next.fetch_add(blk); if (i0 >= n) break; work(i0,std::min(n,i0 + blk)); } });
    for (auto &x : th) x.join();
  }
  if (b->count_mismatch) { b->err = "device traceback counts disagree with the host replay"; return GMAPDP_ERR_ARG; }
  return GMAPDP_OK;
}

/* Puts the device calls back into their queued state (results and pair lists dropped), so that the same batch can be
   completed again: benchmarks time GmapDP_batch_run -- device and replay -- repeatedly on one batch. */
extern "C" void GmapDP_batch_rewind (gmapdp_batch *b) {
  for (Call &c : b->calls) {
    if (c.box < 0) continue;
    c.done = false; c.isnull = true; c.pairs.clear();
    for (int i = 0; i < 10; i++) c.iout[i] = c.iout_q[i];
    c.dout[0] = c.dout_q[0]; c.dout[1] = c.dout_q[1];
  }
  b->count_mismatch = 0;
}

/* The device part of the queued calls, for callers that run the boxes themselves (the streaming runtime of
   gmapdp_stream.h runs them in batches shared with other threads) ... */
extern "C" int GmapDP_batch_device_view (const gmapdp_batch *b, const gmapdp_box **boxes, int *nboxes, const uint8_t **seqpool, size_t *seqbytes,
					 const double **probpool, size_t *nprobs) {
  if (b->overflow) return GMAPDP_ERR_CAPACITY;
  *boxes = b->boxes.data(); *nboxes = (int) b->boxes.size();
  *seqpool = b->seqpool.data(); *seqbytes = b->seqpool.size();
  *probpool = b->probpool.data(); *nprobs = b->probpool.size();
  return GMAPDP_OK;
}
/* ... and their completion from results produced elsewhere: results[k] belongs to box k of the view, its ops start at
   script[results[k].script_off] */
extern "C" int GmapDP_batch_complete (gmapdp_batch *b, const gmapdp_result *results, const uint32_t *script) {
  return finish_all(b,results,script);
}

static int ensure_host_buffers (gmapdp_batch *b) {
  if (b->overflow) { b->err = "sequence pool exceeds 4 GiB: split the batch"; return GMAPDP_ERR_CAPACITY; }
  size_t need = 64;
  for (const gmapdp_box &x : b->boxes) {
    need += (size_t) x.rlenL + x.glenL + 4;
    if (x.mode == GMAPDP_GENOME || x.mode == GMAPDP_CDNA) need += (size_t) x.rlenR + x.glenR + 4;
  }
  /* pinning pays off only for large batches: page-locking costs far more than a small staged copy */
  const bool big = (b->seqpool.size() + need * sizeof(uint32_t) > ((size_t) 8 << 20));
  if (big && !b->pinned) {
    /* the pools are complete: pin them in place so every H2D is a straight DMA */
    /* (a failed registration only costs speed: the copy is staged) */
    if (gmapdp_host_register(b->boxes.data(),b->boxes.size() * sizeof(gmapdp_box)) == GMAPDP_OK) b->pin_boxes = b->boxes.data();
    if (gmapdp_host_register(b->seqpool.data(),b->seqpool.size()) == GMAPDP_OK) b->pin_seq = b->seqpool.data();
    if (!b->probpool.empty() && gmapdp_host_register(b->probpool.data(),b->probpool.size() * sizeof(double)) == GMAPDP_OK) b->pin_probs = b->probpool.data();
    b->pinned = true;
  }
  if (big != b->bufs_pinned && (b->results || b->script)) {
    host_free(b,b->results); b->results = NULL; b->nresults = 0;
    host_free(b,b->script); b->script = NULL; b->script_cap = 0;
  }
  b->bufs_pinned = big;
  if (b->nresults < b->boxes.size()) {
    host_free(b,b->results);
    const size_t n = b->boxes.size() + b->boxes.size() / 2 + 16;
    b->results = (gmapdp_result *) (big ? gmapdp_host_alloc(n * sizeof(gmapdp_result)) : malloc(n * sizeof(gmapdp_result)));
    b->nresults = b->results ? n : 0;
  }
  if (b->script_cap < need) {
    host_free(b,b->script);
    const size_t n = need + need / 2;
    b->script = (uint32_t *) (big ? gmapdp_host_alloc(n * sizeof(uint32_t)) : malloc(n * sizeof(uint32_t)));
    b->script_cap = b->script ? n : 0;
  }
  if (!b->results || !b->script) { b->err = "pinned host allocation failed"; return GMAPDP_ERR_CUDA; }
  return GMAPDP_OK;
}

extern "C" int GmapDP_batch_upload (gmapdp_batch *b) {
  if (b->boxes.empty()) { b->uploaded = true; return GMAPDP_OK; }
  if (!b->ctx) { b->err = "no device context: the DP engine has no CPU fallback"; return GMAPDP_ERR_CUDA; }
  int rc = ensure_host_buffers(b);
  if (rc) return rc;
  rc = gmapdp_upload(b->ctx,b->boxes.data(),(int) b->boxes.size(),b->seqpool.data(),b->seqpool.size(),
		     b->probpool.data(),b->probpool.size());
  if (rc) { b->err = gmapdp_last_error(b->ctx); return rc; }
  b->uploaded = true;
  return GMAPDP_OK;
}

extern "C" int GmapDP_batch_run_resident (gmapdp_batch *b, float *kernel_ms) {
  if (kernel_ms) *kernel_ms = 0.f;
  if (b->boxes.empty()) return GMAPDP_OK;
  if (!b->uploaded) { b->err = "batch not uploaded"; return GMAPDP_ERR_ARG; }
  int rc = gmapdp_run_resident(b->ctx,kernel_ms);
  if (rc) b->err = gmapdp_last_error(b->ctx);
  return rc;
}

/* device boundary only: H2D + kernel + D2H of results and scripts, no pair-list replay */
extern "C" int GmapDP_batch_run_device (gmapdp_batch *b) {
  if (b->boxes.empty()) return GMAPDP_OK;
  if (!b->ctx) { b->err = "no device context: the DP engine has no CPU fallback"; return GMAPDP_ERR_CUDA; }
  int rc = ensure_host_buffers(b);
  if (rc) return rc;
  rc = gmapdp_run_batch(b->ctx,b->boxes.data(),(int) b->boxes.size(),b->seqpool.data(),b->seqpool.size(),
			b->probpool.data(),b->probpool.size(),b->results,b->script,b->script_cap,&b->script_used);
  if (rc) b->err = gmapdp_last_error(b->ctx);
  b->uploaded = (rc == 0);
  return rc;
}

extern "C" int GmapDP_batch_download (gmapdp_batch *b) {
  if (b->boxes.empty()) return GMAPDP_OK;
  int rc = ensure_host_buffers(b);
  if (rc) return rc;
  rc = gmapdp_download(b->ctx,b->results,b->script,b->script_cap,&b->script_used);
  if (rc) b->err = gmapdp_last_error(b->ctx);
  return rc;
}

extern "C" int GmapDP_batch_finish (gmapdp_batch *b) {
  int rc = GmapDP_batch_download(b);
  if (rc) return rc;
  return finish_all(b,b->results,b->script);
}

/* Replay that overlaps the device: the batch runs chunk by chunk (gmapdp_run_batch_chunks), and as soon as a chunk's
   results and scripts are on the host its calls are replayed by the pool's threads while the device works on the next
   chunks.  Chunks complete in box order, so "ready" is a prefix of the boxes. */
namespace {
struct ReplayPool {
  gmapdp_batch *b;
  int nboxes, ready = 0;
  bool stop = false;
  std::mutex mu; std::condition_variable cv;
  std::atomic<int> next{0};
  std::vector<std::thread> th;

  static void on_chunk (void *user, int first, int n) {
    ReplayPool *P = (ReplayPool *) user;
    { std::lock_guard<std::mutex> g(P->mu); if (first + n > P->ready) P->ready = first + n; }
    P->cv.notify_all();
  }
  void work () {
    const int blk = 128;
    for (;;) {
      const int i0 = next.fetch_add(blk);
      if (i0 >= nboxes) return;
      const int i1 = std::min(nboxes,i0 + blk);
      {
	std::unique_lock<std::mutex> g(mu);
	cv.wait(g,[&]() { return ready >= i1 || stop; });
	if (stop) return;
      }
      for (int k = i0; k < i1; k++) {
	Call &c = b->calls[b->box_call[k]];
	if (!c.done) { const gmapdp_result &r = b->results[k]; finish_call(c,r,b->script + r.script_off,&b->count_mismatch); c.done = true; }
      }
    }
  }
  void start (int nthreads) { for (int t = 0; t < nthreads; t++) th.emplace_back([this]() { work(); }); }
  void finish (bool ok) {
    { std::lock_guard<std::mutex> g(mu); if (ok) ready = nboxes; else stop = true; }
    cv.notify_all();
    for (auto &x : th) x.join();
  }
};
}

extern "C" int GmapDP_batch_run (gmapdp_batch *b) {
  if (b->boxes.size() < 4096) {
    int rc = GmapDP_batch_run_device(b);
    if (rc) return rc;
    return finish_all(b,b->results,b->script);
  }
  if (!b->ctx) { b->err = "no device context: the DP engine has no CPU fallback"; return GMAPDP_ERR_CUDA; }
  int rc = ensure_host_buffers(b);
  if (rc) return rc;
  ReplayPool P;
  P.b = b; P.nboxes = (int) b->boxes.size();
  P.start(replay_threads());
  rc = gmapdp_run_batch_chunks(b->ctx,b->boxes.data(),(int) b->boxes.size(),b->seqpool.data(),b->seqpool.size(),
			       b->probpool.data(),b->probpool.size(),b->results,b->script,b->script_cap,&b->script_used,
			       ReplayPool::on_chunk,&P);
  P.finish(rc == GMAPDP_OK);
  if (rc) { b->err = gmapdp_last_error(b->ctx); b->uploaded = false; return rc; }
  b->uploaded = true;
  if (b->count_mismatch) { b->err = "device traceback counts disagree with the host replay"; return GMAPDP_ERR_ARG; }
  return GMAPDP_OK;
}

/* order-independent digest of the device results (scores, best cells, counts, scripts): what the
   1M-box runs compare between repetitions / ranks without materialising pair lists */
extern "C" unsigned long long GmapDP_batch_digest (const gmapdp_batch *b) {
  unsigned long long h = 1469598103934665603ull;
  for (size_t k = 0; k < b->boxes.size(); k++) {
    const gmapdp_result &r = b->results[k];
    const int v[12] = {r.status,r.finalscore,r.bestrL,r.bestcL,r.bestrR,r.bestcR,r.tb_score,r.nmatches,r.nmismatches,r.nopens,r.nindels,r.script_lenA + 4096 * r.script_lenB};
    unsigned long long hb = 1099511628211ull * (k + 1);
    for (int j = 0; j < 12; j++) hb = (hb ^ (unsigned) v[j]) * 1099511628211ull;
    for (int j = 0; j < r.script_lenA + r.script_lenB; j++) hb = (hb ^ b->script[r.script_off + j]) * 1099511628211ull;
    h += hb;
  }
  return h;
}

extern "C" int GmapDP_result (const gmapdp_batch *b, int id, int *iout, double *dout, gmapdp_pair *pairs, int maxpairs) {
  if (id < 0 || id >= (int) b->calls.size()) return -2;
  const Call &c = b->calls[id];
  if (!c.done) return -3;
  const int ni = (c.mode == GMAPDP_GENOME) ? 10 : (c.mode == GMAPDP_CDNA ? 3 : 6);
  if (iout) for (int k = 0; k < ni; k++) iout[k] = c.iout[k];
  if (dout && c.mode == GMAPDP_GENOME) { dout[0] = c.dout[0]; dout[1] = c.dout[1]; }
  if (c.isnull || c.pairs.empty()) return -1;
  const int n = (int) c.pairs.size();
  for (int k = 0; k < n && k < maxpairs; k++) {		/* the compact records spelled out */
    const gmapdp_cpair &s = c.pairs[k];
    gmapdp_pair &p = pairs[k];
    memset(&p,0,sizeof(p));
    p.cdna = s.cdna; p.comp = cpair_comp(s); p.genome = s.genome; p.genomealt = s.genomealt;
    if (cpair_is_gap(s)) {
      const gmapdp_gapinfo &g = c.pairs.gaps[s.grel];
      p.querypos = p.genomepos = -1;
      p.gapp = 1; p.queryjump = g.queryjump; p.genomejump = g.genomejump; p.introntype = g.introntype;
      p.donor_prob = g.donor_prob; p.acceptor_prob = g.acceptor_prob;
    } else {
      const int side = ((unsigned char) s.comp >> 7) & 1;
      p.querypos = c.pairs.base.q[side] + (int) s.qrel; p.genomepos = c.pairs.base.g[side] + (int) s.grel;
      p.dynprogindex = c.dpi;
    }
  }
  return n;
}

/* zero-copy view of a completed call's pair list (head first); valid until the batch is cleared */
extern "C" int GmapDP_result_view (const gmapdp_batch *b, int id, const int **iout, const double **dout, const gmapdp_cpair **pairs,
				   const gmapdp_gapinfo **gaps, const gmapdp_pairbase **base, int *dynprogindex) {
  if (id < 0 || id >= (int) b->calls.size()) return -2;
  const Call &c = b->calls[id];
  if (!c.done) return -3;
  if (iout) *iout = c.iout;
  if (dout) *dout = c.dout;
  if (dynprogindex) *dynprogindex = c.dpi;
  if (gaps) *gaps = c.pairs.gaps.data();
  if (base) *base = &c.pairs.base;
  if (c.isnull || c.pairs.empty()) { if (pairs) *pairs = NULL; return -1; }
  if (pairs) *pairs = c.pairs.data();
  return (int) c.pairs.size();
}

extern "C" const gmapdp_result *GmapDP_device_result (const gmapdp_batch *b, int id) {
  if (id < 0 || id >= (int) b->calls.size()) return NULL;
  const Call &c = b->calls[id];
  if (c.box < 0 || (size_t) c.box >= b->nresults) return NULL;
  return &b->results[c.box];
}
