// TEST HARNESS ONLY.  Runs the lane-level product code of t1k_b200/csrc/t1k_core.cuh sequentially on the
// CPU (compiled by g++ with the CUDA qualifiers defined away) so that its logic can be checked against the
// oracle on the GPU-less build box.  The warp orchestration of the real kernels is re-stated here with plain
// loops; nothing in the product links against this file.
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#define T1K_EMU_COUNTERS 1
#include "../t1k_b200/csrc/t1k_core.cuh"
#include "../t1k_b200/csrc/t1k_host.hpp"

long long t1k_emu_counters[128];
using namespace t1k;

struct Emu {
  bool noFast = false;     // run everything through the general hit-list path (A/B check of the diagonal fast path)
  PackedRef P;
  std::vector<int32_t> covDiff, covPoint;
  std::vector<u16> simThr;
  RefView R;
  std::vector<u8> scratch;
};

struct EmuOverlap { int32_t seqIdx, readStart, readEnd, seqStart, seqEnd, strand, matchCnt, relaxedMatchCnt, leftClip, rightClip; };

extern "C" {

long long *emu_counters() { return t1k_emu_counters; }

Emu *emu_create(int32_t n, const char *bases, const int64_t *off, const int32_t *exonPtr, const int32_t *exonSE,
                double sim, int32_t relax) {
  Emu *e = new Emu;
  if (!pack_reference(n, bases, off, exonPtr, exonSE, e->P)) { delete e; return NULL; }
  e->covDiff.assign(e->P.covEntries, 0);
  e->covPoint.assign(e->P.covEntries, 0);
  RefView &R = e->R;
  R.seq2 = e->P.seq2.data(); R.n2 = e->P.n2.data(); R.ex2 = e->P.ex2.data();
  R.wordOff = e->P.wordOff.data(); R.len = e->P.len.data(); R.hasN = e->P.hasN.data(); R.meta = e->P.meta.data();
  e->simThr.resize(2 * SIM_DEN); sim_threshold_table(sim, e->simThr.data()); R.simThr = e->simThr.data();
  R.kstart = e->P.kstart.data(); R.post = e->P.post.data(); R.kinfo = e->P.kinfo.data(); R.entries = e->P.entries.data();
  R.covDiff = e->covDiff.data(); R.covPoint = e->covPoint.data(); R.covOff = e->P.covOff.data();
  R.nAlleles = n; R.sim = sim; R.relax = relax;
  e->scratch.assign(scr_bytes(MAX_READ_LEN), 0);
  return e;
}
void emu_destroy(Emu *e) { delete e; }
void emu_set_fast(Emu *e, int on) { e->noFast = !on; }

// dp_align / diag_certified on a single pair of strings; returns number of ops, fills ops, *certified
int32_t emu_align(const char *t, int32_t lent, const char *p, int32_t lenp, int8_t *opsOut, int32_t *certified,
                  int32_t *matches) {
  // build a one-allele reference holding t and a read holding p
  // (two pad words in front, as in the packed reference: the certificates read a few bases before the window)
  std::vector<u64> seq2((lent + 31) / 32 + 5, 0), n2(seq2.size(), 0), ex2(seq2.size(), 0);
  for (int j = 0; j < lent; ++j) {
    seq2[2 + (j >> 5)] |= (u64)code_of(t[j]) << ((j & 31) * 2);
    if (t[j] == 'N') n2[2 + (j >> 5)] |= 1ull << ((j & 31) * 2);
  }
  u64 w0 = 2; int32_t len = lent;
  u8 tHasN = memchr(t, 'N', lent) != NULL;
  RefView R; memset(&R, 0, sizeof(R));
  R.seq2 = seq2.data(); R.n2 = n2.data(); R.ex2 = ex2.data(); R.wordOff = &w0; R.len = &len; R.hasN = &tHasN; R.nAlleles = 1;
  u64 fs[MAX_RWORDS], fn[MAX_RWORDS], rs[MAX_RWORDS], rn[MAX_RWORDS];
  if (lenp > MAX_READ_LEN) return -2;
  pack_read(p, lenp, fs, fn, rs, rn, MAX_RWORDS);
  ReadView Q; Q.seq2 = fs; Q.n2 = fn; Q.len = lenp; Q.anyN = memchr(p, 'N', lenp) != NULL;
  std::vector<u8> scr(scr_bytes(MAX_READ_LEN), 0);
  const LaneScratch S = lane_scratch(scr.data(), MAX_READ_LEN);
  int err = 0, mm = 0;
  const AlleleView T = allele_view(R, 0, Q);
  *certified = (lent == lenp && lent > 0) ? (int)diag_certified(T, 0, Q, 0, lent, mm) : 0;
  *matches = align_matches(T, 0, lent, Q, 0, lenp, S, err);
  int n = dp_align(T, 0, lent, Q, 0, lenp, S, err);
  if (err) return -1;
  for (int i = 0; i < n; ++i) opsOut[i] = (int8_t)S.ops()[i];
  return n;
}

// SeqSet::AssignRead through the product's lane code.  The warp orchestration of k_assign (t1k_kernels.cuh) is restated
// sequentially: seeds with the skip rule, the sweep over allele tiles of the tile index (KmerEntry) with the streaming
// per-allele counters of the mismatch-mask path, the hit list only for the alleles that path declines.
int32_t emu_assign(Emu *E, const char *read, int32_t weight, EmuOverlap *out, int32_t cap, int32_t *errOut) {
  const RefView &R = E->R;
  int len = (int)strlen(read);
  *errOut = 0;
  if (len < KMER || len > MAX_READ_LEN) return -1;
  u64 planes[4][MAX_RWORDS];
  if (!pack_read(read, len, planes[0], planes[1], planes[2], planes[3], MAX_RWORDS)) { *errOut = -1; return -1; }
  const LaneScratch S = lane_scratch(E->scratch.data(), MAX_READ_LEN);
  int err = 0;
  std::vector<Cand> cands;
  u64 bestKey = 0;
  int nFwd = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int strand01 = pass == 0 ? 1 : 0;
    ReadView Q; Q.seq2 = planes[pass * 2]; Q.n2 = planes[pass * 2 + 1]; Q.len = len; Q.anyN = strchr(read, 'N') != NULL;
    // seed selection (GetHitsFromRead skip rule)
    u32 prev = 0; int skip = 0;
    const int P = len - KMER + 1;
    u16 seedA[1024]; u32 cur[1024], end[1024]; int nS = 0;
    bool strandFast = !Q.anyN && len <= FAST_MAX_LEN && !E->noFast;     // the kernel's eligibility rule (t1k_kernels.cuh)
    for (int a = 0; a < P; ++a) {
      u32 code = (u32)(fetch32(Q.seq2, a) & 0x3FFFFF);
      bool valid = (fetch32(Q.n2, a) & 0x155555) == 0;
      if (a == 0 || prev != code) {
        u32 lo = 0, hi = 0; int size = 0;
        if (valid) { lo = R.kinfo[code].estart; hi = R.kinfo[code + 1].estart; size = (int)(R.kinfo[code + 1].pstart - R.kinfo[code].pstart); }
        if (size >= 100 && a != 0 && a != P - 1 && skip < KMER / 2) { ++skip; continue; }
        skip = 0;
        if (size > 0) { seedA[nS] = (u16)a; cur[nS] = lo; end[nS] = hi; ++nS; if (kmer_homopolymer(code)) strandFast = false; }
      }
      prev = code;
    }
    t1k_emu_counters[52] += 1; t1k_emu_counters[53] += nS;
    u32 stab[256];
    if (strandFast) seed_table_build(seedA, nS, len, stab);
    u32 lcMemo = 0;
    u32 s2[10], lcp[260];
    u64 S2[5];
    if (strandFast) {
      seed_bits2_build(seedA, nS, s2);
      for (int j = 0; j < 5; ++j) S2[j] = (u64)s2[2 * j] | ((u64)s2[2 * j + 1] << 32);
      lc_prefix_build(Q, lcp);
    }
    for (;;) {
      u32 T = 0xffffffffu;
      for (int k = 0; k < nS; ++k) if (cur[k] < end[k]) T = std::min(T, R.entries[cur[k]].tile);
      if (T == 0xffffffffu) break;
      t1k_emu_counters[50] += 1;
      for (int k = 0; k < nS; ++k) if (cur[k] < end[k] && R.entries[cur[k]].tile == T) t1k_emu_counters[51] += 1 + R.entries[cur[k]].more;
      for (int lane = 0; lane < 32; ++lane) {
        // sweep 1: the allele's hit count, first diagonal, hits on / far off that diagonal
        int n = 0, d0 = 0, onDiag = 0, far = 0;
        std::vector<u32> h;
        for (int k = 0; k < nS; ++k) {
          if (cur[k] >= end[k] || R.entries[cur[k]].tile != T) continue;
          const u32 more = R.entries[cur[k]].more;
          for (u32 j = 0; j <= more; ++j) {
            const KmerEntry &e = R.entries[cur[k] + j];
            if (!((e.mask >> lane) & 1u)) continue;
            const int dg = (int)e.off - (int)seedA[k];
            if (n == 0) d0 = dg;
            const int dd = dg - d0;
            ++n; onDiag += dd == 0; far += dd > RADIUS || dd < -RADIUS;
            h.push_back(hit_make((int)seedA[k], e.off));        // (sweep 2 of the kernel: only for declined alleles)
          }
        }
        if (n < 3) continue;
        const int seqIdx = (int)(T * 32 + lane);
        int nEmit = 0;
        t1k_emu_counters[20] += 1;
        if (strandFast) {
          Cand fc; bool emitted = false;
          // the kernel's two-speed protocol: every allele of a tile in hot mode; the deferred ones again, in full mode
          // The kernel's protocol: every allele of a tile through the bit-parallel diag_hot; the deferred ones through the
          // streaming diag_fast (full mode, k_deferred).  A/B: whenever both finish, they must agree field for field.
          u64 keyHot = bestKey, keyFull = bestKey;
          int df = diag_hot(R, Q, strand01, seqIdx, n, d0, onDiag, far, S2, lcp, fc, emitted, keyHot);
          {
            Cand fc2; bool em2 = false; u32 memo2 = 0;
            memset(&fc2, 0, sizeof(fc2));
            const int df2 = diag_fast(R, Q, strand01, seqIdx, n, d0, onDiag, far, stab, false, fc2, em2, keyFull, memo2, S, err);
            t1k_emu_counters[35 + df] += 1;
            if (df == DF_DEFER) {
              t1k_emu_counters[34] += 1;
              if (df2 != DF_DONE) err |= 1 << 20;          // a deferred allele never declines
              df = df2; fc = fc2; emitted = em2; keyHot = keyFull;
            } else if (df == DF_DONE && df2 == DF_DONE) {
              bool same = emitted == em2 && keyHot == keyFull;
              if (same && emitted && !(fc2.flags & CF_PRE)) {
                // the streaming path leaves an overhang longer than one word to ExtendOverlap proper; the bit-parallel path
                // knows that <= 3 mismatches make it the diagonal whatever its length: same result, no mismatch memo
                extend_cand<false>(R, Q, fc2, S, err); fc2.flags |= CF_PRE;
                if (fc.flags & CF_FA) { fc2.flags |= CF_FA; fc2.mmPos = fc.mmPos; }
              }
              if (same && emitted) {
                same = fc.seqIdx == fc2.seqIdx && fc.seqStart == fc2.seqStart && fc.seqEnd == fc2.seqEnd && fc.readStart == fc2.readStart &&
                       fc.readEnd == fc2.readEnd && fc.strand01 == fc2.strand01 && fc.matchCnt == fc2.matchCnt && fc.flags == fc2.flags;
                if (same && (fc.flags & CF_PRE))
                  same = fc.eSeqStart == fc2.eSeqStart && fc.eSeqEnd == fc2.eSeqEnd && fc.eReadStart == fc2.eReadStart && fc.eReadEnd == fc2.eReadEnd &&
                         fc.leftClip == fc2.leftClip && fc.rightClip == fc2.rightClip && fc.eMatchCnt == fc2.eMatchCnt && fc.relaxed == fc2.relaxed &&
                         fc.mmPos == fc2.mmPos;
              }
              if (!same) {
                err |= 1 << 21; t1k_emu_counters[39] += 1;
                if (getenv("EMU_DEBUG"))
                  fprintf(stderr, "A/B allele %d n=%d d=%d strand %d: em %d/%d key %llx/%llx | rs %d/%d re %d/%d mc %d/%d flags %x/%x | eRs %d/%d eRe %d/%d eMc %d/%d relaxed %d/%d mmPos %x/%x lc %d/%d clips %d,%d/%d,%d\n",
                          seqIdx, n, d0, strand01, (int)emitted, (int)em2, (unsigned long long)keyHot, (unsigned long long)keyFull, fc.readStart, fc2.readStart,
                          fc.readEnd, fc2.readEnd, fc.matchCnt, fc2.matchCnt, fc.flags, fc2.flags, fc.eReadStart, fc2.eReadStart, fc.eReadEnd, fc2.eReadEnd,
                          fc.eMatchCnt, fc2.eMatchCnt, fc.relaxed, fc2.relaxed, fc.mmPos, fc2.mmPos, 0, 0, fc.leftClip, fc.rightClip, fc2.leftClip, fc2.rightClip);
              }
            } else if (df == DF_DECLINED && df2 != DF_DECLINED) err |= 1 << 22;   // the hot path declines nothing the streaming path takes
          }
          bestKey = keyHot;
          if (df == DF_DONE) {
            if (emitted) {
              if (!(fc.flags & CF_PRE)) { extend_cand<false>(R, Q, fc, S, err); fc.flags |= CF_PRE; }     // a long or dirty overhang
              cands.push_back(fc);
            }
            continue;
          }
        }
        chain_allele(R, Q, strand01, seqIdx, h.data(), 1, n, S, nEmit, bestKey, err);
        if (nEmit > 1) sort_emitted(S.emit(), nEmit);
        // the kernel extends the seed overlaps of the hit-list path right where they are emitted
        std::vector<Cand> em(S.emit(), S.emit() + nEmit);
        for (int k = 0; k < nEmit; ++k) { Cand c = em[k]; c.mmPos = 0; extend_cand<false>(R, Q, c, S, err); c.flags |= CF_PRE; cands.push_back(c); }
      }
      for (int k = 0; k < nS; ++k) if (cur[k] < end[k] && R.entries[cur[k]].tile == T) cur[k] += 1 + R.entries[cur[k]].more;
    }
    if (pass == 0) nFwd = (int)cands.size();
  }
  int best01 = (bestKey & 1) ? 0 : 1;
  int c0 = best01 ? 0 : nFwd, c1 = best01 ? nFwd : (int)cands.size();
  if (c1 - c0 <= 0) { *errOut = err; return -1; }
  ReadView Q; Q.seq2 = planes[best01 ? 0 : 2]; Q.n2 = planes[best01 ? 1 : 3]; Q.len = len; Q.anyN = strchr(read, 'N') != NULL;
  // pass 1: extension + first failing key
  u64 fKey = ~0ull; int fIdx = 0x7fffffff;
  for (int i = c0; i < c1; ++i) {
    Cand &c = cands[i];
    if (!(c.flags & CF_SEP) && !(c.flags & CF_RET)) {
      u64 k = cand_key_pre(c);
      if (k < fKey || (k == fKey && i < fIdx)) { fKey = k; fIdx = i; }
    }
  }
  int good = -1;
  for (int i = c0; i < c1; ++i) {
    Cand &c = cands[i];
    if ((c.flags & CF_SEP) || !(c.flags & CF_RET)) continue;
    u64 k = cand_key_pre(c);
    if (k < fKey || (k == fKey && i < fIdx)) good = std::max(good, (int)c.matchCnt);
  }
  int bestMc = -1, nInc = 0;
  for (int i = c0; i < c1; ++i) {
    Cand &c = cands[i];
    if ((c.flags & CF_SEP) || !(c.flags & CF_RET)) continue;
    u64 k = cand_key_pre(c);
    bool before = k < fKey || (k == fKey && i < fIdx);
    double sim = (double)c.matchCnt / (double)cand_denom_pre(c);
    if (!before && (int)c.matchCnt < good && (!(c.flags & CF_NEEDCLIP) || sim < 0.95)) continue;
    c.flags |= CF_INCLUDE;
    bestMc = std::max(bestMc, c.eMatchCnt);
    ++nInc;
  }
  for (int i = c0; i < c1; ++i) {
    Cand &c = cands[i];
    if (!(c.flags & CF_INCLUDE)) continue;
    if (weight >= 0) {
      if (c.eMatchCnt < bestMc - 10) c.relaxed = 0;
      else if (c.flags & CF_FA) full_align_known(R, c, weight);
      else full_align<false>(R, Q, c, weight, S, err);
    }
  }
  bool usePost = nInc > 1000;
  if (usePost) {
    u64 bKey = ~0ull; int bIdx = 0x7fffffff;
    for (int i = c0; i < c1; ++i) if (cands[i].flags & CF_INCLUDE) {
      u64 k = cand_key_post(cands[i]);
      if (k < bKey || (k == bKey && i < bIdx)) { bKey = k; bIdx = i; }
    }
    double bestSim = (double)cands[bIdx].eMatchCnt / (double)cand_denom_post(cands[bIdx]);
    u64 cKey = ~0ull; int cIdx = 0x7fffffff;
    for (int i = c0; i < c1; ++i) if ((cands[i].flags & CF_INCLUDE) && i != bIdx) {
      double sim = (double)cands[i].eMatchCnt / (double)cand_denom_post(cands[i]);
      if (sim < bestSim - 0.1) {
        u64 k = cand_key_post(cands[i]);
        if (k < cKey || (k == cKey && i < cIdx)) { cKey = k; cIdx = i; }
      }
    }
    for (int i = c0; i < c1; ++i) if (cands[i].flags & CF_INCLUDE) {
      u64 k = cand_key_post(cands[i]);
      if (k > cKey || (k == cKey && i >= cIdx)) cands[i].flags &= ~CF_INCLUDE;
    }
  }
  // the records as the kernel's compaction pass stores them (candidate order) and the order t1k_assignment_fetch returns
  std::vector<Rec> recs; std::vector<int> recCand;
  for (int i = c0; i < c1; ++i) if (cands[i].flags & CF_INCLUDE) {
    const Cand &c = cands[i];
    Rec o;
    o.seqIdx = c.seqIdx; o.seqStart = c.eSeqStart; o.seqEnd = c.eSeqEnd;
    o.readStart = c.eReadStart; o.readEnd = c.eReadEnd; o.leftClip = c.leftClip; o.rightClip = c.rightClip;
    o.mcx = rec_mcx(c.eMatchCnt, c.relaxed, c.strand01);
    o.key = usePost ? (cand_key_post(c) | 1ull) : cand_key_pre(c);
    recs.push_back(o); recCand.push_back(i);
  }
  std::vector<std::pair<u64, int> > order;
  {
    std::vector<int> idx(recs.size());
    for (size_t k = 0; k < idx.size(); ++k) idx[k] = (int)k;
    std::sort(idx.begin(), idx.end(), [&](int x, int y) { return rec_before(recs[x], x, recs[y], y); });
    for (size_t k = 0; k < idx.size(); ++k) order.push_back(std::make_pair(recs[idx[k]].key, recCand[idx[k]]));
  }
  int w = 0;
  for (size_t k = 0; k < order.size() && w < cap; ++k, ++w) {
    const Cand &c = cands[order[k].second];
    EmuOverlap &o = out[w];
    o.seqIdx = c.seqIdx; o.readStart = c.eReadStart; o.readEnd = c.eReadEnd; o.seqStart = c.eSeqStart; o.seqEnd = c.eSeqEnd;
    o.strand = c.strand01 ? 1 : -1; o.matchCnt = c.eMatchCnt; o.relaxedMatchCnt = c.relaxed;
    o.leftClip = c.leftClip; o.rightClip = c.rightClip;
  }
  *errOut = err;
  return (int)order.size();
}

// SeqSet::AddOverlapAlignmentInfo through the product's lane code (align_info): edit string of one record of emu_assign.
// Returns the number of ops (< 0: error); *ranDp = 1 when the band DP ran (0: certified diagonal).
int32_t emu_align_info(Emu *E, const char *read, const EmuOverlap *o, int32_t noDiag, int8_t *opsOut, int32_t *ranDpOut) {
  int len = (int)strlen(read);
  if (len > MAX_READ_LEN || o->seqIdx < 0) return -2;
  u64 planes[4][MAX_RWORDS];
  if (!pack_read(read, len, planes[0], planes[1], planes[2], planes[3], MAX_RWORDS)) return -2;
  const int pass = o->strand == 1 ? 0 : 1;
  ReadView Q; Q.seq2 = planes[pass * 2]; Q.n2 = planes[pass * 2 + 1]; Q.len = len; Q.anyN = strchr(read, 'N') != NULL;
  const LaneScratch S = lane_scratch(E->scratch.data(), MAX_READ_LEN);
  const AlleleView T = allele_view(E->R, o->seqIdx, Q);
  int err = 0; bool ranDp = false;
  const int n = align_info(T, o->seqStart, o->seqEnd - o->seqStart + 1, Q, o->readStart, o->readEnd - o->readStart + 1, S, err, noDiag != 0, ranDp);
  *ranDpOut = ranDp ? 1 : 0;
  if (err || n < 0) return -1;
  for (int i = 0; i < n; ++i) opsOut[i] = (int8_t)S.ops()[i];
  return n;
}

// coverage of one allele = prefix(covDiff) + covPoint
void emu_coverage(Emu *E, int32_t allele, int32_t *out) {
  size_t cb = (size_t)E->P.covOff[allele];
  int run = 0;
  for (int j = 0; j < E->P.len[allele]; ++j) { run += E->covDiff[cb + (size_t)COV_STRIDE * j]; out[j] = run + E->covPoint[cb + (size_t)COV_STRIDE * j]; }
}

}  // extern "C"
