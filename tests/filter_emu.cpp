// TEST HARNESS ONLY.  Runs the lane code of the candidate filter (t1k_b200/csrc/t1k_filter_lane.cuh)
// sequentially on the CPU: index build with a runtime k (KmerIndex::BuildIndexFromRead order and quirk), seeds of both
// strands, per-(strand, sequence) buckets, the best bucket, its chaining, the verdict.  The orchestration the future
// kernel will do warp-wide is plain loops here.
#include <map>
#include <string>
#include <vector>

#include "../t1k_b200/csrc/t1k_filter_lane.cuh"
#include "../t1k_b200/csrc/t1k_host.hpp"

using namespace t1k;

struct FEmu {
  std::vector<u32> kstart;
  std::vector<Posting> post;
  int k, hitLenReq;
  double sim;
};

extern "C" {

FEmu *femu_create(int32_t n, const char *bases, const int64_t *off, int32_t k, int32_t hitLenReq, double sim) {
  if (k < 1 || k > 15) return NULL;
  FEmu *E = new FEmu;
  E->k = k; E->hitLenReq = hitLenReq; E->sim = sim;
  const size_t nK = (size_t)1 << (2 * k);
  std::vector<u32> cnt(nK + 1, 0);
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = 0; i < n; ++i) {
      const char *s = bases + off[i];
      const int len = (int)(off[i + 1] - off[i]);
      if (len < k) continue;
      u32 code = 0, prev = 0; int bad = -1;
      const u32 mask = (u32)(nK - 1);
      for (int j = 0; j < len; ++j) {
        if (bad != -1) ++bad;
        code = (code >> 2) | ((u32)code_of(s[j]) << (2 * (k - 1)));      // first base of the window in the low bits
        if (s[j] == 'N') bad = 0;
        if (bad >= k) bad = -1;
        if (j < k - 1) continue;
        code &= mask;
        if (bad == -1 && (j == k || code != prev)) {                      // the i == kl quirk (KmerIndex.hpp:121, Q1)
          if (pass == 0) ++cnt[code + 1];
          else { Posting p; p.idx = (u32)i; p.off = (u32)(j - k + 1); E->post[cnt[code]++] = p; }
        }
        prev = code;
      }
    }
    if (pass == 0) {
      for (size_t c = 0; c < nK; ++c) cnt[c + 1] += cnt[c];
      E->kstart = cnt;
      E->post.resize(cnt[nK] + 1);
    }
  }
  return E;
}
void femu_destroy(FEmu *E) { delete E; }

// IsGoodCandidate (FastqExtractor.cpp:114-119); -1: read not usable by the packed path (invalid character, > 1000 bases)
int32_t femu_good_candidate(FEmu *E, const char *read) {
  const int len = (int)strlen(read);
  if (len > MAX_READ_LEN) return -1;
  u64 fs[MAX_RWORDS], fn[MAX_RWORDS], rs[MAX_RWORDS], rn[MAX_RWORDS];
  if (!pack_read(read, len, fs, fn, rs, rn, MAX_RWORDS)) return -1;
  if (filt::read_low_complexity(fs, fn, len)) return 0;
  if (len < E->k) return 0;
  filt::IndexView I;
  I.kstart = E->kstart.data(); I.post = E->post.data(); I.k = E->k; I.hitLenReq = E->hitLenReq;
  std::map<std::pair<int, u32>, std::vector<u32> > buckets;      // (tag: 0 = reverse strand first, sequence) -> hits in arrival order
  u32 prev = 0;
  for (int pass = 0; pass < 2; ++pass) {
    u16 seedA[1024]; u32 lo[1024], hi[1024];
    const int nS = filt::seed_list(I, pass == 0 ? fs : rs, pass == 0 ? fn : rn, len, prev, seedA, lo, hi);
    for (int s = 0; s < nS; ++s)
      for (u32 j = lo[s]; j < hi[s]; ++j)
        buckets[std::make_pair(pass == 0 ? 1 : 0, E->post[j].idx)].push_back(hit_make((int)seedA[s], E->post[j].off));
  }
  int mx = -1;
  std::vector<u32> *best = NULL;
  for (std::map<std::pair<int, u32>, std::vector<u32> >::iterator it = buckets.begin(); it != buckets.end(); ++it)
    if ((int)it->second.size() > mx) { mx = (int)it->second.size(); best = &it->second; }
  if (!best || E->k * mx < E->hitLenReq) return 0;
  std::vector<u8> scratch((size_t)12 * best->size() + filt::FILTER_USED_BYTES);
  const int bestLen = filt::bucket_best_hit_len(best->data(), (int)best->size(), E->k, E->hitLenReq, scratch.data());
  return filt::hit_length_passes(len, bestLen, E->k, E->sim) ? 1 : 0;
}

}  // extern "C"
