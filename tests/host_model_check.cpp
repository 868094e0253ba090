// TEST HARNESS ONLY.  Exercises the threaded parts of the product's host model (t1k_b200/csrc/t1k_model.hpp: sharded
// coalescing + gather, CSR transposition, equivalence classes, EM inputs, table serialisation / parallel merge) on
// random fragment rows and checks that every thread count gives the single-threaded result, byte for byte.
// Returns 0 when everything agrees; a non-zero code names the first stage that differs.
#include <stdio.h>
#include <stdlib.h>
#include <random>
#include <string>

#include "../t1k_b200/csrc/t1k_model.hpp"

using namespace t1k;

static uint64_t row_hash0(const HostEntry *row, uint32_t n) { return ReadGroups::hash_row(row, n); }

template <class A, class B> static bool same_vec(const A &a, const B &b) {
  return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(a[0])) == 0);
}
static bool same_groups(const ReadGroups &a, const ReadGroups &b) {
  return same_vec(a.ptr, b.ptr) && same_vec(a.ent, b.ent) && same_vec(a.first, b.first) && same_vec(a.hashes, b.hashes) &&
         a.assignedFragments == b.assignedFragments;
}

// unique_read_ends: for every thread count the mapping reproduces each read-end, the weights are the duplicate counts,
// the N flags are right and the number of unique read-ends is the same
extern "C" int dedup_check(int seed, int nFrag, int paired) {
  std::mt19937_64 rng((uint64_t)seed);
  const uint32_t stride = 40;
  const char alpha[5] = {'A', 'C', 'G', 'T', 'N'};
  std::vector<std::string> pool;
  for (int i = 0; i < std::max(3, nFrag / 6); ++i) {
    std::string r;
    const int L = 5 + (int)(rng() % 34);
    for (int j = 0; j < L; ++j) r.push_back(alpha[rng() % (rng() % 11 == 0 ? 5 : 4)]);
    pool.push_back(r);
  }
  std::vector<char> r1((size_t)nFrag * stride, 0), r2((size_t)nFrag * stride, 0);
  for (int f = 0; f < nFrag; ++f) {
    const std::string &a = pool[rng() % pool.size()], &b = pool[rng() % pool.size()];
    memcpy(r1.data() + (size_t)f * stride, a.data(), a.size());
    memcpy(r2.data() + (size_t)f * stride, b.data(), b.size());
  }
  size_t nUnique = 0;
  for (int T = 1; T <= 8; T += (T == 1 ? 2 : 5)) {
    for (int half = 0; half < 2; ++half) {            // a chunk that starts in the middle as well
      const uint32_t f0 = half ? (uint32_t)nFrag / 3 : 0, m = (uint32_t)nFrag - f0;
      ReadEndChunk C;
      unique_read_ends(r1.data(), paired ? r2.data() : nullptr, stride, f0, m, 255, T, C);
      std::vector<int32_t> seen(C.w.size(), 0);
      for (uint32_t i = 0; i < m; ++i) {
        bool hasN = false;
        for (int mate = 0; mate < (paired ? 2 : 1); ++mate) {
          const char *s = (mate ? r2 : r1).data() + (size_t)(f0 + i) * stride;
          const uint32_t u = (mate ? C.e2 : C.e1)[i];
          if (u >= C.w.size()) return 20;
          const size_t L = strlen(s);
          if (C.len[u] != L || memcmp(C.bases.data() + C.off[u], s, L) != 0) return 21;
          ++seen[u];
          hasN |= memchr(s, 'N', L) != nullptr;
        }
        if ((C.hasN[i] != 0) != hasN) return 22;
      }
      for (size_t u = 0; u < seen.size(); ++u) if (seen[u] != C.w[u] || seen[u] == 0) return 23;
      // no two unique entries hold the same sequence
      std::vector<std::string> all;
      for (size_t u = 0; u < C.w.size(); ++u) all.push_back(std::string(C.bases.data() + C.off[u], C.len[u]));
      std::sort(all.begin(), all.end());
      if (std::adjacent_find(all.begin(), all.end()) != all.end()) return 24;
      if (!half) { if (T == 1) nUnique = all.size(); else if (all.size() != nUnique) return 25; }
    }
  }
  return 0;
}

extern "C" int host_model_check(int seed, int nFrag, int nAlleles, int nSets) {
  setenv("T1K_PAR_MIN", "1", 1);                 // the threaded paths on small inputs too
  std::mt19937_64 rng((uint64_t)seed);
  // a pool of allele sets (sorted), fragments draw from it: many fragments share a set -> groups with many members
  std::vector<std::vector<int32_t> > sets((size_t)nSets);
  for (int s = 0; s < nSets; ++s) {
    const int n = 1 + (int)(rng() % 40);
    std::vector<int32_t> v;
    const int base = (int)(rng() % (uint64_t)nAlleles);
    for (int i = 0; i < n; ++i) v.push_back((base + (int)(rng() % 97)) % nAlleles);
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    sets[s] = v;
  }
  std::vector<HostEntry> ent;
  std::vector<uint64_t> off((size_t)nFrag), hash((size_t)2 * nFrag);
  std::vector<uint32_t> cnt((size_t)nFrag);
  for (int f = 0; f < nFrag; ++f) {
    off[f] = ent.size();
    if (rng() % 9 == 0) { cnt[f] = 0; continue; }          // unassigned fragment
    const std::vector<int32_t> &v = sets[rng() % (uint64_t)nSets];
    for (size_t i = 0; i < v.size(); ++i) {
      HostEntry e;
      e.alleleIdx = v[i]; e.start = (int32_t)(rng() % 900); e.end = e.start + 150;
      e.weight = (float)(1 + rng() % 3) * 0.5f; e.qual = 1.0f; e.adjustWeight = e.weight * 0.25f;
      ent.push_back(e);
    }
    cnt[f] = (uint32_t)v.size();
    hash[2 * (size_t)f] = row_hash0(ent.data() + off[f], cnt[f]); hash[2 * (size_t)f + 1] = 0;
  }
  // ---- sharded coalescing + gather: T threads == 1 thread
  ReadGroups ref;
  {
    GroupShards S(1);
    S.add_chunk(ent.data(), off.data(), cnt.data(), hash.data(), (uint32_t)nFrag / 2, 0);
    S.add_chunk(ent.data(), off.data() + nFrag / 2, cnt.data() + nFrag / 2, hash.data() + 2 * (size_t)(nFrag / 2), (uint32_t)(nFrag - nFrag / 2), nFrag / 2);
    S.gather(ref);
  }
  for (int T = 2; T <= 8; T += 3) {
    GroupShards S(T);
    S.add_chunk(ent.data(), off.data(), cnt.data(), hash.data(), (uint32_t)nFrag / 2, 0);
    S.add_chunk(ent.data(), off.data() + nFrag / 2, cnt.data() + nFrag / 2, hash.data() + 2 * (size_t)(nFrag / 2), (uint32_t)(nFrag - nFrag / 2), nFrag / 2);
    ReadGroups g;
    S.gather(g);
    if (!same_groups(g, ref)) return 1;
  }
  // ---- the plain serial coalescing (ReadGroups::add in fragment order) gives the same groups, in the same order
  {
    ReadGroups g;
    for (int f = 0; f < nFrag; ++f) if (cnt[f]) { const uint64_t h = hash[2 * (size_t)f]; g.add(ent.data() + off[f], cnt[f], 1, &h, f); }
    if (!same_vec(g.ptr, ref.ptr) || !same_vec(g.ent, ref.ent)) return 2;
  }
  // ---- transposition, equivalence classes, EM inputs: threads == 1 thread
  std::vector<int32_t> effLen((size_t)nAlleles), seqW((size_t)nAlleles);
  for (int a = 0; a < nAlleles; ++a) { effLen[a] = 900 + (int)(rng() % 300); seqW[a] = 1 + (int)(rng() % 3); }
  EquivalenceClasses ec1;
  ec1.build(ref, nAlleles, 1);
  EmInputs in1;
  in1.build(ref, ec1, effLen.data(), seqW.data(), 1);
  std::vector<int64_t> cp1; std::vector<int32_t> ri1;
  const int32_t *colp = in1.col.data();
  transpose_csr(in1.rowPtr.data(), ref.size(), [colp](int64_t k) { return colp[k]; }, ec1.size(), 1, cp1, ri1);
  for (int T = 2; T <= 8; T += 3) {
    EquivalenceClasses ec;
    ec.build(ref, nAlleles, T);
    if (!same_vec(ec.ecPtr, ec1.ecPtr) || !same_vec(ec.ecAlleles, ec1.ecAlleles) || !same_vec(ec.alleleEc, ec1.alleleEc)) return 3;
    EmInputs in;
    in.build(ref, ec, effLen.data(), seqW.data(), T);
    if (!same_vec(in.rowPtr, in1.rowPtr) || !same_vec(in.col, in1.col) || !same_vec(in.count, in1.count) || !same_vec(in.ecLen, in1.ecLen) ||
        !same_vec(in.x0, in1.x0)) return 4;
    std::vector<int64_t> cp; std::vector<int32_t> ri;
    transpose_csr(in1.rowPtr.data(), ref.size(), [colp](int64_t k) { return colp[k]; }, ec1.size(), T, cp, ri);
    if (!same_vec(cp, cp1) || !same_vec(ri, ri1)) return 5;
  }
  // ---- every column of the transposition lists its rows in ascending order (the EM's fixed summation order)
  for (int32_t e = 0; e < ec1.size(); ++e)
    for (int64_t k = cp1[e] + 1; k < cp1[e + 1]; ++k) if (ri1[k - 1] >= ri1[k]) return 6;
  // ---- read-sharded merge: two "ranks" = the two halves of the fragments; serialise, merge in rank order on T threads
  {
    ReadGroups half[2];
    for (int r = 0; r < 2; ++r) {
      GroupShards S(3);
      const int f0 = r ? nFrag / 2 : 0, m = r ? nFrag - nFrag / 2 : nFrag / 2;
      S.add_chunk(ent.data(), off.data() + f0, cnt.data() + f0, hash.data() + 2 * (size_t)f0, (uint32_t)m, 0);     // rank-local fragment numbers
      S.gather(half[r]);
    }
    std::vector<std::vector<uint8_t> > blobs(2);
    std::vector<GroupBlobView> views(2);
    for (int r = 0; r < 2; ++r) { serialize_groups(half[r], blobs[r]); if (!views[r].parse(blobs[r].data(), blobs[r].size())) return 7; }
    const std::vector<int64_t> fragBase = {0, nFrag / 2};
    ReadGroups m1;
    if (!merge_tables_parallel(views, fragBase, 1, m1)) return 8;
    for (int T = 2; T <= 8; T += 3) {
      ReadGroups m;
      if (!merge_tables_parallel(views, fragBase, T, m)) return 8;
      if (!same_groups(m, m1)) return 9;
    }
    // ---- rank-partitioned merge (T1K_MERGE_PARTITIONED): three "ranks", every rank merges its hash partitions, the
    // partitions are serialised, exchanged and interleaved; equal to the full merge of the same three tables
    {
      const int W = 3;
      const int cut[W + 1] = {0, nFrag / 4, nFrag / 4 + nFrag / 3, nFrag};
      ReadGroups shard[W];
      std::vector<std::vector<uint8_t> > b3(W);
      std::vector<GroupBlobView> v3(W);
      std::vector<int64_t> fb3(W);
      for (int r = 0; r < W; ++r) {
        GroupShards S(2);
        S.add_chunk(ent.data(), off.data() + cut[r], cnt.data() + cut[r], hash.data() + 2 * (size_t)cut[r], (uint32_t)(cut[r + 1] - cut[r]), 0);
        S.gather(shard[r]);
        serialize_groups(shard[r], b3[r]);
        if (!v3[r].parse(b3[r].data(), b3[r].size())) return 12;
        fb3[r] = cut[r];
      }
      ReadGroups full;
      if (!merge_tables_parallel(v3, fb3, 4, full)) return 13;
      for (int T = 1; T <= 5; T += 2) {
        std::vector<std::vector<uint8_t> > pb(W);
        std::vector<GroupBlobView> pv(W);
        for (int r = 0; r < W; ++r) {
          ReadGroups mine;
          if (!merge_tables_partition(v3, fb3, r, W, T, mine)) return 14;
          serialize_groups(mine, pb[r]);
          if (!pv[r].parse(pb[r].data(), pb[r].size())) return 15;
        }
        ReadGroups asm_;
        if (!assemble_partitions(pv, T, asm_)) return 16;
        asm_.assignedFragments = full.assignedFragments;
        if (!same_groups(asm_, full)) return 17;
        // the product's exchange: every rank splits its table by owner (plan_partitions / serialize_partitions), rank r gets
        // only the blobs addressed to it (the all-to-all), merges them, and the merged partitions assemble to the same table
        std::vector<PartitionPlan> plans(W);
        std::vector<std::vector<uint8_t> > sendBuf(W);
        for (int s = 0; s < W; ++s) {
          plan_partitions(shard[s], W, T, plans[s]);
          sendBuf[s].assign(plans[s].total + 16, 0);
          serialize_partitions(shard[s], plans[s], sendBuf[s].data(), T);
        }
        std::vector<std::vector<uint8_t> > pb2(W);
        std::vector<GroupBlobView> pv2(W);
        for (int r = 0; r < W; ++r) {
          std::vector<GroupBlobView> in(W);
          for (int s = 0; s < W; ++s) {
            size_t at = 0;
            for (int q = 0; q < r; ++q) at += plans[s].bytes[q];
            if (!in[s].parse(sendBuf[s].data() + at, plans[s].bytes[r])) return 18;
          }
          ReadGroups mine;
          if (!merge_tables_partition(in, fb3, r, W, T, mine)) return 19;
          serialize_groups(mine, pb2[r]);
          if (!pv2[r].parse(pb2[r].data(), pb2[r].size())) return 20;
        }
        ReadGroups asm2;
        if (!assemble_partitions(pv2, T, asm2)) return 21;
        asm2.assignedFragments = full.assignedFragments;
        if (!same_groups(asm2, full)) return 22;
        // compact form of the second exchange: the host assembly, and the heads-only assembly of the device path (the allele
        // runs gathered from the blobs by their byte offsets, as k_tail_gather does) give the full table's ids / counts
        {
          std::vector<ReadGroups> mine(W);
          std::vector<std::vector<uint8_t> > cb(W);
          std::vector<const uint8_t *> blobs(W), heads(W);
          std::vector<uint64_t> bytes(W), headBytes(W), base(W);
          std::vector<uint8_t> all;
          for (int r = 0; r < W; ++r) {
            std::vector<GroupBlobView> in(W);
            for (int q = 0; q < W; ++q) {
              size_t at = 0;
              for (int z = 0; z < r; ++z) at += plans[q].bytes[z];
              if (!in[q].parse(sendBuf[q].data() + at, plans[q].bytes[r])) return 30;
            }
            if (!merge_tables_partition(in, fb3, r, W, T, mine[r])) return 31;
            cb[r].assign(compact_group_bytes(mine[r]), 0);
            serialize_compact(mine[r], cb[r].data(), T);
            base[r] = all.size();
            all.insert(all.end(), cb[r].begin(), cb[r].end());
            bytes[r] = cb[r].size();
            headBytes[r] = 16 + ((uint64_t)mine[r].size() + 1) * 8 + (uint64_t)mine[r].size() * 16;
          }
          for (int r = 0; r < W; ++r) { blobs[r] = cb[r].data(); heads[r] = all.data() + base[r]; }
          CompactGroups c1, c2;
          std::vector<int64_t> srcOff;
          if (!assemble_compact(blobs, bytes, T, c1)) return 32;
          if (!assemble_compact_heads(heads, headBytes, bytes, base, c2, srcOff)) return 33;
          if (!same_vec(c1.ptr, full.ptr) || !same_vec(c2.ptr, full.ptr) || !same_vec(c1.count, c2.count)) return 34;
          const GroupsView fv = view_of(full);
          for (int32_t g = 0; g < full.size(); ++g) {
            if (c1.count[g] != fv.count_of(g)) return 35;
            for (int64_t k = full.ptr[g]; k < full.ptr[g + 1]; ++k) {
              int32_t fromBlob;
              memcpy(&fromBlob, all.data() + srcOff[g] + (size_t)(k - full.ptr[g]) * 4, 4);
              if (c1.allele[k] != full.ent[k].alleleIdx || fromBlob != full.ent[k].alleleIdx) return 36;
            }
          }
        }
        // covered ranges of the alleles (N4): threaded == serial
        {
          int32_t nA = 0;
          for (size_t k = 0; k < full.ent.size(); ++k) nA = std::max(nA, full.ent[k].alleleIdx + 1);
          std::vector<int32_t> s1, sT;
          allele_spans(full, nA, 1, s1);
          allele_spans(full, nA, T + 2, sT);
          if (!same_vec(s1, sT)) return 37;
        }
      }
    }
    // the merged table has the single-process groups in the single-process order (float32 sums may differ in the last
    // bit: per-rank partial sums), so compare structure and allele ids
    if (!same_vec(m1.ptr, ref.ptr) || m1.assignedFragments != ref.assignedFragments) return 10;
    for (size_t k = 0; k < m1.ent.size(); ++k) if (m1.ent[k].alleleIdx != ref.ent[k].alleleIdx) return 11;
  }
  return 0;
}
