// Host side of the genotyping model between the alignment kernels and the EM kernels: the order-sensitive,
// serial bookkeeping of Genotyper.hpp that SURVEY.md §8b keeps on the host (read-group coalescing, allele
// equivalence classes, EM problem assembly, allele abundances).  Plain C++, no CUDA.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <thread>
#include <unordered_map>
#include <vector>

namespace t1k {

// vector storage that is NOT zero-filled on resize: the big entry arrays are sized once and then filled (and first
// touched) by several threads in parallel
template <class T> struct NoInitAlloc : std::allocator<T> {
  template <class U> struct rebind { typedef NoInitAlloc<U> other; };
  NoInitAlloc() {}
  template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
  template <class U, class... A> void construct(U *p, A &&...a) {
    if (sizeof...(A) == 0) ::new ((void *)p) U; else ::new ((void *)p) U(static_cast<A &&>(a)...);
  }
};

template <class F> inline void run_threads(int n, F fn);

struct HostEntry {   // == PairEntry / T1KReadAssignment
  int32_t alleleIdx, start, end;
  float weight, qual, adjustWeight;
};

// Genotyper::CoalesceReadAssignments (Genotyper.hpp:841-908): fragments with the same allele set (and qual, which
// is always 1 on this path, SeqSet.hpp:2507,2514) merge into one read group; float32 weights accumulate in
// fragment order; start/end follow the reference's update rule verbatim (including end <- start, :893-894).
struct ReadGroups {
  std::vector<int64_t> ptr{0};
  std::vector<HostEntry, NoInitAlloc<HostEntry> > ent;
  std::vector<int64_t> first;          // per group: index of the fragment that created it (orders merged shards)
  std::vector<uint64_t> hashes;        // per group: the allele-set hash it is filed under
  std::unordered_map<uint64_t, std::vector<int32_t> > byHash;
  int64_t assignedFragments = 0;

  int32_t size() const { return (int32_t)ptr.size() - 1; }

  static uint64_t hash_row(const HostEntry *row, uint32_t n) {
    uint64_t h = 0x9e3779b97f4a7c15ull ^ n;
    for (uint32_t i = 0; i < n; ++i) {
      h ^= (uint64_t)(uint32_t)row[i].alleleIdx + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
      h *= 0xff51afd7ed558ccdull;
    }
    return h;
  }

  // row must be sorted by alleleIdx (the pairing kernel emits it that way).  `fragments` = how many fragments the
  // row stands for: 1 for a fragment, the group's own count when another rank's table is merged in.
  // `hash`: a precomputed hash of the allele set (the pairing kernel's), else computed here; `fragIdx`: global index
  // of the fragment (first-appearance order of the groups).
  void add(const HostEntry *row, uint32_t n, int64_t fragments = 1, const uint64_t *hash = nullptr, int64_t fragIdx = -1) {
    if (n == 0) return;
    assignedFragments += fragments;
    const uint64_t h = hash ? *hash : hash_row(row, n);
    std::vector<int32_t> &cand = byHash[h];
    for (size_t c = 0; c < cand.size(); ++c) {
      const int32_t g = cand[c];
      if (ptr[g + 1] - ptr[g] != (int64_t)n) continue;
      HostEntry *t = &ent[ptr[g]];
      bool same = true;
      for (uint32_t j = 0; j < n && same; ++j) same = t[j].alleleIdx == row[j].alleleIdx && t[j].qual == row[j].qual;
      if (!same) continue;
      for (uint32_t j = 0; j < n; ++j) {
        if (row[j].qual == 1) {
          if (row[j].start < t[j].start) t[j].start = row[j].start;
          if (row[j].end < t[j].end) t[j].end = row[j].start;
        }
        t[j].weight += row[j].weight;
        t[j].adjustWeight += row[j].adjustWeight;
      }
      return;
    }
    cand.push_back(size());
    ent.insert(ent.end(), row, row + n);
    ptr.push_back((int64_t)ent.size());
    first.push_back(fragIdx);
    hashes.push_back(h);
  }
};

// What the global tail (equivalence classes, EM inputs) reads of the group table: the allele ids of every group and the group's
// count (largest summed weight, Genotyper.hpp:1159-1163).  Either a view of a full table, or the compact form the ranks of a
// read-sharded run exchange (4 bytes per entry instead of 24).
struct GroupsView {
  int32_t n = 0;
  const int64_t *ptr = nullptr;
  const HostEntry *ent = nullptr;      // full table ...
  const int32_t *allele = nullptr;     // ... or compact: allele ids
  const double *count = nullptr;       //                  and counts
  const int32_t *dAllele = nullptr;    // ... or compact with the allele ids resident on the DEVICE (allele == NULL then; host code
                                       // that needs the ids falls back to fetching them)
  int64_t entries() const { return n > 0 ? ptr[n] : 0; }
  int32_t allele_at(int64_t k) const { return ent ? ent[k].alleleIdx : allele[k]; }
  double count_of(int32_t g) const {
    if (!ent) return count[g];
    float c = ent[ptr[g]].weight;
    for (int64_t k = ptr[g] + 1; k < ptr[g + 1]; ++k) if (ent[k].weight > c) c = ent[k].weight;
    return c;
  }
};
inline GroupsView view_of(const ReadGroups &G) { GroupsView v; v.n = G.size(); v.ptr = G.ptr.data(); v.ent = G.ent.data(); return v; }

// Coalescing on T host threads.  Thread t owns the read groups whose allele-set hash falls into partition t, so every
// group still sees its fragments in fragment order (float32 sums unchanged) while the partitions proceed in parallel;
// gather() interleaves the partitions by the fragment that created each group = the single-threaded group order.
struct GroupShards {
  std::vector<ReadGroups> part;
  explicit GroupShards(int T) : part((size_t)(T < 1 ? 1 : T)) {}
  int threads() const { return (int)part.size(); }

  // rows: entries at ent + off[i], cnt[i] of them, hash[2*i] = allele-set hash, fragments f0 .. f0+m-1
  void add_chunk(const HostEntry *ent, const uint64_t *off, const uint32_t *cnt, const uint64_t *hash, uint32_t m, int64_t f0);

  // partitions interleaved by the fragment that created each group; the entry copy (hundreds of MB) runs on the
  // partitions' threads, every thread touching its own range of the destination first
  void gather(ReadGroups &out) const {
    struct Ref { int64_t first; int32_t t, g; };
    std::vector<Ref> order;
    for (int t = 0; t < threads(); ++t) {
      for (int32_t g = 0; g < part[t].size(); ++g) order.push_back(Ref{part[t].first[g], t, g});
      out.assignedFragments += part[t].assignedFragments;
    }
    std::sort(order.begin(), order.end(), [](const Ref &a, const Ref &b) { return a.first < b.first; });
    const size_t g0 = out.first.size(), nG = order.size();
    out.ptr.resize(g0 + nG + 1); out.first.resize(g0 + nG); out.hashes.resize(g0 + nG);
    for (size_t k = 0; k < nG; ++k) {
      const ReadGroups &P = part[order[k].t];
      const int32_t g = order[k].g;
      out.ptr[g0 + k + 1] = out.ptr[g0 + k] + (P.ptr[g + 1] - P.ptr[g]);
      out.first[g0 + k] = P.first[g];
      out.hashes[g0 + k] = P.hashes[g];
    }
    out.ent.resize((size_t)out.ptr[g0 + nG]);
    const int T = nG < 1024 ? 1 : threads();
    auto copy = [&](int t) {
      for (size_t k = nG * t / T; k < nG * (t + 1) / T; ++k) {
        const ReadGroups &P = part[order[k].t];
        const int32_t g = order[k].g;
        const int64_t n = P.ptr[g + 1] - P.ptr[g];
        if (n) memcpy(out.ent.data() + out.ptr[g0 + k], P.ent.data() + P.ptr[g], (size_t)n * sizeof(HostEntry));
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(copy, t);
    copy(0);
    for (size_t k = 0; k < th.size(); ++k) th[k].join();
  }
};

inline void GroupShards::add_chunk(const HostEntry *ent, const uint64_t *off, const uint32_t *cnt, const uint64_t *hash, uint32_t m, int64_t f0) {
  const int T = threads();
  auto work = [&](int t) {
    ReadGroups &G = part[t];
    for (uint32_t i = 0; i < m; ++i) {
      if (!cnt[i]) continue;
      const uint64_t h = hash[2 * (size_t)i] ^ (hash[2 * (size_t)i + 1] * 0x9e3779b97f4a7c15ull);
      if ((int)((h >> 17) % (uint64_t)T) != t) continue;
      G.add(ent + off[i], cnt[i], 1, &h, f0 + i);
    }
  };
  if (T == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (size_t k = 0; k < th.size(); ++k) th[k].join();
}

// Relocatable image of a group table: [nGroups, nEntries, assignedFragments, 0] then ptr[nGroups+1], hashes[nGroups],
// first[nGroups], entries.
struct GroupBlobView {
  uint64_t nG = 0, nE = 0, assigned = 0;
  const uint8_t *ptr = nullptr, *hashes = nullptr, *first = nullptr, *ent = nullptr;
  bool parse(const uint8_t *blob, uint64_t bytes) {
    uint64_t hdr[4];
    if (bytes < sizeof(hdr)) return false;
    memcpy(hdr, blob, sizeof(hdr));
    nG = hdr[0]; nE = hdr[1]; assigned = hdr[2];
    if (bytes < sizeof(hdr) + (nG + 1) * 8 + nG * 16 + nE * sizeof(HostEntry)) return false;
    ptr = blob + sizeof(hdr); hashes = ptr + (nG + 1) * 8; first = hashes + nG * 8; ent = first + nG * 8;
    return true;
  }
  bool row(uint64_t g, int64_t &b, int64_t &e) const {
    memcpy(&b, ptr + g * 8, 8); memcpy(&e, ptr + (g + 1) * 8, 8);
    return b >= 0 && e >= b && (uint64_t)e <= nE;
  }
  uint64_t hash(uint64_t g) const { uint64_t h; memcpy(&h, hashes + g * 8, 8); return h; }
  int64_t first_frag(uint64_t g) const { int64_t f; memcpy(&f, first + g * 8, 8); return f; }
  const HostEntry *entries(int64_t b) const { return (const HostEntry *)(ent + (size_t)b * sizeof(HostEntry)); }   // 8-byte aligned
};

inline size_t serialized_group_bytes(const ReadGroups &G) {
  return 4 * 8 + G.ptr.size() * 8 + (size_t)G.size() * 16 + G.ent.size() * sizeof(HostEntry);
}
inline void serialize_groups(const ReadGroups &G, uint8_t *p) {
  const uint64_t hdr[4] = {(uint64_t)G.size(), (uint64_t)G.ent.size(), (uint64_t)G.assignedFragments, 0};
  memcpy(p, hdr, sizeof(hdr)); p += sizeof(hdr);
  memcpy(p, G.ptr.data(), G.ptr.size() * 8); p += G.ptr.size() * 8;
  if (G.size()) { memcpy(p, G.hashes.data(), (size_t)G.size() * 8); p += (size_t)G.size() * 8; memcpy(p, G.first.data(), (size_t)G.size() * 8); p += (size_t)G.size() * 8; }
  if (!G.ent.empty()) {
    const size_t bytes = G.ent.size() * sizeof(HostEntry);
    const int T = bytes < ((size_t)8 << 20) ? 1 : 8;
    const uint8_t *src = (const uint8_t *)G.ent.data();
    auto copy = [&](int t) { const size_t a = bytes * t / T, b = bytes * (t + 1) / T; memcpy(p + a, src + a, b - a); };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(copy, t);
    copy(0);
    for (size_t k = 0; k < th.size(); ++k) th[k].join();
  }
}
inline void serialize_groups(const ReadGroups &G, std::vector<uint8_t> &blob) {
  blob.resize(serialized_group_bytes(G));
  serialize_groups(G, blob.data());
}

// The table split by the rank that merges each group: the hash space is cut into world x T partitions, rank r owns the
// partitions r*T .. r*T + T - 1 (thread t of rank r merges partition r*T + t, see merge_tables_partition).  One blob per owner,
// same format as serialize_groups (assigned = 0), written back to back into `out`; bytes[r] = size of the blob for rank r.
struct PartitionPlan {
  std::vector<std::vector<int32_t> > groupsOf;      // [world] group ids, ascending
  std::vector<size_t> bytes;                        // [world]
  size_t total = 0;
};
inline void plan_partitions(const ReadGroups &G, int world, int T, PartitionPlan &pl) {
  const uint64_t P = (uint64_t)world * (uint64_t)(T < 1 ? 1 : T);
  pl.groupsOf.assign((size_t)world, std::vector<int32_t>());
  std::vector<size_t> nE((size_t)world, 0);
  for (int32_t g = 0; g < G.size(); ++g) {
    const int owner = (int)(((G.hashes[g] >> 17) % P) / (uint64_t)(T < 1 ? 1 : T));
    pl.groupsOf[owner].push_back(g);
    nE[owner] += (size_t)(G.ptr[g + 1] - G.ptr[g]);
  }
  pl.bytes.assign((size_t)world, 0);
  pl.total = 0;
  for (int r = 0; r < world; ++r) {
    const size_t nG = pl.groupsOf[r].size();
    pl.bytes[r] = (4 * 8 + (nG + 1) * 8 + nG * 16 + nE[r] * sizeof(HostEntry) + 15) & ~(size_t)15;
    pl.total += pl.bytes[r];
  }
}
inline void serialize_partitions(const ReadGroups &G, const PartitionPlan &pl, uint8_t *out, int threads) {
  const int world = (int)pl.groupsOf.size();
  std::vector<size_t> at((size_t)world + 1, 0);
  for (int r = 0; r < world; ++r) at[r + 1] = at[r] + pl.bytes[r];
  run_threads(std::max(1, std::min(threads, world)), [&](int t) {
    const int nT = std::max(1, std::min(threads, world));
    for (int r = t; r < world; r += nT) {
      const std::vector<int32_t> &ids = pl.groupsOf[r];
      const size_t nG = ids.size();
      uint8_t *p = out + at[r];
      size_t nE = 0;
      for (size_t k = 0; k < nG; ++k) nE += (size_t)(G.ptr[ids[k] + 1] - G.ptr[ids[k]]);
      const uint64_t hdr[4] = {(uint64_t)nG, (uint64_t)nE, 0, 0};
      memcpy(p, hdr, sizeof(hdr)); p += sizeof(hdr);
      int64_t *ptr = (int64_t *)p; p += (nG + 1) * 8;
      uint64_t *hs = (uint64_t *)p; p += nG * 8;
      int64_t *fi = (int64_t *)p; p += nG * 8;
      HostEntry *en = (HostEntry *)p;
      int64_t run = 0;
      for (size_t k = 0; k < nG; ++k) {
        const int32_t g = ids[k];
        const int64_t n = G.ptr[g + 1] - G.ptr[g];
        ptr[k] = run; hs[k] = G.hashes[g]; fi[k] = G.first[g];
        if (n) memcpy(en + run, G.ent.data() + G.ptr[g], (size_t)n * sizeof(HostEntry));
        run += n;
      }
      ptr[nG] = run;
    }
  });
}

// The same split taken straight from the coalescing shards of this rank (no gathered table in between): the groups of a blob
// need no particular order, the owner files them by hash and orders them by their creating fragment.
struct ShardPlan {
  std::vector<std::vector<std::pair<int32_t, int32_t> > > groupsOf;      // [world] (shard, group)
  std::vector<size_t> bytes;
  size_t total = 0;
};
inline void plan_partitions(const GroupShards &S, int world, int T, ShardPlan &pl) {
  const uint64_t P = (uint64_t)world * (uint64_t)(T < 1 ? 1 : T);
  pl.groupsOf.assign((size_t)world, std::vector<std::pair<int32_t, int32_t> >());
  std::vector<size_t> nE((size_t)world, 0);
  for (int t = 0; t < S.threads(); ++t) {
    const ReadGroups &G = S.part[t];
    for (int32_t g = 0; g < G.size(); ++g) {
      const int owner = (int)(((G.hashes[g] >> 17) % P) / (uint64_t)(T < 1 ? 1 : T));
      pl.groupsOf[owner].push_back(std::make_pair((int32_t)t, g));
      nE[owner] += (size_t)(G.ptr[g + 1] - G.ptr[g]);
    }
  }
  pl.bytes.assign((size_t)world, 0);
  pl.total = 0;
  for (int r = 0; r < world; ++r) {
    const size_t nG = pl.groupsOf[r].size();
    pl.bytes[r] = (4 * 8 + (nG + 1) * 8 + nG * 16 + nE[r] * sizeof(HostEntry) + 15) & ~(size_t)15;
    pl.total += pl.bytes[r];
  }
}
inline void serialize_partitions(const GroupShards &S, const ShardPlan &pl, uint8_t *out, int threads) {
  const int world = (int)pl.groupsOf.size();
  std::vector<size_t> at((size_t)world + 1, 0);
  for (int r = 0; r < world; ++r) at[r + 1] = at[r] + pl.bytes[r];
  const int nT = std::max(1, std::min(threads, world));
  run_threads(nT, [&](int t) {
    for (int r = t; r < world; r += nT) {
      const std::vector<std::pair<int32_t, int32_t> > &ids = pl.groupsOf[r];
      const size_t nG = ids.size();
      uint8_t *p = out + at[r];
      size_t nE = 0;
      for (size_t k = 0; k < nG; ++k) { const ReadGroups &G = S.part[ids[k].first]; nE += (size_t)(G.ptr[ids[k].second + 1] - G.ptr[ids[k].second]); }
      const uint64_t hdr[4] = {(uint64_t)nG, (uint64_t)nE, 0, 0};
      memcpy(p, hdr, sizeof(hdr)); p += sizeof(hdr);
      int64_t *ptr = (int64_t *)p; p += (nG + 1) * 8;
      uint64_t *hs = (uint64_t *)p; p += nG * 8;
      int64_t *fi = (int64_t *)p; p += nG * 8;
      HostEntry *en = (HostEntry *)p;
      int64_t run = 0;
      for (size_t k = 0; k < nG; ++k) {
        const ReadGroups &G = S.part[ids[k].first];
        const int32_t g = ids[k].second;
        const int64_t n = G.ptr[g + 1] - G.ptr[g];
        ptr[k] = run; hs[k] = G.hashes[g]; fi[k] = G.first[g];
        if (n) memcpy(en + run, G.ent.data() + G.ptr[g], (size_t)n * sizeof(HostEntry));
        run += n;
      }
      ptr[nG] = run;
    }
  });
}

// Merge of another rank's table (read-sharded path): its groups are added in their own order, so merging the ranks'
// tables in rank order visits the allele sets in the order a single process would first see them shard by shard.
// float32 weights of equal allele sets add as (sum of rank 0) + (sum of rank 1) + ...
inline bool merge_groups(ReadGroups &G, const uint8_t *blob, uint64_t bytes) {
  GroupBlobView V;
  if (!V.parse(blob, bytes)) return false;
  std::vector<HostEntry> row;
  for (uint64_t g = 0; g < V.nG; ++g) {
    int64_t b, e;
    if (!V.row(g, b, e)) return false;
    row.assign(V.entries(b), V.entries(e));
    G.add(row.data(), (uint32_t)(e - b), 0);
  }
  G.assignedFragments += (int64_t)V.assigned;
  return true;
}

// The same merge on T threads: thread t files the groups of hash partition t from rank 0, 1, ... (rank order inside every
// partition), gather() restores the global first-appearance order from fragBase[r] + first.  All tables must carry
// hashes of one kind (the pairing kernel's).
inline bool merge_tables_parallel(const std::vector<GroupBlobView> &tables, const std::vector<int64_t> &fragBase, int T, ReadGroups &out) {
  GroupShards M(T);
  T = M.threads();
  std::vector<char> ok((size_t)T, 1);
  auto work = [&](int t) {
    ReadGroups &G = M.part[t];
    for (size_t r = 0; r < tables.size(); ++r) {
      const GroupBlobView &V = tables[r];
      for (uint64_t g = 0; g < V.nG; ++g) {
        const uint64_t h = V.hash(g);
        if ((int)((h >> 17) % (uint64_t)T) != t) continue;
        int64_t b, e;
        if (!V.row(g, b, e)) { ok[t] = 0; return; }
        G.add(V.entries(b), (uint32_t)(e - b), 0, &h, fragBase[r] + V.first_frag(g));
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (size_t k = 0; k < th.size(); ++k) th[k].join();
  for (int t = 0; t < T; ++t) if (!ok[t]) return false;
  M.gather(out);
  for (size_t r = 0; r < tables.size(); ++r) out.assignedFragments += (int64_t)tables[r].assigned;
  return true;
}

// contiguous row ranges balanced by non-zeros (rank r takes rows [bounds[r], bounds[r+1]))
inline void partition_rows(const int64_t *rowPtr, int32_t nGroups, int32_t world, int32_t *bounds) {
  const int64_t nnz = rowPtr[nGroups];
  bounds[0] = 0;
  int32_t g = 0;
  for (int32_t r = 1; r < world; ++r) {
    const int64_t target = nnz * r / world;
    while (g < nGroups && rowPtr[g] < target) ++g;
    bounds[r] = g;
  }
  bounds[world] = nGroups;
}

// Genotyper::FinalizeReadAssignments + BuildAlleleEquivalentClass (Genotyper.hpp:912-939, 1072-1139).
// RemoveLowMAPQAlleleInEquivalentClass (:1330-1368) keeps the members whose summed qual is maximal; every
// assignment has qual 1 and EC members share their read groups, so it keeps all of them.
// ---- small fork-join helpers for the serial tail of the host model (deterministic: every thread owns a contiguous
// range and results are concatenated in range order, so the output is the single-threaded one)
// inputs smaller than this stay single-threaded (T1K_PAR_MIN overrides: the parity tests force the threaded path)
inline size_t par_min_entries() {
  const char *e = getenv("T1K_PAR_MIN");
  return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)1 << 16;
}
template <class F> inline void run_threads(int n, F fn) {
  if (n <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < n; ++t) th.emplace_back(fn, t);
  fn(0);
  for (size_t t = 0; t < th.size(); ++t) th[t].join();
}
// contiguous row ranges with about equal numbers of entries: bounds[t] .. bounds[t + 1]
inline std::vector<int32_t> balanced_ranges(const int64_t *ptr, int32_t nRows, int threads) {
  std::vector<int32_t> b((size_t)threads + 1, nRows);
  b[0] = 0;
  const int64_t nnz = nRows > 0 ? ptr[nRows] - ptr[0] : 0;
  int32_t r = 0;
  for (int t = 1; t < threads; ++t) {
    const int64_t want = ptr[0] + nnz * t / threads;
    while (r < nRows && ptr[r] < want) ++r;
    b[t] = r;
  }
  return b;
}
// CSR (row -> column ids, colOf(k) for entry k) to CSC with ascending row order inside every column
template <class ColOf>
inline void transpose_csr(const int64_t *rowPtr, int32_t nRows, ColOf colOf, int32_t nCols, int threads, std::vector<int64_t> &colPtr,
                          std::vector<int32_t> &rowIdx) {
  const int64_t nnz = nRows > 0 ? rowPtr[nRows] - rowPtr[0] : 0, k0 = nRows > 0 ? rowPtr[0] : 0;
  if (threads < 1) threads = 1;
  if ((size_t)nnz < par_min_entries()) threads = 1;
  const std::vector<int32_t> b = balanced_ranges(rowPtr, nRows, threads);
  std::vector<std::vector<int64_t> > cnt((size_t)threads, std::vector<int64_t>((size_t)nCols, 0));
  run_threads(threads, [&](int t) {
    std::vector<int64_t> &c = cnt[t];
    for (int64_t k = rowPtr[b[t]]; k < rowPtr[b[t + 1]]; ++k) ++c[colOf(k)];
  });
  colPtr.assign((size_t)nCols + 1, 0);
  for (int32_t c = 0; c < nCols; ++c) {
    int64_t run = colPtr[c];
    for (int t = 0; t < threads; ++t) { const int64_t n = cnt[t][c]; cnt[t][c] = run; run += n; }     // cnt becomes the thread's cursor
    colPtr[c + 1] = run;
  }
  rowIdx.resize((size_t)std::max<int64_t>(nnz, 1));
  run_threads(threads, [&](int t) {
    std::vector<int64_t> &cur = cnt[t];
    for (int32_t r = b[t]; r < b[t + 1]; ++r)
      for (int64_t k = rowPtr[r]; k < rowPtr[r + 1]; ++k) rowIdx[cur[colOf(k)]++] = r;
  });
  (void)k0;
}

// ---- The same merge with the work divided over the ranks (the default of a read-sharded run): rank r merges only the hash
// partitions it owns (partition space = world x T; thread t of rank r takes partition r*T + t) from every rank's table,
// the ranks exchange their merged partitions, and assemble_partitions() interleaves them by the creating fragment.
// Every group is still merged by exactly one thread in rank order, so the result equals merge_tables_parallel's; the
// per-rank merge work is 1/world of it.
inline bool merge_tables_partition(const std::vector<GroupBlobView> &tables, const std::vector<int64_t> &fragBase, int rank, int world, int T,
                                   ReadGroups &mine) {
  GroupShards M(T);
  T = M.threads();
  const uint64_t P = (uint64_t)world * (uint64_t)T;
  std::vector<char> ok((size_t)T, 1);
  run_threads(T, [&](int t) {
    ReadGroups &G = M.part[t];
    const uint64_t own = (uint64_t)rank * (uint64_t)T + (uint64_t)t;
    for (size_t r = 0; r < tables.size(); ++r) {
      const GroupBlobView &V = tables[r];
      for (uint64_t g = 0; g < V.nG; ++g) {
        const uint64_t h = V.hash(g);
        if ((h >> 17) % P != own) continue;
        int64_t b, e;
        if (!V.row(g, b, e)) { ok[t] = 0; return; }
        G.add(V.entries(b), (uint32_t)(e - b), 0, &h, fragBase[r] + V.first_frag(g));
      }
    }
  });
  for (int t = 0; t < T; ++t) if (!ok[t]) return false;
  M.gather(mine);
  return true;
}
// parts[r] = rank r's merged partitions (`first` already global); out = all groups in first-appearance order
inline bool assemble_partitions(const std::vector<GroupBlobView> &parts, int T, ReadGroups &out) {
  struct Ref { int64_t first; uint32_t r; uint64_t g; };
  std::vector<Ref> order;
  for (size_t r = 0; r < parts.size(); ++r)
    for (uint64_t g = 0; g < parts[r].nG; ++g) order.push_back(Ref{parts[r].first_frag(g), (uint32_t)r, g});
  std::sort(order.begin(), order.end(), [](const Ref &a, const Ref &b) { return a.first < b.first; });
  const size_t nG = order.size();
  out.ptr.assign(nG + 1, 0); out.first.resize(nG); out.hashes.resize(nG);
  for (size_t k = 0; k < nG; ++k) {
    int64_t b, e;
    if (!parts[order[k].r].row(order[k].g, b, e)) return false;
    out.ptr[k + 1] = out.ptr[k] + (e - b);
    out.first[k] = order[k].first;
    out.hashes[k] = parts[order[k].r].hash(order[k].g);
  }
  out.ent.resize((size_t)out.ptr[nG]);
  if (T < 1 || nG < 1024) T = 1;
  run_threads(T, [&](int t) {
    for (size_t k = nG * t / T; k < nG * (t + 1) / T; ++k) {
      int64_t b, e;
      parts[order[k].r].row(order[k].g, b, e);
      if (e > b) memcpy(out.ent.data() + out.ptr[k], parts[order[k].r].entries(b), (size_t)(e - b) * sizeof(HostEntry));
    }
  });
  return true;
}

// ---- compact form of a merged partition for the global tail of a read-sharded run: [nG, nE], ptr[nG+1], first[nG], count[nG]
// (double), allele[nE] (int32)
inline size_t compact_group_bytes(const ReadGroups &G) { return (2 * 8 + (size_t)(G.size() + 1) * 8 + (size_t)G.size() * 16 + G.ent.size() * 4 + 15) & ~(size_t)15; }
inline void serialize_compact(const ReadGroups &G, uint8_t *p, int threads) {
  const uint64_t hdr[2] = {(uint64_t)G.size(), (uint64_t)G.ent.size()};
  memcpy(p, hdr, sizeof(hdr)); p += sizeof(hdr);
  memcpy(p, G.ptr.data(), G.ptr.size() * 8); p += G.ptr.size() * 8;
  const int32_t nG = G.size();
  int64_t *first = (int64_t *)p; p += (size_t)nG * 8;
  double *cnt = (double *)p; p += (size_t)nG * 8;
  int32_t *al = (int32_t *)p;
  const GroupsView V = view_of(G);
  const int T = G.ent.size() < par_min_entries() ? 1 : std::max(1, threads);
  const std::vector<int32_t> b = balanced_ranges(G.ptr.data(), nG, T);
  run_threads(T, [&](int t) {
    for (int32_t g = b[t]; g < b[t + 1]; ++g) {
      first[g] = G.first[g]; cnt[g] = V.count_of(g);
      for (int64_t k = G.ptr[g]; k < G.ptr[g + 1]; ++k) al[k] = G.ent[k].alleleIdx;
    }
  });
}
struct CompactGroups {
  std::vector<int64_t> ptr{0};
  std::vector<int32_t> allele;
  std::vector<double> count;
  GroupsView view() const { GroupsView v; v.n = (int32_t)ptr.size() - 1; v.ptr = ptr.data(); v.allele = allele.data(); v.count = count.data(); return v; }
};
// blobs[r] / bytes[r]: rank r's compact partition; out = all groups in first-appearance order
inline bool assemble_compact(const std::vector<const uint8_t *> &blobs, const std::vector<uint64_t> &bytes, int T, CompactGroups &out) {
  struct Ref { int64_t first; uint32_t r; uint32_t g; };
  struct View { uint64_t nG, nE; const int64_t *ptr, *first; const double *cnt; const int32_t *al; };
  std::vector<View> V(blobs.size());
  std::vector<Ref> order;
  for (size_t r = 0; r < blobs.size(); ++r) {
    if (bytes[r] < 16) return false;
    uint64_t hdr[2]; memcpy(hdr, blobs[r], 16);
    View &v = V[r];
    v.nG = hdr[0]; v.nE = hdr[1];
    if (bytes[r] < 16 + (v.nG + 1) * 8 + v.nG * 16 + v.nE * 4) return false;
    v.ptr = (const int64_t *)(blobs[r] + 16); v.first = v.ptr + v.nG + 1; v.cnt = (const double *)(v.first + v.nG); v.al = (const int32_t *)(v.cnt + v.nG);
    for (uint64_t g = 0; g < v.nG; ++g) {
      if (v.ptr[g] < 0 || v.ptr[g + 1] < v.ptr[g] || (uint64_t)v.ptr[g + 1] > v.nE) return false;
      order.push_back(Ref{v.first[g], (uint32_t)r, (uint32_t)g});
    }
  }
  std::sort(order.begin(), order.end(), [](const Ref &a, const Ref &b) { return a.first < b.first; });
  const size_t nG = order.size();
  out.ptr.assign(nG + 1, 0); out.count.resize(nG);
  for (size_t k = 0; k < nG; ++k) {
    const View &v = V[order[k].r];
    out.ptr[k + 1] = out.ptr[k] + (v.ptr[order[k].g + 1] - v.ptr[order[k].g]);
    out.count[k] = v.cnt[order[k].g];
  }
  out.allele.resize((size_t)out.ptr[nG]);
  if (T < 1 || nG < 1024) T = 1;
  run_threads(T, [&](int t) {
    for (size_t k = nG * t / T; k < nG * (t + 1) / T; ++k) {
      const View &v = V[order[k].r];
      const int64_t b = v.ptr[order[k].g], e = v.ptr[order[k].g + 1];
      if (e > b) memcpy(out.allele.data() + out.ptr[k], v.al + b, (size_t)(e - b) * 4);
    }
  });
  return true;
}

// The same for partitions that stay on the device: heads[r] = the start of rank r's compact blob up to (not including) its
// allele ids ([nG, nE], ptr, first, count — all the host needs), blobBase[r] = the blob's byte offset in the device buffer.
// out.ptr / out.count as assemble_compact gives them, out.allele stays empty; srcOff[k] = byte offset of group k's allele run.
inline bool assemble_compact_heads(const std::vector<const uint8_t *> &heads, const std::vector<uint64_t> &headBytes, const std::vector<uint64_t> &blobBytes,
                                   const std::vector<uint64_t> &blobBase, CompactGroups &out, std::vector<int64_t> &srcOff) {
  struct Ref { int64_t first; uint32_t r; uint32_t g; };
  struct View { uint64_t nG, nE; const int64_t *ptr, *first; const double *cnt; uint64_t alOff; };
  std::vector<View> V(heads.size());
  std::vector<Ref> order;
  for (size_t r = 0; r < heads.size(); ++r) {
    if (headBytes[r] < 16) return false;
    uint64_t hdr[2]; memcpy(hdr, heads[r], 16);
    View &v = V[r];
    v.nG = hdr[0]; v.nE = hdr[1];
    v.alOff = 16 + (v.nG + 1) * 8 + v.nG * 16;
    if (headBytes[r] < v.alOff || blobBytes[r] < v.alOff + v.nE * 4) return false;
    v.ptr = (const int64_t *)(heads[r] + 16); v.first = v.ptr + v.nG + 1; v.cnt = (const double *)(v.first + v.nG);
    for (uint64_t g = 0; g < v.nG; ++g) {
      if (v.ptr[g] < 0 || v.ptr[g + 1] < v.ptr[g] || (uint64_t)v.ptr[g + 1] > v.nE) return false;
      order.push_back(Ref{v.first[g], (uint32_t)r, (uint32_t)g});
    }
  }
  std::sort(order.begin(), order.end(), [](const Ref &a, const Ref &b) { return a.first < b.first; });
  const size_t nG = order.size();
  out.ptr.assign(nG + 1, 0); out.count.resize(nG); out.allele.clear();
  srcOff.resize(nG);
  for (size_t k = 0; k < nG; ++k) {
    const View &v = V[order[k].r];
    out.ptr[k + 1] = out.ptr[k] + (v.ptr[order[k].g + 1] - v.ptr[order[k].g]);
    out.count[k] = v.cnt[order[k].g];
    srcOff[k] = (int64_t)(blobBase[order[k].r] + v.alOff + (uint64_t)v.ptr[order[k].g] * 4);
  }
  return true;
}

// ---- unique read-ends of a chunk of fragments (the de-duplication of Genotyper.cpp:450-454: only the grouping matters)
// in two fork-join phases: (1) length, N flag and hash of every read-end (reads split across the threads), (2) one
// open-addressing table per hash partition (a thread owns a partition).  The unique index of a read-end = partition
// offset + rank inside the partition: deterministic for a given thread count, and nothing downstream depends on the
// order of the unique read-ends.  reads1/reads2: fixed-stride, NUL-padded rows (reads2 may be NULL: single-end).
struct ReadEndChunk {
  uint32_t f0 = 0, m = 0;
  std::vector<uint32_t> e1, e2;                 // per fragment: unique read-end of each mate
  std::vector<int32_t> w;                       // per unique read-end: number of duplicates (Genotyper.cpp:149,472)
  std::vector<uint64_t> off; std::vector<uint32_t> len; std::vector<char> bases;     // the unique read-ends, concatenated
  std::vector<uint8_t> hasN;                    // per fragment: some mate holds an N
  std::vector<const char *> rep;
  bool tooLong = false;
  double ms = 0;
};
inline void unique_read_ends(const char *reads1, const char *reads2, uint32_t stride, uint32_t f0, uint32_t m, uint32_t maxLen, int threads,
                             ReadEndChunk &C) {
  typedef uint64_t u64; typedef uint32_t u32; typedef uint8_t u8;
  const int mates = reads2 ? 2 : 1;
  C.f0 = f0; C.m = m; C.tooLong = false;
  C.e1.resize(m); if (reads2) C.e2.resize(m); else C.e2.clear();
  C.hasN.assign(m, 0);
  const size_t nEnds = (size_t)m * mates;
  const int T = nEnds < 4096 ? 1 : std::max(1, threads);
  std::vector<u64> hashes(nEnds);
  std::vector<u32> lens(nEnds);
  std::vector<u8> tooLong((size_t)T, 0);
  auto end_ptr = [&](size_t k) { return ((k % mates) ? reads2 : reads1) + (size_t)(f0 + k / mates) * stride; };
  run_threads(T, [&](int tIdx) {
    // whole fragments per thread, so that a fragment's N flag has one writer
    const size_t k0 = (size_t)m * tIdx / T * mates, k1 = (size_t)m * (tIdx + 1) / T * mates;
    for (size_t k = k0; k < k1; ++k) {
      const char *s = end_ptr(k);
      u32 L = 0; u64 h = 1469598103934665603ull; bool hasN = false;
      while (L < stride && s[L]) { h = (h ^ (u8)s[L]) * 1099511628211ull; hasN |= s[L] == 'N'; ++L; }
      if (L > maxLen) { tooLong[tIdx] = 1; L = maxLen; }
      if (hasN) C.hasN[k / mates] = 1;
      hashes[k] = h ^ (h >> 29); lens[k] = L;
    }
  });
  for (int i = 0; i < T; ++i) if (tooLong[i]) C.tooLong = true;
  struct Part { std::vector<u32> table, first, cnt; };      // first: read-end index of each unique, cnt: duplicates
  std::vector<Part> parts((size_t)T);
  std::vector<u32> local(nEnds);                             // rank of the read-end's unique inside its partition
  run_threads(T, [&](int tIdx) {
    Part &P = parts[tIdx];
    size_t mine = 0;
    for (size_t k = 0; k < nEnds; ++k) mine += (int)((hashes[k] >> 40) % (u64)T) == tIdx;
    size_t tabSize = 16; while (tabSize < mine * 2) tabSize <<= 1;
    P.table.assign(tabSize, 0xffffffffu);
    for (size_t k = 0; k < nEnds; ++k) {
      if ((int)((hashes[k] >> 40) % (u64)T) != tIdx) continue;
      const char *s = end_ptr(k);
      const u32 L = lens[k];
      size_t slot = (size_t)hashes[k] & (tabSize - 1);
      u32 u;
      for (;;) {
        u = P.table[slot];
        if (u == 0xffffffffu) { u = (u32)P.first.size(); P.table[slot] = u; P.first.push_back((u32)k); P.cnt.push_back(0); break; }
        if (lens[P.first[u]] == L && memcmp(end_ptr(P.first[u]), s, L) == 0) break;
        slot = (slot + 1) & (tabSize - 1);
      }
      ++P.cnt[u];
      local[k] = u;
    }
  });
  std::vector<u32> base((size_t)T + 1, 0);
  for (int i = 0; i < T; ++i) base[i + 1] = base[i] + (u32)parts[i].first.size();
  const size_t nU = base[T];
  C.rep.resize(nU); C.len.resize(nU); C.w.resize(nU); C.off.resize(nU);
  for (int i = 0; i < T; ++i)
    for (size_t u = 0; u < parts[i].first.size(); ++u) {
      const size_t k = parts[i].first[u];
      C.rep[base[i] + u] = end_ptr(k); C.len[base[i] + u] = lens[k]; C.w[base[i] + u] = (int32_t)parts[i].cnt[u];
    }
  size_t tot = 0;
  for (size_t k = 0; k < nU; ++k) { C.off[k] = tot; tot += C.len[k]; }
  C.bases.resize(tot + 1);
  run_threads(T, [&](int tIdx) {
    for (size_t k = nEnds * tIdx / T; k < nEnds * (tIdx + 1) / T; ++k) {
      const u32 u = base[(int)((hashes[k] >> 40) % (u64)T)] + local[k];
      ((k % mates) ? C.e2 : C.e1)[k / mates] = u;
    }
    for (size_t k = nU * tIdx / T; k < nU * (tIdx + 1) / T; ++k) memcpy(C.bases.data() + C.off[k], C.rep[k], C.len[k]);
  });
}

struct EquivalenceClasses {
  std::vector<int32_t> ecPtr{0}, ecAlleles, alleleEc;
  std::vector<int64_t> inPtr;          // readsInAllele: the groups of every allele, ascending (kept: the members of a class sit in exactly
  std::vector<int32_t> in;             // the same groups, so column e of the EM's matrix IS the group list of the class's first member)

  void build(const ReadGroups &G, int32_t nAlleles, int threads = 1) { build(view_of(G), nAlleles, threads); }
  void build(const GroupsView &G, int32_t nAlleles, int threads = 1) {
    const int32_t readCnt = G.n;
    // readsInAllele (Genotyper.hpp:912-939): the groups of every allele, ascending
    transpose_csr(G.ptr, readCnt, [&G](int64_t k) { return G.allele_at(k); }, nAlleles, threads, inPtr, in);
    struct FP { int32_t a, b; };
    std::vector<FP> fp(nAlleles);
    {
      const int T = (size_t)G.entries() < par_min_entries() ? 1 : std::max(1, threads);
      const std::vector<int32_t> ab = balanced_ranges(inPtr.data(), nAlleles, T);
      run_threads(T, [&](int t) {
        for (int32_t a = ab[t]; a < ab[t + 1]; ++a) {
          int32_t b = -1;
          if (inPtr[a + 1] > inPtr[a]) {
            b = 0;
            for (int64_t k = inPtr[a]; k < inPtr[a + 1]; ++k)
              b = (int32_t)(((uint32_t)b * (uint32_t)readCnt + (uint32_t)in[k]) % 1000003u);   // Genotyper.hpp:1089
          }
          fp[a].a = a; fp[a].b = b;
        }
      });
    }
    std::sort(fp.begin(), fp.end(), [](const FP &x, const FP &y) { return x.b != y.b ? y.b < x.b : x.a < y.a; });
    alleleEc.assign(nAlleles, -1);
    std::vector<std::vector<int32_t> > ecs;
    for (int32_t i = 0; i < nAlleles; ++i) {
      if (fp[i].b == -1) break;
      const int32_t a = fp[i].a;
      int32_t found = -1;
      for (int32_t j = i - 1; j >= 0 && fp[j].b == fp[i].b; --j) {
        const int32_t o = fp[j].a;
        const int64_t n = inPtr[a + 1] - inPtr[a];
        if (inPtr[o + 1] - inPtr[o] == n && std::equal(in.begin() + inPtr[a], in.begin() + inPtr[a + 1], in.begin() + inPtr[o])) {
          found = o;
          break;
        }
      }
      if (found < 0) { alleleEc[a] = (int32_t)ecs.size(); ecs.push_back(std::vector<int32_t>(1, a)); }
      else { alleleEc[a] = alleleEc[found]; ecs[alleleEc[found]].push_back(a); }
    }
    ecPtr.assign(1, 0); ecAlleles.clear();
    for (size_t e = 0; e < ecs.size(); ++e) {
      ecAlleles.insert(ecAlleles.end(), ecs[e].begin(), ecs[e].end());
      ecPtr.push_back((int32_t)ecAlleles.size());
    }
  }
  int32_t size() const { return (int32_t)ecPtr.size() - 1; }

};

// EM inputs as QuantifyAlleleEquivalentClass assembles them (Genotyper.hpp:1155-1232)
struct EmInputs {
  std::vector<int64_t> rowPtr;
  std::vector<int32_t> col, ecLen;
  std::vector<double> count, x0;

  void build(const ReadGroups &G, const EquivalenceClasses &EC, const int32_t *effectiveLen, const int32_t *seqWeight, int threads = 1) {
    build(view_of(G), EC, effectiveLen, seqWeight, threads);
  }
  void build(const GroupsView &G, const EquivalenceClasses &EC, const int32_t *effectiveLen, const int32_t *seqWeight, int threads = 1) {
    const int32_t n = G.n, E = EC.size();
    count.resize(n);
    if (threads < 1 || (size_t)G.entries() < par_min_entries()) threads = 1;
    const std::vector<int32_t> b = balanced_ranges(G.ptr, n, threads);
    std::vector<std::vector<int32_t> > cols((size_t)threads);
    std::vector<int64_t> rowLen((size_t)n, 0);
    run_threads(threads, [&](int t) {
      std::vector<int32_t> stamp(E, -1);
      std::vector<int32_t> &out = cols[t];
      for (int32_t g = b[t]; g < b[t + 1]; ++g) {
        count[g] = G.count_of(g);
        const size_t before = out.size();
        for (int64_t k = G.ptr[g]; k < G.ptr[g + 1]; ++k) {
          const int32_t e = EC.alleleEc[G.allele_at(k)];
          if (stamp[e] != g) { stamp[e] = g; out.push_back(e); }
        }
        rowLen[g] = (int64_t)(out.size() - before);
      }
    });
    rowPtr.assign((size_t)n + 1, 0);
    for (int32_t g = 0; g < n; ++g) rowPtr[g + 1] = rowPtr[g] + rowLen[g];
    col.resize((size_t)rowPtr[n]);
    run_threads(threads, [&](int t) {
      if (!cols[t].empty()) memcpy(col.data() + rowPtr[b[t]], cols[t].data(), cols[t].size() * sizeof(int32_t));
    });
    build_ec_vectors(EC, effectiveLen, seqWeight);
  }
  // ecInfo[].length and the initial abundances (Genotyper.hpp:1191-1232)
  void build_ec_vectors(const EquivalenceClasses &EC, const int32_t *effectiveLen, const int32_t *seqWeight) {
    const int32_t E = EC.size();
    ecLen.resize(E); x0.resize(E);
    for (int32_t e = 0; e < E; ++e) {
      int32_t len = effectiveLen[EC.ecAlleles[EC.ecPtr[e]]];
      double w = 0;
      for (int32_t k = EC.ecPtr[e]; k < EC.ecPtr[e + 1]; ++k) {
        len = std::min(len, effectiveLen[EC.ecAlleles[k]]);
        w += seqWeight ? seqWeight[EC.ecAlleles[k]] : 1;
      }
      ecLen[e] = len; x0[e] = w;
    }
  }
  // everything but the matrix (which the device tail builds where the EM reads it): group counts and the class vectors
  void build_vectors(const GroupsView &G, const EquivalenceClasses &EC, const int32_t *effectiveLen, const int32_t *seqWeight, int threads = 1) {
    const int32_t n = G.n;
    count.resize(n);
    if (threads < 1 || (size_t)G.entries() < par_min_entries()) threads = 1;
    const std::vector<int32_t> b = balanced_ranges(G.ptr, n, threads);
    run_threads(threads, [&](int t) { for (int32_t g = b[t]; g < b[t + 1]; ++g) count[g] = G.count_of(g); });
    build_ec_vectors(EC, effectiveLen, seqWeight);
  }
};

// Genotyper::SetAlleleAbundance, per-allele part (Genotyper.hpp:957-987)
inline void set_allele_abundance(const double *rc, const int32_t *ecLen, const int32_t *ecPtr, const int32_t *ecAlleles, int32_t nEc,
                                 int32_t nAlleles, double *abundance, double *ecAbundance) {
  for (int32_t i = 0; i < nAlleles; ++i) abundance[i] = ecAbundance[i] = 0;
  for (int32_t e = 0; e < nEc; ++e) {
    const int32_t size = ecPtr[e + 1] - ecPtr[e];
    const double abund = rc[e] / ecLen[e] * 1000.0;
    for (int32_t k = ecPtr[e]; k < ecPtr[e + 1]; ++k) { abundance[ecAlleles[k]] = abund / size; ecAbundance[ecAlleles[k]] = abund; }
  }
}

// ---- Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460; Genotyper.cpp:647), SURVEY.md §8f N4.
// Step 1 (:1395-1416): the covered range of every allele = min start / max end over its entries in the coalesced read groups.
// The reference walks the groups of the class representative and picks the entries of the class members; members of one
// class sit in exactly the same groups, so this is the per-allele min / max over all groups.  Threads own contiguous group
// ranges; min / max commute, so the result does not depend on the thread count.  spans[a] = minStart (INT32_MAX: no entry),
// spans[nAlleles + a] = maxEnd (-1: no entry).
inline void allele_spans(const ReadGroups &G, int32_t nAlleles, int threads, std::vector<int32_t> &spans) {
  const int32_t n = G.size();
  if (threads < 1 || G.ent.size() < par_min_entries()) threads = 1;
  const std::vector<int32_t> b = balanced_ranges(G.ptr.data(), n, threads);
  std::vector<std::vector<int32_t> > part((size_t)threads);
  run_threads(threads, [&](int t) {
    std::vector<int32_t> &sp = part[t];
    sp.assign((size_t)2 * nAlleles, -1);
    std::fill(sp.begin(), sp.begin() + nAlleles, INT32_MAX);
    for (int64_t k = G.ptr[b[t]]; k < G.ptr[b[t + 1]]; ++k) {
      const HostEntry &e = G.ent[k];
      if (e.start < sp[e.alleleIdx]) sp[e.alleleIdx] = e.start;
      if (e.end > sp[nAlleles + e.alleleIdx]) sp[nAlleles + e.alleleIdx] = e.end;
    }
  });
  spans.swap(part[0]);
  for (int t = 1; t < threads; ++t)
    for (int32_t a = 0; a < nAlleles; ++a) {
      spans[a] = std::min(spans[a], part[t][a]);
      spans[nAlleles + a] = std::max(spans[nAlleles + a], part[t][nAlleles + a]);
    }
}
// Step 2 (:1418-1457): likelihood pow(effectiveLen / len, ecAbundance) per member, keep the members within 0.05 of the best
// (or equal to it).  kept[a] = 1 when allele a stays in its class (0 also for alleles without a class).
inline void ec_likelihood_filter(const int32_t *ecPtr, const int32_t *ecAlleles, int32_t nEc, int32_t nAlleles, const int32_t *alleleLen,
                                 const double *ecAbundance, const int32_t *spans, uint8_t *kept) {
  memset(kept, 0, (size_t)nAlleles);
  std::vector<double> ll;
  for (int32_t e = 0; e < nEc; ++e) {
    const int32_t size = ecPtr[e + 1] - ecPtr[e];
    ll.assign((size_t)size, 0.0);
    double maxLikelihood = -1;
    for (int32_t j = 0; j < size; ++j) {
      const int32_t a = ecAlleles[ecPtr[e] + j];
      const int32_t len = alleleLen[a];
      const int32_t minStart = std::min(len, spans[a]), maxEnd = std::max(-1, spans[nAlleles + a]);
      int32_t effectiveLen = maxEnd - minStart + 1;
      if (effectiveLen > len) effectiveLen = len;
      const double v = pow(double(effectiveLen) / len, ecAbundance[a]);
      if (v > maxLikelihood) maxLikelihood = v;
      ll[j] = v;
    }
    const double cutoff = 0.05;
    for (int32_t j = 0; j < size; ++j)
      if (ll[j] / maxLikelihood >= cutoff || ll[j] == maxLikelihood) kept[ecAlleles[ecPtr[e] + j]] = 1;
  }
}

// The every-10-iterations low-abundance mask (Genotyper.hpp:1292-1313 with SetAlleleAbundance :989-1013):
// returns the new ecAbundance0 in x0.
inline void em_mask(const double *rc, const int32_t *ecLen, const int32_t *ecPtr, const int32_t *ecAlleles, int32_t nEc,
                    int32_t nAlleles, const int32_t *alleleMajor, const int32_t *alleleGene, int32_t nMajor, int32_t nGene,
                    double filterFrac, double *x0) {
  std::vector<double> ab(nAlleles), ecAb(nAlleles), major(nMajor, 0.0), gmax(nGene, 0.0);
  set_allele_abundance(rc, ecLen, ecPtr, ecAlleles, nEc, nAlleles, ab.data(), ecAb.data());
  for (int32_t i = 0; i < nAlleles; ++i) major[alleleMajor[i]] += ab[i];
  for (int32_t i = 0; i < nAlleles; ++i) gmax[alleleGene[i]] = std::max(gmax[alleleGene[i]], major[alleleMajor[i]]);
  for (int32_t i = 0; i < nAlleles; ++i)
    if (major[alleleMajor[i]] < filterFrac * 0.5 * gmax[alleleGene[i]]) ecAb[i] = 0;
  for (int32_t e = 0; e < nEc; ++e) x0[e] = ecAb[ecAlleles[ecPtr[e]]];
}

}  // namespace t1k
